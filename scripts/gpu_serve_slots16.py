"""ContinuousBatcher with 16 slots (two n8 batch tiles of the weight-streaming GEMM): 24 teacher-forced TINY requests must
each reproduce their own stand-alone greedy_decode (the 2-slot version of this check is tests/test_decode_gpu.py)."""
import sys
import warnings

import torch

sys.path.insert(0, ".")
warnings.filterwarnings("ignore")
from metamorph_b200.engine.serve import ContinuousBatcher  # noqa: E402
from oracle.weights import TINY, make_weights  # noqa: E402
from tests.helpers import build_product_model  # noqa: E402

model = build_product_model(TINY, make_weights(TINY), num_image_tokens=4)
model.eval()
g = torch.Generator().manual_seed(23)
reqs = []
for i in range(24):
    P, n_new = int(torch.randint(3, 13, (1,), generator=g)), int(torch.randint(5, 21, (1,), generator=g))
    prompt = torch.randint(0, 128000, (1, P), generator=g)
    forced = torch.randint(0, 128000, (n_new + 2,), generator=g).to(torch.int32)
    if i % 5 == 0:
        forced[1] = 128256
    if i % 7 == 3:
        forced[3] = 128009
    reqs.append((model.get_model().embed_tokens(prompt.cuda()), n_new, forced))
srv = ContinuousBatcher(model, max_slots=16, max_context=64, max_new_tokens=24, poll_every=3)
rids = [srv.submit(e, max_new_tokens=n, forced_tokens=f) for e, n, f in reqs]
results = {rid: payload for rid, kind, payload in srv.run() if kind == "done"}
bad = 0
for rid, (emb, n_new, forced) in zip(rids, reqs):
    ids1, img1 = model.greedy_decode(None, None, emb, max_new_tokens=n_new, output_image=True, forced_tokens=forced.reshape(1, -1))
    ids, img = results[rid]
    n1 = img1.shape[0] if img1.dim() == 2 else 0
    ok = ids.cpu().tolist() == ids1[0].cpu().tolist() and img.shape[0] == n1
    if ok and n1:
        ok = (img.float() - img1.float()).abs().max().item() <= 3e-2 * (img1.float().abs().max().item() + 1e-6)
    bad += 0 if ok else 1
print(f"16 slots, {len(rids)} requests, {srv.steps_run} steps: {len(rids) - bad} identical to their stand-alone decode, {bad} different")
