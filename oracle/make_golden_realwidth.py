"""ORACLE (test infrastructure): golden DIGESTS of the reference itself at LLaMA-3-8B layer width.

Run in the build container only (needs /root/reference, read-only; ~45 GB of host RAM, a few minutes of CPU):
    PYTHONPATH=/root/repo python oracle/make_golden_realwidth.py
The tiny fixtures (make_golden.py) pin the arithmetic; this one pins the KERNEL COMBINATION the benchmark runs:
H=4096, 32 query / 8 kv heads (GQA group 4), d=128, I=14336, V=128258, T=4096 (32 key tiles) — case A, one decoder
layer, B=2 with one full-length and one ragged sample — and a two-layer case B whose batch length 1501 is not a
multiple of 4. The reference model (metamorph_llama.py:603-660 -> :285-498) is run in fp32 (forward + backward) and in
bf16 (forward; the error budget of tests/test_model_gpu.py). Only digests are stored: the three losses, logits / hidden
states at 64 sampled positions, and per-parameter gradient norms + 64 sampled entries, so the fixture stays small.
Nothing here is imported by the product or by the GPU box.
"""
import gc
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle.ref_model import activate, build_reference  # noqa: E402
from oracle.weights import REAL_A, REAL_B, make_batch_real, make_weights  # noqa: E402

activate("/root/reference")


def sample_positions(mask_after, g, n_per_sample=32):
    """(b, t) pairs inside the valid region of every sample, always including its first and last position."""
    out = []
    for b in range(mask_after.shape[0]):
        n = int(mask_after[b].sum())
        t = torch.randint(0, n, (n_per_sample - 2,), generator=g).tolist() + [0, n - 1]
        out += [(b, x) for x in t]
    return torch.tensor(out, dtype=torch.long)


def run_case(name, cfg, case):
    t0 = time.time()
    W = make_weights(cfg)
    ids, mask, labs, images = make_batch_real(case, cfg)
    g = torch.Generator().manual_seed(321)
    cols = torch.cat([torch.randint(0, cfg["vocab"], (48,), generator=g),
                      torch.tensor([128256, 128257, 128001, 128009])])
    ref = build_reference(cfg, W, torch.float32)
    with torch.no_grad():
        (_, pos, am, _, embeds, new_labels, impos, tgt) = ref.prepare_inputs_labels_for_multimodal(
            ids, None, mask, None, labs, images)
    T = embeds.shape[1]
    pts = sample_positions(am.bool(), g)
    res = dict(cfg=cfg, case=case, seq_len=T, new_labels=new_labels.to(torch.int32), image_positions=impos.to(torch.int8),
               new_attention_mask=am.bool(), targets_digest=tgt.float()[:, ::9, ::37].clone(), points=pts, logit_cols=cols)
    del embeds, tgt
    ref.zero_grad()
    out = ref(input_ids=ids, attention_mask=mask, labels=labs, images=images)
    res.update(loss=out.loss.detach().float().clone(), loss_language=torch.tensor(ref.loss_language),
               loss_image_ar=torch.tensor(ref.loss_image_ar))
    lg = out.logits.detach()
    res["logits_sub"] = lg[pts[:, 0], pts[:, 1]][:, cols].float().clone()
    res["logits_argmax"] = lg[pts[:, 0], pts[:, 1]].argmax(-1).clone()
    res["hidden_sub"] = out.hidden_states.detach()[pts[:, 0], pts[:, 1], :64].float().clone()
    del lg
    out.loss.backward()
    del out
    digest = {}
    for pname, p in ref.named_parameters():
        if p.grad is None or "vision_tower" in pname:
            continue
        v = p.grad.detach().float()
        flat = v.reshape(-1)
        # half of the samples from the largest-magnitude entries (a systematic error in the signal shows there),
        # half uniformly at random
        k = min(32, flat.numel())
        top = flat.abs().topk(k).indices
        rnd = torch.randint(0, flat.numel(), (32,), generator=g)
        idx = torch.cat([top, rnd])
        digest[pname] = dict(norm=v.norm().clone(), idx=idx, vals=flat[idx].clone(), shape=tuple(v.shape))
    res["grad_digest"] = digest
    print(f"[{name}] fp32 done in {time.time() - t0:.0f}s: loss {float(res['loss']):.6f} "
          f"lang {float(res['loss_language']):.6f} img {float(res['loss_image_ar']):.6f}", flush=True)
    del ref
    gc.collect()
    ref16 = build_reference(cfg, W, torch.bfloat16)
    with torch.no_grad():
        out = ref16(input_ids=ids, attention_mask=mask, labels=labs, images=images.bfloat16())
    lg = out.logits
    res["bf16"] = dict(loss=out.loss.float().clone(), loss_language=torch.tensor(ref16.loss_language),
                       loss_image_ar=torch.tensor(ref16.loss_image_ar),
                       logits_sub=lg[pts[:, 0], pts[:, 1]][:, cols].float().clone(),
                       hidden_sub=out.hidden_states[pts[:, 0], pts[:, 1], :64].float().clone())
    print(f"[{name}] bf16 loss {float(res['bf16']['loss']):.6f}; total {time.time() - t0:.0f}s", flush=True)
    del ref16, out, lg
    gc.collect()
    return res


def main():
    torch.manual_seed(0)
    out_dir = os.path.join(REPO, "tests", "golden")
    only = sys.argv[1] if len(sys.argv) > 1 else None
    for name, cfg, case in (("realwidth_b", REAL_B, "B"), ("realwidth_a", REAL_A, "A")):
        if only and only != case:
            continue
        res = run_case(name, cfg, case)
        path = os.path.join(out_dir, name + ".pt")
        torch.save(res, path)
        print(name, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
