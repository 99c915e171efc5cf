"""Continuous batching vs static batches at LLaMA-3-8B dims on one B200: 24 requests with 128-token
prompts and 96..512 new positions (teacher-forced schedules with visual-token runs, weights are random) through 8 slots."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from metamorph_b200 import synthetic  # noqa: E402
from metamorph_b200.constants import IMAGE_END_TOKEN_ID, IMAGE_START_TOKEN_ID  # noqa: E402
from metamorph_b200.engine.serve import ContinuousBatcher  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    model = synthetic.build_model(synthetic.make_config(), device=dev)
    model.eval()
    g = torch.Generator().manual_seed(7)
    n_req, P = 24, 128
    lens = torch.randint(96, 513, (n_req,), generator=g).tolist()
    reqs = []
    for n in lens:
        prompt = torch.randint(0, 128000, (1, P), generator=g)
        sched = torch.randint(0, 128000, (n + 2,), generator=g).to(torch.int32)
        for s in range(20, n - 70, 150):                       # a 64-embedding image every ~150 positions
            sched[s] = IMAGE_START_TOKEN_ID
            sched[s + 65] = IMAGE_END_TOKEN_ID
        reqs.append((model.get_model().embed_tokens(prompt.to(dev)), n, sched))
    total_positions = sum(n + 1 for n in lens)

    # ---- continuous batching
    srv = ContinuousBatcher(model, max_slots=8, max_context=1024, max_new_tokens=512, poll_every=8)
    for e, n, f in reqs[:8]:                                   # warm-up pass (kernel attributes, graph capture)
        srv.submit(e, max_new_tokens=8, forced_tokens=f)
    srv.run_until_idle()
    torch.cuda.synchronize()
    for e, n, f in reqs:
        srv.submit(e, max_new_tokens=n, forced_tokens=f)
    steps0 = srv.steps_run
    t0 = time.perf_counter()
    out = srv.run_until_idle()
    torch.cuda.synchronize()
    dt_c = time.perf_counter() - t0
    got = sum(int(i.numel() + im.shape[0]) for i, im in out.values())

    # ---- static batches of 8 run to the longest member (DecodeEngine)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    static_steps = 0
    for i in range(0, n_req, 8):
        chunk = reqs[i:i + 8]
        nmax = max(n for _, n, _ in chunk)
        emb = torch.cat([e for e, _, _ in chunk], 0)
        forced = torch.stack([torch.cat([f, f.new_full((nmax + 2 - f.numel(),), 128009)]) for _, _, f in chunk])
        model.greedy_decode(None, None, emb, max_new_tokens=nmax, output_image=True, forced_tokens=forced)
        static_steps += model._decode.last_steps
    torch.cuda.synchronize()
    dt_s = time.perf_counter() - t0
    print(json.dumps({
        "workload": f"{n_req} requests, prompt {P}, new positions {min(lens)}..{max(lens)} (sum {total_positions}), 8 slots",
        "continuous": {"wall_s": dt_c, "device_steps": srv.steps_run - steps0, "positions_out": got,
                       "positions_per_s": got / dt_c},
        "static_batches_of_8": {"wall_s": dt_s, "device_steps": static_steps,
                                "positions_per_s": total_positions / dt_s,
                                "note": "includes prefill and one graph capture per batch, like any greedy_decode call"},
    }))


if __name__ == "__main__":
    main()
