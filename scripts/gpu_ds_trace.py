"""Per-phase timeline of the one-kernel decode step (MM_DS_TRACE=1): globaltimer stamps of every CTA at the
phase boundaries of the last step, summarised per interval (mean / max over CTAs, averaged over layers)."""
import ctypes
import os
import sys

import numpy as np
import torch

os.environ["MM_DS_TRACE"] = "1"
sys.path.insert(0, ".")
from metamorph_b200 import synthetic  # noqa: E402
from metamorph_b200._lib import lib  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    model = synthetic.build_model(synthetic.make_config(), device=dev)
    model.eval()
    model._decode.use_cuda_graph = False
    g = torch.Generator().manual_seed(1)
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    prompts = torch.randint(0, 128000, (8, P), generator=g)
    emb = model.get_model().embed_tokens(prompts.to(dev))
    steps = 6
    forced = torch.randint(0, 128000, (8, steps + 4), generator=g).to(torch.int32)
    model.greedy_decode(None, None, emb, max_new_tokens=steps - 1, output_image=True, forced_tokens=forced)
    torch.cuda.synchronize()
    plan = model._decode.last_stack_plan
    L, H, Hq, Hkv, inter, B = plan.dims
    fn = lib().mm_decode_stack_trace_offset
    fn.restype = ctypes.c_longlong
    off = int(fn(B, H, Hq, Hkv, inter))
    nsm = torch.cuda.get_device_properties(0).multi_processor_count
    raw = plan.workspace[off:off + L * 16 * nsm * 8].cpu().numpy().view(np.uint64).reshape(L, 16, nsm).astype(np.int64)
    names = {0: "qkv tiles start", 1: "qkv tiles done", 2: "barrier", 3: "attention done", 7: "barrier",
             4: "o_proj tiles start (staged)", 5: "o_proj tiles done", 6: "barrier",
             8: "gate/up tiles start (staged)", 9: "gate/up tiles done", 10: "barrier",
             12: "down tiles start", 13: "down tiles done", 14: "barrier"}
    order = [0, 1, 2, 3, 7, 4, 5, 6, 8, 9, 10, 12, 13, 14]
    t0 = raw[0, 0].min()
    print(f"whole stack: {(raw[L - 1, 13].max() - t0) / 1e3:.1f} us over {L} layers")
    rows = []
    for l in range(1, L - 1):
        prev = raw[l - 1, 14]          # barrier after the previous layer's down_proj, per CTA
        for ev in order:
            cur = raw[l, ev]
            rows.append((ev, (cur - prev).mean(), (cur - prev).max(), (cur - prev).min(), cur.max() - cur.min()))
            prev = cur
    print(f"{'interval ending at':34s} {'mean':>8s} {'max':>8s} {'min':>8s} {'skew(max-min of stamp)':>24s}   [us, avg over layers]")
    tot = 0.0
    for ev in order:
        sel = np.array([r[1:] for r in rows if r[0] == ev])
        m = sel.mean(axis=0) / 1e3
        tot += m[0]
        print(f"{names[ev]:34s} {m[0]:8.2f} {m[1]:8.2f} {m[2]:8.2f} {m[3]:24.2f}")
    print(f"sum of mean intervals per layer: {tot:.2f} us")


if __name__ == "__main__":
    main()
