"""Decode-step profile helper: LLaMA-3-8B dims, batch 8, a few stream-launched steps (no CUDA graph)
so that `ncu --metrics gpu__time_duration.sum` lists every kernel of a step."""
import sys

import torch

sys.path.insert(0, ".")
from metamorph_b200 import synthetic  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    dev = torch.device("cuda", 0)
    cfg = synthetic.make_config()
    model = synthetic.build_model(cfg, device=dev)
    model.eval()
    model._decode.use_cuda_graph = False
    g = torch.Generator().manual_seed(1)
    prompts = torch.randint(0, 128000, (8, 128), generator=g)
    emb = model.get_model().embed_tokens(prompts.to(dev))
    forced = torch.randint(0, 128000, (8, steps + 4), generator=g).to(torch.int32)
    forced[:, 2] = 128256  # enter image mode at step 2
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ids, imgs = model.greedy_decode(None, None, emb, max_new_tokens=steps - 1, output_image=True, forced_tokens=forced)
    e1.record()
    torch.cuda.synchronize()
    print("decode", steps, "steps:", e0.elapsed_time(e1), "ms total", [int(x.numel()) for x in ids][:2],
          [int(x.shape[0]) for x in imgs][:2])


if __name__ == "__main__":
    main()
