"""ORACLE (test infrastructure): time THE REFERENCE's own code for the hot path on the host cores.

    python oracle/ref_bench.py train  --steps K --warmup W [--layers 2] [--dtype auto|f32|bf16] [--budget-s 240]
    python oracle/ref_bench.py decode --new-tokens 32 --prompt-len 128 [--layers 2]

Runs the unmodified reference (vendored by oracle/build_ref.py into oracle/_ref, or /root/reference when present):
`MetaMorphLlamaForCausalLM.forward` + `loss.backward()` (metamorph_llama.py:603-660 -> :285-498: SigLIP tower,
projector, prepare_inputs_labels_for_multimodal, the HF LLaMA stack, lm_head + cross entropy, vision head + cosine loss)
and `generate()` -> `greedy_decode` (metamorph_llama.py:666-717, :502-597; no KV cache) — through the reference's public
API, HF random init, on one synthetic sample of the benchmark's shape (B=1, T=4096 interleaved positions, 2 prompt-side
+ 2 answer-side images). A whole 8 B fp32 step needs > 64 GB and tens of minutes on host cores, so the model is
FULL WIDTH but DEPTH-REDUCED: `--layers` LLaMA layers (2, or 1 if the time budget demands it) and 2
SigLIP layers; everything else (lm_head over all T rows, both losses, the index logic) runs as the reference runs it.
Each line of output is one JSON object; the last one is the summary bench.py reads. bench.py extrapolates tokens/s to
the 32-layer / 27-layer step by the ratio of algorithmic FLOPs and labels it as such.
Used only by bench.py's CPU legs; never imported by the product.
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

H, I, HQ, HKV, DH, V = 4096, 14336, 32, 8, 128, 128258
SIGLIP = dict(siglip_width=1152, siglip_inter=4304, siglip_heads=16, image_size=384)


def reference_root():
    for cand in (os.path.join(HERE, "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(cand, "metamorph", "model")):
            return cand
    return None


def pick_threads():
    """All host threads the reference can USE: torch's CPU GEMMs get slower past some thread count on many-core boxes
    (oversubscription, NUMA), so a short calibration picks the fastest of {all, 96, 64, 48, 32, 16} threads."""
    import torch
    n = os.cpu_count() or 1
    cands = sorted({c for c in (n, 96, 64, 48, 32, 16) if 1 <= c <= n}, reverse=True)
    a, b = torch.randn(1024, 4096), torch.randn(4096, 14336)
    best, best_t = n, float("inf")
    for c in cands:
        torch.set_num_threads(c)
        torch.mm(a, b)
        t0 = time.perf_counter()
        torch.mm(a, b)
        torch.mm(a, b)
        dt = time.perf_counter() - t0
        if dt < best_t * 0.97:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cfg_for(layers, siglip_layers):
    return dict(hidden=H, layers=layers, heads=HQ, kv_heads=HKV, head_dim=DH, inter=I, vocab=V, rms_eps=1e-5,
                rope_theta=500000.0, siglip_layers=siglip_layers, image_tokens=64, max_len=8192, vision_coef=1.0, **SIGLIP)


def sample_flops(T, layers, siglip_layers, n_images):
    """Algorithmic FLOPs of one reference forward+backward on the sample (SURVEY.md section 8d conventions: 6 flop per
    dense parameter per token, causal-halved attention, forward-only frozen tower)."""
    per_layer = H * (HQ * DH + 2 * HKV * DH + HQ * DH) + 3 * H * I
    dense = layers * per_layer + H * V
    vision = 2 * 729 * (siglip_layers * (4 * 1152 ** 2 + 2 * 1152 * 4304) + 588 * 1152) + siglip_layers * 4 * 729 ** 2 * 1152
    return T * (6 * dense + 6 * layers * H * T) + n_images * vision


def build(layers, siglip_layers, dtype):
    import torch
    from oracle.ref_model import build_reference
    torch.manual_seed(0)
    m = build_reference(cfg_for(layers, siglip_layers), None, dtype, attn="sdpa")   # 4.45's default attention for LLaMA
    m.train()
    for p in m.get_vision_tower().parameters():   # freeze_vision=True (scripts/*.sh)
        p.requires_grad_(False)
    return m


def train_batch(T):
    import numpy as np
    import torch
    from oracle.weights import interleaved_sample
    rng = np.random.default_rng(1234)
    s, l = interleaved_sample(rng, T, 2, 2)
    ids = torch.tensor([s])
    labs = torch.tensor([l])
    mask = torch.ones_like(ids, dtype=torch.bool)
    images = torch.from_numpy(rng.standard_normal((4, 3, 384, 384), dtype=np.float32))
    return ids, mask, labs, images


def one_train_step(model, batch, dtype):
    ids, mask, labs, images = batch
    model.zero_grad(set_to_none=True)
    t0 = time.perf_counter()
    out = model(input_ids=ids, attention_mask=mask, labels=labs, images=images.to(dtype))
    out.loss.backward()
    dt = time.perf_counter() - t0
    return dt, float(out.loss.detach())


def run_train(a):
    import torch
    threads = a.threads or pick_threads()
    torch.set_num_threads(threads)
    dts = {"f32": torch.float32, "bf16": torch.bfloat16}
    batch = train_batch(a.seq_len)
    total = a.steps + a.warmup
    layers = a.layers
    probe = {}
    order = ["f32", "bf16"] if a.dtype == "auto" else [a.dtype]
    models = {}
    for name in order:   # one step per candidate dtype: warms the pools and decides `auto`
        models[name] = build(layers, a.siglip_layers, dts[name])
        dt, loss = one_train_step(models[name], batch, dts[name])
        probe[name] = dt
        print(json.dumps({"probe": name, "layers": layers, "seconds": dt, "loss": loss}), flush=True)
    dtype = min(probe, key=probe.get)
    for name in list(models):
        if name != dtype:
            del models[name]
    # bounded sample: keep the whole run inside the budget by dropping to one decoder layer if needed
    if layers > 1 and probe[dtype] * total > a.budget_s:
        layers = 1
        models = {dtype: build(layers, a.siglip_layers, dts[dtype])}
        dt, loss = one_train_step(models[dtype], batch, dts[dtype])
        print(json.dumps({"probe": dtype, "layers": layers, "seconds": dt, "loss": loss}), flush=True)
    model = models[dtype]
    secs = []
    for i in range(total):
        dt, loss = one_train_step(model, batch, dts[dtype])
        if i >= a.warmup:
            secs.append(dt)
        print(json.dumps({"step": i, "timed": i >= a.warmup, "seconds": dt, "loss": loss}), flush=True)
    fl = sample_flops(a.seq_len, layers, a.siglip_layers, 4)
    print(json.dumps({"summary": "train", "kind": "reference", "root": reference_root(), "dtype": dtype, "threads": threads,
                      "host_threads": os.cpu_count(), "layers": layers, "siglip_layers": a.siglip_layers,
                      "seq_len": a.seq_len, "batch": 1, "images": 4, "seconds": secs, "probe_seconds": probe,
                      "sample_flops": fl, "attention": "sdpa", "torch": torch.__version__}), flush=True)
    return 0


def run_decode(a):
    import torch
    from oracle.ref_model import pin_decode_mask_semantics
    threads = a.threads or pick_threads()
    torch.set_num_threads(threads)
    dt_t = torch.float32 if a.dtype in ("auto", "f32") else torch.bfloat16
    model = build(a.layers, a.siglip_layers, dt_t)
    model.eval()
    pin_decode_mask_semantics(model)
    g = torch.Generator().manual_seed(4321)
    prompt = torch.randint(0, 128000, (1, a.prompt_len), generator=g)
    t0 = time.perf_counter()
    # eos / image-start ids that never match: exactly `new_tokens` no-cache steps of text decoding
    out = model.generate(prompt, max_new_tokens=a.new_tokens - 1, eos_token_id=[-1], start_image_token_id=-1)
    dt = time.perf_counter() - t0
    n = int(out[0].numel())
    per_layer = H * (HQ * DH + 2 * HKV * DH + HQ * DH) + 3 * H * I
    # algorithmic FLOPs the no-cache loop executes: every step re-runs the whole prefix and lm_head over all positions
    fl = sum(2 * (a.layers * per_layer + H * V) * (a.prompt_len + t) + 4 * a.layers * H * (a.prompt_len + t) ** 2 // 2
             for t in range(n))
    print(json.dumps({"summary": "decode", "kind": "reference", "root": reference_root(), "dtype": "f32" if dt_t == torch.float32 else "bf16",
                      "threads": threads, "host_threads": os.cpu_count(), "layers": a.layers, "prompt_len": a.prompt_len,
                      "new_tokens": n, "seconds": dt, "executed_flops": fl}), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["train", "decode"])
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--siglip-layers", type=int, default=2)
    ap.add_argument("--seq-len", type=int, default=4096)
    ap.add_argument("--dtype", default="auto", choices=["auto", "f32", "bf16"])
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--budget-s", type=float, default=240.0)
    ap.add_argument("--prompt-len", type=int, default=128)
    ap.add_argument("--new-tokens", type=int, default=32)
    a = ap.parse_args()
    root = reference_root()
    if root is None:
        print(json.dumps({"summary": a.mode, "error": "no reference checkout (oracle/_ref missing: run oracle/build_ref.py)"}))
        return 2
    from oracle.ref_model import activate
    activate(root)
    return run_train(a) if a.mode == "train" else run_decode(a)


if __name__ == "__main__":
    sys.exit(main())
