"""Sequence packing vs padding at LLaMA-3-8B dims on one B200: a ragged batch of 4 samples
(4096 / 1800 / 1200 / 900 interleaved positions, 4 images each), full train step (fwd + bwd + AdamW)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from metamorph_b200 import synthetic  # noqa: E402
from metamorph_b200.engine.trainer import TrainEngine  # noqa: E402


def ragged_batch(lengths, seed=1234):
    parts = [synthetic.train_batch(1, T, seed=seed + 17 * i, pin=False) for i, T in enumerate(lengths)]
    L = max(p["input_ids"].shape[1] for p in parts)
    ids = torch.zeros((len(parts), L), dtype=parts[0]["input_ids"].dtype)
    labs = torch.full((len(parts), L), -100, dtype=parts[0]["labels"].dtype)
    mask = torch.zeros((len(parts), L), dtype=torch.bool)
    for i, p in enumerate(parts):
        n = p["input_ids"].shape[1]
        ids[i, :n], labs[i, :n], mask[i, :n] = p["input_ids"][0], p["labels"][0], True
    images = torch.cat([p["images"] for p in parts]).pin_memory()
    return dict(input_ids=ids, labels=labs, attention_mask=mask, images=images)


def timed(engine, batch, warmup=2, steps=3):
    for _ in range(warmup):
        out = engine.step(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = engine.step(batch)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, float(out["loss"]), int(out["tokens"])


def main():
    dev = torch.device("cuda", 0)
    lengths = [4096, 1800, 1200, 900]
    model = synthetic.build_model(synthetic.make_config(), device=dev)
    batch = ragged_batch(lengths)
    # lr = 0: the optimizer runs (full step cost) but the weights stay put, so both layouts see the same model and
    # their losses must agree
    eng = TrainEngine(model, lr=0.0, constant_lr=True, n_save_gu_layers=32)
    ms_pad, loss_pad, pos_pad = timed(eng, batch)
    eng.pack_sequences = True
    ms_pack, loss_pack, pos_pack = timed(eng, batch)
    valid = sum(lengths)
    print(json.dumps({"workload": f"ragged batch {lengths} (valid positions {valid}), LLaMA-3-8B + SigLIP, fwd+bwd+AdamW",
                      "padded": {"ms_per_step": ms_pad, "positions_processed": pos_pad, "valid_tokens_per_s": valid / ms_pad * 1e3,
                                 "loss": loss_pad},
                      "packed": {"ms_per_step": ms_pack, "positions_processed": pos_pack,
                                 "valid_tokens_per_s": valid / ms_pack * 1e3, "loss": loss_pack},
                      "speedup": ms_pad / ms_pack, "loss_rel_diff": abs(loss_pad - loss_pack) / abs(loss_pad)}))


if __name__ == "__main__":
    main()
