"""CPU tests of the N>1 host logic with the gloo backend, world_size 2 (no GPU needed)."""
import os
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, init_file, out_dir):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    from metamorph_b200 import synthetic
    from metamorph_b200.engine import dp
    gb = synthetic.train_batch(4, 256, n_prompt_images=1, n_answer_images=1, image_tokens=16, image_size=28, pin=False)
    # make sample 1 text-only: it still owns one (dummy) image slot
    gb["input_ids"][1][gb["input_ids"][1] == -200] = 5
    gb["images"] = torch.arange(7, dtype=torch.float32).view(7, 1, 1, 1).expand(7, 3, 28, 28).contiguous()
    shard = dp.shard_global_batch(gb, rank, world)
    assert shard["input_ids"].shape[0] == 2
    counts = dp.images_per_sample(gb["input_ids"])
    assert counts == [2, 1, 2, 2]
    exp_imgs = [0, 1, 2] if rank == 0 else [3, 4, 5, 6]
    assert shard["images"][:, 0, 0, 0].tolist() == [float(x) for x in exp_imgs]
    # plan per shard is consistent with the number of images it received
    from metamorph_b200.model.interleave_plan import build_interleave_plan
    plan = build_interleave_plan(shard["input_ids"], shard["attention_mask"], shard["labels"],
                                 shard["images"].shape[0], 16, 4096)
    assert plan.batch == 2
    # gradient bucket: sum across ranks then 1/world == mean of per-rank gradients
    g = torch.full((1000,), float(rank + 1))
    dp.all_reduce_sum_([g])
    assert torch.allclose(g / world, torch.full((1000,), 1.5))
    assert dp.max_over_ranks(float(rank), "cpu") == 1.0
    torch.save(plan.labels, os.path.join(out_dir, f"labels{rank}.pt"))
    dist.destroy_process_group()


def test_dp_world2_gloo():
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "init")
        mp.spawn(_worker, args=(2, init_file, d), nprocs=2, join=True)
        a, b = torch.load(os.path.join(d, "labels0.pt")), torch.load(os.path.join(d, "labels1.pt"))
        assert a.shape[0] == 2 and b.shape[0] == 2


def test_cosine_lr_schedule_matches_hf_formula():
    import math
    from metamorph_b200.engine.trainer import cosine_lr
    total, base = 1000, 6.93e-5
    warm = math.ceil(total * 0.03)
    assert cosine_lr(0, total, base) == 0.0
    assert abs(cosine_lr(warm, total, base) - base) < 1e-12
    assert abs(cosine_lr(total, total, base)) < 1e-12
    mid = warm + (total - warm) // 2
    assert abs(cosine_lr(mid, total, base) - base * 0.5 * (1 + math.cos(math.pi * (mid - warm) / (total - warm)))) < 1e-15
