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


def _adamw_ref(lr, step, scale, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.AdamW arithmetic (weight_decay 0) on explicit state tensors; writes the rounded compute copy."""
    def fn(p16, p32, m, v, g):
        g = g.float() * scale
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        mh, vh = m / (1 - b1 ** step), v / (1 - b2 ** step)
        p32.add_(-lr * mh / (vh.sqrt() + eps))
        p16.copy_(p32)
    return fn


def _shard_worker(rank, world, init_file, out_dir):
    """ZeRO-1 plumbing of engine/trainer.py on CPU tensors: reduce-scatter -> AdamW on the 1/world slice -> all-gather
    must leave EVERY rank with the parameters of the unsharded update on the summed gradient."""
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    from metamorph_b200.engine.trainer import ShardedBucket, bucket_reduce, bucket_update, shard_bounds
    g0 = torch.Generator().manual_seed(0)
    shapes = [(24, 16), (16, 16), (40, 16), (16, 24)]                       # one "layer": four matrices, 1664 elements
    ws = [torch.randn(s, generator=g0).bfloat16() for s in shapes]
    params = [(f"w{i}", torch.nn.Parameter(w.clone())) for i, w in enumerate(ws)]
    b = ShardedBucket("layer0", params, world, rank)
    assert b.numel == 1664 and (b.lo, b.hi) == shard_bounds(1664, world, rank)
    for (_, p), w in zip(params, ws):                                        # parameters are views of the flat buffer now
        assert p.data.data_ptr() >= b.flat.data_ptr() and torch.equal(p.data, w)
    # per-rank gradients (bf16, as the wgrad GEMMs write them)
    gr = torch.Generator().manual_seed(100 + rank)
    grad = torch.randn(b.numel, generator=gr).bfloat16()
    both = [torch.randn(b.numel, generator=torch.Generator().manual_seed(100 + r)).bfloat16() for r in range(world)]
    for step in (1, 2):
        gs = bucket_reduce(b, grad.clone(), world)
        bucket_update(b, gs, _adamw_ref(1e-2, step, 1.0 / world))
    # unsharded reference: same arithmetic on the whole flat vector with the summed gradient
    flat = torch.cat([w.reshape(-1) for w in ws])
    p32, m, v = flat.float(), torch.zeros(b.numel), torch.zeros(b.numel)
    gsum = (both[0].float() + both[1].float()).bfloat16() if world == 2 else both[0]
    out = torch.empty_like(flat)
    for step in (1, 2):
        _adamw_ref(1e-2, step, 1.0 / world)(out, p32, m, v, gsum)
    # gloo reduces bf16 in bf16; the reference above rounds the sum the same way
    assert torch.equal(b.flat, out), float((b.flat.float() - out.float()).abs().max())
    assert torch.equal(b.p32, p32[b.lo:b.hi])
    for (_, p), shp in zip(params, shapes):
        assert tuple(p.shape) == shp
    # tensors wholly inside this rank's slice expose per-tensor state views (what checkpoints at world 1 are made of)
    views = b.param_views()
    assert all(torch.equal(st.p16, dict(params)[n].data) for n, st in views.items())
    torch.save(b.flat.clone(), os.path.join(out_dir, f"flat{rank}.pt"))
    dist.destroy_process_group()


def test_sharded_optimizer_world2_gloo_matches_unsharded_update():
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_shard_worker, args=(2, os.path.join(d, "init"), d), nprocs=2, join=True)
        a, b = torch.load(os.path.join(d, "flat0.pt")), torch.load(os.path.join(d, "flat1.pt"))
        assert torch.equal(a, b)                              # both replicas hold the same updated parameters


def test_shard_bounds_cover_the_bucket():
    from metamorph_b200.engine.trainer import shard_bounds
    n = 218103808                                             # one LLaMA-3-8B decoder layer's four matrices
    for world in (1, 2, 4, 8):
        cuts = [shard_bounds(n, world, r) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
        assert all((hi - lo) % 8 == 0 for lo, hi in cuts)
    assert 128258 * 4096 % 64 == 0                            # embedding / lm_head cut into 8 aligned slices
