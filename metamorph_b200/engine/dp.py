"""Data-parallel host logic: the path shards by sample, images travel with their sample,
one gradient all-reduce (sum, then 1/world inside the fused AdamW) per bucket. Backend-agnostic so the
N>1 logic is testable with gloo on CPU; on the GPU box the backend is NCCL over NVLink/NVSwitch."""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.distributed as dist

from ..constants import IMAGE_TOKEN_INDEX


def images_per_sample(input_ids: torch.Tensor, attention_mask=None) -> List[int]:
    """Number of image slots each sample consumes: one per <image> placeholder, and ONE dummy image for
    a sample without any (metamorph_arch.py:275-284, train.py:1239-1242)."""
    ids = input_ids if attention_mask is None else input_ids.masked_fill(~attention_mask.bool(), 0)
    n = (ids == IMAGE_TOKEN_INDEX).sum(-1).tolist()
    return [max(1, int(x)) for x in n]


def shard_global_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Contiguous split of a global batch by sample; the flat `images` tensor is split so that every
    sample keeps its own images (order = occurrence order, as the reference collator builds it)."""
    B = batch["input_ids"].shape[0]
    assert B % world == 0, f"global batch {B} not divisible by world size {world}"
    per = B // world
    lo, hi = rank * per, (rank + 1) * per
    counts = images_per_sample(batch["input_ids"], batch.get("attention_mask"))
    i0, i1 = sum(counts[:lo]), sum(counts[:hi])
    out = {k: v[lo:hi] for k, v in batch.items() if k != "images" and torch.is_tensor(v)}
    out["images"] = batch["images"][i0:i1]
    return out


def all_reduce_sum_(buffers: List[torch.Tensor], group=None) -> None:
    for b in buffers:
        dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)


def max_over_ranks(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
