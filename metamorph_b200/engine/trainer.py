"""Train step driver: hot path forward/backward + fused AdamW + the data-parallel gradient collective.

Replaces, for this path, HF Trainer + Accelerate + DeepSpeed ZeRO-2 (scripts/zero2.json:16-24:
`stage 2`, `overlap_comm`, `reduce_scatter`, `contiguous_gradients`):
  * data parallel, one process per GPU (torchrun), NCCL over NVLink/NVSwitch, parameters replicated in bf16;
  * OPTIMIZER STATE SHARDED (ZeRO-1/2 style): every large tensor group ("bucket": one decoder layer's four matrices in
    ONE flat bf16 buffer, the embedding table, lm_head) keeps fp32 master weights and both AdamW moments only for this
    rank's 1/world slice of the flat index space. Per bucket: reduce-scatter of the gradient (the rank receives the SUM
    of its slice) -> fused AdamW on the slice -> all-gather of the updated bf16 slice straight into the flat parameter
    buffer every rank computes with. Same wire volume as the all-reduce it replaces; AdamW's HBM traffic and the
    optimizer memory drop by `world` (8 GPUs: 97 -> 12 GB of state per GPU, which is what lets config 3 run its stated
    8 samples per GPU). world == 1 is the same code with a slice that is the whole bucket and no collective;
  * "optimizer in the backward sweep": as soon as a layer's wgrad GEMMs retire, its bucket is reduced and applied on a
    side stream while the main stream continues with the next layer's backward. Without accumulation only two
    layer-sized gradient buffers exist (they rotate), so the 16 GB of full-model gradients are never resident;
  * `gradient_accumulation_steps` = k (HF TrainingArguments; scripts/*.sh pass it): `step()` takes k micro-batches, the
    wgrad GEMMs accumulate into resident per-layer gradient buffers (`accumulate` epilogue), and the collective + AdamW
    run once, during the LAST micro-batch's backward sweep, on the mean over micro-batches (HF divides each micro loss
    by k);
  * optional global-norm clipping (`max_grad_norm`): gradients stay resident, the squared norm is summed over the
    reduce-scattered slices (+ the small replicated tensors once) and the clip factor is applied inside AdamW
    (the reference scripts pass no max_grad_norm);
  * parameters that do not train (stage 1, `tune_mm_mlp_adapter`: everything but the projector is frozen,
    train.py:1516-1519) get no wgrad GEMMs at all — only the dgrad chain down to the projector runs.
Small tensors (norm weights, projector, vision head: ~45 M parameters) keep replicated state and an all-reduce.
Optimizer = torch.optim.AdamW semantics (reference: --optim adamw_torch, train.py:82), cosine LR with 3 % warm-up
(scripts/*.sh: lr_scheduler_type cosine, warmup_ratio 0.03), weight decay 0.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .. import ops
from .hot_path import GradProvider, HotPath
from .llama import LayerGrads


def cosine_lr(step: int, total_steps: int, base_lr: float, warmup_ratio: float = 0.03) -> float:
    """transformers.get_cosine_schedule_with_warmup (HF Trainer default for lr_scheduler_type=cosine)."""
    warm = math.ceil(total_steps * warmup_ratio)
    if step < warm:
        return base_lr * step / max(1, warm)
    prog = (step - warm) / max(1, total_steps - warm)
    return base_lr * max(0.0, 0.5 * (1.0 + math.cos(math.pi * prog)))


def shard_bounds(numel: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of this rank's slice of a flat bucket; numel must be a multiple of 8 * world (16-byte slices)."""
    assert numel % (8 * world) == 0, f"bucket of {numel} elements cannot be cut into {world} 16-byte-aligned slices"
    n = numel // world
    return rank * n, (rank + 1) * n


class _OptState:
    """fp32 master / exp_avg / exp_avg_sq of ONE parameter tensor, as views into its bucket's state when the whole
    tensor lies inside this rank's slice (always at world == 1), standalone tensors for replicated small parameters."""
    __slots__ = ("p16", "p32", "m", "v")

    def __init__(self, p16, p32, m, v):
        self.p16, self.p32, self.m, self.v = p16, p32, m, v


class ShardedBucket:
    """One flat bf16 parameter buffer (the tensors' `.data` are views into it) + this rank's slice of the optimizer state."""

    def __init__(self, name: str, params: Sequence[Tuple[str, torch.nn.Parameter]], world: int, rank: int,
                 grad_dtype=torch.bfloat16, symmetric: bool = False):
        """symmetric=True (world > 1): the flat parameter buffer is allocated as SYMMETRIC memory (same virtual layout on
        every rank, mapped into every peer over NVLink, with an NVSwitch multicast address when the fabric has NVLS) so
        that the optimizer kernel can write its updated slice into all replicas itself (`mm_adamw_step_bcast`)."""
        self.name, self.world, self.rank = name, world, rank
        self.symm = None
        self.names = [n for n, _ in params]
        self.shapes = [tuple(p.shape) for _, p in params]
        self.sizes = [p.numel() for _, p in params]
        self.numel = sum(self.sizes)
        self.grad_dtype = grad_dtype
        dev = params[0][1].device
        if len(params) == 1 and params[0][1].data.is_contiguous() and not (symmetric and world > 1):
            self.flat = params[0][1].data.view(-1)                       # a single tensor is its own flat buffer
        else:
            if symmetric and world > 1:
                import torch.distributed._symmetric_memory as symm_mem
                self.flat = symm_mem.empty(self.numel, dtype=torch.bfloat16, device=dev)
                self.symm = symm_mem.rendezvous(self.flat, dist.group.WORLD)          # collective: same order on all ranks
            else:
                self.flat = torch.empty(self.numel, dtype=torch.bfloat16, device=dev)
            off = 0
            for (_, p), n in zip(params, self.sizes):
                self.flat[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.flat[off:off + n].view(p.shape)          # the model now computes with views of the bucket
                off += n
        self.lo, self.hi = shard_bounds(self.numel, world, rank)
        sl = self.flat[self.lo:self.hi]
        self.p32 = sl.float()
        self.m = torch.zeros_like(self.p32)
        self.v = torch.zeros_like(self.p32)
        # staging of the NCCL path (world > 1): the reduced gradient slice and the updated bf16 slice, allocated on first
        # use (the fused NVSwitch path needs neither)
        self._gshard = self._pshard = None

    @property
    def gshard(self) -> torch.Tensor:
        if self._gshard is None:
            self._gshard = torch.empty(self.hi - self.lo, dtype=self.grad_dtype, device=self.flat.device)
        return self._gshard

    @property
    def pshard(self) -> torch.Tensor:
        if self._pshard is None:
            self._pshard = torch.empty(self.hi - self.lo, dtype=torch.bfloat16, device=self.flat.device)
        return self._pshard


    def param_views(self) -> Dict[str, _OptState]:
        """Per-tensor views of the state for tensors that lie entirely inside this rank's slice."""
        out, off = {}, 0
        for n, shp, sz in zip(self.names, self.shapes, self.sizes):
            if off >= self.lo and off + sz <= self.hi:
                a, b = off - self.lo, off - self.lo + sz
                out[n] = _OptState(self.flat[off:off + sz].view(shp), self.p32[a:b].view(shp), self.m[a:b].view(shp),
                                   self.v[a:b].view(shp))
            off += sz
        return out

    def grad_views(self, flat_grad: torch.Tensor) -> List[torch.Tensor]:
        views, off = [], 0
        for shp, sz in zip(self.shapes, self.sizes):
            views.append(flat_grad[off:off + sz].view(shp))
            off += sz
        return views


def bucket_reduce(b: ShardedBucket, flat_grad: torch.Tensor, world: int) -> torch.Tensor:
    """This rank's slice of the gradient SUM over ranks (the whole buffer when the state is not sharded)."""
    if world == 1:
        return flat_grad
    if b.world == 1:                                   # replicated state: plain all-reduce
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        return flat_grad
    dist.reduce_scatter_tensor(b.gshard, flat_grad, op=dist.ReduceOp.SUM)
    return b.gshard


def bucket_update(b: ShardedBucket, gslice: torch.Tensor, adamw, adamw_bcast=None) -> None:
    """`adamw(p16_out, p32, m, v, grad)` on this rank's slice; the updated bf16 slice then reaches every rank's flat
    parameter buffer — written there by the optimizer kernel itself when the bucket lives in symmetric memory
    (`adamw_bcast(bucket, p32, m, v, grad)`: AdamW + all-gather in ONE kernel over NVSwitch multicast / NVLink P2P), else
    through an NCCL all-gather."""
    if b.world == 1:
        adamw(b.flat, b.p32, b.m, b.v, gslice)
        return
    if b.symm is not None and adamw_bcast is not None:
        adamw_bcast(b, b.p32, b.m, b.v, gslice)
        return
    adamw(b.pshard, b.p32, b.m, b.v, gslice)
    dist.all_gather_into_tensor(b.flat, b.pshard)


def probe_hdl_multicast(device) -> int:
    """Multicast (NVLS) address of a scratch symmetric allocation, 0 when the fabric has none. Collective."""
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1024, dtype=torch.bfloat16, device=device)
    return int(symm_mem.rendezvous(t, dist.group.WORLD).multicast_ptr or 0)


class BucketGradProvider(GradProvider):
    """Hands the hot path its gradient destinations and fires a bucket's collective + AdamW as soon as it is complete.

    rotating mode (no accumulation, no clipping): two layer-sized flat buffers alternate between layers;
    resident mode: one flat buffer per layer (needed to accumulate over micro-batches / to clip by the global norm)."""

    def __init__(self, model, engine: "TrainEngine", resident: bool):
        super().__init__(model)
        self.e = engine
        self.resident = resident
        self.apply_now = True          # False during all but the last micro-batch of an accumulation window
        self.defer = False             # clipping: collect everything, apply at the end of the step
        self.pending: List = []
        lb = engine.layer_buckets
        self.H = model.model.layers[0].input_layernorm.weight.shape[0]
        dev = model.device
        if lb:
            n_sets = len(lb) if resident else min(2, len(lb))
            self.sets = []
            for _ in range(n_sets):
                flat, hdl = engine.alloc_grad(lb[0].numel, torch.bfloat16)
                self.sets.append(dict(flat=flat, hdl=hdl, ln=torch.zeros(2 * self.H, dtype=torch.float32, device=dev),
                                      free=None))
        else:
            self.sets = []
        self.big_hdl: Dict[str, object] = {}       # symmetric-memory handles of the lm_head / embedding gradient buffers

    def get(self, name: str, like: torch.Tensor, fp32: bool = False, zero: bool = False) -> torch.Tensor:
        if name in self.e.big_buckets and name not in self.buffers and self.e.fused_reduce:
            # the gradient of a sharded big tensor lives in symmetric memory: its slice is summed over the ranks inside
            # the switch by the optimizer kernel (collective allocation: every rank reaches this point in the same step)
            flat, hdl = self.e.alloc_grad(like.numel(), torch.float32 if fp32 else like.dtype)
            self.buffers[name] = flat.view(like.shape)
            self.big_hdl[name] = hdl
            return self.buffers[name]
        return super().get(name, like, fp32=fp32, zero=zero)

    def _set(self, i: int):
        return self.sets[i if self.resident else i % len(self.sets)]

    def layer(self, i: int) -> LayerGrads:
        s = self._set(i)
        if s["free"] is not None:                      # bucket still being reduced / applied
            torch.cuda.current_stream().wait_event(s["free"])
            s["free"] = None
        v = self.e.layer_buckets[i].grad_views(s["flat"])
        if not self.accumulate:
            s["ln"].zero_()
        return LayerGrads(v[0], v[1], v[2], v[3], s["ln"][:self.H], s["ln"][self.H:])

    def layer_done(self, i: int, g: LayerGrads):
        if not self.apply_now:
            return
        s = self._set(i)
        l = self.model.model.layers[i]
        small = [(l.input_layernorm.weight, g.ln1), (l.post_attention_layernorm.weight, g.ln2)]
        if self.defer:
            self.pending.append((self.e.layer_buckets[i], s["flat"], small, [s["ln"]]))
            return
        s["free"] = self.e.reduce_and_apply(self.e.layer_buckets[i], s["flat"], small, [s["ln"]], grad_hdl=s["hdl"])

    GROUPS = {
        "heads": ["lm_head.weight", "vision_head.0.weight", "vision_head.0.bias", "vision_head.2.weight",
                  "vision_head.2.bias"],
        "final_norm": ["model.norm.weight"],
        "embed": ["model.embed_tokens.weight"],
        "projector": ["model.mm_projector.0.weight", "model.mm_projector.0.bias",
                      "model.mm_projector.2.weight", "model.mm_projector.2.bias"],
    }

    def group_done(self, group: str, written: Optional[Sequence[str]] = None):
        """`written`: the gradient buffers the hot path actually wrote this step (e.g. the vision head has none on a step
        without answer images) — the others must not be applied: their buffers hold a previous step's gradient, and the
        reference gives such parameters no gradient at all."""
        if not self.apply_now:
            return
        names = [n for n in self.GROUPS[group] if n in self.buffers and (written is None or n in written)]
        named = self.e.named_params
        for n in names:
            if n in self.e.big_buckets:
                if not self.defer:
                    self.e.reduce_and_apply(self.e.big_buckets[n], self.buffers[n].view(-1), [], [],
                                            grad_hdl=self.big_hdl.get(n))
                    continue
                item = (self.e.big_buckets[n], self.buffers[n].view(-1), [], [])
            elif n in self.e.small_state:
                item = (None, None, [(named[n], self.buffers[n])], [self.buffers[n]])
            else:
                continue
            if self.defer:
                self.pending.append(item)
            else:
                self.e.reduce_and_apply(*item)


class TrainEngine:
    def __init__(self, model, lr: float = 6.93e-5, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_grad_norm: Optional[float] = None, total_steps: int = 1000,
                 warmup_ratio: float = 0.03, constant_lr: bool = False, n_save_gu_layers: int = 0,
                 pack_sequences: bool = False, pack_len: Optional[int] = None,
                 gradient_accumulation_steps: int = 1, shard_optimizer: bool = True, fused_allgather: bool = True):
        self.model = model
        self.hot = HotPath(model)
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        self.total_steps, self.warmup_ratio, self.constant_lr = total_steps, warmup_ratio, constant_lr
        self.n_save_gu_layers = n_save_gu_layers
        # lay the samples of a batch end to end (block-diagonal attention) instead of padding them
        self.pack_sequences, self.pack_len = pack_sequences, pack_len
        self.accum = max(1, int(gradient_accumulation_steps))
        self.step_count = 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        # shard_optimizer=False keeps the whole state on every rank (slices = whole buckets, all-reduce instead of
        # reduce-scatter + all-gather): the round-1 behaviour, kept for A/B measurements
        self.shard_world = self.world if shard_optimizer else 1
        self.shard_rank = self.rank if shard_optimizer else 0
        # sharded state: fuse the all-gather of the updated parameters into the AdamW kernel (symmetric-memory buckets);
        # falls back to NCCL all-gather if symmetric memory cannot be set up on this system
        self.fused_allgather = bool(fused_allgather and self.shard_world > 1)
        self.use_multicast = os.environ.get("MM_ADAMW_MULTICAST", "1") != "0"     # 0: per-peer P2P stores (A/B, debugging)
        if self.fused_allgather:
            try:
                import torch.distributed._symmetric_memory as symm_mem   # noqa: F401
                probe = symm_mem.empty(1024, dtype=torch.bfloat16, device=model.device)
                symm_mem.rendezvous(probe, dist.group.WORLD)
            except Exception as e:  # noqa: BLE001
                import warnings
                warnings.warn(f"symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL all-gather")
                self.fused_allgather = False
            flag = torch.tensor([1 if self.fused_allgather else 0], device=model.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)           # every rank must take the same path
            self.fused_allgather = bool(int(flag))
        # ... and the reduce-scatter too: gradients in symmetric memory, summed over the ranks INSIDE the NVSwitch by the
        # optimizer kernel's multimem.ld_reduce -> a bucket's [reduce-scatter -> AdamW -> all-gather] is one kernel and
        # NCCL leaves the data path (needs multicast/NVLS support; not used with clipping, which needs the slice norms first)
        self.fused_reduce = bool(self.fused_allgather and self.use_multicast and max_grad_norm is None and
                                 os.environ.get("MM_FUSED_REDUCE", "1") != "0")
        if self.fused_reduce:
            has_mc = torch.tensor([1 if int(probe_hdl_multicast(model.device)) else 0], device=model.device)
            dist.all_reduce(has_mc, op=dist.ReduceOp.MIN)
            self.fused_reduce = bool(int(has_mc))
        self.named_params: Dict[str, torch.nn.Parameter] = dict(model.named_parameters())
        trainable = {n: p for n, p in self.named_params.items()
                     if p.requires_grad and "vision_tower" not in n and "vision_proj" not in n}

        # ---- buckets
        self.layer_buckets: List[ShardedBucket] = []
        self.big_buckets: Dict[str, ShardedBucket] = {}
        self.small_state: Dict[str, _OptState] = {}
        self.opt: Dict[str, _OptState] = {}
        L = len(model.model.layers)
        self.train_llm = all(f"model.layers.{i}.self_attn.qkv_proj.weight" in trainable for i in range(L)) and L > 0
        claimed = set()
        if self.train_llm:
            for i in range(L):
                p = f"model.layers.{i}."
                names = [p + "self_attn.qkv_proj.weight", p + "self_attn.o_proj.weight", p + "mlp.gate_up_proj.weight",
                         p + "mlp.down_proj.weight"]
                b = ShardedBucket(f"layer{i}", [(n, trainable[n]) for n in names], self.shard_world, self.shard_rank,
                                  symmetric=self.fused_allgather)
                self.layer_buckets.append(b)
                claimed.update(names)
        for n, gd in (("lm_head.weight", torch.float32), ("model.embed_tokens.weight", torch.bfloat16)):
            if n in trainable and trainable[n].numel() % (8 * self.shard_world) == 0 and trainable[n].numel() >= (1 << 22):
                self.big_buckets[n] = ShardedBucket(n, [(n, trainable[n])], self.shard_world, self.shard_rank, grad_dtype=gd,
                                                    symmetric=self.fused_allgather)
                claimed.add(n)
        for n, p in trainable.items():
            if n not in claimed:
                p32 = p.data.float()
                self.small_state[n] = _OptState(p.data, p32, torch.zeros_like(p32), torch.zeros_like(p32))
        for b in list(self.layer_buckets) + list(self.big_buckets.values()):
            self.opt.update(b.param_views())
        self.opt.update(self.small_state)
        self.trainable_names = list(trainable)
        # the tensors' old storages (now replaced by views of the flat buckets) and the fp32 staging copies went back to
        # the caching allocator in odd sizes: hand them to the driver once so that they do not sit on top of the step's
        # working set as reserved-but-unusable HBM (the 8 B model fills 171 of 180 GB)
        torch.cuda.empty_cache()
        self._state_by_ptr = {st.p16.data_ptr(): st for st in self.small_state.values()}
        # side stream: the gradient collectives (N > 1) and the HBM-bound fused AdamW run here, concurrently with the
        # tensor-core-bound backward GEMMs of the next layers on the main stream
        self.comm_stream = torch.cuda.Stream()
        self._bcast_pending = False
        self._fence = torch.zeros(1, dtype=torch.float32, device=model.device)
        resident = self.accum > 1 or max_grad_norm is not None
        self.provider = BucketGradProvider(model, self, resident=resident)
        self.provider.defer = max_grad_norm is not None
        self.kernel_launch_estimate = 0

    # -------------------------------------------------------------- optimizer plumbing
    def alloc_grad(self, numel: int, dtype):
        """A zeroed flat gradient buffer: symmetric memory (+ its handle) when the reduction is fused into the optimizer
        kernel, plain device memory otherwise. Collective in the symmetric case."""
        dev = self.model.device
        if not self.fused_reduce:
            return torch.zeros(numel, dtype=dtype, device=dev), None
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty(numel, dtype=dtype, device=dev)
        t.zero_()
        return t, symm_mem.rendezvous(t, dist.group.WORLD)

    def optimizer_state_bytes(self) -> int:
        n = sum(b.p32.numel() for b in list(self.layer_buckets) + list(self.big_buckets.values()))
        n += sum(st.p32.numel() for st in self.small_state.values())
        return 12 * n

    @torch.no_grad()
    def refresh_compute_copies(self):
        """bf16 compute copies <- rounded fp32 masters (after restoring the optimizer state from a checkpoint)."""
        for b in list(self.layer_buckets) + list(self.big_buckets.values()):
            if self.shard_world == 1:
                b.flat.copy_(b.p32)
            else:
                b.pshard.copy_(b.p32)
                dist.all_gather_into_tensor(b.flat, b.pshard)
        for st in self.small_state.values():
            st.p16.copy_(st.p32)

    def _adamw(self, p16, p32, m, v, grad, lr, grad_scale, scale_tensor=None):
        b1, b2 = self.betas
        ops.adamw_step_(p16.reshape(-1), p32.reshape(-1), m.reshape(-1), v.reshape(-1), grad.reshape(-1), lr=lr, beta1=b1,
                        beta2=b2, eps=self.eps, wd=self.wd, step=self.step_count, grad_scale=grad_scale,
                        grad_scale_tensor=scale_tensor)

    def _apply_small(self, params, lr, grad_scale, scale_tensor=None):
        for p, g in params:
            st = self._state_by_ptr.get(p.data.data_ptr())
            if st is not None:
                self._adamw(st.p16, st.p32, st.m, st.v, g, lr, grad_scale, scale_tensor)

    def _bucket_reduce(self, b: ShardedBucket, flat_grad: torch.Tensor) -> torch.Tensor:
        """(comm stream) -> this rank's slice of the summed gradient."""
        return bucket_reduce(b, flat_grad, self.world)

    def _bucket_update(self, b: ShardedBucket, gslice, lr, grad_scale, scale_tensor=None):
        """(comm stream) fused AdamW on the slice, then the updated bf16 slice goes back into every rank's flat buffer."""
        bucket_update(b, gslice, lambda p16, p32, m, v, g: self._adamw(p16, p32, m, v, g, lr, grad_scale, scale_tensor),
                      (lambda bk, p32, m, v, g: self._adamw_bcast(bk, p32, m, v, g, lr, grad_scale, scale_tensor))
                      if self.fused_allgather else None)
        self._bcast_pending = self._bcast_pending or (self.fused_allgather and b.symm is not None)

    def _adamw_bcast(self, b: ShardedBucket, p32, m, v, grad, lr, grad_scale, scale_tensor=None):
        """AdamW on the slice; the kernel writes the updated bf16 values into every rank's replica of the bucket."""
        b1, b2 = self.betas
        h = b.symm
        mc = int(h.multicast_ptr or 0)        # 0 when the fabric has no multicast (NVLS) support: per-peer stores instead
        if not self.use_multicast:
            mc = 0
        ops.adamw_step_bcast_(mc + 2 * b.lo if mc else 0, int(h.buffer_ptrs_dev), self.shard_world, b.lo, p32.reshape(-1),
                              m.reshape(-1), v.reshape(-1), grad.reshape(-1), lr=lr, beta1=b1, beta2=b2, eps=self.eps,
                              wd=self.wd, step=self.step_count, grad_scale=grad_scale, grad_scale_tensor=scale_tensor)

    def _fence_broadcasts(self):
        """The kernels of THIS rank are ordered by its streams; the slices the PEERS write into this rank's buffers are
        ordered by one tiny NCCL all-reduce at the end of the step's side-stream work: it cannot complete here before
        every rank has finished the broadcast kernels it enqueued ahead of it."""
        if self._bcast_pending:
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(self._fence, op=dist.ReduceOp.SUM)
            self._bcast_pending = False

    def reduce_and_apply(self, bucket: Optional[ShardedBucket], flat_grad, small_params, small_bufs, grad_hdl=None):
        """Collective + AdamW of one completed bucket (and/or of small replicated tensors) on the side stream.
        Returns an event marking when the gradient buffers may be reused."""
        lr = self.current_lr
        scale = 1.0 / (self.world * self.accum)
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ready)
            if bucket is not None and self.fused_reduce and grad_hdl is not None and bucket.symm is not None:
                # ONE kernel per bucket: in-switch sum of every rank's gradient slice (multimem.ld_reduce) -> AdamW ->
                # broadcast of the updated bf16 slice (multimem.st). Two cross-rank barriers order it: all ranks have
                # written their gradients / all ranks have read them (the buffer is reused two layers later).
                grad_hdl.barrier(channel=0)
                b1, b2 = self.betas
                esz = flat_grad.element_size()
                ops.adamw_step_bcast_(int(bucket.symm.multicast_ptr) + 2 * bucket.lo, 0, self.shard_world, bucket.lo,
                                      bucket.p32.reshape(-1), bucket.m.reshape(-1), bucket.v.reshape(-1), None, lr=lr,
                                      beta1=b1, beta2=b2, eps=self.eps, wd=self.wd, step=self.step_count, grad_scale=scale,
                                      grad_multicast_ptr=int(grad_hdl.multicast_ptr) + esz * bucket.lo,
                                      grad_f32=(flat_grad.dtype == torch.float32))
                grad_hdl.barrier(channel=1)
                self._bcast_pending = True
            elif bucket is not None:
                g = self._bucket_reduce(bucket, flat_grad)
                self._bucket_update(bucket, g, lr, scale)
            if small_params:
                if self.world > 1:
                    for t in small_bufs:
                        dist.all_reduce(t, op=dist.ReduceOp.SUM)
                self._apply_small(small_params, lr, scale)
            done = torch.cuda.Event()
            done.record()
        return done

    @property
    def current_lr(self) -> float:
        if self.constant_lr:
            return self.lr
        # HF Trainer steps the scheduler AFTER the optimizer: optimizer step k (1-based) runs with lambda(k - 1), so the
        # very first update has lr 0 (LambdaLR starts at lambda(0)); step_count is already k inside step()
        return cosine_lr(max(self.step_count - 1, 0), self.total_steps, self.lr, self.warmup_ratio)

    # -------------------------------------------------------------- one train step
    def _micro(self, batch: dict, first: bool, last: bool):
        m = self.model
        dev = m.device
        images = batch["images"]
        if not images.is_cuda:
            images = images.to(dev, non_blocking=True)
        if images.dtype != torch.bfloat16:
            images = images.to(torch.bfloat16)
        plan = m.plan_inputs(batch["input_ids"], batch.get("attention_mask"), batch["labels"], images.shape[0])
        if self.pack_sequences:
            from ..model.interleave_plan import pack_plan
            padded_positions = plan.batch * plan.seq_len
            plan = pack_plan(plan, self.pack_len)
            self.last_padding_saved = padded_positions - plan.batch * plan.seq_len
        self.provider.accumulate = not first
        self.provider.apply_now = last
        res, _ = self.hot.forward_backward(plan, images, self.provider, want_grad=True,
                                           n_save_gu=self.n_save_gu_layers,
                                           train_embed=("model.embed_tokens.weight" in self.opt or
                                                        "model.embed_tokens.weight" in self.big_buckets),
                                           train_projector=("model.mm_projector.0.weight" in self.small_state),
                                           train_llm=self.train_llm,
                                           train_lm_head=("lm_head.weight" in self.big_buckets or "lm_head.weight" in self.small_state),
                                           train_vision_head=("vision_head.0.weight" in self.small_state))
        return res, plan.batch * plan.seq_len

    def step(self, batch) -> dict:
        """batch: input_ids [B,L] (-200 at images), labels, attention_mask (host tensors), images [N,3,S,S] (host,
        pinned, or device) — or, with gradient_accumulation_steps = k > 1, a list of k such micro-batches.
        Returns device scalars (no host sync here); losses are the mean over the micro-batches, as HF logs them."""
        micro = list(batch) if isinstance(batch, (list, tuple)) else [batch]
        if len(micro) != self.accum:
            raise ValueError(f"step() needs {self.accum} micro-batch(es) (gradient_accumulation_steps), got {len(micro)}")
        self.step_count += 1
        tot = None
        tokens = 0
        for i, mb in enumerate(micro):
            res, n_tok = self._micro(mb, first=(i == 0), last=(i == len(micro) - 1))
            tokens += n_tok
            vals = torch.cat([res.loss.reshape(1), res.loss_language.reshape(1), res.loss_image_ar.reshape(1)])
            tot = vals if tot is None else tot + vals
        tot = tot / len(micro)
        if self.max_grad_norm is not None:
            self._clipped_update()
        self._fence_broadcasts()
        torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.last_tokens = tokens
        return dict(loss=tot[0], loss_language=tot[1:2], loss_image_ar=tot[2:3], tokens=tokens)

    def _clipped_update(self):
        """Two-phase update with global-norm clipping (torch.nn.utils.clip_grad_norm_ semantics on the averaged
        gradient): reduce every bucket, sum the squared norms of the slices across ranks, then apply with the factor."""
        pending, self.provider.pending = self.provider.pending, []
        dev = self.model.device
        lr = self.current_lr
        inv = 1.0 / (self.world * self.accum)
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ready)
            sq_sharded = torch.zeros(1, dtype=torch.float32, device=dev)
            sq_repl = torch.zeros(1, dtype=torch.float32, device=dev)

            def sumsq(t, acc):
                if t.dtype == torch.bfloat16 and t.numel() % 8 == 0:
                    ops.sumsq_accum(t.reshape(-1), acc)
                else:
                    acc += t.float().pow(2).sum()

            slices = []
            for bucket, flat, small, bufs in pending:
                if bucket is not None:
                    g = self._bucket_reduce(bucket, flat)
                    sumsq(g, sq_sharded if self.shard_world > 1 else sq_repl)
                    slices.append((bucket, g))
                if self.world > 1:
                    for t in bufs:
                        dist.all_reduce(t, op=dist.ReduceOp.SUM)
                for t in bufs:
                    sumsq(t, sq_repl)
            if self.shard_world > 1:
                dist.all_reduce(sq_sharded, op=dist.ReduceOp.SUM)
            coef = ops.clip_coef((sq_sharded + sq_repl) * (inv * inv), self.max_grad_norm)
            self.last_grad_norm = coef[1:2]
            for bucket, g in slices:
                self._bucket_update(bucket, g, lr, inv, scale_tensor=coef[0:1])
            for _, _, small, _ in pending:
                self._apply_small(small, lr, inv, scale_tensor=coef[0:1])
