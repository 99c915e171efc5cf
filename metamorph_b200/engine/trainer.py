"""Train step driver: hot path forward/backward + fused AdamW + data-parallel gradient collective.

Replaces, for this path, HF Trainer + Accelerate + DeepSpeed ZeRO (scripts/zero2.json):
  * pure data parallel, one process per GPU (torchrun), NCCL over NVLink/NVSwitch;
  * "optimizer in the backward sweep": as soon as a layer's wgrad GEMMs retire, its gradient
    bucket (one flat bf16 buffer, ~436 MB for LLaMA-3-8B) is all-reduced on a dedicated comm stream
    and the fused AdamW kernel updates that layer's fp32 master weights / moments / bf16 copy, while
    the main stream continues with the next layer's backward. Only two layer-sized gradient buckets
    exist, so the 16 GB of full-model gradients are never resident (180 GB HBM budget);
  * optional global-norm clipping (`max_grad_norm`) switches to a two-phase step with resident
    gradients (the reference scripts pass no max_grad_norm).
Optimizer = torch.optim.AdamW semantics (reference: --optim adamw_torch, train.py:82), cosine LR with
3 % warm-up (scripts/*.sh: lr_scheduler_type cosine, warmup_ratio 0.03), weight decay 0.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

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


class _OptState:
    __slots__ = ("p16", "p32", "m", "v")

    def __init__(self, p: torch.Tensor):
        self.p16 = p.data
        self.p32 = p.data.float()
        self.m = torch.zeros_like(self.p32)
        self.v = torch.zeros_like(self.p32)


class FusedGradProvider(GradProvider):
    """Gradient buckets + immediate (all-reduce ->) AdamW as each bucket completes."""

    def __init__(self, model, engine: "TrainEngine"):
        super().__init__(model)
        self.e = engine
        l0 = model.model.layers[0]
        self.shapes = [l0.self_attn.qkv_proj.weight.shape, l0.self_attn.o_proj.weight.shape,
                       l0.mlp.gate_up_proj.weight.shape, l0.mlp.down_proj.weight.shape]
        self.sizes = [s[0] * s[1] for s in self.shapes]
        dev = l0.self_attn.qkv_proj.weight.device
        H = l0.input_layernorm.weight.shape[0]
        n = sum(self.sizes)
        self.sets = []
        for _ in range(2):
            flat = torch.zeros(n, dtype=torch.bfloat16, device=dev)
            ln = torch.zeros(2 * H, dtype=torch.float32, device=dev)
            self.sets.append(dict(flat=flat, ln=ln, free=None))
        self.H = H

    def layer(self, i: int) -> LayerGrads:
        s = self.sets[i % 2]
        if s["free"] is not None:                      # bucket still being reduced / applied
            torch.cuda.current_stream().wait_event(s["free"])
            s["free"] = None
        views, off = [], 0
        for shp, sz in zip(self.shapes, self.sizes):
            views.append(s["flat"][off:off + sz].view(shp))
            off += sz
        s["ln"].zero_()
        return LayerGrads(views[0], views[1], views[2], views[3], s["ln"][:self.H], s["ln"][self.H:])

    def layer_done(self, i: int, g: LayerGrads):
        s = self.sets[i % 2]
        l = self.model.model.layers[i]
        params = [(l.self_attn.qkv_proj.weight, g.wqkv), (l.self_attn.o_proj.weight, g.wo),
                  (l.mlp.gate_up_proj.weight, g.wgu), (l.mlp.down_proj.weight, g.wd),
                  (l.input_layernorm.weight, g.ln1), (l.post_attention_layernorm.weight, g.ln2)]
        s["free"] = self.e.reduce_and_apply(params, [s["flat"], s["ln"]])

    def group_done(self, group: str):
        m = self.model
        names = {
            "heads": ["lm_head.weight", "vision_head.0.weight", "vision_head.0.bias", "vision_head.2.weight",
                      "vision_head.2.bias"],
            "final_norm": ["model.norm.weight"],
            "embed": ["model.embed_tokens.weight"],
            "projector": ["model.mm_projector.0.weight", "model.mm_projector.0.bias",
                          "model.mm_projector.2.weight", "model.mm_projector.2.bias"],
        }[group]
        named = self.e.named_params
        params = [(named[n], self.buffers[n]) for n in names if n in self.buffers and n in self.e.opt]
        if params:
            self.e.reduce_and_apply(params, [b for _, b in params])


class TrainEngine:
    def __init__(self, model, lr: float = 6.93e-5, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_grad_norm: Optional[float] = None, total_steps: int = 1000,
                 warmup_ratio: float = 0.03, constant_lr: bool = False, n_save_gu_layers: int = 0,
                 pack_sequences: bool = False, pack_len: Optional[int] = None):
        self.model = model
        self.hot = HotPath(model)
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        self.total_steps, self.warmup_ratio, self.constant_lr = total_steps, warmup_ratio, constant_lr
        self.n_save_gu_layers = n_save_gu_layers
        # lay the samples of a batch end to end (block-diagonal attention) instead of padding them
        self.pack_sequences, self.pack_len = pack_sequences, pack_len
        self.step_count = 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.named_params: Dict[str, torch.nn.Parameter] = dict(model.named_parameters())
        self.opt: Dict[str, _OptState] = {}
        for n, p in self.named_params.items():
            if p.requires_grad and "vision_tower" not in n and "vision_proj" not in n:
                self.opt[n] = _OptState(p)
        self._state_by_ptr = {st.p16.data_ptr(): st for st in self.opt.values()}
        # side stream: gradient all-reduce (N>1) and the HBM-bound fused AdamW run here, concurrently with
        # the tensor-core-bound backward GEMMs of the next layers on the main stream
        self.comm_stream = torch.cuda.Stream()
        self.provider = FusedGradProvider(model, self) if max_grad_norm is None else GradProvider(model)
        self.kernel_launch_estimate = 0

    # -------------------------------------------------------------- optimizer plumbing
    def _apply(self, params, lr: float, grad_scale: float, scale_tensor=None):
        b1, b2 = self.betas
        for p, g in params:
            st = self._state_by_ptr.get(p.data.data_ptr())
            if st is None:
                continue
            ops.adamw_step_(st.p16.view(-1), st.p32.view(-1), st.m.view(-1), st.v.view(-1), g.reshape(-1),
                            lr=lr, beta1=b1, beta2=b2, eps=self.eps, wd=self.wd, step=self.step_count,
                            grad_scale=grad_scale, grad_scale_tensor=scale_tensor)

    def reduce_and_apply(self, params, flat_buffers):
        """(all-reduce the bucket ->) fused AdamW. Returns an event marking bucket reuse safety."""
        lr = self.current_lr
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ready)
            if self.world > 1:
                for b in flat_buffers:
                    dist.all_reduce(b, op=dist.ReduceOp.SUM)
            self._apply(params, lr, 1.0 / self.world)
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
    def step(self, batch: dict) -> dict:
        """batch: input_ids [B,L] (-200 at images), labels, attention_mask (host tensors), images
        [N,3,S,S] (host, pinned, or device). Returns device scalars (no host sync here)."""
        m = self.model
        self.step_count += 1
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
        res, _ = self.hot.forward_backward(plan, images, self.provider, want_grad=True,
                                           n_save_gu=self.n_save_gu_layers,
                                           train_embed=("model.embed_tokens.weight" in self.opt),
                                           train_projector=("model.mm_projector.0.weight" in self.opt))
        if self.max_grad_norm is not None:
            self._clipped_update()
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.last_tokens = plan.batch * plan.seq_len
        return dict(loss=res.loss, loss_language=res.loss_language, loss_image_ar=res.loss_image_ar,
                    tokens=self.last_tokens)

    def _clipped_update(self):
        """Two-phase update with global-norm clipping (torch.nn.utils.clip_grad_norm_ semantics)."""
        bufs = self.provider.buffers
        sumsq = torch.zeros(1, dtype=torch.float32, device=self.model.device)
        items = [(self.named_params[n], g) for n, g in bufs.items() if n in self.opt]
        if self.world > 1:
            for _, g in items:
                dist.all_reduce(g, op=dist.ReduceOp.SUM)
        inv = 1.0 / self.world
        for _, g in items:
            if g.dtype == torch.bfloat16 and g.numel() % 8 == 0:
                ops.sumsq_accum(g.view(-1), sumsq)
            else:
                sumsq += g.float().pow(2).sum()
        coef = ops.clip_coef(sumsq * (inv * inv), self.max_grad_norm)
        self.last_grad_norm = coef[1:2]
        self._apply(items, self.current_lr, inv, scale_tensor=coef[0:1])
