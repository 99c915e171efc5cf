"""Continuous-batching decode server: the per-sequence text / image state machine of
`greedy_decode` (metamorph_llama.py:502-597) for a stream of requests instead of one batch.

The reference serves one request at a time (inference/demo.py:117-179: `generate` -> visualise the returned image
embeddings). Here up to `max_slots` (<= 32) requests share every weight-streaming decode step:
  * each batch slot owns a KV-cache region, its mode / counters / position on the device and its own output limit;
  * a queued request is admitted into a free slot BETWEEN steps: its first P-1 prompt positions are prefilled with the
    full-sequence kernels (tcgen05 GEMMs + flash attention) straight into the slot's cache region, and the last prompt
    position is fed through the ordinary decode step — no special first-token path;
  * the step itself (decoder stack + heads + argmax + state machine + next-input gather) is ONE captured CUDA graph
    replayed for all slots, exactly the kernels of DecodeEngine; idle / finished slots are frozen by the device state;
  * the host polls the tiny state arrays every `poll_every` steps, hands out finished requests, streams the new text
    ids / visual embeddings of running ones (`run()` yields them), and refills the freed slots.
Every request's output equals what `greedy_decode` produces for it alone (tests/test_decode_gpu.py).
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass, field
from typing import Deque, Dict, List, Optional, Tuple

import torch

from .. import ops
from ..constants import EOS_TOKEN_IDS, IMAGE_END_TOKEN_ID, IMAGE_START_TOKEN_ID
from .decode_step import decode_heads, decoder_stack_step
from .llama import StackContext


@dataclass
class _Request:
    rid: int
    embeds: torch.Tensor                 # [P, H] bf16 (device)
    max_new_tokens: int
    forced: Optional[torch.Tensor]       # [n] int32 (host) or None
    slot: int = -1
    sent_ids: int = 0
    sent_img: int = 0
    ids: List[torch.Tensor] = field(default_factory=list)
    img: List[torch.Tensor] = field(default_factory=list)


class ContinuousBatcher:
    def __init__(self, model, max_slots: int = 8, max_context: int = 2048, max_new_tokens: int = 1024,
                 poll_every: int = 8, use_cuda_graph: bool = True, start_image_token_id: int = IMAGE_START_TOKEN_ID,
                 end_image_token_id: int = IMAGE_END_TOKEN_ID, eos_token_id=EOS_TOKEN_IDS):
        assert 1 <= max_slots <= 32, "the weight-streaming step serves at most 32 sequences"
        self.m = model
        self.inner = model.get_model()
        self.stack = model.stack
        d = self.stack.dims
        self.d = d
        self.B, self.Tmax, self.cap = max_slots, max_context, max_new_tokens
        self.poll_every = poll_every
        self.dev = self.inner.embed_tokens.weight.device
        dev, B = self.dev, max_slots
        self.L = len(self.inner.layers)
        self.layers = [l.weights() for l in self.inner.layers]
        self.ntok = model.get_vision_tower().image_token_len if model.get_vision_tower() is not None else 0
        self.C = model.vision_head.fc2.out_features
        eos = list(eos_token_id) if isinstance(eos_token_id, (list, tuple)) else [eos_token_id]
        if len(eos) > 2:
            raise NotImplementedError("at most two EOS ids (the reference's default is [128001, 128009])")
        self.eos0 = eos[0] if eos else -1                       # an empty list never matches
        self.eos1 = eos[1] if len(eos) > 1 else self.eos0
        self.start_id, self.end_id = start_image_token_id, end_image_token_id
        self.stack.ensure_positions(max_context + 1)
        self.kc = torch.zeros((self.L, B, d.n_kv_heads, max_context, d.head_dim), dtype=torch.bfloat16, device=dev)
        self.vc = torch.zeros_like(self.kc)
        self.st = {k: torch.zeros(B, dtype=torch.int32, device=dev) for k in
                   ("in_image_mode", "total_image_tokens", "total_output", "n_ids", "n_img", "append_kind",
                    "next_token")}
        self.st["finished"] = torch.ones(B, dtype=torch.int32, device=dev)          # idle slots are frozen
        self.st["pos"] = torch.ones(B, dtype=torch.int32, device=dev)
        self.max_ids = max_new_tokens + 2
        self.max_img = max(1, ((max_new_tokens + 1) // max(self.ntok, 1) + 1) * max(self.ntok, 1))
        self.st["ids_out"] = torch.full((B, self.max_ids), -1, dtype=torch.int32, device=dev)
        self.img_out = torch.zeros((B, self.max_img, self.C), dtype=torch.bfloat16, device=dev)
        self.forced = torch.full((B, max_new_tokens + 2), -1, dtype=torch.int32, device=dev)   # -1 = free running
        self.max_new_slot = torch.zeros(B, dtype=torch.int32, device=dev)
        self.xin = torch.zeros((B, d.hidden), dtype=torch.bfloat16, device=dev)
        V = model.lm_head.weight.shape[0]
        self.V = V
        self.logits = torch.empty((B, (V + 7) // 8 * 8), dtype=torch.float32, device=dev)
        self.queue: Deque[_Request] = deque()
        self.slots: List[Optional[_Request]] = [None] * B
        self.next_rid = 0
        self.steps_run = 0
        self.graph = None
        self.use_cuda_graph = use_cuda_graph
        self._warm = False

    # ------------------------------------------------------------------ one device step for all slots
    def _step_body(self):
        st = self.st
        x = decoder_stack_step(self.layers, self.xin, self.kc, self.vc, st["pos"] - 1, self.stack)
        tok, pred_z, prediction = decode_heads(self.m, x, st["in_image_mode"], self.logits, self.V)
        ops.decode_state_step_slots(st, tok, self.forced, self.max_new_slot, self.B, self.ntok, self.start_id,
                                    self.end_id, self.eos0, self.eos1, pred_z, self.img_out)
        ops.decode_next_input(st["append_kind"], st["next_token"], self.inner.embed_tokens.weight.data, prediction,
                              self.xin)

    def _device_step(self):
        if self.graph is not None:
            self.graph.replay()
        elif self.use_cuda_graph and self._warm:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g):
                    self._step_body()
                self.graph = g                         # (the capture itself does not execute the step)
                self.graph.replay()
            except Exception:  # noqa: BLE001 - capture unsupported: stay on stream launches
                self.use_cuda_graph = False
                torch.cuda.synchronize()
                self._step_body()
        else:
            self._step_body()                          # first step eager: sets kernel attributes
            self._warm = True
        self.steps_run += 1

    # ------------------------------------------------------------------ requests
    @torch.no_grad()
    def submit(self, inputs_embeds: torch.Tensor, max_new_tokens: Optional[int] = None,
               forced_tokens: Optional[torch.Tensor] = None) -> int:
        """inputs_embeds: [P, H] or [1, P, H] prompt embeddings (text + projected image rows, as `generate` builds
        them). Returns the request id."""
        e = inputs_embeds.reshape(-1, inputs_embeds.shape[-1]).to(self.dev, dtype=torch.bfloat16).contiguous()
        n_new = self.cap if max_new_tokens is None else int(max_new_tokens)
        if n_new > self.cap:
            raise ValueError(f"max_new_tokens {n_new} exceeds the server's limit {self.cap}")
        if e.shape[0] < 1 or e.shape[0] + n_new + 2 > self.Tmax:
            raise ValueError(f"prompt of {e.shape[0]} positions + {n_new} new ones does not fit max_context {self.Tmax}")
        f = None if forced_tokens is None else forced_tokens.reshape(-1).to(torch.int32).cpu()
        rid = self.next_rid
        self.next_rid += 1
        self.queue.append(_Request(rid, e, n_new, f))
        return rid

    @torch.no_grad()
    def _admit(self, req: _Request, b: int):
        d = self.d
        Hq, Hkv, dh = d.n_heads, d.n_kv_heads, d.head_dim
        P = req.embeds.shape[0]
        if P > 1:   # prefill positions 0..P-2 into the slot's cache region
            pos = torch.arange(P - 1, dtype=torch.int32, device=self.dev)
            ctx = StackContext(B=1, T=P - 1, pos=pos, seqlens=None)
            x = req.embeds[:P - 1].contiguous()
            for i, w in enumerate(self.layers):
                x = self.stack.layer_forward(w, x, ctx, save=True, save_gu=False)
                s = ctx.saved.pop()
                ops.kv_prefill(s.qkv, self.kc[i, b:b + 1], self.vc[i, b:b + 1], 1, P - 1, Hq, Hkv, dh)
                del s
        self.xin[b].copy_(req.embeds[P - 1])
        row = torch.full((self.forced.shape[1],), -1, dtype=torch.int32)
        if req.forced is not None:
            n = min(row.numel(), req.forced.numel())
            row[:n] = req.forced[:n]
        self.forced[b].copy_(row.to(self.dev, non_blocking=True))
        for k in ("in_image_mode", "total_image_tokens", "total_output", "n_ids", "n_img", "finished"):
            self.st[k][b] = 0
        self.st["append_kind"][b] = -1
        self.st["pos"][b] = P                       # the fed token sits at position P-1
        self.max_new_slot[b] = req.max_new_tokens
        req.slot, req.sent_ids, req.sent_img = b, 0, 0
        self.slots[b] = req

    def _poll(self) -> Tuple[List[Tuple[int, str, torch.Tensor]], List[int]]:
        """One host sync: new ids / visual embeddings of every running request + the requests that finished."""
        snap = torch.stack([self.st["finished"], self.st["n_ids"], self.st["n_img"]]).cpu()
        events, done = [], []
        for b, req in enumerate(self.slots):
            if req is None:
                continue
            fin, n_ids, n_img = int(snap[0, b]), int(snap[1, b]), int(snap[2, b])
            if n_ids > req.sent_ids:
                chunk = self.st["ids_out"][b, req.sent_ids:n_ids].clone()
                req.ids.append(chunk)
                events.append((req.rid, "ids", chunk))
                req.sent_ids = n_ids
            if n_img > req.sent_img:
                chunk = self.img_out[b, req.sent_img:n_img].clone()
                req.img.append(chunk)
                events.append((req.rid, "image_embeds", chunk))
                req.sent_img = n_img
            if fin:
                done.append(b)
        return events, done

    @torch.no_grad()
    def run(self, max_steps: Optional[int] = None):
        """Generator: serves the queue; yields (rid, 'ids' | 'image_embeds', tensor) as outputs appear and
        (rid, 'done', (ids, image_embeds)) when a request completes. Returns when queue and slots are empty."""
        steps = 0
        while self.queue or any(s is not None for s in self.slots):
            for b in range(self.B):
                if self.slots[b] is None and self.queue:
                    self._admit(self.queue.popleft(), b)
            for _ in range(self.poll_every):
                self._device_step()
                steps += 1
            events, done = self._poll()
            for ev in events:
                yield ev
            for b in done:
                req = self.slots[b]
                self.slots[b] = None
                ids = torch.cat(req.ids) if req.ids else torch.empty(0, dtype=torch.int32, device=self.dev)
                img = torch.cat(req.img) if req.img else torch.empty((0, self.C), dtype=torch.bfloat16, device=self.dev)
                yield (req.rid, "done", (ids, img))
            if max_steps is not None and steps >= max_steps:
                return

    def run_until_idle(self) -> Dict[int, Tuple[torch.Tensor, torch.Tensor]]:
        return {rid: payload for rid, kind, payload in self.run() if kind == "done"}
