"""On-GPU SigLIP image pre-processing + pinned-memory H2D pipeline.

Replaces, bit for bit, the CPU chain the reference runs in its dataset workers
(`metamorph/train/train.py:1189-1209`: `expand2square(image, int(mean*255))` then
`processor.preprocess(image, return_tensors='pt')['pixel_values'][0]`, processor = the SigLIP image processor of
`multimodal_encoder/siglip_encoder.py:113-121`): Pillow BICUBIC resize to 384x384, x 1/255, normalise 0.5/0.5, CHW.
Raw uint8 pixels travel to the GPU (3 bytes per pixel instead of 12 per output pixel and no CPU resampling); the two
resampling passes and the normalisation run in `csrc/preprocess.cu`.

`SiglipGpuImageProcessor.preprocess(images, return_tensors='pt')['pixel_values']` mirrors the HF processor call the
reference makes; `ImageBatchPipeline` double-buffers host staging so that the copy of batch i+1 overlaps step i.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

from ._lib import MetaMorphB200Error, c_float, c_int, c_void_p, call, lib, ptr, stream_ptr  # noqa: F401

SIGLIP_SIZE = 384
PAD_VALUE = 127   # int(0.5 * 255), train.py:1203


def build_resize_coeffs(in_size: int, out_size: int) -> Tuple[torch.Tensor, int]:
    """Host table (uint8 tensor) of one resampling axis + its tap count, through the C ABI (no GPU needed)."""
    fn = lib().mm_resize_coeff_bytes
    fn.restype = ctypes.c_longlong
    nbytes = int(fn(c_int(in_size), c_int(out_size)))
    if nbytes <= 0:
        raise MetaMorphB200Error(f"bad resize sizes {in_size}->{out_size}")
    host = torch.empty(nbytes, dtype=torch.uint8)
    call("mm_resize_coeff_build", c_void_p(host.data_ptr()), c_int(in_size), c_int(out_size))
    ksize = int(host[:16].view(torch.int32)[2])
    return host, ksize


def normalize_lut() -> torch.Tensor:
    """float32(float64(u) * (1/255)) -> (x - 0.5) / 0.5 in float32: the HF slow processor's arithmetic per byte."""
    x = (np.arange(256, dtype=np.float64) * (1 / 255)).astype(np.float32)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return torch.from_numpy(x)


def _as_hwc_u8(image) -> torch.Tensor:
    """PIL image / numpy array / tensor -> contiguous uint8 [H, W, 3] (host or device) ('convert RGB' semantics)."""
    if isinstance(image, torch.Tensor):
        t = image
    else:
        if hasattr(image, "convert"):            # PIL.Image
            image = np.array(image.convert("RGB"))   # (copy: PIL hands out a read-only buffer)
        t = torch.from_numpy(np.ascontiguousarray(image))
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise MetaMorphB200Error(f"expected an RGB uint8 image [H, W, 3], got {tuple(t.shape)} {t.dtype}")
    return t.contiguous()


class SiglipGpuImageProcessor:
    """Drop-in for the `image_processor` the reference takes from the SigLIP AutoProcessor (only the call it makes)."""

    image_mean = [0.5, 0.5, 0.5]
    image_std = [0.5, 0.5, 0.5]
    crop_size = {"height": SIGLIP_SIZE, "width": SIGLIP_SIZE}
    size = {"height": SIGLIP_SIZE, "width": SIGLIP_SIZE}

    def __init__(self, device="cuda", pad_to_square: bool = True, out_dtype=torch.float32, size: int = SIGLIP_SIZE):
        self.device = torch.device(device)
        self.pad = pad_to_square
        self.out_dtype = out_dtype
        self.out_size = size
        self._coeffs: Dict[int, Tuple[torch.Tensor, int]] = {}
        self._lut = None
        self._tmp = None

    def _coeff(self, in_size: int):
        c = self._coeffs.get(in_size)
        if c is None:
            host, ksize = build_resize_coeffs(in_size, self.out_size)
            c = (host.to(self.device), ksize)
            self._coeffs[in_size] = c
        return c

    def _one(self, img: torch.Tensor, out: torch.Tensor):
        H, W = int(img.shape[0]), int(img.shape[1])
        side_w = max(H, W) if self.pad else W
        side_h = side_w if self.pad else H
        cx, kx = self._coeff(side_w)
        cy, ky = self._coeff(side_h)
        if self._lut is None:
            self._lut = normalize_lut().to(self.device)
        need = side_h * self.out_size * 3
        if self._tmp is None or self._tmp.numel() < need:
            self._tmp = torch.empty(need, dtype=torch.uint8, device=self.device)
        call("mm_siglip_preprocess", ptr(img), c_int(H), c_int(W), c_int(1 if self.pad else 0), c_int(PAD_VALUE),
             ptr(cx), ptr(cy), c_int(kx), c_int(ky), c_int(self.out_size), ptr(self._lut), ptr(self._tmp), ptr(out),
             c_int(1 if out.dtype == torch.bfloat16 else 0), stream_ptr())

    def preprocess(self, images, return_tensors="pt", out=None, **_):
        """images: one image or a list (PIL / numpy HWC uint8 / uint8 tensors, host or device).
        Returns {'pixel_values': [N, 3, S, S]} on the device, like `processor.preprocess(...)`; `out` lets the caller
        supply the destination (no allocation on the hot path)."""
        if not isinstance(images, (list, tuple)):
            images = [images]
        if out is None:
            out = torch.empty((len(images), 3, self.out_size, self.out_size), dtype=self.out_dtype, device=self.device)
        for i, im in enumerate(images):
            t = _as_hwc_u8(im)
            if not t.is_cuda:
                t = t.to(self.device, non_blocking=True)
            self._one(t, out[i])
        return {"pixel_values": out}

    __call__ = preprocess


class ImageBatchPipeline:
    """Two pinned staging buffers + a copy stream: `submit(images)` packs the raw uint8 pixels of a batch into pinned
    memory and starts the H2D copy + GPU pre-processing on a side stream; `result()` hands the previous batch to the
    compute stream. Submitting batch i+1 before running step i overlaps its copy and resampling with the step."""

    def __init__(self, processor: SiglipGpuImageProcessor, max_bytes: int = 64 << 20):
        self.p = processor
        self.stream = torch.cuda.Stream(device=processor.device)
        self.host = [torch.empty(max_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.dev = [torch.empty(max_bytes, dtype=torch.uint8, device=processor.device) for _ in range(2)]
        self.out = [None, None]          # per-slot outputs, kept across batches: a training process that fills HBM must
        self.slot = 0                    # not go back to the allocator every step
        self.pending = None
        self.h2d_bytes = 0

    def submit(self, images: Sequence):
        imgs = [_as_hwc_u8(im) for im in images]
        total = sum(int(t.numel()) for t in imgs)
        host, dev = self.host[self.slot], self.dev[self.slot]
        if total > host.numel():
            raise MetaMorphB200Error(f"batch of {total} raw bytes exceeds the staging buffer ({host.numel()})")
        off, views = 0, []
        for t in imgs:
            n = int(t.numel())
            host[off:off + n].copy_(t.reshape(-1))
            views.append((off, tuple(t.shape)))
            off += n
        self.h2d_bytes = total
        self.stream.wait_stream(torch.cuda.current_stream())     # the slot's previous consumer has been enqueued
        with torch.cuda.stream(self.stream):
            dev[:total].copy_(host[:total], non_blocking=True)
            buf = self.out[self.slot]
            if buf is None or buf.shape[0] < len(imgs):
                buf = torch.empty((len(imgs), 3, self.p.out_size, self.p.out_size), dtype=self.p.out_dtype,
                                  device=self.p.device)
                self.out[self.slot] = buf
            out = self.p.preprocess([dev[o:o + h * w * 3].view(h, w, 3) for o, (h, w, _) in views],
                                    out=buf[:len(imgs)])["pixel_values"]
            done = torch.cuda.Event()
            done.record()
        self.pending = (out, done)
        self.slot ^= 1
        return self

    def result(self) -> torch.Tensor:
        out, done = self.pending
        torch.cuda.current_stream().wait_event(done)
        return out          # valid until the same slot is submitted again (two batches later)
