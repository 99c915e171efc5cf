"""Host-side index logic of `prepare_inputs_labels_for_multimodal` (metamorph_arch.py:245-425).

The reference interleaves text and image embeddings with a per-sample Python loop over device
tensors (>= 3 host syncs per sample). Here the same decisions are taken ONCE on the host from the
integer inputs only and returned as an `InterleavePlan`: an int32 row map consumed by the CUDA
gather kernel (csrc/interleave.cu) plus the label / image-position / mask / position-id tensors.
All integer outputs are bit-exact with the reference, including its quirks:
  * a sample without <image> still consumes one (dummy) image slot        (:275-284)
  * `answer_image` = label of the token right before the placeholder == 128256   (:317)
  * an image that would overflow tokenizer_model_max_length stops the sample  (:324-326)
  * truncation to tokenizer_model_max_length after interleaving             (:355-358)
  * right / left padding                                                     (:373-397)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from ..constants import IGNORE_INDEX, IMAGE_START_TOKEN_ID, IMAGE_TOKEN_INDEX

ROW_PAD = -1  # row_map value for padding rows; image rows are encoded as -(2 + flat_image_row)


@dataclass
class InterleavePlan:
    row_map: torch.Tensor          # int32 [B, T]: >=0 token id, -1 pad, <=-2 image row -(2+r)
    labels: torch.Tensor           # int64 [B, T]
    image_positions: torch.Tensor  # int64 [B, T]  (1 on answer-image rows)
    attention_mask: torch.Tensor   # bool  [B, T]
    position_ids: torch.Tensor     # int64 [B, T]
    seqlens: torch.Tensor          # int32 [B]   valid length per sample
    target_image_idx: List[int]    # images (flat index) whose tower features are regression targets
    image_placeholder: List[int]   # complement bookkeeping of the reference (:268, :415-423)
    padding_side: str
    # sequence packing: per packed row, the (start, length) of every sample placed in it; None = one
    # sample per row. Attention is block-diagonal over these segments; everything else is per token.
    segments: Optional[List[List[tuple]]] = None

    @property
    def batch(self) -> int:
        return self.row_map.shape[0]

    @property
    def seq_len(self) -> int:
        return self.row_map.shape[1]


def build_interleave_plan(input_ids, attention_mask, labels, num_images: int, image_len: int,
                          tokenizer_model_max_length: Optional[int], padding_side: str = "right",
                          start_image_token_id: int = IMAGE_START_TOKEN_ID) -> InterleavePlan:
    ids_all = np.asarray(input_ids.detach().cpu().numpy() if torch.is_tensor(input_ids) else input_ids)
    B, L = ids_all.shape
    if attention_mask is None:
        mask_all = np.ones((B, L), dtype=bool)
    else:
        mask_all = np.asarray(attention_mask.detach().cpu().numpy() if torch.is_tensor(attention_mask)
                              else attention_mask).astype(bool)
    if labels is None:
        labels_all = np.full((B, L), IGNORE_INDEX, dtype=np.int64)
    else:
        labels_all = np.asarray(labels.detach().cpu().numpy() if torch.is_tensor(labels) else labels)

    rows_out, labels_out, impos_out = [], [], []
    placeholder: List[int] = []
    cur_image_idx = 0
    for b in range(B):
        ids = ids_all[b][mask_all[b]].astype(np.int64)
        labs = labels_all[b][mask_all[b]].astype(np.int64)
        img_pos = np.nonzero(ids == IMAGE_TOKEN_INDEX)[0]
        n_img = len(img_pos)
        if n_img == 0:
            if cur_image_idx >= num_images:
                raise IndexError(f"index {cur_image_idx} is out of bounds for dimension 0 with size {num_images}")
            placeholder.append(cur_image_idx)
            rows_out.append(ids.copy())
            labels_out.append(labs.copy())
            impos_out.append(np.zeros_like(labs))
            cur_image_idx += 1
            continue
        bounds = [-1] + img_pos.tolist() + [len(ids)]
        r_parts, l_parts, p_parts = [], [], []
        cur_len = 0
        need_to_stop = False
        for i in range(n_img + 1):
            c_ids = ids[bounds[i] + 1:bounds[i + 1]]
            c_labs = labs[bounds[i] + 1:bounds[i + 1]]
            if not need_to_stop:
                r_parts.append(c_ids)
                l_parts.append(c_labs)
                p_parts.append(np.zeros_like(c_labs))
                cur_len += len(c_ids)
            if i < n_img:
                if len(c_labs) == 0:
                    raise IndexError("index -1 is out of bounds for dimension 0 with size 0")
                answer_image = bool(c_labs[-1] == start_image_token_id)
                if cur_image_idx >= num_images:
                    raise IndexError(f"index {cur_image_idx} is out of bounds for dimension 0 with size {num_images}")
                if tokenizer_model_max_length is None:
                    raise TypeError("'>' not supported between instances of 'int' and 'NoneType'")
                if cur_len + image_len > tokenizer_model_max_length:
                    need_to_stop = True
                    placeholder.append(cur_image_idx)
                else:
                    base = cur_image_idx * image_len
                    r_parts.append(-(2 + base + np.arange(image_len, dtype=np.int64)))
                    l_parts.append(np.full((image_len,), IGNORE_INDEX, dtype=np.int64))
                    if answer_image:
                        p_parts.append(np.ones((image_len,), dtype=np.int64))
                    else:
                        placeholder.append(cur_image_idx)
                        p_parts.append(np.zeros((image_len,), dtype=np.int64))
                    cur_len += image_len
                cur_image_idx += 1
        rows_out.append(np.concatenate(r_parts))
        labels_out.append(np.concatenate(l_parts))
        impos_out.append(np.concatenate(p_parts))

    if tokenizer_model_max_length is not None:
        rows_out = [x[:tokenizer_model_max_length] for x in rows_out]
        labels_out = [x[:tokenizer_model_max_length] for x in labels_out]
        impos_out = [x[:tokenizer_model_max_length] for x in impos_out]

    T = max(len(x) for x in rows_out)
    row_map = np.full((B, T), ROW_PAD, dtype=np.int32)
    new_labels = np.full((B, T), IGNORE_INDEX, dtype=np.int64)
    new_impos = np.zeros((B, T), dtype=np.int64)
    new_mask = np.zeros((B, T), dtype=bool)
    pos_ids = np.zeros((B, T), dtype=np.int64)
    seqlens = np.zeros((B,), dtype=np.int32)
    for b in range(B):
        n = len(rows_out[b])
        seqlens[b] = n
        if n == 0:
            continue
        sl = slice(T - n, T) if padding_side == "left" else slice(0, n)
        row_map[b, sl] = rows_out[b]
        new_labels[b, sl] = labels_out[b]
        new_impos[b, sl] = impos_out[b]
        new_mask[b, sl] = True
        pos_ids[b, sl] = np.arange(n)

    ph = set(placeholder)
    targets = [i for i in range(num_images) if i not in ph]
    return InterleavePlan(torch.from_numpy(row_map), torch.from_numpy(new_labels),
                          torch.from_numpy(new_impos), torch.from_numpy(new_mask),
                          torch.from_numpy(pos_ids), torch.from_numpy(seqlens), targets,
                          placeholder, padding_side)


def pack_plan(plan: InterleavePlan, pack_len: Optional[int] = None) -> InterleavePlan:
    """Sequence packing: the samples of a right-padded plan are laid end to end into rows of
    `pack_len` positions (next-fit in the original sample order, so the row-major order of the answer-image rows still
    matches the order of the regression targets, metamorph_arch.py:415-423). The reference pads every sample to the
    batch maximum (metamorph_arch.py:361-399); packing feeds the same tokens, labels and position ids through
    block-diagonal causal attention, so the losses and gradients are those of the padded batch while the padding rows
    (and their FLOPs) disappear. The label at each segment start is forced to IGNORE_INDEX: in the padded layout that
    label is dropped by the shift (metamorph_llama.py:404-405), here it would become the target of the previous
    sample's last token."""
    if plan.padding_side != "right":
        raise NotImplementedError("packing expects a right-padded plan")
    B, T = plan.batch, plan.seq_len
    Tp = int(pack_len or T)
    lens = [int(x) for x in plan.seqlens.tolist()]
    if max(lens) > Tp:
        raise ValueError(f"a sample of {max(lens)} positions does not fit pack_len {Tp}")
    rows: List[List[int]] = [[]]
    used = 0
    for b, n in enumerate(lens):
        if n == 0:
            continue
        if used + n > Tp and rows[-1]:
            rows.append([])
            used = 0
        rows[-1].append(b)
        used += n
    R = len(rows)
    row_map = np.full((R, Tp), ROW_PAD, dtype=np.int32)
    labels = np.full((R, Tp), IGNORE_INDEX, dtype=np.int64)
    ipos = np.zeros((R, Tp), dtype=np.int64)
    mask = np.zeros((R, Tp), dtype=bool)
    pos = np.zeros((R, Tp), dtype=np.int64)
    seqlens = np.zeros((R,), dtype=np.int32)
    segments: List[List[tuple]] = []
    src = dict(row_map=plan.row_map.numpy(), labels=plan.labels.numpy(), ipos=plan.image_positions.numpy(),
               pos=plan.position_ids.numpy())
    for r, members in enumerate(rows):
        off, segs = 0, []
        for b in members:
            n = lens[b]
            row_map[r, off:off + n] = src["row_map"][b, :n]
            labels[r, off:off + n] = src["labels"][b, :n]
            labels[r, off] = IGNORE_INDEX
            ipos[r, off:off + n] = src["ipos"][b, :n]
            pos[r, off:off + n] = src["pos"][b, :n]
            mask[r, off:off + n] = True
            segs.append((off, n))
            off += n
        seqlens[r] = off
        segments.append(segs)
    return InterleavePlan(torch.from_numpy(row_map), torch.from_numpy(labels), torch.from_numpy(ipos),
                          torch.from_numpy(mask), torch.from_numpy(pos), torch.from_numpy(seqlens),
                          list(plan.target_image_idx), list(plan.image_placeholder), "right", segments)
