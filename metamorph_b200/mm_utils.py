"""Host-side input contract of the hot path — same outputs as the reference helpers, written
for this package:
  tokenizer_image_token      metamorph/mm_utils.py:191-214  prompt with `<image>` marks -> ids with -200 placeholders
  preprocess_multimodal      metamorph/train/train.py:309-332  `<image>` -> `<image_start><image><image_end>` text marks
  DataCollatorForSupervisedDataset  metamorph/train/train.py:1253-1284  pad / truncate / mask / flatten the images
They stay Python (nothing to accelerate) but their integer outputs are the bit-exact contract the GPU path consumes;
pinned by tests/test_input_contract.py against fixtures produced by running the reference's own functions
(oracle/make_golden_inputs.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Sequence

import torch

from .constants import (DEFAULT_IM_END_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IMAGE_TOKEN, IGNORE_INDEX,
                        IMAGE_TOKEN_INDEX)


def tokenizer_image_token(prompt: str, tokenizer, image_token_index: int = IMAGE_TOKEN_INDEX, return_tensors=None):
    """Tokenise the text between the `<image>` marks separately and join the pieces with one placeholder id each.
    Every piece is tokenised with the tokenizer's BOS; the BOS is kept once at the very front and stripped from every
    piece (so a prompt that itself starts with the BOS string yields two BOS ids, as the reference does)."""
    pieces = [tokenizer(text).input_ids for text in prompt.split(DEFAULT_IMAGE_TOKEN)]
    has_bos = bool(pieces) and len(pieces[0]) > 0 and pieces[0][0] == tokenizer.bos_token_id
    skip = 1 if has_bos else 0
    ids: List[int] = [pieces[0][0]] if has_bos else []
    for n, piece in enumerate(pieces):
        if n > 0:
            ids.append(image_token_index)
        ids.extend(piece[skip:])
    if return_tensors is None:
        return ids
    if return_tensors == "pt":
        return torch.tensor(ids, dtype=torch.long)
    raise ValueError(f"Unsupported tensor type: {return_tensors}")


def preprocess_multimodal(sources: Sequence[Sequence[Dict[str, Any]]], data_args) -> Sequence:
    """In place: wrap every `<image>` mark of every turn in the start/end marks when `mm_use_im_start_end` is set
    (MetaMorph leaves the mark where the data put it — no move-to-front as in LLaVA)."""
    if not data_args.is_multimodal:
        return sources
    mark = DEFAULT_IMAGE_TOKEN
    if data_args.mm_use_im_start_end:
        mark = DEFAULT_IM_START_TOKEN + DEFAULT_IMAGE_TOKEN + DEFAULT_IM_END_TOKEN
    for conversation in sources:
        for turn in conversation:
            turn["value"] = turn["value"].replace(DEFAULT_IMAGE_TOKEN, mark)
    return sources


@dataclass
class DataCollatorForSupervisedDataset:
    """Right-pads `input_ids` (pad id) and `labels` (-100) to the batch maximum, cuts both at `model_max_length`,
    derives the attention mask from the pad id and stacks every sample's image list into one flat `[N, 3, S, S]`."""

    tokenizer: Any

    def __call__(self, instances: Sequence[Dict]) -> Dict[str, torch.Tensor]:
        pad = self.tokenizer.pad_token_id
        limit = self.tokenizer.model_max_length
        width = max(int(x["input_ids"].shape[0]) for x in instances)
        ids = torch.full((len(instances), width), pad, dtype=instances[0]["input_ids"].dtype)
        labels = torch.full((len(instances), width), IGNORE_INDEX, dtype=instances[0]["labels"].dtype)
        for r, x in enumerate(instances):
            ids[r, :x["input_ids"].shape[0]] = x["input_ids"]
            labels[r, :x["labels"].shape[0]] = x["labels"]
        ids, labels = ids[:, :limit], labels[:, :limit]
        batch = {"input_ids": ids, "labels": labels, "attention_mask": ids.ne(pad)}
        if "image" in instances[0]:
            batch["images"] = torch.stack([img for x in instances for img in x["image"]])
        return batch
