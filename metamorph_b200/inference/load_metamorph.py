"""`load_metamorph_model` (inference/load_metamorph.py:25-65)."""
import os

import torch

from ..model.builder import load_pretrained_model


def get_model_name_from_path(model_path):
    """metamorph/mm_utils.py:218-224."""
    model_path = model_path.strip("/")
    parts = model_path.split("/")
    if parts[-1].startswith("checkpoint-"):
        return parts[-2] + "_" + parts[-1]
    return parts[-1]


def load_metamorph_model(model_path, model_base=None, device="cuda", dtype=torch.float16):
    model_path = os.path.expanduser(model_path)
    model_name = get_model_name_from_path(model_path)
    tokenizer, model, image_processor, context_len = load_pretrained_model(
        model_path, model_base, model_name, device=device, torch_dtype=dtype)
    model.eval()
    return tokenizer, model, image_processor, context_len


load_metamorph = load_metamorph_model  # north-star spelling
