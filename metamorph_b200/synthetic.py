"""Synthetic, seeded inputs of the benchmark configurations: no dataset, tokenizer
or checkpoint is needed. Shapes follow BASELINE.json `configs`."""
from __future__ import annotations

import torch

from .constants import IGNORE_INDEX, IMAGE_END_TOKEN_ID, IMAGE_START_TOKEN_ID, IMAGE_TOKEN_INDEX

LLAMA3_8B = dict(hidden_size=4096, intermediate_size=14336, num_hidden_layers=32, num_attention_heads=32,
                 num_key_value_heads=8, head_dim=128, vocab_size=128258, rms_norm_eps=1e-5, rope_theta=500000.0,
                 max_position_embeddings=8192)
SIGLIP_SO400M = dict(width=1152, inter=4304, n_layers=27, n_heads=16, image_size=384)


def make_config(llama: dict = None, siglip: dict = None, num_image_tokens: int = 64, max_len: int = 4096):
    from .model import MetaMorphConfig
    l = dict(LLAMA3_8B)
    l.update(llama or {})
    c = MetaMorphConfig(attention_bias=False, tie_word_embeddings=False, **l)
    c.rope_theta = l["rope_theta"]
    c.mm_vision_tower = "siglip/CLIP-ViT-SO400M-14-384"
    c.mm_projector_type = "mlp2x_gelu"
    c.mm_hidden_size = 1152
    c.num_image_tokens = num_image_tokens
    c.image_token_reduction = "interpolation"
    c.normalize_vision = True
    c.freeze_vision = True
    c.vision_head_type = "mlp"
    c.mm_vision_select_layer = -1
    c.mm_use_im_start_end = True
    c.tokenizer_model_max_length = max_len
    c.tokenizer_padding_side = "right"
    s = dict(SIGLIP_SO400M)
    s.update(siglip or {})
    c.mm_vision_tower_dims = s
    c.mm_vision_tower_random_init = True    # synthetic configs never have pretrained SigLIP weights
    return c


def build_model(config, device="cuda"):
    """Random-init model at the configured architecture (HF-style init), tower loaded and frozen."""
    from .model import MetaMorphLlamaForCausalLM
    model = MetaMorphLlamaForCausalLM(config, vision_head="mlp", normalize_vision=True, vision_delay_load=True,
                                      device=device)
    model.get_vision_tower().load_model(device=device, allow_random_init=True)   # synthetic: random-init tower
    for p in model.get_vision_tower().parameters():
        p.requires_grad = False
    return model


def interleaved_sample(seq_len: int, n_prompt_images: int, n_answer_images: int, image_tokens: int,
                       seed: int, vocab_text: int = 128000):
    """One sample whose interleaved length is exactly `seq_len`: BOS, text, [<image_start>,IMG,<image_end>]
    x prompt images, then the answer (labels on) with its images (SURVEY.md §8d layout)."""
    g = torch.Generator().manual_seed(seed)
    n_img = n_prompt_images + n_answer_images
    pre_len = seq_len - n_img * (image_tokens - 1)          # each IMG placeholder expands to image_tokens rows
    text_total = pre_len - 1 - 3 * n_img
    assert text_total > 2 * (n_img + 1), "sequence too short for the requested images"
    n_chunks = n_img + 1
    base = text_total // n_chunks
    chunks = [base] * n_chunks
    chunks[-1] += text_total - base * n_chunks
    ids, labels = [128000], [IGNORE_INDEX]
    for i in range(n_chunks):
        t = torch.randint(0, vocab_text, (chunks[i],), generator=g).tolist()
        answer_side = i >= n_prompt_images
        ids += t
        labels += t if answer_side else [IGNORE_INDEX] * len(t)
        if i < n_img:
            img_answer = i >= n_prompt_images
            trip = [IMAGE_START_TOKEN_ID, IMAGE_TOKEN_INDEX, IMAGE_END_TOKEN_ID]
            ids += trip
            labels += trip if img_answer else [IGNORE_INDEX] * 3
    assert len(ids) == pre_len
    return torch.tensor(ids), torch.tensor(labels)


def train_batch(batch: int, seq_len: int, n_prompt_images: int = 2, n_answer_images: int = 2,
                image_tokens: int = 64, image_size: int = 384, seed: int = 1234, pin: bool = True):
    ids, labs = zip(*[interleaved_sample(seq_len, n_prompt_images, n_answer_images, image_tokens, seed + b)
                      for b in range(batch)])
    input_ids = torch.stack(ids)
    labels = torch.stack(labs)
    mask = torch.ones_like(input_ids, dtype=torch.bool)
    g = torch.Generator().manual_seed(99 + seed)
    n_img = batch * (n_prompt_images + n_answer_images)
    images = torch.randn((n_img, 3, image_size, image_size), generator=g).to(torch.bfloat16)
    if pin and torch.cuda.is_available():
        images = images.pin_memory()
    return dict(input_ids=input_ids, labels=labels, attention_mask=mask, images=images)
