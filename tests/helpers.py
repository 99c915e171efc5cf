"""Shared test helpers (tests only): build the product model at the oracle's tiny dims."""
import torch


def product_config(cfg, num_image_tokens=None, max_len=None):
    from metamorph_b200.model import MetaMorphConfig
    c = MetaMorphConfig(hidden_size=cfg["hidden"], intermediate_size=cfg["inter"],
                        num_hidden_layers=cfg["layers"], num_attention_heads=cfg["heads"],
                        num_key_value_heads=cfg["kv_heads"], head_dim=cfg["head_dim"], vocab_size=cfg["vocab"],
                        rms_norm_eps=cfg["rms_eps"], rope_theta=cfg["rope_theta"], max_position_embeddings=8192,
                        attention_bias=False, tie_word_embeddings=False)
    c.rope_theta = cfg["rope_theta"]
    c.mm_vision_tower = "siglip/CLIP-ViT-SO400M-14-384"
    c.mm_projector_type = "mlp2x_gelu"
    c.mm_hidden_size = 1152
    c.num_image_tokens = num_image_tokens or cfg["image_tokens"]
    c.image_token_reduction = "interpolation"
    c.normalize_vision = True
    c.freeze_vision = True
    c.vision_head_type = "mlp"
    c.mm_vision_select_layer = -1
    c.tokenizer_model_max_length = max_len or cfg["max_len"]
    c.tokenizer_padding_side = "right"
    c.mm_vision_tower_dims = dict(width=cfg["siglip_width"], inter=cfg["siglip_inter"],
                                  n_layers=cfg["siglip_layers"], n_heads=cfg["siglip_heads"],
                                  image_size=cfg["image_size"])
    return c


def build_product_model(cfg, weights, device="cuda", **kw):
    from metamorph_b200.model import MetaMorphLlamaForCausalLM
    c = product_config(cfg, **kw)
    model = MetaMorphLlamaForCausalLM(c, vision_head="mlp", normalize_vision=True, vision_delay_load=True,
                                      device=device)
    model.get_vision_tower().load_model(device=device, allow_random_init=True)   # tower tensors follow via load_state_dict
    missing, unexpected = model.load_state_dict({k: v.to(torch.bfloat16) for k, v in weights.items()}, strict=False)
    assert not unexpected, unexpected
    assert not missing, missing
    for p in model.get_vision_tower().parameters():
        p.requires_grad = False
    return model
