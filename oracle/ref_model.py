"""ORACLE (test infrastructure): build THE REFERENCE's own model class on CPU with the SURVEY.md section 8c recipe.

`activate(root)` puts a reference checkout first on sys.path — `/root/reference` in the build container, or the copy
`oracle/build_ref.py` vendors to `oracle/_ref` (git-ignored; it travels to the GPU box) — and checks that `metamorph`
really resolves there and not to this repository's alias package of the same name. Only golden-vector generators,
`oracle/ref_bench.py` (the CPU arm of bench.py) and tests may import this module; the product never does.
"""
import os
import sys

import torch


def activate(root: str) -> str:
    root = os.path.abspath(root)
    if not os.path.isdir(os.path.join(root, "metamorph", "model")):
        raise FileNotFoundError(f"no reference checkout at {root} (run oracle/build_ref.py in the build container)")
    os.environ.setdefault("WANDB_MODE", "disabled")
    os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
    sys.dont_write_bytecode = True
    for name in [m for m in sys.modules if m == "metamorph" or m.startswith("metamorph.")]:
        del sys.modules[name]
    if root in sys.path:
        sys.path.remove(root)
    sys.path.insert(0, root)
    import metamorph
    got = os.path.abspath(os.path.dirname(metamorph.__file__))
    assert got.startswith(root), f"`metamorph` resolved to {got}, not the reference under {root}"
    return root


def build_reference(cfg, weights=None, dtype=torch.float32, num_image_tokens=None, max_len=None, attn="eager"):
    """MetaMorphLlamaForCausalLM (metamorph_llama.py:226) at the dims of `cfg`, tower injected from a config (no hub
    download, siglip_encoder.py:113), mm_vision_select_layer=-1. weights=None keeps the HF random init (timing runs)."""
    from metamorph.model import MetaMorphLlamaForCausalLM
    from metamorph.model.language_model.metamorph_llama import MetaMorphConfig
    from transformers import SiglipVisionConfig, SiglipVisionModel
    c = MetaMorphConfig(hidden_size=cfg["hidden"], intermediate_size=cfg["inter"],
                        num_hidden_layers=cfg["layers"], num_attention_heads=cfg["heads"],
                        num_key_value_heads=cfg["kv_heads"], head_dim=cfg["head_dim"], vocab_size=cfg["vocab"],
                        rms_norm_eps=cfg["rms_eps"], rope_theta=cfg["rope_theta"],
                        max_position_embeddings=8192, attention_bias=False, tie_word_embeddings=False)
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
    c._attn_implementation = attn
    model = MetaMorphLlamaForCausalLM(c, vision_head="mlp", normalize_vision=True)
    vt = model.get_vision_tower()
    vt.vision_tower = SiglipVisionModel(SiglipVisionConfig(
        hidden_size=cfg["siglip_width"], intermediate_size=cfg["siglip_inter"],
        num_hidden_layers=cfg["siglip_layers"], num_attention_heads=cfg["siglip_heads"],
        image_size=cfg["image_size"], patch_size=14))
    vt.is_loaded = True
    if weights is not None:
        sd = {}
        tp = "model.vision_tower.vision_tower."
        for k, v in weights.items():
            if k.startswith(tp):
                sd[tp + "vision_model." + k[len(tp):]] = v
            else:
                sd[k] = v
        missing, unexpected = model.load_state_dict(sd, strict=False)
        missing = [m for m in missing if ".head." not in m and "rotary" not in m]
        assert not unexpected, unexpected
        assert not missing, missing
    model = model.to(dtype)
    model.eval()
    return model


def pin_decode_mask_semantics(model):
    """Version-drift shim: greedy_decode passes a [1,1] all-ones attention_mask together with the
    full-length inputs_embeds (metamorph_llama.py:524). Under the pinned transformers 4.45 that mask is a no-op (pure
    causal attention); transformers 5.x broadcasts it into a different mask. Drop it so the run carries the
    pinned-version semantics. The reference's own code is untouched: only the argument it passes on is filtered."""
    orig = model.llm_forward

    def llm_forward_445(*a, **kw):
        am = kw.get("attention_mask")
        if am is not None and am.shape[-1] == 1 and kw["inputs_embeds"].shape[1] != 1:
            kw["attention_mask"] = None
        return orig(*a, **kw)

    model.llm_forward = llm_forward_445
    return model
