"""CPU: where weights may come from: the SigLIP tower never silently stays random
(reference: `siglip_encoder.py:113` downloads the pretrained tower), and a non-strict checkpoint load REPORTS what it
did not find (a core LLaMA tensor missing is an error, `metamorph_llama.py` from_pretrained / builder.py:79-92)."""
import warnings
from types import SimpleNamespace

import pytest
import torch

from metamorph_b200.model.siglip_tower import SiglipVisionTower

DIMS = dict(width=64, inter=128, n_layers=2, n_heads=2, image_size=28, patch=14)


def _tower(**kw):
    args = SimpleNamespace(mm_vision_tower_dims=DIMS, mm_vision_tower_path="/nonexistent/siglip", **kw)
    return SiglipVisionTower("siglip/CLIP-ViT-SO400M-14-384", args, delay_load=True)   # the name the reference scripts pass


def test_tower_without_weights_raises(monkeypatch):
    monkeypatch.delenv("MM_SIGLIP_PATH", raising=False)
    monkeypatch.setenv("HF_HUB_OFFLINE", "1")
    t = _tower()
    with pytest.raises(RuntimeError, match="no SigLIP weights available"):
        t.load_model(device="cpu")
    assert not t.is_loaded


def test_tower_random_init_only_on_request(monkeypatch):
    monkeypatch.delenv("MM_SIGLIP_PATH", raising=False)
    monkeypatch.setenv("HF_HUB_OFFLINE", "1")
    t = _tower()
    with pytest.warns(UserWarning, match="RANDOMLY INITIALISED"):
        t.load_model(device="cpu", allow_random_init=True)
    assert t.is_loaded and t.weights_source == "random init"
    t2 = _tower(mm_vision_tower_random_init=True)         # the config switch synthetic runs use
    with pytest.warns(UserWarning, match="RANDOMLY INITIALISED"):
        t2.load_model(device="cpu")
    assert t2.weights_source == "random init"


def test_tower_takes_checkpoint_tensors_and_rejects_incomplete_ones(monkeypatch):
    monkeypatch.setenv("HF_HUB_OFFLINE", "1")
    src = _tower()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        src.load_model(device="cpu", allow_random_init=True)
    sd = {k: v.clone() for k, v in src.vision_tower.state_dict().items()}
    # (1) explicit state_dict argument
    t = _tower()
    t.load_model(device="cpu", state_dict=sd)
    assert t.weights_source == "state_dict argument"
    for k, v in t.vision_tower.state_dict().items():
        assert torch.equal(v, sd[k]), k
    # (2) tensors stashed from the model checkpoint while the tower was delay-loaded; pooling-head tensors are kept aside
    t = _tower()
    t.stash_checkpoint_state({**sd, "head.probe": torch.zeros(1, 1, 64)})
    t.load_model(device="cpu")
    assert t.weights_source == "model checkpoint" and "head.probe" in t.vision_tower._extra_state_tensors
    # (3) a checkpoint that lacks encoder tensors must not leave them random
    bad = {k: v for k, v in sd.items() if "layers.1.mlp" not in k}
    t = _tower()
    with pytest.raises(RuntimeError, match="lack"):
        t.load_model(device="cpu", state_dict=bad)


def test_checkpoint_key_report():
    from metamorph_b200.model.metamorph_llama import MetaMorphLlamaForCausalLM as M
    R = SimpleNamespace
    # a core tensor missing (e.g. a renamed decoder weight) is an error, not a silent random init
    with pytest.raises(RuntimeError, match="core tensors are missing"):
        M.check_loaded_keys(M, R(missing_keys=["model.layers.0.self_attn.q_proj.weight"], unexpected_keys=[]), "load")
    with pytest.raises(RuntimeError, match="core tensors are missing"):
        M.check_loaded_keys(M, R(missing_keys=["lm_head.weight"], unexpected_keys=[]), "load")
    # projector / vision head of a plain LLaMA base, unknown extras: reported
    with pytest.warns(UserWarning) as rec:
        M.check_loaded_keys(M, R(missing_keys=["model.vision_proj.0.weight", "vision_head.fc1.weight"],
                                 unexpected_keys=["model.layers.0.self_attn.rotary_emb.inv_freq", "foo.bar"]), "load")
    text = " | ".join(str(w.message) for w in rec)
    assert "model.vision_proj" in text and "vision_head.fc1" in text and "foo.bar" in text and "inv_freq" not in text
    # nothing to report: silent
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        M.check_loaded_keys(M, R(missing_keys=[], unexpected_keys=["model.layers.3.self_attn.rotary_emb.inv_freq"]), "load")
