"""CPU: checkpoint I/O in the reference's formats (metamorph_b200/checkpoint.py) — HF sharded
safetensors written straight from the fused layout, the stage-1 `mm_projector.bin` artefact, checkpoint discovery —
held to the reference model's own state_dict (tests/golden/reference_state_keys.json, made by
oracle/make_golden_checkpoint.py from /root/reference)."""
import json
import os

import pytest
import torch

from metamorph_b200 import checkpoint as ck


@pytest.fixture(scope="module")
def tiny_model():
    from oracle.weights import TINY, make_weights
    from tests.helpers import build_product_model
    torch.manual_seed(0)
    return build_product_model(TINY, make_weights(TINY), device="cpu", num_image_tokens=4)


def test_parse_size():
    assert ck.parse_size("5GB") == 5 * 10 ** 9 and ck.parse_size("200MB") == 2 * 10 ** 8
    assert ck.parse_size("1GiB") == 2 ** 30 and ck.parse_size(123) == 123
    with pytest.raises(ValueError):
        ck.parse_size("lots")


def test_saved_names_and_shapes_are_the_reference_state_dict(tiny_model, tmp_path):
    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_state_keys.json")) as fh:
        ref = json.load(fh)["state"]
    ck.save_model(tiny_model, str(tmp_path), max_shard_size="5GB")
    assert os.path.exists(tmp_path / "model.safetensors") and os.path.exists(tmp_path / "config.json")
    assert not os.path.exists(tmp_path / ck.WEIGHTS_INDEX_NAME)          # one shard -> no index, as HF
    sd = ck.load_model_state(str(tmp_path))
    extra = set(sd) - set(ref)
    assert not extra, f"names the reference does not know: {sorted(extra)[:5]}"
    missing = set(ref) - set(sd)
    # the SigLIP pooling head is never used by the path (siglip_encoder.py takes hidden states); it is carried through
    # only when the source checkpoint has it
    assert all(".vision_tower.vision_tower.head." in k for k in missing), sorted(missing)[:5]
    for k, v in sd.items():
        assert list(v.shape) == ref[k]["shape"], k
    assert not any("qkv_proj" in k or "gate_up_proj" in k for k in sd)   # fused device names never reach disk


def test_sharded_save_round_trip(tiny_model, tmp_path):
    wm = ck.save_model(tiny_model, str(tmp_path), max_shard_size="40MB")
    files = sorted(set(wm.values()))
    assert len(files) > 2 and all(f.startswith("model-0000") and f.endswith(f"-of-{len(files):05d}.safetensors") for f in files)
    with open(tmp_path / ck.WEIGHTS_INDEX_NAME) as fh:
        index = json.load(fh)
    want = dict(tiny_model.state_dict())
    assert index["weight_map"] == wm and set(wm) == set(want)
    assert index["metadata"]["total_size"] == sum(v.numel() * v.element_size() for v in want.values())
    from safetensors import safe_open
    limit = ck.parse_size("40MB")
    order = list(want)
    last = -1
    for f in files:
        with safe_open(str(tmp_path / f), framework="pt") as sf:
            assert sf.metadata() == {"format": "pt"}
            names = list(sf.keys())
        nbytes = sum(want[n].numel() * want[n].element_size() for n in names)
        assert nbytes <= limit or len(names) == 1
        idxs = sorted(order.index(n) for n in names)
        assert idxs[0] == last + 1 and idxs == list(range(idxs[0], idxs[0] + len(idxs)))   # greedy, in state_dict order
        last = idxs[-1]
    got = ck.load_model_state(str(tmp_path))
    for k, v in want.items():
        assert torch.equal(got[k], v.cpu()), k
    # and back into the fused layout through the class API
    from metamorph_b200.model import MetaMorphLlamaForCausalLM
    clone = MetaMorphLlamaForCausalLM.from_pretrained(str(tmp_path), torch_dtype=torch.bfloat16, device="cpu",
                                                      vision_head="mlp", normalize_vision=True)
    a, b = torch.nn.Module.state_dict(tiny_model), torch.nn.Module.state_dict(clone)
    for k, v in a.items():
        if "vision_tower" in k:
            continue                                          # the tower is delay-loaded (vision_delay_load=True)
        assert torch.equal(v, b[k]), k


def test_mm_projector_artifact_paths_and_round_trip(tiny_model, tmp_path):
    out = tmp_path / "run"
    dst = ck.save_mm_projector(tiny_model, str(out))
    assert dst == str(out / "mm_projector.bin") and os.path.exists(out / "config.json")
    w = torch.load(dst, map_location="cpu")
    assert sorted(w) == ["model.mm_projector.0.bias", "model.mm_projector.0.weight", "model.mm_projector.2.bias",
                         "model.mm_projector.2.weight"]
    dst2 = ck.save_mm_projector(tiny_model, str(out / "checkpoint-40"), use_im_start_end=True)
    assert dst2 == str(out / "mm_projector" / "checkpoint-40.bin")          # train.py:199-204
    w2 = torch.load(dst2, map_location="cpu")
    assert "model.embed_tokens.weight" in w2 and len(w2) == 5
    dst3 = ck.save_mm_projector_checkpoint(tiny_model, str(out), 60)
    assert dst3 == str(out / "checkpoint-60" / "mm_projector.bin") and os.path.exists(out / "checkpoint-60" / "config.json")
    proj = tiny_model.get_model().mm_projector
    saved = {k: v.detach().clone() for k, v in proj.state_dict().items()}
    with torch.no_grad():
        for p in proj.parameters():
            p.add_(1.0)
    res = ck.load_mm_projector(tiny_model, str(out))
    assert not res.unexpected_keys
    for k, v in proj.state_dict().items():
        assert torch.equal(v, saved[k]), k


def test_latest_checkpoint_discovery(tmp_path):
    assert ck.latest_checkpoint(str(tmp_path / "nope")) is None
    for n, complete in [(5, True), (20, True), (100, False), (7, True)]:
        d = tmp_path / f"checkpoint-{n}"
        d.mkdir()
        if complete:
            (d / ck.TRAINER_STATE_NAME).write_text("{}")
    (tmp_path / "checkpoint-final").mkdir()
    assert ck.latest_checkpoint(str(tmp_path)) == str(tmp_path / "checkpoint-20")   # 100 is incomplete (no state file)


def test_hf_loader_reads_our_shards(tiny_model, tmp_path):
    """HF's own `from_pretrained` (index parsing, shard loading, safetensors metadata) must accept the files: load the
    LLaMA part of the checkpoint into a stock `LlamaForCausalLM` and compare every tensor."""
    from transformers import LlamaConfig, LlamaForCausalLM
    from oracle.weights import TINY
    ck.save_model(tiny_model, str(tmp_path), max_shard_size="40MB")
    cfg = LlamaConfig(hidden_size=TINY["hidden"], intermediate_size=TINY["inter"], num_hidden_layers=TINY["layers"],
                      num_attention_heads=TINY["heads"], num_key_value_heads=TINY["kv_heads"], head_dim=TINY["head_dim"],
                      vocab_size=TINY["vocab"], rms_norm_eps=TINY["rms_eps"], tie_word_embeddings=False,
                      attention_bias=False)
    hf = LlamaForCausalLM.from_pretrained(str(tmp_path), config=cfg, torch_dtype=torch.bfloat16)
    ours = tiny_model.state_dict()
    n = 0
    for k, v in hf.state_dict().items():
        if "rotary" in k:
            continue
        assert torch.equal(v.cpu(), ours[k].cpu()), k
        n += 1
    assert n == 3 + 9 * TINY["layers"]          # embed, norm, lm_head + 9 tensors per decoder layer
