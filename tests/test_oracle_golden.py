"""CPU tests (no GPU): the oracle restatement and the product's host index logic against golden
vectors produced by the REFERENCE ITSELF (oracle/make_golden.py, committed under tests/golden/)."""
import os

import pytest
import torch

from oracle import restatement as R
from oracle.weights import TINY, make_batch, make_weights

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def fb():
    return torch.load(os.path.join(G, "forward_backward_tiny.pt"), weights_only=False)


@pytest.fixture(scope="module")
def weights():
    return make_weights(TINY)


def test_oracle_full_forward_matches_reference(fb, weights):
    ids, mask, labs, images = make_batch(TINY)
    assert torch.equal(ids, fb["input_ids"]) and torch.equal(labs, fb["labels"])
    out = R.full_forward(weights, TINY, ids, mask, labs, images)
    plan = out["plan"]
    assert torch.equal(torch.tensor(plan["labels"]), fb["new_labels"])
    assert torch.equal(torch.tensor(plan["image_positions"]), fb["image_positions"])
    assert torch.equal(torch.tensor(plan["mask"]), fb["new_attention_mask"].bool())
    torch.testing.assert_close(out["inputs_embeds"].sum(-1), fb["inputs_embeds_sum"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(out["feats"][plan["targets"]], fb["targets"], rtol=1e-4, atol=1e-5)
    valid = fb["new_attention_mask"].bool()   # rows at padded positions are unspecified in the reference
    torch.testing.assert_close(out["logits"][..., fb["logit_cols"]][valid], fb["logits_sub"][valid], rtol=1e-3, atol=2e-4)
    torch.testing.assert_close(out["hidden"][..., :32][valid], fb["hidden_sub"][valid], rtol=1e-3, atol=2e-4)
    for k in ("loss", "loss_language", "loss_image_ar"):
        assert abs(float(out[k]) - float(fb[k])) < 1e-4, k


def test_oracle_gradients_match_reference(fb, weights):
    ids, mask, labs, images = make_batch(TINY)
    names = [k for k in fb["grad_digest"] if k in weights]
    for k in names:
        weights[k] = weights[k].clone().requires_grad_(True)
    out = R.full_forward(weights, TINY, ids, mask, labs, images)
    out["loss"].backward()
    checked = 0
    for k in names:
        d = fb["grad_digest"][k]
        g = weights[k].grad
        if g is None:
            continue
        torch.testing.assert_close(g.reshape(-1)[d["idx"]], d["vals"], rtol=2e-3, atol=1e-6)
        assert abs(float(g.norm()) - float(d["norm"])) <= 2e-3 * float(d["norm"]) + 1e-7, k
        checked += 1
    assert checked >= 20


def test_oracle_index_cases_match_reference():
    cases = torch.load(os.path.join(G, "interleave_cases.pt"), weights_only=False)
    for c in cases:
        plan = R.interleave_reference(c["input_ids"].tolist(), c["attention_mask"].tolist(), c["labels"].tolist(),
                                      c["n_images"], 64, c["max_len"], c["padding_side"])
        assert torch.equal(torch.tensor(plan["labels"]), c["new_labels"])
        assert torch.equal(torch.tensor(plan["image_positions"]), c["image_positions"])
        assert torch.equal(torch.tensor(plan["mask"]), c["new_attention_mask"].bool())
        feats0 = torch.arange(c["n_images"] * 64 * 1152, dtype=torch.float32).reshape(c["n_images"], 64, 1152)[:, 0, 0] / 1e6
        torch.testing.assert_close(feats0[plan["targets"]], c["target_first"])


def test_product_plan_matches_reference_and_oracle(fb):
    from metamorph_b200.model.interleave_plan import build_interleave_plan
    cases = torch.load(os.path.join(G, "interleave_cases.pt"), weights_only=False)
    cases.append(dict(input_ids=fb["input_ids"], attention_mask=fb["attention_mask"], labels=fb["labels"],
                      n_images=4, max_len=4096, padding_side="right", new_labels=fb["new_labels"],
                      image_positions=fb["image_positions"], new_attention_mask=fb["new_attention_mask"]))
    for c in cases:
        plan = build_interleave_plan(c["input_ids"], c["attention_mask"], c["labels"], c["n_images"], 64,
                                     c["max_len"], c["padding_side"])
        assert torch.equal(plan.labels, c["new_labels"])
        assert torch.equal(plan.image_positions, c["image_positions"])
        assert torch.equal(plan.attention_mask, c["new_attention_mask"].bool())
        ora = R.interleave_reference(c["input_ids"].tolist(), c["attention_mask"].tolist(), c["labels"].tolist(),
                                     c["n_images"], 64, c["max_len"], c["padding_side"])
        assert plan.target_image_idx == ora["targets"]
        assert plan.image_placeholder == ora["placeholder"]
        assert plan.position_ids.tolist() == ora["position_ids"]
        # row map vs oracle rows
        for b, rows in enumerate(ora["rows"]):
            for t, r in enumerate(rows):
                v = int(plan.row_map[b, t])
                if r[0] == "t":
                    assert v == r[1]
                elif r[0] == "i":
                    assert v == -(2 + r[1] * 64 + r[2])
                else:
                    assert v == -1


def test_product_plan_left_padding_and_errors():
    from metamorph_b200.model.interleave_plan import build_interleave_plan
    ids = torch.tensor([[1, 2, 128256, -200, 128257, 3], [4, 5, 6, 0, 0, 0]])
    mask = torch.tensor([[1, 1, 1, 1, 1, 1], [1, 1, 1, 0, 0, 0]]).bool()
    labs = ids.clone()
    plan = build_interleave_plan(ids, mask, labs, 2, 4, 4096, "left")
    ora = R.interleave_reference(ids.tolist(), mask.tolist(), labs.tolist(), 2, 4, 4096, "left")
    assert plan.labels.tolist() == ora["labels"] and plan.attention_mask.tolist() == ora["mask"]
    assert plan.position_ids.tolist() == ora["position_ids"]
    with pytest.raises(IndexError):  # <image> as the very first token: reference indexes an empty chunk
        build_interleave_plan(torch.tensor([[-200, 1]]), None, torch.tensor([[-100, 1]]), 1, 4, 4096)
    with pytest.raises(TypeError):   # reference compares int > None when tokenizer_model_max_length is unset
        build_interleave_plan(torch.tensor([[1, -200, 1]]), None, torch.tensor([[1, -200, 1]]), 1, 4, None)


def test_oracle_greedy_decode_matches_reference(weights):
    d = torch.load(os.path.join(G, "greedy_decode_tiny.pt"), weights_only=False)
    cfg = dict(TINY, image_tokens=d["num_image_tokens"])
    x = weights["model.embed_tokens.weight"][d["prompt"]]
    ids, img = R.greedy_decode_nocache(weights, cfg, x, d["max_new_tokens"], start_id=d["start_image_token_id"])
    assert ids == d["ids"].tolist()
    torch.testing.assert_close(img, d["image_embeds"], rtol=1e-3, atol=1e-5)
