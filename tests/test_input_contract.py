"""CPU: the host-side input contract (tokenizer_image_token, preprocess_multimodal, the collator)
against fixtures produced by the reference's own functions (tests/golden/input_contract.json,
oracle/make_golden_inputs.py) — integer outputs bit-exact."""
import copy
import json
import os
from types import SimpleNamespace

import pytest
import torch

from metamorph_b200.mm_utils import DataCollatorForSupervisedDataset, preprocess_multimodal, tokenizer_image_token
from oracle.input_cases import SOURCES, ToyTokenizer, collator_cases

GOLD = os.path.join(os.path.dirname(__file__), "golden", "input_contract.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as fh:
        return json.load(fh)


def test_tokenizer_image_token(gold):
    assert len(gold["tokenizer_image_token"]) >= 16
    for case in gold["tokenizer_image_token"]:
        tok = ToyTokenizer(add_bos=case["add_bos"])
        assert tokenizer_image_token(case["prompt"], tok) == case["ids"], case["prompt"]
        t = tokenizer_image_token(case["prompt"], tok, return_tensors="pt")
        assert t.dtype == torch.long and t.tolist() == case["ids"]
    with pytest.raises(ValueError):
        tokenizer_image_token("x", ToyTokenizer(), return_tensors="np")


def test_preprocess_multimodal(gold):
    for case in gold["preprocess_multimodal"]:
        src = copy.deepcopy(SOURCES)
        args = SimpleNamespace(is_multimodal=case["is_multimodal"], mm_use_im_start_end=case["mm_use_im_start_end"])
        assert preprocess_multimodal(src, args) == case["result"]


def test_collator(gold):
    for name, (instances, max_len) in collator_cases().items():
        want = gold["collator"][name]
        b = DataCollatorForSupervisedDataset(tokenizer=ToyTokenizer(model_max_length=max_len))(instances)
        assert set(b) == {k for k in want if k != "images_sum"}
        assert b["input_ids"].tolist() == want["input_ids"] and b["labels"].tolist() == want["labels"]
        assert b["attention_mask"].dtype == torch.bool and b["attention_mask"].tolist() == want["attention_mask"]
        if "images" in want:
            assert list(b["images"].shape) == want["images"]
            assert float(b["images"].double().sum()) == want["images_sum"]
