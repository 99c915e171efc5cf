"""Inputs of the input-contract fixtures (tests/golden/input_contract.json): shared by the generator
(oracle/make_golden_inputs.py, runs the reference) and the test (tests/test_input_contract.py). No side effects."""
from types import SimpleNamespace

import torch


class ToyTokenizer:
    def __init__(self, add_bos=True, pad_token_id=0, model_max_length=32):
        self.bos_token_id = 1
        self.pad_token_id = pad_token_id
        self.model_max_length = model_max_length
        self.add_bos = add_bos

    def __call__(self, text):
        ids = [2 + (sum(ord(c) * (i + 1) for i, c in enumerate(w)) % 997) for w in text.split()]
        return SimpleNamespace(input_ids=([self.bos_token_id] if self.add_bos else []) + ids)


PROMPTS = [
    "hello world", "<image>", "<image> describe this", "look at <image> and <image> then answer", "tail image <image>",
    "<image><image>", "", "a b c <image>\nnew line <image> end", "<image_start><image><image_end> wrapped",
]
SOURCES = [
    [{"from": "human", "value": "<image>\nWhat is this?"}, {"from": "gpt", "value": "A cat."}],
    [{"from": "human", "value": "Draw a dog"}, {"from": "gpt", "value": "Here: <image> and again <image>"}],
    [{"from": "human", "value": "no image here"}, {"from": "gpt", "value": "ok"}],
]


def collator_cases():
    g = torch.Generator().manual_seed(0)
    def inst(n, n_img):
        d = {"input_ids": torch.randint(2, 900, (n,), generator=g), "labels": torch.randint(-100, 900, (n,), generator=g)}
        if n_img is not None:
            d["image"] = [torch.randn(3, 4, 4, generator=g) for _ in range(n_img)]
        return d
    return {"ragged_with_images": ([inst(5, 1), inst(9, 2), inst(3, 1)], 32),
            "truncated": ([inst(40, 1), inst(12, 1)], 16),
            "text_only_keys": ([inst(4, None), inst(6, None)], 32)}
