"""Generates tests/golden/input_contract.json by running the REFERENCE's own host helpers:
  * `tokenizer_image_token` imported from /root/reference/metamorph/mm_utils.py;
  * `preprocess_multimodal` and `DataCollatorForSupervisedDataset`: metamorph/train/train.py cannot be imported in this
    container, so their definitions are cut out of the reference file with `ast` and executed unmodified in a
    namespace that provides the names they use.
A toy tokenizer (deterministic word -> id map, optional BOS) stands in for the LLaMA-3 tokenizer: the helpers only use
`tokenizer(text).input_ids`, `bos_token_id`, `pad_token_id` and `model_max_length`.

    cd /tmp && PYTHONPATH=/root/repo python /root/repo/oracle/make_golden_inputs.py
"""
import ast
import copy
import json
import os
import sys
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, Sequence

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


from oracle.input_cases import PROMPTS, SOURCES, ToyTokenizer, collator_cases  # noqa: E402


def reference_train_symbols():
    src = open("/root/reference/metamorph/train/train.py").read()
    tree = ast.parse(src)
    want = {"preprocess_multimodal", "DataCollatorForSupervisedDataset"}
    nodes = [n for n in tree.body if getattr(n, "name", None) in want]
    assert {n.name for n in nodes} == want
    import transformers
    from metamorph.constants import (DEFAULT_IM_END_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IMAGE_TOKEN, IGNORE_INDEX)
    ns = dict(torch=torch, transformers=transformers, dataclass=dataclass, Dict=Dict, Sequence=Sequence,
              DataArguments=object, IGNORE_INDEX=IGNORE_INDEX, DEFAULT_IMAGE_TOKEN=DEFAULT_IMAGE_TOKEN,
              DEFAULT_IM_START_TOKEN=DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN=DEFAULT_IM_END_TOKEN)
    exec(compile(ast.Module(body=nodes, type_ignores=[]), "reference:train.py", "exec"), ns)
    return ns["preprocess_multimodal"], ns["DataCollatorForSupervisedDataset"]


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference")          # `metamorph` must be the reference, not this repo's alias package
    os.environ.setdefault("WANDB_MODE", "disabled")
    from metamorph.mm_utils import tokenizer_image_token
    preprocess_multimodal, Collator = reference_train_symbols()
    out = {"tokenizer_image_token": [], "preprocess_multimodal": [], "collator": {}}
    for add_bos in (True, False):
        tok = ToyTokenizer(add_bos=add_bos)
        for p in PROMPTS:
            out["tokenizer_image_token"].append({"prompt": p, "add_bos": add_bos,
                                                 "ids": tokenizer_image_token(p, tok)})
    for flag in (True, False):
        for mm in (True, False):
            src = copy.deepcopy(SOURCES)
            res = preprocess_multimodal(src, SimpleNamespace(is_multimodal=mm, mm_use_im_start_end=flag))
            out["preprocess_multimodal"].append({"mm_use_im_start_end": flag, "is_multimodal": mm, "result": res})
    for name, (instances, max_len) in collator_cases().items():
        b = Collator(tokenizer=ToyTokenizer(model_max_length=max_len))(instances)
        out["collator"][name] = {k: (v.tolist() if k != "images" else list(v.shape)) for k, v in b.items()}
        if "images" in b:
            out["collator"][name]["images_sum"] = float(b["images"].double().sum())
    path = os.path.join(REPO, "tests", "golden", "input_contract.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
        fh.write("\n")
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
