"""ORACLE (test infrastructure): the reference's own forward on the TINY batch with tokenizer_padding_side == "left"
(metamorph_arch.py:373-386) -> tests/golden/leftpad_tiny.pt (losses fp32 + bf16, padded labels / mask / image positions).
    PYTHONPATH=/root/repo python oracle/make_golden_leftpad.py      (build container; needs /root/reference)
"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle.ref_model import activate, build_reference  # noqa: E402
from oracle.weights import TINY, make_batch, make_weights  # noqa: E402

activate("/root/reference")


def main():
    W = make_weights(TINY)
    ids, mask, labs, images = make_batch(TINY)
    out = {}
    for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        ref = build_reference(TINY, W, dt)
        ref.config.tokenizer_padding_side = "left"
        with torch.no_grad():
            (_, pos, am, _, emb, nl, ip, tg) = ref.prepare_inputs_labels_for_multimodal(ids, None, mask, None, labs, images.to(dt))
            o = ref(input_ids=ids, attention_mask=mask, labels=labs, images=images.to(dt))
        out[name] = dict(loss=o.loss.float().clone(), loss_language=torch.tensor(ref.loss_language),
                         loss_image_ar=torch.tensor(ref.loss_image_ar))
        if name == "fp32":
            out.update(new_labels=nl.clone(), new_attention_mask=am.clone(), image_positions=ip.clone(),
                       hidden_last=o.hidden_states[:, -1, :32].float().clone())
        print(name, float(o.loss), ref.loss_language, ref.loss_image_ar)
    torch.save(out, os.path.join(REPO, "tests", "golden", "leftpad_tiny.pt"))


if __name__ == "__main__":
    main()
