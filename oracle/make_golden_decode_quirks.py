"""ORACLE (test infrastructure): golden vectors of the reference's `greedy_decode` for its two state-machine quirks.

    PYTHONPATH=/root/repo python oracle/make_golden_decode_quirks.py        (build container; needs /root/reference)

metamorph_llama.py:502-597 has two behaviours a faithful re-implementation must reproduce and that ordinary prompts
do not reach:
  Q1 "EOS inside an image": while emitting an image's visual tokens the loop still tests `next_token in eos_token_id`
     (:583) on the argmax of logits computed from the hidden state that the decoding branch has OVERWRITTEN with the
     projector's prediction (:363-377); a hit ends generation in the middle of the image (fewer than num_image_tokens
     embeddings are returned).
  Q2 "<image_start> twice without <image_end>": after an image, `total_image_tokens` stays at num_image_tokens until an
     <image_end> resets it (:568-572). A second <image_start> before that sets `in_image_mode` (:549-553) but the image
     branch (:556) is closed, so tokens keep being appended as TEXT while every forward runs with decoding=True, i.e.
     every later token is the argmax over the overwritten hidden state.
The model weights are random (oracle/weights.py TINY, 4 visual tokens per image; lm_head reduced to 16 live vocabulary
rows by `with_sparse_lm_head` so that logit margins exceed bf16 noise), so the token ids that play
<image_start> / <eos> are chosen after the fact, exactly as the first fixture does (make_golden.py case 4): a search
over seeded prompts with the CPU restatement (oracle/restatement.py, itself pinned to the reference) finds prompts
whose free-running trajectory hits the quirk with a top-1/top-2 logit margin far above bf16 noise at EVERY step, so
that a bf16 implementation can be held to token-exactness. The fixture itself is then produced by THE REFERENCE's own
generate(); the restatement's trajectory must agree with it or the script aborts.
"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import restatement as R  # noqa: E402
from oracle.ref_model import activate, build_reference, pin_decode_mask_semantics  # noqa: E402
from oracle.weights import TINY, make_weights, with_sparse_lm_head  # noqa: E402

NTOK = 4
MAXNEW = 12
MIN_MARGIN = 0.08      # (top1 - top2) / max|logit| at EVERY step; bf16 noise is ~0.5 % of the logit scale
LIVE_ROWS = 16


def decode(W, cfg, prompt, start, eos, max_new=MAXNEW):
    tr = []
    emb = W["model.embed_tokens.weight"][prompt].float()
    ids, imgs = R.greedy_decode_nocache(W, cfg, emb, max_new, start_id=start, end_id=-1, eos=tuple(eos), trace=tr)
    return ids, imgs, tr


def search(W, cfg):
    best = {"q1": None, "q2": None}
    for seed in range(300):
        g = torch.Generator().manual_seed(1000 + seed)
        prompt = torch.tensor([[128000] + torch.randint(0, 128000, (11,), generator=g).tolist()])
        _, _, tr0 = decode(W, cfg, prompt, -1, ())
        for i in range(0, 4):                       # candidate <image_start> = the token first emitted at step i
            s = tr0[i]["tok"]
            if s in [t["tok"] for t in tr0[:i]]:
                continue
            ids, imgs, tr = decode(W, cfg, prompt, s, ())
            margin = min(t["margin"] / t["scale"] for t in tr)
            if margin < MIN_MARGIN:
                continue
            # Q2: s re-emitted after the image block with n_img_tok == NTOK, at least 3 more steps follow
            hits = [k for k, t in enumerate(tr) if k > i + NTOK and t["tok"] == s and not t["in_image"] and t["n_img_tok"] == NTOK]
            if hits and hits[0] + 3 < len(tr):
                if best["q2"] is None or margin > best["q2"]["margin"]:
                    best["q2"] = dict(seed=seed, prompt=prompt, start=s, eos=[], margin=margin, enter=hits[0])
            # Q1: EOS := the argmax at the 3rd visual-token step of the first image (must not occur before that step)
            k = i + 3
            e = tr[k]["tok"]
            if tr[k]["in_image"] and e not in [t["tok"] for t in tr[:k]]:
                if best["q1"] is None or margin > best["q1"]["margin"]:
                    best["q1"] = dict(seed=seed, prompt=prompt, start=s, eos=[e], margin=margin, enter=k)
    return best


def main():
    cfg = dict(TINY, image_tokens=NTOK)
    W, live = with_sparse_lm_head(make_weights(TINY), LIVE_ROWS)
    print("live vocabulary rows:", live)
    best = search(W, cfg)
    print({k: (None if v is None else {x: v[x] for x in ("seed", "start", "eos", "margin", "enter")}) for k, v in best.items()})
    assert best["q1"] is not None and best["q2"] is not None, "search found no case with a safe margin"
    activate("/root/reference")
    ref = build_reference(TINY, W, torch.float32, num_image_tokens=NTOK)
    pin_decode_mask_semantics(ref)
    out = {}
    for name, c in best.items():
        ids_o, imgs_o, tr = decode(W, cfg, c["prompt"], c["start"], c["eos"])
        with torch.no_grad():
            ids_r, img_r = ref.generate(c["prompt"], output_image=True, max_new_tokens=MAXNEW, start_image_token_id=c["start"],
                                        end_image_token_id=-1, eos_token_id=list(c["eos"]))
        assert ids_r[0].tolist() == ids_o, (name, ids_r[0].tolist(), ids_o)
        assert tuple(img_r.shape) == tuple(imgs_o.shape)
        assert float((img_r.float() - imgs_o).abs().max()) < 1e-4
        out[name] = dict(prompt=c["prompt"], start_image_token_id=c["start"], end_image_token_id=-1, eos_token_id=list(c["eos"]),
                         max_new_tokens=MAXNEW, num_image_tokens=NTOK, live_rows=LIVE_ROWS, ids=ids_r[0].clone(), image_embeds=img_r.float().clone(),
                         min_margin=min(t["margin"] / t["scale"] for t in tr), trace=[(t["tok"], t["in_image"], t["n_img_tok"]) for t in tr])
        print(name, "ids", ids_r[0].tolist(), "embeds", tuple(img_r.shape), "min relative margin %.4f" % out[name]["min_margin"],
              [(t["tok"], int(t["in_image"]), t["n_img_tok"]) for t in tr])
    torch.save(out, os.path.join(REPO, "tests", "golden", "greedy_decode_quirks.pt"))


if __name__ == "__main__":
    main()
