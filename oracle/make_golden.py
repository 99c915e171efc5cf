"""ORACLE (test infrastructure): generate golden vectors by running THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, read-only):
    PYTHONPATH=/root/repo python oracle/make_golden.py
Imports `metamorph.model.MetaMorphLlamaForCausalLM` from /root/reference, builds it at tiny LLaMA
dims + real-width 2-layer SigLIP with the deterministic weights of oracle/weights.py (SURVEY.md §8c
recipe: config-built tower injected to bypass the hub download; mm_vision_select_layer=-1), runs
forward / backward / generate on CPU and stores the results under tests/golden/.
Nothing here is imported by the product or by the GPU box (which has no /root/reference).
"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle.ref_model import activate, build_reference, pin_decode_mask_semantics  # noqa: E402
from oracle.weights import TINY, make_batch, make_weights  # noqa: E402

activate("/root/reference")   # first on sys.path: `metamorph` must be the reference, not this repo's alias package


def run_forward_case(model, ids, mask, labs, images, with_grads):
    model.zero_grad()
    out = model(input_ids=ids, attention_mask=mask, labels=labs, images=images.to(next(model.parameters()).dtype))
    res = dict(loss=out.loss.detach().float(), loss_language=torch.tensor(model.loss_language),
               loss_image_ar=torch.tensor(model.loss_image_ar), logits=out.logits.detach().float(),
               hidden=out.hidden_states.detach().float())
    if with_grads:
        out.loss.backward()
        g = {}
        for name, p in model.named_parameters():
            if p.grad is not None and "vision_tower" not in name:
                g[name] = p.grad.detach().float()
        res["grads"] = g
    return res


def main():
    torch.manual_seed(0)
    cfg = TINY
    W = make_weights(cfg)
    ids, mask, labs, images = make_batch(cfg)
    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    # ---------------- case 1: fp32 forward + backward of the reference
    ref = build_reference(cfg, W, torch.float32)
    with torch.no_grad():
        (_, pos, am, _, embeds, new_labels, impos, tgt) = ref.prepare_inputs_labels_for_multimodal(
            ids, None, mask, None, labs, images)
    r32 = run_forward_case(ref, ids, mask, labs, images, with_grads=True)
    # ---------------- case 2: same in bf16 (what the reference runs in production) for error budgets
    ref16 = build_reference(cfg, W, torch.bfloat16)
    with torch.no_grad():
        r16 = run_forward_case(ref16, ids, mask, labs, images, with_grads=False)

    g = torch.Generator().manual_seed(123)
    cols = torch.randint(0, cfg["vocab"], (48,), generator=g)
    cols = torch.cat([cols, torch.tensor([128256, 128257, 128001, 128009])])
    grads = r32["grads"]
    grad_digest = {}
    for k, v in grads.items():
        flat = v.reshape(-1)
        idx = torch.randint(0, flat.numel(), (64,), generator=g)
        grad_digest[k] = dict(norm=v.norm(), idx=idx, vals=flat[idx].clone())
    torch.save(dict(
        cfg=cfg, input_ids=ids, attention_mask=mask, labels=labs,
        new_labels=new_labels, image_positions=impos, new_attention_mask=am, targets=tgt.float(),
        inputs_embeds_sum=embeds.float().sum(-1), inputs_embeds_rows=embeds.float()[:, ::7, :16].clone(),
        loss=r32["loss"], loss_language=r32["loss_language"], loss_image_ar=r32["loss_image_ar"],
        logit_cols=cols, logits_sub=r32["logits"][..., cols].clone(), logits_argmax=r32["logits"].argmax(-1),
        hidden_sub=r32["hidden"][..., :32].clone(),
        bf16=dict(loss=r16["loss"], loss_language=r16["loss_language"], loss_image_ar=r16["loss_image_ar"],
                  logits_sub=r16["logits"][..., cols].clone(), hidden_sub=r16["hidden"][..., :32].clone()),
        grad_digest=grad_digest), os.path.join(out_dir, "forward_backward_tiny.pt"))
    print("forward/backward:", float(r32["loss"]), float(r32["loss_language"]), float(r32["loss_image_ar"]),
          "| bf16:", float(r16["loss"]))

    # ---------------- case 3: index-logic cases (integers only)
    idx_cases = []
    rng = torch.Generator().manual_seed(7)
    ref_small = build_reference(cfg, W, torch.float32, max_len=100)
    START, END, IMG = 128256, 128257, -200

    def mk(parts):
        s, l = [], []
        for kind, n, lab in parts:
            if kind == "t":
                t = torch.randint(0, 128000, (n,), generator=rng).tolist()
                s += t
                l += t if lab else [-100] * n
            elif kind == "img":
                s += [START, IMG, END]
                l += ([START, IMG, END] if lab else [-100] * 3)
        return s, l

    samples = [
        [mk([("t", 10, False), ("img", 0, False), ("t", 5, True), ("img", 0, True), ("t", 3, True)])],
        [mk([("t", 30, False), ("img", 0, False), ("t", 20, False), ("img", 0, True), ("t", 10, True)]),   # overflow at max_len=100
         mk([("t", 12, True)])],
        [mk([("t", 5, False), ("img", 0, True), ("t", 28, True), ("img", 0, True), ("t", 9, True)]),       # truncation
         mk([("t", 40, False), ("img", 0, True), ("t", 2, True)]),
         mk([("t", 7, True)])],
    ]
    for batch in samples:
        L = max(len(s) for s, _ in batch)
        bi = torch.full((len(batch), L), 128001, dtype=torch.long)
        bl = torch.full((len(batch), L), -100, dtype=torch.long)
        bm = torch.zeros((len(batch), L), dtype=torch.bool)
        n_img = 0
        for b, (s, l) in enumerate(batch):
            bi[b, :len(s)] = torch.tensor(s)
            bl[b, :len(l)] = torch.tensor(l)
            bm[b, :len(s)] = True
            n_img += max(1, s.count(IMG))
        feats = torch.arange(n_img * 64 * 1152, dtype=torch.float32).reshape(n_img, 64, 1152) / 1e6
        for side, model in (("right", ref_small),):
            with torch.no_grad():
                (_, pos, am, _, emb, nl, ip, tg) = model.prepare_inputs_labels_for_multimodal(
                    bi, None, bm, None, bl, None, image_embeds=feats)
            idx_cases.append(dict(input_ids=bi, attention_mask=bm, labels=bl, n_images=n_img, max_len=100,
                                  padding_side=side, new_labels=nl, image_positions=ip, new_attention_mask=am,
                                  target_first=tg[:, 0, 0].clone(), embeds_shape=tuple(emb.shape)))
    torch.save(idx_cases, os.path.join(out_dir, "interleave_cases.pt"))
    print("index cases:", len(idx_cases), [c["embeds_shape"] for c in idx_cases])

    # ---------------- case 4: greedy decode (no cache) of the reference, 4 visual tokens per image
    ref_dec = build_reference(cfg, W, torch.float32, num_image_tokens=4)
    pin_decode_mask_semantics(ref_dec)   # transformers 4.45 semantics of the [1,1] mask (oracle/ref_model.py)
    prompt = torch.tensor([[128000] + torch.randint(0, 128000, (11,), generator=rng).tolist()])
    with torch.no_grad():
        first = ref_dec.generate(prompt, max_new_tokens=0)[0]
        t0 = int(first[0])
        out_ids, img = ref_dec.generate(prompt, output_image=True, max_new_tokens=9, start_image_token_id=t0)
    torch.save(dict(prompt=prompt, start_image_token_id=t0, ids=out_ids[0].clone(), image_embeds=img.float(),
                    max_new_tokens=9, num_image_tokens=4), os.path.join(out_dir, "greedy_decode_tiny.pt"))
    print("decode:", out_ids[0].tolist(), tuple(img.shape))


if __name__ == "__main__":
    main()
