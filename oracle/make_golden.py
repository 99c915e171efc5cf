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

os.environ.setdefault("WANDB_MODE", "disabled")
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
sys.dont_write_bytecode = True
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")   # first: `metamorph` must be the reference, not this repo's alias package

from oracle.weights import TINY, make_batch, make_weights  # noqa: E402


def build_reference(cfg, weights, dtype=torch.float32, num_image_tokens=None, max_len=None):
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
    c._attn_implementation = "eager"
    model = MetaMorphLlamaForCausalLM(c, vision_head="mlp", normalize_vision=True)
    vt = model.get_vision_tower()
    vt.vision_tower = SiglipVisionModel(SiglipVisionConfig(
        hidden_size=cfg["siglip_width"], intermediate_size=cfg["siglip_inter"],
        num_hidden_layers=cfg["siglip_layers"], num_attention_heads=cfg["siglip_heads"],
        image_size=cfg["image_size"], patch_size=14))
    vt.is_loaded = True
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
    # Version-drift shim: greedy_decode passes a [1,1] all-ones attention_mask together
    # with the full-length inputs_embeds (metamorph_llama.py:524). Under the pinned transformers 4.45
    # that mask is a no-op (pure causal attention); transformers 5.x broadcasts it into a different
    # mask. Drop it so the golden vectors carry the pinned-version semantics.
    _orig_llm_forward = ref_dec.llm_forward

    def _llm_forward_445(*a, **kw):
        am = kw.get("attention_mask")
        if am is not None and am.shape[-1] == 1 and kw["inputs_embeds"].shape[1] != 1:
            kw["attention_mask"] = None
        return _orig_llm_forward(*a, **kw)

    ref_dec.llm_forward = _llm_forward_445
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
