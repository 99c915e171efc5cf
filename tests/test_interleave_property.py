"""CPU property test: the product's interleave plan (metamorph_b200/model/interleave_plan.py, zero device work) against
the oracle restatement of `prepare_inputs_labels_for_multimodal` (metamorph_arch.py:259-423; the restatement itself is
pinned to outputs of the reference on tests/golden/interleave_cases.pt) over randomly drawn ragged batches: any number of
images per sample, with and without <image_start>/<image_end>, prompt vs answer images, text-only samples (dummy image),
right / left padding, truncation at model_max_length inside or in front of an image — and the reference's error cases
(an <image> with no text before it: IndexError, as the reference's `cur_labels_noim[i][-1]` on an empty chunk)."""
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import restatement as R

IMG, START, END = -200, 128256, 128257


@st.composite
def batches(draw):
    B = draw(st.integers(1, 4))
    image_len = draw(st.sampled_from([1, 4, 7]))
    rows = []
    for _ in range(B):
        n_seg = draw(st.integers(1, 5))
        ids, labs = [], []
        answer_from = draw(st.integers(0, n_seg))           # segments >= answer_from carry labels (the "answer")
        for s in range(n_seg):
            n_text = draw(st.integers(0 if s else 1, 6))    # the very first token is never <image> (separate error test)
            toks = draw(st.lists(st.integers(0, 999), min_size=n_text, max_size=n_text))
            ids += toks
            labs += toks if s >= answer_from else [-100] * n_text
            if s < n_seg - 1 and draw(st.booleans()):
                wrap = draw(st.booleans())                 # mm_use_im_start_end on / off
                if wrap:
                    ids.append(START)
                    labs.append(START if s + 1 >= answer_from else -100)
                ids.append(IMG)
                labs.append(-100)
                if wrap:
                    ids.append(END)
                    labs.append(END if s + 1 >= answer_from else -100)
        rows.append((ids, labs))
    L = max(len(r[0]) for r in rows)
    ids = torch.zeros(B, L, dtype=torch.long)
    labs = torch.full((B, L), -100, dtype=torch.long)
    mask = torch.zeros(B, L, dtype=torch.bool)
    for b, (i, l) in enumerate(rows):
        ids[b, :len(i)] = torch.tensor(i)
        labs[b, :len(l)] = torch.tensor(l)
        mask[b, :len(i)] = True
    n_images = sum(max(1, r[0].count(IMG)) for r in rows)   # text-only samples consume one dummy image (:275-284)
    max_len = draw(st.sampled_from([4096, 12, 9, 5]))
    side = draw(st.sampled_from(["right", "left"]))
    return ids, mask, labs, n_images, image_len, max_len, side


@settings(max_examples=300, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(batches())
def test_plan_equals_reference_restatement(case):
    from metamorph_b200.model.interleave_plan import build_interleave_plan
    ids, mask, labs, n_images, image_len, max_len, side = case
    try:
        ora = R.interleave_reference(ids.tolist(), mask.tolist(), labs.tolist(), n_images, image_len, max_len, side)
    except IndexError:
        with pytest.raises(IndexError):
            build_interleave_plan(ids, mask, labs, n_images, image_len, max_len, side)
        return
    plan = build_interleave_plan(ids, mask, labs, n_images, image_len, max_len, side)
    assert plan.labels.tolist() == ora["labels"]
    assert plan.image_positions.tolist() == ora["image_positions"]
    assert plan.attention_mask.tolist() == ora["mask"]
    assert plan.position_ids.tolist() == ora["position_ids"]
    assert plan.target_image_idx == ora["targets"]
    assert plan.image_placeholder == ora["placeholder"]
    for b, rows in enumerate(ora["rows"]):
        for t, r in enumerate(rows):
            v = int(plan.row_map[b, t])
            want = r[1] if r[0] == "t" else (-(2 + r[1] * image_len + r[2]) if r[0] == "i" else -1)
            assert v == want, (b, t, r, v)
