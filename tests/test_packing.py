"""CPU: sequence packing of an interleave plan is a pure re-layout of the padded plan."""
import numpy as np
import pytest
import torch

from metamorph_b200.constants import IGNORE_INDEX
from metamorph_b200.model.interleave_plan import ROW_PAD, build_interleave_plan, pack_plan
from oracle.weights import TINY, make_batch


def _plan(max_len=4096):
    ids, mask, labs, images = make_batch(TINY)
    return build_interleave_plan(ids, mask, labs, images.shape[0], TINY["image_tokens"], max_len)


@pytest.mark.parametrize("pack_len", [None, 160, 300])
def test_pack_plan_is_a_relayout(pack_len):
    plan = _plan()
    lens = plan.seqlens.tolist()
    if pack_len is not None and max(lens) > pack_len:
        with pytest.raises(ValueError):
            pack_plan(plan, pack_len)
        return
    packed = pack_plan(plan, pack_len)
    Tp = pack_len or plan.seq_len
    assert packed.seq_len == Tp and packed.segments is not None
    flat = [(r, off, n) for r, segs in enumerate(packed.segments) for off, n in segs]
    assert [n for _, _, n in flat] == [n for n in lens if n > 0]            # sample order preserved (next-fit)
    for b, (r, off, n) in enumerate(flat):
        np.testing.assert_array_equal(packed.row_map[r, off:off + n].numpy(), plan.row_map[b, :n].numpy())
        np.testing.assert_array_equal(packed.position_ids[r, off:off + n].numpy(), plan.position_ids[b, :n].numpy())
        np.testing.assert_array_equal(packed.image_positions[r, off:off + n].numpy(), plan.image_positions[b, :n].numpy())
        want = plan.labels[b, :n].clone()
        want[0] = IGNORE_INDEX                                              # dropped by the shift in the padded layout
        np.testing.assert_array_equal(packed.labels[r, off:off + n].numpy(), want.numpy())
    for r, segs in enumerate(packed.segments):
        used = sum(n for _, n in segs)
        assert int(packed.seqlens[r]) == used <= Tp
        assert [off for off, _ in segs] == list(np.cumsum([0] + [n for _, n in segs[:-1]]))
        assert bool(packed.attention_mask[r, :used].all()) and not bool(packed.attention_mask[r, used:].any())
        assert (packed.row_map[r, used:] == ROW_PAD).all() and (packed.labels[r, used:] == IGNORE_INDEX).all()
    # the shifted targets (what the loss sees) are the same multiset of (token, target) pairs
    def pairs(p):
        out = []
        for r in range(p.batch):
            for t in range(p.seq_len - 1):
                if int(p.labels[r, t + 1]) != IGNORE_INDEX:
                    out.append((int(p.row_map[r, t]), int(p.position_ids[r, t]), int(p.labels[r, t + 1])))
        return sorted(out)
    assert pairs(packed) == pairs(plan)
    assert packed.target_image_idx == plan.target_image_idx
    assert packed.batch * packed.seq_len <= plan.batch * plan.seq_len or pack_len is not None


def test_segment_tables_cover_every_tile_once_heaviest_first():
    """Work lists of the one-launch packed attention (ops.SegmentTables): each (sequence, 128-row tile) appears exactly
    once in the query list (forward / dQ) and in the key list (dK/dV), ordered by the number of tiles it will visit."""
    from metamorph_b200.ops import SegmentTables
    segs = [(0, 300), (300, 1), (301, 128), (429, 1501), (1930, 129)]
    t = SegmentTables(segs, "cpu")
    assert t.n_seg == 5 and t.max_len == 1501 and t.total == 300 + 1 + 128 + 1501 + 129
    assert t.start.tolist() == [a for a, _ in segs] and t.length.tolist() == [n for _, n in segs]
    want = {(s, j) for s, (_, n) in enumerate(segs) for j in range((n + 127) // 128)}
    wq = t.work_q.view(-1, 2).tolist()
    wk = t.work_k.view(-1, 2).tolist()
    assert t.n_work_q == len(wq) == len(want) and t.n_work_k == len(wk) == len(want)
    assert {tuple(x) for x in wq} == want and {tuple(x) for x in wk} == want
    ntile = [(n + 127) // 128 for _, n in segs]
    cost_q = [j + 1 for s, j in wq]                     # a causal query tile attends to j+1 key tiles
    cost_k = [ntile[s] - j for s, j in wk]              # a key tile is visited by the query tiles at or after it
    assert cost_q == sorted(cost_q, reverse=True) and cost_k == sorted(cost_k, reverse=True)
