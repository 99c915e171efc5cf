"""CPU: image pre-processing oracle (oracle/preprocess.py) against the fixtures generated from Pillow + the HF PIL
SigLIP processor (tests/golden/preprocess.npz), against Pillow itself when importable, and the host-side coefficient
tables of the C ABI (mm_resize_coeff_build needs no GPU) against the oracle, bit for bit."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import preprocess as op

GOLD = os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz")


def _sha(a: np.ndarray) -> bytes:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_fixture_cases_match_generator(gold):
    assert gold["cases"].tolist() == [list(c) for c in op.GOLDEN_CASES]


@pytest.mark.parametrize("i", range(len(op.GOLDEN_CASES)))
def test_oracle_matches_reference_fixture(gold, i):
    h, w, seed = op.GOLDEN_CASES[i]
    img = op.synthetic_image(h, w, seed)
    assert _sha(img) == gold[f"sha_input_{i}"].tobytes(), "synthetic input differs from the one the fixture was made from"
    padded = op.siglip_preprocess(img, pad=True)
    assert padded.dtype == np.float32 and padded.shape == (3, 384, 384)
    np.testing.assert_array_equal(padded[:, 190:194, :], gold[f"rows_padded_f32_{i}"])
    assert _sha(padded) == gold[f"sha_padded_f32_{i}"].tobytes()          # bit-exact vs expand2square + HF processor
    assert _sha(op.pil_bicubic_resize_u8(img, 384, 384)) == gold[f"sha_plain_u8_{i}"].tobytes()   # vs bare Pillow


def test_oracle_matches_pillow_live():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(123)
    for h, w, oh, ow in [(37, 91, 384, 384), (640, 480, 384, 384), (384, 384, 384, 384), (200, 300, 64, 48),
                         (5, 7, 384, 384), (1000, 333, 384, 384)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), resample=Image.BICUBIC))
        np.testing.assert_array_equal(op.pil_bicubic_resize_u8(img, oh, ow), ref)


@pytest.mark.parametrize("sizes", [(384, 384), (50, 384), (1200, 384), (385, 384), (1031, 384), (1, 384), (768, 384),
                                   (500, 64)])
def test_abi_coefficient_table_matches_oracle(sizes):
    from metamorph_b200.preprocess import build_resize_coeffs
    in_size, out_size = sizes
    host, ksize = build_resize_coeffs(in_size, out_size)
    ks, bounds, kk = op.precompute_coeffs(in_size, out_size)
    assert ksize == ks
    words = host.view(torch.int32).numpy()
    assert words[:3].tolist() == [in_size, out_size, ks]
    got_bounds = words[4:4 + 2 * out_size].reshape(out_size, 2)
    got_kk = words[4 + 2 * out_size:].reshape(out_size, ks)
    np.testing.assert_array_equal(got_bounds, bounds)
    np.testing.assert_array_equal(got_kk, kk)


def test_normalize_table_is_the_reference_arithmetic():
    from metamorph_b200.preprocess import normalize_lut
    np.testing.assert_array_equal(normalize_lut().numpy(), op.normalize_lut())
    assert float(normalize_lut()[0]) == -1.0 and float(normalize_lut()[255]) == 1.0


def test_product_rejects_cpu_execution():
    from metamorph_b200._lib import MetaMorphB200Error
    from metamorph_b200.preprocess import _as_hwc_u8
    with pytest.raises(MetaMorphB200Error):
        _as_hwc_u8(np.zeros((4, 4), dtype=np.uint8))
    with pytest.raises(MetaMorphB200Error):
        _as_hwc_u8(np.zeros((4, 4, 3), dtype=np.float32))
