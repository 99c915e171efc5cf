"""Generates tests/golden/preprocess.npz from the REAL dependencies the reference calls (run in the build container):
Pillow's Image.resize(BICUBIC) and the HF PIL-backend SigLIP image processor, driven exactly like
metamorph/train/train.py:1189-1209 (expand2square to int(mean*255), then processor.preprocess). The inputs are the
deterministic images of oracle.preprocess.synthetic_image, so only digests + a few rows are stored.

    python -m oracle.make_golden_preprocess
"""
import hashlib
import os

import numpy as np
from PIL import Image

from oracle.preprocess import GOLDEN_CASES, synthetic_image


def expand2square(pil_img, background_color):          # behaviour of train.py:1191-1202
    width, height = pil_img.size
    if width == height:
        return pil_img
    side = max(width, height)
    result = Image.new(pil_img.mode, (side, side), background_color)
    if width > height:
        result.paste(pil_img, (0, (width - height) // 2))
    else:
        result.paste(pil_img, ((height - width) // 2, 0))
    return result


def main():
    import PIL
    import transformers
    from transformers.models.siglip.image_processing_pil_siglip import SiglipImageProcessorPil
    proc = SiglipImageProcessorPil(do_resize=True, size={"height": 384, "width": 384}, resample=3, do_rescale=True,
                                   rescale_factor=1 / 255, do_normalize=True, image_mean=[0.5] * 3, image_std=[0.5] * 3)
    out = {"cases": np.array(GOLDEN_CASES, dtype=np.int64),
           "versions": np.array([f"Pillow {PIL.__version__}", f"transformers {transformers.__version__} (PIL backend)"])}
    for i, (h, w, seed) in enumerate(GOLDEN_CASES):
        img = synthetic_image(h, w, seed)
        pil = Image.fromarray(img)
        sq = expand2square(pil, tuple(int(x * 255) for x in proc.image_mean))
        px = proc.preprocess(sq, return_tensors="pt")["pixel_values"][0].numpy()
        assert px.dtype == np.float32 and px.shape == (3, 384, 384)
        plain = np.asarray(pil.resize((384, 384), resample=Image.BICUBIC))          # no padding: bare Pillow resize
        out[f"sha_padded_f32_{i}"] = np.frombuffer(hashlib.sha256(px.tobytes()).digest(), dtype=np.uint8)
        out[f"sha_plain_u8_{i}"] = np.frombuffer(hashlib.sha256(plain.tobytes()).digest(), dtype=np.uint8)
        out[f"rows_padded_f32_{i}"] = px[:, 190:194, :].copy()                      # 4 centre rows of every channel
        out[f"sha_input_{i}"] = np.frombuffer(hashlib.sha256(img.tobytes()).digest(), dtype=np.uint8)
    path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "preprocess.npz")
    np.savez_compressed(path, **out)
    print("wrote", os.path.abspath(path), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
