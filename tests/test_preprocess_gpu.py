"""GPU: csrc/preprocess.cu through the C ABI against the CPU oracle, bit for bit (integer / table arithmetic)."""
import numpy as np
import pytest
import torch

from oracle import preprocess as op

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("i", range(len(op.GOLDEN_CASES)))
def test_gpu_preprocess_bit_exact(cuda_device, i):
    from metamorph_b200.preprocess import SiglipGpuImageProcessor
    h, w, seed = op.GOLDEN_CASES[i]
    img = op.synthetic_image(h, w, seed)
    want = op.siglip_preprocess(img, pad=True)
    proc = SiglipGpuImageProcessor(device="cuda")
    got = proc.preprocess(img, return_tensors="pt")["pixel_values"]
    assert got.shape == (1, 3, 384, 384) and got.dtype == torch.float32 and got.is_cuda
    np.testing.assert_array_equal(got[0].cpu().numpy(), want)
    nopad = SiglipGpuImageProcessor(device="cuda", pad_to_square=False)(torch.from_numpy(img).cuda())["pixel_values"]
    np.testing.assert_array_equal(nopad[0].cpu().numpy(), op.siglip_preprocess(img, pad=False))


def test_gpu_preprocess_batch_bf16_and_pipeline(cuda_device):
    from metamorph_b200.preprocess import ImageBatchPipeline, SiglipGpuImageProcessor
    imgs = [op.synthetic_image(h, w, s) for h, w, s in [(480, 640, 11), (333, 222, 12), (384, 384, 13), (900, 1200, 14)]]
    want = np.stack([op.siglip_preprocess(im) for im in imgs])
    proc16 = SiglipGpuImageProcessor(device="cuda", out_dtype=torch.bfloat16)
    got16 = proc16.preprocess(imgs)["pixel_values"]
    assert got16.dtype == torch.bfloat16
    assert torch.equal(got16.cpu(), torch.from_numpy(want).bfloat16())      # one rounding of the exact fp32 value
    proc = SiglipGpuImageProcessor(device="cuda")
    pipe = ImageBatchPipeline(proc)
    pipe.submit(imgs)
    first = pipe.result()
    pipe.submit(list(reversed(imgs)))                                       # second slot while `first` is still in use
    second = pipe.result()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(first.cpu().numpy(), want)
    np.testing.assert_array_equal(second.cpu().numpy(), want[::-1])
    assert pipe.h2d_bytes == sum(im.size for im in imgs)


def test_gpu_preprocess_accepts_pil(cuda_device):
    Image = pytest.importorskip("PIL.Image")
    from metamorph_b200.preprocess import SiglipGpuImageProcessor
    img = op.synthetic_image(211, 317, 21)
    got = SiglipGpuImageProcessor(device="cuda").preprocess(Image.fromarray(img))["pixel_values"][0]
    np.testing.assert_array_equal(got.cpu().numpy(), op.siglip_preprocess(img))
