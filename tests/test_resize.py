"""the resize step of the reference's pre-processing (T.Resize on PIL images = Pillow BILINEAR).
CPU: the numpy restatement (oracle/resize_pil.py) and the product's coefficient tables are pinned against the installed
Pillow.  GPU: yms_resample_u8 is bit-exact against Pillow; the uint8 pipeline feeds the model."""
import numpy as np
import pytest
import torch
from PIL import Image

from oracle.resize_pil import precompute_coeffs, resize_bilinear_u8
from yolo_ms_b200.preprocess import bilinear_coeffs

CASES = [(137, 138, 640, 640), (480, 640, 640, 640), (1080, 1920, 640, 640), (33, 47, 64, 96), (640, 640, 320, 320),
         (700, 500, 640, 640), (5, 7, 3, 2), (64, 64, 64, 64), (100, 640, 640, 640), (640, 100, 640, 640)]


def _img(h, w, seed):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def _pil(img, oh, ow):
    return np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))


@pytest.mark.parametrize("h,w,oh,ow", CASES)
def test_oracle_matches_pillow(h, w, oh, ow):
    img = _img(h, w, h * 1000 + w)
    assert np.array_equal(resize_bilinear_u8(img, oh, ow), _pil(img, oh, ow))


def test_oracle_matches_pillow_on_reference_sample():
    # the only image the reference ships is a 137x138 RGBA png; same geometry, RGB
    img = _img(137, 138, 5)
    assert np.array_equal(resize_bilinear_u8(img, 640, 640), _pil(img, 640, 640))


@pytest.mark.parametrize("a,b", [(137, 640), (1920, 640), (640, 320), (47, 96), (7, 2), (480, 640), (640, 641)])
def test_product_coefficients_match_oracle(a, b):
    k1, b1, c1 = bilinear_coeffs(a, b)
    k2, b2, c2 = precompute_coeffs(a, b)
    assert k1 == k2 and np.array_equal(b1, b2) and np.array_equal(c1, c2)
    assert (c1.sum(1) - (1 << 22)).__abs__().max() <= c1.shape[1]          # weights sum to one (22-bit fixed point)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,oh,ow", CASES)
def test_gpu_resize_is_bit_exact_vs_pillow(h, w, oh, ow):
    from yolo_ms_b200.preprocess import resize_u8
    img = _img(h, w, 7 * h + w)
    out = torch.empty((oh, ow, 3), dtype=torch.uint8, device="cuda")
    resize_u8(torch.from_numpy(img).cuda(), out)
    assert np.array_equal(out.cpu().numpy(), _pil(img, oh, ow))


@pytest.mark.gpu
def test_gpu_preprocess_batch_writes_batch_slices():
    from yolo_ms_b200.preprocess import preprocess_batch
    imgs = [_img(90, 120, 1), _img(200, 64, 2), _img(64, 96, 3)]
    batch = preprocess_batch(imgs, (64, 96))
    assert batch.shape == (3, 64, 96, 3) and batch.dtype == torch.uint8
    for i, im in enumerate(imgs):
        assert np.array_equal(batch[i].cpu().numpy(), _pil(im, 64, 96))


@pytest.mark.gpu
def test_resize_rejects_cpu_tensors():
    from yolo_ms_b200 import YmsError
    from yolo_ms_b200.preprocess import resize_u8
    with pytest.raises(YmsError):
        resize_u8(torch.zeros(4, 4, 3, dtype=torch.uint8), torch.zeros(2, 2, 3, dtype=torch.uint8))
