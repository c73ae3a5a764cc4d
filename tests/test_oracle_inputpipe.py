"""CPU: oracle/inputpipe.py pinned against PIL and torchvision themselves -- the libraries the reference's dataset code
runs (handobjset.py:336-379 through libyana's thin wrappers): affine crop with nearest sampling, the four colour
adjustments, to_tensor + normalize, the jitter mask.  Bit for bit."""
import numpy as np
import pytest
import torch

from oracle import inputpipe as oip

PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402
from torchvision.transforms import functional as F  # noqa: E402


def _img(rng, h, w):
    a = (rng.random((h, w, 3)) * 255).astype(np.uint8)
    a[0, :6] = [[10, 10, 10], [255, 0, 0], [0, 255, 0], [0, 0, 255], [0, 0, 0], [255, 255, 255]]
    return a


def test_affine_nearest_matches_pil():
    rng = np.random.default_rng(0)
    for _ in range(25):
        a = _img(rng, int(rng.integers(20, 70)), int(rng.integers(20, 90)))
        th, sc = rng.uniform(-3.1, 3.1), rng.uniform(0.4, 2.5)
        coef = (sc * np.cos(th), -sc * np.sin(th), rng.uniform(-20, 40), sc * np.sin(th), sc * np.cos(th), rng.uniform(-20, 40))
        size = (int(rng.integers(8, 64)), int(rng.integers(8, 64)))
        ref = np.asarray(Image.fromarray(a).transform(size, Image.AFFINE, coef))
        got, inside = oip.affine_nearest(a, coef, size)
        np.testing.assert_array_equal(got, ref)
        white = np.asarray(Image.new("RGB", (a.shape[1], a.shape[0]), (255, 255, 255)).transform(size, Image.AFFINE, coef))
        np.testing.assert_array_equal(inside, white[..., 0] == 255)


@pytest.mark.parametrize("factor", [0.5, 0.83, 1.0, 1.27, 1.5])
def test_colour_adjustments_match_torchvision_on_pil(factor):
    rng = np.random.default_rng(1)
    a = _img(rng, 37, 53)
    im = Image.fromarray(a)
    np.testing.assert_array_equal(oip.adjust_brightness(a, factor), np.asarray(F.adjust_brightness(im, factor)))
    np.testing.assert_array_equal(oip.adjust_saturation(a, factor), np.asarray(F.adjust_saturation(im, factor)))
    np.testing.assert_array_equal(oip.adjust_contrast(a, factor), np.asarray(F.adjust_contrast(im, factor)))
    hue = factor - 1.0  # in [-0.5, 0.5]
    np.testing.assert_array_equal(oip.adjust_hue(a, hue), np.asarray(F.adjust_hue(im, hue)))


def test_hsv_round_trip_matches_pil():
    rng = np.random.default_rng(2)
    a = _img(rng, 64, 64)
    hsv = np.asarray(Image.fromarray(a).convert("HSV"))
    np.testing.assert_array_equal(oip.rgb2hsv(a), hsv)
    np.testing.assert_array_equal(oip.hsv2rgb(hsv), np.asarray(Image.fromarray(hsv, "HSV").convert("RGB")))


def test_whole_frame_matches_the_reference_sequence():
    """handobjset.py:340-379 with PIL / torchvision, every op order the shuffle can produce a prefix of."""
    rng = np.random.default_rng(3)
    a = _img(rng, 54, 96)
    res = (48, 32)
    affine, _ = oip.get_affine_transform((50.0, 25.0), 60.0, res, rot=0.4)
    funcs = {oip.OP_BRIGHTNESS: lambda im, c: F.adjust_brightness(im, c["brightness"]),
             oip.OP_SATURATION: lambda im, c: F.adjust_saturation(im, c["saturation"]),
             oip.OP_HUE: lambda im, c: F.adjust_hue(im, c["hue"]),
             oip.OP_CONTRAST: lambda im, c: F.adjust_contrast(im, c["contrast"])}
    for order in ([0, 1, 2, 3], [3, 2, 1, 0], [2, 0, 3, 1], [1, 3, 0, 2]):
        color = dict(brightness=1.2, saturation=0.7, hue=-0.08, contrast=1.35, order=order)
        im = Image.fromarray(a)
        for op in order:
            im = funcs[op](im, color)
        coef = oip.transform_coefficients(affine)
        white = Image.new("RGB", im.size, (255, 255, 255))
        crop = im.transform(res, Image.AFFINE, coef).crop((0, 0, res[0], res[1]))
        ref_img = F.normalize(F.to_tensor(crop).float(), [0.5, 0.5, 0.5], [1, 1, 1])
        ref_mask = F.to_tensor(white.transform(res, Image.AFFINE, coef).crop((0, 0, res[0], res[1]))).float()
        img, mask = oip.frame_to_tensors(a, affine, res, color)
        assert torch.equal(torch.from_numpy(img), ref_img)
        assert torch.equal(torch.from_numpy(mask), ref_mask)
    img, mask = oip.frame_to_tensors(a, affine, res, None)
    crop = Image.fromarray(a).transform(res, Image.AFFINE, oip.transform_coefficients(affine))
    assert torch.equal(torch.from_numpy(img), F.normalize(F.to_tensor(crop).float(), [0.5] * 3, [1] * 3))
    assert 0.2 < mask.mean() < 1.0


def test_batched_coefficients_equal_the_per_sample_computation():
    """inputpipe.affine_fixed_coefficients inverts the whole batch at once; the reference inverts one matrix at a time
    (handutils.transform_img): same integers."""
    from handobjectconsist_b200 import inputpipe
    rng = np.random.default_rng(4)
    aff = np.stack([oip.get_affine_transform((rng.uniform(100, 300), rng.uniform(50, 200)), rng.uniform(100, 400),
                                             (256, 256), rot=rng.uniform(-3, 3))[0] for _ in range(64)])
    got = inputpipe.affine_fixed_coefficients(aff).numpy()
    ref = np.array([oip.affine_fixed_coeffs(oip.transform_coefficients(a)) for a in aff])
    np.testing.assert_array_equal(got, ref)
    a1, a2 = inputpipe.get_affine_transform((120.0, 80.0), 200.0, (256, 192), 0.3), oip.get_affine_transform((120.0, 80.0), 200.0, (256, 192), 0.3)
    np.testing.assert_array_equal(a1[0], a2[0])
    np.testing.assert_array_equal(a1[1], a2[1])
