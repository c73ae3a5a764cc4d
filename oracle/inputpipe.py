"""oracle/inputpipe.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU restatement (numpy) of the image side of ``HandObjSet.get_sample`` for a frame pair
(/root/reference/meshreg/datasets/handobjset.py:336-379): colour jitter of the full source frame, affine crop /
rotation to ``inp_res`` with nearest sampling, ``to_tensor`` + ``normalize(0.5, 1)``, and the jitter mask (a white image
through the same affine transform) -- what SURVEY.md section 8(f) row f3 moves onto the GPU.

Where the arithmetic lives: the reference calls ``libyana.transformutils.handutils.transform_img`` /
``colortrans.apply_jitter`` (libyana@v0.2.0, absent from /root/reference), which are thin wrappers over PIL and
torchvision: ``Image.transform(res, Image.AFFINE, inv(affinetrans)[:2].flatten())`` (PIL's default NEAREST filter) and
``torchvision.transforms.functional.adjust_{brightness,saturation,hue,contrast}`` applied in a random order.
PINNED: tests/test_oracle_inputpipe.py checks every function below bit for bit against PIL / torchvision themselves
(both are in this image), so the GPU kernels are compared with PIL's own results.
NOT restated: the Gaussian blur of handobjset.py:338-339 (PIL's extended-box approximation) -- out of scope of the
kernel; ``get_affine_transform`` below restates libyana's helper as published in hassony2/obman_train (unpinned).

PIL's arithmetic, as restated here:
* AFFINE + NEAREST runs in 16.16 fixed point: every coefficient is rounded to floor(v * 65536 + 0.5), the constant
  terms include the half-pixel offsets, source coordinates are (c + x a + y b) >> 16 (Geometry.c, affine_fixed);
* brightness / contrast / saturation are ``Image.blend(degenerate, image, factor)``: float32
  ``degenerate + factor * (image - degenerate)``, clipped to [0, 255], truncated; the degenerate image is black, the
  rounded mean of the grey-scale image, the grey-scale image (L = (19595 R + 38470 G + 7471 B + 32768) >> 16);
* hue goes through PIL's 8-bit HSV conversion and back (Convert.c rgb2hsv_row / hsv2rgb_row) with a wrapping uint8
  shift of the H channel.
"""
import numpy as np

OP_BRIGHTNESS, OP_SATURATION, OP_HUE, OP_CONTRAST = 0, 1, 2, 3


def fix16(v):
    return int(np.floor(np.float64(v) * 65536.0 + 0.5))


def affine_fixed_coeffs(coef):
    """PIL's six 16.16 integers (a0, a1, a2, a3, a4, a5) for output -> input coefficients (a, b, c, d, e, f)."""
    a, b, c, d, e, f = [np.float64(v) for v in coef]
    return (fix16(a), fix16(b), fix16(c + a * 0.5 + b * 0.5), fix16(d), fix16(e), fix16(f + d * 0.5 + e * 0.5))


def affine_nearest(src, coef, size, fill=0):
    """``Image.fromarray(src).transform(size, Image.AFFINE, coef)`` for an [H,W,C] uint8 array; also returns the mask of
    output pixels whose source exists."""
    w, h = size
    a0, a1, a2, a3, a4, a5 = affine_fixed_coeffs(coef)
    x = np.arange(w, dtype=np.int64)[None, :]
    y = np.arange(h, dtype=np.int64)[:, None]
    xin = (a2 + y * a1 + x * a0) >> 16
    yin = (a5 + y * a4 + x * a3) >> 16
    inside = (xin >= 0) & (xin < src.shape[1]) & (yin >= 0) & (yin < src.shape[0])
    out = np.full((h, w, src.shape[2]), fill, dtype=np.uint8)
    out[inside] = src[yin[inside], xin[inside]]
    return out, inside


def gray_l(rgb):
    r, g, b = [rgb[..., i].astype(np.int64) for i in range(3)]
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.int64)


def _blend(deg, img, factor):
    v = deg.astype(np.float32) + np.float32(factor) * (img.astype(np.float32) - deg.astype(np.float32))
    return np.clip(v, 0, 255).astype(np.uint8)


def adjust_brightness(rgb, factor):
    return _blend(np.zeros_like(rgb), rgb, factor)


def adjust_saturation(rgb, factor):
    return _blend(np.repeat(gray_l(rgb)[..., None], 3, -1), rgb, factor)


def gray_mean(rgb):
    """int(ImageStat.Stat(img.convert('L')).mean[0] + 0.5)"""
    return int(gray_l(rgb).sum() / float(rgb.shape[0] * rgb.shape[1]) + 0.5)


def adjust_contrast(rgb, factor, mean=None):
    mean = gray_mean(rgb) if mean is None else mean
    return _blend(np.full_like(rgb, mean), rgb, factor)


def rgb2hsv(rgb):
    r, g, b = [rgb[..., i].astype(np.int32) for i in range(3)]
    maxc = np.maximum(r, np.maximum(g, b))
    minc = np.minimum(r, np.minimum(g, b))
    cr = (maxc - minc).astype(np.float32)
    with np.errstate(all="ignore"):
        s = cr / maxc.astype(np.float32)
        rc = (maxc - r).astype(np.float32) / cr
        gc = (maxc - g).astype(np.float32) / cr
        bc = (maxc - b).astype(np.float32) / cr
        d = lambda t: t.astype(np.float64)
        h = np.where(r == maxc, bc - gc, np.where(g == maxc, (2.0 + d(rc) - d(bc)).astype(np.float32),
                                                  (4.0 + d(gc) - d(rc)).astype(np.float32)))
        h = np.fmod(d(h) / 6.0 + 1.0, 1.0).astype(np.float32)
        uh = np.clip((d(h) * 255.0).astype(np.int64), 0, 255)
        us = np.clip((d(s) * 255.0).astype(np.int64), 0, 255)
    grey = minc == maxc
    return np.stack([np.where(grey, 0, uh), np.where(grey, 0, us), maxc], -1).astype(np.uint8)


def hsv2rgb(hsv):
    h, s, v = [hsv[..., i].astype(np.int32) for i in range(3)]
    f32 = np.float32
    fh = h.astype(f32) * f32(6.0) / f32(255.0)
    fs = s.astype(f32) / f32(255.0)
    i = np.floor(fh).astype(np.int32)
    f = fh - i.astype(f32)
    vf = v.astype(f32)
    rnd = lambda x: np.floor(x + f32(0.5)).astype(np.int32)
    p = rnd(vf * (f32(1.0) - fs))
    q = rnd(vf * (f32(1.0) - fs * f))
    t = rnd(vf * (f32(1.0) - fs * (f32(1.0) - f)))
    i = i % 6
    out = np.stack([np.choose(i, [v, q, p, p, t, v]), np.choose(i, [t, v, v, q, p, p]), np.choose(i, [p, p, t, v, v, q])], -1)
    out = np.where((s == 0)[..., None], np.stack([v, v, v], -1), out)
    return np.clip(out, 0, 255).astype(np.uint8)


def hue_shift_u8(factor):
    """The wrapping uint8 torchvision adds to the H channel: uint8(hue_factor * 255)."""
    return int(factor * 255) & 0xFF


def adjust_hue(rgb, factor):
    hsv = rgb2hsv(rgb).astype(np.int32)
    hsv[..., 0] = (hsv[..., 0] + hue_shift_u8(factor)) & 0xFF
    return hsv2rgb(hsv.astype(np.uint8))


def apply_jitter(rgb, brightness, saturation, hue, contrast, order):
    """The four adjustments in ``order`` (a permutation of the OP_* ids; colortrans.apply_jitter shuffles them)."""
    for op in order:
        if op == OP_BRIGHTNESS:
            rgb = adjust_brightness(rgb, brightness)
        elif op == OP_SATURATION:
            rgb = adjust_saturation(rgb, saturation)
        elif op == OP_HUE:
            rgb = adjust_hue(rgb, hue)
        elif op == OP_CONTRAST:
            rgb = adjust_contrast(rgb, contrast)
    return rgb


def get_affine_transform(center, scale, res, rot=0.0):
    """libyana.transformutils.handutils.get_affine_transform as published in hassony2/obman_train (UNPINNED: libyana is
    absent): (total affine transform source -> crop, the same without the rotation), ``res`` = (width, height)."""
    rot_mat = np.array([[np.cos(rot), -np.sin(rot), 0], [np.sin(rot), np.cos(rot), 0], [0, 0, 1]], dtype=np.float64)
    c = np.array([center[0], center[1], 1.0])
    origin_rot_center = rot_mat.dot(c)[:2]
    t_mat = np.eye(3)
    t_mat[0, 2] = -res[1] / 2
    t_mat[1, 2] = -res[0] / 2
    t_inv = t_mat.copy()
    t_inv[:2, 2] *= -1
    transformed_center = t_inv.dot(rot_mat).dot(t_mat).dot(c)

    def no_rot(origin):
        m = np.zeros((3, 3))
        m[0, 0] = float(res[1]) / scale
        m[1, 1] = float(res[0]) / scale
        m[0, 2] = res[1] * (-float(origin[0]) / scale + 0.5)
        m[1, 2] = res[0] * (-float(origin[1]) / scale + 0.5)
        m[2, 2] = 1
        return m

    total = no_rot(origin_rot_center).dot(rot_mat)
    return total.astype(np.float32), no_rot(transformed_center[:2]).astype(np.float32)


def transform_coefficients(affinetrans):
    """handutils.transform_img: the inverse of the 3x3 transform, first two rows (PIL's output -> input coefficients)."""
    inv = np.linalg.inv(np.asarray(affinetrans))
    return (inv[0, 0], inv[0, 1], inv[0, 2], inv[1, 0], inv[1, 1], inv[1, 2])


def frame_to_tensors(src_rgb, affinetrans, inp_res, color=None):
    """One frame through handobjset.py:340-379 (without the blur): returns (image [3,H,W] float32 in [-0.5, 0.5],
    jitter_mask [3,H,W] float32).  ``color``: dict(brightness, saturation, hue, contrast, order) or None."""
    img = src_rgb if color is None else apply_jitter(src_rgb, color["brightness"], color["saturation"], color["hue"],
                                                     color["contrast"], color["order"])
    coef = transform_coefficients(affinetrans)
    crop, inside = affine_nearest(img, coef, inp_res)
    image = crop.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0) - np.float32(0.5)
    mask = np.repeat(inside[None].astype(np.float32), 3, 0)   # white image: 255 / 255 = 1 inside, fill 0 outside
    return image, mask
