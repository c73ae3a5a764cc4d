"""Differentiable rasterization -- same surface as /root/reference/meshreg/neurender/rasterize.py.

``RasterizeFunction`` (rasterize.py:16-315), ``Rasterize`` (:318-359) and ``rasterize_rgbad``
(:362-448) keep the reference's names, argument order, defaults, return tuple / dict and error
behaviour (CUDA tensors only).  The five ``neural_renderer.cuda.rasterize`` entry points the
reference wraps are replaced by two C-ABI calls into libhoc_b200.so: ``hoc_raster_forward`` and
``hoc_raster_backward`` (include/hoc_b200.h).

Differences that do not change results:
  * nothing but ``faces``, ``textures``, ``face_index_map`` and ``rgb`` is saved for backward --
    the kernels recompute weights / depth / taps bit-identically instead of storing
    ``sampling_index_map``, ``sampling_weight_map`` (64 B/px) and reading ``face_inv_map``;
  * ``rasterize_rgbad`` asks the kernel to write ``rgb`` as NCHW with rows already flipped
    (HOC_LAYOUT_IMAGE), so the reference's permute + three list-index gathers (rasterize.py:417-428)
    and their ``index_put`` backward never run;
  * the geometry gradient is skipped when ``faces`` does not require grad and the texture gradient
    when ``textures`` does not (the reference computes and discards them).
"""
import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from .. import _lib

DEFAULT_IMAGE_SIZE = 256
DEFAULT_ANTI_ALIASING = True
DEFAULT_NEAR = 0.1
DEFAULT_FAR = 100
DEFAULT_EPS = 1e-4
DEFAULT_BACKGROUND_COLOR = (0, 0, 0)


def _background_args(background_color, batch_size, device):
    """(host float[3] pointer, device [B,3] tensor or None) from the reference's argument."""
    if background_color is None:  # rasterize_silhouettes / rasterize_depth pass None (rasterize.py:505,534)
        background_color = DEFAULT_BACKGROUND_COLOR
    bg = torch.as_tensor(background_color, dtype=torch.float32)
    if bg.dim() == 1:
        if bg.numel() != 3:
            raise ValueError("background_color needs 3 values")
        host = (ctypes.c_float * 3)(*[float(v) for v in bg.tolist()])
        return host, None
    if bg.dim() == 2:
        if bg.shape != (batch_size, 3):
            raise ValueError(f"per-sample background_color must be [{batch_size}, 3], got {tuple(bg.shape)}")
        host = (ctypes.c_float * 3)(0.0, 0.0, 0.0)
        return host, bg.to(device).contiguous()
    raise ValueError("background_color must have 1 or 2 dimensions")


def _forward_impl(ctx, faces, textures, image_size, near, far, eps, background_color, return_rgb, return_alpha,
                  return_depth, layout, want_face_inv, want_weight):
    _lib.require_cuda(faces, textures if return_rgb else None, what="Rasterize")
    L = _lib.lib()
    if faces.dim() != 4 or faces.shape[2:] != (3, 3):
        raise ValueError(f"faces must be [B, F, 3, 3], got {tuple(faces.shape)}")
    faces_c = faces.detach().contiguous().float()
    B, Fn = faces_c.shape[:2]
    S = int(image_size)
    dev = faces_c.device
    ts = 0
    tex_c = None
    if return_rgb:
        if textures is None:
            raise ValueError("return_rgb=True needs textures")
        tex_c = textures.detach().contiguous().float()
        if tex_c.dim() != 6 or tex_c.shape[0] != B or tex_c.shape[1] != Fn or tex_c.shape[-1] != 3:
            raise ValueError(f"textures must be [B, F, ts, ts, ts, 3], got {tuple(tex_c.shape)}")
        ts = int(tex_c.shape[2])

    with torch.cuda.device(dev):
        face_index_map = torch.empty((B, S, S), dtype=torch.int32, device=dev)
        weight_map = torch.empty((B, S, S, 3), dtype=torch.float32, device=dev) if want_weight else None
        rgb = alpha = depth = None
        if return_rgb:
            shape = (B, S, S, 3) if layout == _lib.HOC_LAYOUT_RAW else (B, 3, S, S)
            rgb = torch.empty(shape, dtype=torch.float32, device=dev)
        if return_alpha:
            alpha = torch.empty((B, S, S), dtype=torch.float32, device=dev)
        if return_depth:
            depth = torch.empty((B, S, S), dtype=torch.float32, device=dev)
        if return_depth and want_face_inv:
            face_inv_map = torch.empty((B, S, S, 3, 3), dtype=torch.float32, device=dev)
            inv_arg = face_inv_map
        else:
            face_inv_map = torch.zeros(1, dtype=torch.float32, device=dev)  # rasterize.py:84 dummy
            inv_arg = None
        ws_bytes = L.hoc_raster_forward_workspace_bytes(B, Fn, S)
        ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
        bg_host, bg_dev = _background_args(background_color, B, dev)
        code = L.hoc_raster_forward(
            _lib.ptr(faces_c), _lib.ptr(tex_c), B, Fn, S, ts, float(near), float(far), float(eps), bg_host,
            _lib.ptr(bg_dev), layout, _lib.ptr(rgb), _lib.ptr(alpha), _lib.ptr(depth), _lib.ptr(face_index_map),
            _lib.ptr(weight_map), _lib.ptr(inv_arg), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
        _lib.check(code, "hoc_raster_forward")

    ctx.image_size, ctx.near, ctx.far, ctx.eps = S, float(near), float(far), float(eps)
    ctx.return_rgb, ctx.return_alpha, ctx.return_depth = return_rgb, return_alpha, return_depth
    ctx.layout, ctx.texture_size = layout, ts
    ctx.batch_size, ctx.num_faces = B, Fn
    ctx.save_for_backward(faces_c, tex_c if return_rgb else None, face_index_map, rgb,
                          weight_map if (want_weight and return_depth) else None,
                          depth if (want_weight and return_depth) else None)
    ctx.mark_non_differentiable(face_index_map)
    ctx.set_materialize_grads(False)

    # disabled outputs are empty CPU tensors like the reference's (rasterize.py:118); one object per output
    return (rgb if return_rgb else torch.tensor([]), alpha if return_alpha else torch.tensor([]),
            depth if return_depth else torch.tensor([]), face_index_map, face_inv_map,
            weight_map if want_weight else torch.zeros(1, device=dev))


def _backward_impl(ctx, grad_rgb, grad_alpha, grad_depth):
    faces, textures, face_index_map, rgb, weight_map, depth = ctx.saved_tensors
    L = _lib.lib()
    B, Fn, S = ctx.batch_size, ctx.num_faces, ctx.image_size
    dev = faces.device
    need_faces = ctx.needs_input_grad[0]
    need_tex = ctx.return_rgb and ctx.needs_input_grad[1]
    if not (need_faces or need_tex):
        return None, None

    def prep(g, enabled):
        if not enabled or g is None:
            return None
        return g.contiguous().float()

    g_rgb = prep(grad_rgb, ctx.return_rgb)
    g_alpha = prep(grad_alpha, ctx.return_alpha)
    g_depth = prep(grad_depth, ctx.return_depth)
    with torch.cuda.device(dev):
        grad_faces = torch.empty_like(faces) if need_faces else None
        grad_textures = torch.empty_like(textures) if need_tex else None
        ws_bytes = L.hoc_raster_backward_workspace_bytes_ex(B, Fn, S, max(int(ctx.texture_size), 1), _lib.HOC_TEX_GRAD_CUBE)
        ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
        code = L.hoc_raster_backward(
            _lib.ptr(faces), _lib.ptr(textures), _lib.ptr(face_index_map), _lib.ptr(rgb), _lib.ptr(weight_map),
            _lib.ptr(depth), _lib.ptr(g_rgb), _lib.ptr(g_alpha), _lib.ptr(g_depth), B, Fn, S, ctx.texture_size, ctx.near, ctx.far, ctx.eps, ctx.layout,
            int(ctx.return_alpha), _lib.HOC_TEX_GRAD_CUBE, _lib.ptr(grad_faces), _lib.ptr(grad_textures), _lib.ptr(ws),
            ws_bytes,
            _lib.stream_ptr())
        _lib.check(code, "hoc_raster_backward")
    return grad_faces, grad_textures


class RasterizeFunction(Function):
    """
    Definition of differentiable rasterize operation (drop-in for rasterize.py:16-315).
    Implemented in CUDA (sm_100a); only for cuda Tensors.
    """

    @staticmethod
    def forward(ctx, faces, textures, image_size, near, far, eps, background_color, return_rgb=False,
                return_alpha=False, return_depth=False):
        return _forward_impl(ctx, faces, textures, image_size, near, far, eps, background_color, return_rgb,
                             return_alpha, return_depth, _lib.HOC_LAYOUT_RAW, True, True)

    @staticmethod
    def backward(ctx, grad_rgb_map, grad_alpha_map, grad_depth_map, grad_face_index_map, grad_face_inv_map,
                 grad_weight_map):
        grad_faces, grad_textures = _backward_impl(ctx, grad_rgb_map, grad_alpha_map, grad_depth_map)
        return grad_faces, grad_textures, None, None, None, None, None, None, None, None


class _RasterizeImageFunction(Function):
    """Same operator, outputs written directly as rasterize_rgbad returns them (NCHW rgb, rows flipped)."""

    @staticmethod
    def forward(ctx, faces, textures, image_size, near, far, eps, background_color, return_rgb, return_alpha,
                return_depth, return_face_inv_map, return_weight_map):
        return _forward_impl(ctx, faces, textures, image_size, near, far, eps, background_color, return_rgb,
                             return_alpha, return_depth, _lib.HOC_LAYOUT_IMAGE, return_face_inv_map,
                             return_weight_map)

    @staticmethod
    def backward(ctx, grad_rgb, grad_alpha, grad_depth, grad_face_index_map, grad_face_inv_map, grad_weight_map):
        grad_faces, grad_textures = _backward_impl(ctx, grad_rgb, grad_alpha, grad_depth)
        return (grad_faces, grad_textures) + (None,) * 10


class Rasterize(nn.Module):
    """
    Wrapper around the autograd function RasterizeFunction (rasterize.py:318-359).
    Currently implemented only for cuda Tensors
    """

    def __init__(self, image_size, near, far, eps, background_color, return_rgb=False, return_alpha=False,
                 return_depth=False):
        super(Rasterize, self).__init__()
        self.image_size = image_size
        self.near = near
        self.far = far
        self.eps = eps
        self.background_color = background_color
        self.return_rgb = return_rgb
        self.return_alpha = return_alpha
        self.return_depth = return_depth

    def forward(self, faces, textures):
        if not faces.is_cuda or (textures is not None and not textures.is_cuda):
            raise TypeError("Rasterize module supports only cuda Tensors")
        return RasterizeFunction.apply(faces, textures, self.image_size, self.near, self.far, self.eps,
                                       self.background_color, self.return_rgb, self.return_alpha, self.return_depth)


def rasterize_rgbad(faces, textures=None, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING,
                    near=DEFAULT_NEAR, far=DEFAULT_FAR, eps=DEFAULT_EPS, background_color=DEFAULT_BACKGROUND_COLOR,
                    return_rgb=True, return_alpha=True, return_depth=True, return_face_inv_map=True,
                    return_weight_map=True):
    """
    Generate RGB, alpha channel, and depth images from faces and textures (rasterize.py:362-448).

    Returns the reference's dict: 'rgb' [B,3,S,S], 'alpha' [B,S,S], 'depth' [B,S,S] (rows flipped so
    that row 0 is the image top), plus 'face_inv_map', 'face_index_map', 'weight_map' in raster row
    order (NOT flipped, like the reference).  ``return_face_inv_map`` / ``return_weight_map`` are
    extensions: callers that never read those maps (get_opticalflow) can skip 36 + 12 B/px of writes.
    """
    if not faces.is_cuda or (textures is not None and not textures.is_cuda):
        raise TypeError("Rasterize module supports only cuda Tensors")
    size = image_size * 2 if anti_aliasing else image_size
    rgb, alpha, depth, face_index_map, face_inv_map, weight_map = _RasterizeImageFunction.apply(
        faces, textures, size, near, far, eps, background_color, return_rgb, return_alpha, return_depth,
        return_face_inv_map, return_weight_map)

    if anti_aliasing:
        # 0.5x down-sampling (row flip and pooling commute)
        if return_rgb:
            rgb = F.avg_pool2d(rgb, kernel_size=(2, 2))
        if return_alpha:
            alpha = F.avg_pool2d(alpha[:, None, :, :], kernel_size=(2, 2))[:, 0]
        if return_depth:
            depth = F.avg_pool2d(depth[:, None, :, :], kernel_size=(2, 2))[:, 0]

    return {
        "rgb": rgb if return_rgb else None,
        "alpha": alpha if return_alpha else None,
        "depth": depth if return_depth else None,
        "face_inv_map": face_inv_map,
        "face_index_map": face_index_map,
        "weight_map": weight_map,
    }


def rasterize(faces, textures, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING, near=DEFAULT_NEAR,
              far=DEFAULT_FAR, eps=DEFAULT_EPS, background_color=DEFAULT_BACKGROUND_COLOR):
    """RGB images [B,3,S,S] from faces and textures (rasterize.py:449-478)."""
    return rasterize_rgbad(faces, textures, image_size, anti_aliasing, near, far, eps, background_color, True, False,
                           False)["rgb"]


def rasterize_silhouettes(faces, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING,
                          near=DEFAULT_NEAR, far=DEFAULT_FAR, eps=DEFAULT_EPS):
    """Alpha channels [B,S,S] from faces (rasterize.py:481-507)."""
    return rasterize_rgbad(faces, None, image_size, anti_aliasing, near, far, eps, None, False, True, False)["alpha"]


def rasterize_depth(faces, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING, near=DEFAULT_NEAR,
                    far=DEFAULT_FAR, eps=DEFAULT_EPS):
    """Depth images [B,S,S] from faces (rasterize.py:510-536)."""
    return rasterize_rgbad(faces, None, image_size, anti_aliasing, near, far, eps, None, False, False, True)["depth"]
