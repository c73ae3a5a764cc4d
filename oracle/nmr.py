"""oracle/nmr.py -- TEST INFRASTRUCTURE ONLY (checker; never imported by the product path).

numpy front-end of the C restatement in ``nmr_oracle.c``.  It follows the *wrapper* semantics of
/root/reference/meshreg/neurender/rasterize.py line by line:

* buffer shapes / initial values ............ rasterize.py:58-85  (``-1`` / ``0`` / ``far``)
* forward = K1+K2, K3, background, alpha .... rasterize.py:87-104, 246-260
* backward = K4, K5, K6 ..................... rasterize.py:128-197
* ``rasterize_rgbad`` (NCHW + row flip + 2x SSAA) rasterize.py:362-448

PARITY UNPINNED for the kernel arithmetic (see nmr_oracle_impl.h).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import
this module.
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile the checker library with gcc (no GPU, no reference sources needed)."""
    so = os.path.join(_HERE, "libnmr_oracle.so")
    srcs = [os.path.join(_HERE, n) for n in ("nmr_oracle.c", "nmr_oracle_impl.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libnmr_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libnmr_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
    return _LIB


def set_threads(n):
    """Host threads used by the forward pixel loop and the pseudo-gradient face loop (default 1)."""
    lib().nmr_oracle_set_threads(int(n))


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _sfx(dtype):
    return "f32" if np.dtype(dtype) == np.float32 else "f64"


def _real(dtype):
    return ctypes.c_float if np.dtype(dtype) == np.float32 else ctypes.c_double


def rasterize_forward(faces, textures=None, image_size=256, near=0.1, far=100.0, eps=1e-4,
                      background_color=(0, 0, 0), return_rgb=True, return_alpha=True, return_depth=True,
                      dtype=np.float32):
    """RasterizeFunction.forward (rasterize.py:24-125) on the CPU.

    faces [B,F,3,3] (NDC xy, metric z), textures [B,F,ts,ts,ts,3].  Returns a dict with the raw
    (un-flipped, NHWC) maps plus the tensors the reference saves for backward.
    """
    L = lib()
    sfx, R = _sfx(dtype), _real(dtype)
    faces = np.ascontiguousarray(faces, dtype=dtype)
    B, F = faces.shape[:2]
    S = int(image_size)
    face_index_map = np.full((B, S, S), -1, dtype=np.int32)
    weight_map = np.zeros((B, S, S, 3), dtype=dtype)
    depth_map = np.full((B, S, S), far, dtype=dtype)
    face_inv_map = np.zeros((B, S, S, 3, 3), dtype=dtype) if return_depth else np.zeros((1,), dtype=dtype)
    faces_inv = np.zeros_like(faces)
    getattr(L, "nmr_forward_face_index_map_" + sfx)(
        _p(faces), _p(face_index_map), _p(weight_map), _p(depth_map), _p(face_inv_map), _p(faces_inv),
        ctypes.c_int(B), ctypes.c_int(F), ctypes.c_int(S), R(near), R(far), ctypes.c_int(int(return_depth)))
    out = dict(faces=faces, face_index_map=face_index_map, weight_map=weight_map, depth_map=depth_map,
               face_inv_map=face_inv_map, faces_inv=faces_inv, image_size=S, near=near, far=far, eps=eps,
               return_rgb=return_rgb, return_alpha=return_alpha, return_depth=return_depth, dtype=dtype)
    if return_rgb:
        textures = np.ascontiguousarray(textures, dtype=dtype)
        ts = textures.shape[2]
        rgb_map = np.zeros((B, S, S, 3), dtype=dtype)
        sampling_index_map = np.zeros((B, S, S, 8), dtype=np.int32)
        sampling_weight_map = np.zeros((B, S, S, 8), dtype=dtype)
        getattr(L, "nmr_forward_texture_sampling_" + sfx)(
            _p(faces), _p(textures), _p(face_index_map), _p(weight_map), _p(depth_map), _p(rgb_map),
            _p(sampling_index_map), _p(sampling_weight_map), ctypes.c_int(B), ctypes.c_int(F), ctypes.c_int(S),
            ctypes.c_int(ts), R(eps))
        # forward_background, rasterize.py:252-260
        bg = np.asarray(background_color, dtype=dtype)
        mask = (face_index_map >= 0).astype(dtype)[..., None]
        if bg.ndim == 1:
            rgb_map = rgb_map * mask + (1 - mask) * bg[None, None, None, :]
        else:
            rgb_map = rgb_map * mask + (1 - mask) * bg[:, None, None, :]
        out.update(textures=textures, rgb_map=np.ascontiguousarray(rgb_map), sampling_index_map=sampling_index_map,
                   sampling_weight_map=sampling_weight_map, texture_size=ts)
    if return_alpha:
        # forward_alpha_map, rasterize.py:246-249
        out["alpha_map"] = (face_index_map >= 0).astype(dtype)
    return out


def rasterize_backward(fwd, grad_rgb_map=None, grad_alpha_map=None, grad_depth_map=None, dtype=None):
    """RasterizeFunction.backward (rasterize.py:128-197): K4 store, K5 accumulate, K6 accumulate.

    ``dtype=np.float64`` re-runs the accumulations in double on the (float-decided) maps, which is
    what the gradient-tolerance tests compare the CUDA kernels against.
    """
    L = lib()
    dtype = fwd["dtype"] if dtype is None else dtype
    sfx, R = _sfx(dtype), _real(dtype)
    c = lambda a: np.ascontiguousarray(a, dtype=dtype)
    faces = c(fwd["faces"])
    B, F = faces.shape[:2]
    S = fwd["image_size"]
    idx = fwd["face_index_map"]
    grad_faces = np.zeros_like(faces)
    grad_textures = None
    rr, ra, rd = fwd["return_rgb"], fwd["return_alpha"], fwd["return_depth"]
    z1 = np.zeros((1,), dtype=dtype)
    g_rgb = c(grad_rgb_map) if (rr and grad_rgb_map is not None) else (np.zeros((B, S, S, 3), dtype) if rr else z1)
    g_alpha = c(grad_alpha_map) if (ra and grad_alpha_map is not None) else (np.zeros((B, S, S), dtype) if ra else z1)
    g_depth = c(grad_depth_map) if (rd and grad_depth_map is not None) else (np.zeros((B, S, S), dtype) if rd else z1)
    if rr or ra:
        getattr(L, "nmr_backward_pixel_map_" + sfx)(
            _p(faces), _p(idx), _p(c(fwd["rgb_map"]) if rr else z1), _p(c(fwd["alpha_map"]) if ra else z1),
            _p(g_rgb), _p(g_alpha), _p(grad_faces), ctypes.c_int(B), ctypes.c_int(F), ctypes.c_int(S),
            R(fwd["eps"]), ctypes.c_int(int(rr)), ctypes.c_int(int(ra)))
    if rr:
        grad_textures = np.zeros(fwd["textures"].shape, dtype=dtype)
        getattr(L, "nmr_backward_textures_" + sfx)(
            _p(idx), _p(c(fwd["sampling_weight_map"])), _p(fwd["sampling_index_map"]), _p(g_rgb),
            _p(grad_textures), ctypes.c_int(B), ctypes.c_int(F), ctypes.c_int(S), ctypes.c_int(fwd["texture_size"]))
    if rd:
        getattr(L, "nmr_backward_depth_map_" + sfx)(
            _p(faces), _p(c(fwd["depth_map"])), _p(idx), _p(c(fwd["face_inv_map"])), _p(c(fwd["weight_map"])),
            _p(g_depth), _p(grad_faces), ctypes.c_int(B), ctypes.c_int(F), ctypes.c_int(S))
    return grad_faces, grad_textures


def flip_rows(a, axis):
    return np.flip(a, axis=axis)


def rasterize_rgbad(faces, textures=None, image_size=256, anti_aliasing=True, near=0.1, far=100.0, eps=1e-4,
                    background_color=(0, 0, 0), return_rgb=True, return_alpha=True, return_depth=True,
                    dtype=np.float32):
    """rasterize.py:362-448: NCHW rgb, row flip of rgb/alpha/depth (NOT of the index/weight maps),
    optional 2x super-sampling followed by a 2x2 average pool."""
    S = image_size * 2 if anti_aliasing else image_size
    fwd = rasterize_forward(faces, textures, S, near, far, eps, background_color, return_rgb, return_alpha,
                            return_depth, dtype)
    rgb = alpha = depth = None
    if return_rgb:
        rgb = np.flip(fwd["rgb_map"].transpose(0, 3, 1, 2), axis=2)
    if return_alpha:
        alpha = np.flip(fwd["alpha_map"], axis=1)
    if return_depth:
        depth = np.flip(fwd["depth_map"], axis=1)
    if anti_aliasing:
        def pool(a):  # F.avg_pool2d(kernel 2), last two axes
            sh = a.shape
            return a.reshape(sh[:-2] + (sh[-2] // 2, 2, sh[-1] // 2, 2)).mean(axis=(-3, -1), dtype=a.dtype)
        rgb = pool(rgb) if return_rgb else None
        alpha = pool(alpha) if return_alpha else None
        depth = pool(depth) if return_depth else None
    return dict(rgb=None if rgb is None else np.ascontiguousarray(rgb),
                alpha=None if alpha is None else np.ascontiguousarray(alpha),
                depth=None if depth is None else np.ascontiguousarray(depth),
                face_inv_map=fwd["face_inv_map"], face_index_map=fwd["face_index_map"],
                weight_map=fwd["weight_map"], _fwd=fwd)


def parallel_over_batch(fn, arrays, threads):
    """Run ``fn(*per-sample slices)`` on a thread pool (ctypes drops the GIL): how the
    cpu_baseline / --impl reference leg of bench.py uses all host cores."""
    B = arrays[0].shape[0]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(lambda b: fn(*[None if a is None else a[b:b + 1] for a in arrays]), range(B)))
