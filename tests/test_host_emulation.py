"""CPU: the restructured rasterizer algorithm (face-parallel z-buffer with a packed min key, pixel
resolve, face-owned gradient gather, clipped outward scans), emulated serially on the host with the
product's own per-face/per-pixel functions (csrc/raster_math.h), against the oracle's literal
restatement of the reference.  Forward maps must be bit-identical; gradients within 1e-3 relative of
the oracle's float64 accumulation (BASELINE.json north_star tolerance)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import helpers
from helpers import onmr

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "emul", "libraster_emul.so")


def _build():
    src = os.path.join(HERE, "emul", "raster_emul.cpp")
    hdr = os.path.join(helpers.ROOT, "handobjectconsist_b200", "csrc", "raster_math.h")
    if (not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", _SO, src])
    return ctypes.CDLL(_SO)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def emul_forward(faces, tex, S, near=0.1, far=100.0, eps=1e-3, bg=(0, 0, 0), layout=0):
    L = _build()
    B, F = faces.shape[:2]
    ts = tex.shape[2]
    rgb = np.zeros((B, S, S, 3) if layout == 0 else (B, 3, S, S), np.float32)
    alpha = np.zeros((B, S, S), np.float32)
    depth = np.zeros((B, S, S), np.float32)
    idx = np.zeros((B, S, S), np.int32)
    wmap = np.zeros((B, S, S, 3), np.float32)
    inv = np.zeros((B, S, S, 3, 3), np.float32)
    bgv = np.asarray(bg, np.float32)
    L.emul_raster_forward(_p(faces), _p(tex), B, F, S, ts, ctypes.c_float(near), ctypes.c_float(far),
                          ctypes.c_float(eps), _p(bgv), layout, _p(rgb), _p(alpha), _p(depth), _p(idx), _p(wmap),
                          _p(inv))
    return dict(rgb=rgb, alpha=alpha, depth=depth, idx=idx, weight=wmap, inv=inv)


def emul_backward(faces, idx, rgb, g_rgb, g_alpha, g_depth, S, ts=2, near=0.1, far=100.0, eps=1e-3, layout=0,
                  use_alpha=1):
    L = _build()
    B, F = faces.shape[:2]
    gf = np.full((B, F, 3, 3), np.nan, np.float32)
    gt = np.full((B, F, ts, ts, ts, 3), np.nan, np.float32)
    L.emul_raster_backward(_p(faces), _p(idx), _p(rgb), _p(g_rgb), _p(g_alpha), _p(g_depth), B, F, S, ts,
                           ctypes.c_float(near), ctypes.c_float(far), ctypes.c_float(eps), layout, use_alpha, _p(gf),
                           _p(gt))
    return gf, gt


@pytest.mark.parametrize("S,seed", [(64, 0), (48, 3), (33, 5)])
def test_forward_bit_exact(S, seed):
    faces, tex, _ = helpers.scene_faces(2, S, seed=seed)
    ora = onmr.rasterize_forward(faces, tex, S, 0.1, 100.0, 1e-3, (0.1, 0.2, 0.3), True, True, True)
    emu = emul_forward(faces, tex, S, bg=(0.1, 0.2, 0.3))
    assert (ora["face_index_map"] >= 0).mean() > 0.02  # the scene covers something
    np.testing.assert_array_equal(emu["idx"], ora["face_index_map"])
    np.testing.assert_array_equal(emu["depth"], ora["depth_map"])
    np.testing.assert_array_equal(emu["weight"], ora["weight_map"])
    np.testing.assert_array_equal(emu["alpha"], ora["alpha_map"])
    np.testing.assert_array_equal(emu["inv"], ora["face_inv_map"])
    np.testing.assert_array_equal(emu["rgb"], ora["rgb_map"])


def test_forward_image_layout_is_flipped_nchw():
    faces, tex, _ = helpers.scene_faces(1, 40, seed=2)
    raw = emul_forward(faces, tex, 40, layout=0)
    img = emul_forward(faces, tex, 40, layout=1)
    np.testing.assert_array_equal(img["rgb"], np.flip(raw["rgb"].transpose(0, 3, 1, 2), axis=2))
    np.testing.assert_array_equal(img["alpha"], np.flip(raw["alpha"], axis=1))
    np.testing.assert_array_equal(img["depth"], np.flip(raw["depth"], axis=1))
    np.testing.assert_array_equal(img["idx"], raw["idx"])


@pytest.mark.parametrize("dense,S", [(True, 48), (False, 48), (False, 96)])
def test_backward_matches_oracle(dense, S):
    faces, tex, _ = helpers.scene_faces(2, S, seed=1)
    ora = onmr.rasterize_forward(faces, tex, S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True)
    rng = np.random.default_rng(0)
    g_rgb = rng.normal(size=ora["rgb_map"].shape).astype(np.float32)
    g_alpha = rng.normal(size=ora["alpha_map"].shape).astype(np.float32)
    g_depth = rng.normal(size=ora["depth_map"].shape).astype(np.float32)
    if not dense:  # gradients only on covered pixels, like the photometric loss produces
        cov = (ora["face_index_map"] >= 0).astype(np.float32)
        g_rgb *= cov[..., None]
        g_alpha *= cov
    gf32, gt32 = onmr.rasterize_backward(ora, g_rgb, g_alpha, g_depth)
    gf64, gt64 = onmr.rasterize_backward(ora, g_rgb, g_alpha, g_depth, dtype=np.float64)
    gf, gt = emul_backward(faces, ora["face_index_map"], ora["rgb_map"], g_rgb, g_alpha, g_depth, S)
    assert np.isfinite(gf).all() and np.isfinite(gt).all()
    assert np.abs(gf32).max() > 0 and np.abs(gt32).max() > 0
    # same fp32 decisions as the reference restatement -> element-wise 1e-3
    assert helpers.rel_err(gt, gt32) < 1e-3
    assert helpers.rel_err(gf, gf32) < 1e-3
    # fp64 re-evaluation flips a few floor/ceil / sign gates: norm-wise bound only
    assert np.linalg.norm(gf - gf64) / np.linalg.norm(gf64) < 2e-2
    assert np.linalg.norm(gt - gt64) / np.linalg.norm(gt64) < 2e-2
    # the line pass adds the reference's +-eps as one constant per outward scan: every scanned pixel must have had
    # the sign that constant assumes
    L = _build()
    L.emul_sign_violations.restype = ctypes.c_long
    assert L.emul_sign_violations() == 0


def test_backward_float_oracle_agrees_with_double():
    """The f32 oracle itself stays within tolerance of its f64 twin on this scene (sanity of the bar)."""
    S = 48
    faces, tex, _ = helpers.scene_faces(1, S, seed=1)
    ora = onmr.rasterize_forward(faces, tex, S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True)
    rng = np.random.default_rng(0)
    g_rgb = rng.normal(size=ora["rgb_map"].shape).astype(np.float32)
    gf32, gt32 = onmr.rasterize_backward(ora, g_rgb, None, None)
    gf64, gt64 = onmr.rasterize_backward(ora, g_rgb, None, None, dtype=np.float64)
    assert helpers.rel_err(gt32, gt64) < 1e-3
