"""CPU: the C restatement of the rasterizer (oracle/) against (a) its committed golden vectors and (b)
analytic unit scenes that pin every convention of SURVEY.md Appendix A / C (pixel centres, y-up raster
rows, back-face rule, inclusive edges, clamp + renormalise, perspective-correct depth, near/far, z-order
and ties, texture-cube sampling, true-derivative K5/K6, the row flip of rasterize_rgbad)."""
import os

import numpy as np
import pytest

import helpers
from helpers import onmr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["hand64", "handobj48", "handobj40"])
def test_oracle_matches_committed_golden(name):
    g = np.load(os.path.join(GOLD, f"raster_{name}.npz"))
    B, S, seed, with_obj = [int(v) for v in g["cfg"]]
    faces, tex, _ = helpers.scene_faces(B, S, seed=seed, with_object=bool(with_obj))
    ora = onmr.rasterize_forward(faces, tex, S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True)
    np.testing.assert_array_equal(ora["face_index_map"], g["face_index_map"].astype(np.int32))
    np.testing.assert_array_equal(ora["depth_map"], g["depth_map"])
    np.testing.assert_array_equal(ora["weight_map"], g["weight_map"])
    np.testing.assert_array_equal(ora["rgb_map"], g["rgb_map"])
    rng = np.random.default_rng(7)
    g_rgb = rng.normal(size=ora["rgb_map"].shape).astype(np.float32)
    g_alpha = rng.normal(size=ora["alpha_map"].shape).astype(np.float32)
    g_depth = rng.normal(size=ora["alpha_map"].shape).astype(np.float32)
    hit = (g["hit_b"].astype(np.int64), g["hit_f"].astype(np.int64))
    gf32, gt32 = onmr.rasterize_backward(ora, g_rgb, g_alpha, g_depth)
    gf64, gt64 = onmr.rasterize_backward(ora, g_rgb, g_alpha, g_depth, dtype=np.float64)
    np.testing.assert_array_equal(gf32[hit], g["gf32"])
    np.testing.assert_array_equal(gt32[hit], g["gt32"])
    np.testing.assert_allclose(gf64[hit], g["gf64"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(gt64[hit], g["gt64"], rtol=1e-12, atol=1e-12)
    rest = np.ones(gf32.shape[:2], bool)
    rest[hit] = False
    assert not gf32[rest].any() and not gt32[rest].any()


def _tri(points, z=(1.0, 1.0, 1.0)):
    return np.array([[[[points[k][0], points[k][1], z[k]] for k in range(3)]]], dtype=np.float32)


def _const_tex(F=1, value=(1.0, 0.5, 0.25)):
    t = np.zeros((1, F, 2, 2, 2, 3), np.float32)
    t[...] = np.asarray(value, np.float32)
    return t


def test_axis_aligned_triangle_coverage_and_barycentrics():
    S = 8
    # right triangle with the right angle at the lower-left image corner, legs = full width/height (NDC y up)
    faces = _tri([(-1, -1), (1, -1), (-1, 1)])
    o = onmr.rasterize_forward(faces, _const_tex(), S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True)
    idx = o["face_index_map"][0]
    # pixel (yi, xi) centre in NDC = ((2xi+1-S)/S, (2yi+1-S)/S); inside <=> x + y <= 0 <=> xi + yi <= S-1
    yy, xx = np.mgrid[0:S, 0:S]
    np.testing.assert_array_equal(idx >= 0, (xx + yy) <= S - 1)
    assert idx[0, 0] == 0 and idx[S - 1, S - 1] == -1          # raster row 0 is the BOTTOM of the image
    w = o["weight_map"][0]
    # barycentrics are affine in pixel-index space with p = 0.5 (x S + S - 1): vertex 0 at p = (-0.5, -0.5)
    np.testing.assert_allclose(w[0, 0], [1 - 1.0 / S, 0.5 / S, 0.5 / S], atol=1e-6)
    np.testing.assert_allclose(w[..., 0][idx >= 0] + w[..., 1][idx >= 0] + w[..., 2][idx >= 0], 1.0, atol=1e-6)
    np.testing.assert_array_equal(o["alpha_map"][0], (idx >= 0).astype(np.float32))
    np.testing.assert_allclose(o["rgb_map"][0][idx >= 0], np.tile([1.0, 0.5, 0.25], (int((idx >= 0).sum()), 1)), atol=1e-6)
    assert (o["depth_map"][0][idx < 0] == 100.0).all() and np.allclose(o["depth_map"][0][idx >= 0], 1.0)


def test_back_face_rule_and_fill_back():
    S = 8
    ccw = _tri([(-1, -1), (1, -1), (-1, 1)])
    cw = ccw[:, :, ::-1].copy()
    assert (onmr.rasterize_forward(ccw, None, S, 0.1, 100.0, 1e-3, (0, 0, 0), False, True, False)["face_index_map"] >= 0).any()
    assert not (onmr.rasterize_forward(cw, None, S, 0.1, 100.0, 1e-3, (0, 0, 0), False, True, False)["face_index_map"] >= 0).any()
    both = np.concatenate([cw, ccw], axis=1)   # fill_back: exactly one copy of every triangle survives
    idx = onmr.rasterize_forward(both, None, S, 0.1, 100.0, 1e-3, (0, 0, 0), False, True, False)["face_index_map"]
    assert set(np.unique(idx)) == {-1, 1}


def test_perspective_correct_depth_and_near_far():
    S = 16
    faces = _tri([(-1, -1), (1, -1), (-1, 1)], z=(1.0, 2.0, 4.0))
    o = onmr.rasterize_forward(faces, None, S, 0.1, 100.0, 1e-3, (0, 0, 0), False, False, True)
    w, d, idx = o["weight_map"][0], o["depth_map"][0], o["face_index_map"][0]
    wc = w[idx >= 0]
    expect = 1.0 / (wc[..., 0] / 1.0 + wc[..., 1] / 2.0 + wc[..., 2] / 4.0)
    np.testing.assert_allclose(d[idx >= 0], expect, rtol=1e-6)
    for z, near, far in ((0.05, 0.1, 100.0), (150.0, 0.1, 100.0)):
        f = _tri([(-1, -1), (1, -1), (-1, 1)], z=(z, z, z))
        assert not (onmr.rasterize_forward(f, None, S, near, far, 1e-3, (0, 0, 0), False, True, False)["face_index_map"] >= 0).any()


def test_z_order_and_exact_ties():
    S = 8
    near_t = _tri([(-1, -1), (1, -1), (-1, 1)], z=(1.0, 1.0, 1.0))
    far_t = _tri([(-1, -1), (1, -1), (-1, 1)], z=(2.0, 2.0, 2.0))
    idx = onmr.rasterize_forward(np.concatenate([far_t, near_t], 1), None, S, 0.1, 100.0, 1e-3, (0, 0, 0), False, True,
                                 False)["face_index_map"]
    assert set(np.unique(idx)) == {-1, 1}       # nearer face wins regardless of order
    idx = onmr.rasterize_forward(np.concatenate([near_t, near_t], 1), None, S, 0.1, 100.0, 1e-3, (0, 0, 0), False, True,
                                 False)["face_index_map"]
    assert set(np.unique(idx)) == {-1, 0}       # strict `<`: the first of two equal depths stays


def test_texture_cube_corners_and_vertex_colours():
    S = 16
    faces = _tri([(-1, -1), (1, -1), (-1, 1)])
    # cube whose trilinear sample is b0*c0 + b1*c1 + b2*c2 (helpers' batch_vertex_textures construction)
    c = np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]], np.float32)
    tex = np.zeros((1, 1, 2, 2, 2, 3), np.float32)
    for i in range(2):
        for j in range(2):
            for k in range(2):
                tex[0, 0, i, j, k] = i * c[0] + j * c[1] + k * c[2]
    o = onmr.rasterize_forward(faces, tex, S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True)
    idx, w = o["face_index_map"][0], o["weight_map"][0]
    clipped = np.minimum(w, 1.0 - 1e-3)         # texture coordinates are clamped to ts-1-eps
    np.testing.assert_allclose(o["rgb_map"][0][idx >= 0], clipped[idx >= 0], atol=1e-6)


def test_rgbad_row_flip_and_background():
    S = 8
    faces = _tri([(-1, -1), (1, -1), (-1, 1)])
    out = onmr.rasterize_rgbad(faces, _const_tex(), S, False, 0.1, 100.0, 1e-3, (0.2, 0.4, 0.6))
    assert out["rgb"].shape == (1, 3, S, S)
    assert out["alpha"][0, S - 1, 0] == 1 and out["alpha"][0, 0, 0] == 1 and out["alpha"][0, 0, S - 1] == 0
    assert out["alpha"][0, S - 1, S - 1] == 1 and out["alpha"][0, 0, 1] == 0   # image rows: row 0 = top
    np.testing.assert_allclose(out["rgb"][0, :, 0, S - 1], [0.2, 0.4, 0.6])
    assert out["face_index_map"][0, 0, 0] == 0                                     # index map is NOT flipped
    aa = onmr.rasterize_rgbad(faces, _const_tex(), S, True, 0.1, 100.0, 1e-3, (0, 0, 0))
    assert aa["alpha"].shape == (1, S, S) and set(np.unique(aa["alpha"])) <= {0.0, 0.25, 0.5, 0.75, 1.0}


def test_texture_and_depth_gradients_are_true_derivatives():
    """K5 is the exact derivative of rgb w.r.t. textures (rgb is linear in them); K6 the derivative of the
    interpolated depth w.r.t. vertex z (checked by central differences in float64 on interior pixels)."""
    S = 16
    faces, tex, _ = helpers.scene_faces(1, S, seed=2, with_object=False)
    fwd = onmr.rasterize_forward(faces, tex, S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True, dtype=np.float64)
    rng = np.random.default_rng(0)
    g_rgb = rng.normal(size=fwd["rgb_map"].shape)
    _, gt = onmr.rasterize_backward(fwd, g_rgb, None, None, dtype=np.float64)
    d_tex = rng.normal(size=tex.shape) * 1e-3
    fwd2 = onmr.rasterize_forward(faces, tex.astype(np.float64) + d_tex, S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True,
                                  dtype=np.float64)
    lhs = ((fwd2["rgb_map"] - fwd["rgb_map"]) * g_rgb).sum()
    np.testing.assert_allclose(lhs, (gt * d_tex).sum(), rtol=1e-9)
    # depth: perturb z of every vertex of every face; compare only where coverage does not change
    g_depth = rng.normal(size=fwd["depth_map"].shape)
    gf, _ = onmr.rasterize_backward(dict(fwd, return_rgb=False, return_alpha=False), None, None, g_depth, dtype=np.float64)
    dz = np.zeros(faces.shape)
    dz[..., 2] = rng.normal(size=faces.shape[:3]) * 1e-7
    fp = onmr.rasterize_forward(faces.astype(np.float64) + dz, None, S, 0.1, 100.0, 1e-3, (0, 0, 0), False, False, True,
                                dtype=np.float64)
    fm = onmr.rasterize_forward(faces.astype(np.float64) - dz, None, S, 0.1, 100.0, 1e-3, (0, 0, 0), False, False, True,
                                dtype=np.float64)
    same = (fp["face_index_map"] == fwd["face_index_map"]) & (fm["face_index_map"] == fwd["face_index_map"])
    lhs = (((fp["depth_map"] - fm["depth_map"]) / 2) * g_depth)[same].sum()
    g_masked = np.where(same, g_depth, 0.0)
    gf_m, _ = onmr.rasterize_backward(dict(fwd, return_rgb=False, return_alpha=False), None, None, g_masked,
                                      dtype=np.float64)
    np.testing.assert_allclose(lhs, (gf_m * dz).sum(), rtol=1e-5)
    assert np.abs(gf[..., 2]).max() > 0
