"""GPU parity: the CUDA rasterizer (through the reference-shaped python surface, i.e. through the
C ABI of libhoc_b200.so) against the CPU oracle on the same seeded inputs.
Bar (BASELINE.json north_star): forward maps bit-exact (integer index map exact, fp32 maps 1e-4 abs --
they are in fact identical); gradients within 1e-3 relative of the fp32 oracle, element by element.
The pseudo-gradient is full of discrete decisions (floor/ceil of edge crossings, delta > 0 gates) that
the reference takes in fp32; an fp64 re-evaluation flips a few of them, so the fp64 oracle is only used
for a norm-wise sanity bound (GRAD_F64_NORM_TOL)."""
import numpy as np
import pytest
import torch

import helpers
from helpers import onmr

pytestmark = pytest.mark.gpu

ABS_TOL = 1e-4   # pixels (north_star)
REL_TOL = 1e-3   # gradients (north_star), vs the fp32 oracle
GRAD_F64_NORM_TOL = 2e-2  # || g - g64 || / || g64 ||


def _check_grads(g, g32, g64=None):
    g = np.asarray(g)
    assert np.isfinite(g).all()
    assert helpers.rel_err(g, g32) < REL_TOL
    if g64 is not None:
        n = np.linalg.norm(g.astype(np.float64) - g64) / max(np.linalg.norm(g64), 1e-30)
        assert n < GRAD_F64_NORM_TOL, n


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _run_function(faces, tex, S, bg=(0, 0, 0), eps=1e-3, rr=True, ra=True, rd=True):
    from handobjectconsist_b200.neurender.rasterize import RasterizeFunction
    f = _cuda(faces).requires_grad_(True)
    t = _cuda(tex).requires_grad_(True) if tex is not None else None
    out = RasterizeFunction.apply(f, t, S, 0.1, 100.0, eps, bg, rr, ra, rd)
    return f, t, out


@pytest.mark.parametrize("S,B,seed", [(64, 2, 0), (48, 3, 3), (33, 1, 5), (128, 2, 7)])
def test_forward_matches_oracle(S, B, seed):
    faces, tex, _ = helpers.scene_faces(B, S, seed=seed)
    ora = onmr.rasterize_forward(faces, tex, S, 0.1, 100.0, 1e-3, (0.1, 0.2, 0.3), True, True, True)
    _, _, (rgb, alpha, depth, idx, inv, wmap) = _run_function(faces, tex, S, bg=(0.1, 0.2, 0.3))
    assert (ora["face_index_map"] >= 0).mean() > 0.02
    np.testing.assert_array_equal(idx.cpu().numpy(), ora["face_index_map"])
    np.testing.assert_array_equal(depth.detach().cpu().numpy(), ora["depth_map"])
    np.testing.assert_array_equal(wmap.detach().cpu().numpy(), ora["weight_map"])
    np.testing.assert_array_equal(alpha.detach().cpu().numpy(), ora["alpha_map"])
    np.testing.assert_array_equal(inv.detach().cpu().numpy(), ora["face_inv_map"])
    assert np.abs(rgb.detach().cpu().numpy() - ora["rgb_map"]).max() <= ABS_TOL
    np.testing.assert_array_equal(rgb.detach().cpu().numpy(), ora["rgb_map"])


def test_forward_flags_and_disabled_outputs():
    S = 32
    faces, tex, _ = helpers.scene_faces(1, S, seed=2)
    _, _, (rgb, alpha, depth, idx, inv, wmap) = _run_function(faces, None, S, rr=False, ra=True, rd=False)
    assert rgb.numel() == 0 and depth.numel() == 0 and inv.numel() == 1
    ora = onmr.rasterize_forward(faces, None, S, 0.1, 100.0, 1e-3, (0, 0, 0), False, True, False)
    np.testing.assert_array_equal(alpha.detach().cpu().numpy(), ora["alpha_map"])
    np.testing.assert_array_equal(idx.cpu().numpy(), ora["face_index_map"])


def test_forward_edge_cases():
    from handobjectconsist_b200.neurender.rasterize import RasterizeFunction
    S = 16
    # empty batch of faces: everything is background
    f = torch.zeros(2, 0, 3, 3, device="cuda")
    t = torch.zeros(2, 0, 2, 2, 2, 3, device="cuda")
    rgb, alpha, depth, idx, inv, w = RasterizeFunction.apply(f, t, S, 0.1, 100.0, 1e-3, (0.5, 0.25, 0.125), True, True, True)
    assert (idx == -1).all() and (alpha == 0).all() and (depth == 100.0).all()
    assert torch.equal(rgb[0, 0, 0].cpu(), torch.tensor([0.5, 0.25, 0.125]))
    # full-screen triangle, a degenerate one, one with NaN, one behind `near`, one beyond `far`
    faces = np.array([[[[-3, -3, 1.0], [3, -3, 1.0], [0, 3, 2.0]],
                       [[0, 0, 1.0], [0, 0, 1.0], [0, 0, 1.0]],
                       [[np.nan, 0, 1.0], [1, 0, 1.0], [0, 1, 1.0]],
                       [[-1, -1, 0.05], [1, -1, 0.05], [0, 1, 0.05]],
                       [[-1, -1, 200.0], [1, -1, 200.0], [0, 1, 200.0]]]], dtype=np.float32)
    tex = np.random.default_rng(0).uniform(size=(1, 5, 2, 2, 2, 3)).astype(np.float32)
    ora = onmr.rasterize_forward(faces, tex, S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True)
    _, _, (rgb, alpha, depth, idx, inv, wmap) = _run_function(faces, tex, S)
    np.testing.assert_array_equal(idx.cpu().numpy(), ora["face_index_map"])
    np.testing.assert_array_equal(depth.detach().cpu().numpy(), ora["depth_map"])
    np.testing.assert_array_equal(rgb.detach().cpu().numpy(), ora["rgb_map"])
    # CPU tensors are rejected like the reference (rasterize.py:346-347)
    from handobjectconsist_b200.neurender.rasterize import Rasterize
    with pytest.raises(TypeError):
        Rasterize(S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True)(torch.zeros(1, 1, 3, 3), torch.zeros(1, 1, 2, 2, 2, 3))


def test_z_ties_lowest_face_index_wins():
    S = 24
    tri = [[-0.8, -0.8, 1.0], [0.8, -0.8, 1.0], [0.0, 0.9, 1.0]]
    faces = np.array([[tri, tri, tri]], dtype=np.float32)  # three coplanar copies -> exact ties
    tex = np.zeros((1, 3, 2, 2, 2, 3), np.float32)
    for i in range(3):
        tex[0, i] = i + 1
    _, _, (rgb, alpha, depth, idx, inv, wmap) = _run_function(faces, tex, S)
    covered = idx.cpu().numpy() >= 0
    assert covered.any()
    assert (idx.cpu().numpy()[covered] == 0).all()
    ora = onmr.rasterize_forward(faces, tex, S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True)
    np.testing.assert_array_equal(idx.cpu().numpy(), ora["face_index_map"])


@pytest.mark.parametrize("dense,S,B", [(True, 48, 2), (False, 48, 2), (False, 96, 2), (True, 64, 1)])
def test_backward_matches_oracle(dense, S, B):
    faces, tex, _ = helpers.scene_faces(B, S, seed=1)
    ora = onmr.rasterize_forward(faces, tex, S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True)
    rng = np.random.default_rng(0)
    g_rgb = rng.normal(size=ora["rgb_map"].shape).astype(np.float32)
    g_alpha = rng.normal(size=ora["alpha_map"].shape).astype(np.float32)
    g_depth = rng.normal(size=ora["depth_map"].shape).astype(np.float32)
    if not dense:
        cov = (ora["face_index_map"] >= 0).astype(np.float32)
        g_rgb *= cov[..., None]
        g_alpha *= cov
    gf64, gt64 = onmr.rasterize_backward(ora, g_rgb, g_alpha, g_depth, dtype=np.float64)
    gf32, gt32 = onmr.rasterize_backward(ora, g_rgb, g_alpha, g_depth)
    f, t, (rgb, alpha, depth, idx, inv, wmap) = _run_function(faces, tex, S)
    loss = (rgb * _cuda(g_rgb)).sum() + (alpha * _cuda(g_alpha)).sum() + (depth * _cuda(g_depth)).sum()
    loss.backward()
    gf, gt = f.grad.cpu().numpy(), t.grad.cpu().numpy()
    _check_grads(gt, gt32, gt64)
    _check_grads(gf, gf32, gf64)
    # a second run reproduces the gradients to rounding (float atomics reorder the sums): bounded against the
    # gradient scale here, bit for bit in the reproducible mode (test_backward_reproducible_mode)
    f2, t2, (rgb2, alpha2, depth2, _, _, _) = _run_function(faces, tex, S)
    ((rgb2 * _cuda(g_rgb)).sum() + (alpha2 * _cuda(g_alpha)).sum() + (depth2 * _cuda(g_depth)).sum()).backward()
    assert np.abs(f2.grad.cpu().numpy() - gf).max() <= 1e-3 * np.abs(gf).max()
    assert np.abs(t2.grad.cpu().numpy() - gt).max() <= 1e-3 * np.abs(gt).max()


def test_backward_on_a_large_raster():
    """S = 1400: the line pass stages (S + 32) x 36 bytes per CTA = 51.6 KB, above the 48 KB a kernel gets without
    opting in (cudaFuncAttributeMaxDynamicSharedMemorySize), and cuts its scans into 32-pixel chunks (rasters above
    320).  A handful of large triangles keeps the oracle fast."""
    S, F = 1400, 10
    rng = np.random.default_rng(5)
    faces = np.zeros((1, F, 3, 3), np.float32)
    for k in range(F):
        c = rng.uniform(-0.6, 0.6, 2)
        ang = np.sort(rng.uniform(0, 2 * np.pi, 3))  # counter-clockwise: front-facing
        r = rng.uniform(0.15, 0.5, 3)
        faces[0, k, :, 0] = c[0] + r * np.cos(ang)
        faces[0, k, :, 1] = c[1] + r * np.sin(ang)
        faces[0, k, :, 2] = rng.uniform(1.0, 3.0, 3)
    tex = rng.uniform(0.1, 1.0, (1, F, 2, 2, 2, 3)).astype(np.float32)
    ora = onmr.rasterize_forward(faces, tex, S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True)
    assert (ora["face_index_map"] >= 0).mean() > 0.05
    g_rgb = rng.normal(size=ora["rgb_map"].shape).astype(np.float32)
    g_rgb *= (ora["face_index_map"] >= 0).astype(np.float32)[..., None]
    g_alpha = np.zeros_like(ora["alpha_map"])
    g_depth = np.zeros_like(ora["depth_map"])
    gf64, gt64 = onmr.rasterize_backward(ora, g_rgb, g_alpha, g_depth, dtype=np.float64)
    gf32, gt32 = onmr.rasterize_backward(ora, g_rgb, g_alpha, g_depth)
    f, t, (rgb, alpha, depth, idx, inv, wmap) = _run_function(faces, tex, S)
    np.testing.assert_array_equal(idx.cpu().numpy(), ora["face_index_map"])
    (rgb * _cuda(g_rgb)).sum().backward()
    # triangles of ~10^5 pixels: every gradient is a sum of 10^4 .. 10^5 terms that cancel to 1e-3 of their size, and
    # the fp32 summation orders differ (serial per face in the oracle, chunks and atomics here): the bar is relative
    # to the gradient scale, as for every comparison of float-atomic sums, plus the norm-wise fp64 bound
    for g, g32, g64 in ((t.grad.cpu().numpy(), gt32, gt64), (f.grad.cpu().numpy(), gf32, gf64)):
        assert np.isfinite(g).all() and np.abs(g32).max() > 0
        assert np.abs(g - g32).max() <= 1e-3 * np.abs(g32).max()
        assert np.linalg.norm(g - g64) / np.linalg.norm(g64) < 2e-2


@pytest.mark.parametrize("S,B", [(48, 2), (96, 2)])
def test_backward_reproducible_mode(det_mode, S, B):
    """HOC_TUNE_DETERMINISTIC: same parity bar against the oracle, and two runs give the same bits."""
    faces, tex, _ = helpers.scene_faces(B, S, seed=1)
    ora = onmr.rasterize_forward(faces, tex, S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True)
    rng = np.random.default_rng(0)
    g_rgb = rng.normal(size=ora["rgb_map"].shape).astype(np.float32)
    g_alpha = rng.normal(size=ora["alpha_map"].shape).astype(np.float32)
    g_depth = rng.normal(size=ora["depth_map"].shape).astype(np.float32)
    gf32, gt32 = onmr.rasterize_backward(ora, g_rgb, g_alpha, g_depth)
    runs = []
    for _ in range(3):
        f, t, (rgb, alpha, depth, _, _, _) = _run_function(faces, tex, S)
        ((rgb * _cuda(g_rgb)).sum() + (alpha * _cuda(g_alpha)).sum() + (depth * _cuda(g_depth)).sum()).backward()
        runs.append((f.grad.clone(), t.grad.clone()))
    _check_grads(runs[0][0].cpu().numpy(), gf32)
    _check_grads(runs[0][1].cpu().numpy(), gt32)
    for gf, gt in runs[1:]:
        assert torch.equal(gf, runs[0][0]) and torch.equal(gt, runs[0][1])


def test_backward_partial_outputs():
    """Only some outputs carry gradient (None for the others), silhouette-only mode, texture size 3."""
    S = 40
    faces, tex, _ = helpers.scene_faces(1, S, seed=4)
    rng = np.random.default_rng(1)
    # rgb only
    ora = onmr.rasterize_forward(faces, tex, S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True)
    g_rgb = rng.normal(size=ora["rgb_map"].shape).astype(np.float32)
    gf32, gt32 = onmr.rasterize_backward(ora, g_rgb, None, None)
    f, t, out = _run_function(faces, tex, S)
    (out[0] * _cuda(g_rgb)).sum().backward()
    _check_grads(f.grad.cpu().numpy(), gf32)
    _check_grads(t.grad.cpu().numpy(), gt32)
    # alpha only, no textures at all
    ora = onmr.rasterize_forward(faces, None, S, 0.1, 100.0, 1e-3, (0, 0, 0), False, True, False)
    g_alpha = rng.normal(size=ora["alpha_map"].shape).astype(np.float32)
    gf32, _ = onmr.rasterize_backward(ora, None, g_alpha, None)
    f, _, out = _run_function(faces, None, S, rr=False, ra=True, rd=False)
    (out[1] * _cuda(g_alpha)).sum().backward()
    _check_grads(f.grad.cpu().numpy(), gf32)
    # texture size 3 (generic path)
    tex3 = rng.uniform(size=(1, faces.shape[1], 3, 3, 3, 3)).astype(np.float32)
    ora = onmr.rasterize_forward(faces, tex3, S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True)
    g_rgb = rng.normal(size=ora["rgb_map"].shape).astype(np.float32)
    gf32, gt32 = onmr.rasterize_backward(ora, g_rgb, None, None)
    f, t, out = _run_function(faces, tex3, S)
    np.testing.assert_array_equal(out[0].detach().cpu().numpy(), ora["rgb_map"])
    (out[0] * _cuda(g_rgb)).sum().backward()
    _check_grads(t.grad.cpu().numpy(), gt32)
    _check_grads(f.grad.cpu().numpy(), gf32)


@pytest.mark.parametrize("aa", [False, True])
def test_rasterize_rgbad_matches_oracle(aa):
    from handobjectconsist_b200.neurender.rasterize import rasterize_rgbad
    S = 32
    faces, tex, _ = helpers.scene_faces(2, S * (2 if aa else 1), seed=6)
    ora = onmr.rasterize_rgbad(faces, tex, S, aa, 0.1, 100.0, 1e-3, (0, 0, 0))
    out = rasterize_rgbad(_cuda(faces), _cuda(tex), S, aa, 0.1, 100.0, 1e-3, (0, 0, 0))
    for k in ("rgb", "alpha", "depth"):
        assert np.abs(out[k].cpu().numpy() - ora[k]).max() <= ABS_TOL, k
    np.testing.assert_array_equal(out["face_index_map"].cpu().numpy(), ora["face_index_map"])
    np.testing.assert_array_equal(out["weight_map"].cpu().numpy(), ora["weight_map"])
    np.testing.assert_array_equal(out["face_inv_map"].cpu().numpy(), ora["face_inv_map"])


def test_image_layout_backward_matches_raw_layout(det_mode):
    """The fused NCHW/flipped output path gives the same gradients as the reference-shaped path
    followed by permute + flip.  (Reproducible mode: both layouts add the same terms, only in another order.)"""
    from handobjectconsist_b200.neurender.rasterize import rasterize_rgbad
    S = 48
    faces, tex, _ = helpers.scene_faces(2, S, seed=8)
    rng = np.random.default_rng(3)
    g_rgb = _cuda(rng.normal(size=(2, 3, S, S)).astype(np.float32))
    g_a = _cuda(rng.normal(size=(2, S, S)).astype(np.float32))
    g_d = _cuda(rng.normal(size=(2, S, S)).astype(np.float32))
    f1 = _cuda(faces).requires_grad_(True)
    t1 = _cuda(tex).requires_grad_(True)
    o = rasterize_rgbad(f1, t1, S, False, 0.1, 100.0, 1e-3, (0, 0, 0))
    ((o["rgb"] * g_rgb).sum() + (o["alpha"] * g_a).sum() + (o["depth"] * g_d).sum()).backward()
    f2, t2, (rgb, alpha, depth, _, _, _) = _run_function(faces, tex, S)
    rgb = rgb.permute(0, 3, 1, 2).flip(2)
    assert torch.equal(rgb, o["rgb"])
    ((rgb * g_rgb).sum() + (alpha.flip(1) * g_a).sum() + (depth.flip(1) * g_d).sum()).backward()
    assert helpers.rel_err(f1.grad.cpu().numpy(), f2.grad.cpu().numpy()) < 1e-6
    assert helpers.rel_err(t1.grad.cpu().numpy(), t2.grad.cpu().numpy()) < 1e-6
