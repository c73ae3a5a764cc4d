"""GPU: BASELINE.json's full sizes (configs[1]/[2]: 16-32 meshes of 9104 faces at 256x256; configs[4]: 480x480
raster cropped to 480x270), checked through size-independent properties (the oracle comparisons at these sizes are
in test_gpu_parity_fullsize.py): coverage / alpha / depth consistency, fill_back renders every triangle once,
batch-permutation invariance, linearity of the backward in the incoming gradient, zero gradient for untouched
faces, rectangular crops.  Runs LAST (tests/conftest.py)."""
import numpy as np
import pytest
import torch

import helpers
from handobjectconsist_b200 import synth

pytestmark = pytest.mark.gpu


def _ndc_faces(sc, S, dev):
    from handobjectconsist_b200.neurender import nrfuncs as nr
    v = sc["verts1"].to(dev)
    ndc = nr.projection(v, sc["K"].to(dev), torch.eye(3, device=dev)[None], torch.zeros(1, 1, 3, device=dev),
                        torch.zeros(1, 5, device=dev), float(S))
    f = sc["faces"].to(dev)
    f2 = torch.cat([f, f.flip(-1)], 1)
    return nr.vertices_to_faces(ndc, f2)


@pytest.mark.parametrize("B,S", [(32, 256), (4, 480)])
def test_full_size_forward_properties(B, S):
    from handobjectconsist_b200.neurender.rasterize import RasterizeFunction
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, S, seed=42)
    faces = _ndc_faces(sc, S, dev)
    Fn = faces.shape[1]
    tex = torch.rand(B, Fn, 2, 2, 2, 3, device=dev)
    rgb, alpha, depth, idx, inv, w = RasterizeFunction.apply(faces, tex, S, 0.1, 100.0, 1e-3, (0.25, 0.5, 0.75), True, True, True)
    cov = idx >= 0
    assert 0.02 < cov.float().mean().item() < 0.5
    assert torch.equal(alpha, cov.float())
    assert (depth[~cov] == 100.0).all() and (depth[cov] > 0.1).all() and (depth[cov] < 1.0).all()
    assert torch.equal(rgb[~cov], torch.tensor([0.25, 0.5, 0.75], device=dev).expand(int((~cov).sum()), 3))
    assert torch.allclose(w[cov].sum(-1), torch.ones(int(cov.sum()), device=dev), atol=1e-5) and (w[~cov] == 0).all()
    # fill_back: a triangle and its reversed copy are coplanar, exactly one of them is front-facing
    half = Fn // 2
    assert int(idx.max()) < Fn
    # swapping the two halves of the face list renders the same image with indices moved by +-half
    faces_sw = torch.cat([faces[:, half:], faces[:, :half]], 1)
    tex_sw = torch.cat([tex[:, half:], tex[:, :half]], 1)
    rgb2, alpha2, depth2, idx2, _, w2 = RasterizeFunction.apply(faces_sw, tex_sw, S, 0.1, 100.0, 1e-3, (0.25, 0.5, 0.75), True, True, True)
    assert torch.equal(depth2, depth) and torch.equal(alpha2, alpha) and torch.equal(rgb2, rgb)
    moved = torch.where(idx >= half, idx - half, idx + half)
    assert torch.equal(idx2[cov], moved[cov])
    # batch permutation
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0)).to(dev)
    rgb3, _, depth3, idx3, _, _ = RasterizeFunction.apply(faces[perm], tex[perm], S, 0.1, 100.0, 1e-3, (0.25, 0.5, 0.75), True, True, True)
    assert torch.equal(idx3, idx[perm]) and torch.equal(depth3, depth[perm]) and torch.equal(rgb3, rgb[perm])


def test_full_size_backward_properties(det_mode):
    """Reproducible mode: doubling the incoming gradient doubles every term exactly (a power of two), and the
    fixed-point sums do not depend on the order of the additions -- so linearity can be asserted tightly.  (With the
    production float atomics the same comparison carries the run-to-run rounding of sums with heavy cancellation:
    up to ~1e-4 of the gradient scale at this size; test_gpu_determinism.py bounds that noise.)"""
    from handobjectconsist_b200.neurender.rasterize import rasterize_rgbad
    B, S = 16, 256
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, S, seed=43)
    faces0 = _ndc_faces(sc, S, dev)
    Fn = faces0.shape[1]
    tex0 = torch.rand(B, Fn, 2, 2, 2, 3, device=dev)
    gen = torch.Generator().manual_seed(1)
    g_rgb = torch.randn(B, 3, S, S, generator=gen).to(dev)
    g_d = torch.randn(B, S, S, generator=gen).to(dev)

    def grads(scale_rgb, scale_d):
        f = faces0.clone().requires_grad_(True)
        t = tex0.clone().requires_grad_(True)
        o = rasterize_rgbad(f, t, S, False, 0.1, 100.0, 1e-3, (0, 0, 0))
        ((o["rgb"] * g_rgb).sum() * scale_rgb + (o["depth"] * g_d).sum() * scale_d).backward()
        return f.grad, t.grad, o["face_index_map"]

    gf1, gt1, idx = grads(1.0, 1.0)
    gf2, gt2, _ = grads(2.0, 2.0)
    assert torch.isfinite(gf1).all() and torch.isfinite(gt1).all()
    # the texture / depth gradients are linear in the incoming gradient; the pseudo-gradient is positively
    # homogeneous (its delta > 0 gate is scale invariant)
    assert helpers.rel_err(gt2.cpu().numpy(), 2 * gt1.cpu().numpy()) < 1e-6
    assert helpers.rel_err(gf2.cpu().numpy(), 2 * gf1.cpu().numpy()) < 1e-6
    # and a repeated run reproduces the first one bit for bit
    gf3, gt3, _ = grads(1.0, 1.0)
    assert torch.equal(gf3, gf1) and torch.equal(gt3, gt1)
    # faces that own no pixel get exactly zero gradient, and the texture gradient of every face sums the
    # incoming colour gradient over its pixels (trilinear weights sum to one)
    owned = torch.zeros(B, Fn, dtype=torch.bool, device=dev)
    cov = idx >= 0
    bidx = torch.arange(B, device=dev)[:, None, None].expand_as(idx)
    owned[bidx[cov], idx[cov].long()] = True
    assert (gt1[~owned] == 0).all() and (gf1[~owned] == 0).all()
    g_flip = g_rgb.flip(2).permute(0, 2, 3, 1)      # back to raster order, NHWC
    expect = torch.zeros(B, Fn, 3, device=dev)
    expect.index_put_((bidx[cov], idx[cov].long()), g_flip[cov], accumulate=True)
    got = gt1.reshape(B, Fn, 8, 3).sum(2)
    assert helpers.rel_err(got.cpu().numpy(), expect.cpu().numpy()) < 1e-3


def test_rectangular_crop_480x270_consist_step():
    """configs[4]-shaped frames: 480x270 images inside a 480x480 raster, training setting."""
    from handobjectconsist_b200 import warpbranch
    from handobjectconsist_b200.neurender.renderer import Renderer
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
    B, W, H = 4, 480, 270
    S = max(W, H)
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, W, H, seed=44)
    g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}
    r = Renderer(image_size=S, R=torch.eye(3, device=dev)[None], t=torch.zeros(1, 3, device=dev),
                 K=torch.ones(1, 3, 3, device=dev), orig_size=S, anti_aliasing=False, fill_back=True, near=0.1, no_light=True)
    v1 = g["verts1"].clone().requires_grad_(True)
    loss, res = warpbranch.consist_step(v1, g["verts2"], g["faces"], g["K"], g["image_ref"], g["image"],
                                        g["jitter_mask_ref"], g["jitter_mask"], r, PyramidCriterion("l1"), (W, H),
                                        sc["hand_ignore_faces"], detach_renders=True, use_backward=True)
    loss.backward()
    assert res["flows"][0].shape == (B, H, W, 2)
    assert torch.isfinite(loss) and loss.item() > 0
    assert torch.isfinite(v1.grad).all() and v1.grad.abs().max().item() > 0
    # valid pixels only where a flow was rendered, and the loss is the masked mean of the returned diffs
    for i, (img_mask) in enumerate(res["masks"]):
        m = img_mask["full_mask"]
        assert (res["flows"][1 - i][..., 0][m] != 0).all()
        d = res["diffs"][i]
        per = (d * m.unsqueeze(1)).flatten(1).sum(1) / (3 * m.flatten(1).sum(1)).clamp(min=1)
        if i == 0:
            l0 = per
        else:
            assert torch.allclose(l0 + per, res["loss"], atol=1e-5)
