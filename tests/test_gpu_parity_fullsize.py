"""GPU parity against the CPU oracle AT BASELINE.json's SIZES (slices of the configs the bench times):

* configs[1]: hand + object meshes (9104 faces after fill_back) at 256 x 256 -- a 4-sample slice of the 32: forward
  maps identical to the oracle, backward within 1e-3 (rasterize.py:24-197);
* configs[2]: frame pair -> render -> depth-guided warp -> masked L1, 256 x 256 -- a 4-pair slice of the 16: flows /
  loss 1e-4 abs, valid masks exact, vertex gradients 1e-3 (opticalflow.py:51-156, imgflowarp.py:58-115);
* configs[4]: 480 x 480 raster cropped to 480 x 270 -- one pair;
* the whole chain MANO -> mesh -> render -> warp -> loss -> pose gradient at 256 x 256 (SURVEY 8a, a1 -> a14).

The oracle's C restatement runs on all host cores here (seconds per case).  Gradient comparisons are made twice: the
production path (float atomics: bounded against the gradient scale, `max |a - b| <= 1e-3 max |b|`) and the reproducible
mode (order-independent sums: element-wise 1e-3 with the floor of helpers.rel_err) -- the latter's outcome cannot
change from run to run."""
import os

import numpy as np
import pytest
import torch

import helpers
from helpers import onmr
from handobjectconsist_b200 import _lib, synth
from oracle import pipeline as opipe

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _oracle_threads():
    onmr.set_threads(os.cpu_count() or 1)
    yield
    onmr.set_threads(1)


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _renderer(S, dev):
    from handobjectconsist_b200.neurender.renderer import Renderer
    return Renderer(image_size=S, R=torch.eye(3, device=dev)[None], t=torch.zeros(1, 3, device=dev),
                    K=torch.ones(1, 3, 3, device=dev), orig_size=S, anti_aliasing=False, fill_back=True, near=0.1,
                    no_light=True)


def _max_rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


def test_config1_slice_raster_forward_backward_256():
    """4 of configs[1]'s 32 meshes (same generator, same seed) at 256 x 256, all outputs, all incoming gradients."""
    from handobjectconsist_b200.neurender.rasterize import RasterizeFunction
    S, B = 256, 4
    faces, tex, _ = helpers.scene_faces(B, S, seed=0)
    assert faces.shape[1] == 9104
    ora = onmr.rasterize_forward(faces, tex, S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True)
    assert 0.02 < (ora["face_index_map"] >= 0).mean() < 0.5
    rng = np.random.default_rng(0)
    g_rgb = rng.normal(size=ora["rgb_map"].shape).astype(np.float32)
    g_alpha = rng.normal(size=ora["alpha_map"].shape).astype(np.float32)
    g_depth = rng.normal(size=ora["depth_map"].shape).astype(np.float32)
    gf32, gt32 = onmr.rasterize_backward(ora, g_rgb, g_alpha, g_depth)

    def run():
        f = _cuda(faces).requires_grad_(True)
        t = _cuda(tex).requires_grad_(True)
        out = RasterizeFunction.apply(f, t, S, 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True)
        ((out[0] * _cuda(g_rgb)).sum() + (out[1] * _cuda(g_alpha)).sum() + (out[2] * _cuda(g_depth)).sum()).backward()
        return out, f.grad.cpu().numpy(), t.grad.cpu().numpy()

    (rgb, alpha, depth, idx, inv, wmap), gf, gt = run()
    np.testing.assert_array_equal(idx.cpu().numpy(), ora["face_index_map"])
    np.testing.assert_array_equal(depth.detach().cpu().numpy(), ora["depth_map"])
    np.testing.assert_array_equal(wmap.detach().cpu().numpy(), ora["weight_map"])
    np.testing.assert_array_equal(alpha.detach().cpu().numpy(), ora["alpha_map"])
    np.testing.assert_array_equal(rgb.detach().cpu().numpy(), ora["rgb_map"])
    np.testing.assert_array_equal(inv.detach().cpu().numpy(), ora["face_inv_map"])
    # production path (float atomics)
    assert np.isfinite(gf).all() and np.isfinite(gt).all()
    assert _max_rel(gf, gf32) <= 1e-3 and _max_rel(gt, gt32) <= 1e-3
    # reproducible mode: element-wise bar of the north star
    with _lib.deterministic(True):
        _, gf_d, gt_d = run()
        _, gf_d2, gt_d2 = run()
    assert np.array_equal(gf_d, gf_d2) and np.array_equal(gt_d, gt_d2)
    assert helpers.rel_err(gt_d, gt32) < 1e-3
    assert helpers.rel_err(gf_d, gf32) < 1e-3


def _consist_both(sc, S, crop, detach, use_bwd, dev):
    from handobjectconsist_b200 import warpbranch
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
    g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}

    def run():
        v1 = g["verts1"].clone().requires_grad_(True)
        loss, res = warpbranch.consist_step(v1, g["verts2"], g["faces"], g["K"], g["image_ref"], g["image"],
                                            g["jitter_mask_ref"], g["jitter_mask"], _renderer(S, dev),
                                            PyramidCriterion("l1"), crop, sc["hand_ignore_faces"],
                                            detach_renders=detach, use_backward=use_bwd)
        loss.backward()
        return loss, res, v1.grad.cpu().numpy()

    loss, res, grad = run()
    with _lib.deterministic(True):
        _, _, grad_d = run()
        # the fused frame-pair path (consist.py), which warpbranch.forward / the captured step take
        loss_p, res_p, v1_p = helpers.pair_step(sc, S, crop, dev, detach, use_bwd, return_visuals=False)
        loss_p.backward()
    c1 = sc["verts1"].clone().requires_grad_(True)
    loss_o, res_o = opipe.consist_step(c1, sc["verts2"], sc["faces"], sc["K"], sc["image_ref"], sc["image"],
                                       sc["jitter_mask_ref"], sc["jitter_mask"], S, crop, sc["hand_ignore_faces"],
                                       detach_renders=detach, use_backward=use_bwd, warp_device=dev)
    loss_o.backward()
    B = sc["verts1"].shape[0]
    for i in range(2):
        assert res["flows"][i].shape == (B, crop[1], crop[0], 2)
        assert (res["flows"][i].detach() - res_o["flows"][i].detach()).abs().max().item() <= 1e-4
        assert torch.equal(res["masks"][i]["full_mask"], res_o["masks"][i]["full_mask"])
    assert res_o["masks"][0]["full_mask"].float().mean().item() > 0.005
    assert abs(loss.item() - loss_o.item()) <= 1e-4
    go = c1.grad.numpy()
    assert np.abs(go).max() > 0
    assert _max_rel(grad, go) <= 1e-3
    assert helpers.rel_err(grad_d, go) < 1e-3
    for i in range(2):
        assert torch.equal(res_p["flows"][i], res["flows"][i].detach())
        assert torch.equal(res_p["masks"][i]["full_mask"], res_o["masks"][i]["full_mask"])
    assert abs(loss_p.item() - loss_o.item()) <= 1e-4
    assert helpers.rel_err(v1_p.grad.cpu().numpy(), go) < 1e-3


@pytest.mark.parametrize("detach,use_bwd", [(False, True), (True, False)])
def test_config2_slice_consist_step_256(detach, use_bwd):
    """4 of configs[2]'s 16 frame pairs at 256 x 256: the bench's setting (full backward, both directions) and the
    reference's training setting (detach_renders, forward direction only)."""
    dev = torch.device("cuda:0")
    sc = synth.make_scene(4, 256, 256, seed=0)
    _consist_both(sc, 256, (256, 256), detach, use_bwd, dev)


def test_config4_pair_480_raster_270_crop():
    """One of configs[4]'s pairs: 480 x 270 frames inside the 480 x 480 raster (warpreg.py:29,40-45;
    opticalflow.py:152-154)."""
    dev = torch.device("cuda:0")
    W, H = 480, 270
    S = max(W, H)
    sc = synth.make_scene(1, W, H, seed=3)
    _consist_both(sc, S, (W, H), False, True, dev)


def test_mano_to_loss_and_back_256():
    """pose / shape -> ManoLayer -> hand + object mesh -> rendered flows -> warp -> masked L1 at 256 x 256, and back
    to the pose (a1 -> a14).  The checker pushes the oracle pipeline's vertex gradient through the float64 MANO."""
    from handobjectconsist_b200 import warpbranch
    from handobjectconsist_b200.mano.manolayer import ManoLayer
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
    from oracle import mano as omano

    S, B, hv = 256, 2, 778
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, S, seed=7)
    model = synth.mano_model(seed=3)
    dbl = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in model.items()}
    layer = ManoLayer(center_idx=9, flat_hand_mean=False, ncomps=15, use_pca=True, model=model).to(dev)
    g = torch.Generator().manual_seed(5)
    pose = torch.randn(B, 18, generator=g) * 0.4
    betas = torch.randn(B, 10, generator=g) * 0.5
    offset = sc["verts1"][:, :hv].mean(1, keepdim=True)
    gsc = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}

    def run():
        p = pose.to(dev).requires_grad_(True)
        b = betas.to(dev).requires_grad_(True)
        verts_mm, _ = layer(p, th_betas=b)
        v1 = torch.cat([verts_mm / 1000.0 + offset.to(dev), gsc["verts1"][:, hv:]], 1)
        loss, _ = warpbranch.consist_step(v1, gsc["verts2"], gsc["faces"], gsc["K"], gsc["image_ref"], gsc["image"],
                                          gsc["jitter_mask_ref"], gsc["jitter_mask"], _renderer(S, dev),
                                          PyramidCriterion("l1"), (S, S), sc["hand_ignore_faces"], detach_renders=False,
                                          use_backward=True)
        loss.backward()
        return loss, v1.detach(), p.grad.cpu().numpy(), b.grad.cpu().numpy()

    with _lib.deterministic(True):
        loss, v1, gp, gb = run()
    c1 = v1.cpu().clone().requires_grad_(True)
    loss_o, _ = opipe.consist_step(c1, sc["verts2"], sc["faces"], sc["K"], sc["image_ref"], sc["image"],
                                   sc["jitter_mask_ref"], sc["jitter_mask"], S, (S, S), sc["hand_ignore_faces"],
                                   detach_renders=False, use_backward=True, grad_dtype=np.float32, warp_device=dev)
    loss_o.backward()
    assert abs(loss.item() - loss_o.item()) <= 1e-4
    assert c1.grad[:, :hv].abs().max().item() > 0
    po = pose.double().requires_grad_(True)
    bo = betas.double().requires_grad_(True)
    vo, _ = omano.mano_forward(dbl, po, bo, None, True, 9)
    ((vo / 1000.0 + offset.double()) * c1.grad[:, :hv].double()).sum().backward()
    # hundreds of vertex gradients (each within 1e-3 of its scale) add up in every pose / shape coefficient
    assert _max_rel(gp, po.grad.numpy()) <= 5e-3
    assert _max_rel(gb, bo.grad.numpy()) <= 5e-3
