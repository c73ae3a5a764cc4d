"""GPU parity of the MANO skinning kernels (ManoLayer -> hoc_mano_forward / hoc_mano_backward) against
oracle/mano.py, the torch restatement of manopth's ManoLayer.forward, on the synthetic MANO-shaped model.
Bar: vertices / joints 1e-4 relative to the hand size in mm (fp32), gradients 1e-3 relative (vs float64 autograd)."""
import numpy as np
import pytest
import torch

import helpers
from handobjectconsist_b200 import synth
from oracle import mano as omano

pytestmark = pytest.mark.gpu


def _dbl(model):
    return {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in model.items()}


@pytest.mark.parametrize("use_pca,ncomps,center,flat,side,with_betas,with_trans", [
    (True, 15, 9, False, "right", True, False),    # ManoBranch (manobranch.py:70-85)
    (True, 15, 9, False, "left", False, False),
    (True, 6, None, True, "right", True, False),
    (False, 45, None, True, "right", True, False),  # WarpRegNet's layer (warpreg.py:54-60)
    (True, 15, 20, False, "right", True, False),   # centred on a fingertip joint
    (True, 15, 9, False, "right", True, True),     # explicit translation
])
def test_mano_matches_oracle(use_pca, ncomps, center, flat, side, with_betas, with_trans):
    from handobjectconsist_b200.mano.manolayer import ManoLayer, TIP_IDS
    model = synth.mano_model(seed=3)
    model["tip_ids"] = TIP_IDS[side]
    B = 5
    g = torch.Generator().manual_seed(1)
    pose = torch.randn(B, 3 + ncomps, generator=g) * 0.6
    betas = torch.randn(B, 10, generator=g) * 0.8 if with_betas else None
    trans = torch.randn(B, 3, generator=g) * 0.1 if with_trans else None
    layer = ManoLayer(center_idx=center, flat_hand_mean=flat, ncomps=ncomps, side=side, use_pca=use_pca, model=model).cuda()
    assert layer.th_faces.shape == (1538, 3)
    p = pose.cuda().requires_grad_(True)
    b = betas.cuda().requires_grad_(True) if with_betas else None
    t = trans.cuda().requires_grad_(True) if with_trans else None
    verts, joints = layer(p, th_betas=b if with_betas else torch.Tensor([0]), th_trans=t if with_trans else torch.Tensor([0]))

    md = _dbl(model)
    if flat:
        md["hands_mean"] = torch.zeros(45, dtype=torch.float64)
    po = pose.double().requires_grad_(True)
    bo = betas.double().requires_grad_(True) if with_betas else None
    to = trans.double().requires_grad_(True) if with_trans else None
    vo, jo = omano.mano_forward(md, po, bo, to, use_pca, center)
    scale = vo.abs().max().item()
    assert (verts.detach().cpu().double() - vo.detach()).abs().max().item() <= 1e-4 * scale
    assert (joints.detach().cpu().double() - jo.detach()).abs().max().item() <= 1e-4 * scale

    gv = torch.randn(vo.shape, generator=g, dtype=torch.float64)
    gj = torch.randn(jo.shape, generator=g, dtype=torch.float64)
    ((vo * gv).sum() + (jo * gj).sum()).backward()
    ((verts * gv.float().cuda()).sum() + (joints * gj.float().cuda()).sum()).backward()
    assert helpers.rel_err(p.grad.cpu().numpy(), po.grad.numpy()) < 1e-3
    if with_betas:
        assert helpers.rel_err(b.grad.cpu().numpy(), bo.grad.numpy()) < 1e-3
    if with_trans:
        assert helpers.rel_err(t.grad.cpu().numpy(), to.grad.numpy()) < 1e-3


def test_mano_only_joint_gradient_and_errors():
    from handobjectconsist_b200.mano.manolayer import ManoLayer
    model = synth.mano_model(seed=3)
    layer = ManoLayer(center_idx=9, flat_hand_mean=False, ncomps=15, use_pca=True, model=model).cuda()
    pose = (torch.randn(3, 18, generator=torch.Generator().manual_seed(2)) * 0.5)
    p = pose.cuda().requires_grad_(True)
    verts, joints = layer(p)
    joints[:, 4].sum().backward()
    po = pose.double().requires_grad_(True)
    _, jo = omano.mano_forward(_dbl(model), po, None, None, True, 9)
    jo[:, 4].sum().backward()
    assert helpers.rel_err(p.grad.cpu().numpy(), po.grad.numpy()) < 1e-3
    with pytest.raises(TypeError):
        layer(pose)  # CPU tensor
    with pytest.raises(ValueError):
        layer(torch.zeros(2, 7, device="cuda"))
    with pytest.raises(FileNotFoundError):
        ManoLayer(mano_root="/nonexistent")


def test_mano_through_consist_step_pose_gradient():
    """The whole path of SURVEY section 8a, a1 -> a14: pose / shape -> ManoLayer -> metres + translation -> hand + object
    mesh -> rendered flows -> warp -> masked L1, and back to the pose.  The checker runs the oracle's pipeline on the
    SAME hand vertices (the kernels' own, so that both sides rasterise the same geometry) and pushes its vertex
    gradient through the oracle's MANO."""
    from handobjectconsist_b200 import warpbranch
    from handobjectconsist_b200.mano.manolayer import ManoLayer
    from handobjectconsist_b200.neurender.renderer import Renderer
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
    from oracle import pipeline as opipe

    S, B, hv = 64, 2, 778
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, S, seed=7)
    model = synth.mano_model(seed=3)
    layer = ManoLayer(center_idx=9, flat_hand_mean=False, ncomps=15, use_pca=True, model=model).to(dev)
    g = torch.Generator().manual_seed(5)
    pose = torch.randn(B, 18, generator=g) * 0.4
    betas = torch.randn(B, 10, generator=g) * 0.5
    # put the MANO hand where the scene's hand template sits
    offset = sc["verts1"][:, :hv].mean(1, keepdim=True)

    def hand_from(verts_mm, off):
        return verts_mm / 1000.0 + off

    p = pose.to(dev).requires_grad_(True)
    b = betas.to(dev).requires_grad_(True)
    verts_mm, _ = layer(p, th_betas=b)
    hand = hand_from(verts_mm, offset.to(dev))
    gsc = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}
    v1 = torch.cat([hand, gsc["verts1"][:, hv:]], 1)
    renderer = Renderer(image_size=S, R=torch.eye(3, device=dev)[None], t=torch.zeros(1, 3, device=dev),
                        K=torch.ones(1, 3, 3, device=dev), orig_size=S, anti_aliasing=False, fill_back=True, near=0.1,
                        no_light=True)
    loss, _ = warpbranch.consist_step(v1, gsc["verts2"], gsc["faces"], gsc["K"], gsc["image_ref"], gsc["image"],
                                      gsc["jitter_mask_ref"], gsc["jitter_mask"], renderer, PyramidCriterion("l1"),
                                      (S, S), sc["hand_ignore_faces"], detach_renders=False, use_backward=True)
    loss.backward()

    # checker: same vertices into the oracle pipeline, its vertex gradient through the oracle's MANO (float64)
    c1 = v1.detach().cpu().clone().requires_grad_(True)
    loss_o, _ = opipe.consist_step(c1, sc["verts2"], sc["faces"], sc["K"], sc["image_ref"], sc["image"],
                                   sc["jitter_mask_ref"], sc["jitter_mask"], S, (S, S), sc["hand_ignore_faces"],
                                   detach_renders=False, use_backward=True, grad_dtype=np.float32, warp_device=dev)
    loss_o.backward()
    assert abs(loss.item() - loss_o.item()) <= 1e-4
    assert c1.grad[:, :hv].abs().max().item() > 0
    po = pose.double().requires_grad_(True)
    bo = betas.double().requires_grad_(True)
    vo, _ = omano.mano_forward(_dbl(model), po, bo, None, True, 9)
    (hand_from(vo, offset.double()) * c1.grad[:, :hv].double()).sum().backward()
    for got, want in ((p.grad.cpu().numpy(), po.grad.numpy()), (b.grad.cpu().numpy(), bo.grad.numpy())):
        # the vertex gradients of the two sides agree to 1e-3 of their maximum (float atomics, fp32 vs fp32);
        # hundreds of them add up in every pose / shape coefficient
        assert np.abs(got - want).max() <= 5e-3 * np.abs(want).max()
