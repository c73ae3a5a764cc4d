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
