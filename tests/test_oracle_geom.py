"""CPU: the geometry-head oracle (oracle/geom.py) against the golden vectors generated from the REFERENCE's own
meshreg/models/project.py (tests/golden/make_geom_golden.py, run where /root/reference exists) -- bit for bit, values
and gradients -- and the hand-derived adjoint of the kernels (tests/emul/geom_emul.py, the decomposition
csrc/geom_head.cu uses) against float64 autograd of the oracle."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import helpers  # noqa: F401  (sys.path)
from oracle import geom as ogeom
from oracle import mano as omano

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_mk = _load("make_geom_golden", os.path.join(GOLD, "make_geom_golden.py"))
_emul = _load("geom_emul", os.path.join(HERE, "emul", "geom_emul.py"))


@pytest.mark.parametrize("name", sorted(_mk.CASES))
def test_oracle_recover_3d_proj_matches_reference_golden(name):
    gold = np.load(os.path.join(GOLD, "geom_recover3d.npz"))
    B, N, seed, res = _mk.CASES[name]
    pts, K, scale, trans, w_rec, w_c = _mk.inputs(B, N, seed, res)
    pts, scale, trans = [t.clone().requires_grad_(True) for t in (pts, scale, trans)]
    rec, c3d = ogeom.recover_3d_proj(pts, K, scale, trans, input_res=res)
    ((rec * w_rec).sum() + (c3d * w_c).sum()).backward()
    np.testing.assert_array_equal(rec.detach().numpy(), gold[f"{name}_recons3d"])
    np.testing.assert_array_equal(c3d.detach().numpy(), gold[f"{name}_c3d"])
    np.testing.assert_array_equal(pts.grad.numpy(), gold[f"{name}_g_pts"])
    np.testing.assert_array_equal(scale.grad.numpy(), gold[f"{name}_g_scale"])
    np.testing.assert_array_equal(trans.grad.numpy(), gold[f"{name}_g_trans"])


def _camera(B, g, res=(256, 256)):
    f = 300.0 + 400.0 * torch.rand(B, generator=g, dtype=torch.float64)
    K = torch.zeros(B, 3, 3, dtype=torch.float64)
    K[:, 0, 0], K[:, 1, 1], K[:, 2, 2] = f, f * 1.01, 1
    K[:, 0, 1] = 0.3  # a skew term: the kernels use the full matrix
    K[:, 0, 2] = res[0] / 2 + 10 * torch.randn(B, generator=g, dtype=torch.float64)
    K[:, 1, 2] = res[1] / 2 + 10 * torch.randn(B, generator=g, dtype=torch.float64)
    scale = (torch.randn(B, 1, generator=g, dtype=torch.float64) * 2e-4).requires_grad_(True)
    trans = (torch.randn(B, 2, generator=g, dtype=torch.float64) * 30).requires_grad_(True)
    return K, scale, trans


@pytest.mark.parametrize("with_adaptor", [True, False])
def test_hand_head_adjoint_decomposition(with_adaptor):
    g = torch.Generator().manual_seed(3)
    B, V, J, ci, res = 3, 50, 21, 9, (256, 192)
    verts = (torch.randn(B, V, 3, generator=g, dtype=torch.float64) * 0.05).requires_grad_(True)
    joints = (torch.randn(B, J, 3, generator=g, dtype=torch.float64) * 0.05).requires_grad_(True)
    W = torch.rand(J, V, generator=g, dtype=torch.float64) / V if with_adaptor else None
    K, scale, trans = _camera(B, g, res)
    out = ogeom.recover_mano_geometry(verts, joints, K, scale, trans, adaptor_weight=W, center_idx=ci,
                                      trans_factor=100.0, scale_factor=1e-4, input_res=res)
    keys = {"joints3d": "joints3d", "verts3d": "verts3d", "recov_joints3d": "recov_joints3d",
            "recov_verts3d": "recov_handverts3d", "joints2d": "joints2d", "verts2d": "verts2d", "center3d": "center3d"}
    gr = {k: torch.randn(out[o].shape, generator=g, dtype=torch.float64) for k, o in keys.items()}
    sum((out[o] * gr[k]).sum() for k, o in keys.items()).backward()
    gnp = {k: v.numpy().reshape(B, 3) if k == "center3d" else v.numpy() for k, v in gr.items()}
    gv, ga, gs, gt = _emul.hand_head_backward(
        out["recov_handverts3d"].detach().numpy(), out["recov_joints3d"].detach().numpy(),
        None if W is None else W.numpy(), ci if with_adaptor else -1, K.numpy(), scale.detach().numpy().reshape(B),
        trans.detach().numpy(), 1e-4, 100.0, 0.4, res, gnp)
    np.testing.assert_allclose(gv, verts.grad.numpy(), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(gs, scale.grad.numpy().reshape(B), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(gt, trans.grad.numpy(), rtol=1e-9, atol=1e-9)
    if not with_adaptor:
        np.testing.assert_allclose(ga, joints.grad.numpy(), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("with_rot", [True, False])
def test_recover_points_adjoint_decomposition(with_rot):
    g = torch.Generator().manual_seed(5)
    B, N, res = 2, 40, (256, 256)
    pts = (torch.randn(B, N, 3, generator=g, dtype=torch.float64) * 0.05).requires_grad_(True)
    rot = (torch.randn(B, 3, generator=g, dtype=torch.float64) * 0.8).requires_grad_(True)
    K, scale, trans = _camera(B, g, res)
    if with_rot:
        out = ogeom.obj_branch(pts, K, scale, trans, rot, trans_factor=100.0, scale_factor=1e-4, input_res=res)
        outs = {"rot_points": out["obj_verts3d"], "recov_points": out["recov_objverts3d"],
                "points2d": out["obj_verts2d"], "center3d": out["center3d"]}
    else:
        rec, c3d = ogeom.recover_3d_proj(pts, K, scale * 1e-4, trans * 100.0, input_res=res)
        outs = {"recov_points": rec, "center3d": c3d}
    gr = {k: torch.randn(v.shape, generator=g, dtype=torch.float64) for k, v in outs.items()}
    sum((outs[k] * gr[k]).sum() for k in outs).backward()
    gnp = {k: v.numpy().reshape(B, 3) if k == "center3d" else v.numpy() for k, v in gr.items()}
    R = dR = None
    if with_rot:
        R = omano.batch_rodrigues(rot.detach()).reshape(B, 3, 3).numpy()
        jac = torch.autograd.functional.jacobian(lambda r: omano.batch_rodrigues(r).reshape(B, 3, 3), rot.detach())
        dR = np.stack([jac[b, :, :, b].permute(2, 0, 1).numpy() for b in range(B)])  # [B,k,3,3]
    gp, grot, gs, gt = _emul.recover_points_backward(pts.detach().numpy(), R, dR, K.numpy(),
                                                      scale.detach().numpy().reshape(B), trans.detach().numpy(), 1e-4,
                                                      100.0, 0.4, res, gnp)
    np.testing.assert_allclose(gp, pts.grad.numpy(), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(gs, scale.grad.numpy().reshape(B), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(gt, trans.grad.numpy(), rtol=1e-9, atol=1e-9)
    if with_rot:
        np.testing.assert_allclose(grot, rot.grad.numpy(), rtol=1e-9, atol=1e-9)


def test_oracle_geometry_gradcheck():
    g = torch.Generator().manual_seed(7)
    B, N = 2, 6
    pts = (torch.randn(B, N, 3, generator=g, dtype=torch.float64) * 0.05).requires_grad_(True)
    rot = (torch.randn(B, 3, generator=g, dtype=torch.float64) * 0.8).requires_grad_(True)
    K, scale, trans = _camera(B, g)

    def f(p, r, s, t):
        o = ogeom.obj_branch(p, K, s, t, r, trans_factor=100.0, scale_factor=1e-4)
        return o["obj_verts2d"], o["recov_objverts3d"], o["obj_verts3d"]

    assert torch.autograd.gradcheck(f, (pts, rot, scale, trans), eps=1e-7, atol=1e-5)


def test_mano_adaptor_initialisation_follows_the_reference():
    """ManoAdaptor built from a ManoLayer (meshregnet.py:34-45): the 16 regressor rows + 5 one-hot fingertip rows in
    the reference's 21-joint order; built from a pickle (meshregnet.py:27-33): the stored [21,778] matrix."""
    import pickle
    import tempfile

    from handobjectconsist_b200 import synth
    from handobjectconsist_b200.mano.manolayer import ManoLayer
    from handobjectconsist_b200.meshregnet import ManoAdaptor

    model = synth.mano_model(seed=3)
    layer = ManoLayer(center_idx=9, flat_hand_mean=False, ncomps=15, use_pca=True, model=model)
    ad = ManoAdaptor(layer)
    reg = ad.J_regressor
    assert reg.shape == (21, 778) and ad.adaptor.weight.shape == (21, 778)
    for pos, vert in zip((4, 8, 12, 16, 20), (745, 317, 444, 556, 673)):
        row = torch.zeros(778)
        row[vert] = 1
        assert torch.equal(reg[pos], row)
    order = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]
    for pos, src in enumerate(order):
        if src < 16:
            assert torch.equal(reg[pos], model["j_regressor"][src])
    assert torch.equal(ad.pinned_weight().detach(), reg)
    # drifted wrist / fingertip rows are reset by pinned_weight (meshregnet.py:48-50), the others are kept
    with torch.no_grad():
        ad.adaptor.weight.add_(0.5)
    w = ad.pinned_weight().detach()
    assert torch.equal(w[[0, 4, 8, 12, 16, 20]], reg[[0, 4, 8, 12, 16, 20]])
    assert torch.allclose(w[1], reg[1] + 0.5)
    with tempfile.NamedTemporaryFile(suffix=".pkl") as fh:
        stored = np.random.default_rng(0).random((21, 778)).astype(np.float32)
        pickle.dump({"adaptor": stored, "shape": np.zeros(10)}, fh)
        fh.flush()
        ad2 = ManoAdaptor(None, load_path=fh.name)
    assert torch.equal(ad2.J_regressor, torch.from_numpy(stored)) and torch.equal(ad2.adaptor.weight.detach(), ad2.J_regressor)
    with pytest.raises(TypeError):
        ad2(torch.zeros(1, 778, 3))  # CPU tensor: no CPU path


def test_obj_branch_host_logic_with_foreign_query_enums(monkeypatch):
    """ObjBranch's host side (argument handling, sample lookup by member NAME so that a batch keyed by the reference's
    own enums works, result dict) with the kernel call replaced by the oracle -- no GPU involved."""
    from enum import Enum, auto

    from handobjectconsist_b200 import objbranch

    class BaseQueries(Enum):  # stands for meshreg.datasets.queries.BaseQueries: same member names, another class
        OBJCANVERTS = auto()
        OBJCANCORNERS = auto()
        OBJCORNERS3D = auto()

    class TransQueries(Enum):
        IMAGE = auto()
        CAMINTR = auto()

    class FakeFunction:
        @staticmethod
        def apply(points, rot, camintr, scale, trans, scale_factor, trans_factor, off_z, res_w, res_h):
            assert off_z == 0.4 and scale.shape == (points.shape[0],) and trans.shape == (points.shape[0], 2)
            o = ogeom.obj_branch(points, camintr, scale, trans, rot, trans_factor=trans_factor,
                                 scale_factor=scale_factor, input_res=(res_w, res_h))
            return o["obj_verts3d"], o["recov_objverts3d"], o["obj_verts2d"], o["center3d"].reshape(-1, 3)

    monkeypatch.setattr(objbranch, "_RecoverPointsFunction", FakeFunction)
    g = torch.Generator().manual_seed(2)
    B = 3
    can = torch.randn(B, 12, 3, generator=g) * 0.05
    corners = torch.randn(B, 8, 3, generator=g) * 0.05
    K = torch.tensor([[[400.0, 0, 128], [0, 400.0, 96], [0, 0, 1]]]).repeat(B, 1, 1)
    st = torch.cat([torch.randn(B, 1, generator=g), torch.randn(B, 2, generator=g) * 0.3,
                    torch.randn(B, 3, generator=g)], 1)
    sample = {BaseQueries.OBJCANVERTS: can.double(), BaseQueries.OBJCANCORNERS: corners, BaseQueries.OBJCORNERS3D: corners,
              TransQueries.IMAGE: torch.zeros(B, 3, 192, 256), TransQueries.CAMINTR: K}
    branch = objbranch.ObjBranch(trans_factor=100, scale_factor=0.0001)
    want = ogeom.obj_branch(can, K, st[:, :1], st[:, 1:3], st[:, 3:], corners, trans_factor=100, scale_factor=0.0001,
                            input_res=(256, 192))
    for out in (branch(sample, st), branch(sample, None, scale=st[:, :1], trans=st[:, 1:3], rotaxisang=st[:, 3:])):
        assert set(out) == {"obj_verts2d", "obj_verts3d", "recov_objverts3d", "recov_objcorners3d", "obj_scale",
                            "obj_prescale", "obj_prerot", "obj_trans", "obj_pretrans", "obj_corners2d", "obj_corners3d"}
        for k in ("obj_verts2d", "obj_verts3d", "recov_objverts3d", "recov_objcorners3d", "obj_corners2d",
                  "obj_corners3d", "obj_scale", "obj_trans"):
            assert out[k].shape == want[k].shape and torch.allclose(out[k], want[k]), k
        assert out["obj_prescale"].shape == (B, 1) and out["obj_pretrans"].shape == (B, 2) and out["obj_prerot"].shape == (B, 3)
    del sample[BaseQueries.OBJCORNERS3D]
    out = branch(sample, st)
    assert out["obj_corners2d"] is None and out["recov_objcorners3d"] is None and out["obj_corners3d"] is None


@pytest.mark.parametrize("with_adaptor", [True, False])
def test_recover_mano_geometry_host_logic(monkeypatch, with_adaptor):
    """recover_mano_geometry / recover_3d_proj host side (shapes handed to the kernels, centring switch, result keys)
    with the kernel calls replaced by the oracle -- no GPU involved."""
    from handobjectconsist_b200 import meshregnet, project

    class FakeHandHead:
        @staticmethod
        def apply(verts, joints_in, weight, camintr, scale, trans, center_idx, scale_factor, trans_factor, off_z, res_w,
                  res_h):
            B = verts.shape[0]
            assert off_z == 0.4 and scale.shape == (B,) and trans.shape == (B, 2)
            assert (joints_in is None) == (weight is not None) and (center_idx == -1) == (weight is None)
            o = ogeom.recover_mano_geometry(verts, joints_in, camintr, scale, trans, adaptor_weight=weight,
                                            center_idx=center_idx, trans_factor=trans_factor, scale_factor=scale_factor,
                                            input_res=(res_w, res_h))
            return (o["joints3d"], o["verts3d"], o["recov_joints3d"], o["recov_handverts3d"], o["joints2d"],
                    o["verts2d"], o["center3d"].reshape(B, 3))

    class FakeRecover:
        @staticmethod
        def apply(points, rot, camintr, scale, trans, scale_factor, trans_factor, off_z, res_w, res_h):
            assert rot is None and scale_factor == 1.0 and trans_factor == 1.0
            rec, c3d = ogeom.recover_3d_proj(points, camintr, scale, trans, off_z=off_z, input_res=(res_w, res_h))
            return None, rec, None, c3d.reshape(-1, 3)

    monkeypatch.setattr(meshregnet, "_HandHeadFunction", FakeHandHead)
    monkeypatch.setattr(project, "_RecoverPointsFunction", FakeRecover)
    g = torch.Generator().manual_seed(4)
    B = 2
    verts = torch.randn(B, 778, 3, generator=g) * 0.05
    joints = torch.randn(B, 21, 3, generator=g) * 0.05
    W = torch.rand(21, 778, generator=g) / 778 if with_adaptor else None
    K = torch.tensor([[[400.0, 0, 128], [0, 400.0, 96], [0, 0, 1]]]).repeat(B, 1, 1)
    scale, trans = torch.randn(B, 1, generator=g), torch.randn(B, 2, generator=g) * 0.3
    out = meshregnet.recover_mano_geometry({"verts3d": verts, "joints3d": joints, "shape": "kept"}, K, scale, trans,
                                           adaptor=W, mano_center_idx=9, trans_factor=100, scale_factor=0.0001,
                                           input_res=(256, 192))
    want = ogeom.recover_mano_geometry(verts, joints, K, scale, trans, adaptor_weight=W, center_idx=9, trans_factor=100,
                                       scale_factor=0.0001, input_res=(256, 192))
    assert out["shape"] == "kept" and out["hand_pretrans"] is trans and out["hand_prescale"] is scale
    for k in ("joints3d", "verts3d", "joints2d", "recov_joints3d", "recov_handverts3d", "verts2d", "hand_trans",
              "hand_scale"):
        assert out[k].shape == want[k].shape and torch.allclose(out[k], want[k]), k
    # recover_3d_proj: the reference's argument shapes ([B,1,1] scale, [B,1,2] translation) and return shapes
    rec, c3d = project.recover_3d_proj(joints, K, scale.view(B, 1, 1) * 1e-4, trans.unsqueeze(1) * 100, input_res=(256, 192))
    wrec, wc3d = ogeom.recover_3d_proj(joints, K, scale.view(B, 1, 1) * 1e-4, trans.unsqueeze(1) * 100, input_res=(256, 192))
    assert rec.shape == (B, 21, 3) and c3d.shape == (B, 1, 3)
    assert torch.allclose(rec, wrec) and torch.allclose(c3d, wc3d)
