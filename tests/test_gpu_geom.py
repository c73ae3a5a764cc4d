"""GPU parity of the geometry-head kernels (hoc_hand_head_*, hoc_recover_points_* behind
``recover_3d_proj``, ``ObjBranch``, ``ManoAdaptor`` and ``recover_mano_geometry``) against

* the golden vectors generated from the REFERENCE's own project.py (values and gradients), and
* oracle/geom.py evaluated in float64 (ManoAdaptor / recover_mano / ObjBranch, whose modules cannot be imported).

Bar: values 2e-5 relative to the largest magnitude of the tensor (fp32: a few ulp; 2-D projections are pixel
COORDINATES of a few hundred, so this is ~5e-3 px), gradients 1e-3 relative (helpers.rel_err, the north star's bar)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import helpers
from oracle import geom as ogeom

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_spec = importlib.util.spec_from_file_location("make_geom_golden", os.path.join(GOLD, "make_geom_golden.py"))
_mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mk)


def _close(got, want, tol=2e-5):
    got = got.detach().cpu().double().numpy()
    want = want.detach().double().numpy() if torch.is_tensor(want) else np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    err = np.abs(got - want).max() if got.size else 0.0
    assert err <= tol * max(np.abs(want).max() if want.size else 0.0, 1e-30), err


def _camera(B, g, res, batched=True):
    n = B if batched else 1
    f = 300.0 + 400.0 * torch.rand(n, generator=g)
    K = torch.zeros(n, 3, 3)
    K[:, 0, 0], K[:, 1, 1], K[:, 2, 2] = f, f * 1.01, 1
    K[:, 0, 1] = 0.3
    K[:, 0, 2] = res[0] / 2 + 10 * torch.randn(n, generator=g)
    K[:, 1, 2] = res[1] / 2 + 10 * torch.randn(n, generator=g)
    scale = torch.randn(B, 1, generator=g) * 1.0
    trans = torch.randn(B, 2, generator=g) * 0.3
    return K, scale, trans


@pytest.mark.parametrize("name", sorted(_mk.CASES))
def test_recover_3d_proj_matches_reference_golden(name):
    from handobjectconsist_b200.project import recover_3d_proj
    gold = np.load(os.path.join(GOLD, "geom_recover3d.npz"))
    B, N, seed, res = _mk.CASES[name]
    pts, K, scale, trans, w_rec, w_c = _mk.inputs(B, N, seed, res)
    pts, scale, trans = [t.cuda().requires_grad_(True) for t in (pts, scale, trans)]
    rec, c3d = recover_3d_proj(pts, K.cuda(), scale, trans, input_res=res)
    assert rec.shape == (B, N, 3) and c3d.shape == (B, 1, 3)
    ((rec * w_rec.cuda()).sum() + (c3d * w_c.cuda()).sum()).backward()
    _close(rec, gold[f"{name}_recons3d"], 1e-6)
    _close(c3d, gold[f"{name}_c3d"], 1e-6)
    assert helpers.rel_err(pts.grad.cpu().numpy(), gold[f"{name}_g_pts"]) < 1e-5
    assert helpers.rel_err(scale.grad.cpu().numpy(), gold[f"{name}_g_scale"]) < 1e-4
    assert helpers.rel_err(trans.grad.cpu().numpy(), gold[f"{name}_g_trans"]) < 1e-4


@pytest.mark.parametrize("with_corners,batched_K,N", [(True, True, 1502), (False, False, 300), (True, True, 1)])
def test_obj_branch_matches_oracle(with_corners, batched_K, N):
    from handobjectconsist_b200.objbranch import ObjBranch
    from handobjectconsist_b200.queries import BaseQueries, TransQueries
    g = torch.Generator().manual_seed(11)
    B, res = 5, (256, 192)  # (width, height)
    can = torch.randn(B, N, 3, generator=g) * 0.05
    corners = torch.randn(B, 8, 3, generator=g) * 0.08
    K, scale, trans = _camera(B, g, res, batched_K)
    rot = torch.randn(B, 3, generator=g) * 0.9
    scaletrans = torch.cat([scale, trans, rot], 1)
    sample = {BaseQueries.OBJCANVERTS: can.double(), TransQueries.IMAGE: torch.zeros(B, 3, res[1], res[0]),
              TransQueries.CAMINTR: K}
    if with_corners:
        sample[BaseQueries.OBJCORNERS3D] = corners
        sample[BaseQueries.OBJCANCORNERS] = corners
    branch = ObjBranch(trans_factor=100, scale_factor=0.0001)
    st = scaletrans.cuda().requires_grad_(True)
    out = branch(sample, st)

    so = scaletrans.double().requires_grad_(True)
    Ko = K.double().expand(B, 3, 3)
    ref = ogeom.obj_branch(can.double(), Ko, so[:, :1], so[:, 1:3], so[:, 3:], corners.double() if with_corners else None,
                           trans_factor=100, scale_factor=0.0001, input_res=res)
    keys = ["obj_verts2d", "obj_verts3d", "recov_objverts3d", "obj_scale", "obj_trans"]
    if with_corners:
        keys += ["recov_objcorners3d", "obj_corners2d", "obj_corners3d"]
    else:
        assert out["obj_corners2d"] is None and out["recov_objcorners3d"] is None and out["obj_corners3d"] is None
    assert set(out) == {"obj_verts2d", "obj_verts3d", "recov_objverts3d", "recov_objcorners3d", "obj_scale",
                        "obj_prescale", "obj_prerot", "obj_trans", "obj_pretrans", "obj_corners2d", "obj_corners3d"}
    for k in keys:
        _close(out[k], ref[k])
    diff = [k for k in keys if k not in ("obj_scale", "obj_trans")]
    ws = {k: torch.randn(ref[k].shape, generator=g, dtype=torch.float64) for k in diff}
    sum((ref[k] * ws[k]).sum() for k in diff).backward()
    sum((out[k] * ws[k].float().cuda()).sum() for k in diff).backward()
    assert helpers.rel_err(st.grad.cpu().numpy(), so.grad.numpy()) < 1e-3


@pytest.mark.parametrize("with_adaptor,batched_K", [(True, True), (False, True), (True, False)])
def test_recover_mano_geometry_matches_oracle(with_adaptor, batched_K):
    from handobjectconsist_b200.meshregnet import recover_mano_geometry
    g = torch.Generator().manual_seed(13)
    B, V, J, ci, res = 6, 778, 21, 9, (256, 256)
    verts = torch.randn(B, V, 3, generator=g) * 0.05
    joints = torch.randn(B, J, 3, generator=g) * 0.05
    W = torch.rand(J, V, generator=g) * (torch.rand(J, V, generator=g) > 0.9).float()
    W = W / W.sum(1, keepdim=True)
    K, scale, trans = _camera(B, g, res, batched_K)
    vc, jc, sc, tc = [t.cuda().requires_grad_(True) for t in (verts, joints, scale, trans)]
    out = recover_mano_geometry({"verts3d": vc, "joints3d": jc, "pose": None}, K.cuda(), sc, tc,
                                adaptor=W.cuda() if with_adaptor else None, mano_center_idx=ci, trans_factor=100,
                                scale_factor=0.0001, input_res=res)
    vo, jo, so, to = [t.double().requires_grad_(True) for t in (verts, joints, scale, trans)]
    ref = ogeom.recover_mano_geometry(vo, jo, K.double().expand(B, 3, 3), so, to,
                                      adaptor_weight=W.double() if with_adaptor else None, center_idx=ci,
                                      trans_factor=100, scale_factor=0.0001, input_res=res)
    keys = ["joints3d", "verts3d", "joints2d", "recov_joints3d", "recov_handverts3d", "verts2d", "hand_trans",
            "hand_scale"]
    assert "pose" in out and out["hand_pretrans"] is tc and out["hand_prescale"] is sc
    for k in keys:
        _close(out[k], ref[k])
    diff = keys[:6]
    ws = {k: torch.randn(ref[k].shape, generator=g, dtype=torch.float64) for k in diff}
    sum((ref[k] * ws[k]).sum() for k in diff).backward()
    sum((out[k] * ws[k].float().cuda()).sum() for k in diff).backward()
    assert helpers.rel_err(vc.grad.cpu().numpy(), vo.grad.numpy()) < 1e-3
    assert helpers.rel_err(sc.grad.cpu().numpy(), so.grad.numpy()) < 1e-3
    assert helpers.rel_err(tc.grad.cpu().numpy(), to.grad.numpy()) < 1e-3
    if with_adaptor:
        assert jc.grad is None  # the adapted joints replace the MANO joints (meshregnet.py:192-195)
    else:
        assert helpers.rel_err(jc.grad.cpu().numpy(), jo.grad.numpy()) < 1e-3


def test_recover_mano_partial_gradients_and_adaptor_module():
    """Only one output used (the 2-D joints, as with mano_lambda_joints2d alone), and ManoAdaptor's own forward
    (meshregnet.py:47-51) incl. the gradient of an unfrozen weight."""
    from handobjectconsist_b200.meshregnet import ManoAdaptor, recover_mano_geometry
    from handobjectconsist_b200 import synth
    from handobjectconsist_b200.mano.manolayer import ManoLayer
    g = torch.Generator().manual_seed(17)
    B, res = 3, (256, 256)
    layer = ManoLayer(center_idx=9, flat_hand_mean=False, ncomps=15, use_pca=True, model=synth.mano_model(seed=3))
    adaptor = ManoAdaptor(layer).cuda()
    assert adaptor.J_regressor.shape == (21, 778)
    verts = torch.randn(B, 778, 3, generator=g) * 0.05
    K, scale, trans = _camera(B, g, res)
    vc = verts.cuda().requires_grad_(True)
    out = recover_mano_geometry({"verts3d": vc, "joints3d": torch.zeros(B, 21, 3, device="cuda")}, K.cuda(),
                                scale.cuda(), trans.cuda(), adaptor=adaptor, trans_factor=100, scale_factor=0.0001,
                                input_res=res)
    w2 = torch.randn(B, 21, 2, generator=g, dtype=torch.float64)
    (out["joints2d"] * w2.float().cuda()).sum().backward()
    vo = verts.double().requires_grad_(True)
    Wd = adaptor.J_regressor.double().cpu()
    ref = ogeom.recover_mano_geometry(vo, None, K.double(), scale.double(), trans.double(), adaptor_weight=Wd,
                                      center_idx=9, trans_factor=100, scale_factor=0.0001, input_res=res)
    (ref["joints2d"] * w2).sum().backward()
    _close(out["joints2d"], ref["joints2d"])
    assert helpers.rel_err(vc.grad.cpu().numpy(), vo.grad.numpy()) < 1e-3

    # the module's own forward: [B,3,21] joints and the weight drift; weight gradient when not frozen
    adaptor.adaptor.weight.requires_grad_(True)
    adaptor.adaptor.weight.grad = None
    vc2 = verts.cuda().requires_grad_(True)
    joints, drift = adaptor(vc2)
    assert joints.shape == (B, 3, 21) and float(drift.abs().max()) == 0.0
    wj = torch.randn(B, 3, 21, generator=g, dtype=torch.float64)
    (joints * wj.float().cuda()).sum().backward()
    Wo = Wd.clone().requires_grad_(True)
    vo2 = verts.double().requires_grad_(True)
    jo = ogeom.mano_adaptor(Wo, vo2).transpose(1, 2)
    (jo * wj).sum().backward()
    _close(joints, jo)
    assert helpers.rel_err(vc2.grad.cpu().numpy(), vo2.grad.numpy()) < 1e-3
    assert helpers.rel_err(adaptor.adaptor.weight.grad.cpu().numpy(), Wo.grad.numpy()) < 1e-3


def test_geometry_head_rejects_cpu_tensors_and_bad_shapes():
    from handobjectconsist_b200.meshregnet import recover_mano_geometry
    from handobjectconsist_b200.project import recover_3d_proj
    K = torch.eye(3)[None]
    with pytest.raises(TypeError):
        recover_3d_proj(torch.zeros(1, 4, 3), K, torch.zeros(1, 1, 1), torch.zeros(1, 1, 2))
    with pytest.raises(ValueError):
        recover_3d_proj(torch.zeros(2, 4, 3).cuda(), torch.eye(3)[None].repeat(3, 1, 1).cuda(),
                        torch.zeros(2, 1, 1).cuda(), torch.zeros(2, 1, 2).cuda())
    with pytest.raises(ValueError):
        recover_mano_geometry({"verts3d": torch.zeros(1, 10, 3).cuda(), "joints3d": torch.zeros(1, 21, 3).cuda()},
                              K.cuda(), torch.zeros(1, 1).cuda(), torch.zeros(1, 2).cuda(),
                              adaptor=torch.zeros(21, 778).cuda())


def test_network_outputs_to_photometric_loss_and_back():
    """Rows f1 -> a1 -> a14 chained: pose / shape -> ManoLayer -> ManoAdaptor + recover_3d_proj (hand head), object
    scale / translation / rotation -> ObjBranch, both meshes -> warpbranch's consist step -> masked L1, and back to
    every network output.  Checked in two links: (1) loss and the gradient that reaches the mesh vertices against the
    oracle's pipeline run on the same vertices; (2) that vertex gradient pushed through the oracle's float64 front end
    (MANO, adaptor, recover_3d_proj, Rodrigues) against the gradients the kernels return for the network outputs."""
    from handobjectconsist_b200 import synth, warpbranch
    from handobjectconsist_b200.mano.manolayer import ManoLayer
    from handobjectconsist_b200.meshregnet import ManoAdaptor, recover_mano_geometry
    from handobjectconsist_b200.neurender.renderer import Renderer
    from handobjectconsist_b200.objbranch import ObjBranch
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
    from handobjectconsist_b200.queries import BaseQueries, TransQueries
    from oracle import mano as omano
    from oracle import pipeline as opipe

    S, B, hv, sf, tf = 64, 2, 778, 1e-4, 100.0
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, S, seed=7)
    model = synth.mano_model(seed=3)
    layer = ManoLayer(center_idx=9, flat_hand_mean=False, ncomps=15, use_pca=True, model=model).to(dev)
    adaptor = ManoAdaptor(layer).to(dev)
    for p_ in adaptor.parameters():
        p_.requires_grad_(False)  # rec_freeze(self.adaptor), meshregnet.py:147
    g = torch.Generator().manual_seed(5)
    pose = torch.randn(B, 18, generator=g) * 0.4
    betas = torch.randn(B, 10, generator=g) * 0.5
    K = sc["K"]

    def head_inputs(centre):
        """scale / translation (network units) that put est_c3d at `centre` (inverse of project.py:15-20)."""
        f, cc = K[:, 0, 0], K[:, :2, 2]
        s = (centre[:, 2] - 0.4) / (f * sf)
        t = (centre[:, :2] * (f / centre[:, 2])[:, None] - S / 2.0 + cc) / tf
        return torch.cat([s[:, None], t], 1)

    hand_st = head_inputs(sc["verts1"][:, :hv].mean(1))
    obj_centre = sc["verts1"][:, hv:].mean(1)
    can = sc["verts1"][:, hv:] - obj_centre[:, None]
    obj_st = torch.cat([head_inputs(obj_centre), torch.randn(B, 3, generator=g) * 0.2], 1)

    p, b, hs, os_ = [t.to(dev).requires_grad_(True) for t in (pose, betas, hand_st, obj_st)]
    verts_mm, joints_mm = layer(p, th_betas=b)
    mano_results = recover_mano_geometry({"verts3d": verts_mm / 1000, "joints3d": joints_mm / 1000}, K.to(dev),
                                         hs[:, :1], hs[:, 1:], adaptor=adaptor, mano_center_idx=9, trans_factor=tf,
                                         scale_factor=sf, input_res=(S, S))
    sample = {BaseQueries.OBJCANVERTS: can, TransQueries.IMAGE: sc["image"], TransQueries.CAMINTR: K}
    obj_results = ObjBranch(trans_factor=tf, scale_factor=sf)(sample, os_)
    v1 = torch.cat([mano_results["recov_handverts3d"], obj_results["recov_objverts3d"]], 1)
    v1.retain_grad()
    gsc = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}
    renderer = Renderer(image_size=S, R=torch.eye(3, device=dev)[None], t=torch.zeros(1, 3, device=dev),
                        K=torch.ones(1, 3, 3, device=dev), orig_size=S, anti_aliasing=False, fill_back=True, near=0.1,
                        no_light=True)
    loss, _ = warpbranch.consist_step(v1, gsc["verts2"], gsc["faces"], gsc["K"], gsc["image_ref"], gsc["image"],
                                      gsc["jitter_mask_ref"], gsc["jitter_mask"], renderer, PyramidCriterion("l1"),
                                      (S, S), sc["hand_ignore_faces"], detach_renders=False, use_backward=True)
    loss.backward()

    # link 1: same vertices into the oracle's pipeline
    c1 = v1.detach().cpu().clone().requires_grad_(True)
    loss_o, _ = opipe.consist_step(c1, sc["verts2"], sc["faces"], sc["K"], sc["image_ref"], sc["image"],
                                   sc["jitter_mask_ref"], sc["jitter_mask"], S, (S, S), sc["hand_ignore_faces"],
                                   detach_renders=False, use_backward=True, grad_dtype=np.float32, warp_device=dev)
    loss_o.backward()
    assert abs(loss.item() - loss_o.item()) <= 1e-4
    gv = v1.grad.cpu()
    assert gv[:, :hv].abs().max().item() > 0 and gv[:, hv:].abs().max().item() > 0
    assert (gv - c1.grad).abs().max().item() <= 1e-3 * c1.grad.abs().max().item()

    # link 2: the kernels' vertex gradient through the oracle's front end, float64
    dbl = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in model.items()}
    po, bo, hso, oso = [t.double().requires_grad_(True) for t in (pose, betas, hand_st, obj_st)]
    vo, jo = omano.mano_forward(dbl, po, bo, None, True, 9)
    ref_h = ogeom.recover_mano_geometry(vo / 1000, jo / 1000, K.double(), hso[:, :1], hso[:, 1:],
                                        adaptor_weight=adaptor.J_regressor.double().cpu(), center_idx=9,
                                        trans_factor=tf, scale_factor=sf, input_res=(S, S))
    ref_o = ogeom.obj_branch(can.double(), K.double(), oso[:, :1], oso[:, 1:3], oso[:, 3:], trans_factor=tf,
                             scale_factor=sf, input_res=(S, S))
    vref = torch.cat([ref_h["recov_handverts3d"], ref_o["recov_objverts3d"]], 1)
    _close(v1, vref, 2e-5)  # the MANO kernels' bar (test_gpu_mano.py: 1e-4 of the hand size in mm)
    (vref * gv.double()).sum().backward()
    for got, want in ((p, po), (b, bo), (hs, hso), (os_, oso)):
        assert helpers.rel_err(got.grad.cpu().numpy(), want.grad.numpy()) < 1e-3


def test_graphed_step_with_the_head_inside_matches_eager(det_mode):
    """GraphedHeadConsistStep: MANO + ManoAdaptor + recover_3d_proj + ObjBranch + the consistency step, forward and
    backward, captured in ONE CUDA graph -- same loss and the same gradients of the network outputs as the eager
    chain, also after loading other network outputs into its static buffers."""
    from handobjectconsist_b200 import synth, warpbranch
    from handobjectconsist_b200.graphed import GraphedHeadConsistStep
    from handobjectconsist_b200.mano.manolayer import ManoLayer
    from handobjectconsist_b200.meshregnet import recover_mano_geometry
    from handobjectconsist_b200.neurender.renderer import Renderer
    from handobjectconsist_b200.objbranch import ObjBranch
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
    from handobjectconsist_b200.queries import BaseQueries, TransQueries

    S, B, hv, sf, tf = 64, 2, 778, 1e-4, 100.0
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, S, seed=20)
    layer = ManoLayer(center_idx=9, flat_hand_mean=False, ncomps=15, use_pca=True, model=synth.mano_model(seed=3)).to(dev)
    g = torch.Generator().manual_seed(9)
    W = torch.rand(21, hv, generator=g) * (torch.rand(21, hv, generator=g) > 0.9).float()
    W = (W / W.sum(1, keepdim=True)).to(dev)
    K = sc["K"]
    f, cc = K[:, 0, 0], K[:, :2, 2]

    def head_units(centre):
        s = (centre[:, 2] - 0.4) / (f * sf)
        t = (centre[:, :2] * (f / centre[:, 2])[:, None] - S / 2.0 + cc) / tf
        return torch.cat([s[:, None], t], 1)

    obj_centre = sc["verts1"][:, hv:].mean(1)
    can = (sc["verts1"][:, hv:] - obj_centre[:, None]).to(dev)
    Kd = K.to(dev)
    obj_branch = ObjBranch(trans_factor=tf, scale_factor=sf)
    shape_only = torch.empty(0, 0, S, S)

    def head(inp):
        verts_mm, joints_mm = layer(inp["pose"], th_betas=inp["betas"])
        hand = recover_mano_geometry({"verts3d": verts_mm / 1000, "joints3d": joints_mm / 1000}, Kd,
                                     inp["hand_st"][:, :1], inp["hand_st"][:, 1:], adaptor=W, mano_center_idx=9,
                                     trans_factor=tf, scale_factor=sf, input_res=(S, S))["recov_handverts3d"]
        sample = {BaseQueries.OBJCANVERTS: can, TransQueries.IMAGE: shape_only, TransQueries.CAMINTR: Kd}
        return hand, obj_branch(sample, inp["obj_st"])["recov_objverts3d"]

    def net_outputs(seed):
        gg = torch.Generator().manual_seed(seed)
        return {"pose": torch.randn(B, 18, generator=gg) * 0.4, "betas": torch.randn(B, 10, generator=gg) * 0.5,
                "hand_st": head_units(sc["verts1"][:, :hv].mean(1)) + torch.randn(B, 3, generator=gg) * 0.02,
                "obj_st": torch.cat([head_units(obj_centre), torch.randn(B, 3, generator=gg) * 0.2], 1)}

    mv = lambda t: t.to(dev)
    obj_faces = mv(sc["faces"][:, 1552:] - hv)
    samples = []
    for verts, img, jit in ((sc["verts1"], sc["image_ref"], sc["jitter_mask_ref"]),
                            (sc["verts2"], sc["image"], sc["jitter_mask"])):
        samples.append({TransQueries.IMAGE: mv(img), TransQueries.JITTERMASK: mv(jit), TransQueries.CAMINTR: Kd,
                        BaseQueries.OBJFACES: obj_faces, BaseQueries.OBJVERTS3D: mv(verts[:, hv:]),
                        BaseQueries.HANDVERTS3D: mv(verts[:, :hv])})
    hand_face = sc["faces"][0, :1552].to(dev)
    crit = PyramidCriterion("l1")
    # the second frame's entry is read and then replaced by ground truth (gt_refs, warpbranch.py:38-44)
    second = {"recov_handverts3d": mv(sc["verts2"][:, :hv]), "recov_objverts3d": mv(sc["verts2"][:, hv:])}

    def renderer():
        return Renderer(image_size=S, R=torch.eye(3, device=dev)[None], t=torch.zeros(1, 3, device=dev),
                        K=torch.ones(1, 3, 3, device=dev), orig_size=S, anti_aliasing=False, fill_back=True, near=0.1,
                        no_light=True)

    def eager(outs):
        inp = {k: v.to(dev).requires_grad_(True) for k, v in outs.items()}
        hand, obj = head(inp)
        loss, _ = warpbranch.forward(samples, [{"recov_handverts3d": hand, "recov_objverts3d": obj}, second], hand_face,
                                     renderer(), (S, S), crit, hand_ignore_faces=sc["hand_ignore_faces"],
                                     detach_renders=False)
        loss.backward()
        return loss.detach(), {k: v.grad for k, v in inp.items()}, hand.detach(), obj.detach()

    first = net_outputs(1)
    _, _, hand0, obj0 = eager(first)
    gstep = GraphedHeadConsistStep(head, first, renderer(), crit, (S, S), hand_face, samples,
                                   [{"recov_handverts3d": hand0, "recov_objverts3d": obj0}, second],
                                   hand_ignore_faces=sc["hand_ignore_faces"], detach_renders=False)
    for seed in (1, 2, 3):
        outs = net_outputs(seed)
        loss_g, grads_g = gstep(samples, [{}, second], outs)
        loss_g, grads_g = loss_g.clone(), {k: v.clone() for k, v in grads_g.items()}
        loss_e, grads_e, _, _ = eager(outs)
        assert loss_e.item() > 0 and abs(loss_g.item() - loss_e.item()) <= 1e-6
        assert set(grads_g) == {"pose", "betas", "hand_st", "obj_st"}
        for k in grads_g:
            assert grads_e[k].abs().max().item() > 0
            # same kernels on both sides, reproducible mode (order-independent sums): bit-level agreement
            assert helpers.rel_err(grads_g[k].cpu().numpy(), grads_e[k].cpu().numpy()) < 1e-6, k
