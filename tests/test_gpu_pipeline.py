"""GPU end-to-end parity: vertices -> rendered flows -> occlusion -> warp -> masked L1 -> backward, through
the reference-shaped surface (Renderer, get_opticalflow, pair_consist, warpbranch.forward), against the
oracle pipeline (C restatement of the rasterizer on the CPU + the reference's warp ops run by ATen on
CUDA, where the reference runs them).  Bar: flows / loss 1e-4 abs, vertex gradients 1e-3 relative."""
import numpy as np
import pytest
import torch

import helpers
from handobjectconsist_b200 import synth
from oracle import pipeline as opipe

pytestmark = pytest.mark.gpu


def _renderer(S, dev):
    from handobjectconsist_b200.neurender.renderer import Renderer
    return Renderer(image_size=S, R=torch.eye(3, device=dev)[None], t=torch.zeros(1, 3, device=dev),
                    K=torch.ones(1, 3, 3, device=dev), orig_size=S, anti_aliasing=False, fill_back=True, near=0.1,
                    no_light=True)


@pytest.mark.parametrize("S,crop,detach,use_bwd,seed", [
    (64, (64, 64), False, True, 0),
    (64, (64, 64), True, False, 1),     # the reference's training setting (warpbranch.py:59-68, CLI default)
    (96, (96, 54), True, True, 2),      # rectangular image inside a square raster
    (128, (128, 128), False, True, 3),
])
def test_consist_step_matches_oracle(S, crop, detach, use_bwd, seed):
    from handobjectconsist_b200 import warpbranch
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
    B = 2
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, crop[0], crop[1], seed=seed)
    sc["K"] = synth.camera_intrinsics(B, S, S)  # camera of the square raster
    g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}
    v1 = g["verts1"].clone().requires_grad_(True)
    loss, res = warpbranch.consist_step(v1, g["verts2"], g["faces"], g["K"], g["image_ref"], g["image"],
                                        g["jitter_mask_ref"], g["jitter_mask"], _renderer(S, dev),
                                        PyramidCriterion("l1"), crop, sc["hand_ignore_faces"], detach_renders=detach,
                                        use_backward=use_bwd)
    loss.backward()
    c1 = sc["verts1"].clone().requires_grad_(True)
    loss_o, res_o = opipe.consist_step(c1, sc["verts2"], sc["faces"], sc["K"], sc["image_ref"], sc["image"],
                                       sc["jitter_mask_ref"], sc["jitter_mask"], S, crop, sc["hand_ignore_faces"],
                                       detach_renders=detach, use_backward=use_bwd, warp_device=dev)
    loss_o.backward()
    for i in range(2):
        assert res["flows"][i].shape == (B, crop[1], crop[0], 2)
        assert (res["flows"][i].detach() - res_o["flows"][i].detach()).abs().max().item() <= 1e-4
        assert torch.equal(res["masks"][i]["full_mask"], res_o["masks"][i]["full_mask"])
    assert res_o["masks"][0]["full_mask"].float().mean().item() > 0.005
    assert abs(loss.item() - loss_o.item()) <= 1e-4
    go = c1.grad.numpy()
    assert np.abs(go).max() > 0
    assert helpers.rel_err(v1.grad.cpu().numpy(), go) < 1e-3


@pytest.mark.parametrize("S,crop,detach,use_bwd,visuals,seed", [
    (64, (64, 64), False, True, True, 0),
    (64, (64, 64), True, False, True, 1),    # the reference's training setting
    (96, (96, 52), True, True, False, 2),    # rectangular crop, no visualisation returns
    (128, (128, 128), False, True, False, 3),
    (64, (64, 64), False, False, True, 4),   # geometry gradient without the backward direction
    (96, (96, 52), False, True, True, 5),    # raster row window + geometry gradient: the meshes reach below the crop
    (128, (128, 72), False, True, False, 6),
])
def test_pair_path_matches_oracle(S, crop, detach, use_bwd, visuals, seed):
    """The fused frame-pair path (consist.py: both renders stacked along the batch, one autograd node) against the
    oracle pipeline: flows / loss 1e-4, every mask exact, warps / diffs 1e-6, vertex gradients 1e-3."""
    B = 2
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, crop[0], crop[1], seed=seed)
    sc["K"] = synth.camera_intrinsics(B, S, S)
    loss, res, v1 = helpers.pair_step(sc, S, crop, dev, detach, use_bwd, visuals)
    loss.backward()
    c1 = sc["verts1"].clone().requires_grad_(True)
    loss_o, res_o = opipe.consist_step(c1, sc["verts2"], sc["faces"], sc["K"], sc["image_ref"], sc["image"],
                                       sc["jitter_mask_ref"], sc["jitter_mask"], S, crop, sc["hand_ignore_faces"],
                                       detach_renders=detach, use_backward=use_bwd, warp_device=dev)
    loss_o.backward()
    for i in range(2):
        assert res["flows"][i].shape == (B, crop[1], crop[0], 2)
        assert (res["flows"][i].detach() - res_o["flows"][i].detach()).abs().max().item() <= 1e-4
        assert torch.equal(res["masks"][i]["full_mask"], res_o["masks"][i]["full_mask"])
        assert torch.equal(res["masks"][i]["flow_mask"], res_o["masks"][i]["flow_mask"])
        if visuals:
            assert torch.equal(res["masks"][i]["warp_mask"], res_o["masks"][i]["warp_mask"])
            assert (res["warps"][i] - res_o["warps"][i].detach()).abs().max().item() <= 1e-6
            assert (res["diffs"][i] - res_o["diffs"][i].detach()).abs().max().item() <= 1e-6
        else:
            assert res["warps"][i] is None and res["diffs"][i] is None and res["masks"][i]["warp_mask"] is None
    assert res_o["masks"][0]["full_mask"].float().mean().item() > 0.005
    assert (res["loss"].detach() - res_o["loss"].detach()).abs().max().item() <= 1e-4
    go = c1.grad.numpy()
    assert np.abs(go).max() > 0
    assert helpers.rel_err(v1.grad.cpu().numpy(), go) < 1e-3


@pytest.mark.parametrize("detach", [False, True])
def test_raster_row_window_is_exact(det_mode, detach):
    """the frame-pair path rasterises only the rows that can matter for the cropped frame (the crop, the
    reach of the occlusion check, and with the geometry gradient the rows of the meshes).  Same bits as the reference's
    full square followed by the crop -- flows, masks, loss and (reproducible mode) gradients."""
    from handobjectconsist_b200 import _config
    S, B, crop = 160, 3, (160, 88)
    dev = torch.device("cuda:0")
    outs = []
    for seed, K_square in ((21, True), (22, False)):  # meshes centred in the square (half below the crop) / in the frame
        sc = synth.make_scene(B, crop[0], crop[1], seed=seed)
        if K_square:
            sc["K"] = synth.camera_intrinsics(B, S, S)
        res = []
        for window in (False, True):
            _config.raster_window = window
            try:
                loss, r, v1 = helpers.pair_step(sc, S, crop, dev, detach, True, False)
                loss.backward()
            finally:
                _config.raster_window = True
            res.append((loss.detach(), r, v1.grad.clone()))
        (l0, r0, g0), (l1, r1, g1) = res
        assert l0.item() > 0 and torch.equal(l0, l1)
        for i in range(2):
            assert torch.equal(r0["flows"][i], r1["flows"][i])
            assert torch.equal(r0["masks"][i]["full_mask"], r1["masks"][i]["full_mask"])
        assert g0.abs().max().item() > 0 and torch.equal(g0, g1)


def test_pair_path_equals_operator_path(det_mode):
    """consist.py against the operator-by-operator surface (get_opticalflow + pair_consist): same device functions,
    so flows and masks are bit-identical; losses to rounding of the partial sums; gradients (reproducible mode: the
    same terms in another order, through different kernels on the vertex side) to 1e-6."""
    from handobjectconsist_b200 import warpbranch
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
    S, B = 96, 3
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, 64, seed=13)
    sc["K"] = synth.camera_intrinsics(B, S, S)
    g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}
    for detach, use_bwd in ((False, True), (True, False)):
        loss, res, v1 = helpers.pair_step(sc, S, (S, 64), dev, detach, use_bwd, True)
        loss.backward()
        w1 = g["verts1"].clone().requires_grad_(True)
        loss_m, res_m = warpbranch.consist_step(w1, g["verts2"], g["faces"], g["K"], g["image_ref"], g["image"],
                                                g["jitter_mask_ref"], g["jitter_mask"], _renderer(S, dev),
                                                PyramidCriterion("l1"), (S, 64), sc["hand_ignore_faces"],
                                                detach_renders=detach, use_backward=use_bwd)
        loss_m.backward()
        for i in range(2):
            assert torch.equal(res["flows"][i], res_m["flows"][i].detach())
            for key in ("full_mask", "flow_mask", "warp_mask"):
                assert torch.equal(res["masks"][i][key], res_m["masks"][i][key])
            assert torch.equal(res["warps"][i], res_m["warps"][i].detach())
            assert torch.equal(res["diffs"][i], res_m["diffs"][i].detach())
        assert (res["loss"].detach() - res_m["loss"].detach()).abs().max().item() <= 1e-6
        assert helpers.rel_err(v1.grad.cpu().numpy(), w1.grad.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("B,S", [(1, 63), (3, 33)])
def test_odd_batch_and_image_size(B, S):
    """B * S * S odd (last-batch remainder at an odd image size): the z-buffer key fill of the fused gather used to
    reject the 8-byte tail."""
    from handobjectconsist_b200 import warpbranch
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, S, seed=9)
    g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}
    v1 = g["verts1"].clone().requires_grad_(True)
    loss, res = warpbranch.consist_step(v1, g["verts2"], g["faces"], g["K"], g["image_ref"], g["image"],
                                        g["jitter_mask_ref"], g["jitter_mask"], _renderer(S, dev),
                                        PyramidCriterion("l1"), (S, S), sc["hand_ignore_faces"], detach_renders=False,
                                        use_backward=True)
    loss.backward()
    c1 = sc["verts1"].clone().requires_grad_(True)
    loss_o, res_o = opipe.consist_step(c1, sc["verts2"], sc["faces"], sc["K"], sc["image_ref"], sc["image"],
                                       sc["jitter_mask_ref"], sc["jitter_mask"], S, (S, S), sc["hand_ignore_faces"],
                                       detach_renders=False, use_backward=True, warp_device=dev)
    loss_o.backward()
    for i in range(2):
        assert (res["flows"][i].detach() - res_o["flows"][i].detach()).abs().max().item() <= 1e-4
    assert abs(loss.item() - loss_o.item()) <= 1e-4
    assert np.abs(v1.grad.cpu().numpy() - c1.grad.numpy()).max() <= 1e-3 * np.abs(c1.grad.numpy()).max()


def test_warpbranch_forward_batch_dicts():
    """The reference's calling convention: sample dicts keyed by TransQueries/BaseQueries + result dicts."""
    from handobjectconsist_b200 import warpbranch
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
    from handobjectconsist_b200.queries import BaseQueries, TransQueries
    S, B = 64, 2
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, S, seed=5)
    hv = 778
    obj_faces = sc["faces"][:, 1552:] - hv
    samples, results = [], []
    for verts, img, jit in ((sc["verts1"], sc["image_ref"], sc["jitter_mask_ref"]),
                            (sc["verts2"], sc["image"], sc["jitter_mask"])):
        samples.append({TransQueries.IMAGE: img, TransQueries.JITTERMASK: jit, TransQueries.CAMINTR: sc["K"],
                        BaseQueries.OBJFACES: obj_faces, BaseQueries.OBJVERTS3D: verts[:, hv:],
                        BaseQueries.HANDVERTS3D: verts[:, :hv]})
        results.append({"recov_handverts3d": verts[:, :hv].to(dev), "recov_objverts3d": verts[:, hv:].to(dev)})
    h = results[0]["recov_handverts3d"].requires_grad_(True)
    o = results[0]["recov_objverts3d"].requires_grad_(True)
    # second frame prediction is garbage on purpose: gt_refs must replace it by the sample's GT vertices
    results[1] = {k: torch.zeros_like(v) for k, v in results[1].items()}
    loss, pair = warpbranch.forward(samples, results, sc["faces"][0, :1552].to(dev), _renderer(S, dev), (S, S),
                                    PyramidCriterion("l1"), gt_refs=True, first_only=True,
                                    hand_ignore_faces=sc["hand_ignore_faces"], use_backward=True)
    loss.backward()
    c1 = sc["verts1"].clone().requires_grad_(True)
    loss_o, _ = opipe.consist_step(c1, sc["verts2"], sc["faces"], sc["K"], sc["image_ref"], sc["image"],
                                   sc["jitter_mask_ref"], sc["jitter_mask"], S, (S, S), sc["hand_ignore_faces"],
                                   detach_renders=True, use_backward=True, warp_device=dev)
    loss_o.backward()
    assert abs(loss.item() - loss_o.item()) <= 1e-4
    g = torch.cat([h.grad, o.grad], 1).cpu().numpy()
    assert helpers.rel_err(g, c1.grad.numpy()) < 1e-3
    assert set(pair.keys()) == {"masks", "warps", "recons_flows", "diffs", "diff_losses"}


@pytest.mark.parametrize("detach", [False, True])
def test_fused_flow_path_equals_op_by_op_path(detach):
    """get_opticalflow's fused kernels (mesh gather/scatter + flow finalize) against the line-by-line
    mirror of the reference built on Renderer / get_occlusion_mask."""
    from handobjectconsist_b200.warping.opticalflow import get_opticalflow
    S, B = 96, 2
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, S, seed=11)
    g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}
    outs = []
    for unfused in (False, True):
        r = _renderer(S, dev)
        r.force_unfused = unfused
        v1 = g["verts1"].clone().requires_grad_(True)
        v2 = g["verts2"].clone().requires_grad_(True)
        flows = get_opticalflow([v1, v2], g["faces"], [g["K"], g["K"]], r, (S, 64), detach_renders=detach,
                                ignore_face_idxs=sc["hand_ignore_faces"])
        w = [torch.randn(f.shape, generator=torch.Generator().manual_seed(i)).to(dev) for i, f in enumerate(flows)]
        (flows[0] * w[0]).sum().add((flows[1] * w[1]).sum()).backward()
        outs.append((flows, v1.grad, v2.grad))
    for i in range(2):
        assert outs[0][0][i].shape == (B, 64, S, 2)
        assert (outs[0][0][i] - outs[1][0][i]).abs().max().item() <= 1e-4  # north_star pixel tolerance
        assert torch.equal(outs[0][0][i] == 0, outs[1][0][i] == 0)
    for k in (1, 2):  # vertex gradients: sums over many faces with cancellation -> relative to the gradient scale
        a, b = outs[0][k].cpu().numpy(), outs[1][k].cpu().numpy()
        assert np.abs(a - b).max() <= 1e-3 * np.abs(b).max()


def test_graphed_step_matches_eager(det_mode):
    """CUDA-graph capture of forward+backward (handobjectconsist_b200.graphed) reproduces the eager step,
    also after loading a different batch into its static buffers, from pinned host memory.  Reproducible mode:
    the graph and the eager path launch the same kernels, so their gradients must agree to the last bit."""
    from handobjectconsist_b200 import warpbranch
    from handobjectconsist_b200.graphed import GraphedConsistStep
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
    from handobjectconsist_b200.queries import BaseQueries, TransQueries
    S, B, hv = 64, 2, 778
    dev = torch.device("cuda:0")

    def batch(seed, device):
        sc = synth.make_scene(B, S, S, seed=seed)
        mv = (lambda t: t.to(device)) if device is not None else (lambda t: t.contiguous().pin_memory())
        obj_faces = mv(sc["faces"][:, 1552:] - hv)
        samples, results = [], []
        for verts, img, jit in ((sc["verts1"], sc["image_ref"], sc["jitter_mask_ref"]),
                                (sc["verts2"], sc["image"], sc["jitter_mask"])):
            samples.append({TransQueries.IMAGE: mv(img), TransQueries.JITTERMASK: mv(jit),
                            TransQueries.CAMINTR: mv(sc["K"]), BaseQueries.OBJFACES: obj_faces,
                            BaseQueries.OBJVERTS3D: mv(verts[:, hv:]), BaseQueries.HANDVERTS3D: mv(verts[:, :hv])})
            results.append({"recov_handverts3d": mv(verts[:, :hv]), "recov_objverts3d": mv(verts[:, hv:])})
        return sc, samples, results

    sc, samples, results = batch(20, dev)
    hand_face = sc["faces"][0, :1552].to(dev)
    crit = PyramidCriterion("l1")
    gstep = GraphedConsistStep(_renderer(S, dev), crit, (S, S), hand_face, samples, results,
                               hand_ignore_faces=sc["hand_ignore_faces"], detach_renders=False)
    for seed, device in ((20, dev), (21, None), (22, dev)):
        sc, samples, results = batch(seed, device)
        loss_g, gh_g, go_g = [t.clone() for t in gstep(samples, results)]
        dres = [{k: v.to(dev) for k, v in r.items()} for r in results]
        h = dres[0]["recov_handverts3d"].requires_grad_(True)
        o = dres[0]["recov_objverts3d"].requires_grad_(True)
        loss_e, _ = warpbranch.forward(samples, dres, hand_face, _renderer(S, dev), (S, S), crit,
                                       hand_ignore_faces=sc["hand_ignore_faces"], detach_renders=False)
        loss_e.backward()
        assert loss_g.item() == loss_e.item()
        assert torch.equal(gh_g, h.grad) and torch.equal(go_g, o.grad)
    # autograd entry: loss as a differentiable function of the predicted vertices
    h2 = dres[0]["recov_handverts3d"].detach().clone().requires_grad_(True)
    dres[0] = {"recov_handverts3d": h2, "recov_objverts3d": dres[0]["recov_objverts3d"].detach()}
    (gstep.apply(samples, dres) * 3.0).backward()
    assert helpers.rel_err(h2.grad.cpu().numpy(), 3.0 * h.grad.cpu().numpy()) < 1e-6


def test_vertex_texture_mode_is_bit_identical_to_cubes():
    """hoc_mesh_gather_clear(VERTEX) + hoc_raster_forward(HOC_LAYOUT_TEX_VERTEX) evaluates the cube texels on the fly:
    rgb must equal, bit for bit, the render of the materialised [B,F',2,2,2,3] cubes (and the gather's key fill must
    leave the forward's result unchanged)."""
    import ctypes

    from handobjectconsist_b200 import _lib
    from oracle import nrfuncs

    L = _lib.lib()
    S, B = 64, 2
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, S, seed=11)
    R, t, dist = torch.eye(3)[None], torch.zeros(1, 1, 3), torch.zeros(1, 5)
    ndc = nrfuncs.projection(sc["verts1"], sc["K"], R, t, dist, float(S)).to(dev).contiguous()
    attrs = torch.randn(B, ndc.shape[1], 3, generator=torch.Generator().manual_seed(0)).to(dev)
    fi = sc["faces"].to(dev).long().contiguous()
    V, Fn = ndc.shape[1], fi.shape[1]
    Fo = 2 * Fn
    st = _lib.stream_ptr()
    bg = (ctypes.c_float * 3)(0.0, 0.0, 0.0)
    outs = []
    for mode in (_lib.HOC_TEX_GRAD_CUBE, _lib.HOC_TEX_GRAD_VERTEX):
        faces = torch.empty((B, Fo, 3, 3), device=dev)
        tex = torch.empty((B, Fo, 2, 2, 2, 3) if mode == _lib.HOC_TEX_GRAD_CUBE else (B, Fo, 3, 3), device=dev)
        n = L.hoc_raster_forward_workspace_bytes(B, Fo, S)
        ws = torch.empty(n, dtype=torch.uint8, device=dev)
        layout = _lib.HOC_LAYOUT_IMAGE
        if mode == _lib.HOC_TEX_GRAD_VERTEX:
            _lib.check(L.hoc_mesh_gather_clear(_lib.ptr(ndc), _lib.ptr(attrs), _lib.ptr(fi), B, V, Fn, 1, mode,
                                               _lib.ptr(faces), _lib.ptr(tex), _lib.ptr(ws), n, st), "gather_clear")
            layout |= _lib.HOC_LAYOUT_KEYS_CLEARED | _lib.HOC_LAYOUT_TEX_VERTEX
        else:
            _lib.check(L.hoc_mesh_gather(_lib.ptr(ndc), _lib.ptr(attrs), _lib.ptr(fi), B, V, Fn, 1, _lib.ptr(faces),
                                         _lib.ptr(tex), st), "gather")
        rgb = torch.empty((B, 3, S, S), device=dev)
        alpha, depth = torch.empty((B, S, S), device=dev), torch.empty((B, S, S), device=dev)
        idx = torch.empty((B, S, S), dtype=torch.int32, device=dev)
        _lib.check(L.hoc_raster_forward(_lib.ptr(faces), _lib.ptr(tex), B, Fo, S, 2, 0.1, 100.0, 1e-3, bg, None, layout,
                                        _lib.ptr(rgb), _lib.ptr(alpha), _lib.ptr(depth), _lib.ptr(idx), None, None,
                                        _lib.ptr(ws), n, st), "forward")
        outs.append((rgb, alpha, depth, idx))
    torch.cuda.synchronize()
    assert (outs[0][3] >= 0).float().mean().item() > 0.02  # the scene covers something
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


def test_graphed_step_accepts_uint8_frames(det_mode):
    """GraphedConsistStep.load with uint8 IMAGE / JITTERMASK (pinned host memory) gives what the equivalent fp32
    tensors (x / 255 - 0.5, x / 255, computed on the CPU like the reference's dataset workers do) give."""
    from handobjectconsist_b200.graphed import GraphedConsistStep
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
    from handobjectconsist_b200.queries import BaseQueries, TransQueries
    S, B, hv = 64, 2, 778
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, S, seed=31)
    obj_faces = (sc["faces"][:, 1552:] - hv).contiguous()

    def batch(as_u8):
        samples, results = [], []
        for verts, img, jit in ((sc["verts1"], sc["image_ref"], sc["jitter_mask_ref"]),
                                (sc["verts2"], sc["image"], sc["jitter_mask"])):
            iu = ((img + 0.5) * 255.0).round().clamp(0, 255).to(torch.uint8)
            ju = (jit * 255.0).round().clamp(0, 255).to(torch.uint8)
            if as_u8:
                im, jm = iu.pin_memory(), ju.pin_memory()
            else:
                im, jm = iu.float().div(255.0).sub(0.5), ju.float().div(255.0)
            samples.append({TransQueries.IMAGE: im, TransQueries.JITTERMASK: jm, TransQueries.CAMINTR: sc["K"],
                            BaseQueries.OBJFACES: obj_faces, BaseQueries.OBJVERTS3D: verts[:, hv:].contiguous(),
                            BaseQueries.HANDVERTS3D: verts[:, :hv].contiguous()})
            results.append({"recov_handverts3d": verts[:, :hv].contiguous(), "recov_objverts3d": verts[:, hv:].contiguous()})
        return samples, results

    fs, fr = batch(False)
    gstep = GraphedConsistStep(_renderer(S, dev), PyramidCriterion("l1"), (S, S), sc["faces"][0, :1552].to(dev), fs, fr,
                               hand_ignore_faces=sc["hand_ignore_faces"], detach_renders=False)
    ref = [t.clone() for t in gstep(fs, fr)]
    us, ur = batch(True)
    got = [t.clone() for t in gstep(us, ur)]
    assert ref[0].item() > 0
    assert abs(ref[0].item() - got[0].item()) <= 1e-6
    for a, b in zip(ref[1:], got[1:]):  # reproducible mode: the widened frames are bit-identical, so are the sums
        assert (a - b).abs().max().item() <= 1e-6 * a.abs().max().item()
    # and a step captured from a uint8 example batch
    gstep2 = GraphedConsistStep(_renderer(S, dev), PyramidCriterion("l1"), (S, S), sc["faces"][0, :1552].to(dev), us, ur,
                                hand_ignore_faces=sc["hand_ignore_faces"], detach_renders=False)
    got2 = gstep2(us, ur)
    assert torch.allclose(ref[0], got2[0], rtol=0, atol=1e-6)
