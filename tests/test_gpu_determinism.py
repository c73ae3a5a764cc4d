"""GPU: reproducibility of the gradient sums.

Production accumulates the per-face / per-vertex gradient sums with float atomics (like the reference's
backward_textures / backward_depth_map / index_put(accumulate)), so two runs differ by the rounding of sums whose
terms cancel heavily; the reproducible mode (hoc_set_tuning(HOC_TUNE_DETERMINISTIC, 1), csrc/hoc_det.cuh) adds the
same terms in 128-bit fixed point with integer atomics and must give the same bits every time.  Both are checked at
the bench's size (16 pairs, 9104 faces, 256 x 256)."""
import numpy as np
import pytest
import torch

from handobjectconsist_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _step(g, sc, S, dev):
    from handobjectconsist_b200 import warpbranch
    from handobjectconsist_b200.neurender.renderer import Renderer
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
    r = Renderer(image_size=S, R=torch.eye(3, device=dev)[None], t=torch.zeros(1, 3, device=dev),
                 K=torch.ones(1, 3, 3, device=dev), orig_size=S, anti_aliasing=False, fill_back=True, near=0.1,
                 no_light=True)
    v1 = g["verts1"].clone().requires_grad_(True)
    loss, _ = warpbranch.consist_step(v1, g["verts2"], g["faces"], g["K"], g["image_ref"], g["image"],
                                      g["jitter_mask_ref"], g["jitter_mask"], r, PyramidCriterion("l1"), (S, S),
                                      sc["hand_ignore_faces"], detach_renders=False, use_backward=True)
    loss.backward()
    return loss.detach().clone(), v1.grad.clone()


def test_reproducible_mode_is_bit_exact_at_bench_size():
    S, B = 256, 16
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, S, seed=0)
    g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}
    with _lib.deterministic(True):
        runs = [_step(g, sc, S, dev) for _ in range(4)]
    for loss, grad in runs[1:]:
        assert torch.equal(loss, runs[0][0])
        assert torch.equal(grad, runs[0][1])
    # production mode: same loss (its sums are integer-valued and commute), gradients equal to rounding -- bounded
    # against the gradient scale, and close to the reproducible mode's
    prod = [_step(g, sc, S, dev) for _ in range(3)]
    scale = runs[0][1].abs().max().item()
    assert scale > 0
    worst = 0.0
    for loss, grad in prod:
        assert torch.equal(loss, runs[0][0])
        worst = max(worst, (grad - runs[0][1]).abs().max().item() / scale)
    print(f"production vs reproducible mode, max |dg| / max |g| over 3 runs: {worst:.3e}")
    assert worst <= 1e-3


def test_pair_path_is_reproducible_at_bench_size():
    """The fused frame-pair path (consist.py): its two-phase kernels deal the covered pixels to threads in the order of
    shared-memory atomics, which changes from run to run -- the loss must not (its terms are added as integers), and
    in the reproducible mode neither must the gradients.  Freed NaN-filled blocks poison the allocator's cache between
    runs, so that nothing can lean on stale contents of `torch.empty` buffers (rows outside the raster window)."""
    import helpers
    S, B = 256, 16
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, 144, seed=0)   # 256 x 144 frames on a 256 x 256 raster: the row window is active
    runs = []
    with _lib.deterministic(True):
        for i in range(4):
            junk = torch.full((48, 1024, 1024), float("nan"), device=dev)
            del junk
            loss, res, v1 = helpers.pair_step(sc, S, (S, 144), dev, False, True, False)
            loss.backward()
            runs.append((loss.detach().clone(), res["loss"].detach().clone(), v1.grad.clone()))
    for mean, per_sample, grad in runs[1:]:
        assert torch.equal(mean, runs[0][0]) and torch.equal(per_sample, runs[0][1])
        assert torch.equal(grad, runs[0][2])
    assert runs[0][0].item() > 0 and torch.isfinite(runs[0][2]).all()
    # production mode: the loss is still bit-identical
    loss, res, _ = helpers.pair_step(sc, S, (S, 144), dev, False, True, False)
    assert torch.equal(loss.detach(), runs[0][0])


def test_tuning_switches_do_not_change_the_result():
    """hoc_set_tuning switches that only change HOW the frame-pair step is executed must not change its result, bit
    for bit in the reproducible mode: programmatic dependent launch of its kernels (HOC_TUNE_PDL), and the texture
    gradient through the cover pass instead of the line pass's row CTAs (HOC_TUNE_TEX_IN_LINE = 0: scan pass lists the
    pixels, three gradient planes, hoc_raster_bwd_cover_kernel), and several image lines per CTA of the line pass
    (HOC_TUNE_LINE_LINES / HOC_TUNE_LINE_FOLD)."""
    import helpers
    S, B = 128, 3
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, 96, seed=3)
    L = _lib.lib()

    def run():
        loss, res, v1 = helpers.pair_step(sc, S, (S, 96), dev, False, True, False)
        loss.backward()
        return loss.detach().clone(), res["loss"].detach().clone(), v1.grad.clone()

    with _lib.deterministic(True):
        base = run()
        LINES, FOLD = _lib.HOC_TUNE_LINE_LINES, _lib.HOC_TUNE_LINE_FOLD
        for settings in (((_lib.HOC_TUNE_PDL, 1, 0),), ((_lib.HOC_TUNE_TEX_IN_LINE, 0, 1),),
                         # several image lines per CTA of the line pass (3 does not divide 128: ragged last CTA), folded
                         # (the default) or not; 0 = by raster size (here: 1)
                         ((LINES, 1, 0),), ((LINES, 2, 0), (FOLD, 0, 1)), ((LINES, 3, 0),), ((LINES, 4, 0),),
                         ((LINES, 8, 0), (FOLD, 0, 1))):
            for key, value, _ in settings:
                _lib.check(L.hoc_set_tuning(key, value), "hoc_set_tuning")
            try:
                other = run()
            finally:
                for key, _, back in settings:
                    _lib.check(L.hoc_set_tuning(key, back), "hoc_set_tuning")
            for a, b in zip(base, other):
                assert torch.equal(a, b), settings
        # ... and neither does asking for the loss only (flows / flow masks not handed out: written sparsely)
        loss, res, v1 = helpers.pair_step(sc, S, (S, 96), dev, False, True, False, loss_only=True)
        loss.backward()
        assert res["flows"] == [None, None] and res["masks"][0]["flow_mask"] is None
        assert torch.equal(loss.detach(), base[0]) and torch.equal(res["loss"].detach(), base[1])
        assert torch.equal(v1.grad, base[2])
    assert base[0].item() > 0 and base[2].abs().max().item() > 0
    assert L.hoc_set_tuning(4, 1) != 0 and b"hoc_set_tuning" in L.hoc_last_error()  # (a key of an earlier round: gone)


def test_reproducible_mode_needs_its_workspace():
    """In the reproducible mode the plain hoc_mesh_scatter (no workspace) refuses to run instead of silently falling
    back to float atomics, and the rasterizer backward asks for the larger workspace."""
    L = _lib.lib()
    B, V, F = 2, 50, 30
    dev = torch.device("cuda:0")
    assert L.hoc_mesh_scatter_workspace_bytes(B, V) == 0
    small = L.hoc_raster_backward_workspace_bytes_ex(B, F, 32, 2, _lib.HOC_TEX_GRAD_VERTEX)
    assert small == L.hoc_raster_backward_workspace_bytes(B, F, 32)
    gf = torch.randn(B, 2 * F, 3, 3, device=dev)
    fi = torch.randint(0, V, (B, F, 3), device=dev)
    gv = torch.empty(B, V, 3, device=dev)
    st = _lib.stream_ptr()
    with _lib.deterministic(True):
        need = L.hoc_mesh_scatter_workspace_bytes(B, V)
        assert need == 2 * 16 * 3 * B * V
        assert L.hoc_raster_backward_workspace_bytes_ex(B, F, 32, 2, _lib.HOC_TEX_GRAD_VERTEX) > small
        rc = L.hoc_mesh_scatter(_lib.ptr(gf), None, _lib.ptr(fi), B, V, F, 1, _lib.HOC_TEX_GRAD_VERTEX, _lib.ptr(gv),
                                None, st)
        assert rc == -3 and b"workspace" in L.hoc_last_error()
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        outs = []
        for _ in range(2):
            _lib.check(L.hoc_mesh_scatter_ws(_lib.ptr(gf), None, _lib.ptr(fi), B, V, F, 1, _lib.HOC_TEX_GRAD_VERTEX,
                                             _lib.ptr(gv), None, 0, _lib.ptr(ws), need, st), "scatter")
            outs.append(gv.clone())
    assert torch.equal(outs[0], outs[1])
    # against index_add in float64
    ref = torch.zeros(B, V, 3, dtype=torch.float64, device=dev)
    f2 = torch.cat([fi, fi.flip(-1)], 1)
    for b in range(B):
        ref[b].index_add_(0, f2[b].reshape(-1), gf[b].reshape(-1, 3).double())
    assert (outs[0].double() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


def test_out_of_range_face_indices_are_skipped():
    """A face table that points outside [0, V) must not read or write out of bounds: the gather reads
    vertex 0 for such an index, the scatter drops its contribution."""
    L = _lib.lib()
    B, V, F = 1, 10, 4
    dev = torch.device("cuda:0")
    verts = torch.randn(B, V, 3, device=dev)
    attrs = torch.randn(B, V, 3, device=dev)
    fi = torch.tensor([[[0, 1, 2], [3, 4, 5], [6, 7, 99], [-1, 8, 9]]], device=dev)
    faces = torch.empty(B, 2 * F, 3, 3, device=dev)
    tex = torch.empty(B, 2 * F, 2, 2, 2, 3, device=dev)
    st = _lib.stream_ptr()
    _lib.check(L.hoc_mesh_gather(_lib.ptr(verts), _lib.ptr(attrs), _lib.ptr(fi), B, V, F, 1, _lib.ptr(faces),
                                 _lib.ptr(tex), st), "gather")
    torch.cuda.synchronize()
    assert torch.equal(faces[0, 0], verts[0, :3]) and torch.equal(faces[0, 2, 2], verts[0, 0])
    gf = torch.ones(B, 2 * F, 3, 3, device=dev)
    gv = torch.empty(B, V, 3, device=dev)
    _lib.check(L.hoc_mesh_scatter(_lib.ptr(gf), None, _lib.ptr(fi), B, V, F, 1, _lib.HOC_TEX_GRAD_CUBE, _lib.ptr(gv),
                                  None, st), "scatter")
    torch.cuda.synchronize()
    assert torch.isfinite(gv).all() and gv.sum().item() == pytest.approx(3 * 2 * (3 * F - 2))
