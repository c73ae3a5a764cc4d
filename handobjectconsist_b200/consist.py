"""The photometric-consistency step of ONE frame pair behind ONE autograd node.

``warpbranch.forward`` for a pair is, in the reference, batch_cat_meshes -> get_opticalflow (two renders, masks,
occlusion check) -> pair_consist (four warps, mask algebra, two masked L1 means)
(/root/reference/meshreg/models/warpbranch.py:45-88, meshreg/warping/opticalflow.py:51-156,
meshreg/warping/imgflowarp.py:58-115).  The operator-by-operator mirrors of those functions live in
``warping/opticalflow.py`` and ``warping/imgflowarp.py``; this module is the training fast path: the same arithmetic
(same device functions, bit-identical flows / masks) as FIVE launches forward and FOUR backward, with the two renders
of the pair stacked along the batch ([2B]) so that every rasterizer kernel runs once:

    forward   hoc_pair_front                 vertices of both frames -> faces / vertex textures of both renders, key fill
              hoc_raster_forward  [2B]       z-buffer pass + resolve pass
              hoc_flow_finalize_warp         masks, ignore faces, occlusion check, crop -> flow12, flow21, and -- same
                                             pixel, same pass -- both warp directions, valid masks, |warp - target| sums
                                             (with visuals: hoc_flow_finalize, then hoc_warp_photo_forward_pair)
              hoc_pair_loss_mean             masked means -> loss [B] and its batch mean
    backward  hoc_pair_backward_raster [2B]  scan pass FUSED with the backward of pair_consist (d loss / d flow x d flow /
                                             d rgb computed from the valid masks), then the line pass (pseudo-gradient
                                             of the first render, texture gradient of both)
              hoc_mesh_scatter       [2B]    faces -> vertices
              hoc_pair_back                  projection adjoints -> d loss / d (hand, object vertices)

No memset and no ATen kernel in between: every zero-fill a kernel needs is done by the kernel before it (the loss
sums by hoc_pair_front, the rasterizer backward's counters by hoc_pair_loss_mean, the scatter's outputs by the
rasterizer backward's streaming pass), and the batch mean and its adjoint live inside hoc_pair_loss_mean / the
rasterizer backward's streaming pass.

``return_visuals=False`` (what ``GraphedConsistStep`` asks for) skips the three visualisation returns of pair_consist
(``warps``, ``diffs``, ``warp_mask``): nothing the loss or its gradient depends on.
"""
import ctypes

import torch
from torch.autograd import Function

from . import _config, _lib
from .warping.opticalflow import _fused_path_ok, _ignore_tensor
from .warping.imgflowarp import _criterion_is_fused_l1

_CAM_DEFAULTS = {}

# benchmarking hook: a dict here receives the measured coverage of the last forward (a device -> host read: never set
# inside a captured step)
STATS = None


def _default_cams(device):
    """R = I, t = 0, no distortion (what WarpRegNet's renderer uses), created once per device -- not per step."""
    key = str(device)
    if key not in _CAM_DEFAULTS:
        _CAM_DEFAULTS[key] = (torch.eye(3, dtype=torch.float32, device=device)[None].contiguous(),
                              torch.zeros(1, 3, dtype=torch.float32, device=device),
                              torch.zeros(1, 5, dtype=torch.float32, device=device))
    return _CAM_DEFAULTS[key]


def pair_path_ok(renderer, criterion, images, jitter_masks, image_size):
    """The fused pair path covers the configuration WarpRegNet builds (warpreg.py:40-51): projection camera, no
    anti-aliasing, no lighting, L1 criterion at one scale, RGB frames with 3-channel jitter masks; widths that are
    multiples of four (16-byte accesses).  Anything else takes the operator-by-operator path."""
    if not (_fused_path_ok(renderer) and _criterion_is_fused_l1(criterion)):
        return False
    S = int(renderer.image_size)
    W, H = min(int(image_size[0]), S), min(int(image_size[1]), S)
    if S % 4 or W % 4 or W < 4:
        return False
    for t in list(images) + list(jitter_masks):
        if t.dim() != 4 or t.shape[1] != 3 or t.shape[2] != H or t.shape[3] != W or t.dtype != torch.float32:
            return False
    return True


class _PairConsistFunction(Function):
    @staticmethod
    def forward(ctx, hand1, obj1, hand2, obj2, hand_faces, obj_faces, K1, K2, image_ref, image, jitter_ref, jitter, cfg):
        _lib.require_cuda(hand1, obj1, hand2, obj2, hand_faces, obj_faces, K1, K2, image_ref, image, jitter_ref, jitter,
                          what="consist pair")
        L = _lib.lib()
        r = cfg["renderer"]
        c = lambda t: t.detach().contiguous().float()
        h1, o1, h2, o2 = c(hand1), c(obj1), c(hand2), c(obj2)
        ir, im, jr, jm = c(image_ref), c(image), c(jitter_ref), c(jitter)
        hf = hand_faces.detach().contiguous().long()
        of = obj_faces.detach().contiguous().long()
        if hf.dim() == 3 and hf.shape[0] == 1:
            hf = hf[0]
        B, Vh, Vo = h1.shape[0], h1.shape[1], o1.shape[1]
        Fh, Fo = hf.shape[-2], of.shape[1]
        V, Fn = Vh + Vo, Fh + Fo
        fill_back = bool(r.fill_back)
        Fr = 2 * Fn if fill_back else Fn
        S = int(r.image_size)
        W, H = cfg["wh"]
        dev = h1.device
        dt = torch.float32
        dR, dtr, ddist = _default_cams(dev)
        R = c(r.R) if r.R is not None else dR
        t = c(r.t) if r.t is not None else dtr
        dist = c(r.dist_coeffs) if r.dist_coeffs is not None else ddist
        cams = (c(K1).reshape(-1, 9), c(K2).reshape(-1, 9), R.reshape(-1, 9), t.reshape(-1, 3), dist.reshape(-1, 5))
        for cam in cams:
            if cam.shape[0] not in (1, B):
                raise ValueError("camera tensors must have batch dimension 1 or B")
        cam_args = []
        for cam in cams:
            cam_args += [_lib.ptr(cam), int(cam.shape[0] == B and B > 1)]
        ignore = cfg["ignore"]
        n_ign = 0 if ignore is None else ignore.numel()
        visuals = bool(cfg["visuals"])
        loss_only = bool(cfg.get("loss_only")) and not visuals
        bg = (ctypes.c_float * 3)(*[float(x) for x in r.background_color])
        near, far, eps = float(r.near), float(r.far), float(r.rasterizer_eps)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            e = lambda *shape, dtype=dt: torch.empty(shape, dtype=dtype, device=dev)
            faces, tex = e(2 * B, Fr, 3, 3), e(2 * B, Fr, 3, 3)
            table = e(2 * B, Fn, 3, dtype=torch.int64)
            ws_bytes = L.hoc_raster_forward_workspace_bytes(2 * B, Fr, S)
            ws = e(max(ws_bytes, 16), dtype=torch.uint8)
            sums = e(2, B, 2, dtype=torch.float64)  # 32 B bytes: a multiple of 16; zero-filled by the front kernel
            # raster row window: only when the frame is lower than the square raster; the geometry gradient
            # (if any input can ask for it) extends it to the rows of the meshes
            want_geom = (not cfg["detach_renders"]) and any(ctx.needs_input_grad[:4])
            row_lo = e(2 * B, dtype=torch.int32) if (_config.raster_window and H < S) else None
            _lib.check(L.hoc_pair_front(_lib.ptr(h1), _lib.ptr(o1), _lib.ptr(h2), _lib.ptr(o2), _lib.ptr(hf),
                                        int(hf.dim() == 3), _lib.ptr(of), *cam_args, float(r.orig_size), B, Vh, Vo, Fh,
                                        Fo, int(fill_back), _lib.ptr(faces), _lib.ptr(tex), _lib.ptr(table),
                                        _lib.ptr(ws), ws_bytes, _lib.ptr(sums), sums.numel() * 8, _lib.ptr(row_lo), S, H,
                                        int(want_geom), st), "hoc_pair_front")
            rgb, alpha, idx = e(2 * B, 3, S, S), e(2 * B, S, S), e(2 * B, S, S, dtype=torch.int32)
            # depth / weight_map only feed the backward, at covered pixels (HOC_LAYOUT_SPARSE_SAVED)
            depth, wmap = e(2 * B, S, S), e(2 * B, S, S, 3)
            layout = (_lib.HOC_LAYOUT_IMAGE | _lib.HOC_LAYOUT_KEYS_CLEARED | _lib.HOC_LAYOUT_TEX_VERTEX
                      | _lib.HOC_LAYOUT_SPARSE_SAVED)
            _lib.check(L.hoc_raster_forward_ex(_lib.ptr(faces), _lib.ptr(tex), 2 * B, Fr, S, 2, near, far, eps, bg, None,
                                               layout, _lib.ptr(row_lo), _lib.ptr(rgb), _lib.ptr(alpha), _lib.ptr(depth),
                                               _lib.ptr(idx), _lib.ptr(wmap), None, _lib.ptr(ws), ws_bytes, st),
                       "hoc_raster_forward")
            if STATS is not None:  # (rows below the window are undefined: counted through the z-buffer keys instead)
                keys = ws[: 2 * B * S * S * 8].view(torch.int64)
                STATS["coverage"] = float((keys != -1).float().mean())
            flow12, flow21 = e(B, H, W, 2), e(B, H, W, 2)
            mult = e(2, B, H, W)
            valid = e(2, B, H, W, dtype=torch.bool)
            flow_mask = None if loss_only else e(2, B, H, W, 2, dtype=torch.bool)
            pp = _lib.ptr_pair
            renders = (_lib.ptr(rgb[:B]), _lib.ptr(alpha[:B]), _lib.ptr(idx[:B]), _lib.ptr(rgb[B:]), _lib.ptr(alpha[B:]),
                       _lib.ptr(idx[B:]))
            vis = None
            if not visuals:
                # loss_only: flows / mult are written where the valid mask is set (all the backward reads) and left
                # undefined elsewhere, no flow masks: 24 B/px of zero stores less
                _lib.check(L.hoc_flow_finalize_warp_ex(
                    *renders, _lib.ptr(ir), _lib.ptr(im), _lib.ptr(jr), _lib.ptr(jm), B, S, H, W, _lib.ptr(ignore), n_ign,
                    0.03, float(cfg["thresh"]), _lib.ptr(flow12), _lib.ptr(flow21), _lib.ptr(mult[0]), _lib.ptr(mult[1]),
                    pp(valid[0], valid[1]), None if loss_only else pp(flow_mask[0], flow_mask[1]), _lib.ptr(sums),
                    int(loss_only), st), "hoc_flow_finalize_warp")
            else:
                _lib.check(L.hoc_flow_finalize(*renders, B, S, H, W, _lib.ptr(ignore), n_ign, 1, 0.03, _lib.ptr(flow12),
                                               _lib.ptr(flow21), _lib.ptr(mult[0]), _lib.ptr(mult[1]), st),
                           "hoc_flow_finalize")
                vis = e(6, B, 3, H, W)  # warped, warp_mask, diff of direction 0, then of direction 1
                _lib.check(L.hoc_warp_photo_forward_pair(
                    _lib.ptr(ir), _lib.ptr(im), _lib.ptr(flow12), _lib.ptr(flow21), _lib.ptr(jr), _lib.ptr(jm), B, H, W,
                    float(cfg["thresh"]), 1, pp(vis[0], vis[3]), pp(vis[1], vis[4]), pp(vis[2], vis[5]),
                    pp(valid[0], valid[1]), pp(flow_mask[0], flow_mask[1]), _lib.ptr(sums), st),
                    "hoc_warp_photo_forward_pair")
            loss, mean = e(B), e()
            # the workspace of the step's backward is allocated here, so that this launch can zero-fill its counters
            # (one memset node less in a captured step); a second backward over the same graph falls back to a memset
            bws = bws_zero = None
            if any(ctx.needs_input_grad[:4]):
                n_b = 2 * B if cfg["use_backward"] else B
                bws_bytes = L.hoc_raster_backward_workspace_bytes_ex(n_b, Fr, S, 2, _lib.HOC_TEX_GRAD_VERTEX)
                bws = e(max(bws_bytes, 16), dtype=torch.uint8)
                bws_zero = L.hoc_pair_backward_zero_bytes(n_b, Fr, S)  # (no depth gradient: spans and counters only)
            _lib.check(L.hoc_pair_loss_mean(_lib.ptr(sums[0]), _lib.ptr(sums[1]) if cfg["use_backward"] else None, B,
                                            _lib.ptr(loss), _lib.ptr(mean), _lib.ptr(bws), bws_zero or 0, st),
                       "hoc_pair_loss_mean")
            ctx.bws, ctx.bws_clean = bws, bws is not None
        ctx.row_lo = row_lo
        ctx.save_for_backward(h1, o1, h2, o2, faces, table, idx, rgb, wmap, depth, ir, im, flow12, flow21, valid, sums,
                              mult, *cams)
        ctx.cfg = dict(B=B, Vh=Vh, Vo=Vo, Fn=Fn, Fr=Fr, S=S, H=H, W=W, near=near, far=far, eps=eps, fill_back=fill_back,
                       orig_size=float(r.orig_size), detach_renders=bool(cfg["detach_renders"]),
                       use_backward=bool(cfg["use_backward"]), cam_flags=[a for a in cam_args[1::2]])
        if loss_only:  # (flows are undefined outside the valid pixels: not handed out)
            outs = (loss, mean, None, None, valid[0], valid[1], None, None)
        else:
            outs = (loss, mean, flow12, flow21, valid[0], valid[1], flow_mask[0], flow_mask[1])
        if visuals:
            outs = outs + tuple(vis[k] for k in range(6))
        ctx.mark_non_differentiable(*[o for o in outs[2:] if o is not None])
        ctx.set_materialize_grads(False)
        return outs

    @staticmethod
    def backward(ctx, g_loss, g_mean, *unused):
        need = ctx.needs_input_grad
        need1, need2 = need[0] or need[1], need[2] or need[3]
        if (g_loss is None and g_mean is None) or not (need1 or need2):
            return (None,) * 13
        (h1, o1, h2, o2, faces, table, idx, rgb, wmap, depth, ir, im, flow12, flow21, valid, sums, mult, K1, K2, R, t,
         dist) = ctx.saved_tensors
        k = ctx.cfg
        B, Vh, Vo, Fn, Fr, S, H, W = k["B"], k["Vh"], k["Vo"], k["Fn"], k["Fr"], k["S"], k["H"], k["W"]
        V = Vh + Vo
        L = _lib.lib()
        dev = h1.device
        row_lo = ctx.row_lo
        rl = (lambda off: None) if row_lo is None else (lambda off: _lib.ptr(row_lo[off:]))
        gl = None if g_loss is None else g_loss.contiguous().float()
        gm = None if g_mean is None else g_mean.contiguous().float()
        use_backward = k["use_backward"]
        # rows of the stacked batch that receive a gradient: without the backward direction the render of mesh 1
        # (rows 0..B-1) has none (imgflowarp.py:108-114)
        lo = 0 if use_backward else B
        n = 2 * B - lo
        if k["detach_renders"]:
            geom = 0
        elif need2:
            geom = n
        else:
            geom = B - lo  # the pseudo-gradient of mesh 1 only (first_only, warpbranch.py:45-55)
        cam_args = []
        for cam, flag in zip((K1, K2, R, t, dist), k["cam_flags"]):
            cam_args += [_lib.ptr(cam), flag]
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            e = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
            ws_bytes = L.hoc_raster_backward_workspace_bytes_ex(n, Fr, S, 2, _lib.HOC_TEX_GRAD_VERTEX)
            ws = ctx.bws
            clean = ctx.bws_clean and ws is not None and ws.numel() >= ws_bytes
            if not clean:  # (a second backward over the same graph, or the reproducible mode switched on in between)
                ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
            ctx.bws_clean = False
            grad_rgb = e(n, 3, S, S)  # scratch of the fused scan pass: the incoming gradient of the renders' rgb maps
            # grad_faces | grad_textures | (grad of the NDC vertices, grad of the vertex attributes), back to back: ONE
            # region for the zero-fill of the rasterizer backward's streaming pass (the last one is accumulated by the
            # scatter after it)
            n_f = n * Fr * 9
            flat = e((n_f if geom > 0 else 0) + n_f + 2 * 2 * B * V * 3)
            o = n_f if geom > 0 else 0
            grad_faces = flat[:n_f].view(n, Fr, 3, 3) if geom > 0 else None
            grad_tex = flat[o:o + n_f].view(n, Fr, 3, 3)
            both = flat[o + n_f:].view(2, 2 * B, V, 3)
            _lib.check(L.hoc_pair_backward_raster(
                _lib.ptr(ir), _lib.ptr(im), _lib.ptr(flow12), _lib.ptr(flow21), _lib.ptr_pair(valid[0], valid[1]),
                _lib.ptr(sums), _lib.ptr(mult[0]), _lib.ptr(mult[1]), _lib.ptr(gl), _lib.ptr(gm), B, H, W,
                int(use_backward), lo, _lib.ptr(faces[lo:]), _lib.ptr(idx[lo:]), _lib.ptr(rgb[lo:]), _lib.ptr(wmap[lo:]),
                _lib.ptr(depth[lo:]), _lib.ptr(grad_rgb), n, Fr, S, k["near"], k["far"], k["eps"], geom,
                _lib.HOC_BWD_WORKSPACE_ZEROED if clean else 0, _lib.ptr(both), both.numel() * 4, rl(lo),
                _lib.ptr(grad_faces), _lib.ptr(grad_tex), _lib.ptr(ws), ws.numel(), st), "hoc_pair_backward_raster")
            g_ndc, g_attr = (both[0] if geom > 0 else None), both[1]
            sc_bytes = L.hoc_mesh_scatter_workspace_bytes(n, V)
            sc_ws = torch.empty(sc_bytes, dtype=torch.uint8, device=dev) if sc_bytes else None
            gv_arg = _lib.ptr(both[0, lo:]) if geom > 0 else None
            ga_arg = _lib.ptr(both[1, lo:])
            _lib.check(L.hoc_mesh_scatter_ws(_lib.ptr(grad_faces), _lib.ptr(grad_tex), _lib.ptr(table[lo:]), n, V, Fn,
                                             int(k["fill_back"]), _lib.HOC_TEX_GRAD_VERTEX, gv_arg, ga_arg, 1,
                                             _lib.ptr(sc_ws), sc_bytes, st), "hoc_mesh_scatter")
            gv1 = e(B, V, 3) if need1 else None
            gv2 = e(B, V, 3) if need2 else None
            has_ndc1 = int(geom > 0 and lo == 0)
            has_ndc2 = int(geom > 0 and need2)
            _lib.check(L.hoc_pair_back(_lib.ptr(h1), _lib.ptr(o1), _lib.ptr(h2), _lib.ptr(o2), *cam_args,
                                       k["orig_size"], B, Vh, Vo, _lib.ptr(g_ndc), _lib.ptr(g_attr), has_ndc1, has_ndc2,
                                       int(use_backward), 1, _lib.ptr(gv1), _lib.ptr(gv2), st), "hoc_pair_back")
        gh1 = gv1[:, :Vh] if (gv1 is not None and need[0]) else None
        go1 = gv1[:, Vh:] if (gv1 is not None and need[1]) else None
        gh2 = gv2[:, :Vh] if (gv2 is not None and need[2]) else None
        go2 = gv2[:, Vh:] if (gv2 is not None and need[3]) else None
        return (gh1, go1, gh2, go2) + (None,) * 9


def pair_consist_step(hand1, obj1, hand2, obj2, hand_faces, obj_faces, K1, K2, image_ref, image, jitter_mask_ref,
                      jitter_mask, renderer, image_size, hand_ignore_faces=None, detach_renders=True, use_backward=True,
                      return_visuals=True, thresh=0.99999, loss_only=False):
    """Loss [B] of one frame pair and the reference's result structures.

    Returns ``((loss, mean), flows, masks, warps, diffs)``: ``loss`` [B] and its batch mean (scalar, computed by the
    same launch); ``flows = [flow12, flow21]`` ([B,H,W,2]); ``masks`` the two dicts
    of pair_consist (``warp_mask`` is None without visuals); ``warps`` / ``diffs`` lists of two tensors (None entries
    without visuals).  Only ``loss`` / ``mean`` are differentiable -- w.r.t. the four vertex tensors.
    ``loss_only=True`` (with ``return_visuals=False``: what a captured training step asks for): the flows and flow
    masks are not handed out either (None) -- the step then writes them only where the backward reads them."""
    S = int(renderer.image_size)
    wh = (min(int(image_size[0]), S), min(int(image_size[1]), S)) if image_size is not None else (S, S)
    ignore = None if hand_ignore_faces is None else _ignore_tensor(hand_ignore_faces, hand1.device)
    cfg = dict(renderer=renderer, wh=wh, ignore=ignore, detach_renders=detach_renders, use_backward=use_backward,
               visuals=return_visuals, thresh=thresh, loss_only=bool(loss_only) and not return_visuals)
    outs = _PairConsistFunction.apply(hand1, obj1, hand2, obj2, hand_faces, obj_faces, K1, K2, image_ref, image,
                                      jitter_mask_ref, jitter_mask, cfg)
    loss, mean, flow12, flow21, valid1, valid2, fmask1, fmask2 = outs[:8]
    vis = outs[8:] if return_visuals else (None,) * 6
    masks = [{"warp_mask": vis[1], "full_mask": valid1, "flow_mask": fmask1},
             {"warp_mask": vis[4], "full_mask": valid2, "flow_mask": fmask2}]
    return (loss, mean), [flow12, flow21], masks, [vis[0], vis[3]], [vis[2], vis[5]]
