"""Flow-guided warp, pair consistency and occlusion check -- same surface as
/root/reference/meshreg/warping/imgflowarp.py (``get_spatial_meshgrid`` :8-28, ``warp`` :31-55,
``pair_consist`` :58-115, ``get_occlusion_mask`` :118-146, ``occlusion_mask_from_warped_grid``
:149-172), backed by the kernels of csrc/warp_photo.cu through the C ABI.

The reference's coordinate quirk is kept on purpose: coordinates are normalised with the
align_corners=True formula but sampled with align_corners=False, so a zero flow is not the identity.
"""
import torch
from torch.autograd import Function

from .. import _config, _lib


_side_stream = _config.side_stream


def get_spatial_meshgrid(x: torch.Tensor, scale=False):
    """Grid of pixel coordinates [B,2,H,W] (imgflowarp.py:8-28), built on the device of ``x``."""
    batch_size, _, height, width = x.size()
    xx = torch.arange(0, width, device=x.device).view(1, -1).repeat(height, 1)
    yy = torch.arange(0, height, device=x.device).view(-1, 1).repeat(1, width)
    xx = xx.view(1, 1, height, width).repeat(batch_size, 1, 1, 1)
    yy = yy.view(1, 1, height, width).repeat(batch_size, 1, 1, 1)
    grid = torch.cat((xx, yy), 1).float()
    if scale:
        grid[:, 0] = grid[:, 0] / width
        grid[:, 1] = grid[:, 1] / height
    return grid


class _WarpFunction(Function):
    """out, mask = warp(x, flow); differentiable w.r.t. ``flow`` (bilinear mode)."""

    @staticmethod
    def forward(ctx, x, flow, thresh, mode):
        _lib.require_cuda(x, flow, what="warp")
        L = _lib.lib()
        xc = x.detach().contiguous().float()
        fc = flow.detach().contiguous().float()
        B, C, H, W = xc.shape
        if fc.shape != (B, 2, H, W):
            raise ValueError(f"flow must be [{B}, 2, {H}, {W}], got {tuple(fc.shape)}")
        mode_id = {"bilinear": 0, "nearest": 1}[mode]
        with torch.cuda.device(xc.device):
            out = torch.empty_like(xc)
            mask = torch.empty_like(xc)
            _lib.check(L.hoc_warp(_lib.ptr(xc), _lib.ptr(fc), B, C, H, W, float(thresh), mode_id, _lib.ptr(out),
                                  _lib.ptr(mask), _lib.stream_ptr()), "hoc_warp")
        ctx.save_for_backward(xc, fc)
        ctx.thresh, ctx.mode_id = float(thresh), mode_id
        ctx.mark_non_differentiable(mask)
        ctx.set_materialize_grads(False)
        return out, mask

    @staticmethod
    def backward(ctx, grad_out, grad_mask):
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("warp: gradient w.r.t. the sampled image is not provided "
                                      "(the reference never differentiates through its images)")
        if not ctx.needs_input_grad[1] or grad_out is None:
            return None, None, None, None
        xc, fc = ctx.saved_tensors
        if ctx.mode_id != 0:
            return None, torch.zeros_like(fc), None, None
        L = _lib.lib()
        B, C, H, W = xc.shape
        go = grad_out.contiguous().float()
        with torch.cuda.device(xc.device):
            grad_flow = torch.empty_like(fc)
            _lib.check(L.hoc_warp_backward(_lib.ptr(xc), _lib.ptr(fc), _lib.ptr(go), B, C, H, W, ctx.thresh,
                                           _lib.ptr(grad_flow), _lib.stream_ptr()), "hoc_warp_backward")
        return None, grad_flow, None, None


def warp(x, flow, thresh=0.99999, mode="bilinear"):
    """
    warp an image/tensor (im2) back to im1, according to the optical flow (imgflowarp.py:31-55)

    x: [batch_size, channels, height, width] (im2)
    flow: [batch_size, 2, height, width] flow
    returns (output * mask, mask)
    """
    return _WarpFunction.apply(x, flow, thresh, mode)


class _WarpPhotoFunction(Function):
    """One direction of pair_consist with the L1 criterion, fused (hoc_warp_photo_forward/backward)."""

    @staticmethod
    def forward(ctx, flow, src, target, jitter, thresh):
        _lib.require_cuda(flow, src, target, jitter, what="pair_consist")
        L = _lib.lib()
        fc = flow.detach().contiguous().float()
        sc = src.detach().contiguous().float()
        tc = target.detach().contiguous().float()
        jc = jitter.detach().contiguous().float()
        B, C, H, W = sc.shape
        Cj = jc.shape[1]
        if fc.shape != (B, H, W, 2):
            raise ValueError(f"flow must be [{B}, {H}, {W}, 2], got {tuple(fc.shape)}")
        if tc.shape != sc.shape or jc.shape[0] != B or jc.shape[2:] != (H, W):
            raise ValueError("pair_consist: image / jitter mask shapes do not match")
        with torch.cuda.device(sc.device):
            warped = torch.empty_like(sc)
            warp_mask = torch.empty_like(sc)
            diff = torch.empty_like(sc)
            valid = torch.empty((B, H, W), dtype=torch.bool, device=sc.device)      # written as 0/1 bytes
            flow_mask = torch.empty((B, H, W, 2), dtype=torch.bool, device=sc.device)
            sums = torch.empty((B, 2), dtype=torch.float64, device=sc.device)
            loss = torch.empty((B,), dtype=torch.float32, device=sc.device)
            _lib.check(L.hoc_warp_photo_forward(_lib.ptr(sc), _lib.ptr(tc), _lib.ptr(fc), _lib.ptr(jc), B, C, Cj, H, W,
                                                float(thresh), _lib.ptr(warped), _lib.ptr(warp_mask), _lib.ptr(valid),
                                                _lib.ptr(flow_mask), _lib.ptr(diff), _lib.ptr(sums), _lib.ptr(loss),
                                                _lib.stream_ptr()), "hoc_warp_photo_forward")
        ctx.save_for_backward(sc, tc, fc, valid, sums)
        ctx.thresh = float(thresh)
        ctx.mark_non_differentiable(warp_mask, valid, flow_mask)
        ctx.set_materialize_grads(False)
        return loss, warped, warp_mask, valid, flow_mask, diff

    @staticmethod
    def backward(ctx, grad_loss, grad_warped, grad_mask, grad_valid, grad_flow_mask, grad_diff):
        if grad_warped is not None or grad_diff is not None:
            raise NotImplementedError("pair_consist: only the loss output is differentiable in the fused path")
        if not ctx.needs_input_grad[0] or grad_loss is None:
            return None, None, None, None, None
        sc, tc, fc, valid, sums = ctx.saved_tensors
        L = _lib.lib()
        B, C, H, W = sc.shape
        gl = grad_loss.contiguous().float()
        with torch.cuda.device(sc.device):
            grad_flow = torch.empty_like(fc)
            _lib.check(L.hoc_warp_photo_backward(_lib.ptr(sc), _lib.ptr(tc), _lib.ptr(fc), _lib.ptr(valid),
                                                 _lib.ptr(sums), _lib.ptr(gl), B, C, H, W, ctx.thresh,
                                                 _lib.ptr(grad_flow), _lib.stream_ptr()), "hoc_warp_photo_backward")
        return grad_flow, None, None, None, None


class _PairWarpPhotoFunction(Function):
    """Both directions of pair_consist with the L1 criterion behind ONE autograd node: two
    hoc_warp_photo_forward launches (the second on the side stream) and hoc_pair_loss; the backward hands
    d loss straight to the two hoc_warp_photo_backward launches (no per-direction loss tensors, no add node)."""

    @staticmethod
    def forward(ctx, flow12, flow21, image_ref, image, jitter_ref, jitter, thresh, use_backward):
        _lib.require_cuda(flow12, flow21, image_ref, image, jitter_ref, jitter, what="pair_consist")
        L = _lib.lib()
        c = lambda t: t.detach().contiguous().float()
        f12, f21, ir, im, jr, jm = c(flow12), c(flow21), c(image_ref), c(image), c(jitter_ref), c(jitter)
        B, C, H, W = ir.shape
        if f12.shape != (B, H, W, 2) or f21.shape != (B, H, W, 2):
            raise ValueError(f"flows must be [{B}, {H}, {W}, 2], got {tuple(f12.shape)} / {tuple(f21.shape)}")
        if im.shape != ir.shape or jr.shape[0] != B or jm.shape[0] != B or jr.shape[2:] != (H, W) or jm.shape[2:] != (H, W):
            raise ValueError("pair_consist: image / jitter mask shapes do not match")
        dev = ir.device
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev) if _config.overlap_streams else main
        main_capturing = torch.cuda.is_current_stream_capturing()
        outs = []
        with torch.cuda.device(dev):
            # the sums are zero-filled on the side stream BEFORE it joins the main one: the fill does not depend on
            # the flows, so it leaves the critical path (flow_finalize -> warp kernel) of the captured step
            with torch.cuda.stream(side):
                if torch.cuda.is_current_stream_capturing() != main_capturing:
                    side.wait_stream(main)  # a capture the side stream has not joined yet: fork it first
                sums = torch.zeros((2, B, 2), dtype=torch.float64, device=dev)
                sums.record_stream(main)
            main.wait_stream(side)
            loss = torch.empty((B,), dtype=torch.float32, device=dev)
            side.wait_stream(main)
            # direction 1 (index 0): warp(image_ref, flow21) against image, jitter_mask warped with flow21
            # direction 2 (index 1): warp(image, flow12) against image_ref, jitter_mask_ref warped with flow12
            for k, (src, tgt, fl, jit) in enumerate(((ir, im, f21, jm), (im, ir, f12, jr))):
                with torch.cuda.stream(main if k == 0 else side):
                    warped, warp_mask, diff = torch.empty_like(src), torch.empty_like(src), torch.empty_like(src)
                    valid = torch.empty((B, H, W), dtype=torch.bool, device=dev)
                    flow_mask = torch.empty((B, H, W, 2), dtype=torch.bool, device=dev)
                    _lib.check(L.hoc_warp_photo_forward_acc(_lib.ptr(src), _lib.ptr(tgt), _lib.ptr(fl), _lib.ptr(jit), B,
                                                            C, jit.shape[1], H, W, float(thresh), _lib.ptr(warped),
                                                            _lib.ptr(warp_mask), _lib.ptr(valid), _lib.ptr(flow_mask),
                                                            _lib.ptr(diff), _lib.ptr(sums[k]), None,
                                                            _lib.stream_ptr()), "hoc_warp_photo_forward_acc")
                    for t in (warped, warp_mask, diff, valid, flow_mask):
                        t.record_stream(main)
                    outs.append((warped, warp_mask, valid, flow_mask, diff))
            main.wait_stream(side)
            _lib.check(L.hoc_pair_loss(_lib.ptr(sums[0]), _lib.ptr(sums[1]) if use_backward else None, B,
                                       _lib.ptr(loss), _lib.stream_ptr()), "hoc_pair_loss")
        ctx.save_for_backward(ir, im, f12, f21, outs[0][2], outs[1][2], sums)
        ctx.cfg = (float(thresh), bool(use_backward))
        flat = outs[0] + outs[1]
        ctx.mark_non_differentiable(*[t for o in outs for t in o[1:4]])
        ctx.set_materialize_grads(False)
        return (loss,) + flat

    @staticmethod
    def backward(ctx, grad_loss, *others):
        if any(g is not None for g in others):
            raise NotImplementedError("pair_consist: only the loss output is differentiable in the fused path")
        ir, im, f12, f21, valid1, valid2, sums = ctx.saved_tensors
        thresh, use_backward = ctx.cfg
        if grad_loss is None or not (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            return (None,) * 8
        L = _lib.lib()
        B, C, H, W = ir.shape
        gl = grad_loss.contiguous().float()
        dev = ir.device
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev) if _config.overlap_streams else main
        g12 = g21 = None
        with torch.cuda.device(dev):
            side.wait_stream(main)
            if ctx.needs_input_grad[1]:  # d loss_fwd / d flow21
                g21 = torch.empty_like(f21)
                _lib.check(L.hoc_warp_photo_backward(_lib.ptr(ir), _lib.ptr(im), _lib.ptr(f21), _lib.ptr(valid1),
                                                     _lib.ptr(sums[0]), _lib.ptr(gl), B, C, H, W, thresh, _lib.ptr(g21),
                                                     _lib.stream_ptr()), "hoc_warp_photo_backward")
            if ctx.needs_input_grad[0] and use_backward:  # d loss_bwd / d flow12
                with torch.cuda.stream(side):
                    g12 = torch.empty_like(f12)
                    _lib.check(L.hoc_warp_photo_backward(_lib.ptr(im), _lib.ptr(ir), _lib.ptr(f12), _lib.ptr(valid2),
                                                         _lib.ptr(sums[1]), _lib.ptr(gl), B, C, H, W, thresh,
                                                         _lib.ptr(g12), _lib.stream_ptr()), "hoc_warp_photo_backward")
                    g12.record_stream(main)
                    gl.record_stream(side)
            main.wait_stream(side)
        return g12, g21, None, None, None, None, None, None


def _criterion_is_fused_l1(criterion):
    return getattr(criterion, "name", None) == "l1" and getattr(criterion, "level_nb", 1) == 1


def _one_direction(flow, src, target, jitter, criterion, thresh=0.99999):
    """(loss [B], warp, warp_mask, valid_mask, flow_mask, diff) for one direction."""
    if _criterion_is_fused_l1(criterion):
        loss, warped, warp_mask, valid, flow_mask, diff = _WarpPhotoFunction.apply(flow, src, target, jitter, thresh)
        return loss, warped, warp_mask, valid, flow_mask, diff
    flow_mask = ~(flow == 0)
    # other criteria (l2 / ssim / pyramids): warp kernels + the criterion's own torch ops
    flow_nchw = flow.permute(0, 3, 1, 2)
    warped, warp_mask = warp(src, flow_nchw, thresh)
    warpjitter, _ = warp(jitter, flow_nchw, thresh)
    warp_mask = warp_mask * (warpjitter == 1).float()
    valid = warp_mask[:, 0].bool() & flow_mask[:, :, :, 0] & (jitter[:, 0] == 1)
    _, _, losses, diffs, _ = criterion.compute(warped, target, mask=valid.unsqueeze(1).repeat(1, 3, 1, 1))
    return losses, warped, warp_mask, valid, flow_mask, diffs[0][:, :3]


def pair_consist(recons_flow, image_ref: torch.Tensor, image: torch.Tensor, jitter_mask_ref: torch.Tensor,
                 jitter_mask: torch.Tensor, criterion, use_backward: bool = False):
    """
    Photometric consistency of a frame pair under the mesh-rendered flows (imgflowarp.py:58-115).

    Args:
        recons_flow: [flow12, flow21], each [B,H,W,2]
        image_ref: image of the first frame of the pair, image: image of the second one
        jitter_mask(_ref): 1 where the augmented image holds original pixels
    Returns:
        (warp_loss [B], masks, warps, diffs) exactly like the reference.
    """
    image_ref, image = image_ref.cuda(), image.cuda()
    jitter_mask_ref, jitter_mask = jitter_mask_ref.cuda(), jitter_mask.cuda()
    if _criterion_is_fused_l1(criterion):
        (warp_loss, warp1, warp_mask1, valid_mask1, flow_mask1, diffs_fwd, warp2, warp_mask2, valid_mask2, flow_mask2,
         diffs_bwd) = _PairWarpPhotoFunction.apply(recons_flow[0], recons_flow[1], image_ref, image, jitter_mask_ref,
                                                   jitter_mask, 0.99999, use_backward)
        masks = [
            {"warp_mask": warp_mask1, "full_mask": valid_mask1, "flow_mask": flow_mask1},
            {"warp_mask": warp_mask2, "full_mask": valid_mask2, "flow_mask": flow_mask2},
        ]
        return warp_loss, masks, [warp1, warp2], [diffs_fwd, diffs_bwd]
    # the two directions are independent: the second runs on a side stream (its kernels overlap the first's)
    dev = image.device
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev) if _config.overlap_streams else main
    side.wait_stream(main)
    with torch.cuda.stream(side):
        # direction 2: warp(image, flow12) against image_ref; jitter_mask_ref warped with flow12
        dir2 = _one_direction(recons_flow[0], image, image_ref, jitter_mask_ref, criterion)
        for t in dir2:
            t.record_stream(main)
    # direction 1: warp(image_ref, flow21) against image; jitter_mask warped with flow21 (imgflowarp.py:80-95)
    losses_fwd, warp1, warp_mask1, valid_mask1, flow_mask1, diffs_fwd = _one_direction(
        recons_flow[1], image_ref, image, jitter_mask, criterion)
    main.wait_stream(side)
    losses_bwd, warp2, warp_mask2, valid_mask2, flow_mask2, diffs_bwd = dir2
    masks = [
        {"warp_mask": warp_mask1, "full_mask": valid_mask1, "flow_mask": flow_mask1},
        {"warp_mask": warp_mask2, "full_mask": valid_mask2, "flow_mask": flow_mask2},
    ]
    warps = [warp1, warp2]
    diffs = [diffs_fwd, diffs_bwd]
    if use_backward:
        warp_loss = losses_bwd + losses_fwd
    else:
        warp_loss = losses_fwd
    return warp_loss, masks, warps, diffs


def get_occlusion_mask(mask_flow1, mask_flow2, flow12, flow21):
    """
    Forward-backward consistency check (imgflowarp.py:118-146): a pixel is kept when warping its
    coordinates to the other frame and back (nearest sampling) moves it by less than 0.03 in
    normalised image units and every mask on the way is set.  One fused launch, no gradient.
    masks: [B,1,H,W]; flows: [B,>=2,H,W].  Returns (occl_mask1, occl_mask2), each [B,H,W].
    """
    _lib.require_cuda(mask_flow1, mask_flow2, flow12, flow21, what="get_occlusion_mask")
    L = _lib.lib()
    with torch.no_grad():
        m1 = mask_flow1[:, 0].contiguous().float()
        m2 = mask_flow2[:, 0].contiguous().float()
        f12 = flow12.contiguous().float()
        f21 = flow21.contiguous().float()
        B, Cf, H, W = f12.shape
        with torch.cuda.device(m1.device):
            occl1 = torch.empty_like(m1)
            occl2 = torch.empty_like(m2)
            _lib.check(L.hoc_occlusion_mask(_lib.ptr(m1), _lib.ptr(m2), _lib.ptr(f12), _lib.ptr(f21), B, Cf, H, W,
                                            0.03, _lib.ptr(occl1), _lib.ptr(occl2), _lib.stream_ptr()),
                       "hoc_occlusion_mask")
    return occl1, occl2


def occlusion_mask_from_warped_grid(grid: torch.Tensor, warped_grid: torch.Tensor, distance_thresh=0.03):
    """imgflowarp.py:149-172 on already-warped grids (element-wise torch ops; get_occlusion_mask does
    not go through this -- it is kept for callers that build their own grids)."""
    mask = grid[:, :, :, 2] * warped_grid[:, :, :, 2]
    grid_displs = ((warped_grid - grid) * mask.unsqueeze(-1))[:, :, :, :2].norm(2, -1)
    motion_mask = (grid_displs < distance_thresh).float()
    return mask * motion_mask
