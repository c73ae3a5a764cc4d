"""autograd nodes over the geometry-head kernels of ``libhoc_b200.so`` (``csrc/geom_head.cu``;
``include/hoc_b200.h``: hoc_hand_head_*, hoc_recover_points_*).  One launch per direction; CUDA tensors only."""
import torch
from torch.autograd import Function

from . import _lib


def _f32c(t):
    return None if t is None else t.contiguous().float()


def _cam(camintr, scale, trans, B):
    """Kernel-side views of the shared camera arguments: camintr [B or 1,3,3], scale [B], trans [B,2]."""
    if camintr.dim() != 3 or camintr.shape[1:] != (3, 3) or camintr.shape[0] not in (1, B):
        raise ValueError(f"camintr must be [{B},3,3] or [1,3,3], got {tuple(camintr.shape)}")
    if scale.numel() != B or trans.numel() != 2 * B:
        raise ValueError(f"scale / trans must hold {B} / {2 * B} values, got {tuple(scale.shape)} / {tuple(trans.shape)}")
    return _f32c(camintr), int(camintr.shape[0] == B)


class _RecoverPointsFunction(Function):
    """points [B,N,3], rot [B,3] or None, camintr, scale [B], trans [B,2] ->
    (rot_points or None, recov_points, points2d, center3d [B,3])."""

    @staticmethod
    def forward(ctx, points, rot, camintr, scale, trans, scale_factor, trans_factor, off_z, res_w, res_h):
        _lib.require_cuda(points, rot, camintr, scale, trans, what="recover_3d_proj / ObjBranch")
        L = _lib.lib()
        B, N = points.shape[:2]
        points, rot, scale, trans = _f32c(points), _f32c(rot), _f32c(scale), _f32c(trans)
        camintr, batched = _cam(camintr, scale, trans, B)
        dev = points.device
        with torch.cuda.device(dev):
            rot_points = torch.empty_like(points) if rot is not None else None
            recov = torch.empty_like(points)
            pts2d = torch.empty(B, N, 2, device=dev)
            center = torch.empty(B, 3, device=dev)
            _lib.check(L.hoc_recover_points_forward(
                _lib.ptr(points), _lib.ptr(rot), B, N, _lib.ptr(camintr), batched, _lib.ptr(scale), _lib.ptr(trans),
                scale_factor, trans_factor, off_z, res_w, res_h, _lib.ptr(rot_points), _lib.ptr(recov),
                _lib.ptr(pts2d), _lib.ptr(center), _lib.stream_ptr()), "hoc_recover_points_forward")
        ctx.save_for_backward(points, rot, camintr, scale, trans)
        ctx.consts = (batched, scale_factor, trans_factor, off_z, res_w, res_h)
        ctx.set_materialize_grads(False)
        return rot_points, recov, pts2d, center

    @staticmethod
    def backward(ctx, g_rot_points, g_recov, g_2d, g_center):
        points, rot, camintr, scale, trans = ctx.saved_tensors
        batched, scale_factor, trans_factor, off_z, res_w, res_h = ctx.consts
        if g_rot_points is None and g_recov is None and g_2d is None and g_center is None:
            return (None,) * 10
        L = _lib.lib()
        B, N = points.shape[:2]
        need = ctx.needs_input_grad
        with torch.cuda.device(points.device):
            gp = torch.empty_like(points) if need[0] else None
            gr = torch.empty_like(rot) if (rot is not None and need[1]) else None
            gs = torch.empty_like(scale) if need[3] else None
            gt = torch.empty_like(trans) if need[4] else None
            _lib.check(L.hoc_recover_points_backward(
                _lib.ptr(points), _lib.ptr(rot), B, N, _lib.ptr(camintr), batched, _lib.ptr(scale), _lib.ptr(trans),
                scale_factor, trans_factor, off_z, res_w, res_h, _lib.ptr(_f32c(g_rot_points)),
                _lib.ptr(_f32c(g_recov)), _lib.ptr(_f32c(g_2d)), _lib.ptr(_f32c(g_center)), _lib.ptr(gp), _lib.ptr(gr),
                _lib.ptr(gs), _lib.ptr(gt), _lib.stream_ptr()), "hoc_recover_points_backward")
        return gp, gr, None, gs, gt, None, None, None, None, None


class _HandHeadFunction(Function):
    """verts [B,V,3], joints_in [B,J,3] or None, adaptor [J,V] or None, camintr, scale [B], trans [B,2] ->
    (joints3d, verts3d, recov_joints3d, recov_verts3d, joints2d, verts2d, center3d)."""

    @staticmethod
    def forward(ctx, verts, joints_in, adaptor, camintr, scale, trans, center_idx, scale_factor, trans_factor, off_z,
                res_w, res_h):
        _lib.require_cuda(verts, joints_in, adaptor, camintr, scale, trans, what="recover_mano geometry")
        L = _lib.lib()
        B, V = verts.shape[:2]
        verts, joints_in, adaptor, scale, trans = (_f32c(verts), _f32c(joints_in), _f32c(adaptor), _f32c(scale),
                                                   _f32c(trans))
        if adaptor is not None:
            if adaptor.dim() != 2 or adaptor.shape[1] != V:
                raise ValueError(f"adaptor weight must be [J,{V}], got {tuple(adaptor.shape)}")
            J = adaptor.shape[0]
        else:
            if joints_in is None or joints_in.dim() != 3 or joints_in.shape[0] != B:
                raise ValueError("without an adaptor the joints [B,J,3] must be given")
            J = joints_in.shape[1]
        camintr, batched = _cam(camintr, scale, trans, B)
        dev = verts.device
        with torch.cuda.device(dev):
            joints3d = torch.empty(B, J, 3, device=dev)
            verts3d = torch.empty_like(verts)
            recov_j = torch.empty(B, J, 3, device=dev)
            recov_v = torch.empty_like(verts)
            j2d = torch.empty(B, J, 2, device=dev)
            v2d = torch.empty(B, V, 2, device=dev)
            center = torch.empty(B, 3, device=dev)
            _lib.check(L.hoc_hand_head_forward(
                _lib.ptr(verts), _lib.ptr(joints_in if adaptor is None else None), _lib.ptr(adaptor), B, V, J,
                center_idx, _lib.ptr(camintr), batched, _lib.ptr(scale), _lib.ptr(trans), scale_factor, trans_factor,
                off_z, res_w, res_h, _lib.ptr(joints3d), _lib.ptr(verts3d), _lib.ptr(recov_j), _lib.ptr(recov_v),
                _lib.ptr(j2d), _lib.ptr(v2d), _lib.ptr(center), _lib.stream_ptr()), "hoc_hand_head_forward")
        ctx.save_for_backward(verts if adaptor is not None else None, adaptor, camintr, scale, trans, recov_j, recov_v)
        ctx.consts = (batched, center_idx, scale_factor, trans_factor, off_z, res_w, res_h, J)
        ctx.set_materialize_grads(False)
        return joints3d, verts3d, recov_j, recov_v, j2d, v2d, center

    @staticmethod
    def backward(ctx, g_j3d, g_v3d, g_rj, g_rv, g_j2d, g_v2d, g_c):
        verts, adaptor, camintr, scale, trans, recov_j, recov_v = ctx.saved_tensors
        batched, center_idx, scale_factor, trans_factor, off_z, res_w, res_h, J = ctx.consts
        if all(g is None for g in (g_j3d, g_v3d, g_rj, g_rv, g_j2d, g_v2d, g_c)):
            return (None,) * 12
        L = _lib.lib()
        B, V = recov_v.shape[:2]
        need = ctx.needs_input_grad
        dev = recov_v.device
        with torch.cuda.device(dev):
            gv = torch.empty_like(recov_v) if need[0] else None
            gj = torch.empty_like(recov_j) if (adaptor is None and need[1]) else None
            ga = torch.empty_like(recov_j) if (adaptor is not None and need[2]) else None
            gs = torch.empty_like(scale) if need[4] else None
            gt = torch.empty_like(trans) if need[5] else None
            _lib.check(L.hoc_hand_head_backward(
                _lib.ptr(recov_v), _lib.ptr(recov_j), _lib.ptr(adaptor), B, V, J, center_idx, _lib.ptr(camintr),
                batched, _lib.ptr(scale), _lib.ptr(trans), scale_factor, trans_factor, off_z, res_w, res_h,
                _lib.ptr(_f32c(g_j3d)), _lib.ptr(_f32c(g_v3d)), _lib.ptr(_f32c(g_rj)), _lib.ptr(_f32c(g_rv)),
                _lib.ptr(_f32c(g_j2d)), _lib.ptr(_f32c(g_v2d)), _lib.ptr(_f32c(g_c)), _lib.ptr(gv), _lib.ptr(gj),
                _lib.ptr(ga), _lib.ptr(gs), _lib.ptr(gt), _lib.stream_ptr()), "hoc_hand_head_backward")
            # the reference freezes the adaptor (meshregnet.py:146-147); an unfrozen one gets its weight gradient
            # from d L / d adapted joints: one small contraction over the batch
            gw = torch.einsum("bjc,bvc->jv", ga, verts) if ga is not None else None
        return gv, gj, gw, None, gs, gt, None, None, None, None, None, None
