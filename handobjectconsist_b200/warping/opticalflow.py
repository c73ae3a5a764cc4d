"""Mesh-pair -> rendered optical flow -- same surface as /root/reference/meshreg/warping/opticalflow.py
(``get_opticalflows`` :10-48, ``get_opticalflow`` :51-156).

Per-vertex 2-D displacements between the two frames are painted on the mesh as a texture
([dx, dy, 1]) and rendered at each frame; the rendered field drives the image warp.  Reference
quirks are kept: ``mask_flow2`` is re-bound to the un-thresholded alpha inside the no-grad block
(opticalflow.py:139), ``face_index_map`` is not row-flipped by the rasterizer so the ignore mask is
flipped here (:114,132), and only front copies of the ignored faces (indices < F) are matched.
"""
import ctypes
from typing import List

import torch
from torch.autograd import Function

from .. import _config, _lib
from ..meshutils import batch_proj2d, batch_vertex_textures
from . import imgflowarp

_IGNORE_CACHE = {}
_side_stream = _config.side_stream


def _ignore_tensor(ignore_face_idxs, device):
    key = (tuple(int(i) for i in ignore_face_idxs), str(device))
    if key not in _IGNORE_CACHE:
        _IGNORE_CACHE[key] = torch.tensor(key[0], dtype=torch.int32, device=device)
    return _IGNORE_CACHE[key]


class _MeshRasterFunction(Function):
    """Render per-vertex attributes of a mesh: hoc_mesh_gather -> hoc_raster_forward (image layout), and in
    the backward hoc_raster_backward -> hoc_mesh_scatter.  Replaces batch_vertex_textures + fill_back +
    vertices_to_faces + rasterize_rgbad and their autograd adjoints with four launches."""

    @staticmethod
    def forward(ctx, verts_ndc, attrs, faces_idx, image_size, near, far, eps, background_color, fill_back):
        _lib.require_cuda(verts_ndc, attrs, faces_idx, what="mesh render")
        L = _lib.lib()
        v = verts_ndc.detach().contiguous().float()
        a = attrs.detach().contiguous().float()
        fi = faces_idx.detach().contiguous().long()
        B, V = v.shape[:2]
        Fn = fi.shape[1]
        Fo = 2 * Fn if fill_back else Fn
        S = int(image_size)
        dev = v.device
        bg = (ctypes.c_float * 3)(*[float(c) for c in background_color])
        with torch.cuda.device(dev):
            faces = torch.empty((B, Fo, 3, 3), dtype=torch.float32, device=dev)
            # three vertex values per face: the forward evaluates the cube T[i,j,k] = i c0 + j c1 + k c2 on the fly
            tex = torch.empty((B, Fo, 3, 3), dtype=torch.float32, device=dev)
            st = _lib.stream_ptr()
            # the gather also 0xff-fills the z-buffer keys of the forward below (one graph node less between them)
            ws_bytes = L.hoc_raster_forward_workspace_bytes(B, Fo, S)
            ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
            _lib.check(L.hoc_mesh_gather_clear(_lib.ptr(v), _lib.ptr(a), _lib.ptr(fi), B, V, Fn, int(fill_back),
                                               _lib.HOC_TEX_GRAD_VERTEX, _lib.ptr(faces), _lib.ptr(tex), _lib.ptr(ws),
                                               ws_bytes, st),
                       "hoc_mesh_gather_clear")
            rgb = torch.empty((B, 3, S, S), dtype=torch.float32, device=dev)
            alpha = torch.empty((B, S, S), dtype=torch.float32, device=dev)
            depth = torch.empty((B, S, S), dtype=torch.float32, device=dev)
            idx = torch.empty((B, S, S), dtype=torch.int32, device=dev)
            # depth and wmap are saved for the backward, which reads them at covered pixels: the forward writes them
            # there only (HOC_LAYOUT_SPARSE_SAVED) -- the depth this Function returns is NOT a full depth map
            wmap = torch.empty((B, S, S, 3), dtype=torch.float32, device=dev)
            _lib.check(L.hoc_raster_forward(_lib.ptr(faces), _lib.ptr(tex), B, Fo, S, 2, float(near), float(far),
                                            float(eps), bg, None,
                                            _lib.HOC_LAYOUT_IMAGE | _lib.HOC_LAYOUT_KEYS_CLEARED | _lib.HOC_LAYOUT_TEX_VERTEX
                                            | _lib.HOC_LAYOUT_SPARSE_SAVED,
                                            _lib.ptr(rgb),
                                            _lib.ptr(alpha), _lib.ptr(depth), _lib.ptr(idx), _lib.ptr(wmap), None,
                                            _lib.ptr(ws), ws_bytes, st), "hoc_raster_forward")
        ctx.save_for_backward(faces, fi, idx, rgb, wmap, depth)
        ctx.cfg = (B, V, Fn, Fo, S, float(near), float(far), float(eps), bool(fill_back))
        ctx.mark_non_differentiable(idx)
        ctx.set_materialize_grads(False)
        return rgb, alpha, depth, idx

    @staticmethod
    def backward(ctx, g_rgb, g_alpha, g_depth, g_idx):
        faces, fi, idx, rgb, wmap, depth = ctx.saved_tensors
        B, V, Fn, Fo, S, near, far, eps, fill_back = ctx.cfg
        need_v, need_a = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_v or need_a) or (g_rgb is None and g_alpha is None and g_depth is None):
            return (None,) * 9
        L = _lib.lib()
        dev = faces.device
        c = lambda g: None if g is None else g.contiguous().float()
        g_rgb, g_alpha, g_depth = c(g_rgb), c(g_alpha), c(g_depth)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            grad_faces = torch.empty_like(faces) if need_v else None
            # the cubes come from three vertex values per face: ask for d loss / d vertex value directly
            grad_tex = torch.empty((B, Fo, 3, 3), dtype=torch.float32, device=dev) if need_a else None
            ws_bytes = L.hoc_raster_backward_workspace_bytes_ex(B, Fo, S, 2, _lib.HOC_TEX_GRAD_VERTEX)
            ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
            _lib.check(L.hoc_raster_backward(_lib.ptr(faces), None, _lib.ptr(idx), _lib.ptr(rgb), _lib.ptr(wmap),
                                             _lib.ptr(depth), _lib.ptr(g_rgb),
                                             _lib.ptr(g_alpha), _lib.ptr(g_depth), B, Fo, S, 2, near, far, eps,
                                             _lib.HOC_LAYOUT_IMAGE, 1, _lib.HOC_TEX_GRAD_VERTEX, _lib.ptr(grad_faces),
                                             _lib.ptr(grad_tex), _lib.ptr(ws), ws_bytes, st), "hoc_raster_backward")
            if need_v and need_a:  # adjacent: hoc_mesh_scatter zero-fills both with one memset
                both = torch.empty((2, B, V, 3), dtype=torch.float32, device=dev)
                grad_verts, grad_attrs = both[0], both[1]
            else:
                grad_verts = torch.empty((B, V, 3), dtype=torch.float32, device=dev) if need_v else None
                grad_attrs = torch.empty((B, V, 3), dtype=torch.float32, device=dev) if need_a else None
            sc_bytes = L.hoc_mesh_scatter_workspace_bytes(B, V)  # 0 unless the reproducible mode is on
            sc_ws = torch.empty(sc_bytes, dtype=torch.uint8, device=dev) if sc_bytes else None
            _lib.check(L.hoc_mesh_scatter_ws(_lib.ptr(grad_faces), _lib.ptr(grad_tex), _lib.ptr(fi), B, V, Fn,
                                             int(fill_back), _lib.HOC_TEX_GRAD_VERTEX, _lib.ptr(grad_verts),
                                             _lib.ptr(grad_attrs), 0, _lib.ptr(sc_ws), sc_bytes, st),
                       "hoc_mesh_scatter")
        return (grad_verts, grad_attrs) + (None,) * 7


class _FlowFinalizeFunction(Function):
    """opticalflow.py:109-154 after the two renders, one launch (hoc_flow_finalize)."""

    @staticmethod
    def forward(ctx, rgb1, alpha1, idx1, rgb2, alpha2, idx2, ignore, out_hw, mask_occlusions):
        L = _lib.lib()
        B, _, S, _ = rgb1.shape
        H, W = out_hw
        dev = rgb1.device
        with torch.cuda.device(dev):
            flow12 = torch.empty((B, H, W, 2), dtype=torch.float32, device=dev)
            flow21 = torch.empty((B, H, W, 2), dtype=torch.float32, device=dev)
            mult1 = torch.empty((B, H, W), dtype=torch.float32, device=dev)
            mult2 = torch.empty((B, H, W), dtype=torch.float32, device=dev)
            n_ign = 0 if ignore is None else ignore.numel()
            _lib.check(L.hoc_flow_finalize(_lib.ptr(rgb1), _lib.ptr(alpha1), _lib.ptr(idx1), _lib.ptr(rgb2),
                                           _lib.ptr(alpha2), _lib.ptr(idx2), B, S, H, W, _lib.ptr(ignore), n_ign,
                                           int(mask_occlusions), 0.03, _lib.ptr(flow12), _lib.ptr(flow21),
                                           _lib.ptr(mult1), _lib.ptr(mult2), _lib.stream_ptr()), "hoc_flow_finalize")
        ctx.save_for_backward(mult1, mult2)
        ctx.cfg = (B, S, H, W)
        ctx.set_materialize_grads(False)
        return flow12, flow21

    @staticmethod
    def backward(ctx, g12, g21):
        mult1, mult2 = ctx.saved_tensors
        B, S, H, W = ctx.cfg
        L = _lib.lib()
        dev = mult1.device
        outs = []
        with torch.cuda.device(dev):
            if g12 is not None and g21 is not None and ctx.needs_input_grad[0] and ctx.needs_input_grad[3]:
                grad_rgb = torch.empty((2, B, 3, S, S), dtype=torch.float32, device=dev)  # both directions, one launch
                _lib.check(L.hoc_flow_finalize_backward_pair(_lib.ptr(g12.contiguous().float()), _lib.ptr(mult1),
                                                             _lib.ptr(g21.contiguous().float()), _lib.ptr(mult2), B, S, H,
                                                             W, _lib.ptr(grad_rgb[0]), _lib.ptr(grad_rgb[1]),
                                                             _lib.stream_ptr()), "hoc_flow_finalize_backward_pair")
                return grad_rgb[0], None, None, grad_rgb[1], None, None, None, None, None
            for g, mult, need in ((g12, mult1, ctx.needs_input_grad[0]), (g21, mult2, ctx.needs_input_grad[3])):
                if g is None or not need:
                    outs.append(None)
                    continue
                grad_rgb = torch.empty((B, 3, S, S), dtype=torch.float32, device=dev)
                _lib.check(L.hoc_flow_finalize_backward(_lib.ptr(g.contiguous().float()), _lib.ptr(mult), B, S, H, W,
                                                        _lib.ptr(grad_rgb), _lib.stream_ptr()),
                           "hoc_flow_finalize_backward")
                outs.append(grad_rgb)
        return outs[0], None, None, outs[1], None, None, None, None, None


class _FlowVertexFunction(Function):
    """batch_proj2d x2 + displacement attributes + nr.projection x2 in one launch each way
    (hoc_flow_vertices / hoc_flow_vertices_backward)."""

    @staticmethod
    def forward(ctx, verts1, verts2, K1, K2, R, t, dist, orig_size):
        _lib.require_cuda(verts1, verts2, K1, K2, what="get_opticalflow")
        L = _lib.lib()
        c = lambda x: x.detach().contiguous().float()
        v1, v2, K1, K2, R, t, dist = c(verts1), c(verts2), c(K1), c(K2), c(R), c(t), c(dist)
        B, V = v1.shape[:2]
        cams = (K1.reshape(-1, 9), K2.reshape(-1, 9), R.reshape(-1, 9), t.reshape(-1, 3), dist.reshape(-1, 5))
        for cam in cams:
            if cam.shape[0] not in (1, B):
                raise ValueError("camera tensors must have batch dimension 1 or B")
        flags = [int(cam.shape[0] == B and B > 1) for cam in cams]
        dev = v1.device
        with torch.cuda.device(dev):
            outs = [torch.empty((B, V, 3), dtype=torch.float32, device=dev) for _ in range(4)]
            _lib.check(L.hoc_flow_vertices(_lib.ptr(v1), _lib.ptr(v2), _lib.ptr(cams[0]), flags[0], _lib.ptr(cams[1]),
                                           flags[1], _lib.ptr(cams[2]), flags[2], _lib.ptr(cams[3]), flags[3],
                                           _lib.ptr(cams[4]), flags[4], float(orig_size), B, V, *[_lib.ptr(o) for o in outs],
                                           _lib.stream_ptr()), "hoc_flow_vertices")
        ctx.save_for_backward(v1, v2, *cams)
        ctx.cfg = (B, V, flags, float(orig_size))
        ctx.set_materialize_grads(False)
        # ndc_k depends on verts_k only: a detached frame (first_only / gt_refs, warpbranch.py:45-55) must not
        # make its render look differentiable, or its geometry backward would run for nothing
        nondiff = [o for o, need in ((outs[0], ctx.needs_input_grad[0]), (outs[1], ctx.needs_input_grad[1])) if not need]
        if nondiff:
            ctx.mark_non_differentiable(*nondiff)
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_ndc1, g_ndc2, g_a12, g_a21):
        v1, v2, K1, K2, R, t, dist = ctx.saved_tensors
        B, V, flags, orig_size = ctx.cfg
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need1 or need2):
            return (None,) * 8
        L = _lib.lib()
        c = lambda g: None if g is None else g.contiguous().float()
        g_ndc1, g_ndc2, g_a12, g_a21 = c(g_ndc1), c(g_ndc2), c(g_a12), c(g_a21)
        dev = v1.device
        with torch.cuda.device(dev):
            gv1 = torch.empty_like(v1) if need1 else None
            gv2 = torch.empty_like(v2) if need2 else None
            _lib.check(L.hoc_flow_vertices_backward(
                _lib.ptr(v1), _lib.ptr(v2), _lib.ptr(K1), flags[0], _lib.ptr(K2), flags[1], _lib.ptr(R), flags[2],
                _lib.ptr(t), flags[3], _lib.ptr(dist), flags[4], orig_size, B, V, _lib.ptr(g_ndc1), _lib.ptr(g_ndc2),
                _lib.ptr(g_a12), _lib.ptr(g_a21), _lib.ptr(gv1), _lib.ptr(gv2), _lib.stream_ptr()),
                "hoc_flow_vertices_backward")
        return (gv1, gv2) + (None,) * 6


def _fused_path_ok(neurenderer):
    from ..neurender.renderer import Renderer
    return (isinstance(neurenderer, Renderer) and neurenderer.camera_mode == "projection"
            and not neurenderer.anti_aliasing and neurenderer.no_light)


def _get_opticalflow_fused(verts_cam, faces, camintrs, neurenderer, orig_img_size, mask_occlusions, detach_textures,
                           detach_renders, ignore_face_idxs):
    """Same results as the op-by-op path below with ~10 launches instead of ~150."""
    S = neurenderer.image_size
    dev, dt = verts_cam[0].device, verts_cam[0].dtype
    R = neurenderer.R if neurenderer.R is not None else torch.eye(3, dtype=dt, device=dev)[None]
    t = neurenderer.t if neurenderer.t is not None else torch.zeros(1, 3, dtype=dt, device=dev)
    dist = neurenderer.dist_coeffs if neurenderer.dist_coeffs is not None else torch.zeros(1, 5, dtype=dt, device=dev)
    ndc1, ndc2, attrs12, attrs21 = _FlowVertexFunction.apply(verts_cam[0], verts_cam[1], camintrs[0], camintrs[1], R,
                                                             t, dist, neurenderer.orig_size)
    if detach_textures:
        attrs12 = attrs12.detach()  # the reference detaches only the first set (opticalflow.py:104-105)
    # the two renders are independent: the second one runs on a side stream so that their (small, tail-heavy)
    # kernels overlap; autograd replays the same stream assignment in the backward
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev) if _config.overlap_streams else main
    renders = [None, None]
    if detach_renders:
        ndc1, ndc2 = ndc1.detach(), ndc2.detach()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        renders[1] = _MeshRasterFunction.apply(ndc2, attrs21, faces, S, neurenderer.near, neurenderer.far,
                                               neurenderer.rasterizer_eps, neurenderer.background_color,
                                               neurenderer.fill_back)
        for t in renders[1]:
            t.record_stream(main)
    renders[0] = _MeshRasterFunction.apply(ndc1, attrs12, faces, S, neurenderer.near, neurenderer.far,
                                           neurenderer.rasterizer_eps, neurenderer.background_color,
                                           neurenderer.fill_back)
    main.wait_stream(side)
    W, H = (S, S) if orig_img_size is None else (min(orig_img_size[0], S), min(orig_img_size[1], S))
    ignore = None if ignore_face_idxs is None else _ignore_tensor(ignore_face_idxs, verts_cam[0].device)
    (rgb1, alpha1, _, idx1), (rgb2, alpha2, _, idx2) = renders
    flow12, flow21 = _FlowFinalizeFunction.apply(rgb1, alpha1, idx1, rgb2, alpha2, idx2, ignore, (H, W),
                                                 mask_occlusions)
    return [flow12, flow21]


def get_opticalflows(verts_cam: List[torch.Tensor], faces: torch.Tensor, camintrs: List[torch.Tensor], neurenderer,
                     orig_img_size=None, detach_textures: bool = False, detach_renders: bool = False,
                     ignore_face_idxs=None):
    """Flows between the first mesh and every other one (opticalflow.py:10-48)."""
    all_flows = []
    for vert_world, camintr in zip(verts_cam[1:], camintrs[1:]):
        flows = get_opticalflow([verts_cam[0], vert_world], faces, [camintrs[0], camintr], neurenderer,
                                orig_img_size=orig_img_size, detach_textures=detach_textures,
                                detach_renders=detach_renders, ignore_face_idxs=ignore_face_idxs)
        all_flows.append(flows)
    return all_flows


def _ignore_mask(face_index_map, ignore_face_idxs):
    """1 where the covering face is NOT in the ignore list, rows flipped to image order
    (opticalflow.py:110-115; ``isin`` replaces the [B,S,S,len(ignore)] abs-min temporary)."""
    ign = torch.as_tensor(list(ignore_face_idxs), dtype=face_index_map.dtype, device=face_index_map.device)
    keep = ~torch.isin(face_index_map, ign)
    return keep.flip(1).float().unsqueeze(1)


def get_opticalflow(verts_cam: List[torch.Tensor], faces: torch.Tensor, camintrs: List[torch.Tensor], neurenderer,
                    orig_img_size=None, mask_occlusions: bool = True, detach_textures: bool = False,
                    detach_renders: bool = True, ignore_face_idxs=None):
    """
    Rendered flow 1->2 at the pixels of mesh 1 and 2->1 at the pixels of mesh 2 (opticalflow.py:51-156).
    Returns [pred_flow12, pred_flow21], each [B,H,W,2] (cropped to ``orig_img_size`` = (W, H)).

    With the renderer WarpRegNet builds (projection camera, no anti-aliasing, no lighting --
    /root/reference/meshreg/models/warpreg.py:40-51) the whole function runs as fused kernels
    (``_get_opticalflow_fused``); any other renderer configuration takes the op-by-op path below, which
    mirrors the reference line by line on top of ``Renderer`` / ``get_occlusion_mask``.
    """
    if _fused_path_ok(neurenderer) and not getattr(neurenderer, "force_unfused", False):
        return _get_opticalflow_fused(verts_cam, faces, camintrs, neurenderer, orig_img_size, mask_occlusions,
                                      detach_textures, detach_renders, ignore_face_idxs)
    locs2d_1 = batch_proj2d(verts_cam[0], camintrs[0])
    locs2d_2 = batch_proj2d(verts_cam[1], camintrs[1])
    displ_12 = locs2d_2 - locs2d_1
    sample_flows = torch.cat([displ_12, torch.ones_like(displ_12[:, :, :1])], -1)
    all_textures = batch_vertex_textures(faces, sample_flows)
    if detach_textures:
        all_textures = all_textures.detach()

    renderout = neurenderer(verts_cam[0], faces, all_textures, K=camintrs[0], detach_renders=detach_renders)
    mask_flow1 = (renderout["alpha"].unsqueeze(1) > 0.99999).float()
    if ignore_face_idxs is not None:
        mask_flow1 = mask_flow1 * _ignore_mask(renderout["face_index_map"], ignore_face_idxs)
    pred_flow12 = renderout["rgb"] * mask_flow1

    displ_21 = locs2d_1 - locs2d_2
    sample_flows = torch.cat([displ_21, torch.ones_like(displ_21[:, :, :1])], -1)
    all_textures = batch_vertex_textures(faces, sample_flows)

    renderout = neurenderer(verts_cam[1], faces, all_textures, K=camintrs[1], detach_renders=detach_renders)
    mask_flow2 = (renderout["alpha"].unsqueeze(1) > 0.99999).float()
    if ignore_face_idxs is not None:
        mask_flow2 = mask_flow2 * _ignore_mask(renderout["face_index_map"], ignore_face_idxs)
    pred_flow21 = renderout["rgb"] * mask_flow2

    if mask_occlusions:
        with torch.no_grad():
            mask_flow2 = renderout["alpha"].unsqueeze(1)  # sic (opticalflow.py:139)
            occl_mask1, occl_mask2 = imgflowarp.get_occlusion_mask(mask_flow1, mask_flow2, pred_flow12, pred_flow21)
        mask_flow1 = mask_flow1 * occl_mask1.unsqueeze(1)
        mask_flow2 = mask_flow2 * occl_mask2.unsqueeze(1)
        pred_flow12 = pred_flow12 * mask_flow1
        pred_flow21 = pred_flow21 * mask_flow2
    pred_flow12 = pred_flow12.permute(0, 2, 3, 1)[:, :, :, :2]
    pred_flow21 = pred_flow21.permute(0, 2, 3, 1)[:, :, :, :2]
    if orig_img_size is not None:
        pred_flow12 = pred_flow12[:, : orig_img_size[1], : orig_img_size[0]]
        pred_flow21 = pred_flow21[:, : orig_img_size[1], : orig_img_size[0]]
    return [pred_flow12, pred_flow21]
