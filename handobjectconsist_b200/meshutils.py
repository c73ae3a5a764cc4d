"""Mesh / camera glue the reference takes from ``libyana@v0.2.0`` (absent from /root/reference):
``batch_proj2d`` (libyana.camutils.project, called at meshreg/warping/opticalflow.py:98-99),
``batch_vertex_textures`` (libyana.renderutils.textutils, opticalflow.py:103,123) and
``batch_cat_meshes`` (libyana.renderutils.catmesh, meshreg/models/warpbranch.py:50-52).
Differentiable torch ops on a few thousand vertices; they run on the device of their inputs.
"""
import torch
from torch.autograd import Function

from . import _lib


def batch_proj2d(verts, camintr):
    """Pinhole projection [B,V,3] x [B,3,3] -> pixel coordinates [B,V,2]."""
    hom2d = camintr.bmm(verts.transpose(1, 2)).transpose(1, 2)
    return hom2d[:, :, :2] / hom2d[:, :, 2:]


def batch_vertex_textures(faces, vertex_colors):
    """Per-vertex 3-vectors -> texture cubes [B,F,2,2,2,3] whose trilinear sample at the
    barycentric coordinates (b0,b1,b2) is b0*c0 + b1*c1 + b2*c2, i.e. T[i,j,k] = i*c0 + j*c1 + k*c2
    (the multilinear extension)."""
    B, Fn = faces.shape[:2]
    V = vertex_colors.shape[1]
    idx = faces.long() + (torch.arange(B, device=faces.device) * V)[:, None, None]
    cols = vertex_colors.reshape(B * V, 3)[idx]  # [B,F,vertex,channel]
    basis = vertex_colors.new_zeros(2, 2, 2, 3)
    basis[1, :, :, 0] = 1
    basis[:, 1, :, 1] = 1
    basis[:, :, 1, 2] = 1
    return torch.einsum("ijkv,bfvc->bfijkc", basis, cols)


def batch_cat_meshes(verts_list, faces_list):
    """Concatenate meshes along the vertex axis, offsetting each face block; returns
    (verts [B,sum V,3], faces [B,sum F,3], None)."""
    offset = 0
    faces_out = []
    for verts, faces in zip(verts_list, faces_list):
        faces_out.append(faces + offset)
        offset += verts.shape[1]
    return torch.cat(verts_list, 1), torch.cat(faces_out, 1), None


class _CatHandObjectFunction(Function):
    """``batch_cat_meshes([hand, obj], [hand_faces, obj_faces])`` for both frames of a pair in one launch
    (hoc_cat_meshes); the adjoint of a concatenation is a pair of slices, so the backward launches nothing."""

    @staticmethod
    def forward(ctx, hand_a, obj_a, hand_b, obj_b, hand_faces, obj_faces):
        L = _lib.lib()
        c = lambda t: t.detach().contiguous().float()
        ha, oa, hb, ob = c(hand_a), c(obj_a), c(hand_b), c(obj_b)
        hf, of = hand_faces.detach().contiguous().long(), obj_faces.detach().contiguous().long()
        B, Vh, Vo = ha.shape[0], ha.shape[1], oa.shape[1]
        if hf.dim() == 3 and hf.shape[0] == 1:
            hf = hf[0]
        Fh, Fo = hf.shape[-2], of.shape[1]
        dev = ha.device
        with torch.cuda.device(dev):
            va = torch.empty((B, Vh + Vo, 3), dtype=torch.float32, device=dev)
            vb = torch.empty((B, Vh + Vo, 3), dtype=torch.float32, device=dev)
            faces = torch.empty((B, Fh + Fo, 3), dtype=torch.int64, device=dev)
            _lib.check(L.hoc_cat_meshes(_lib.ptr(ha), _lib.ptr(oa), _lib.ptr(hb), _lib.ptr(ob), _lib.ptr(hf),
                                        int(hf.dim() == 3), _lib.ptr(of), B, Vh, Vo, Fh, Fo, _lib.ptr(va), _lib.ptr(vb),
                                        _lib.ptr(faces), _lib.stream_ptr()), "hoc_cat_meshes")
        ctx.vh = Vh
        ctx.mark_non_differentiable(faces)
        ctx.set_materialize_grads(False)
        return va, vb, faces

    @staticmethod
    def backward(ctx, g_a, g_b, g_faces):
        Vh = ctx.vh
        sl = lambda g: (None, None) if g is None else (g[:, :Vh], g[:, Vh:])
        (gha, goa), (ghb, gob) = sl(g_a), sl(g_b)
        return gha, goa, ghb, gob, None, None


def cat_hand_object_pair(hand_a, obj_a, hand_b, obj_b, hand_faces, obj_faces):
    """(verts_a, verts_b, faces) of a frame pair: what two ``batch_cat_meshes`` calls return (warpbranch.py:50-52)."""
    _lib.require_cuda(hand_a, obj_a, hand_b, obj_b, hand_faces, obj_faces, what="cat_hand_object_pair")
    return _CatHandObjectFunction.apply(hand_a, obj_a, hand_b, obj_b, hand_faces, obj_faces)
