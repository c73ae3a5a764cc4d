"""Mesh / camera glue the reference takes from ``libyana@v0.2.0`` (absent from /root/reference):
``batch_proj2d`` (libyana.camutils.project, called at meshreg/warping/opticalflow.py:98-99),
``batch_vertex_textures`` (libyana.renderutils.textutils, opticalflow.py:103,123) and
``batch_cat_meshes`` (libyana.renderutils.catmesh, meshreg/models/warpbranch.py:50-52).
Differentiable torch ops on a few thousand vertices; they run on the device of their inputs.
"""
import torch


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
