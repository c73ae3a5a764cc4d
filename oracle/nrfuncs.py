"""oracle/nrfuncs.py -- TEST INFRASTRUCTURE ONLY.

torch (CPU, any float dtype, autograd-differentiable) restatements of the small out-of-tree helpers
the reference's Renderer / opticalflow call around the rasterizer.  The
sources (`neural_renderer`, `libyana@v0.2.0`) are absent from /root/reference, so these follow the
published behaviour and the reference's call sites:

* ``projection`` ............ nr.projection, called at renderer.py:144,187
* ``vertices_to_faces`` ..... nr.vertices_to_faces, renderer.py:147,160,198,224,256,282
* ``lighting`` .............. nr.lighting, renderer.py:199,257
* ``fill_back`` ............. renderer.py:250-252
* ``batch_proj2d`` .......... libyana.camutils.project.batch_proj2d, opticalflow.py:98-99
* ``batch_vertex_textures`` . libyana.renderutils.textutils, opticalflow.py:103,123 (exact-linear
  fill, SURVEY.md Appendix B-1: corner fill of the upstream helper is unverified)
* ``batch_cat_meshes`` ...... libyana.renderutils.catmesh, warpbranch.py:50-52
PARITY UNPINNED (no reference tests / golden vectors exist for these helpers).
"""
import torch


def projection(vertices, K, R, t, dist_coeffs, orig_size, eps=1e-9):
    vertices = torch.matmul(vertices, R.transpose(2, 1)) + t
    x, y, z = vertices[:, :, 0], vertices[:, :, 1], vertices[:, :, 2]
    x_ = x / (z + eps)
    y_ = y / (z + eps)
    k1, k2, p1, p2, k3 = [dist_coeffs[:, None, i] for i in range(5)]
    r = torch.sqrt(x_ ** 2 + y_ ** 2)
    x__ = x_ * (1 + k1 * (r ** 2) + k2 * (r ** 4) + k3 * (r ** 6)) + 2 * p1 * x_ * y_ + p2 * (r ** 2 + 2 * x_ ** 2)
    y__ = y_ * (1 + k1 * (r ** 2) + k2 * (r ** 4) + k3 * (r ** 6)) + p1 * (r ** 2 + 2 * y_ ** 2) + 2 * p2 * x_ * y_
    vertices = torch.stack([x__, y__, torch.ones_like(z)], dim=-1)
    vertices = torch.matmul(vertices, K.transpose(1, 2))
    u, v = vertices[:, :, 0], vertices[:, :, 1]
    v = orig_size - v
    u = 2 * (u - orig_size / 2.0) / orig_size
    v = 2 * (v - orig_size / 2.0) / orig_size
    return torch.stack([u, v, z], dim=-1)


def vertices_to_faces(vertices, faces):
    bs, nv = vertices.shape[:2]
    faces = faces + (torch.arange(bs, dtype=faces.dtype, device=faces.device) * nv)[:, None, None]
    return vertices.reshape(bs * nv, 3)[faces.long()]


def lighting(faces, textures, intensity_ambient=0.5, intensity_directional=0.5, color_ambient=(1, 1, 1),
             color_directional=(1, 1, 1), direction=(0, 1, 0)):
    bs, nf = faces.shape[:2]
    light = torch.zeros(bs, nf, 3, dtype=faces.dtype)
    color_ambient = torch.as_tensor(color_ambient, dtype=faces.dtype)
    color_directional = torch.as_tensor(color_directional, dtype=faces.dtype)
    direction = torch.as_tensor(direction, dtype=faces.dtype)
    if intensity_ambient != 0:
        light = light + intensity_ambient * color_ambient[None, None, :]
    if intensity_directional != 0:
        v10 = faces[:, :, 0] - faces[:, :, 1]
        v12 = faces[:, :, 2] - faces[:, :, 1]
        normals = torch.nn.functional.normalize(torch.cross(v10, v12, dim=-1), eps=1e-5, dim=-1)
        cos = torch.relu(torch.sum(normals * direction[None, None, :], dim=2))
        light = light + intensity_directional * (color_directional[None, None, :] * cos[:, :, None])
    return textures * light[:, :, None, None, None, :]


def fill_back(faces, textures=None):
    faces = torch.cat((faces, faces.flip(-1)), dim=1)
    if textures is not None:
        textures = torch.cat((textures, textures.permute((0, 1, 4, 3, 2, 5))), dim=1)
    return faces, textures


def batch_proj2d(verts, camintr):
    hom = camintr.bmm(verts.transpose(1, 2)).transpose(1, 2)
    return hom[:, :, :2] / hom[:, :, 2:]


def batch_vertex_textures(faces, vertex_colors):
    """ts=2 cube whose trilinear sample at (b0,b1,b2) is b0*c0+b1*c1+b2*c2:
    T[i,j,k] = i*c0 + j*c1 + k*c2."""
    B, F = faces.shape[:2]
    idx = faces.long()
    cols = torch.stack([vertex_colors[b][idx[b]] for b in range(B)])  # [B,F,3(vertex),3(ch)]
    T = vertex_colors.new_zeros(B, F, 2, 2, 2, 3)
    for i in range(2):
        for j in range(2):
            for k in range(2):
                T[:, :, i, j, k] = i * cols[:, :, 0] + j * cols[:, :, 1] + k * cols[:, :, 2]
    return T


def batch_cat_meshes(verts_list, faces_list):
    offset = 0
    faces_out = []
    for v, f in zip(verts_list, faces_list):
        faces_out.append(f + offset)
        offset += v.shape[1]
    return torch.cat(verts_list, 1), torch.cat(faces_out, 1), None
