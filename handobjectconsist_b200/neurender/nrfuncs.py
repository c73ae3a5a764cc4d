"""Camera / mesh helpers the reference takes from ``neural_renderer`` (absent from /root/reference):
``projection``, ``vertices_to_faces``, ``lighting``, ``look_at``, ``look``, ``perspective``
(call sites /root/reference/meshreg/neurender/renderer.py:124-147,187-199,224,256-282).

Small differentiable torch ops on the device of their inputs (plumbing around the kernels; they
run on a few thousand vertices).  They restate the published behaviour of
daniilidis-group/neural_renderer.
"""
import math

import torch
import torch.nn.functional as F


def projection(vertices, K, R, t, dist_coeffs, orig_size, eps=1e-9):
    """Pinhole projection to NDC: v' = v R^T + t; x/z, y/z; OpenCV distortion; K; y flip; [-1,1]."""
    vertices = torch.matmul(vertices, R.transpose(2, 1)) + t
    x, y, z = vertices[:, :, 0], vertices[:, :, 1], vertices[:, :, 2]
    x_ = x / (z + eps)
    y_ = y / (z + eps)
    k1 = dist_coeffs[:, None, 0]
    k2 = dist_coeffs[:, None, 1]
    p1 = dist_coeffs[:, None, 2]
    p2 = dist_coeffs[:, None, 3]
    k3 = dist_coeffs[:, None, 4]
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
    """[B,V,3], [B,F,3] int -> [B,F,3,3]."""
    assert vertices.ndimension() == 3 and faces.ndimension() == 3
    assert vertices.shape[0] == faces.shape[0] and vertices.shape[2] == 3 and faces.shape[2] == 3
    bs, nv = vertices.shape[:2]
    faces = faces + (torch.arange(bs, dtype=faces.dtype, device=faces.device) * nv)[:, None, None]
    return vertices.reshape((bs * nv, 3))[faces.long()]


def lighting(faces, textures, intensity_ambient=0.5, intensity_directional=0.5, color_ambient=(1, 1, 1),
             color_directional=(1, 1, 1), direction=(0, 1, 0)):
    """Ambient + ReLU(n . dir) directional light, multiplied into the per-face texture cubes."""
    bs, nf = faces.shape[:2]
    device, dtype = faces.device, faces.dtype
    light = torch.zeros(bs, nf, 3, dtype=dtype, device=device)
    if intensity_ambient != 0:
        color_ambient = torch.as_tensor(color_ambient, dtype=dtype, device=device)
        if color_ambient.ndimension() == 1:
            color_ambient = color_ambient[None, :]
        light = light + intensity_ambient * color_ambient[:, None, :]
    if intensity_directional != 0:
        color_directional = torch.as_tensor(color_directional, dtype=dtype, device=device)
        direction = torch.as_tensor(direction, dtype=dtype, device=device)
        if color_directional.ndimension() == 1:
            color_directional = color_directional[None, :]
        if direction.ndimension() == 1:
            direction = direction[None, :]
        faces = faces.reshape((bs * nf, 3, 3))
        v10 = faces[:, 0] - faces[:, 1]
        v12 = faces[:, 2] - faces[:, 1]
        normals = F.normalize(torch.cross(v10, v12, dim=1), eps=1e-5)
        normals = normals.reshape((bs, nf, 3))
        if direction.ndimension() == 2:
            direction = direction[:, None, :]
        cos = F.relu(torch.sum(normals * direction, dim=2))
        light = light + intensity_directional * (color_directional[:, None, :] * cos[:, :, None])
    light = light[:, :, None, None, None, :]
    return textures * light


def _as_batch_vec(v, like):
    v = torch.as_tensor(v, dtype=like.dtype, device=like.device)
    if v.ndimension() == 1:
        v = v[None, :]
    return v


def look_at(vertices, eye, at=(0, 0, 0), up=(0, 1, 0)):
    """Rotate/translate so the camera at ``eye`` looks at ``at``."""
    eye, at, up = _as_batch_vec(eye, vertices), _as_batch_vec(at, vertices), _as_batch_vec(up, vertices)
    bs = vertices.shape[0]
    z_axis = F.normalize(at - eye, eps=1e-5)
    x_axis = F.normalize(torch.cross(up.expand_as(z_axis), z_axis, dim=1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=1), eps=1e-5)
    r = torch.cat((x_axis[:, None, :], y_axis[:, None, :], z_axis[:, None, :]), dim=1)
    if r.shape[0] != bs:
        r = r.expand(bs, 3, 3)
    vertices = vertices - eye[:, None, :]
    return torch.matmul(vertices, r.transpose(1, 2))


def look(vertices, eye, direction=(0, 1, 0), up=None):
    """Rotate/translate so the camera at ``eye`` looks along ``direction``."""
    if up is None:
        up = (0, 1, 0)
    eye, direction, up = _as_batch_vec(eye, vertices), _as_batch_vec(direction, vertices), _as_batch_vec(up, vertices)
    bs = vertices.shape[0]
    z_axis = F.normalize(direction, eps=1e-5)
    x_axis = F.normalize(torch.cross(up.expand_as(z_axis), z_axis, dim=1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=1), eps=1e-5)
    r = torch.cat((x_axis[:, None, :], y_axis[:, None, :], z_axis[:, None, :]), dim=1)
    if r.shape[0] != bs:
        r = r.expand(bs, 3, 3)
    vertices = vertices - eye[:, None, :]
    return torch.matmul(vertices, r.transpose(1, 2))


def perspective(vertices, angle=30.0):
    """x, y divided by z * tan(angle)."""
    angle = torch.as_tensor(angle, dtype=vertices.dtype, device=vertices.device)
    if angle.ndimension() == 0:
        angle = angle[None]
    width = torch.tan(angle / 180 * math.pi)[:, None]
    z = vertices[:, :, 2]
    x = vertices[:, :, 0] / z / width
    y = vertices[:, :, 1] / z / width
    return torch.stack((x, y, z), dim=2)
