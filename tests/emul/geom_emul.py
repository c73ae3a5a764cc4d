"""Host restatement (numpy, float64) of the ADJOINT the geometry-head kernels compute by hand
(handobjectconsist_b200/csrc/geom_head.cu: gh_project_adjoint, gh_camera_adjoint, the centring adjoint of
hoc_hand_head_backward_kernel and the G_R / dR contraction of hoc_recover_points_backward_kernel), step for step
in the kernels' own decomposition.  tests/test_oracle_geom.py checks it against autograd of oracle/geom.py, so the
hand derivation is verified where no GPU exists; the GPU tests then check the kernels themselves."""
import numpy as np


def camera(K, scale, trans, scale_factor, trans_factor, off_z, res):
    """gh_camera: per-sample (C [B,3], f [B], d [B,2])."""
    f = K[:, 0, 0]
    Z0 = f * (scale * scale_factor) + off_z
    d = trans * trans_factor + np.asarray(res, dtype=np.float64)[None] / 2.0 - K[:, :2, 2]
    C = np.concatenate([d * (Z0 / f)[:, None], Z0[:, None]], 1)
    return C, f, d


def camera_adjoint(gC, C, f, d, scale_factor, trans_factor):
    """gh_camera_adjoint: d est_c3d [B,3] -> (d scale [B], d trans [B,2])."""
    gZ0 = gC[:, 2] + (gC[:, 0] * d[:, 0] + gC[:, 1] * d[:, 1]) / f
    return gZ0 * f * scale_factor, gC[:, :2] * (C[:, 2] / f)[:, None] * trans_factor


def project_adjoint(K, p, g2d):
    """gh_project_adjoint: p [B,N,3] (the points the projection was taken at), g2d [B,N,2] -> [B,N,3]."""
    h = np.einsum("bij,bnj->bni", K, p)
    a0, a1 = g2d[..., 0] / h[..., 2], g2d[..., 1] / h[..., 2]
    a2 = -(g2d[..., 0] * (h[..., 0] / h[..., 2]) + g2d[..., 1] * (h[..., 1] / h[..., 2])) / h[..., 2]
    return np.einsum("bji,bnj->bni", K, np.stack([a0, a1, a2], -1))


def hand_head_backward(recov_v, recov_j, W, center_idx, K, scale, trans, scale_factor, trans_factor, off_z, res, g):
    """g: dict of output gradients (joints3d, verts3d, recov_joints3d, recov_verts3d, joints2d, verts2d, center3d),
    any missing.  Returns (grad_verts, grad_joints_in or grad_adapt, grad_scale, grad_trans)."""
    z = lambda k, like: np.zeros_like(like) if g.get(k) is None else g[k]
    C, f, d = camera(K, scale, trans, scale_factor, trans_factor, off_z, res)
    G_rv = z("recov_verts3d", recov_v) + (project_adjoint(K, recov_v, g["verts2d"]) if g.get("verts2d") is not None else 0)
    G_v3d = z("verts3d", recov_v) + G_rv
    G_rj = z("recov_joints3d", recov_j) + (project_adjoint(K, recov_j, g["joints2d"]) if g.get("joints2d") is not None else 0)
    G_a = z("joints3d", recov_j) + G_rj
    gC = G_rv.sum(1) + G_rj.sum(1) + z("center3d", C)
    gs, gt = camera_adjoint(gC, C, f, d, scale_factor, trans_factor)
    if center_idx >= 0:
        G_a = G_a.copy()
        G_a[:, center_idx] -= G_v3d.sum(1) + (z("joints3d", recov_j) + G_rj).sum(1)
    gv = G_v3d + (np.einsum("jv,bjc->bvc", W, G_a) if W is not None else 0)
    return gv, G_a, gs, gt


def recover_points_backward(points, R, dR, K, scale, trans, scale_factor, trans_factor, off_z, res, g):
    """R [B,3,3] or None, dR [B,3(k),3,3] = d R / d rot_k.  g: dict (rot_points, recov_points, points2d, center3d).
    Returns (grad_points, grad_rot or None, grad_scale, grad_trans)."""
    z = lambda k, like: np.zeros_like(like) if g.get(k) is None else g[k]
    C, f, d = camera(K, scale, trans, scale_factor, trans_factor, off_z, res)
    p = points if R is None else np.einsum("bij,bnj->bni", R, points)
    G_rec = z("recov_points", points) + (project_adjoint(K, p + C[:, None], g["points2d"]) if g.get("points2d") is not None else 0)
    gC = G_rec.sum(1) + z("center3d", C)
    gs, gt = camera_adjoint(gC, C, f, d, scale_factor, trans_factor)
    if R is None:
        return G_rec, None, gs, gt
    G_r = G_rec + z("rot_points", points)
    G_R = np.einsum("bni,bnk->bik", G_r, points)
    grad_rot = np.einsum("bik,bcik->bc", G_R, dR)
    return np.einsum("bji,bnj->bni", R, G_r), grad_rot, gs, gt
