"""baseline/ref_equiv/geomhead.py -- BENCHMARK BASELINE ONLY: the geometry head as the reference composes it, op by
op, out of ATen kernels on CUDA tensors (MeshRegNet.recover_mano, /root/reference/meshreg/models/meshregnet.py:191-229;
ObjBranch.forward, objbranch.py:46-77; recover_3d_proj, project.py:5-23; manopth's batch_rodrigues; libyana's
batch_proj2d).  A RESTATEMENT (manopth / libyana are absent); `bench.py` times it beside hoc_hand_head_* /
hoc_recover_points_*.  Never imported by the product package."""
import torch


def _rodrigues(r):
    angle = torch.norm(r + 1e-8, p=2, dim=1, keepdim=True)
    axis = r / angle
    half = angle * 0.5
    q = torch.cat([torch.cos(half), torch.sin(half) * axis], 1)
    q = q / q.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w.pow(2), x.pow(2), y.pow(2), z.pow(2)
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz, 2 * wz + 2 * xy, w2 - x2 + y2 - z2,
                        2 * yz - 2 * wx, 2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], 1).view(-1, 3, 3)


def _proj2d(verts, camintr):
    hom = camintr.bmm(verts.transpose(1, 2)).transpose(1, 2)
    return hom[:, :, :2] / hom[:, :, 2:]


def _recover_3d_proj(points, camintr, est_scale, est_trans, input_res, off_z=0.4):
    B = points.shape[0]
    focal = camintr[:, :1, :1].view(B, 1)
    est_Z0 = focal * est_scale.view(B, 1) + off_z
    cam_centers = camintr[:, :2, 2]
    img_centers = (cam_centers.new_tensor(input_res) / 2).view(1, 2).repeat(B, 1)
    est_XY0 = (est_trans.view(B, 2) + img_centers - cam_centers) * est_Z0 / focal
    c3d = torch.cat([est_XY0, est_Z0], -1).unsqueeze(1)
    return c3d + points, c3d


def hand_head(verts, adaptor_linear, center_idx, camintr, scale, trans, scale_factor, trans_factor, input_res):
    adapt = adaptor_linear(verts.transpose(2, 1)).transpose(1, 2)
    joints3d = adapt - adapt[:, center_idx].unsqueeze(1)
    verts3d = verts - adapt[:, center_idx].unsqueeze(1)
    final_trans = trans.unsqueeze(1) * trans_factor
    final_scale = scale.view(scale.shape[0], 1, 1) * scale_factor
    recov_joints3d, c3d = _recover_3d_proj(joints3d, camintr, final_scale, final_trans, input_res)
    recov_verts3d = verts3d + c3d
    return recov_joints3d, recov_verts3d, _proj2d(recov_joints3d, camintr), _proj2d(verts3d + c3d, camintr)


def obj_head(canverts, camintr, scale, trans, rot, scale_factor, trans_factor, input_res):
    rotmat = _rodrigues(rot)
    rotverts = rotmat.bmm(canverts.float().transpose(1, 2)).transpose(1, 2)
    final_trans = trans.unsqueeze(1) * trans_factor
    final_scale = scale.view(scale.shape[0], 1, 1) * scale_factor
    verts3d, _ = _recover_3d_proj(rotverts, camintr, final_scale, final_trans, input_res)
    return rotverts, verts3d, _proj2d(verts3d, camintr)
