"""oracle/mano.py -- TEST INFRASTRUCTURE ONLY (checker; never imported by the product path).

torch (CPU, any float dtype, autograd-differentiable) restatement of ``manopth.manolayer.ManoLayer.forward``
in the configuration the reference uses (``use_pca=True`` or axis-angle input, ``root_rot_mode='axisang'``,
``flat_hand_mean`` either way, ``center_idx`` optional), as called from
/root/reference/meshreg/models/manobranch.py:70-85,139-145 and meshreg/models/warpreg.py:54-60.

PARITY UNPINNED: manopth is an un-pinned git dependency (/root/reference/environment.yml:34) that is absent from
/root/reference and not installable here, and the licensed MANO_RIGHT.pkl is absent too
(/root/reference/README.md:36-53).  The function below restates the published algorithm (Romero et al. 2017,
SMPL-style linear blend skinning; Rodrigues via quaternions with the `+1e-8` of manopth's rodrigues_layer;
16-joint kinematic tree with three levels per finger; fingertip vertices appended; joint reorder table of
meshregnet.py:41-43; centring on ``center_idx``; millimetre outputs) and is exercised on a synthetic
MANO-shaped parameter set (handobjectconsist_b200.synth.mano_model).
"""
import torch

LEV1, LEV2, LEV3 = [1, 4, 7, 10, 13], [2, 5, 8, 11, 14], [3, 6, 9, 12, 15]
REORDER_TRANSFORMS = [0, 1, 6, 11, 2, 7, 12, 3, 8, 13, 4, 9, 14, 5, 10, 15]
REORDER_JOINTS = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]


def quat2mat(quat):
    nq = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = nq[:, 0], nq[:, 1], nq[:, 2], nq[:, 3]
    w2, x2, y2, z2 = w.pow(2), x.pow(2), y.pow(2), z.pow(2)
    wx, wy, wz = w * x, w * y, w * z
    xy, xz, yz = x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).view(-1, 3, 3)


def batch_rodrigues(axisang):
    """[N,3] axis-angle -> [N,9] (manopth rodrigues_layer.batch_rodrigues; also objbranch.py:46)."""
    angle = torch.norm(axisang + 1e-8, p=2, dim=1).unsqueeze(-1)
    normalized = axisang / angle
    angle = angle * 0.5
    quat = torch.cat([torch.cos(angle), torch.sin(angle) * normalized], dim=1)
    return quat2mat(quat).reshape(-1, 9)


def _with_zeros(t34):
    pad = t34.new_zeros(t34.shape[0], 1, 4)
    pad[:, 0, 3] = 1
    return torch.cat([t34, pad], 1)


def mano_forward(model, pose, betas=None, trans=None, use_pca=True, center_idx=None):
    """model: dict of tensors (v_template [V,3], shapedirs [V,3,10], posedirs [V,3,135], j_regressor [16,V],
    weights [V,16], hands_components [45,45], hands_mean [45], tip_ids (5 ints)); pose [B,3+ncomps] (PCA) or
    [B,48] (axis-angle).  Returns (verts [B,V,3], joints [B,21,3]) in millimetres."""
    B = pose.shape[0]
    if use_pca:
        ncomps = pose.shape[1] - 3
        hand = pose[:, 3:].mm(model["hands_components"][:ncomps])
    else:
        hand = pose[:, 3:]
    full_pose = torch.cat([pose[:, :3], model["hands_mean"] + hand], 1)
    rots = batch_rodrigues(full_pose.reshape(-1, 3)).reshape(B, 16, 3, 3)
    root_rot = rots[:, 0]
    eye = torch.eye(3, dtype=pose.dtype)
    pose_map = (rots[:, 1:] - eye).reshape(B, 135)
    if betas is None:
        betas = pose.new_zeros(B, 10)
    v_shaped = torch.matmul(model["shapedirs"], betas.transpose(1, 0)).permute(2, 0, 1) + model["v_template"]
    J = torch.matmul(model["j_regressor"], v_shaped)
    v_posed = v_shaped + torch.matmul(model["posedirs"], pose_map.transpose(0, 1)).permute(2, 0, 1)

    root_j = J[:, 0].reshape(B, 3, 1)
    root_trans = _with_zeros(torch.cat([root_rot, root_j], 2))
    all_rots = rots[:, 1:]
    l1r, l2r, l3r = (all_rots[:, [i - 1 for i in L]] for L in (LEV1, LEV2, LEV3))
    l1j, l2j, l3j = J[:, LEV1], J[:, LEV2], J[:, LEV3]
    transforms = [root_trans.unsqueeze(1)]
    rel = _with_zeros(torch.cat([l1r, (l1j - root_j.transpose(1, 2)).unsqueeze(3)], 3).reshape(-1, 3, 4))
    root_flt = root_trans.unsqueeze(1).repeat(1, 5, 1, 1).reshape(B * 5, 4, 4)
    lev1 = torch.matmul(root_flt, rel)
    transforms.append(lev1.reshape(B, 5, 4, 4))
    rel = _with_zeros(torch.cat([l2r, (l2j - l1j).unsqueeze(3)], 3).reshape(-1, 3, 4))
    lev2 = torch.matmul(lev1, rel)
    transforms.append(lev2.reshape(B, 5, 4, 4))
    rel = _with_zeros(torch.cat([l3r, (l3j - l2j).unsqueeze(3)], 3).reshape(-1, 3, 4))
    lev3 = torch.matmul(lev2, rel)
    transforms.append(lev3.reshape(B, 5, 4, 4))
    results = torch.cat(transforms, 1)[:, REORDER_TRANSFORMS]
    results_global = results

    joint_js = torch.cat([J, J.new_zeros(B, 16, 1)], 2)
    tmp2 = torch.matmul(results, joint_js.unsqueeze(3))
    results2 = (results - torch.cat([tmp2.new_zeros(B, 16, 4, 3), tmp2], 3)).permute(0, 2, 3, 1)
    T = torch.matmul(results2, model["weights"].transpose(0, 1))
    rest_h = torch.cat([v_posed.transpose(2, 1), v_posed.new_ones(B, 1, v_posed.shape[1])], 1)
    verts = (T * rest_h.unsqueeze(1)).sum(2).transpose(2, 1)[:, :, :3]
    jtr = results_global[:, :, :3, 3]
    tips = verts[:, list(model["tip_ids"])]
    jtr = torch.cat([jtr, tips], 1)[:, REORDER_JOINTS]
    if trans is None or bool(torch.norm(trans) == 0):
        if center_idx is not None:
            center = jtr[:, center_idx].unsqueeze(1)
            jtr = jtr - center
            verts = verts - center
    else:
        jtr = jtr + trans.unsqueeze(1)
        verts = verts + trans.unsqueeze(1)
    return verts * 1000, jtr * 1000
