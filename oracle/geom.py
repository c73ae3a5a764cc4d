"""oracle/geom.py -- TEST INFRASTRUCTURE ONLY (checker; never imported by the product path).

torch (CPU, any float dtype, autograd-differentiable) restatement of the "geometry head" of the reference: what
sits between the network outputs and the meshes ``warpbranch.forward`` renders.

* ``recover_3d_proj`` ...... /root/reference/meshreg/models/project.py:5-23.  **PINNED**: the reference's module is
  importable; ``tests/golden/geom_recover3d.npz`` holds its outputs (``tests/golden/make_geom_golden.py``) and
  ``tests/test_oracle_geom.py`` compares this restatement with them bit for bit.
* ``mano_adaptor`` ......... ManoAdaptor.forward, meshregnet.py:23-51 (a bias-free ``Linear(778, 21)`` applied to the
  transposed vertices; returned joints are ``[B,21,3]``, i.e. after the caller's ``transpose(1, 2)`` at
  meshregnet.py:194).
* ``recover_mano_geometry``  MeshRegNet.recover_mano, meshregnet.py:191-235 without the loss terms (manopth /
  libyana imports make the module itself unimportable here: PARITY UNPINNED beyond ``recover_3d_proj``).
* ``obj_branch`` ........... ObjBranch.forward, objbranch.py:27-81 (``batch_rodrigues`` restated in oracle/mano.py).
"""
import torch

from .mano import batch_rodrigues
from .nrfuncs import batch_proj2d


def recover_3d_proj(objpoints3d, camintr, est_scale, est_trans, off_z=0.4, input_res=(128, 128)):
    """project.py:5-23: centred points + pixel-space scale / translation -> camera-space points and centre."""
    B = objpoints3d.shape[0]
    focal = camintr[:, :1, :1].reshape(B, 1)
    est_scale = est_scale.reshape(B, 1)
    est_trans = est_trans.reshape(B, 2)
    est_Z0 = focal * est_scale + off_z
    cam_centers = camintr[:, :2, 2]
    img_centers = (torch.tensor(input_res, dtype=cam_centers.dtype) / 2).reshape(1, 2).repeat(B, 1)
    est_XY0 = (est_trans + img_centers - cam_centers) * est_Z0 / focal
    est_c3d = torch.cat([est_XY0, est_Z0], -1).unsqueeze(1)
    return est_c3d + objpoints3d, est_c3d


def mano_adaptor(weight, verts):
    """meshregnet.py:47-51 followed by the transpose of :194 -> adapted joints [B,J,3]."""
    return torch.nn.functional.linear(verts.transpose(2, 1), weight).transpose(1, 2)


def recover_mano_geometry(verts3d, joints3d, camintr, scale, trans, adaptor_weight=None, center_idx=9,
                          trans_factor=1.0, scale_factor=1.0, input_res=(256, 256)):
    """meshregnet.py:191-235.  Returns the entries the reference adds to ``mano_results``."""
    res = {}
    if adaptor_weight is not None:
        adapt = mano_adaptor(adaptor_weight, verts3d)
        joints3d = adapt - adapt[:, center_idx].unsqueeze(1)
        verts3d = verts3d - adapt[:, center_idx].unsqueeze(1)
    res["joints3d"], res["verts3d"] = joints3d, verts3d
    final_trans = trans.unsqueeze(1) * trans_factor
    final_scale = scale.reshape(scale.shape[0], 1, 1) * scale_factor
    recov_joints3d, center3d = recover_3d_proj(joints3d, camintr, final_scale, final_trans, input_res=input_res)
    recov_verts3d = verts3d + center3d
    res.update(joints2d=batch_proj2d(recov_joints3d, camintr), recov_joints3d=recov_joints3d,
               recov_handverts3d=recov_verts3d, verts2d=batch_proj2d(recov_verts3d, camintr), hand_pretrans=trans,
               hand_prescale=scale, hand_trans=final_trans, hand_scale=final_scale, center3d=center3d)
    return res


def obj_branch(canverts, camintr, scale, trans, rotaxisang, cancorners=None, trans_factor=1.0, scale_factor=1.0,
               input_res=(256, 256)):
    """objbranch.py:27-81."""
    B = scale.shape[0]
    rotmat = batch_rodrigues(rotaxisang).reshape(B, 3, 3)
    rotverts = rotmat.bmm(canverts.transpose(1, 2)).transpose(1, 2)
    final_trans = trans.unsqueeze(1) * trans_factor
    final_scale = scale.reshape(B, 1, 1) * scale_factor
    objverts3d, center3d = recover_3d_proj(rotverts, camintr, final_scale, final_trans, input_res=input_res)
    res = {"obj_verts2d": batch_proj2d(objverts3d, camintr), "obj_verts3d": rotverts, "recov_objverts3d": objverts3d,
           "obj_scale": final_scale, "obj_prescale": scale, "obj_prerot": rotaxisang, "obj_trans": final_trans,
           "obj_pretrans": trans, "center3d": center3d}
    if cancorners is not None:
        rotcorners = rotmat.bmm(cancorners.transpose(1, 2)).transpose(1, 2)
        res.update(recov_objcorners3d=rotcorners + center3d, obj_corners2d=batch_proj2d(rotcorners + center3d, camintr),
                   obj_corners3d=rotcorners)
    else:
        res.update(recov_objcorners3d=None, obj_corners2d=None, obj_corners3d=None)
    return res
