"""``recover_3d_proj`` with the signature of /root/reference/meshreg/models/project.py:5-23, on
``hoc_recover_points_forward/backward`` (one launch per direction instead of ~12 ATen kernels)."""
from ._geomhead import _RecoverPointsFunction


def recover_3d_proj(objpoints3d, camintr, est_scale, est_trans, off_z=0.4, input_res=(128, 128)):
    """Given estimated centred points ``[B,N,3]``, camera intrinsics ``[B,3,3]`` and the predicted scale ``[B,...]`` /
    translation ``[B,...,2]`` in pixel space, the points in the camera coordinate system and their centre:
    ``(recons3d [B,N,3], est_c3d [B,1,3])``."""
    B = objpoints3d.shape[0]
    _, recons3d, _, c3d = _RecoverPointsFunction.apply(objpoints3d, None, camintr, est_scale.reshape(B),
                                                       est_trans.reshape(B, 2), 1.0, 1.0, float(off_z),
                                                       float(input_res[0]), float(input_res[1]))
    return recons3d, c3d.unsqueeze(1)
