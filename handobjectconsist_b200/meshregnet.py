"""The geometry of ``MeshRegNet.recover_mano`` (/root/reference/meshreg/models/meshregnet.py:179-235) and
``ManoAdaptor`` (meshregnet.py:23-51): MANO output -> adapted, centred joints and vertices -> camera-space hand
(recover_3d_proj) -> pixel projections, in ONE launch per direction (``hoc_hand_head_forward/backward``).
The backbone, the MLP heads and the supervised losses around it are out of scope."""
import pickle

import torch

from ._geomhead import _HandHeadFunction

# manopth's joint order -> the 21-joint order of the reference (meshregnet.py:41-43)
JOINT_REORDER = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]
TIP_VERTS = (745, 317, 444, 556, 673)


class ManoAdaptor(torch.nn.Module):
    """Bias-free ``Linear(778, 21)`` from the MANO vertices to the dataset's joints (meshregnet.py:23-51), initialised
    from ``load_path`` (a pickle with an ``adaptor`` [21,778] entry, e.g. assets/mano/fhb_skel_centeridx9.pkl) or
    from the layer's joint regressor + fingertip vertices.  ``forward`` keeps the reference's return value
    (``[B,3,21]`` joints, weight drift); the fused path reads ``weight`` directly."""

    def __init__(self, mano_layer, load_path=None):
        super().__init__()
        self.adaptor = torch.nn.Linear(778, 21, bias=False)
        if load_path is not None:
            with open(load_path, "rb") as p_f:
                regressor = torch.Tensor(pickle.load(p_f)["adaptor"])
        else:
            base = mano_layer._buffers["th_J_regressor"]
            tips = base.new_zeros(5, base.shape[1])
            for row, vert in enumerate(TIP_VERTS):
                tips[row, vert] = 1
            regressor = torch.cat([base, tips])[JOINT_REORDER]
        self.register_buffer("J_regressor", regressor)
        self.adaptor.weight.data = self.J_regressor.clone()

    def pinned_weight(self):
        """The weight with the wrist / fingertip rows reset to the regressor (meshregnet.py:48-50)."""
        for idx in (0, 4, 8, 12, 16, 20):
            self.adaptor.weight.data[idx] = self.J_regressor[idx]
        return self.adaptor.weight

    def forward(self, inp):
        from . import _lib

        _lib.require_cuda(inp, what="ManoAdaptor")
        weight = self.pinned_weight()
        B = inp.shape[0]
        zeros = inp.new_zeros(B)
        eye = torch.eye(3, device=inp.device)[None]
        # adapted joints only: the head with no centring (the camera part is unused)
        joints = _HandHeadFunction.apply(inp, None, weight, eye, zeros, inp.new_zeros(B, 2), -1, 1.0, 1.0, 0.4, 0.0,
                                         0.0)[0]
        return joints.transpose(1, 2), weight - self.J_regressor


def recover_mano_geometry(mano_results, camintr, scale, trans, adaptor=None, mano_center_idx=9, trans_factor=1,
                          scale_factor=1, input_res=(256, 256)):
    """meshregnet.py:191-235 without the loss terms.  ``mano_results``: dict with ``verts3d`` [B,778,3] and
    ``joints3d`` [B,21,3] (ManoBranch output, metres); ``scale`` [B,1], ``trans`` [B,2]: the scaletrans head's
    output; ``input_res`` = (width, height).  Returns ``mano_results`` updated with the reference's keys."""
    verts, joints = mano_results["verts3d"], mano_results["joints3d"]
    B = verts.shape[0]
    weight = None
    if adaptor is not None:
        weight = adaptor.pinned_weight() if hasattr(adaptor, "pinned_weight") else adaptor
    j3d, v3d, recov_j, recov_v, j2d, v2d, _ = _HandHeadFunction.apply(
        verts, None if weight is not None else joints, weight, camintr, scale.reshape(B), trans.reshape(B, 2),
        int(mano_center_idx) if weight is not None else -1, float(scale_factor), float(trans_factor), 0.4,
        float(input_res[0]), float(input_res[1]))
    mano_results = dict(mano_results)
    mano_results.update(joints3d=j3d, verts3d=v3d, joints2d=j2d, recov_joints3d=recov_j, recov_handverts3d=recov_v,
                        verts2d=v2d, hand_pretrans=trans, hand_prescale=scale,
                        hand_trans=trans.unsqueeze(1) * trans_factor,
                        hand_scale=scale.view(B, 1, 1) * scale_factor)
    return mano_results
