"""``ObjBranch`` with the interface of /root/reference/meshreg/models/objbranch.py:10-81: predicted scale /
2-D translation / axis-angle rotation + the canonical object vertices of the sample -> object vertices in the
camera frame and their projections.  batch_rodrigues, the rotation, recover_3d_proj and batch_proj2d are ONE
launch per direction (``hoc_recover_points_forward/backward``).  Like ``warpbranch.forward`` the sample may be keyed by
the reference's own ``BaseQueries`` / ``TransQueries`` enums or by ours: the lookup goes by member NAME."""
from torch import nn

from ._geomhead import _RecoverPointsFunction
from .warpbranch import _base, _trans


class ObjBranch(nn.Module):
    def __init__(self, trans_factor=1, scale_factor=1):
        """trans_factor / scale_factor: scalings that keep the updates of translation and scale comparable
        during training (objbranch.py:11-21)."""
        super().__init__()
        self.trans_factor = trans_factor
        self.scale_factor = scale_factor
        self.inp_res = [256, 256]

    def forward(self, sample, scaletrans=None, scale=None, trans=None, rotaxisang=None):
        """scaletrans ``[B,6]``: scale, 2-D translation, axis-angle rotation (channels 0, 1:3, 3:6); any of them can
        be given separately instead."""
        batch_size = scale.shape[0] if scaletrans is None else scaletrans.shape[0]
        if scale is None:
            scale = scaletrans[:, :1]
        if trans is None:
            trans = scaletrans[:, 1:3]
        if rotaxisang is None:
            rotaxisang = scaletrans[:, 3:]
        dev = rotaxisang.device
        height, width = tuple(_trans(sample, "IMAGE").shape[2:])
        camintr = _trans(sample, "CAMINTR").to(dev)
        consts = (float(self.scale_factor), float(self.trans_factor), 0.4, float(width), float(height))
        flat_scale, flat_trans = scale.reshape(batch_size), trans.reshape(batch_size, 2)

        def head(points):
            return _RecoverPointsFunction.apply(points.to(dev).float(), rotaxisang, camintr, flat_scale, flat_trans,
                                                *consts)

        rotobjverts, objverts3d, pred_objverts2d, center3d = head(_base(sample, "OBJCANVERTS"))
        if any(getattr(key, "name", key) == "OBJCORNERS3D" and type(key).__name__ == "BaseQueries" for key in sample):
            rotobjcorners, recov_objcorners3d, pred_objcorners2d, _ = head(_base(sample, "OBJCANCORNERS"))
        else:
            pred_objcorners2d = recov_objcorners3d = rotobjcorners = None
        return {
            "obj_verts2d": pred_objverts2d,
            "obj_verts3d": rotobjverts,
            "recov_objverts3d": objverts3d,
            "recov_objcorners3d": recov_objcorners3d,
            "obj_scale": scale.view(batch_size, 1, 1) * self.scale_factor,
            "obj_prescale": scale,
            "obj_prerot": rotaxisang,
            "obj_trans": trans.unsqueeze(1) * self.trans_factor,
            "obj_pretrans": trans,
            "obj_corners2d": pred_objcorners2d,
            "obj_corners3d": rotobjcorners,
        }
