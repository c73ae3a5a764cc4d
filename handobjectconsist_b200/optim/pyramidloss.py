"""``PyramidCriterion`` -- same surface as /root/reference/meshreg/optim/pyramidloss.py:8-63.

``WarpRegNet`` always builds ``PyramidCriterion(criterion)`` with ``level_nb=1`` (warpreg.py:34), i.e.
``compute`` is |inp - target| (or its square) followed by ``batch_masked_mean_loss``.  For "l1" at
level 1 ``pair_consist`` does not call ``compute`` at all: the criterion is fused into
``hoc_warp_photo_forward``.  Scale pyramids and SSIM need kornia, which the hot path never reaches
at the reference's defaults -- they raise here.
"""
import torch

from . import lossutils


class PyramidCriterion:
    def __init__(self, criterion, geom_weight=1, level_nb=1):
        self.level_nb = level_nb
        self.name = criterion
        if criterion == "l2":
            self.criterion = torch.nn.MSELoss(reduction="none")
        elif criterion == "l1":
            self.criterion = torch.nn.L1Loss(reduction="none")
        elif criterion == "ssim":
            raise NotImplementedError("ssim needs kornia (out of scope of the accelerated path)")
        else:
            raise ValueError(f"{criterion} not in [l2, l1, ssim]")
        if level_nb != 1:
            raise NotImplementedError("scale pyramids (level_nb > 1) need kornia's ScalePyramid")
        self.geom_weight = geom_weight

    def compute(self, inp, target, mask=None):
        diff = self.criterion(inp, target)
        losses = lossutils.batch_masked_mean_loss(diff, mask)
        return [inp], [target], losses, [diff], [mask]
