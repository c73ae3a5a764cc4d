"""``WarpRegNet`` -- same surface as /root/reference/meshreg/models/warpreg.py:12-127: the module trainmeshwarp.py
optimises.  It runs the single-frame network (``model``: MeshRegNet) on every frame of a batch, adds the photometric
consistency loss between the frames (``warp_forward`` -> ``warpbranch.forward``: this package's kernels) and mixes the
two with the progressive schedule of the reference.

Differences to the reference, all opt-in:
* ``mano_faces=`` / ``mano_layer=``: where the closed hand topology comes from when the licence-gated MANO pickle is
  not under ``mano_root`` (the reference always reads it from there, warpreg.py:54-63);
* ``detach_renders=`` (default True = the reference, warpbranch.py:66) lets the rasterizer's geometry gradient flow;
* ``graphed=True``: the consistency step is captured once as a CUDA graph (``graphed.GraphedConsistStep``) and
  replayed -- same loss and gradients, ~5x less host time per step at 16 pairs of 256 x 256.
"""
import torch

from . import manoutils, warpbranch
from .neurender import renderer
from .optim import pyramidloss


class _FaceTable(torch.nn.Module):
    """Holder of ``th_faces`` when no ManoLayer is available (the reference only reads that buffer, warpreg.py:70)."""

    def __init__(self, faces):
        super().__init__()
        self.register_buffer("th_faces", faces)


class WarpRegNet(torch.nn.Module):
    def __init__(self, image_size, model, fill_back=True, use_backward=True, lambda_data=1, lambda_consist=1,
                 criterion="l1", consist_scale=1, first_only=True, gt_refs=True, progressive_consist=True,
                 progressive_steps=1000, mano_faces=None, mano_layer=None, mano_root="assets/mano",
                 detach_renders=True, graphed=False):
        super().__init__()
        self.fill_back = fill_back
        self.use_backward = use_backward
        self.image_size = image_size
        self.lambda_data = lambda_data
        self.lambda_consist = lambda_consist
        self.consist_scale = consist_scale
        self.criterion = pyramidloss.PyramidCriterion(criterion)
        self.first_only = first_only
        self.gt_refs = gt_refs
        self.progressive_consist = progressive_consist
        self.progressive_steps = progressive_steps
        self.detach_renders = detach_renders
        self.step_count = 0
        # the renderer works on the square that contains the image (warpreg.py:29,40-51); flows are cropped back
        side = max(image_size)
        dev = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        self.renderer = renderer.Renderer(image_size=side, R=torch.eye(3, device=dev).unsqueeze(0),
                                          t=torch.zeros(1, 3, device=dev), K=torch.ones(1, 3, 3, device=dev),
                                          orig_size=side, anti_aliasing=False, fill_back=fill_back, near=0.1,
                                          no_light=True, light_intensity_ambient=0.8)
        self.model = model
        if mano_layer is None and mano_faces is None:
            from .mano.manolayer import ManoLayer

            mano_layer = ManoLayer(joint_rot_mode="axisang", use_pca=False, mano_root=mano_root, center_idx=None,
                                   flat_hand_mean=True)
        if mano_layer is not None and mano_faces is None:
            mano_faces = mano_layer.th_faces
        closed_faces, hand_ignore_faces = manoutils.get_closed_faces(mano_faces)
        if mano_layer is None:
            mano_layer = _FaceTable(closed_faces)
        else:
            mano_layer.register_buffer("th_faces", closed_faces.to(mano_layer.th_faces.device))
        self.mano_layer = mano_layer
        self.hand_ignore_faces = hand_ignore_faces
        self.graphed = graphed
        self._graph_step = None

    # ------------------------------------------------------------------------------------------------------------
    def consist_weights(self):
        """(lambda_data, lambda_consist) for the CURRENT step_count (warpreg.py:102-110): the consistency weight
        ramps linearly from 0 to ``lambda_consist`` over ``progressive_steps`` optimisation steps and is taken out of
        the data weight."""
        if not self.progressive_consist:
            return self.lambda_data, self.lambda_consist
        ramp = min(self.lambda_consist * self.step_count / self.progressive_steps, self.lambda_consist)
        return self.lambda_data - ramp, ramp

    def warp_forward(self, samples, all_results):
        kw = dict(gt_refs=self.gt_refs, first_only=self.first_only, hand_ignore_faces=self.hand_ignore_faces,
                  use_backward=self.use_backward)
        if self.graphed and len(samples) == 2:
            from .graphed import GraphedConsistStep

            if self._graph_step is None:
                self._graph_step = GraphedConsistStep(self.renderer, self.criterion, self.image_size,
                                                      self.mano_layer.th_faces, samples, all_results,
                                                      detach_renders=self.detach_renders, **kw)
            return self._graph_step.apply(samples, all_results), None
        return warpbranch.forward(samples, all_results, self.mano_layer.th_faces, self.renderer, self.image_size,
                                  self.criterion, detach_renders=self.detach_renders, **kw)

    def forward(self, batch):
        samples = batch["data"]
        supervision = batch["supervision"]
        per_sample = [self.model(sample) for sample in samples]  # (loss, results, losses) per frame
        mesh_losses = [out[0] for out in per_sample]
        all_results = [out[1] for out in per_sample]
        all_losses = [out[2] for out in per_sample]

        pair_results = None
        if "consist" in supervision:
            warp_loss, pair_results = self.warp_forward(samples, all_results)

        aggregate_losses = {}
        for key, first in all_losses[0].items():
            if first is not None:
                aggregate_losses[key] = torch.stack([losses[key] for losses in all_losses]).mean()

        lambda_data, lambda_consist = self.consist_weights()
        loss = 0
        if "data" in supervision:
            reg_loss = torch.cat(mesh_losses).mean()
            aggregate_losses["reg_loss"] = reg_loss
            loss = loss + lambda_data * reg_loss
        if "consist" in supervision:
            # pose and shape regularisation of every frame, then the consistency term
            mano_reg = torch.stack([losses["mano_reg_loss"] for losses in all_losses]).mean()
            loss = loss + lambda_data * mano_reg
            loss = loss + lambda_consist * warp_loss
            aggregate_losses["warp_consist"] = warp_loss
            self.step_count += 1
        return loss, aggregate_losses, all_results, pair_results
