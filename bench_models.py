"""bench_models.py -- the CALLER of the hot path for BASELINE.json configs[3] ("full trainmeshwarp step: ResNet-18
backbone + MANO + render + warp, synthetic FPHAB-shape batch=64, 8xB200 DDP/NCCL").  Benchmark infrastructure: the
backbone and the MLP heads are out of scope as kernels and stay torch / cuDNN; what this file
provides is a network of the reference's shape that produces what ``WarpRegNet`` consumes, so that bench.py can time
the optimisation step the way trainmeshwarp.py runs it (/root/reference/trainmeshwarp.py:181-255,
meshreg/netscripts/epochpassconsist.py:56-68, meshreg/models/warpreg.py:81-127).

``BenchMeshRegNet`` mirrors ``MeshRegNet.forward`` (meshreg/models/meshregnet.py:237-330) for the consistency setting:
image -> ResNet-18 features (512) -> ManoBranch MLP [512, 512] -> pose (3 + 15 PCA) + shape (10) -> ManoLayer
(handobjectconsist_b200.mano, one launch) -> scaletrans head (512 -> 256 -> 3) -> recover_mano geometry
(hoc_hand_head_*) ; object head (512 -> 256 -> 6) -> ObjBranch (hoc_recover_points_*).  Returns the reference's
``(loss, results, losses)`` triple with ``recov_handverts3d`` / ``recov_objverts3d`` and ``mano_reg_loss``.
"""
import torch
from torch import nn

from handobjectconsist_b200 import synth
from handobjectconsist_b200.mano.manolayer import ManoLayer
from handobjectconsist_b200.meshregnet import recover_mano_geometry
from handobjectconsist_b200.objbranch import ObjBranch


def _mlp(sizes):
    layers = []
    for a, b in zip(sizes[:-1], sizes[1:]):
        layers += [nn.Linear(a, b), nn.ReLU()]
    return nn.Sequential(*layers)


def _name(key):
    return getattr(key, "name", key)


def _get(sample, name, kind):
    for key, val in sample.items():
        if _name(key) == name and type(key).__name__ == kind:
            return val
    raise KeyError(name)


class BenchMeshRegNet(nn.Module):
    def __init__(self, mano_comps=15, center_idx=9, trans_factor=100.0, scale_factor=1e-4, lambda_pose_reg=1e-6,
                 lambda_shape=5e-7):
        super().__init__()
        import torchvision

        net = torchvision.models.resnet18(weights=None)  # random init: no network for checkpoints (bench.py `data`)
        net.fc = nn.Identity()
        self.base_net = net
        self.mano_base = _mlp([512, 512, 512])          # ManoBranch.base_layer (manobranch.py:43-50)
        self.pose_reg = nn.Linear(512, mano_comps + 3)   # manobranch.py:53
        self.shape_reg = nn.Linear(512, 10)              # manobranch.py:64
        self.scaletrans_branch = nn.Sequential(_mlp([512, 256]), nn.Linear(256, 3))       # absolutebranch.py
        self.scaletrans_branch_obj = nn.Sequential(_mlp([512, 256]), nn.Linear(256, 6))
        for lin in (self.pose_reg, self.shape_reg, self.scaletrans_branch[1], self.scaletrans_branch_obj[1]):
            nn.init.normal_(lin.weight, std=1e-3)        # small outputs: the heads' biases place the meshes in view
            nn.init.zeros_(lin.bias)
        self.mano_layer = ManoLayer(center_idx=center_idx, flat_hand_mean=False, ncomps=mano_comps, use_pca=True,
                                    model=synth.mano_model(seed=3))
        self.obj_branch = ObjBranch(trans_factor=trans_factor, scale_factor=scale_factor)
        self.center_idx, self.trans_factor, self.scale_factor = center_idx, trans_factor, scale_factor
        self.lambda_pose_reg, self.lambda_shape = lambda_pose_reg, lambda_shape
        # frozen ManoAdaptor (meshregnet.py:147): sparse random rows that sum to one
        g = torch.Generator().manual_seed(0)
        W = torch.rand(21, 778, generator=g) * (torch.rand(21, 778, generator=g) > 0.9).float()
        self.register_buffer("adaptor", W / W.sum(1, keepdim=True))

    def freeze_batchnorm(self):
        """BatchNorm in inference mode (the reference fine-tunes with frozen statistics; no SyncBN under DDP)."""
        for m in self.base_net.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
        return self

    def set_head_bias(self, hand_st, obj_st):
        """Initialise the output biases of the scale / translation heads so that a random-init network places its meshes
        where the synthetic scene has them (mean over the batch); the gradients still flow to every layer."""
        with torch.no_grad():
            self.scaletrans_branch[1].bias.copy_(hand_st.mean(0))
            self.scaletrans_branch_obj[1].bias.copy_(obj_st.mean(0))

    def forward(self, sample):
        image = _get(sample, "IMAGE", "TransQueries").cuda(non_blocking=True)
        camintr = _get(sample, "CAMINTR", "TransQueries").cuda(non_blocking=True)
        height, width = image.shape[2:]
        feats = self.base_net(image)
        mano_feats = self.mano_base(feats)
        pose, shape = self.pose_reg(mano_feats), self.shape_reg(mano_feats)
        verts_mm, joints_mm = self.mano_layer(pose, th_betas=shape)
        st = self.scaletrans_branch(feats)
        mano = recover_mano_geometry({"verts3d": verts_mm / 1000, "joints3d": joints_mm / 1000}, camintr, st[:, :1],
                                     st[:, 1:], adaptor=self.adaptor, mano_center_idx=self.center_idx,
                                     trans_factor=self.trans_factor, scale_factor=self.scale_factor,
                                     input_res=(width, height))
        obj = self.obj_branch(sample, self.scaletrans_branch_obj(feats))
        # pose / shape regularisation (ManoLoss, manobranch.py: the terms the consistency setting keeps)
        mano_reg = self.lambda_pose_reg * (pose[:, 3:] ** 2).mean() + self.lambda_shape * (shape ** 2).mean()
        results = {"recov_handverts3d": mano["recov_handverts3d"], "recov_objverts3d": obj["recov_objverts3d"],
                   "pose": pose, "shape": shape}
        losses = {"mano_reg_loss": mano_reg}
        return mano_reg.reshape(1), results, losses
