"""Generates tests/golden/geom_recover3d.npz by importing the REFERENCE's own
/root/reference/meshreg/models/project.py (recover_3d_proj; torch + an enum module only, so it imports on
torch 2.11) in the build container and running it on the CPU, forward and backward.  /root/reference does
not travel to the GPU box; the fixture does.  The rest of the geometry head (ManoAdaptor, ObjBranch) lives in
modules that import manopth / libyana and cannot be imported.

    python tests/golden/make_geom_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
CASES = {"fphab": (4, 778, 0, (256, 256)), "ho3d": (3, 21, 1, (640, 480)), "tiny": (1, 1, 2, (128, 128))}


def inputs(B, N, seed, input_res):
    """Centred points, FPHAB-like / HO3D-like intrinsics, network-sized scale / translation."""
    g = torch.Generator().manual_seed(seed)
    pts = torch.randn(B, N, 3, generator=g) * 0.05
    f = 300.0 + 400.0 * torch.rand(B, generator=g)
    K = torch.zeros(B, 3, 3)
    K[:, 0, 0] = f
    K[:, 1, 1] = f * (1 + 0.01 * torch.randn(B, generator=g))
    K[:, 0, 2] = input_res[0] / 2 + 20 * torch.randn(B, generator=g)
    K[:, 1, 2] = input_res[1] / 2 + 20 * torch.randn(B, generator=g)
    K[:, 2, 2] = 1
    scale = torch.randn(B, 1, 1, generator=g) * 2e-4
    trans = torch.randn(B, 1, 2, generator=g) * 30
    w_rec = torch.randn(B, N, 3, generator=g)
    w_c = torch.randn(B, 1, 3, generator=g)
    return pts, K, scale, trans, w_rec, w_c


def main():
    sys.path.insert(0, REF)
    from meshreg.models import project

    out = {}
    for name, (B, N, seed, res) in CASES.items():
        pts, K, scale, trans, w_rec, w_c = inputs(B, N, seed, res)
        pts, scale, trans = [t.clone().requires_grad_(True) for t in (pts, scale, trans)]
        rec, c3d = project.recover_3d_proj(pts, K, scale, trans, input_res=res)
        ((rec * w_rec).sum() + (c3d * w_c).sum()).backward()
        out.update({f"{name}_recons3d": rec.detach().numpy(), f"{name}_c3d": c3d.detach().numpy(),
                    f"{name}_g_pts": pts.grad.numpy(), f"{name}_g_scale": scale.grad.numpy(),
                    f"{name}_g_trans": trans.grad.numpy()})
    np.savez_compressed(os.path.join(HERE, "geom_recover3d.npz"), **out)
    print("wrote geom_recover3d.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
