"""Generates tests/golden/warp_*.npz by importing the REFERENCE's own modules
(/root/reference/meshreg/warping/imgflowarp.py, meshreg/optim/{pyramidloss,lossutils}.py) in the build
container and running them on the CPU (ATen CPU kernels).  /root/reference does not travel to the GPU box;
the fixtures do.  `kornia` (imported by pyramidloss for SSIM / pyramids that the default path never
reaches) is absent and stubbed; the hard-coded `.cuda()` calls of pair_consist are neutralised.

    python tests/golden/make_warp_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _stub_kornia():
    k = types.ModuleType("kornia")
    kl = types.ModuleType("kornia.losses")
    kg = types.ModuleType("kornia.geometry")
    kt = types.ModuleType("kornia.geometry.transform")

    class SSIM:  # never constructed for "l1"
        def __init__(self, *a, **k):
            raise NotImplementedError

    class ScalePyramid:
        def __init__(self, *a, **k):
            pass

    kl.SSIM = SSIM
    kt.ScalePyramid = ScalePyramid
    k.losses, k.geometry, kg.transform = kl, kg, kt
    sys.modules.update({"kornia": k, "kornia.losses": kl, "kornia.geometry": kg, "kornia.geometry.transform": kt})


def inputs(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    img_ref = torch.rand(B, 3, H, W, generator=g) - 0.5
    img = torch.rand(B, 3, H, W, generator=g) - 0.5
    flows = []
    for _ in range(2):
        f = torch.randn(B, H, W, 2, generator=g) * 2.5
        blob = (torch.rand(B, H, W, generator=g) > 0.35).float()
        flows.append(f * blob[..., None])
    jits = []
    for k in range(2):
        j = torch.ones(B, 3, H, W)
        j[:, :, : 2 + k] = 0
        j[:, :, :, W - 3 - k:] = 0
        jits.append(j)
    m1 = (torch.rand(B, 1, H, W, generator=g) > 0.3).float()
    m2 = (torch.rand(B, 1, H, W, generator=g) > 0.3).float()
    c = torch.randn(B, 3, 1, 1, generator=g) * 2
    f12 = (c + torch.randn(B, 3, H, W, generator=g) * 0.4) * m1
    f21 = (-c + torch.randn(B, 3, H, W, generator=g) * 0.4) * m2
    return img_ref, img, flows, jits, m1, m2, f12, f21


def main():
    sys.path.insert(0, REF)
    _stub_kornia()
    torch.Tensor.cuda = lambda self, *a, **k: self  # pair_consist hard-codes .cuda() (imgflowarp.py:80-85)
    from meshreg.optim import pyramidloss
    from meshreg.warping import imgflowarp

    for name, (B, H, W, seed) in {"a": (2, 24, 32, 0), "b": (1, 27, 48, 1)}.items():
        img_ref, img, flows, jits, m1, m2, f12, f21 = inputs(B, H, W, seed)
        out = {}
        for mode in ("bilinear", "nearest"):
            o, m = imgflowarp.warp(img_ref, flows[0].permute(0, 3, 1, 2).contiguous(), mode=mode)
            out[f"warp_{mode}_out"], out[f"warp_{mode}_mask"] = o.numpy(), m.numpy()
        crit = pyramidloss.PyramidCriterion("l1")
        for ub in (False, True):
            fl = [f.clone().requires_grad_(True) for f in flows]
            loss, masks, warps, diffs = imgflowarp.pair_consist(fl, img_ref, img, jits[0], jits[1], crit, use_backward=ub)
            (loss * torch.arange(1, B + 1).float()).sum().backward()
            tag = f"pc{int(ub)}"
            out[f"{tag}_loss"] = loss.detach().numpy()
            for i in range(2):
                out[f"{tag}_warp_mask{i}"] = masks[i]["warp_mask"].detach().numpy()
                out[f"{tag}_full_mask{i}"] = masks[i]["full_mask"].numpy()
                out[f"{tag}_flow_mask{i}"] = masks[i]["flow_mask"].numpy()
                out[f"{tag}_warp{i}"] = warps[i].detach().numpy()
                out[f"{tag}_diff{i}"] = diffs[i].detach().numpy()
                out[f"{tag}_grad{i}"] = (fl[i].grad if fl[i].grad is not None else torch.zeros_like(fl[i])).numpy()
        o1, o2 = imgflowarp.get_occlusion_mask(m1, m2, f12, f21)
        out["occl1"], out["occl2"] = o1.numpy(), o2.numpy()
        np.savez_compressed(os.path.join(HERE, f"warp_{name}.npz"), shape=np.array([B, H, W, seed]), **out)
        print("wrote", f"warp_{name}.npz", {k: v.shape for k, v in list(out.items())[:3]})


if __name__ == "__main__":
    main()
