"""CPU: oracle/warp.py against the golden vectors generated from the REFERENCE's own imgflowarp.py /
lossutils.py (tests/golden/make_warp_golden.py, run where /root/reference exists).  Both sides run ATen's
CPU kernels on identical inputs, so everything is compared bit for bit.  This is what pins the warp /
mask / loss part of the oracle to the reference."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import helpers  # noqa: F401  (sys.path)
from oracle import warp as owarp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_spec = importlib.util.spec_from_file_location("make_warp_golden", os.path.join(GOLD, "make_warp_golden.py"))
_mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mk)


@pytest.mark.parametrize("name", ["a", "b"])
def test_oracle_warp_matches_reference_golden(name):
    gold = np.load(os.path.join(GOLD, f"warp_{name}.npz"))
    B, H, W, seed = [int(v) for v in gold["shape"]]
    img_ref, img, flows, jits, m1, m2, f12, f21 = _mk.inputs(B, H, W, seed)
    for mode in ("bilinear", "nearest"):
        o, m = owarp.warp(img_ref, flows[0].permute(0, 3, 1, 2).contiguous(), mode=mode)
        np.testing.assert_array_equal(o.numpy(), gold[f"warp_{mode}_out"])
        np.testing.assert_array_equal(m.numpy(), gold[f"warp_{mode}_mask"])
    for ub in (False, True):
        fl = [f.clone().requires_grad_(True) for f in flows]
        loss, masks, warps, diffs = owarp.pair_consist(fl, img_ref, img, jits[0], jits[1], use_backward=ub)
        (loss * torch.arange(1, B + 1).float()).sum().backward()
        tag = f"pc{int(ub)}"
        np.testing.assert_array_equal(loss.detach().numpy(), gold[f"{tag}_loss"])
        for i in range(2):
            np.testing.assert_array_equal(masks[i]["warp_mask"].detach().numpy(), gold[f"{tag}_warp_mask{i}"])
            np.testing.assert_array_equal(masks[i]["full_mask"].numpy(), gold[f"{tag}_full_mask{i}"])
            np.testing.assert_array_equal(masks[i]["flow_mask"].numpy(), gold[f"{tag}_flow_mask{i}"])
            np.testing.assert_array_equal(warps[i].detach().numpy(), gold[f"{tag}_warp{i}"])
            np.testing.assert_array_equal(diffs[i].detach().numpy(), gold[f"{tag}_diff{i}"])
            g = fl[i].grad if fl[i].grad is not None else torch.zeros_like(fl[i])
            np.testing.assert_array_equal(g.numpy(), gold[f"{tag}_grad{i}"])
    o1, o2 = owarp.get_occlusion_mask(m1, m2, f12, f21)
    np.testing.assert_array_equal(o1.numpy(), gold["occl1"])
    np.testing.assert_array_equal(o2.numpy(), gold["occl2"])
    assert 0.0 < gold["occl1"].mean() < 1.0 and gold["pc1_full_mask0"].mean() > 0.05


def test_zero_flow_is_not_identity_f6():
    """align_corners=True normalisation + align_corners=False sampling: zero flow samples at
    x*W/(W-1) - 0.5 and masks the border."""
    x = torch.arange(16.0).view(1, 1, 1, 16).repeat(1, 1, 8, 1)
    out, mask = owarp.warp(x, torch.zeros(1, 2, 8, 16))
    assert mask[0, 0, :, 0].sum() == 0 and mask[0, 0, 0, :].sum() == 0  # left column / top row masked
    xs = torch.arange(16.0) * 16 / 15 - 0.5
    inner = mask[0, 0, 4] == 1
    assert torch.allclose(out[0, 0, 4][inner], xs[inner], atol=1e-5)
