"""GPU parity of the warp / photometric / occlusion kernels against oracle/warp.py run on CUDA, i.e.
against the very ATen kernels (grid_sampler_2d, element-wise ops) the reference runs in production.
Masks hold exact floating-point tests in the reference (== 1, >= 0.99999), so they are compared
exactly; values 1e-4 abs (they are expected to be identical); flow gradients 1e-3 relative."""
import numpy as np
import pytest
import torch

import helpers
from oracle import warp as owarp

pytestmark = pytest.mark.gpu


def _inputs(B, H, W, seed, flow_scale=3.0):
    g = torch.Generator().manual_seed(seed)
    img_ref = torch.rand(B, 3, H, W, generator=g) - 0.5
    img = torch.rand(B, 3, H, W, generator=g) - 0.5
    flows = []
    for _ in range(2):
        f = torch.randn(B, H, W, 2, generator=g) * flow_scale
        blob = (torch.rand(B, H, W, generator=g) > 0.4).float()  # rendered flows are exactly 0 off the mesh
        flows.append((f * blob[..., None]))
    jit = []
    for _ in range(2):
        j = torch.ones(B, 3, H, W)
        j[:, :, :3] = 0
        j[:, :, :, -5:] = 0
        jit.append(j)
    c = lambda t: t.cuda()
    return c(img_ref), c(img), [c(flows[0]), c(flows[1])], c(jit[0]), c(jit[1])


@pytest.mark.parametrize("B,H,W,seed", [(2, 64, 64, 0), (3, 45, 80, 1), (1, 128, 128, 2)])
@pytest.mark.parametrize("mode", ["bilinear", "nearest"])
def test_warp_matches_aten(B, H, W, seed, mode):
    from handobjectconsist_b200.warping.imgflowarp import warp
    img_ref, _, flows, _, _ = _inputs(B, H, W, seed)
    flow = flows[0].permute(0, 3, 1, 2).contiguous()
    out_o, mask_o = owarp.warp(img_ref, flow, mode=mode)
    out, mask = warp(img_ref, flow, mode=mode)
    assert torch.equal(mask, mask_o)
    assert (out - out_o).abs().max().item() <= 1e-6
    ones = torch.ones_like(img_ref)
    o1, _ = warp(ones, flow, mode=mode)
    o2, _ = owarp.warp(ones, flow, mode=mode)
    assert torch.equal(o1 == 1, o2 == 1)  # the reference's `warpjitter == 1` test is reproduced exactly


@pytest.mark.parametrize("use_backward", [False, True])
@pytest.mark.parametrize("B,H,W,seed", [(2, 64, 64, 0), (3, 45, 80, 1)])
def test_pair_consist_matches_oracle(B, H, W, seed, use_backward):
    from handobjectconsist_b200.warping.imgflowarp import pair_consist
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
    img_ref, img, flows, jit_ref, jit = _inputs(B, H, W, seed)
    fo = [f.clone().requires_grad_(True) for f in flows]
    fm = [f.clone().requires_grad_(True) for f in flows]
    loss_o, masks_o, warps_o, diffs_o = owarp.pair_consist(fo, img_ref, img, jit_ref, jit, use_backward)
    loss, masks, warps, diffs = pair_consist(fm, img_ref, img, jit_ref, jit, PyramidCriterion("l1"), use_backward)
    for i in range(2):
        assert torch.equal(masks[i]["full_mask"], masks_o[i]["full_mask"])
        assert torch.equal(masks[i]["warp_mask"], masks_o[i]["warp_mask"])
        assert torch.equal(masks[i]["flow_mask"], masks_o[i]["flow_mask"])
        assert (warps[i] - warps_o[i]).abs().max().item() <= 1e-6
        assert (diffs[i] - diffs_o[i]).abs().max().item() <= 1e-6
    assert masks_o[0]["full_mask"].float().mean().item() > 0.05
    assert (loss - loss_o).abs().max().item() <= 1e-6
    w = torch.arange(1, B + 1, device="cuda", dtype=torch.float32)
    (loss_o * w).sum().backward()
    (loss * w).sum().backward()
    for i in range(2):
        if fo[i].grad is None:
            assert fm[i].grad is None or fm[i].grad.abs().max().item() == 0
            continue
        assert helpers.rel_err(fm[i].grad.cpu().numpy(), fo[i].grad.cpu().numpy()) < 1e-3


def test_pair_consist_l2_generic_path():
    from handobjectconsist_b200.warping.imgflowarp import pair_consist
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
    img_ref, img, flows, jit_ref, jit = _inputs(2, 40, 56, 4)
    fm = [f.clone().requires_grad_(True) for f in flows]
    fo = [f.clone().requires_grad_(True) for f in flows]
    loss, _, _, _ = pair_consist(fm, img_ref, img, jit_ref, jit, PyramidCriterion("l2"), True)
    # oracle with an L2 criterion
    w1, m1 = owarp.warp(img_ref, fo[1].permute(0, 3, 1, 2))
    wj, _ = owarp.warp(jit, fo[1].permute(0, 3, 1, 2))
    valid = (m1 * (wj == 1).float())[:, 0].bool() & ~(fo[1] == 0)[..., 0] & (jit[:, 0] == 1)
    m = valid.unsqueeze(1).repeat(1, 3, 1, 1).float()
    l_fwd = (m * (w1 - img) ** 2).flatten(1).sum(1) / m.flatten(1).sum(1).clamp(min=1)
    w2, m2 = owarp.warp(img, fo[0].permute(0, 3, 1, 2))
    wj2, _ = owarp.warp(jit_ref, fo[0].permute(0, 3, 1, 2))
    valid2 = (m2 * (wj2 == 1).float())[:, 0].bool() & ~(fo[0] == 0)[..., 0] & (jit_ref[:, 0] == 1)
    mm = valid2.unsqueeze(1).repeat(1, 3, 1, 1).float()
    l_bwd = (mm * (w2 - img_ref) ** 2).flatten(1).sum(1) / mm.flatten(1).sum(1).clamp(min=1)
    assert (loss - (l_fwd + l_bwd)).abs().max().item() <= 1e-6
    loss.sum().backward()
    (l_fwd + l_bwd).sum().backward()
    for i in range(2):
        assert helpers.rel_err(fm[i].grad.cpu().numpy(), fo[i].grad.cpu().numpy()) < 1e-3


@pytest.mark.parametrize("B,H,W,seed", [(2, 64, 64, 0), (2, 45, 80, 5)])
def test_occlusion_mask_matches_oracle(B, H, W, seed):
    from handobjectconsist_b200.warping.imgflowarp import get_occlusion_mask
    g = torch.Generator().manual_seed(seed)
    m1 = (torch.rand(B, 1, H, W, generator=g) > 0.3).float().cuda()
    m2 = (torch.rand(B, 1, H, W, generator=g) > 0.3).float().cuda()
    # mostly consistent flows: a per-sample translation there and back, plus noise
    c = (torch.randn(B, 3, 1, 1, generator=g) * 3).cuda()
    f12 = (c + (torch.randn(B, 3, H, W, generator=g) * 0.5).cuda()) * m1
    f21 = (-c + (torch.randn(B, 3, H, W, generator=g) * 0.5).cuda()) * m2
    o1, o2 = owarp.get_occlusion_mask(m1, m2, f12, f21)
    r1, r2 = get_occlusion_mask(m1, m2, f12, f21)
    assert 0.02 < o1.mean().item() < 0.98
    assert (r1 != o1).float().mean().item() <= 1e-4  # thresholded at |d| < 0.03: allow ulp-level ties
    assert (r2 != o2).float().mean().item() <= 1e-4


def test_unpack_u8_matches_to_tensor_normalize_bit_for_bit():
    """hoc_unpack_u8 = torchvision's to_tensor (x / 255) followed by normalize(mean, 1) (x - mean) as the reference's
    dataset workers compute them -- on the CPU, where ATen divides (its CUDA kernel multiplies by the reciprocal) -- bit
    for bit, also for sizes and offsets that are not multiples of the 16-byte vector path."""
    from handobjectconsist_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    for n, off in ((16 * 1000, 0), (12345, 0), (777, 3), (5, 1)):
        raw = torch.randint(0, 256, (n + off,), dtype=torch.uint8, generator=g).to(dev)
        src = raw[off:]
        for div, sub in ((255.0, 0.5), (255.0, 0.0)):
            dst = torch.full((n,), float("nan"), device=dev)
            _lib.check(L.hoc_unpack_u8(_lib.ptr(src), _lib.ptr(dst), n, div, sub, _lib.stream_ptr()), "hoc_unpack_u8")
            want = src.cpu().float().div(div).sub(sub)
            assert torch.equal(dst.cpu(), want)
