"""CPU: host-side logic of the reference-shaped modules that needs no kernel -- closed-hand topology
(meshreg/models/manoutils.py:6-33), WarpRegNet's loss mixing / progressive schedule / step counter
(meshreg/models/warpreg.py:81-127) and batch_masked_mean_loss's contract (meshreg/optim/lossutils.py:1-8)."""
import pytest
import torch

from handobjectconsist_b200 import manoutils, synth, warpreg
from handobjectconsist_b200.optim.lossutils import batch_masked_mean_loss


def test_closed_faces_topology():
    hv, hf = synth.hand_template()
    mano_faces = torch.from_numpy(hf[: manoutils.MANO_FACE_NB])
    closed, ignore = manoutils.get_closed_faces(mano_faces)
    assert closed.shape == (1552, 3) and closed.dtype == torch.int64
    assert ignore == list(range(1538, 1552))                       # manoutils.py:31
    assert torch.equal(closed[:1538], mano_faces)
    fan = closed[1538:]
    assert fan[0].tolist() == [92, 38, 122] and fan[-1].tolist() == [214, 215, 121] and fan[5].tolist() == [215, 122, 118]
    assert int(fan.max()) < 778
    # the fan closes a ring: every ring edge (an edge used once inside the fan) is traversed once
    edges = {}
    for a, b, c in fan.tolist():
        for e in ((a, b), (b, c), (c, a)):
            edges[tuple(sorted(e))] = edges.get(tuple(sorted(e)), 0) + 1
    boundary = [e for e, n in edges.items() if n == 1]
    assert len(boundary) == 16 and all(n <= 2 for n in edges.values())
    with pytest.raises(FileNotFoundError):
        manoutils.get_closed_faces(mano_root="/nonexistent")
    with pytest.raises(ValueError):
        manoutils.get_closed_faces(torch.zeros(5, 4))


class _StubNet(torch.nn.Module):
    """Stands in for MeshRegNet: (loss [1], results, losses) per sample."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.tensor(2.0))

    def forward(self, sample):
        reg = self.w * sample["x"]
        return reg.reshape(1), {"recov_handverts3d": None}, {"mano_reg_loss": 0.1 * reg, "shape": None, "pose": reg * 3}


def _net(**kw):
    hv, hf = synth.hand_template()
    net = warpreg.WarpRegNet((48, 27), _StubNet(), mano_faces=torch.from_numpy(hf[:1538]), **kw)
    net.warp_forward = lambda samples, results: (net.model.w * 5.0, {"masks": []})  # no kernels on the CPU box
    return net


def test_warpregnet_progressive_schedule_and_step_count():
    net = _net(lambda_data=1, lambda_consist=0.5, progressive_steps=4)
    assert net.renderer.image_size == 48 and net.renderer.orig_size == 48 and not net.renderer.anti_aliasing
    assert net.hand_ignore_faces == list(range(1538, 1552)) and net.mano_layer.th_faces.shape == (1552, 3)
    batch = {"data": [{"x": torch.tensor(1.0)}, {"x": torch.tensor(3.0)}], "supervision": ["consist"]}
    expect_lc = [0.0, 0.125, 0.25, 0.375, 0.5, 0.5]            # min(0.5 * step / 4, 0.5)
    for step, lc in enumerate(expect_lc):
        assert net.step_count == step
        assert net.consist_weights() == pytest.approx((1 - lc, lc))
        loss, agg, results, pair = net(batch)
        mano_reg = (0.1 * 2.0 * 1.0 + 0.1 * 2.0 * 3.0) / 2
        assert loss.item() == pytest.approx((1 - lc) * mano_reg + lc * 10.0)
        assert agg["warp_consist"].item() == pytest.approx(10.0)
        assert "shape" not in agg and agg["pose"].item() == pytest.approx((6.0 + 18.0) / 2)
        assert "reg_loss" not in agg and pair == {"masks": []} and len(results) == 2
    # data-only batches do not advance the schedule and use the data weight only
    before = net.step_count
    loss, agg, _, pair = net({"data": batch["data"], "supervision": ["data"]})
    assert net.step_count == before and pair is None
    assert agg["reg_loss"].item() == pytest.approx(4.0) and loss.item() == pytest.approx(0.5 * 4.0)
    loss.backward()
    assert net.model.w.grad is not None


def test_warpregnet_fixed_weights():
    net = _net(lambda_data=2, lambda_consist=3, progressive_consist=False)
    batch = {"data": [{"x": torch.tensor(1.0)}, {"x": torch.tensor(1.0)}], "supervision": ["data", "consist"]}
    loss, agg, _, _ = net(batch)
    assert net.consist_weights() == (2, 3) and net.step_count == 1
    assert loss.item() == pytest.approx(2 * 2.0 + 2 * 0.2 + 3 * 10.0)


def test_batch_masked_mean_loss_contract():
    d = torch.arange(2 * 3 * 2 * 2, dtype=torch.float32).reshape(2, 3, 2, 2)
    m = torch.tensor([[[[1, 0], [0, 1]]], [[[0, 0], [0, 0]]]], dtype=torch.bool)      # [2,1,2,2]: broadcast over channels
    out = batch_masked_mean_loss(d, m)
    # the reference divides by the mask's OWN element count (lossutils.py:4): 2, not 6
    assert out[0].item() == pytest.approx(float(d[0, :, 0, 0].sum() + d[0, :, 1, 1].sum()) / 2)
    assert out[1].item() == 0.0                                                       # empty mask: sum 0 / 1
    full = m.expand(2, 3, 2, 2)
    assert batch_masked_mean_loss(d, full)[0].item() == pytest.approx(float((d[0] * full[0]).sum()) / 6)
    soft = torch.full((2, 3, 2, 2), 0.01)                                             # soft weights summing to 0.12 < 1
    assert batch_masked_mean_loss(d, soft)[0].item() == pytest.approx(float(d[0].mean()), rel=1e-5)


def test_line_pass_deals_every_line_once():
    """The line pass of the rasterizer backward (csrc/raster_bwd.cu, hoc_raster_bwd_line_kernel) hands image lines to
    CTAs by centre-out rank: with `lines` lines per CTA (HOC_TUNE_LINE_LINES), line k of CTA y has rank k * G + y -- or,
    folded (HOC_TUNE_LINE_FOLD), k * G + (G - 1 - y) for odd k -- with G = ceil(S / lines) CTAs; ranks >= S are skipped.
    Restated here (same integer formulas, hoc_centre_out of csrc/hoc_common.cuh included): every line of the raster is
    visited exactly once, whatever S (ragged last CTA), `lines` and the fold."""
    def centre_out(k, n):  # hoc_centre_out: lines from the image centre outwards
        return (n >> 1) + (-((k + 1) >> 1) if (k & 1) else (k >> 1))

    for S in (4, 63, 64, 128, 255, 256, 480, 1400):
        assert sorted(centre_out(k, S) for k in range(S)) == list(range(S))
        for lines in range(1, 9):
            G = (S + lines - 1) // lines
            for fold in (0, 1):
                seen = []
                for y in range(G):
                    for k in range(lines):
                        li = k * G + ((G - 1 - y) if (fold and (k & 1)) else y)
                        if li < S:
                            seen.append(centre_out(li, S))
                assert sorted(seen) == list(range(S)), (S, lines, fold)
