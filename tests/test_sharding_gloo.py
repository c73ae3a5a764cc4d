"""CPU, world_size 2 over gloo: the multi-GPU host logic -- equal batch shards in the reference's sample layout,
global mean of per-rank mean losses, max-over-ranks timing reduction (what bench.py does under torchrun)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers  # noqa: F401
from handobjectconsist_b200 import sharding
from handobjectconsist_b200.queries import BaseQueries, TransQueries


def _global_batch(n=8):
    g = torch.Generator().manual_seed(0)
    samples, results = [], []
    for _ in range(2):
        samples.append({TransQueries.IMAGE: torch.rand(n, 3, 4, 4, generator=g), TransQueries.JITTERMASK: torch.ones(n, 3, 4, 4),
                        TransQueries.CAMINTR: torch.rand(n, 3, 3, generator=g), BaseQueries.OBJFACES: torch.zeros(n, 5, 3).long(),
                        BaseQueries.OBJVERTS3D: torch.rand(n, 6, 3, generator=g), BaseQueries.HANDVERTS3D: torch.rand(n, 7, 3, generator=g)})
        results.append({"recov_handverts3d": torch.rand(n, 7, 3, generator=g), "recov_objverts3d": torch.rand(n, 6, 3, generator=g)})
    return samples, results


def _per_sample_loss(samples, results):
    # stand-in for the per-sample consistency loss: any function with no cross-sample data flow
    return (samples[0][TransQueries.IMAGE].flatten(1).mean(1) + results[0]["recov_handverts3d"].flatten(1).sum(1)
            - samples[1][BaseQueries.OBJVERTS3D].flatten(1).mean(1))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        samples, results = _global_batch()
        s, r = sharding.shard_samples(samples, results, rank, world)
        assert s[0][TransQueries.IMAGE].shape[0] == 4 and r[1]["recov_objverts3d"].shape[0] == 4
        lo, hi = sharding.shard_range(8, rank, world)
        assert torch.equal(s[1][BaseQueries.HANDVERTS3D], samples[1][BaseQueries.HANDVERTS3D][lo:hi])
        local = _per_sample_loss(s, r).mean()
        glob = sharding.global_mean_loss(local)
        expect = _per_sample_loss(samples, results).mean()
        assert torch.allclose(glob, expect, atol=1e-6), (glob, expect)
        tmax = sharding.max_over_ranks([1.0 + rank, 5.0 - rank])
        assert tmax == [2.0, 5.0]
        out[rank] = float(glob)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    world = 2
    port = 29600 + os.getpid() % 300
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert len(out) == 2 and abs(out[0] - out[1]) < 1e-7


def test_shard_range_rejects_ragged_batches():
    assert sharding.shard_range(32, 3, 8) == (12, 16)
    with pytest.raises(ValueError):
        sharding.shard_range(10, 0, 4)
    assert sharding.max_over_ranks([1.5]) == [1.5]  # no process group: identity


class _TinyNet(torch.nn.Module):
    """Stands in for MeshRegNet under DDP: a per-sample regression loss and 'vertices' that depend on the weights."""

    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Linear(3, 3)

    def forward(self, sample):
        x = sample["x"]
        verts = self.lin(x)
        reg = (verts ** 2).mean()
        return reg.reshape(1), {"recov_handverts3d": verts}, {"mano_reg_loss": 0.1 * reg}


def _ddp_worker(rank, world, port, out):
    """WarpRegNet (warpreg.py:81-127) wrapped in DistributedDataParallel over gloo, the consistency term replaced by a
    per-sample function of the predicted vertices (no kernels on a CPU box): every rank sees its shard of the global
    batch, the all-reduced gradients equal the gradients of the global-batch loss, step_count advances in lock step."""
    from torch.nn.parallel import DistributedDataParallel as DDP

    from handobjectconsist_b200 import synth, warpreg

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        hv, hf = synth.hand_template()

        def make():
            net = warpreg.WarpRegNet((32, 32), _TinyNet(), mano_faces=torch.from_numpy(hf[:1538]), lambda_data=1,
                                     lambda_consist=0.5, progressive_steps=2)
            net.warp_forward = lambda samples, results: ((results[0]["recov_handverts3d"] - samples[1]["x"]).abs().mean(), None)
            return net

        g = torch.Generator().manual_seed(1)
        x0, x1 = torch.rand(8, 5, 3, generator=g), torch.rand(8, 5, 3, generator=g)
        lo, hi = sharding.shard_range(8, rank, world)
        net = make()
        ddp = DDP(net, broadcast_buffers=False)
        single = make()
        single.load_state_dict(net.state_dict())
        for step in range(3):
            batch = {"data": [{"x": x0[lo:hi]}, {"x": x1[lo:hi]}], "supervision": ["consist"]}
            loss, agg, _, _ = ddp(batch)
            ddp.zero_grad()
            loss.backward()
            # the same step on the global batch, one process
            loss_s, _, _, _ = single({"data": [{"x": x0}, {"x": x1}], "supervision": ["consist"]})
            single.zero_grad()
            loss_s.backward()
            for (n, p), (_, q) in zip(net.named_parameters(), single.named_parameters()):
                assert torch.allclose(p.grad, q.grad, atol=1e-6), n
            assert net.step_count == single.step_count == step + 1
            assert abs(float(sharding.global_mean_loss(loss)) - float(loss_s)) < 1e-6
        out[rank] = net.step_count
    finally:
        dist.destroy_process_group()


def test_warpregnet_under_ddp_over_gloo():
    world = 2
    port = 29300 + os.getpid() % 250
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_ddp_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: 3, 1: 3}
