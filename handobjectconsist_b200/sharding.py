"""Batch-axis sharding of the consistency path across the GPUs of one box.

Every frame pair is rendered, warped and reduced to its own loss with no cross-sample data flow
(/root/reference/meshreg/optim/lossutils.py:3-7 reduces per sample; /root/reference/meshreg/models/
warpbranch.py:88 means the per-sample values), so rank r simply takes samples [r*B/G, (r+1)*B/G): there is no
collective inside the path.  The only exchanges are the ones data-parallel training needs anyway: the mean of the
per-rank losses (equal shards => global mean) and, outside this package, the backbone gradients.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Half-open sample range of ``rank``; shards are equal (drop_last semantics, trainmeshwarp.py:118,152)."""
    if n % world != 0:
        raise ValueError(f"global batch {n} is not divisible by the number of ranks {world}")
    per = n // world
    return rank * per, (rank + 1) * per


def _slice(value, lo, hi):
    return value[lo:hi] if torch.is_tensor(value) and value.dim() > 0 else value


def shard_samples(samples, all_results, rank, world):
    """Slice the reference's batch layout (list of sample dicts + list of result dicts) along the batch axis."""
    n = next(v.shape[0] for v in samples[0].values() if torch.is_tensor(v) and v.dim() > 0)
    lo, hi = shard_range(n, rank, world)
    s = [{k: _slice(v, lo, hi) for k, v in sample.items()} for sample in samples]
    r = [{k: _slice(v, lo, hi) for k, v in res.items()} for res in all_results]
    return s, r


def global_mean_loss(local_mean_loss):
    """Mean over the global batch from each rank's mean over its equal shard (one scalar all-reduce)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_mean_loss
    out = local_mean_loss.detach().clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out / dist.get_world_size()


def max_over_ranks(values, device=None):
    """Element-wise maximum over ranks of a list of python floats (device timings)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()
