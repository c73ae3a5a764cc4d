"""``batch_masked_mean_loss`` -- same contract as /root/reference/meshreg/optim/lossutils.py:1-8: the mean of a
per-element distance over the elements a mask selects, one value per sample, 0 for a sample whose mask is empty.
(Element-wise torch ops for the generic criteria; the fused L1 path of ``pair_consist`` computes the same quantity
inside ``hoc_warp_photo_forward`` / ``hoc_pair_loss``.)"""


def batch_masked_mean_loss(dists, mask):
    """dists, mask: [B, ...] of the same shape (mask boolean or 0/1).  Returns [B]."""
    weights = mask.float()
    per_sample = weights.flatten(1)
    selected_sum = (per_sample * dists.flatten(1)).sum(dim=1)
    selected_count = per_sample.sum(dim=1)
    # an empty mask would divide 0 by 0: its count is replaced by 1, which makes the sample's loss exactly 0
    return selected_sum / selected_count.clamp(min=1.0)
