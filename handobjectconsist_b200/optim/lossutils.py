"""``batch_masked_mean_loss`` -- same contract as /root/reference/meshreg/optim/lossutils.py:1-8: the mean of a
per-element distance over the elements a mask selects, one value per sample, 0 for a sample whose mask is empty.
(Element-wise torch ops for the generic criteria; the fused L1 path of ``pair_consist`` computes the same quantity
inside ``hoc_warp_photo_forward`` / ``hoc_pair_loss``.)"""
import torch


def batch_masked_mean_loss(dists, mask):
    """dists [B, ...]; mask broadcastable against it over the same number of dimensions (e.g. [B,1,H,W] against
    [B,3,H,W], like the reference), boolean, 0/1 or soft weights.  Returns [B]: sum(mask * dists) / sum(mask),
    both over every axis but the first; the denominator counts the mask's OWN elements (a broadcast mask is not
    re-counted per channel) and only an exactly-zero denominator is replaced by 1."""
    weights = mask.float()
    axes = list(range(1, dists.dim()))
    selected_sum = (weights * dists).sum(dim=axes)
    selected_count = weights.sum(dim=axes)
    selected_count = torch.where(selected_count == 0, torch.ones_like(selected_count), selected_count)
    return selected_sum / selected_count
