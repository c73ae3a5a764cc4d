"""GPU: the frame-pair input pipeline (hoc_augment_frame_pair) against oracle/inputpipe.py, which
tests/test_oracle_inputpipe.py pins bit for bit to PIL / torchvision -- the libraries the reference's dataset code runs
(handobjset.py:336-379).  Everything is integer / exactly-rounded float work: results must be IDENTICAL."""
import numpy as np
import pytest
import torch

from oracle import inputpipe as oip

pytestmark = pytest.mark.gpu


def _frames(rng, B, Hs, Ws):
    f = (rng.random((2, B, Hs, Ws, 3)) * 255).astype(np.uint8)
    f[:, :, 0, :6] = [[10, 10, 10], [255, 0, 0], [0, 255, 0], [0, 0, 255], [0, 0, 0], [255, 255, 255]]
    return f


@pytest.mark.parametrize("B,src,res,jitter", [(3, (54, 96), (48, 32), True), (2, (270, 480), (256, 256), True),
                                              (2, (60, 80), (64, 64), False), (1, (33, 47), (20, 11), True)])
def test_augment_frame_pair_matches_pil_oracle(B, src, res, jitter):
    from handobjectconsist_b200 import inputpipe
    rng = np.random.default_rng(B * 7 + res[0])
    Hs, Ws = src
    frames = _frames(rng, B, Hs, Ws)
    affine = np.stack([oip.get_affine_transform((rng.uniform(0.3, 0.7) * Ws, rng.uniform(0.3, 0.7) * Hs),
                                                rng.uniform(0.5, 1.2) * max(Hs, Ws), res, rot=rng.uniform(-3.1, 3.1))[0]
                       for _ in range(B)])
    color = orders = None
    if jitter:
        color = dict(brightness=rng.uniform(0.5, 1.5, B), saturation=rng.uniform(0.5, 1.5, B),
                     hue=rng.uniform(-0.15, 0.15, B), contrast=rng.uniform(0.5, 1.5, B))
        orders = np.stack([[rng.permutation(4) for _ in range(2)] for _ in range(B)])
        orders[0, 1, 2:] = -1  # a shortened chain
    images, masks = inputpipe.augment_frame_pair([torch.from_numpy(frames[0]), torch.from_numpy(frames[1])], affine, res,
                                                 color=color, orders=orders)
    for fr in range(2):
        assert images[fr].shape == (B, 3, res[1], res[0]) and masks[fr].shape == (B, 3, res[1], res[0])
        for b in range(B):
            c = None
            if jitter:
                c = dict(brightness=float(np.float32(color["brightness"][b])), saturation=float(np.float32(color["saturation"][b])),
                         hue=float(color["hue"][b]), contrast=float(np.float32(color["contrast"][b])),
                         order=[int(o) for o in orders[b, fr] if o >= 0])
            img, mask = oip.frame_to_tensors(frames[fr, b], affine[b], res, c)
            np.testing.assert_array_equal(images[fr][b].cpu().numpy(), img)
            np.testing.assert_array_equal(masks[fr][b].cpu().numpy(), mask)
    assert 0.05 < masks[0].mean().item() <= 1.0


def test_augment_writes_into_the_static_buffers_of_a_captured_step():
    """out=: the crop lands directly in preallocated [B,3,H,W] buffers (what GraphedConsistStep.load_frames uses)."""
    from handobjectconsist_b200 import inputpipe
    rng = np.random.default_rng(5)
    B, res = 2, (64, 48)
    frames = _frames(rng, B, 54, 96)
    affine = np.stack([oip.get_affine_transform((48.0, 27.0), 70.0, res, rot=0.2)[0] for _ in range(B)])
    dev = torch.device("cuda:0")
    bufs = ([torch.full((B, 3, 48, 64), 7.0, device=dev) for _ in range(2)], [torch.full((B, 3, 48, 64), 7.0, device=dev) for _ in range(2)])
    pinned = [torch.from_numpy(frames[i]).pin_memory() for i in range(2)]
    images, masks = inputpipe.augment_frame_pair(pinned, affine, res, out=bufs)
    assert images[0] is bufs[0][0] and masks[1] is bufs[1][1]
    img, mask = oip.frame_to_tensors(frames[1, 1], affine[1], res, None)
    np.testing.assert_array_equal(images[1][1].cpu().numpy(), img)
    np.testing.assert_array_equal(masks[1][1].cpu().numpy(), mask)
    with pytest.raises(ValueError):
        inputpipe.augment_frame_pair([torch.zeros(2, 8, 8, 3), torch.zeros(2, 8, 8, 3)], affine, res)
