"""Frame-pair input pipeline on the GPU -- the image side of ``HandObjSet.get_sample``
(/root/reference/meshreg/datasets/handobjset.py:336-379) for the two frames of a pair.

The reference's dataset workers decode a frame with PIL, colour-jitter it, crop / rotate it to ``inp_res`` with
``handutils.transform_img`` (``Image.transform(res, Image.AFFINE, inv(affinetrans))``, nearest sampling), tensorize and
normalize it, and push a white image through the same transform to get the jitter mask -- per sample, on the CPU -- and
then ship four float tensors per pair to the GPU.  ``augment_frame_pair`` does all of it in two launches from the
decoded uint8 frames (a quarter of the bytes over PCIe) and is bit-compatible with the PIL / torchvision calls
(``csrc/input_pipe.cu``; oracle/inputpipe.py is pinned against PIL itself).  The Gaussian blur of
handobjset.py:338-339 is not part of it.

Both frames of a sample share the space augmentation (``affinetrans``) and the colour factors
(handobjset.py:417-421); the ORDER of the four colour adjustments is shuffled per call by
``colortrans.apply_jitter``, so it is an input here, per frame.
"""
import ctypes

import numpy as np
import torch

from . import _lib

OP_BRIGHTNESS, OP_SATURATION, OP_HUE, OP_CONTRAST = 0, 1, 2, 3


def get_affine_transform(center, scale, res, rot=0.0):
    """``libyana.transformutils.handutils.get_affine_transform`` (called at handobjset.py:162; libyana is absent from
    /root/reference -- restated from its published source, hassony2/obman_train): the source -> crop transform for a
    crop of ``scale`` source pixels around ``center``, rotated by ``rot`` radians, to ``res`` = (width, height).
    Returns ``(affinetrans, post_rot_trans)`` as float32 [3,3] arrays like the reference."""
    rot_mat = np.array([[np.cos(rot), -np.sin(rot), 0], [np.sin(rot), np.cos(rot), 0], [0, 0, 1]], dtype=np.float64)
    c = np.array([center[0], center[1], 1.0])
    origin_rot_center = rot_mat.dot(c)[:2]
    shift = np.eye(3)
    shift[0, 2], shift[1, 2] = -res[1] / 2, -res[0] / 2
    unshift = shift.copy()
    unshift[:2, 2] *= -1
    transformed_center = unshift.dot(rot_mat).dot(shift).dot(c)

    def no_rot(origin):
        m = np.zeros((3, 3))
        m[0, 0], m[1, 1], m[2, 2] = float(res[1]) / scale, float(res[0]) / scale, 1
        m[0, 2] = res[1] * (-float(origin[0]) / scale + 0.5)
        m[1, 2] = res[0] * (-float(origin[1]) / scale + 0.5)
        return m

    return no_rot(origin_rot_center).dot(rot_mat).astype(np.float32), no_rot(transformed_center[:2]).astype(np.float32)


def affine_fixed_coefficients(affinetrans):
    """[B,3,3] source -> crop transforms (``TransQueries.AFFINETRANS``) -> [B,6] int32, PIL's 16.16 fixed-point
    coefficients of the inverse map.  Host arithmetic exactly as the reference runs it: ``np.linalg.inv`` in the
    array's own precision (handutils.transform_img), then PIL's double-precision rounding (Geometry.c)."""
    a = affinetrans.detach().cpu().numpy() if torch.is_tensor(affinetrans) else np.asarray(affinetrans)
    inv = np.linalg.inv(a).astype(np.float64)  # (batched: the same LAPACK solve per matrix as the reference's single call)
    ca, cb, cc, cd, ce, cf = inv[:, 0, 0], inv[:, 0, 1], inv[:, 0, 2], inv[:, 1, 0], inv[:, 1, 1], inv[:, 1, 2]
    fix = lambda v: np.floor(v * 65536.0 + 0.5)
    out = np.stack([fix(ca), fix(cb), fix(cc + ca * 0.5 + cb * 0.5), fix(cd), fix(ce), fix(cf + cd * 0.5 + ce * 0.5)], 1)
    if not np.isfinite(out).all() or np.abs(out).max() >= 2 ** 31:
        raise ValueError("affine transform outside PIL's 16.16 fixed-point range")
    return torch.from_numpy(out.astype(np.int32))


class _PinnedRing:
    """A few reusable pinned staging buffers per size: the small per-batch parameters cross PCIe in ONE async copy, and
    a buffer is only rewritten after the copy that read it has completed."""

    def __init__(self, depth=8):
        self.depth, self.slots, self.next = depth, {}, {}

    def get(self, n_int32):
        ring = self.slots.setdefault(n_int32, [])
        if len(ring) < self.depth:
            ring.append((torch.empty(n_int32, dtype=torch.int32).pin_memory(), torch.cuda.Event()))
            return ring[-1]
        i = self.next.get(n_int32, 0)
        self.next[n_int32] = (i + 1) % self.depth
        ring[i][1].synchronize()
        return ring[i]


_PARAM_RING = _PinnedRing()


def augment_frame_pair(frames, affinetrans, inp_res, color=None, orders=None, out=None):
    """
    Args:
        frames: two uint8 tensors [B,Hs,Ws,3] (decoded source frames, HWC); host (pinned: copied asynchronously) or cuda
        affinetrans: [B,3,3] source -> crop transforms (host; numpy or tensor)
        inp_res: (width, height) of the crop; width must be a multiple of 4
        color: None (evaluation: no jitter) or dict of length-B sequences ``brightness``, ``saturation``, ``hue``,
            ``contrast`` (the factors colortrans.get_color_params draws; shared by the pair)
        orders: [B,2,4] ints, per (sample, frame) the order of the adjustments (OP_* ids, -1 = skip); default: the
            identity order for both frames
        out: optional ``(images, masks)``: two lists of two preallocated cuda tensors [B,3,H,W] (e.g. the static buffers
            of a GraphedConsistStep) to write into
    Returns ``(images, jitter_masks)``: two lists of two [B,3,H,W] float32 cuda tensors.
    """
    L = _lib.lib()
    W, H = int(inp_res[0]), int(inp_res[1])
    dev = torch.device("cuda", torch.cuda.current_device())
    if out is not None:
        dev = out[0][0].device
    if frames[0].dtype != torch.uint8 or frames[0].dim() != 4 or frames[0].shape[3] != 3 or frames[1].shape != frames[0].shape:
        raise ValueError("frames must be two uint8 tensors of the same shape [B,Hs,Ws,3]")
    f0, f1 = [f.to(dev, non_blocking=True).contiguous() for f in frames]
    B, Hs, Ws = f0.shape[:3]
    # the per-batch parameters, packed [coef B*6 | colour B*3 (float bits) | hue B | order B*8] -> one pinned buffer
    jitter = color is not None
    n = B * 6 + (B * 12 if jitter else 0)
    pack = np.empty(n, dtype=np.int32)
    pack[: B * 6] = affine_fixed_coefficients(affinetrans).numpy().reshape(-1)
    if jitter:
        f32 = lambda v: np.asarray(v, dtype=np.float32).reshape(B)
        col = np.stack([f32(color["brightness"]), f32(color["saturation"]), f32(color["contrast"])], 1)
        pack[B * 6: B * 9] = col.reshape(-1).view(np.int32)
        hue = np.asarray(color["hue"], dtype=np.float64).reshape(B)
        pack[B * 9: B * 10] = (hue * 255).astype(np.int64) & 0xFF                       # uint8(hue * 255), wrapping
        if orders is None:
            orders = np.tile(np.array([OP_BRIGHTNESS, OP_SATURATION, OP_HUE, OP_CONTRAST]), (B, 2, 1))
        o = orders.cpu().numpy() if torch.is_tensor(orders) else np.asarray(orders)
        pack[B * 10:] = o.astype(np.int32).reshape(B * 8)
    with torch.cuda.device(dev):
        host, event = _PARAM_RING.get(n)
        host.numpy()[:] = pack
        params = host.to(dev, non_blocking=True)
        event.record()
        base = params.data_ptr()
        p = lambda off: ctypes.c_void_p(base + 4 * off)
        if out is None:
            images = [torch.empty((B, 3, H, W), dtype=torch.float32, device=dev) for _ in range(2)]
            masks = [torch.empty((B, 3, H, W), dtype=torch.float32, device=dev) for _ in range(2)]
        else:
            images, masks = out
        ws_bytes = L.hoc_augment_frame_pair_workspace_bytes(B)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        _lib.check(L.hoc_augment_frame_pair(_lib.ptr(f0), _lib.ptr(f1), B, Hs, Ws, p(0), p(B * 6) if jitter else None,
                                            p(B * 9) if jitter else None, p(B * 10) if jitter else None, H, W,
                                            _lib.ptr(images[0]), _lib.ptr(images[1]), _lib.ptr(masks[0]),
                                            _lib.ptr(masks[1]), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()),
                   "hoc_augment_frame_pair")
        params.record_stream(torch.cuda.current_stream(dev))
    return images, masks
