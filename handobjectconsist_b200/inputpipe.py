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
import numpy as np
import torch

from . import _lib

OP_BRIGHTNESS, OP_SATURATION, OP_HUE, OP_CONTRAST = 0, 1, 2, 3


def affine_fixed_coefficients(affinetrans):
    """[B,3,3] source -> crop transforms (``TransQueries.AFFINETRANS``) -> [B,6] int32, PIL's 16.16 fixed-point
    coefficients of the inverse map.  Host arithmetic exactly as the reference runs it: ``np.linalg.inv`` in the
    array's own precision (handutils.transform_img), then PIL's double-precision rounding (Geometry.c)."""
    a = affinetrans.detach().cpu().numpy() if torch.is_tensor(affinetrans) else np.asarray(affinetrans)
    out = np.empty((a.shape[0], 6), dtype=np.int64)
    for i in range(a.shape[0]):
        inv = np.linalg.inv(a[i])
        ca, cb, cc, cd, ce, cf = [np.float64(v) for v in (inv[0, 0], inv[0, 1], inv[0, 2], inv[1, 0], inv[1, 1], inv[1, 2])]
        fix = lambda v: int(np.floor(v * 65536.0 + 0.5))
        out[i] = (fix(ca), fix(cb), fix(cc + ca * 0.5 + cb * 0.5), fix(cd), fix(ce), fix(cf + cd * 0.5 + ce * 0.5))
    if np.abs(out).max() >= 2 ** 31:
        raise ValueError("affine transform outside PIL's 16.16 fixed-point range")
    return torch.from_numpy(out.astype(np.int32))


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
    f0, f1 = [f.to(dev, non_blocking=True).contiguous() for f in frames]
    if f0.dtype != torch.uint8 or f0.dim() != 4 or f0.shape[3] != 3 or f1.shape != f0.shape:
        raise ValueError("frames must be two uint8 tensors of the same shape [B,Hs,Ws,3]")
    B, Hs, Ws = f0.shape[:3]
    coef = affine_fixed_coefficients(affinetrans).to(dev, non_blocking=True)
    col = hue = order = None
    if color is not None:
        t = lambda v: torch.as_tensor(v, dtype=torch.float32).reshape(B)
        col = torch.stack([t(color["brightness"]), t(color["saturation"]), t(color["contrast"])], 1).contiguous().to(dev)
        hue_f = torch.as_tensor(color["hue"], dtype=torch.float64).reshape(B)
        hue = torch.tensor([int(float(h) * 255) & 0xFF for h in hue_f], dtype=torch.int32).to(dev)  # uint8(hue * 255)
        if orders is None:
            orders = torch.tensor([OP_BRIGHTNESS, OP_SATURATION, OP_HUE, OP_CONTRAST]).repeat(B, 2, 1)
        order = torch.as_tensor(orders, dtype=torch.int32).reshape(B, 2, 4).contiguous().to(dev)
    with torch.cuda.device(dev):
        if out is None:
            images = [torch.empty((B, 3, H, W), dtype=torch.float32, device=dev) for _ in range(2)]
            masks = [torch.empty((B, 3, H, W), dtype=torch.float32, device=dev) for _ in range(2)]
        else:
            images, masks = out
        ws_bytes = L.hoc_augment_frame_pair_workspace_bytes(B)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        _lib.check(L.hoc_augment_frame_pair(_lib.ptr(f0), _lib.ptr(f1), B, Hs, Ws, _lib.ptr(coef), _lib.ptr(col),
                                            _lib.ptr(hue), _lib.ptr(order), H, W, _lib.ptr(images[0]),
                                            _lib.ptr(images[1]), _lib.ptr(masks[0]), _lib.ptr(masks[1]), _lib.ptr(ws),
                                            ws_bytes, _lib.stream_ptr()), "hoc_augment_frame_pair")
    return images, masks
