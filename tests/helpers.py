"""Shared test helpers: scenes in rasterizer space, oracle wrappers, comparison utilities."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import nmr as onmr  # noqa: E402  (checker)
from oracle import nrfuncs as onr  # noqa: E402
from handobjectconsist_b200 import synth  # noqa: E402


def scene_faces(batch, size, seed=0, with_object=True, fill_back=True, frame=1):
    """NDC faces [B,F',3,3] + flow textures [B,F',2,2,2,3] (numpy float32) of a synthetic scene,
    built with the ORACLE's helpers (projection, vertices_to_faces, fill_back, vertex textures)."""
    sc = synth.make_scene(batch, size, size, seed=seed, with_object=with_object)
    v1, v2 = (sc["verts1"], sc["verts2"]) if frame == 1 else (sc["verts2"], sc["verts1"])
    K = sc["K"]
    loc1 = onr.batch_proj2d(v1, K)
    loc2 = onr.batch_proj2d(v2, K)
    displ = loc2 - loc1
    cols = torch.cat([displ, torch.ones_like(displ[:, :, :1])], -1)
    tex = onr.batch_vertex_textures(sc["faces"], cols)
    faces = sc["faces"]
    if fill_back:
        faces, tex = onr.fill_back(faces, tex)
    R = torch.eye(3)[None]
    t = torch.zeros(1, 1, 3)
    dist = torch.zeros(1, 5)
    ndc = onr.projection(v1, K, R, t, dist, float(size))
    f = onr.vertices_to_faces(ndc, faces)
    return np.ascontiguousarray(f.numpy(), dtype=np.float32), np.ascontiguousarray(tex.numpy(), dtype=np.float32), sc


def rel_err(a, b):
    """max |a-b| / max(|b|, floor) with the floor at 1e-3 of the largest reference magnitude."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    return float((np.abs(a - b) / np.maximum(np.abs(b), 1e-3 * scale)).max())


HAND_VERTS, HAND_FACES = 778, 1552


def pair_step(sc, S, crop, dev, detach_renders, use_backward, return_visuals=True, loss_only=False):
    """The fused frame-pair path (handobjectconsist_b200.consist) on a synth scene: hand / object split like
    warpbranch.forward receives them.  Returns (mean loss, dict like warpbranch.consist_step's, leaf vertices)."""
    from handobjectconsist_b200 import consist
    from handobjectconsist_b200.neurender.renderer import Renderer
    g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}
    r = Renderer(image_size=S, R=torch.eye(3, device=dev)[None], t=torch.zeros(1, 3, device=dev),
                 K=torch.ones(1, 3, 3, device=dev), orig_size=S, anti_aliasing=False, fill_back=True, near=0.1,
                 no_light=True)
    v1 = g["verts1"].clone().requires_grad_(True)
    hv = HAND_VERTS
    hand_faces = g["faces"][0, :HAND_FACES]
    obj_faces = g["faces"][:, HAND_FACES:] - hv
    (loss, mean), flows, masks, warps, diffs = consist.pair_consist_step(
        v1[:, :hv], v1[:, hv:], g["verts2"][:, :hv], g["verts2"][:, hv:], hand_faces, obj_faces, g["K"], g["K"],
        g["image_ref"], g["image"], g["jitter_mask_ref"], g["jitter_mask"], r, crop,
        hand_ignore_faces=sc["hand_ignore_faces"], detach_renders=detach_renders, use_backward=use_backward,
        return_visuals=return_visuals, loss_only=loss_only)
    return mean, dict(flows=flows, loss=loss, masks=masks, warps=warps, diffs=diffs), v1
