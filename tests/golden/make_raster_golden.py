"""Generates tests/golden/raster_*.npz: outputs of the C restatement of the rasterizer (oracle/) on three
small seeded scenes, fp32 maps + fp32 and fp64 gradients.  The reference holds no golden vectors for this
path (its arithmetic lives in the absent `neural_renderer` wheel -- parity unpinned); these
self-made pins freeze the oracle so that later edits to it cannot silently move the target.

    python tests/golden/make_raster_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402
from helpers import onmr  # noqa: E402

SCENES = {"hand64": dict(B=1, S=64, seed=0, with_object=False), "handobj48": dict(B=2, S=48, seed=1, with_object=True),
          "handobj40": dict(B=1, S=40, seed=4, with_object=True)}


def grads_in(shape_rgb, shape_plane, seed):
    rng = np.random.default_rng(seed)
    return (rng.normal(size=shape_rgb).astype(np.float32), rng.normal(size=shape_plane).astype(np.float32),
            rng.normal(size=shape_plane).astype(np.float32))


def main():
    for name, cfg in SCENES.items():
        faces, tex, _ = helpers.scene_faces(cfg["B"], cfg["S"], seed=cfg["seed"], with_object=cfg["with_object"])
        ora = onmr.rasterize_forward(faces, tex, cfg["S"], 0.1, 100.0, 1e-3, (0, 0, 0), True, True, True)
        g_rgb, g_alpha, g_depth = grads_in(ora["rgb_map"].shape, ora["alpha_map"].shape, 7)
        gf32, gt32 = onmr.rasterize_backward(ora, g_rgb, g_alpha, g_depth)
        gf64, gt64 = onmr.rasterize_backward(ora, g_rgb, g_alpha, g_depth, dtype=np.float64)
        # sparse storage: only the faces that received any gradient
        hit = np.nonzero(np.abs(gf64).reshape(gf64.shape[0], gf64.shape[1], -1).sum(-1) +
                         np.abs(gt64).reshape(gt64.shape[0], gt64.shape[1], -1).sum(-1))
        np.savez_compressed(
            os.path.join(HERE, f"raster_{name}.npz"), cfg=np.array([cfg["B"], cfg["S"], cfg["seed"], int(cfg["with_object"])]),
            face_index_map=ora["face_index_map"].astype(np.int16), depth_map=ora["depth_map"],
            weight_map=ora["weight_map"].astype(np.float32), rgb_map=ora["rgb_map"],
            hit_b=hit[0].astype(np.int16), hit_f=hit[1].astype(np.int16),
            gf32=gf32[hit], gt32=gt32[hit], gf64=gf64[hit], gt64=gt64[hit])
        print(name, "covered", float((ora["face_index_map"] >= 0).mean()), "faces with grad", len(hit[0]))


if __name__ == "__main__":
    main()
