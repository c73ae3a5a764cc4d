"""CPU: the C-ABI library loads without a GPU and exports every symbol include/hoc_b200.h declares; the
ctypes table of the python binding names exactly the same entry points; argument errors come back as
codes + messages, never as exceptions or crashes.  (No kernel is launched here.)"""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers
from handobjectconsist_b200 import _lib

HEADER = os.path.join(helpers.ROOT, "include", "hoc_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hoc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 12
    L = _lib.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/hoc_b200.h but not exported by libhoc_b200.so"
    assert sorted(_lib.SIGNATURES) == names, "python binding table and header disagree"
    assert L.hoc_abi_version() == 1


def test_every_entry_point_cites_the_reference():
    src = open(HEADER).read()
    for tag in ("rasterize.py:202", "rasterize.py:269", "imgflowarp.py:80", "imgflowarp.py:31", "imgflowarp.py:118"):
        assert tag in src


def test_argument_errors_are_codes_not_crashes():
    L = _lib.lib()
    assert L.hoc_raster_forward_workspace_bytes(2, 10, 16) == 2 * 16 * 16 * 8
    # line spans (4 ints per line) + counters + per-face depth sums + list of (pixel, face) pairs
    assert L.hoc_raster_backward_workspace_bytes(2, 10, 16) >= (2 * 4 * 16 * 4 + 2 * 4 + 2 * 10 * 3 * 4 + 2 * 16 * 16 * 8)
    assert L.hoc_mano_backward_workspace_bytes(3) >= 3 * (192 + 135 + 10) * 4
    bg = (ctypes.c_float * 3)(0, 0, 0)
    # image size out of range / missing index map: rejected before anything touches the device
    code = L.hoc_raster_forward(None, None, 1, 0, 4096, 0, 0.1, 100.0, 1e-3, bg, None, 0, None, None, None, None, None,
                                None, None, 0, None)
    assert code == -1 and b"image_size" in L.hoc_last_error()
    code = L.hoc_raster_forward(None, None, 1, 0, 16, 0, 0.1, 100.0, 1e-3, bg, None, 0, None, None, None, None, None,
                                None, None, 0, None)
    assert code == -1 and b"face_index_map" in L.hoc_last_error()
    code = L.hoc_warp(None, None, 1, 3, 8, 8, 0.99999, 7, None, None, None)
    assert code == -1 and b"mode" in L.hoc_last_error()
    # the newer entry points: same contract (a code and a message, nothing launched)
    assert L.hoc_set_tuning(99, 1) == -1 and b"hoc_set_tuning" in L.hoc_last_error()
    assert L.hoc_set_tuning(1, 100) == -1            # threads per CTA must be a multiple of 32
    assert L.hoc_set_tuning(1, 128) == 0 and L.hoc_set_tuning(2, 16) == 0 and L.hoc_set_tuning(2, 0) == 0  # (0: by raster size)
    assert L.hoc_unpack_u8(None, None, 16, 255.0, 0.5, None) == -1 and b"hoc_unpack_u8" in L.hoc_last_error()
    assert L.hoc_unpack_u8(None, None, 0, 255.0, 0.5, None) == 0   # nothing to do
    assert L.hoc_cat_meshes(None, None, None, None, None, 0, None, 0, 778, 1502, 1552, 3000, None, None, None, None) == 0
    assert L.hoc_cat_meshes(None, None, None, None, None, 0, None, 2, 778, 1502, 1552, 3000, None, None, None, None) == -1
    assert L.hoc_pair_loss(None, None, 4, None, None) == -1 and b"hoc_pair_loss" in L.hoc_last_error()
    assert L.hoc_mesh_gather_clear(None, None, None, 1, 4, 2, 1, 7, None, None, None, 0, None) == -1
    assert b"tex_mode" in L.hoc_last_error()
    assert L.hoc_raster_forward(None, None, 1, 0, 16, 3, 0.1, 100.0, 1e-3, bg, None, 0x200, None, None, None, None, None,
                                None, None, 0, None) == -1      # vertex textures need texture_size 2
    assert b"texture_size 2" in L.hoc_last_error()
    # geometry head (geom_head.cu)
    cam = (None, 1, None, None, 1.0, 1.0, 0.4, 256.0, 256.0)
    assert L.hoc_hand_head_forward(None, None, None, 0, 778, 21, 9, *cam, *([None] * 8)) == 0       # empty batch
    assert L.hoc_hand_head_forward(None, None, None, 2, 778, 21, 9, *cam, *([None] * 8)) == -1
    assert b"camintr" in L.hoc_last_error()
    assert L.hoc_hand_head_forward(None, None, None, 0, 5000, 21, 9, *cam, *([None] * 8)) == -1
    assert b"vertices" in L.hoc_last_error()
    assert L.hoc_hand_head_forward(None, None, None, 0, 778, 21, 21, *cam, *([None] * 8)) == -1
    assert b"center_idx" in L.hoc_last_error()
    assert L.hoc_hand_head_backward(None, None, None, 0, 778, 40, 9, *cam, *([None] * 13)) == -1
    assert b"joints" in L.hoc_last_error()
    assert L.hoc_recover_points_forward(None, None, 0, 10, *cam, *([None] * 5)) == 0
    assert L.hoc_recover_points_forward(None, None, 1, -1, *cam, *([None] * 5)) == -1
    assert L.hoc_recover_points_backward(None, None, 3, 10, *cam, *([None] * 9)) == -1
    assert b"hoc_recover_points_backward" in L.hoc_last_error()


def test_pixel_centre_float_equals_reference_double_formula():
    """raster_math.h computes (2i+1-S)/S as a correctly rounded float quotient; the reference evaluates it in
    double and rounds -- identical for every S the ABI accepts."""
    for S in list(range(1, 300)) + [480, 512, 960, 1024, 2047, 2048]:
        i = np.arange(S)
        ref = ((2.0 * i + 1 - S) / S).astype(np.float32)
        ours = (2 * i + 1 - S).astype(np.float32) / np.float32(S)
        np.testing.assert_array_equal(ours, ref)


def test_cpu_tensors_are_rejected_loudly():
    import torch
    from handobjectconsist_b200.neurender.rasterize import Rasterize, rasterize_rgbad
    from handobjectconsist_b200.warping.imgflowarp import warp
    with pytest.raises(TypeError):
        Rasterize(8, 0.1, 100, 1e-3, (0, 0, 0), True, True, True)(torch.zeros(1, 1, 3, 3), torch.zeros(1, 1, 2, 2, 2, 3))
    with pytest.raises(TypeError):
        rasterize_rgbad(torch.zeros(1, 1, 3, 3), torch.zeros(1, 1, 2, 2, 2, 3))
    with pytest.raises(TypeError):
        warp(torch.zeros(1, 3, 4, 4), torch.zeros(1, 2, 4, 4))
    from handobjectconsist_b200.project import recover_3d_proj
    with pytest.raises(TypeError):
        recover_3d_proj(torch.zeros(1, 4, 3), torch.eye(3)[None], torch.zeros(1, 1, 1), torch.zeros(1, 1, 2))
