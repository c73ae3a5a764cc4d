"""ctypes binding of ``libhoc_b200.so`` (C ABI declared in ``include/hoc_b200.h``).

The library is the product: there is no CPU or PyTorch fallback.  If it has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C handobjectconsist_b200/csrc``)
every operator of this package raises ``HocLibraryError``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HOC_B200_LIB: another build of the same library (A/B sweeps of compile-time knobs); default: the in-tree build
LIB_PATH = os.environ.get("HOC_B200_LIB") or os.path.join(_HERE, "libhoc_b200.so")

HOC_LAYOUT_RAW = 0
HOC_LAYOUT_IMAGE = 1
HOC_LAYOUT_KEYS_CLEARED = 0x100
HOC_LAYOUT_TEX_VERTEX = 0x200
HOC_LAYOUT_SPARSE_SAVED = 0x400
HOC_TEX_GRAD_CUBE = 0
HOC_TEX_GRAD_VERTEX = 1
HOC_TUNE_LINE_THREADS = 1
HOC_TUNE_LINE_SEGMENT = 2
HOC_TUNE_DETERMINISTIC = 3
HOC_TUNE_PDL = 5
HOC_TUNE_COVER_CTAS = 6
HOC_TUNE_LINE_LINES = 7
HOC_TUNE_LINE_FOLD = 8
HOC_TUNE_TEX_IN_LINE = 9
HOC_BWD_WORKSPACE_ZEROED = 1

_c_float_p = ctypes.c_void_p  # device pointers travel as integers
_vp = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float
_sz = ctypes.c_size_t


class HocLibraryError(RuntimeError):
    pass


# name -> (restype, argtypes); mirrors include/hoc_b200.h one to one
SIGNATURES = {
    "hoc_abi_version": (_i, []),
    "hoc_last_error": (ctypes.c_char_p, []),
    "hoc_launch_count": (ctypes.c_ulonglong, [_i]),
    "hoc_timer_begin": (_i, [ctypes.c_ulonglong]),
    "hoc_timer_end": (_i, [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int), _i]),
    "hoc_timer_pause": (_i, []),
    "hoc_timer_peek": (_i, [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int), _i]),
    "hoc_raster_forward_workspace_bytes": (_sz, [_i, _i, _i]),
    "hoc_raster_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _f, _f, ctypes.POINTER(ctypes.c_float), _vp, _i,
                                _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hoc_raster_forward_ex": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _f, _f, ctypes.POINTER(ctypes.c_float), _vp, _i, _vp,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hoc_raster_backward_workspace_bytes": (_sz, [_i, _i, _i]),
    "hoc_raster_backward_workspace_bytes_ex": (_sz, [_i, _i, _i, _i, _i]),
    "hoc_mesh_scatter_workspace_bytes": (_sz, [_i, _i]),
    "hoc_mesh_scatter_ws": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _sz, _vp]),
    "hoc_set_tuning": (_i, [_i, _i]),
    "hoc_unpack_u8": (_i, [_vp, _vp, ctypes.c_longlong, _f, _f, _vp]),
    "hoc_flow_finalize_backward_pair": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "hoc_mesh_gather_clear": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "hoc_pair_loss": (_i, [_vp, _vp, _i, _vp, _vp]),
    "hoc_cat_meshes": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "hoc_raster_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _i, _i, _i,
                                 _vp, _vp, _vp, _sz, _vp]),
    "hoc_raster_backward_ex": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _i, _i, _i,
                                    _i, _i, _vp, _sz, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hoc_raster_backward_zero_bytes": (_sz, [_i, _i, _i]),
    "hoc_flow_finalize_warp": (_i, [_vp] * 10 + [_i] * 4 + [_vp, _i, _f, _f] + [_vp] * 8),
    "hoc_flow_finalize_warp_ex": (_i, [_vp] * 10 + [_i] * 4 + [_vp, _i, _f, _f] + [_vp] * 7 + [_i, _vp]),
    "hoc_pair_backward_zero_bytes": (_sz, [_i, _i, _i]),
    "hoc_pair_loss_mean": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "hoc_pair_backward_raster": (_i, [_vp] * 10 + [_i] * 5 + [_vp] * 6 + [_i] * 3 + [_f] * 3 + [_i, _i, _vp, _sz, _vp, _vp,
                                                                                                 _vp, _vp, _sz, _vp]),
    "hoc_warp_photo_forward_pair": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp,
                                         _vp]),
    "hoc_warp_photo_backward_pair": (_i, [_vp] * 10 + [_i] * 5 + [_vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "hoc_pair_front": (_i, [_vp] * 5 + [_i, _vp] + [_vp, _i] * 5 + [_f] + [_i] * 6
                       + [_vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _i, _i, _i, _vp]),
    "hoc_pair_back": (_i, [_vp] * 4 + [_vp, _i] * 5 + [_f] + [_i] * 3 + [_vp, _vp] + [_i] * 4 + [_vp, _vp, _vp]),
    "hoc_warp_photo_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                    _vp]),
    "hoc_warp_photo_forward_acc": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                    _vp]),
    "hoc_warp_photo_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp]),
    "hoc_warp": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp]),
    "hoc_warp_backward": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp]),
    "hoc_occlusion_mask": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "hoc_mesh_gather": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "hoc_mesh_scatter": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "hoc_flow_finalize": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _f, _vp, _vp, _vp, _vp,
                               _vp]),
    "hoc_flow_finalize_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "hoc_flow_vertices": (_i, [_vp, _vp, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _f, _i, _i, _vp, _vp, _vp, _vp,
                               _vp]),
    "hoc_flow_vertices_backward": (_i, [_vp, _vp, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _f, _i, _i, _vp, _vp,
                                        _vp, _vp, _vp, _vp, _vp]),
}

KERNEL_IDS = {"raster_zbuf": 0, "raster_resolve": 1, "grad_extent": 2, "raster_backward": 3, "warp_photo_fwd": 4,
              "warp_photo_bwd": 5, "warp": 6, "warp_bwd": 7, "occlusion": 8, "mesh_gather": 9, "mesh_scatter": 10,
              "flow_finalize": 11, "flow_finalize_bwd": 12, "raster_bwd_pixel": 13, "raster_bwd_line": 14, "flow_vertices": 15, "flow_vertices_bwd": 16, "mano_fwd": 17,
              "mano_bwd": 18, "raster_bwd_pixel_k4": 19, "raster_bwd_cover": 20, "cat_meshes": 21, "pair_loss": 22, "unpack_u8": 23,
              "hand_head_fwd": 24, "hand_head_bwd": 25, "recover_points_fwd": 26, "recover_points_bwd": 27,
              "pair_front": 28, "pair_back": 29, "augment_stats": 30, "augment_frames": 31,
              "raster_bwd_group": 32}



class ManoModelStruct(ctypes.Structure):
    """``hoc_mano_model`` of include/hoc_b200.h."""
    _fields_ = [("v_template", _vp), ("shapedirs", _vp), ("posedirs", _vp), ("j_regressor", _vp), ("posedirs_t", _vp),
                ("j_template", _vp), ("j_shapedirs", _vp), ("weights", _vp),
                ("hands_components", _vp), ("hands_mean", _vp), ("num_verts", _i), ("ncomps", _i), ("use_pca", _i),
                ("center_idx", _i), ("tip_ids", _i * 5)]


SIGNATURES["hoc_mano_forward"] = (_i, [ctypes.POINTER(ManoModelStruct), _vp, _vp, _vp, _i, _vp, _vp, _vp])
SIGNATURES["hoc_mano_backward_workspace_bytes"] = (_sz, [_i])
SIGNATURES["hoc_mano_backward"] = (_i, [ctypes.POINTER(ManoModelStruct), _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp,
                                        _vp, _sz, _vp])

SIGNATURES["hoc_augment_frame_pair_workspace_bytes"] = (_sz, [_i])
SIGNATURES["hoc_augment_frame_pair"] = (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz,
                                             _vp])

_GH_CAM = [_vp, _i, _vp, _vp] + [_f] * 5  # camintr, camintr_batched, scale, trans, factors / off_z / input_res
SIGNATURES["hoc_hand_head_forward"] = (_i, [_vp] * 3 + [_i] * 4 + _GH_CAM + [_vp] * 7 + [_vp])
SIGNATURES["hoc_hand_head_backward"] = (_i, [_vp] * 3 + [_i] * 4 + _GH_CAM + [_vp] * 7 + [_vp] * 5 + [_vp])
SIGNATURES["hoc_recover_points_forward"] = (_i, [_vp] * 2 + [_i] * 2 + _GH_CAM + [_vp] * 4 + [_vp])
SIGNATURES["hoc_recover_points_backward"] = (_i, [_vp] * 2 + [_i] * 2 + _GH_CAM + [_vp] * 4 + [_vp] * 4 + [_vp])

_LIB = None


def lib():
    """Load the shared library once; raise loudly when it is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise HocLibraryError(
                f"{LIB_PATH} not found: the sm_100a kernels are not built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "There is no CPU / PyTorch fallback for this path.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
        # measurement switch: HOC_B200_TUNE="key=value,key=value" (hoc_set_tuning keys of include/hoc_b200.h)
        for item in filter(None, os.environ.get("HOC_B200_TUNE", "").split(",")):
            key, value = item.split("=")
            if L.hoc_set_tuning(int(key), int(value)) != 0:
                raise HocLibraryError(f"HOC_B200_TUNE: {item}: {L.hoc_last_error().decode()}")
    return _LIB


class deterministic:
    """Context manager / switch of the library's reproducible mode (``HOC_TUNE_DETERMINISTIC``, include/hoc_b200.h):
    gradient sums are accumulated in 128-bit fixed point with integer atomics, so that two runs -- or a captured
    graph and the eager path -- give the same bits.  Slower; meant for tests and debugging.  Process-wide."""

    def __init__(self, on=True):
        self.on = bool(on)

    def __enter__(self):
        check(lib().hoc_set_tuning(HOC_TUNE_DETERMINISTIC, int(self.on)), "hoc_set_tuning")
        return self

    def __exit__(self, *exc):
        check(lib().hoc_set_tuning(HOC_TUNE_DETERMINISTIC, 0), "hoc_set_tuning")
        return False


def check(code, what):
    if code != 0:
        msg = lib().hoc_last_error().decode("utf-8", "replace")
        raise HocLibraryError(f"{what} failed with code {code}: {msg}")


def ptr_pair(a, b):
    """C array of two device pointers (per-direction outputs of the frame-pair kernels); None -> NULL entries."""
    arr = (ctypes.c_void_p * 2)()
    arr[0] = None if a is None else a.data_ptr()
    arr[1] = None if b is None else b.data_ptr()
    return arr


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors, what="operator"):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise TypeError(f"{what} supports only cuda Tensors (got a {t.device} tensor); "
                            "this package has no CPU path")
