import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# Order of the GPU files under `-x`: the oracle-parity tests of every row of the hot path first, the
# BASELINE-size oracle comparisons next, self-consistency properties and baselines last -- so that a failure in a
# property test can never hide the parity tests.
_GPU_ORDER = ["test_gpu_raster", "test_gpu_warp", "test_gpu_pipeline", "test_gpu_mano", "test_gpu_geom",
              "test_gpu_renderer", "test_gpu_inputpipe", "test_gpu_warpreg", "test_gpu_parity_fullsize",
              "test_gpu_determinism", "test_gpu_ref_equiv", "test_gpu_fullsize"]


def _file_rank(item):
    name = os.path.splitext(os.path.basename(str(item.fspath)))[0]
    if name in _GPU_ORDER:
        return (1, _GPU_ORDER.index(name))
    return (0, 0) if not name.startswith("test_gpu") else (1, len(_GPU_ORDER))


@pytest.fixture
def det_mode():
    """The library's reproducible mode (HOC_TUNE_DETERMINISTIC): gradient sums are accumulated in fixed point with
    integer atomics, so two GPU runs (or a captured graph and the eager path) can be compared bit for bit."""
    from handobjectconsist_b200 import _lib

    with _lib.deterministic(True):
        yield


@pytest.fixture(autouse=True)
def _poison_cuda_allocator(request):
    """Before every GPU test: fill a few hundred MB of the caching allocator's free blocks with NaN patterns.  A kernel
    that reads a `torch.empty` buffer it was supposed to fill (e.g. rows outside the raster window) then sees NaNs
    instead of whatever correct values an earlier test left at that address -- such bugs fail every time, not once in
    a while on a fresh box."""
    if "gpu" in request.keywords:
        import torch

        if torch.cuda.is_available():
            junk = [torch.full((64, 1024, 1024), float("nan"), device="cuda") for _ in range(2)]
            del junk
    yield


def pytest_collection_modifyitems(config, items):
    import torch

    items.sort(key=_file_rank)  # stable: keeps the order inside a file
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
