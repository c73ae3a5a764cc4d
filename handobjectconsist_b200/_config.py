"""Run-time switches of the python layer."""

# Run the two independent renders / the two warp directions of a frame pair on two CUDA streams so that their
# kernels overlap (eager mode and inside captured graphs).  bench.py turns it off while it captures the
# instrumented graph it uses to time kernels one at a time.
overlap_streams = True

# warpbranch.forward takes the fused frame-pair path (consist.py) when the configuration allows it; False forces the
# operator-by-operator path (get_opticalflow + pair_consist), e.g. to compare the two in tests.
fused_pair = True


_SIDE_STREAMS = {}


def side_stream(device):
    """The side stream that accompanies the CURRENT stream on `device` (one per launching stream, so that
    independent chains never meet on it).  Shared by the flow renders and pair_consist: inside a captured step the
    stream forked for the second render is the one pair_consist reuses."""
    import torch

    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]

# The frame-pair path rasterises only the rows of the square raster that can matter for the (cropped) frame: the crop,
# the reach of the occlusion check below it and -- when the geometry gradient is wanted -- the rows of the meshes
# (hoc_pair_front computes the window on the device; include/hoc_b200.h).  False = the reference's full square.
raster_window = True
