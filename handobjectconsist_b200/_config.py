"""Run-time switches of the python layer."""

# Run the two independent renders / the two warp directions of a frame pair on two CUDA streams so that their
# kernels overlap (eager mode and inside captured graphs).  bench.py turns it off while it captures the
# instrumented graph it uses to time kernels one at a time.
overlap_streams = True
