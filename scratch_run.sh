#!/bin/bash
# usage: scratch_run.sh tag  -- gpu tests + bench summary
tag=$1
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_$tag.log; cat gpurun_out/pytest_$tag.log
python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("value",round(d["value"]),"ms",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"]),"eager",round(d["eager"]["value"]))
for k in d["kernels"]: print(k["kernel"],k["launches_per_step"],round(k["avg_ms"]*1000,1))
print(d["roofline_raster_backward"]["ms_per_render"], d["roofline_raster_backward"]["frac"])
PY
