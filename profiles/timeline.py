"""Kernel timeline of ONE replay of the captured consist step (torch.profiler / CUPTI), bench.py's workload.

    python profiles/timeline.py [pairs width height] > profiles/timeline_rN.txt      (default: configs[2], 16 256 256)

Columns: start (us, relative to the first kernel of the replay), duration (us), kernel name.  Unlike ncu's per-launch
numbers these are warm-cache and overlapped exactly as in production (two streams inside the graph)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from handobjectconsist_b200.graphed import GraphedConsistStep  # noqa: E402
from handobjectconsist_b200.neurender.renderer import Renderer  # noqa: E402
from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion  # noqa: E402

P, W, H = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (bench.PAIRS, bench.WIDTH, bench.HEIGHT)
S = max(W, H)
dev = torch.device("cuda:0")
renderer = Renderer(image_size=S, R=torch.eye(3, device=dev)[None], t=torch.zeros(1, 3, device=dev),
                    K=torch.ones(1, 3, 3, device=dev), orig_size=S, anti_aliasing=False, fill_back=True, near=0.1,
                    no_light=True)
sets = bench._make_sets(2, P, (W, H), dev)
hand_face = sets[0]["faces"][0, :1552].clone()
batches = [bench._samples_from_scene(sc) for sc in sets]
g = GraphedConsistStep(renderer, PyramidCriterion("l1"), (W, H), hand_face, *batches[0],
                       hand_ignore_faces=sets[0]["hand_ignore_faces"], gt_refs=True, first_only=True, use_backward=True,
                       detach_renders=False, warmup=2)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
REPLAYS = 40
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for i in range(REPLAYS):
        g.load(*batches[i % 2])
        g.replay()
    torch.cuda.synchronize()
ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
every = {}
for e in ev:
    if "Memcpy" not in e.name and "Memset" not in e.name:
        every.setdefault(e.name.split("(")[0][:60], []).append(e.time_range.end - e.time_range.start)
# last replay = everything after the last input copy of g.load()
cut = max(i for i, e in enumerate(ev) if "Memcpy" in e.name)
ev = ev[cut + 1:]
t0 = ev[0].time_range.start
span = max(e.time_range.end for e in ev) - t0
print(f"{P} frame pairs of {W}x{H} on a {S}x{S} raster: kernels in replay {len(ev)} span us {span:.3f}")
agg = {}
for e in ev:
    d = e.time_range.end - e.time_range.start
    print(f"{e.time_range.start - t0:8.1f} {d:7.1f}  {e.name[:64]}")
    a = agg.setdefault(e.name.split("(")[0][:60], [0, 0.0])
    a[0] += 1
    a[1] += d
print("\nsum of kernel durations by name (overlapped streams: sums exceed the span)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1]:8.1f} us  x{a[0]:<3d} {k}")

print(f"\nmean / min duration over {REPLAYS} replays (us)")
tot = 0.0
for k, v in sorted(every.items(), key=lambda kv: -sum(kv[1])):
    per = len(v) / REPLAYS
    tot += sum(v) / REPLAYS
    print(f"{sum(v) / len(v):8.2f} {min(v):8.2f}  x{per:<4.1f} {k}")
print(f"{tot:8.2f}           sum of the means")
