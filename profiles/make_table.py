"""Markdown roofline table of DESIGN.md section 4 from a bench line, the ncu traffic summary and the CUPTI timeline.

    python profiles/make_table.py profiles/bench_r2_1gpu.json profiles/kernel_traffic_r2.json profiles/timeline_r2.txt
"""
import json
import re
import sys

d = json.load(open(sys.argv[1]))
traffic = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else {}
cupti = {}
if len(sys.argv) > 3:
    for line in open(sys.argv[3]):
        m = re.match(r"\s*([0-9.]+)\s+([0-9.]+)\s+(?:void )?(hoc_\w+)", line)  # one replay: start, duration, name
        if m:
            cupti.setdefault(m.group(3), float(m.group(2)))
    for line in open(sys.argv[3]):  # mean over the replays (preferred): mean, min, count, name
        m = re.match(r"\s*([0-9.]+)\s+([0-9.]+)\s+x[0-9.]+\s+(?:void )?(hoc_\w+)", line)
        if m:
            cupti[m.group(3)] = float(m.group(1))
peak = d["roofline"]["peak"]
names = {"raster_zbuf": "hoc_raster_zbuf_kernel", "raster_resolve": "hoc_raster_resolve4_kernel",
         "raster_bwd_pixel": "hoc_raster_bwd_scan_pair_kernel",
         "raster_bwd_cover": "hoc_raster_bwd_cover_kernel", "raster_bwd_line": "hoc_raster_bwd_line_kernel",
         "warp_photo_bwd": "hoc_warp_photo_pair_backward_kernel", "flow_finalize": "hoc_flow_finalize_warp_kernel",
         "mesh_scatter": "hoc_mesh_scatter_kernel", "pair_front": "hoc_pair_front_kernel",
         "pair_back": "hoc_pair_back_kernel", "pair_loss": "hoc_pair_loss_mean_kernel"}
print(f"| Kernel (one launch per step each) | algorithmic MB | in-graph, event nodes (us) | CUPTI, plain replay, mean of 40 (us) | "
      f"achieved GB/s (CUPTI) | frac of {peak:.1f} | ncu DRAM traffic MB |")
print("|---|---|---|---|---|---|---|")
tot_ev = tot_cu = 0.0
for k in d["kernels"]:
    ab = k["algorithmic_bytes_per_launch"]
    kn = names.get(k["kernel"], k["kernel"])
    t = traffic.get(kn)
    cu = cupti.get(kn)
    tot_ev += k["avg_ms"] * 1e3 * k["launches_per_step"]
    tot_cu += cu or 0.0
    gbs = (ab / (cu * 1e-6) / 1e9) if (ab and cu) else None
    print(f"| `{kn}` | {'' if not ab else f'{ab / 1e6:.1f}'} | {k['avg_ms'] * 1e3:.1f} | {'' if cu is None else f'{cu:.1f}'} | "
          f"{'' if gbs is None else f'{gbs:.0f}'} | {'' if gbs is None else f'{gbs / peak:.2f}'} | "
          f"{'' if t is None else f'{t / 1e6:.1f}'} |")
r = d["roofline"]
bw = [cupti.get(n) for n in ("hoc_raster_bwd_scan_pair_kernel", "hoc_raster_bwd_line_kernel")]
cu_b = sum(bw) if all(bw) else None
ab = r["algorithmic_bytes_per_launch"]
print(f"| **rasterizer backward: scan + line** | {ab / 1e6:.1f} | {r['avg_launch_ms'] * 1e3:.1f} | "
      f"{'' if cu_b is None else f'{cu_b:.1f}'} | {'' if cu_b is None else f'{ab / (cu_b * 1e-6) / 1e9:.0f}'} | "
      f"**{r['frac']:.3f}** (events) / **{'' if cu_b is None else f'{ab / (cu_b * 1e-6) / 1e9 / peak:.3f}'}** (CUPTI) | "
      f"{'' if not r['traffic'] else f'{r['traffic'] / 1e6:.1f}'} |")
print(f"| sum over the step | | {tot_ev:.0f} | {tot_cu:.0f} | | | |")
