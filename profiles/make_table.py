"""Markdown roofline table of DESIGN.md section 4 from a bench line and the ncu traffic summary.

    python profiles/make_table.py profiles/bench_r1_1gpu.json profiles/raster_backward_traffic.json
"""
import json
import sys

d = json.load(open(sys.argv[1]))
traffic = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else {}
peak = d["roofline"]["peak"]
names = {"raster_zbuf": "hoc_raster_zbuf_kernel", "raster_resolve": "hoc_raster_resolve_kernel",
         "raster_bwd_pixel": "hoc_raster_bwd_scan_kernel", "raster_bwd_pixel_k4": "hoc_raster_bwd_cover_kernel",
         "raster_bwd_cover": "hoc_raster_bwd_cover_kernel", "raster_bwd_line": "hoc_raster_bwd_line_kernel",
         "warp_photo_fwd": "hoc_warp_photo_forward_kernel", "warp_photo_bwd": "hoc_warp_photo_backward_kernel",
         "flow_finalize": "hoc_flow_finalize_kernel", "flow_finalize_bwd": "hoc_flow_finalize_backward_kernel",
         "mesh_gather": "hoc_mesh_gather_kernel", "mesh_scatter": "hoc_mesh_scatter_kernel",
         "cat_meshes": "hoc_cat_meshes_kernel", "flow_vertices": "hoc_flow_vertices_kernel",
         "flow_vertices_bwd": "hoc_flow_vertices_backward_kernel", "pair_loss": "hoc_pair_loss_kernel"}
print(f"| Kernel (launches/step) | algorithmic MB | in-graph avg (us) | achieved GB/s | frac of {peak:.1f} | ncu DRAM traffic MB |")
print("|---|---|---|---|---|---|")
for k in d["kernels"]:
    ab = k["algorithmic_bytes_per_launch"]
    t = traffic.get(names.get(k["kernel"], ""), None)
    print(f"| {k['kernel']} ({k['launches_per_step']:.0f}) | {ab / 1e6:.1f} | {k['avg_ms'] * 1e3:.1f} | "
          f"{k['achieved_gbs']:.0f} | {k['achieved_gbs'] / peak:.2f} | {'' if t is None else f'{t / 1e6:.1f}'} |"
          if ab else f"| {k['kernel']} ({k['launches_per_step']:.0f}) | | {k['avg_ms'] * 1e3:.1f} | | | {'' if t is None else f'{t / 1e6:.1f}'} |")
r = d["roofline_raster_backward"]
print(f"| **raster backward, {len(r['kernels'])} launches together** | {r['algorithmic_bytes_per_render'] / 1e6:.1f} | "
      f"{r['ms_per_render'] * 1e3:.1f} | {r['achieved']:.0f} | **{r['frac']:.3f}** | {'' if not r['traffic'] else f'{r['traffic'] / 1e6:.1f}'} |")
