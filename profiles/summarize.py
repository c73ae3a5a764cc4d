"""Turns the raw ncu outputs under gpurun_out/ (scratch) into the tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv profiles/launches_r1.md
    python profiles/summarize.py full gpurun_out/prof_r1_full.ncu-rep profiles/ncu_full_r1.md profiles/raster_backward_traffic.json
"""
import collections
import csv
import json
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1000.0 if r["Metric Unit"] == "ns" else (v * 1000.0 if r["Metric Unit"] == "ms" else v)
        a = agg.setdefault(r["Kernel Name"].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    own = sum(a[1] for k, a in agg.items() if "hoc_" in k)
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over "
                f"`python bench.py --eager-only --steps 2 --warmup 3` (the eager arm of the frame-pair path: coverage probe, "
                f"warm-up and timed steps issued from Python; the ATen kernels are bench.py's own cloning of its inputs "
                f"into fresh leaf tensors and the one-off measured-coverage reduction, not part of a captured step).  "
                f"Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n"
                f"{len(rows)} launches, {tot:.0f} us of kernel time, of which this library's kernels: {own:.0f} us "
                f"({own / tot * 100:.1f} %).\n\n| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / tot * 100:.1f} % | {a[1] / a[0]:.1f} |\n")


WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instr"),
        ("launch__registers_per_thread", "registers"), ("smsp__inst_executed.sum", "warp instructions"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__cycles_active.avg", "sm cycles active (avg)"), ("sm__cycles_elapsed.max", "sm cycles elapsed (max)"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait")]


def _num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def full(rep, dst, traffic_json=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    per = collections.OrderedDict()
    for r in rows[2:]:
        per.setdefault(r[hdr.index("Kernel Name")].split("(")[0], []).append(r)
    traffic = {}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({rep})\n\n`ncu --set full --clock-control none --import-source on -k regex:hoc_` "
                f"over the eager arm of bench.py (16 frame pairs, 256x256: both renders of a pair stacked, 32 samples per rasterizer launch).  Values are the mean over the captured launches of "
                f"each kernel.\n\n")
        for k, rs in per.items():
            f.write(f"## `{k}`  ({len(rs)} launches, grid {rs[0][hdr.index('Grid Size')]}, block {rs[0][hdr.index('Block Size')]})\n\n"
                    f"| metric | value | unit |\n|---|---|---|\n")
            vals = {}
            for m, label in WANT:
                if m not in hdr:
                    continue
                i = hdr.index(m)
                xs = [_num(r[i]) for r in rs if _num(r[i]) is not None]
                if not xs:
                    continue
                vals[m] = (sum(xs) / len(xs), units[i])
                f.write(f"| {label} (`{m}`) | {vals[m][0]:.3f} | {units[i]} |\n")
            f.write("\n")
            if "dram__bytes_read.sum" in vals and "dram__bytes_write.sum" in vals:
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd = vals["dram__bytes_read.sum"][0] * scale.get(vals["dram__bytes_read.sum"][1], 1)
                wr = vals["dram__bytes_write.sum"][0] * scale.get(vals["dram__bytes_write.sum"][1], 1)
                traffic[k.replace("void ", "").split("<")[0]] = rd + wr
    if traffic_json:
        with open(traffic_json, "w") as f:
            json.dump(traffic, f, indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
