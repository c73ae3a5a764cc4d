"""usage: ncu_lines.py <rep> <kernel regex> <cubin> [top]
Executed warp instructions and stall samples of one kernel aggregated by SOURCE LINE: ncu's SASS page is aligned (by
instruction order) with nvdisasm -g of the same build's cubin."""
import collections
import csv
import glob
import os
import re
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "handobjectconsist_b200", "csrc")
rep, kre, cubin = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]; h = rows[hi]
kname = rows[hi - 1][1]
ia, isamp, isrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
end = his[1] - 1 if len(his) > 1 else len(rows)
data = [(r[isrc].strip(), int(r[ia]), int(r[isamp])) for r in rows[hi + 1:end] if len(r) > ia and r[ia].isdigit()]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# find the function section whose name matches
short = re.match(r"(?:void )?(\w+)", kname).group(1)
tmpl = re.search(r"<\(?\w*\)?(\d+)>", kname)
cur = None; line = None; fn_lines = []
active = False
for l in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        name = m.group(1)
        active = short in name and (tmpl is None or ("ILi%sE" % tmpl.group(1)) in name or ("ILb%sE" % tmpl.group(1)) in name)
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
    if m:
        fn_lines.append((line, m.group(1).strip()))
print(kname[:90]); print("ncu sass", len(data), "nvdisasm sass", len(fn_lines))
n = min(len(data), len(fn_lines))
agg = collections.defaultdict(lambda: [0, 0])
tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
for (src, cnt, smp), (ln, txt) in zip(data[:n], fn_lines[:n]):
    agg[ln][0] += cnt; agg[ln][1] += smp
print("total warp instr", tot, "samples", ts)
srcs = {}
for (ln, (c, s)) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if ln:
        for cand in glob.glob(os.path.join(CSRC, ln[0])):
            if cand not in srcs: srcs[cand] = open(cand).read().splitlines()
            if ln[1] - 1 < len(srcs[cand]): text = srcs[cand][ln[1] - 1].strip()[:90]
    print(f"{c:9d} {100*c/tot:5.1f}%  smp {100*s/max(ts,1):5.1f}%  {ln[0] if ln else '?'}:{ln[1] if ln else 0}  {text}")
