"""Region-level view of one kernel's SASS from `ncu -i rep --page source --csv [--kernel-name K] > f.csv`:
warp instructions, active lanes, stall samples by reason per block of N SASS lines.
    python profiles/hot_sass.py f.csv [block=40]"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
col = {k: hdr.index(k) for k in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples")}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[1:] if r[col["Instructions Executed"]].isdigit()]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tot = sum(int(r[col["Instructions Executed"]]) for r in data)
tots = sum(int(r[col["# Samples"]]) for r in data)
print(f"{len(data)} SASS lines, {tot} warp instructions, {tots} samples")
for s in range(0, len(data), step):
    seg = data[s:s + step]
    c = sum(int(r[col["Instructions Executed"]]) for r in seg)
    t = sum(int(r[col["Thread Instructions Executed"]]) for r in seg)
    sm = sum(int(r[col["# Samples"]]) for r in seg)
    if c == 0:
        continue
    st = {k: sum(int(r[hdr.index(k)] or 0) for r in seg) for k in stalls}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    ops = [(r[col["Source"]].split()[0] if not r[col["Source"]].strip().startswith("@") else r[col["Source"]].split()[1]) for r in seg]
    key = [o for o in ops if o.startswith(("LDG", "LDS", "STS", "STG", "BAR", "RED", "ATOM", "MUFU", "EXIT"))][:6]
    print(f"{s:5d} inst {c / tot * 100:5.1f}%  lanes {t / c:5.1f}  samples {sm / max(tots, 1) * 100:5.1f}%  "
          f"{' '.join(f'{k[6:]}={v}' for k, v in top if v)}  | {' '.join(key)}")
