"""Hot SASS lines of one kernel from `ncu -i rep --page source --csv --kernel-name K` (stdin or file)."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
data = [r for r in rows[1:] if r[hdr.index("Instructions Executed")].isdigit()]
isrc, iinst, ithr, isamp = (hdr.index(k) for k in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
tot = sum(int(r[iinst]) for r in data)
tots = sum(int(r[isamp]) for r in data)
print("SASS instr", len(data), "warp instr", tot, "samples", tots)
cum = 0
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
for n, r in enumerate(data):
    c = int(r[iinst])
    cum += c
    if c > tot * thr or int(r[isamp]) > tots * 0.01:
        print(n, r[isrc].strip()[:64].ljust(64), c, round(int(r[ithr]) / max(c, 1), 1), r[isamp], round(cum / tot, 3))
