"""Distribution of the line pass's work on bench.py's workload (reads the counters left in the backward workspace)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from handobjectconsist_b200 import consist, warpbranch
from handobjectconsist_b200.neurender.renderer import Renderer
from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
dev = torch.device("cuda", 0)
W = H = S = 256
renderer = Renderer(image_size=S, R=torch.eye(3, device=dev)[None], t=torch.zeros(1, 3, device=dev),
                    K=torch.ones(1, 3, 3, device=dev), orig_size=S, anti_aliasing=False, fill_back=True, near=0.1, no_light=True)
crit = PyramidCriterion("l1")
sc = bench._make_sets(1, 16, (W, H), dev)[0]
samples, results = bench._samples_from_scene(sc)
hv = results[0]["recov_handverts3d"].detach().requires_grad_(True)
ov = results[0]["recov_objverts3d"].detach().requires_grad_(True)
res = [{"recov_handverts3d": hv, "recov_objverts3d": ov}, results[1]]
st = {}
consist.STATS = st
loss, _ = warpbranch.forward(samples, res, sc["faces"][0, :1552].clone(), renderer, (W, H), crit,
                             hand_ignore_faces=sc["hand_ignore_faces"], gt_refs=True, first_only=True, use_backward=True, detach_renders=False)
loss.backward()
torch.cuda.synchronize()
ws, n, F, S, geom = st["backward_workspace"]
up = lambda v: (v + 255) & ~255
raw = ws.cpu().numpy()
off = 0
ext = raw[off:off + 4 * 4 * n * S].view(np.int32).reshape(n, 4, S); off = up(off + 4 * 4 * n * S)
cov = raw[off:off + 4 * n].view(np.int32); off = up(off + 4 * n)
lc = raw[off:off + 4 * 2 * n * S].view(np.int32).reshape(n, 2, S); off = up(off + 4 * 2 * n * S)
off = up(off + 4)
off = up(off + 4 * 3 * n * F)   # acc_d
off = up(off + 4 * 2 * n * S)   # line_list
off = up(off + 8 * n * S * S)   # cov_list
em = raw[off:off + 4 * 2 * n * S * 3 * S].view(np.uint32).reshape(n, 2, S, 3 * S)
print("n", n, "geom", geom, "cov per sample", cov[:geom].mean(), "max", cov.max())
lc = lc[:geom]
print("scans total", lc.sum(), "non-empty lines", (lc > 0).sum(), "of", lc.size, "max per line", lc.max(), "mean over non-empty", lc[lc > 0].mean())
row_lo = S - ext[:geom, 0]; row_hi = ext[:geom, 1] - 1; col_lo = S - ext[:geom, 2]; col_hi = ext[:geom, 3] - 1
span = np.stack([np.maximum(col_hi - col_lo + 1, 0), np.maximum(row_hi - row_lo + 1, 0)], 1)  # axis 0: column x uses col span
print("span mean over non-empty", span[lc > 0].mean(), "max", span.max())
work = lc * span
print("upper bound scans x span: total", work.sum(), "max line", work.max(), "p99", np.percentile(work[lc > 0], 99), "median", np.median(work[lc > 0]))
# exact chunk counts need the geometry; the upper bound n * span / 2 on average.  Histogram of scans per line:
print("hist scans/line", np.histogram(lc[lc > 0], bins=[1, 8, 16, 32, 64, 128, 256, 512, 1024])[0])
top = np.sort(work.ravel())[::-1][:10]
print("top-10 line work (scans x span)", top)
