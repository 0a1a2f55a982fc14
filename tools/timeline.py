import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SWNE_TC_TIMELINE"] = os.path.join(ROOT, "gpurun_out", "timeline.txt")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
import numpy as np, torch
from swne_b200 import NMFHandle, synth
c = synth.CONFIGS["cfg3"]
dA = synth.make_matrix_torch(c["n"], c["m"], c["kt"], c["scale"], c["seed"])
h = NMFHandle(0); h.set_matrix(dA)
W0, H0 = synth.random_init(c["n"], c["m"], c["k"], 0)
imi = int(sys.argv[1]) if len(sys.argv) > 1 else 50
r = h.fit(c["k"], W0, H0, max_iter=8, rel_tol=-1.0, engine="tc", inner_max_iter=imi)
rows = [list(map(int, l.split()[2:])) for l in open(os.environ["SWNE_TC_TIMELINE"])]
a = np.array(rows, dtype=float) / 1000.0   # us
names = {0: "start", 1: "prod_p1_end", 2: "mma_p1_end", 3: "mma_p3_end", 32: "drain0", 33: "drain1", 37: "drain5"}
for s, nm in names.items():
    v = a[:, s]; v = v[v >= 0]
    print("%-12s min %.1f med %.1f max %.1f" % (nm, v.min(), np.median(v), v.max()))
for c_ in (0, 73, 147):
    print("cta", c_, "job pick", a[c_, 8:16].round(0), "job done", a[c_, 16:24].round(0), "ps0 tile done", a[c_, 24:32].round(0))
wpath = os.environ["SWNE_TC_TIMELINE"] + ".wside"
if os.path.exists(wpath):
    w = np.array([list(map(int, l.split()[2:])) for l in open(wpath)], dtype=float) / 1000.0
    print("W-side kernel marks (us): start, prologue done, init pushes done, sweeps done, gram atomics done, [last CTA:] ticket, finalize, split, end")
    for c_ in range(w.shape[0]):
        print("  cta %2d" % c_, w[c_].round(1))
