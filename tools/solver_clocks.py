"""Solver clock ring of the fused pass (SWNE_TC_CLOCKS=<file>): per-phase SM-cycle breakdown of the block updates.

    SWNE_TC_CLOCKS=gpurun_out/clocks.bin python tools/solver_clocks.py [inner_max_iter]
"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
path = os.environ.setdefault("SWNE_TC_CLOCKS", os.path.join(ROOT, "gpurun_out", "clocks.bin"))
os.makedirs(os.path.dirname(path), exist_ok=True)
import numpy as np
from swne_b200 import NMFHandle, synth

c = synth.CONFIGS["cfg3"]
m = int(os.environ.get("CELLS", c["m"]))
dA = synth.make_matrix_torch(c["n"], m, c["kt"], c["scale"], c["seed"])
h = NMFHandle(0)
h.set_matrix(dA)
W0, H0 = synth.random_init(c["n"], m, c["k"], 0)
imi = int(sys.argv[1]) if len(sys.argv) > 1 else 50
r = h.fit(c["k"], W0, H0, max_iter=8, rel_tol=-1.0, engine="tc", inner_max_iter=imi)
print("ms/iter %.3f" % (r.stats["loop_ms"] / 8), "epochs", r.average_epochs)
raw = open(path, "rb").read()
hdr = np.frombuffer(raw[:16], dtype=np.int32)
grid, ng, nrec, w = [int(x) for x in hdr]
a = np.frombuffer(raw[16:], dtype=np.int64).reshape(grid, ng, nrec, w)
blk = a[a[..., 7] == 1]
tile = a[a[..., 7] == 2]
names = ["wait_mma", "tmem_ld", "steps16", "push(split+st+bar+issue)", "-", "-"]
d = np.diff(blk[:, :7], axis=1)
print("block updates recorded: %d" % len(blk))
for i, nm in enumerate(names):
    print("  %-16s mean %7.0f  p50 %7.0f  p90 %7.0f cycles" % (nm, d[:, i].mean(), np.median(d[:, i]), np.percentile(d[:, i], 90)))
print("  %-16s mean %7.0f cycles per block update" % ("total", (blk[:, 6] - blk[:, 0]).mean()))
# gap between consecutive block updates of the same group (includes bar.or / loop overhead)
for g in range(ng):
    x = a[0, g]
    x = x[x[:, 7] == 1]
    if len(x) > 2:
        print("  cta0 group %d: mean period %.0f cycles over %d updates" % (g, np.diff(x[:, 0]).mean(), len(x)))
if len(tile):
    t = tile
    print("tiles recorded: %d" % len(t))
    print("  wait C-full %7.0f  init(mu) %7.0f  sweeps %7.0f  epilogue %7.0f cycles (means); sweeps per tile %.1f" % (
        (t[:, 1] - t[:, 0]).mean(), (t[:, 2] - t[:, 1]).mean(), (t[:, 3] - t[:, 2]).mean(), (t[:, 4] - t[:, 3]).mean(), t[:, 5].mean()))
    for cta in (0, grid // 2):
        for g in range(ng):
            x = a[cta, g]
            x = x[x[:, 7] == 2]
            base = a[cta, :, :, 0][a[cta, :, :, 7] == 2].min()
            print("  cta %d group %d tiles:" % (cta, g), [(int(r[6]), int(r[0] - base), int(r[1] - base), int(r[3] - base), int(r[4] - base), int(r[5])) for r in x])
# block-update duration against time (cta 0): is the solver slower while the two streams run?
x = a[0].reshape(-1, w)
x = x[x[:, 7] == 1]
t0 = a[0][..., 0][a[0][..., 7] == 2].min()
edges = np.arange(0, 1_200_000, 50_000)
print("cta 0: block updates by start time (kcycles): count, mean total, mean steps16, mean push, mean wait")
for lo, hi in zip(edges[:-1], edges[1:]):
    sel = x[(x[:, 0] - t0 >= lo) & (x[:, 0] - t0 < hi)]
    if len(sel):
        print("  %5d-%5d  n=%4d  total %6.0f  steps %6.0f  push %6.0f  wait %6.0f  ld %5.0f" % (
            lo // 1000, hi // 1000, len(sel), (sel[:, 6] - sel[:, 0]).mean(), (sel[:, 3] - sel[:, 2]).mean(),
            (sel[:, 6] - sel[:, 3]).mean(), (sel[:, 1] - sel[:, 0]).mean(), (sel[:, 2] - sel[:, 1]).mean()))
