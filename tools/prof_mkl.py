"""Per-kernel-family event times of the mkl fit on cfg1's matrix (BASELINE cfg2)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from swne_b200 import NMFHandle, synth
from swne_b200._capi import PROF_NAMES
c = synth.CONFIGS[os.environ.get("CFG", "cfg1")]
m = int(os.environ.get("CELLS", c["m"]))
dA = synth.make_matrix_torch(c["n"], m, c["kt"], c["scale"], c["seed"])
h = NMFHandle(0); h.set_matrix(dA)
A = dA.cpu().numpy()
nzc = (A != 0).sum(1); nzg = (A != 0).sum(0)
print("cells: nnz/col mean %.0f max %d; genes: mean %.0f max %d" % (nzc.mean(), nzc.max(), nzg.mean(), nzg.max()))
W0, H0 = synth.random_init(c["n"], m, c["k"], 0)
warm = h.fit(c["k"], W0, H0, loss="mkl", max_iter=6, rel_tol=-1.0)
for imi in (1, 4):
    for prof in (False, True):
        r = h.fit(c["k"], warm.W, warm.H, loss="mkl", max_iter=20, rel_tol=-1.0, inner_max_iter=imi, trace=100, profile=prof)
        print("imi", imi, "profile", prof, "us/iter %.1f" % (1e3 * r.stats["loop_ms"] / 20),
              {n: round(1e3 * v / 20, 1) for n, v in zip(PROF_NAMES, r.stats["prof_ms"]) if v > 0},
              "launches", r.stats["gpu_launches"], flush=True)
