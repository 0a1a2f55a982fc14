"""Per-kernel-family event times of the mse fit on cfg1 (2000 x 3000, k=16)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from swne_b200 import NMFHandle, synth
from swne_b200._capi import PROF_NAMES
c = synth.CONFIGS["cfg1"]
dA = synth.make_matrix_torch(c["n"], c["m"], c["kt"], c["scale"], c["seed"])
h = NMFHandle(0); h.set_matrix(dA)
W0, H0 = synth.random_init(c["n"], c["m"], c["k"], 0)
warm = h.fit(c["k"], W0, H0, max_iter=6, rel_tol=-1.0)
for eng in ("tc", "simt"):
    for prof in (False, True):
        r = h.fit(c["k"], warm.W, warm.H, max_iter=100, rel_tol=-1.0, engine=eng, trace=1000, profile=prof, full_traces=-1)
        print(eng, "profile", prof, "us/iter %.1f" % (1e3 * r.stats["loop_ms"] / 100),
              {n: round(1e3 * v / 100, 1) for n, v in zip(PROF_NAMES, r.stats["prof_ms"]) if v > 0},
              "launches/iter", r.stats["gpu_launches"] / 100, flush=True)
