"""Per-iteration time of single fits at cfg4's shape (3000 x 20000) for several k and both engines."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from swne_b200 import NMFHandle, synth
from swne_b200._capi import PROF_NAMES
c = synth.CONFIGS["cfg4"]
dA = synth.make_matrix_torch(c["n"], c["m"], c["kt"], c["scale"], c["seed"])
h = NMFHandle(0); h.set_matrix(dA)
for k in (24, 32, 36, 40):
    W0, H0 = synth.random_init(c["n"], c["m"], k, 0)
    for eng in ("tc", "simt"):
        h.fit(k, W0, H0, max_iter=3, rel_tol=-1.0, engine=eng, full_traces=-1)
        t0 = time.perf_counter()
        r = h.fit(k, W0, H0, max_iter=50, rel_tol=-1.0, engine=eng, full_traces=-1)
        wall = time.perf_counter() - t0
        rp = h.fit(k, W0, H0, max_iter=50, rel_tol=-1.0, engine=eng, full_traces=-1, profile=True)
        print("k=%d %s: %.1f us/iter (wall %.1f ms for 50)" % (k, eng, 1e3 * r.stats["loop_ms"] / 50, 1e3 * wall),
              {n: round(1e3 * v / 50, 1) for n, v in zip(PROF_NAMES, rp.stats["prof_ms"]) if v > 0}, flush=True)
