"""mkl at cfg3's shape (3000 x CELLS, k = 30): gene rows are far longer than the shared-memory staging of scd_kl_kernel, so the
gene side takes the coordinate-stepped kernels (grid-wide pass per coordinate, chosen automatically), and there are enough cell
columns for the warp-per-column kernel.  Per-iteration time (solve_H = cell side, mkl_update = gene side) of the default
against the block kernel (SWNE_KL_WARP=0) and the warp kernel (SWNE_KL_WARP=1) on the cell side, and their agreement."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import rel_fro, max_rel
from swne_b200 import NMFHandle, synth
from swne_b200._capi import PROF_NAMES
c = synth.CONFIGS["cfg3"]
for m in [int(x) for x in os.environ.get("CELLS", "20000,100000").split(",")]:
    dA = synth.make_matrix_torch(c["n"], m, c["kt"], c["scale"], c["seed"])
    h = NMFHandle(0); h.set_matrix(dA)
    W0, H0 = synth.random_init(c["n"], m, c["k"], 0)
    base = None
    for name, env in (("default", {}), ("block kernel", {"SWNE_KL_WARP": "0"}), ("warp kernel", {"SWNE_KL_WARP": "1"})):
        os.environ.update(env)
        try:
            h.fit(c["k"], W0, H0, loss="mkl", max_iter=1, rel_tol=-1.0, trace=100)
            r = h.fit(c["k"], W0, H0, loss="mkl", max_iter=4, rel_tol=-1.0, trace=100, profile=True)
        finally:
            for e in env: del os.environ[e]
        base = base or r
        print("mkl 3000 x %d k=30 %-12s %.2f ms/iteration %s | vs default: loss %.1e W %.1e H %.1e" % (
            m, name, r.stats["loop_ms"] / 4, {n: round(v / 4, 2) for n, v in zip(PROF_NAMES, r.stats["prof_ms"]) if v > 0},
            max_rel(r.target_loss, base.target_loss), rel_fro(r.W, base.W), rel_fro(r.H, base.H)), flush=True)
    h.close()
