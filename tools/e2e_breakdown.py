"""Where the end-to-end time of one host-buffer fit goes (cfg3)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from swne_b200 import NMFHandle, synth
c = synth.CONFIGS["cfg3"]
n, m, k = c["n"], c["m"], c["k"]
dA = synth.make_matrix_torch(n, m, c["kt"], c["scale"], c["seed"])
h = NMFHandle(0)
hostA = torch.empty((m, n), dtype=torch.float32, pin_memory=True); hostA.copy_(dA)
An = hostA.numpy().T
W0, H0 = synth.random_init(n, m, k, 0)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h.set_matrix(An); torch.cuda.synchronize(); t1 = time.perf_counter()
    r = h.fit(k, W0, H0, max_iter=30, rel_tol=-1.0, full_traces=-1); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("set_matrix %.2f ms (%.1f GB/s), fit wall %.2f ms: loop %.2f upload %.2f download %.2f -> python/other %.2f" % (
        1e3 * (t1 - t0), An.nbytes / (t1 - t0) / 1e9, 1e3 * (t2 - t1), r.stats["loop_ms"], r.stats["upload_ms"], r.stats["download_ms"],
        1e3 * (t2 - t1) - r.stats["loop_ms"] - r.stats["upload_ms"] - r.stats["download_ms"]), flush=True)
