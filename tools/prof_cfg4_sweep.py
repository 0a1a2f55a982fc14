"""Where the wall time of the cfg4 FindNumFactors sweep goes."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from swne_b200 import NMFHandle, synth
c = synth.CONFIGS["cfg4"]
dA = synth.make_matrix_torch(c["n"], c["m"], c["kt"], c["scale"], c["seed"])
A = dA.cpu().numpy().T
n, m = A.shape
h = NMFHandle(0)
rng = np.random.default_rng(4)
t = time.perf_counter(); A_rand = rng.permutation(A.ravel(order="F")).reshape((n, m), order="F"); t_perm = time.perf_counter() - t
tot = dict(perm=t_perm, set=0.0, init=0.0, fit=0.0, loop=0.0, recon=0.0)
iters = []
for mat in (A, A_rand):
    t = time.perf_counter(); h.set_matrix(mat); tot["set"] += time.perf_counter() - t
    for k in range(2, 41):
        t = time.perf_counter(); W0 = rng.uniform(size=(n, k)) * 0.01; H0 = rng.uniform(size=(k, m)) * 0.01; tot["init"] += time.perf_counter() - t
        t = time.perf_counter(); z = h.fit(k, W0, H0, loss="mse", max_iter=250); tot["fit"] += time.perf_counter() - t
        tot["loop"] += z.stats["loop_ms"] / 1e3; iters.append(z.n_iteration)
        t = time.perf_counter(); h.recon_error(z.W, z.H); tot["recon"] += time.perf_counter() - t
print({k: round(v, 3) for k, v in tot.items()}, "iterations total", sum(iters), "max", max(iters))
