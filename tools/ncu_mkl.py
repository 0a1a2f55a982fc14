"""short cfg2 (mkl, 2000 x 3000, k = 16) run for ncu captures of scd_kl_kernel: ITERS outer iterations from the random start"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swne_b200 import NMFHandle, synth
c = synth.CONFIGS["cfg1"]
dA = synth.make_matrix_torch(c["n"], c["m"], c["kt"], c["scale"], c["seed"])
h = NMFHandle(0); h.set_matrix(dA)
W0, H0 = synth.random_init(c["n"], c["m"], c["k"], 0)
r = h.fit(c["k"], W0, H0, loss="mkl", beta=[0, 0, 0.1], max_iter=int(os.environ.get("ITERS", "6")), rel_tol=-1.0, trace=100)
print("us/iter %.1f" % (1e3 * r.stats["loop_ms"] / r.n_iteration), "launches", r.stats["gpu_launches"])
