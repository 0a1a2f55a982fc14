"""short cfg3 run for ncu captures of the fused pass: ITERS outer iterations from the random start"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swne_b200 import NMFHandle, synth
c = synth.CONFIGS["cfg3"]
m = int(os.environ.get("CELLS", c["m"]))
dA = synth.make_matrix_torch(c["n"], m, c["kt"], c["scale"], c["seed"])
h = NMFHandle(0); h.set_matrix(dA)
W0, H0 = synth.random_init(c["n"], m, c["k"], 0)
r = h.fit(c["k"], W0, H0, max_iter=int(os.environ.get("ITERS", "6")), rel_tol=-1.0, engine="tc", full_traces=-1)
print("ms/iter %.3f" % (r.stats["loop_ms"] / r.n_iteration), "launches", r.stats["gpu_launches"])
