import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from swne_b200 import NMFHandle, synth
c = synth.CONFIGS["cfg3"]
dA = synth.make_matrix_torch(c["n"], c["m"], c["kt"], c["scale"], c["seed"])
h = NMFHandle(0); h.set_matrix(dA)
W0, H0 = synth.random_init(c["n"], c["m"], c["k"], 0)
r = h.fit(c["k"], W0, H0, max_iter=60, rel_tol=-1.0, trace=1, full_traces=-1)
print("epochs/iter (random init):", np.round(r.average_epochs, 1).tolist())
print("loop ms/iter %.3f" % (r.stats["loop_ms"] / 60))
r = h.fit(c["k"], W0, H0, max_iter=500, rel_tol=1e-4)
print("to tol: iters", r.n_iteration, "loop ms", r.stats["loop_ms"], "mkl", r.mkl[-1])
Wn, Hn, sv = h.nnsvd_init(c["k"])
rng = np.random.default_rng(0)
from swne_b200 import zero_replace
ini = zero_replace(dict(W=Wn, H=Hn), float(h.matrix_info()["sum"]) / (c["n"] * c["m"]), rng)
r = h.fit(c["k"], ini["W"], ini["H"], max_iter=60, rel_tol=-1.0, trace=1, full_traces=-1)
print("epochs/iter (nnsvd init):", np.round(r.average_epochs, 1).tolist())
print("loop ms/iter %.3f" % (r.stats["loop_ms"] / 60))
r = h.fit(c["k"], ini["W"], ini["H"], max_iter=500, rel_tol=1e-4)
print("to tol (nnsvd): iters", r.n_iteration, "loop ms", r.stats["loop_ms"])
