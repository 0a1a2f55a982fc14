"""Wide engine (k = 40) slab pipeline: ms per iteration for several slab counts and the event timeline of one iteration
(SWNE_TCX_TIMELINE, printed by the library on stderr).  CELLS (default 650000) x 3000 genes, dense on the device."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from swne_b200 import NMFHandle, synth
c = synth.CONFIGS["cfg5"]; m = int(os.environ.get("CELLS", "650000"))
dA = synth.make_matrix_torch(c["n"], m, c["kt"], c["scale"], c["seed"])
h = NMFHandle(0); h.set_matrix(dA)
W0, H0 = synth.random_init(c["n"], m, c["k"], 0)
w = h.fit(c["k"], W0, H0, max_iter=4, rel_tol=-1.0, full_traces=-1)
for slabs in sys.argv[1:] or ("1", "2", "4", "8"):
    os.environ["SWNE_TCX_SLABS"] = slabs
    r = h.fit(c["k"], w.W, w.H, max_iter=10, rel_tol=-1.0, full_traces=-1)
    print("slabs %s: %.3f ms/iteration" % (slabs, r.stats["loop_ms"] / 10), flush=True)
os.environ["SWNE_TCX_SLABS"] = os.environ.get("TL_SLABS", "4")
os.environ["SWNE_TCX_TIMELINE"] = "1"
sys.stderr.flush()
r = h.fit(c["k"], w.W, w.H, max_iter=5, rel_tol=-1.0, full_traces=-1)
