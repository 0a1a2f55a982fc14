import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from swne_b200 import NMFHandle, synth
c = synth.CONFIGS["cfg3"]
m = int(os.environ.get("CELLS", c["m"]))
dA = synth.make_matrix_torch(c["n"], m, c["kt"], c["scale"], c["seed"])
h = NMFHandle(0)
h.set_matrix(dA)
W0, H0 = synth.random_init(c["n"], m, c["k"], 0)
warm = h.fit(c["k"], W0, H0, max_iter=6, rel_tol=-1.0)
for eng in sys.argv[1:] or ["tc"]:
    for imi in (1, 5, 50):
        r = h.fit(c["k"], warm.W, warm.H, max_iter=10, rel_tol=-1.0, engine=eng, inner_max_iter=imi, trace=5)
        print(eng, "inner_max_iter", imi, "ms/iter %.3f" % (r.stats["loop_ms"] / 10), "epochs", r.average_epochs, "launches", r.stats["gpu_launches"], flush=True)
