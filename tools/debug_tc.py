import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import small_matrix, seeded_init, rel_fro, max_rel
from swne_b200 import NMFHandle
shapes = [(256, 384, 5, 8), (300, 700, 6, 8), (3000, 4500, 9, 30), (1000, 5000, 9, 16)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
h = NMFHandle(0)
for (n, m, kt, k) in shapes:
    A = small_matrix(n, m, kt, 1)
    W0, H0 = seeded_init(A, k, 1)
    h.set_matrix(A)
    for iters in (1, 5):
        rs = h.fit(k, W0, H0, max_iter=iters, trace=1, rel_tol=-1.0, engine="simt")
        t0 = time.time()
        rt = h.fit(k, W0, H0, max_iter=iters, trace=1, rel_tol=-1.0, engine="tc")
        print((n, m, k), "iters", iters, "engine", rt.options["engine"], "loss rel", max_rel(rt.target_loss, rs.target_loss),
              "W", rel_fro(rt.W, rs.W), "H", rel_fro(rt.H, rs.H), "t=%.2fs" % (time.time() - t0), flush=True)
print("done")
