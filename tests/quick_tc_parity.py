import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from helpers import small_matrix, seeded_init, rel_fro, max_rel
from oracle import nnmf_ref
from swne_b200 import NMFHandle
h = NMFHandle(0)
for (n, m, kt, k, seed) in [(256, 384, 5, 8, 1), (3000, 4500, 9, 30, 2), (301, 517, 6, 30, 2), (1000, 4133, 6, 3, 4)]:
    A = small_matrix(n, m, kt, seed); W0, H0 = seeded_init(A, k, seed)
    ref = nnmf_ref(A, k, W0, H0, max_iter=15, trace=1, rel_tol=-1.0)
    h.set_matrix(A)
    r = h.fit(k, W0, H0, max_iter=15, trace=1, rel_tol=-1.0, engine="tc")
    print(n, m, k, r.options["engine"], "loss", max_rel(r.target_loss, ref["target_loss"]), "W", rel_fro(r.W, ref["W"]), "H", rel_fro(r.H, ref["H"]),
          "epochs", max_rel(r.average_epochs, ref["average_epochs"]), flush=True)
    kw = dict(alpha=[0.05, 0.02, 0.01], beta=[0.03, 0.01, 0.02], max_iter=8, trace=1, rel_tol=-1.0)
    r = h.fit(k, W0, H0, engine="tc", **kw); ref = nnmf_ref(A, k, W0, H0, **kw)
    print("   pen: loss", max_rel(r.target_loss, ref["target_loss"]), "W", rel_fro(r.W, ref["W"]), "H", rel_fro(r.H, ref["H"]), flush=True)
