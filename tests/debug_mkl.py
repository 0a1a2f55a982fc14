import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import small_matrix, seeded_init, rel_fro
from oracle import nnmf_ref
from swne_b200 import NMFHandle
A = small_matrix(200, 300, 5, 6)
W0, H0 = seeded_init(A, 6, 6)
h = NMFHandle(0)
h.set_matrix(A)
for it in (1, 2, 3):
    ref = nnmf_ref(A, 6, W0, H0, loss="mkl", max_iter=it, trace=1, rel_tol=-1.0)
    r = h.fit(6, W0, H0, loss="mkl", max_iter=it, trace=1, rel_tol=-1.0)
    print(it, "mkl", r.mkl, ref["mkl"], "mse", r.mse, ref["mse"], "W", rel_fro(r.W, ref["W"]), "H", rel_fro(r.H, ref["H"]))
    W, H = r.W, r.H
    Ah = W @ H
    mkl = np.mean((A + 1e-16) * np.log(A + 1e-16) - A) + np.mean(-(A + 1e-16) * np.log(Ah + 1e-16) + Ah)
    print("   recomputed from our W,H:", mkl, " sum ahat", Ah.sum(), " min Ah on support", Ah[A > 0].min())
