"""Sensitivity of the NMF solution to storing A as fp16 x 2^e (an 11-bit mantissa, i.e. TF32-grade input): the oracle run on A
against the oracle run on the rounded matrix, same init.  This is the deviation a pass that streams only the fp16 hi part of
A (half the HBM traffic) would add to the parity errors; DESIGN.md section 5 cites the result (profiles/r2_a16_sensitivity.txt).
Not a test (pytest does not collect it): python tests/exp_a16_sensitivity.py"""
import sys, time, numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import small_matrix, seeded_init, rel_fro
from oracle import nnmf_ref

def round16(A):
    e = 14 - int(np.ceil(np.log2(A.max())))
    return (A * 2.0 ** e).astype(np.float16).astype(np.float64) / 2.0 ** e

cases = [("cfg1 2000x3000 k16 kt9", 2000, 3000, 9, 16, 1, [10, 50, 200, 500]),
         ("3000x4500 k30 kt20", 3000, 4500, 20, 30, 3, [10, 30, 100]),
         ("600x3000 k16 kt6", 600, 3000, 6, 16, 5, [10, 40, 200]),
         ("cfg3 slice 3000x20000 k30 kt20", 3000, 20000, 20, 30, 3, [10, 30])]
for name, n, m, kt, k, seed, iters_list in cases:
    A = small_matrix(n, m, kt, seed)
    A16 = np.asfortranarray(round16(A))
    W0, H0 = seeded_init(A, k, seed)
    for iters in iters_list:
        t0 = time.time()
        r = nnmf_ref(A, k, W0, H0, max_iter=iters, rel_tol=-1.0, trace=1)
        q = nnmf_ref(A16, k, W0, H0, max_iter=iters, rel_tol=-1.0, trace=1)
        le, lq = np.asarray(r["mse"] if "mse" in r else r["target_loss"]), np.asarray(q["mse"] if "mse" in q else q["target_loss"])
        print(name, "iters", iters, "loss rel", float(np.max(np.abs(lq - le) / le)), "W", rel_fro(q["W"], r["W"]), "H", rel_fro(q["H"], r["H"]),
              "(%.0fs)" % (time.time() - t0), flush=True)
