"""Generates tests/golden/*.npz from the CPU oracle (oracle/nnmf_ref.c + oracle/ref.py).

The reference ships no golden vectors for this path and cannot run here (no R, NNLM un-vendored), so these
fixtures pin the ORACLE's outputs (regression pin for the oracle, parity target for the CUDA path on the
GPU box where /root/reference and the generating environment do not exist).  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from helpers import seeded_init, small_matrix  # noqa: E402
from oracle import (get_sample_coords_ref, factor_distance_ref, nnlm_ref, nnmf_ref, nnsvd_init_ref)  # noqa: E402

CASES = {
    # name: (n, m, kt, k, seed, kwargs)
    "mse_plain": (60, 90, 4, 5, 11, dict(loss="mse", max_iter=30, trace=1)),
    "mse_pen": (60, 90, 4, 5, 12, dict(loss="mse", alpha=[0.02, 0.01, 0.005], beta=[0.01, 0.0, 0.02], max_iter=30, trace=1)),
    "mkl_plain": (60, 90, 4, 5, 13, dict(loss="mkl", max_iter=30, trace=1)),
    "mkl_l1h": (60, 90, 4, 5, 14, dict(loss="mkl", beta=[0.0, 0.0, 0.1], max_iter=30, trace=1)),
    "mse_k3_ragged": (37, 53, 3, 3, 15, dict(loss="mse", max_iter=20, trace=2)),
}


def main():
    for name, (n, m, kt, k, seed, kw) in CASES.items():
        A = small_matrix(n, m, kt, seed)
        W0, H0 = seeded_init(A, k, seed)
        r = nnmf_ref(A, k, W0, H0, rel_tol=-1.0, n_threads=1, **kw)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), A=A, W0=W0, H0=H0, W=r["W"], H=r["H"], mse=r["mse"],
                            mkl=r["mkl"], target_loss=r["target_loss"], average_epochs=r["average_epochs"],
                            n_iteration=r["n_iteration"], k=k,
                            alpha=np.asarray(kw.get("alpha", [0, 0, 0]), dtype=float),
                            beta=np.asarray(kw.get("beta", [0, 0, 0]), dtype=float), loss=kw["loss"],
                            max_iter=kw["max_iter"], trace=kw["trace"])
    # nnsvd / consumers / nnlm
    A = small_matrix(60, 90, 4, 21)
    Wn, Hn = nnsvd_init_ref(A, 5)
    rng = np.random.default_rng(5)
    H = rng.gamma(0.5, 1.0, size=(6, 40)); H[:, 3] = 0; H[2, 7] = H[4, 7]        # an empty cell and a tie
    coords = rng.uniform(size=(6, 2))
    sc = get_sample_coords_ref(H, coords, alpha=1.3, n_pull=3)
    W = rng.gamma(0.5, 1.0, size=(60, 5))
    lm = nnlm_ref(W, A, init=np.full((5, 90), 0.005))
    np.savez_compressed(os.path.join(HERE, "aux.npz"), A=A, Wn=Wn, Hn=Hn, H=H, coords=coords, sample_coords=sc,
                        d_cos=factor_distance_ref(H, "cosine"), d_pear=factor_distance_ref(H, "pearson"),
                        d_euc=factor_distance_ref(H, "euclidean"), W=W, nnlm_coef=lm["coefficients"])
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
