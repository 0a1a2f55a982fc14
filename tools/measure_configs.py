"""Round-1 measurements of the BASELINE.json configurations other than the bench workload (cfg3).
    python tools/measure_configs.py [cfg1] [cfg2] [cfg4] [cfg5]
Prints one line per measurement; copy into profiles/."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from swne_b200 import NMFHandle, synth, zero_replace

which = sys.argv[1:] or ["cfg1", "cfg2", "cfg4"]
h = NMFHandle(0)


def seeded(h, k, seed=0):
    Wn, Hn, _ = h.nnsvd_init(k, seed=seed + 1)
    info = h.matrix_info()
    return zero_replace(dict(W=Wn, H=Hn), info["sum"] / (float(info["n"]) * info["m"]), np.random.default_rng(seed))


if "cfg1" in which or "cfg2" in which:
    c = synth.CONFIGS["cfg1"]
    dA = synth.make_matrix_torch(c["n"], c["m"], c["kt"], c["scale"], c["seed"])
    h.set_matrix(dA)
    info = h.matrix_info()
    t0 = time.perf_counter(); ini = seeded(h, c["k"]); torch.cuda.synchronize(); t_seed = time.perf_counter() - t0
    print("cfg1 matrix %dx%d nnz %.1f%%, nnsvd seed (GPU rSVD + host split) %.1f ms" % (c["n"], c["m"], 100.0 * info["nnz"] / (c["n"] * c["m"]), 1e3 * t_seed))
    if "cfg1" in which:
        for eng in ("tc", "simt"):
            r = h.fit(c["k"], ini["W"], ini["H"], max_iter=500, engine=eng)
            r2 = h.fit(c["k"], ini["W"], ini["H"], max_iter=200, rel_tol=-1.0, engine=eng, full_traces=-1)
            print("cfg1 mse k=16 engine=%s: to tol %d iterations, loop %.2f ms; fixed 200 iterations %.1f us/iteration (%.0f it/s), final loss %.6g" % (
                eng, r.n_iteration, r.stats["loop_ms"], 1e3 * r2.stats["loop_ms"] / 200, 200 / (r2.stats["loop_ms"] * 1e-3), r.target_loss[-1]))
    if "cfg2" in which:
        for beta, name in (([0, 0, 0.1], "beta=(0,0,0.1) L1 on H"), ([0, 0, 0], "no penalty")):
            r = h.fit(c["k"], ini["W"], ini["H"], loss="mkl", beta=beta, max_iter=200, rel_tol=-1.0, trace=100)
            print("cfg2 mkl k=16 %s: fixed 200 iterations %.1f us/iteration (%.0f it/s), mkl %.6g" % (
                name, 1e3 * r.stats["loop_ms"] / 200, 200 / (r.stats["loop_ms"] * 1e-3), r.mkl[-1]))
        r = h.fit(c["k"], ini["W"], ini["H"], loss="mkl", alpha=0.1, max_iter=200, rel_tol=-1.0, trace=100)
        print("cfg2 mkl k=16 scalar alpha=0.1 (L2 on W, what RunNMF can express): %.1f us/iteration" % (1e3 * r.stats["loop_ms"] / 200))

if "cfg4" in which:
    from swne_b200 import FindNumFactors
    c = synth.CONFIGS["cfg4"]
    dA = synth.make_matrix_torch(c["n"], c["m"], c["kt"], c["scale"], c["seed"])
    A = dA.cpu().numpy().T          # genes x cells view
    t0 = time.perf_counter()
    r = FindNumFactors(A, k_range=list(range(2, 41)), seed=4, max_iter=250, handle=h)
    dt = time.perf_counter() - t0
    print("cfg4 FindNumFactors k=2..40 on %dx%d (78 fits, mse, random init, max.iter 250): %.2f s wall, picked k=%d" % (c["n"], c["m"], dt, r["k"]))

if "cfg5" in which:
    c = synth.CONFIGS["cfg5"]
    m = int(os.environ.get("CFG5_CELLS", c["m"]))
    dA = synth.make_matrix_torch(c["n"], m, c["kt"], c["scale"], c["seed"])
    h.set_matrix(dA)
    info = h.matrix_info()
    W0, H0 = synth.random_init(c["n"], m, c["k"], 0)
    r = h.fit(c["k"], W0, H0, max_iter=10, rel_tol=-1.0, full_traces=-1)
    print("cfg5 dense fp32 %dx%d (nnz %.1f%%) k=40 engine=%s: %.2f ms/iteration on 1 GPU" % (c["n"], m, 100.0 * info["nnz"] / (c["n"] * float(m)), r.options["engine"], r.stats["loop_ms"] / 10))
h.close()
