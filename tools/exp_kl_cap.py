"""mkl at BASELINE cfg2 (2000 x 3000, k = 16): time per iteration and GPU-vs-GPU agreement for several bounds on the staged
kernel variants: the default (one warp per column, scd_kl_warp_kernel), the one-block-per-column kernel (SWNE_KL_WARP=0) at
several thread counts (SWNE_KL_NT) and with / without its two length classes (SWNE_KL_SPLIT=0), and the coordinate-stepped
gene side that a cell-sharded fit uses (SWNE_KL_SHARD_STEP).  Earlier runs of this script (profiles/r2_kl_cap_experiment.txt)
also bounded the staged column length (SWNE_KL_CAP)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from helpers import rel_fro, max_rel
from swne_b200 import NMFHandle, synth, zero_replace

c = synth.CONFIGS["cfg1"]
dA = synth.make_matrix_torch(c["n"], c["m"], c["kt"], c["scale"], c["seed"])
A = dA.cpu().numpy()
print("cells: nnz/col mean %.0f max %d; genes: nnz/row mean %.0f max %d" % ((A != 0).sum(1).mean(), (A != 0).sum(1).max(),
                                                                         (A != 0).sum(0).mean(), (A != 0).sum(0).max()), flush=True)
h = NMFHandle(0); h.set_matrix(dA)
Wn, Hn, _ = h.nnsvd_init(c["k"], seed=1)
info = h.matrix_info()
ini = zero_replace(dict(W=Wn, H=Hn), info["sum"] / (float(info["n"]) * info["m"]), np.random.default_rng(0))
kw = dict(loss="mkl", beta=[0, 0, 0.1], rel_tol=-1.0)
base = None
for name, env in [("default", {}), ("warp kernel", {"SWNE_KL_WARP": "1"}), ("block kernel", {"SWNE_KL_WARP": "0"}), ("block NT 512", {"SWNE_KL_WARP": "0", "SWNE_KL_NT": "512"}),
                  ("block NT 256", {"SWNE_KL_WARP": "0", "SWNE_KL_NT": "256"}), ("block NT 64", {"SWNE_KL_WARP": "0", "SWNE_KL_NT": "64"}),
                  ("block one launch", {"SWNE_KL_WARP": "0", "SWNE_KL_SPLIT": "0"}),
                  ("block 512 one launch", {"SWNE_KL_WARP": "0", "SWNE_KL_SPLIT": "0", "SWNE_KL_NT": "512"}),
                  ("sharded step", {"SWNE_KL_SHARD_STEP": "1"})]:
    os.environ.update(env)
    try:
        h.fit(c["k"], ini["W"], ini["H"], max_iter=5, trace=100, **kw)
        r = h.fit(c["k"], ini["W"], ini["H"], max_iter=200, trace=100, **kw)
        p = h.fit(c["k"], ini["W"], ini["H"], max_iter=40, trace=1, **kw)
    finally:
        for e in env:
            del os.environ[e]
    if base is None:
        base = p
    print("cfg2 mkl %-22s %.1f us/iteration | vs default after 40 iterations: loss %.1e W %.1e H %.1e" % (
        name, 1e3 * r.stats["loop_ms"] / 200, max_rel(p.target_loss, base.target_loss), rel_fro(p.W, base.W), rel_fro(p.H, base.H)), flush=True)
h.close()
