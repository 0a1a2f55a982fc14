"""Wall time of the NNDSVD seed path on cfg3 (GPU rSVD + host split + zero replacement)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from swne_b200 import NMFHandle, synth, zero_replace
c = synth.CONFIGS[os.environ.get("CFG", "cfg3")]
dA = synth.make_matrix_torch(c["n"], c["m"], c["kt"], c["scale"], c["seed"])
h = NMFHandle(0); h.set_matrix(dA)
k = c["k"]
info = h.matrix_info()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    Wn, Hn, sv = h.nnsvd_init(k, seed=1); t1 = time.perf_counter()
    ini = zero_replace(dict(W=Wn, H=Hn), info["sum"] / (float(info["n"]) * info["m"]), np.random.default_rng(0)); t2 = time.perf_counter()
    r = h.fit(k, ini["W"], ini["H"], max_iter=500, rel_tol=1e-4); t3 = time.perf_counter()
    print("nnsvd_init %.1f ms, zero_replace %.1f ms, fit wall %.1f ms (loop %.1f ms, %d iterations, loss %.6g), sv[:3] %s" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), r.stats["loop_ms"], r.n_iteration, r.target_loss[-1], np.round(sv[:3], 3)), flush=True)
os.environ["SWNE_NNSVD_SIMT"] = "1"
t0 = time.perf_counter(); Wn2, Hn2, sv2 = h.nnsvd_init(k, seed=1); t1 = time.perf_counter()
print("fp32-kernel sketch: nnsvd_init %.1f ms; max |sv diff| rel %.2e; W rel diff %.2e" % (
    1e3 * (t1 - t0), np.max(np.abs(sv - sv2) / sv2), np.linalg.norm(Wn - Wn2) / np.linalg.norm(Wn2)))
