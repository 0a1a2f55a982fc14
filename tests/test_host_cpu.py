"""CPU tests (no GPU): the C-ABI library loads and exports every declared symbol, fails loudly without a
GPU (no CPU fallback), the host-side logic (sammon, icafast, sharding) matches the oracle, and the N>1
cell-sharded algorithm equals the unsharded one (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helpers import rel_fro, seeded_init, small_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from swne_b200 import _capi
    lib = _capi.load()
    header = open(os.path.join(ROOT, "include", "swne_nmf.h")).read()
    declared = set(re.findall(r"\b(swne_[a-z0-9_]+)\s*\(", header))
    declared -= {"swne_nmf_progress_fn"}
    assert declared == set(_capi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.swne_nmf_abi_version() == 2


def test_struct_layout_matches_header():
    from swne_b200 import _capi
    lib = _capi.load()
    p = _capi.Params()
    lib.swne_nmf_default_params(ctypes.byref(p), 7)
    assert p.k == 7 and p.max_iter == 500 and abs(p.rel_tol - 1e-4) < 1e-20 and abs(p.inner_rel_tol - 1e-9) < 1e-20
    assert p.loss == 0 and p.engine == 0 and list(p.alpha) == [0, 0, 0] and p.profile == 0 and p.nnlm_variant == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from swne_b200 import NMFHandle
    from swne_b200._capi import SwneError
    with pytest.raises(SwneError) as e:
        NMFHandle(0)
    assert e.value.code == -2


def test_zero_replace_host_entry():
    from swne_b200 import zero_replace
    rng = np.random.default_rng(0)
    init = dict(W=np.array([[0.0, 5e-7], [2e-6, 1.0]]), H=np.array([[0.0, 3.0]]))
    out = zero_replace(init, 2.0, rng)
    assert np.all(out["W"] > 0) and out["W"][1, 1] == 1.0 and out["W"][1, 0] == 2e-6 and out["H"][0, 1] == 3.0
    assert out["W"][0, 0] <= 0.02 and out["H"][0, 0] <= 0.02


def test_host_sammon_matches_oracle():
    from oracle import sammon_ref
    from swne_b200 import sammon
    rng = np.random.default_rng(4)
    P = rng.uniform(size=(8, 2))
    D = np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1)) + 0.05 * (1 - np.eye(8))
    assert np.max(np.abs(sammon(D) - sammon_ref(D))) < 1e-9


def test_ica_whitening_is_the_truncated_svd_of_the_centred_matrix():
    """the identity the device ICA seed rests on (swne_nmf.cu ica_seed_device): icafast's whitened data, built from the
    eigenpairs of the gene covariance, are sqrt(m) V of the SVD A - rowMeans(A) 1' = U S V' (up to column signs), and its
    mixing matrix is U diag(s / sqrt(m)) R."""
    A = small_matrix(40, 120, 4, 9)
    n, m = A.shape
    k = 4
    X = A.T - A.T.mean(axis=0, keepdims=True)
    vals, vecs = np.linalg.eigh(X.T @ X / m)
    order = np.argsort(vals)[::-1][:k]
    Xw = X @ (vecs[:, order] / np.sqrt(vals[order])[None, :])
    U, s, Vt = np.linalg.svd(A - A.mean(axis=1, keepdims=True), full_matrices=False)
    Y = np.sqrt(m) * Vt[:k].T
    sign = np.sign(np.sum(Xw * Y, axis=0))
    assert np.max(np.abs(Xw - Y * sign[None, :])) < 1e-8
    assert np.max(np.abs(vals[order] - s[:k] ** 2 / m)) < 1e-10


def test_shard_bounds():
    from swne_b200.dist import shard_bounds, shard_bounds_nnz
    assert shard_bounds(10, 4) == [0, 3, 6, 8, 10]
    assert shard_bounds(3, 4) == [0, 1, 2, 3, 3]
    b = shard_bounds_nnz(np.array([0, 5, 5, 9, 20, 21, 30]), 3)
    assert b[0] == 0 and b[-1] == 6 and b == sorted(b)


def test_bench_reference_arm_runs():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cells", "2000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"


WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as td
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from helpers import small_matrix, seeded_init
from swne_b200.dist import shard_bounds
from sharded_model import sharded_nnmf_mse
td.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
A = small_matrix(50, 81, 4, 3); W0, H0 = seeded_init(A, 5, 3)
b = shard_bounds(81, 2); r = td.get_rank()
W, H, terr = sharded_nnmf_mse(A[:, b[r]:b[r+1]], W0, H0[:, b[r]:b[r+1]], 12, A.shape[1])
np.savez(sys.argv[4] + ".%d.npz" % r, W=W, H=H, terr=terr)
td.destroy_process_group()
'''


def test_sharded_algorithm_equals_unsharded_gloo(tmp_path):
    """world_size-2 gloo run of the cell-sharded schedule (local H update, all-reduce of A H', H H' and the
    loss partial, replicated W update) against the unsharded oracle."""
    from oracle import nnmf_ref
    port = 29000 + os.getpid() % 2000
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), str(tmp_path / "out")]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    A = small_matrix(50, 81, 4, 3); W0, H0 = seeded_init(A, 5, 3)
    ref = nnmf_ref(A, 5, W0, H0, max_iter=12, trace=1, rel_tol=-1.0)
    r0, r1 = np.load(str(tmp_path / "out") + ".0.npz"), np.load(str(tmp_path / "out") + ".1.npz")
    assert np.array_equal(r0["W"], r1["W"])                              # W is replicated bit for bit
    assert rel_fro(r0["W"], ref["W"]) < 1e-9
    assert rel_fro(np.concatenate([r0["H"], r1["H"]], axis=1), ref["H"]) < 1e-9
    assert np.max(np.abs(r0["terr"] - ref["target_loss"]) / ref["target_loss"]) < 1e-10


MKL_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as td
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from helpers import small_matrix, seeded_init
from swne_b200.dist import shard_bounds
from sharded_model import sharded_nnmf_mkl
td.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
A = small_matrix(40, 61, 4, 3); W0, H0 = seeded_init(A, 5, 3)
b = shard_bounds(61, 2); r = td.get_rank()
out = {}
for inner in (1, 3):
    W, H, mkl = sharded_nnmf_mkl(A[:, b[r]:b[r+1]], W0, H0[:, b[r]:b[r+1]], 8, A.shape[1], (0.05, 0.01, 0.02), (0.02, 0.0, 0.05), inner)
    out.update({"W%d" % inner: W, "H%d" % inner: H, "mkl%d" % inner: mkl})
np.savez(sys.argv[4] + ".%d.npz" % r, **out)
td.destroy_process_group()
'''


def test_sharded_mkl_schedule_equals_unsharded_gloo(tmp_path):
    """world_size-2 gloo run of the cell-sharded mkl schedule (swne_nmf.cu launch_kl_sharded: per coordinate, partial Newton
    sums on the rank's cells -> all-reduce of 2 n numbers -> replicated step with the lazy A-hat update; local cell side)
    against the unsharded oracle, with all three penalties, for one and for three inner sweeps."""
    from oracle import nnmf_ref
    port = 31000 + os.getpid() % 2000
    script = tmp_path / "worker.py"
    script.write_text(MKL_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), str(tmp_path / "out")]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    A = small_matrix(40, 61, 4, 3); W0, H0 = seeded_init(A, 5, 3)
    r0, r1 = np.load(str(tmp_path / "out") + ".0.npz"), np.load(str(tmp_path / "out") + ".1.npz")
    for inner in (1, 3):
        ref = nnmf_ref(A, 5, W0, H0, alpha=(0.05, 0.01, 0.02), beta=(0.02, 0.0, 0.05), loss="mkl", max_iter=8, trace=1, rel_tol=-1.0,
                       inner_max_iter=inner)
        assert np.array_equal(r0["W%d" % inner], r1["W%d" % inner])          # W is replicated bit for bit
        assert rel_fro(r0["W%d" % inner], ref["W"]) < 1e-9
        assert rel_fro(np.concatenate([r0["H%d" % inner], r1["H%d" % inner]], axis=1), ref["H"]) < 1e-9
        assert np.max(np.abs(r0["mkl%d" % inner] - ref["mkl"]) / np.abs(ref["mkl"])) < 1e-10


REPLICA_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as td
sys.path.insert(0, sys.argv[1])
from swne_b200.dist import find_num_factors_replicas
td.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)

class FakeHandle:
    """stands in for NMFHandle: the errors are a fixed function of k, as they are for the library (fits seeded by (seed, k))"""
    def __init__(self): self.calls = []
    def set_matrix(self, A): self.shape = A.shape
    def find_num_factors(self, ks, loss="mse", max_iter=250, seed=0, errors_only=False):
        assert errors_only
        self.calls.append(list(ks))
        ks = np.asarray(ks, dtype=float)
        return dict(k_err=np.stack([10.0 / ks + 0.3 * (ks > 5), 12.0 / np.sqrt(ks)]))

h = FakeHandle()
r = find_num_factors_replicas(h, np.zeros((4, 6)), [2, 3, 4, 5, 6, 8, 10], seed=3)
np.savez(sys.argv[4] + ".%d.npz" % td.get_rank(), k_err=r["k_err"], err=r["err"], k=r["k"], mine=np.array(h.calls[0]))
td.destroy_process_group()
'''


def test_find_num_factors_replicas_gloo(tmp_path):
    """FindNumFactors over two ranks as replicas (dist.find_num_factors_replicas): every rank sweeps a strided share of
    k.range, the errors are gathered, and all ranks apply R/nmf.R:186-194's selection to the same table."""
    port = 31000 + os.getpid() % 2000
    script = tmp_path / "replica_worker.py"
    script.write_text(REPLICA_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), str(tmp_path / "rep")]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    r0, r1 = np.load(str(tmp_path / "rep") + ".0.npz"), np.load(str(tmp_path / "rep") + ".1.npz")
    ks = np.array([2, 3, 4, 5, 6, 8, 10], dtype=float)
    want = np.stack([10.0 / ks + 0.3 * (ks > 5), 12.0 / np.sqrt(ks)])
    assert list(r0["mine"]) == [2, 4, 6, 10] and list(r1["mine"]) == [3, 5, 8]
    for r in (r0, r1):
        assert np.allclose(r["k_err"], want, rtol=0, atol=1e-15)
        d = (want[0, :-1] - want[0, 1:]) - (want[1, :-1] - want[1, 1:])
        assert np.allclose(r["err"], d)
        neg = np.nonzero(d < 0)[0]
        assert int(r["k"]) == int(ks[int(neg[0]) if neg.size else int(np.argmin(d)) + 1])


def test_build_digest_is_location_independent(tmp_path):
    """the tree is copied to another absolute path on the GPU box: the shipped .so must still count as current there
    (a rebuild on the box costs GPU minutes, and eight ranks rebuilding at once used to race)"""
    import importlib.util, shutil
    src = os.path.join(ROOT, "swne_b200")
    dst = tmp_path / "elsewhere" / "swne_b200"
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("*.so", "*.o", "build", "__pycache__"))
    os.makedirs(tmp_path / "elsewhere" / "include")
    shutil.copy(os.path.join(ROOT, "include", "swne_nmf.h"), tmp_path / "elsewhere" / "include" / "swne_nmf.h")
    spec = importlib.util.spec_from_file_location("build_elsewhere", str(dst / "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from swne_b200 import build as here
    assert mod._digest() == here._digest()


def test_permuted_entries_is_a_permutation():
    """A_rand of FindNumFactors (R/nmf.R:163) -- the host fallback taken where there is no CUDA device: same shape,
    same multiset of entries, different arrangement"""
    from swne_b200.nmf import _permuted_entries
    rng = np.random.default_rng(3)
    A = rng.gamma(0.3, 1.0, size=(40, 70))
    P = np.asarray(_permuted_entries(A, np.random.default_rng(5)))
    assert P.shape == A.shape
    assert np.array_equal(np.sort(P.ravel()), np.sort(A.ravel())) and not np.array_equal(P, A)


def test_block_deferred_gauss_seidel_equals_sequential_sweeps():
    """The tensor-core solver (pass_tc.cuh tc_block_gs / contract_tc.cuh tcx_solve_kernel) defers, inside a block of
    16 coordinates, the rank-1 updates of all mu entries outside the block to one rank-16 update at the block's end,
    and works in the unit-diagonal scaled variables h' = d o h, mu' = mu / d, G' = G / (d d').  Both rewrites leave
    the iterates of scd_ls_update unchanged: numpy model of exactly that schedule against the plain loop."""
    rng = np.random.default_rng(7)
    K, sweeps = 40, 6
    X = rng.gamma(0.5, 1.0, size=(K, 90))
    G = X @ X.T + 0.05 * np.eye(K) + 1e-16 * np.eye(K)
    a = rng.gamma(0.5, 1.0, size=90)
    h0 = rng.uniform(size=K)

    def plain(h):
        h = h.copy(); mu = G @ h - X @ a
        for _ in range(sweeps):
            for kk in range(K):
                tmp = max(h[kk] - mu[kk] / G[kk, kk], 0.0)
                if tmp != h[kk]:
                    mu += (tmp - h[kk]) * G[:, kk]
                    h[kk] = tmp
        return h

    def blocked(h):
        d = np.sqrt(np.diag(G)); Gs = G / np.outer(d, d)
        hs = h * d; mus = (G @ h - X @ a) / d
        for _ in range(sweeps):
            for b0 in range(0, K, 16):
                b1 = min(K, b0 + 16)
                m16 = mus[b0:b1].copy(); delta = np.zeros(b1 - b0)
                for i in range(b1 - b0):
                    kk = b0 + i
                    tmp = max(hs[kk] - m16[i], 0.0)
                    delta[i] = tmp - hs[kk]
                    hs[kk] = tmp
                    m16[i + 1:] += delta[i] * Gs[kk, b0 + i + 1:b1]          # triangular in-block part, in registers
                mus += delta @ Gs[b0:b1, :]                                  # the deferred rank-16 update (one MMA)
        return hs / d

    hp, hb = plain(h0), blocked(h0)
    assert np.max(np.abs(hp - hb)) <= 1e-12 * max(1.0, np.max(np.abs(hp)))
    assert np.array_equal(hp == 0, hb == 0)                                   # the same active set


def test_lazy_ahat_update_of_the_kl_sweep_equals_the_eager_one():
    """scd_kl_kernel applies A-hat += delta * w_kk inside the accumulation loop of the NEXT coordinate (one pass over
    the column per coordinate instead of two) and never applies the last one: same iterates as scd_kl_update."""
    rng = np.random.default_rng(8)
    K, R = 6, 50
    Wt = rng.gamma(0.5, 1.0, size=(K, R)); a = rng.gamma(0.5, 1.0, size=R) * (rng.uniform(size=R) < 0.5)
    sumW = Wt.sum(1); pen = (0.03, 0.01, 0.02); T = 1e-16

    def step(h, kk, Ajt, sumH):
        mu = Wt[kk] / (Ajt + T)
        A2 = a @ (mu * mu) + pen[0]
        B = a @ mu - sumW[kk] + A2 * h[kk] - pen[2] - pen[1] * (sumH - h[kk])
        return max(B / (A2 + T), 0.0)

    def eager(h):
        h = h.copy(); Ajt = Wt.T @ h; sumH = h.sum()
        for _ in range(3):
            for kk in range(K):
                tmp = step(h, kk, Ajt, sumH)
                Ajt = Ajt + (tmp - h[kk]) * Wt[kk]; sumH += tmp - h[kk]; h[kk] = tmp
        return h

    def lazy(h):
        h = h.copy(); Ajt = Wt.T @ h; sumH = h.sum(); d_prev, k_prev = 0.0, 0
        for _ in range(3):
            for kk in range(K):
                if d_prev != 0.0:
                    Ajt = np.maximum(Ajt + d_prev * Wt[k_prev], 0.0)          # applied on the way into the next coordinate
                tmp = step(h, kk, Ajt, sumH)
                d_prev, k_prev = tmp - h[kk], kk
                sumH += d_prev; h[kk] = tmp
        return h

    h0 = rng.uniform(size=K) + 0.1
    assert np.allclose(eager(h0), lazy(h0), rtol=1e-12, atol=1e-14)


def test_split_fp16_three_product_scheme_is_fp32_grade():
    """numpy model of the tensor-core operands (pass_tc.cuh / contract_tc.cuh): x * 2^e = hi + lo in fp16 with the
    maximum scaled to ~2^14, product = hi*hi + hi*lo + lo*hi accumulated in fp32.  Its error must sit at fp32
    level (the parity bar is 1e-4 on the loss, 1e-3 on W/H), where a single fp16 or TF32 pass does not."""
    rng = np.random.default_rng(9)
    A = (rng.gamma(0.6, 1.0, size=(256, 600)) * (rng.uniform(size=(256, 600)) < 0.25)).astype(np.float32)   # cells x genes
    W = rng.gamma(0.5, 0.3, size=(600, 30)).astype(np.float32)

    def split(x):
        e = 14 - int(np.frexp(float(np.abs(x).max()))[1])
        xs = x.astype(np.float64) * 2.0 ** e
        hi = xs.astype(np.float16)
        lo = (xs - hi.astype(np.float64)).astype(np.float16)
        return hi.astype(np.float32), lo.astype(np.float32), e

    ah, al, ea = split(A); wh, wl, ew = split(W)
    assert np.isfinite(ah).all() and np.isfinite(wh).all() and float(ah.max()) < 65504
    emu = (ah @ wh + ah @ wl + al @ wh).astype(np.float64) * 2.0 ** (-(ea + ew))
    exact = A.astype(np.float64) @ W.astype(np.float64)
    scale = np.abs(exact).max()
    err3 = np.abs(emu - exact).max() / scale
    err1 = np.abs((ah @ wh).astype(np.float64) * 2.0 ** (-(ea + ew)) - exact).max() / scale
    tf32 = lambda x: (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    errt = np.abs((tf32(A.copy()) @ tf32(W.copy())).astype(np.float64) - exact).max() / scale
    assert err3 < 2e-6, err3                      # fp32-grade (2^-21 operand error, fp32 accumulation)
    assert err1 > 20 * err3 and errt > 20 * err3  # one fp16 pass / one TF32 pass are 2-3 orders worse


def test_bench_clock_sampler_summary_and_peaks(tmp_path, monkeypatch):
    """bench.py's clock line: median SM clock under load, throttle reasons by name, only samples after mark();
    the roofline denominator comes from MEASURED_PEAKS.json when present, else the profiling guide's fallback"""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    cs = bench.ClockSampler(0)
    cs.samples = [(1.0, ["210", "1965", "80.1", "Not Active", "Not Active", "Not Active", "Not Active"])]
    cs.t_start = 2.0
    cs.samples += [(2.1, ["1965", "1965", "410.5", "Not Active", "Not Active", "Not Active", "Active"]),
                   (2.2, ["1950", "1965", "455.0", "Not Active", "Not Active", "Not Active", "Not Active"]),
                   (2.3, ["1965", "1965", "430.0", "Not Active", "Not Active", "Not Active", "Not Active"])]
    assert cs.n_loaded() == 3
    out = cs.summary()
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
    assert out["power_w_max"] == 455.0 and out["samples"] == 3
    empty = bench.ClockSampler(0).summary()
    assert empty["samples"] == 0 and empty["sm_mhz"] is None
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.peaks() == (6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)")
    (tmp_path / "MEASURED_PEAKS.json").write_text('{"hbm_gbs": 6549.1, "bf16_tflops": 1645}')
    assert bench.peaks()[0] == 6549.1


def test_rcpp_shim_type_checks_against_the_c_abi():
    """r/src/swne_nmf_cuda.cpp (the Rcpp glue a maintainer adds next to src/swne_cpp_funcs.cpp) cannot be built here -- no
    R -- but every call it makes into include/swne_nmf.h can be type-checked: g++ -fsyntax-only against a declarations-only
    stand-in for <Rcpp.h> (tests/rcpp_stub).  A changed C-ABI signature that the shim does not follow fails here."""
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "rcpp_stub"),
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "r", "src", "swne_nmf_cuda.cpp")]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]


def test_r_overlay_keeps_the_reference_formals():
    """The R overlay (r/R/nmf_cuda.R) must keep the formals of the functions it replaces: compared as text with
    the signatures in R/nmf.R:96-97, :157-158, :222, :241-242 and R/swne_plotting.R:69 (recorded here: the reference tree is
    not available on the GPU box)."""
    src = open(os.path.join(ROOT, "r", "R", "nmf_cuda.R")).read()
    squeeze = lambda s: "".join(s.split())
    for sig in ('RunNMF <- function(A, k, alpha = 0, init = "ica", n.cores = 1, loss = "mse", max.iter = 500, ica.fast = F)',
                'FindNumFactors <- function(A, k.range = seq(2, 12, 2), n.cores = 1, do.plot = T, seed = NULL, loss = "mse", max.iter = 250)',
                'ProjectFeatures <- function(newdata, H, alpha = rep(0, 3), loss = "mse", n.cores = 1)',
                'ProjectSamples <- function(newdata, W, features.use = NULL, alpha = rep(0, 3), loss = "mse", n.cores = 1)',
                'get_sample_coords <- function(H, H.coords, alpha, n_pull)',
                'get_factor_coords <- function(H, method, distance, n.neighbors, min.dist)',
                'nnsvd_init <- function(A, k, LINPACK)',
                'ica_init <- function(A, k, ica.fast = F)'):
        assert squeeze(sig) in squeeze(src), sig
