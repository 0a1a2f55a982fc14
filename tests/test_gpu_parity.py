"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI, against the CPU oracle on the same
seeded inputs and against the committed golden fixtures.  Bars (BASELINE.json north_star): per-iteration
reconstruction loss within 1e-4 relative, final W/H within 1e-3 relative Frobenius (fp32 vs the fp64 oracle)."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse

from helpers import max_rel, rel_fro, seeded_init, small_matrix, within

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOSS_TOL = 1e-4     # relative, per trace entry
FACTOR_TOL = 1e-3   # relative Frobenius of final W and H


def _check_against(res, ref, loss_tol=LOSS_TOL, factor_tol=FACTOR_TOL, key="target_loss"):
    assert res["n_iteration"] == ref["n_iteration"]
    assert len(res[key]) == len(ref[key])
    within("loss trace, max rel", max_rel(res[key], ref[key]), loss_tol)
    within("W rel Frobenius", rel_fro(res["W"], ref["W"]), factor_tol)
    within("H rel Frobenius", rel_fro(res["H"], ref["H"]), factor_tol)


@pytest.mark.parametrize("path", sorted(f for f in glob.glob(os.path.join(GOLD, "mse*.npz")) if not f.endswith(".nnlm.npz")))
def test_golden_mse(handle, path):
    g = np.load(path)
    handle.set_matrix(g["A"])
    r = handle.fit(int(g["k"]), g["W0"], g["H0"], alpha=g["alpha"], beta=g["beta"], loss="mse",
                   max_iter=int(g["max_iter"]), trace=int(g["trace"]), rel_tol=-1.0, full_traces=True)
    _check_against(r, g)
    assert max_rel(r["mse"], g["mse"]) < LOSS_TOL
    assert max_rel(r["mkl"], g["mkl"]) < LOSS_TOL


def test_true_nnlm_goldens(handle):
    """The CUDA path against outputs of the real NNLM package (tests/golden/<name>.nnlm.npz, written by tools/make_golden.R
    where R + NNLM exist).  None are committed -- neither can run in the build image -- so parity is UNPINNED and this test
    skips saying so; with the files present it holds the CUDA path to the north-star bars directly against NNLM."""
    files = sorted(glob.glob(os.path.join(GOLD, "*.nnlm.npz")))
    if not files:
        pytest.skip("parity unpinned: no true NNLM goldens (run tools/make_golden.R where R + NNLM exist)")
    for f in files:
        t, g = np.load(f), np.load(f.replace(".nnlm.npz", ".npz"))
        handle.set_matrix(g["A"])
        r = handle.fit(int(g["k"]), g["W0"], g["H0"], alpha=g["alpha"], beta=g["beta"], loss=str(g["loss"]),
                       max_iter=int(g["max_iter"]), trace=int(g["trace"]), rel_tol=-1.0)
        within("true NNLM golden: loss trace, max rel", max_rel(r["target_loss"], t["target_loss"]), LOSS_TOL)
        within("true NNLM golden: W rel Frobenius", rel_fro(r["W"], t["W"]), FACTOR_TOL)
        within("true NNLM golden: H rel Frobenius", rel_fro(r["H"], t["H"]), FACTOR_TOL)


@pytest.mark.parametrize("path", sorted(f for f in glob.glob(os.path.join(GOLD, "mkl*.npz")) if not f.endswith(".nnlm.npz")))
def test_golden_mkl(handle, path):
    g = np.load(path)
    handle.set_matrix(g["A"])
    r = handle.fit(int(g["k"]), g["W0"], g["H0"], alpha=g["alpha"], beta=g["beta"], loss="mkl",
                   max_iter=int(g["max_iter"]), trace=int(g["trace"]), rel_tol=-1.0)
    _check_against(r, g)          # measured round 2: loss 7e-7, W/H 4e-6 (profiles/r2_parity_measured.json)
    within("mse of the mkl fit, max rel", max_rel(r["mse"], g["mse"]), LOSS_TOL)


@pytest.mark.parametrize("n,m,kt,k,seed", [(200, 300, 5, 5, 1), (2000, 3000, 9, 16, 1), (301, 517, 6, 30, 2), (128, 257, 6, 40, 3)])
def test_mse_vs_oracle(handle, n, m, kt, k, seed):
    from oracle import nnmf_ref
    A = small_matrix(n, m, kt, seed)
    W0, H0 = seeded_init(A, k, seed)
    iters = 20
    ref = nnmf_ref(A, k, W0, H0, max_iter=iters, trace=1, rel_tol=-1.0)
    handle.set_matrix(A)
    r = handle.fit(k, W0, H0, max_iter=iters, trace=1, rel_tol=-1.0, engine="simt")
    _check_against(r, ref)
    assert r["H"].min() >= 0 and r["W"].min() >= 0
    within("average_epochs, max rel", max_rel(r["average_epochs"], ref["average_epochs"]), 0.35)      # fp32 stagnates earlier than fp64


@pytest.mark.parametrize("n,m,kt,k,seed", [(256, 384, 5, 8, 1), (2000, 3000, 9, 16, 1), (3000, 4500, 9, 30, 2), (301, 517, 6, 30, 2),
                                           (1000, 4133, 6, 3, 4)])
def test_tcgen05_engine_vs_oracle(handle, n, m, kt, k, seed):
    """the fused TMA/tcgen05 pass (split-fp16 3-product contractions) against the fp64 oracle"""
    from oracle import nnmf_ref
    A = small_matrix(n, m, kt, seed)
    W0, H0 = seeded_init(A, k, seed)
    iters = 15
    ref = nnmf_ref(A, k, W0, H0, max_iter=iters, trace=1, rel_tol=-1.0)
    handle.set_matrix(A)
    r = handle.fit(k, W0, H0, max_iter=iters, trace=1, rel_tol=-1.0, engine="tc")
    assert r.options["engine"] == "tc"
    _check_against(r, ref)
    assert r["H"].min() >= 0 and r["W"].min() >= 0
    # penalties go through the same pass; with full traces the mkl of the mse fit is evaluated at every block
    kw = dict(alpha=[0.05, 0.02, 0.01], beta=[0.03, 0.01, 0.02], max_iter=8, trace=1, rel_tol=-1.0)
    rp, refp = handle.fit(k, W0, H0, engine="tc", full_traces=True, **kw), nnmf_ref(A, k, W0, H0, **kw)
    _check_against(rp, refp)
    assert max_rel(rp["mkl"], refp["mkl"]) < LOSS_TOL and max_rel(rp["mse"], refp["mse"]) < LOSS_TOL


@pytest.mark.parametrize("n,m,kt,k,seed", [(300, 700, 12, 40, 31), (520, 1000, 20, 64, 32), (257, 385, 10, 33, 33)])
def test_tcgen05_wide_contractions_vs_oracle(handle, n, m, kt, k, seed):
    """32 < k <= 64: un-fused tcgen05 contractions (contract_tc.cuh) + CUDA-core solves against the fp64 oracle,
    including partial tiles in both directions and padded factor columns"""
    from oracle import nnmf_ref
    A = small_matrix(n, m, kt, seed)
    W0, H0 = seeded_init(A, k, seed)
    iters = 10
    ref = nnmf_ref(A, k, W0, H0, max_iter=iters, trace=1, rel_tol=-1.0)
    handle.set_matrix(A)
    r = handle.fit(k, W0, H0, max_iter=iters, trace=1, rel_tol=-1.0, engine="tc")
    assert r.options["engine"] == "tc"
    _check_against(r, ref)
    assert r["H"].min() >= 0 and r["W"].min() >= 0
    # the automatic choice takes the same engine, and the fp32 engine agrees to fp32 accuracy
    ra = handle.fit(k, W0, H0, max_iter=iters, trace=1, rel_tol=-1.0)
    assert ra.options["engine"] == "tc"
    rs = handle.fit(k, W0, H0, max_iter=iters, trace=1, rel_tol=-1.0, engine="simt")
    assert max_rel(r["target_loss"], rs["target_loss"]) < 1e-5
    kw = dict(alpha=[0.05, 0.02, 0.01], beta=[0.03, 0.01, 0.02], max_iter=6, trace=1, rel_tol=-1.0)
    _check_against(handle.fit(k, W0, H0, engine="tc", **kw), nnmf_ref(A, k, W0, H0, **kw))


def test_mse_default_stopping_rule_matches(handle):
    from oracle import nnmf_ref
    A = small_matrix(300, 500, 6, 4)
    W0, H0 = seeded_init(A, 8, 4)
    ref = nnmf_ref(A, 8, W0, H0)                       # NNLM defaults: 500 / 1e-4 / trace 2
    handle.set_matrix(A)
    r = handle.fit(8, W0, H0)
    assert abs(r["n_iteration"] - ref["n_iteration"]) <= 2   # the stopping test sits on a 1e-4 threshold
    assert abs(r["target_loss"][-1] - ref["target_loss"][-1]) / ref["target_loss"][-1] < 1e-4
    assert not r["not_converged"]


def test_mse_penalties_vs_oracle(handle):
    from oracle import nnmf_ref
    A = small_matrix(200, 300, 5, 5)
    W0, H0 = seeded_init(A, 6, 5)
    kw = dict(alpha=[0.05, 0.02, 0.01], beta=[0.03, 0.01, 0.02], max_iter=25, trace=1, rel_tol=-1.0)
    ref = nnmf_ref(A, 6, W0, H0, **kw)
    handle.set_matrix(A)
    r = handle.fit(6, W0, H0, **kw)
    _check_against(r, ref)
    # RunNMF's scalar alpha is an L2 penalty on W
    ref2 = nnmf_ref(A, 6, W0, H0, alpha=0.1, max_iter=10, trace=1, rel_tol=-1.0)
    r2 = handle.fit(6, W0, H0, alpha=0.1, max_iter=10, trace=1, rel_tol=-1.0)
    _check_against(r2, ref2)


def test_mkl_vs_oracle_with_self_noise(handle):
    """mkl parity is reported next to the oracle's own summation-order noise on the same input."""
    from oracle import nnmf_ref
    A = small_matrix(200, 300, 5, 6)
    W0, H0 = seeded_init(A, 6, 6)
    kw = dict(loss="mkl", beta=[0, 0, 0.1], max_iter=30, trace=1, rel_tol=-1.0)
    ref = nnmf_ref(A, 6, W0, H0, **kw)
    p = np.random.default_rng(0).permutation(200)
    ref_perm = nnmf_ref(A[p], 6, W0[p], H0, **kw)
    noise = max(rel_fro(ref_perm["W"], ref["W"][p]), rel_fro(ref_perm["H"], ref["H"]), 1e-7)
    handle.set_matrix(A)
    r = handle.fit(6, W0, H0, **kw)
    within("mkl target_loss, max rel", max_rel(r["target_loss"], ref["target_loss"]), LOSS_TOL)
    within("mkl trace, max rel", max_rel(r["mkl"], ref["mkl"]), LOSS_TOL)
    tol = max(FACTOR_TOL, 1e4 * noise)     # fp32 (6e-8) vs the oracle's fp64 (1e-16) perturbation, same amplification
    within("oracle self-noise (row permutation)", noise, 1.0)
    within("W rel Frobenius", rel_fro(r["W"], ref["W"]), tol)
    within("H rel Frobenius", rel_fro(r["H"], ref["H"]), tol)


def test_unpinned_nnlm_details_both_readings(handle):
    """The two details of NNLM that Appendix A recalls without certainty are switches on both sides (oracle and C-ABI): the
    CUDA path must match the oracle under BOTH readings, so a future golden from real NNLM decides with no code change.
      1. the default of `trace` (2 = 100 / inner.max.iter, or 10 as in NNLM 0.4.1): it sets where the stopping rule is tested
      2. scd_kl_update's Newton step with an L2 penalty (a += beta0 before or after b += a*h)"""
    from oracle import nnmf_ref
    A = small_matrix(300, 500, 6, 4)
    W0, H0 = seeded_init(A, 8, 4)
    handle.set_matrix(A)
    for trace in (2, 10):
        ref = nnmf_ref(A, 8, W0, H0, trace=trace)                 # NNLM's stopping rule, 500 / 1e-4
        r = handle.fit(8, W0, H0, trace=trace)
        assert abs(r["n_iteration"] - ref["n_iteration"]) <= trace   # the test sits on a 1e-4 threshold, checked every `trace`
        assert abs(r["target_loss"][-1] - ref["target_loss"][-1]) / ref["target_loss"][-1] < 1e-4
        assert r["options"]["trace"] == trace
    refs = {}
    for late in (False, True):
        kw = dict(loss="mkl", alpha=[0.2, 0, 0], beta=[0.1, 0, 0.05], max_iter=30, trace=1, rel_tol=-1.0)
        refs[late] = nnmf_ref(A, 8, W0, H0, kl_l2_late=late, **kw)
        r = handle.fit(8, W0, H0, kl_l2_late=late, **kw)
        within("mkl target_loss, max rel (kl_l2_late=%d)" % late, max_rel(r["target_loss"], refs[late]["target_loss"]), LOSS_TOL)
        within("W rel Frobenius (kl_l2_late=%d)" % late, rel_fro(r["W"], refs[late]["W"]), FACTOR_TOL)
        within("H rel Frobenius (kl_l2_late=%d)" % late, rel_fro(r["H"], refs[late]["H"]), FACTOR_TOL)
    # the two readings are distinguishable on this input (otherwise the test above proves nothing)
    assert rel_fro(refs[True]["H"], refs[False]["H"]) > 1e-3


def test_mkl_sharded_gene_side_kernels_on_one_gpu(handle):
    """With cells sharded over ranks the gene side of an mkl fit runs coordinate by coordinate around an all-reduce
    (kl_shard_* kernels, swne_nmf.cu launch_kl_sharded).  SWNE_KL_SHARD_STEP takes exactly those kernels on one GPU (the
    all-reduce of one rank is the identity): same bars against the oracle as the one-block-per-gene kernel, for one inner
    sweep (NNLM's default for mkl) and for three (genes leave the sweeps one by one), and for nnlm on the gene side."""
    from oracle import nnmf_ref, nnlm_ref
    A = small_matrix(300, 500, 6, 4)
    W0, H0 = seeded_init(A, 8, 4)
    handle.set_matrix(A)
    # three inner sweeps start from the oracle's state after 30 default iterations: from the NNDSVD start a second sweep drives
    # a-hat to ~1e-16 on entries with a > 0, where a log(a-hat) turns 1e-7 of fp32 rounding into 1e-3 of loss (measured on
    # both kernels; the reported loss equals the fp64 loss of the returned W, H to 3e-9 -- conditioning, not bookkeeping)
    pen = dict(alpha=[0.2, 0.01, 0.02], beta=[0.1, 0, 0.05])
    warm = nnmf_ref(A, 8, W0, H0, loss="mkl", max_iter=30, trace=1, rel_tol=-1.0, **pen)
    os.environ["SWNE_KL_SHARD_STEP"] = "1"
    try:
        for inner, Wi, Hi, iters in ((1, W0, H0, 20), (3, warm["W"], warm["H"], 10)):
            kw = dict(loss="mkl", max_iter=iters, trace=1, rel_tol=-1.0, inner_max_iter=inner, **pen)
            ref = nnmf_ref(A, 8, Wi, Hi, **kw)
            r = handle.fit(8, Wi, Hi, **kw)
            within("sharded-step mkl target_loss, max rel (inner=%d)" % inner, max_rel(r["target_loss"], ref["target_loss"]), LOSS_TOL)
            within("sharded-step mse of the mkl fit, max rel (inner=%d)" % inner, max_rel(r["mse"], ref["mse"]), LOSS_TOL)
            within("sharded-step W rel Frobenius (inner=%d)" % inner, rel_fro(r["W"], ref["W"]), FACTOR_TOL)
            within("sharded-step H rel Frobenius (inner=%d)" % inner, rel_fro(r["H"], ref["H"]), FACTOR_TOL)
            within("sharded-step average_epochs, max rel (inner=%d)" % inner, max_rel(r["average_epochs"], ref["average_epochs"]), 0.05)
        # ProjectFeatures with loss = mkl: nnlm on the gene side
        Hfix = np.random.default_rng(4).gamma(0.5, 1.0, size=(6, 500))
        refW = nnlm_ref(Hfix.T, A.T, loss="mkl", init=np.full((6, 300), 0.005), max_iter=200)["coefficients"].T
        rW = handle.nnlm(1, Hfix, init=np.full((300, 6), 0.005), loss="mkl", max_iter=200)["coefficients"]
        within("sharded-step nnlm gene side (mkl)", rel_fro(rW, refW), FACTOR_TOL)
    finally:
        del os.environ["SWNE_KL_SHARD_STEP"]


def test_mkl_warp_per_column_kernel_vs_oracle(handle):
    """scd_kl_warp_kernel (one warp per column, factor-major gathers) is chosen by column count (>= 9472 columns); SWNE_KL_WARP=1
    forces it on a small case and on the golden fixtures' shapes: same bars against the oracle as the block kernel, with all
    three penalties on both sides, the alternative reading of the L2 step, and nnlm on both sides."""
    from oracle import nnmf_ref, nnlm_ref
    A = small_matrix(300, 500, 6, 4)
    W0, H0 = seeded_init(A, 8, 4)
    handle.set_matrix(A)
    os.environ["SWNE_KL_WARP"] = "1"
    try:
        for late in (False, True):
            kw = dict(loss="mkl", alpha=[0.2, 0.01, 0.02], beta=[0.1, 0, 0.05], max_iter=30, trace=1, rel_tol=-1.0)
            ref = nnmf_ref(A, 8, W0, H0, kl_l2_late=late, **kw)
            r = handle.fit(8, W0, H0, kl_l2_late=late, **kw)
            within("warp-kernel mkl target_loss, max rel (kl_l2_late=%d)" % late, max_rel(r["target_loss"], ref["target_loss"]), LOSS_TOL)
            within("warp-kernel mse of the mkl fit, max rel (kl_l2_late=%d)" % late, max_rel(r["mse"], ref["mse"]), LOSS_TOL)
            within("warp-kernel W rel Frobenius (kl_l2_late=%d)" % late, rel_fro(r["W"], ref["W"]), FACTOR_TOL)
            within("warp-kernel H rel Frobenius (kl_l2_late=%d)" % late, rel_fro(r["H"], ref["H"]), FACTOR_TOL)
            within("warp-kernel average_epochs, max rel", max_rel(r["average_epochs"], ref["average_epochs"]), 0.05)
        Hfix = np.random.default_rng(4).gamma(0.5, 1.0, size=(6, 500))
        refW = nnlm_ref(Hfix.T, A.T, loss="mkl", init=np.full((6, 300), 0.005), max_iter=200)["coefficients"].T
        rW = handle.nnlm(1, Hfix, init=np.full((300, 6), 0.005), loss="mkl", max_iter=200)["coefficients"]
        within("warp-kernel nnlm gene side (mkl)", rel_fro(rW, refW), FACTOR_TOL)
        refH = nnlm_ref(refW, A, loss="mkl", init=np.full((6, 500), 0.005), max_iter=200)["coefficients"]
        rH = handle.nnlm(0, refW, init=np.full((6, 500), 0.005), loss="mkl", max_iter=200)["coefficients"]
        within("warp-kernel nnlm cell side (mkl)", rel_fro(rH, refH), FACTOR_TOL)
    finally:
        del os.environ["SWNE_KL_WARP"]


def _match_components(S, Sref):
    """|correlation| of every reference component with its best match among S (components are defined up to sign/order)"""
    S = (S - S.mean(axis=1, keepdims=True)); Sref = (Sref - Sref.mean(axis=1, keepdims=True))
    S /= np.linalg.norm(S, axis=1, keepdims=True); Sref /= np.linalg.norm(Sref, axis=1, keepdims=True)
    return np.max(np.abs(Sref @ S.T), axis=1)


def test_device_ica_init_vs_oracle(handle):
    """ica_init on the device (centred randomized SVD + FastICA fixed point) against the oracle's icafast restatement
    (exact eigen-decomposition): same independent components up to sign and order, same mixing matrix."""
    from oracle import ica_init_ref
    A = small_matrix(300, 2000, 5, 13, density_scale=2.0)
    k = 5
    Wr, Hr = ica_init_ref(A, k)
    handle.set_matrix(A)
    W, H = handle.ica_init(k, seed=3)
    c = _match_components(H.copy(), Hr.copy())
    assert c.min() > 0.999, c
    cw = _match_components(W.T.copy(), Wr.T.copy())
    assert cw.min() > 0.999, cw
    # scale: icafast's sources have unit variance
    assert np.max(np.abs(H.std(axis=1) - 1.0)) < 1e-3
    # ica.fast = T: the reference picks k of 50 equal-eigenvalue directions (numerically arbitrary); the device version
    # takes the k leading ones.  Property checks: unit-variance sources, W = (A - rowMeans) S
    Wf, Hf = handle.ica_init(k, ica_fast=True, seed=3)
    assert np.max(np.abs(Hf.std(axis=1) - 1.0)) < 1e-3
    assert rel_fro(Wf, (A - A.mean(axis=1, keepdims=True)) @ Hf.T) < 1e-3


def test_run_nmf_device_seeds(handle):
    """RunNMF with every seed computed on the device (swne_run_nmf): the fit must start from a valid point (monotone
    loss from iteration 1) and end as well as the oracle's RunNMF with the same kind of seed."""
    from oracle import run_nmf_ref
    from swne_b200 import RunNMF
    A = small_matrix(300, 1500, 6, 14, density_scale=2.0)
    k = 6
    for init in ("random", "nnsvd", "ica"):
        r = RunNMF(A, k, init=init, max_iter=60, seed=2, handle=handle, rel_tol=-1.0, trace=1)
        t = r["target_loss"]
        assert r["W"].shape == (300, k) and r["H"].shape == (k, 1500) and r["W"].min() >= 0 and r["H"].min() >= 0
        assert np.all(np.diff(t) <= 1e-6 * t[:-1])
        ref = run_nmf_ref(A, k, init=init, max_iter=60, seed=2, rel_tol=-1.0, trace=1)
        assert t[-1] < 1.02 * ref["target_loss"][-1], (init, t[-1], ref["target_loss"][-1])
    # the NNDSVD start is far better than the random one after one iteration (that is its point)
    r1 = RunNMF(A, k, init="nnsvd", max_iter=1, seed=2, handle=handle, rel_tol=-1.0, trace=1)
    r2 = RunNMF(A, k, init="random", max_iter=1, seed=2, handle=handle, rel_tol=-1.0, trace=1)
    assert r1["target_loss"][0] < r2["target_loss"][0]


def test_csc_input_equals_dense(handle):
    A = small_matrix(150, 220, 5, 7)
    W0, H0 = seeded_init(A, 5, 7)
    handle.set_matrix(A)
    info_d = handle.matrix_info()
    # the CUDA-core engine is run-to-run deterministic at this size (one cell chunk, no atomics race)
    rd = handle.fit(5, W0, H0, max_iter=10, trace=1, rel_tol=-1.0, engine="simt")
    S = scipy.sparse.csc_matrix(A)
    handle.set_matrix(S)
    info_s = handle.matrix_info()
    rs = handle.fit(5, W0, H0, max_iter=10, trace=1, rel_tol=-1.0, engine="simt")
    assert info_d["nnz"] == info_s["nnz"] == S.nnz
    assert abs(info_d["sumsq"] - info_s["sumsq"]) < 1e-9 * info_d["sumsq"]
    assert np.array_equal(rd["H"], rs["H"]) and np.array_equal(rd["W"], rs["W"])
    assert max_rel(rs["mse"], rd["mse"]) < 1e-12
    # float32 CSC, an empty column and an empty matrix row
    S32 = S.astype(np.float32).tolil(); S32[:, 3] = 0; S32[5, :] = 0
    S32 = S32.tocsc()
    handle.set_matrix(S32)
    r32 = handle.fit(5, W0, H0, max_iter=5, trace=1, rel_tol=-1.0, engine="simt")
    r32t = handle.fit(5, W0, H0, max_iter=5, trace=1, rel_tol=-1.0, engine="tc")
    handle.set_matrix(np.asfortranarray(S32.toarray()))
    rdd = handle.fit(5, W0, H0, max_iter=5, trace=1, rel_tol=-1.0, engine="simt")
    assert np.array_equal(r32["H"], rdd["H"])
    assert rel_fro(r32t["H"], rdd["H"]) < 1e-4 and rel_fro(r32t["W"], rdd["W"]) < 1e-4


def test_counts_ingest_fuses_scale_counts(handle):
    """swne_nmf_set_matrix_csc_counts: ScaleCounts (R/data_processing.R:275-307: depth scaling, gene variance factor, log or
    Freeman-Tukey on the stored entries) applied during the upload gives the matrix the host-side transform gives."""
    rng = np.random.default_rng(8)
    n, m = 150, 260
    full = rng.poisson(rng.gamma(0.3, 2.0, size=(400, m)))           # all genes: defines the depth
    counts = scipy.sparse.csc_matrix(full[:n].astype(np.float64))   # the variable genes handed to RunNMF
    depth = full.sum(axis=0).astype(np.float64)
    gsf = rng.uniform(0.5, 2.0, size=n)
    W0, H0 = rng.uniform(size=(n, 5)) * 0.01, rng.uniform(size=(5, m)) * 0.01
    for method, f in (("log", np.log1p), ("ft", lambda v: np.sqrt(v) + np.sqrt(v + 1.0)), ("none", lambda v: v)):
        for gs in (None, gsf):
            norm = full[:n] / depth[None, :] * np.median(depth) * (1.0 if gs is None else gs[:, None])
            ref = np.where(full[:n] > 0, f(norm), 0.0)                # the reference transforms the @x slot only
            handle.set_matrix(ref)
            info_r = handle.matrix_info()
            rr = handle.fit(5, W0, H0, max_iter=8, trace=1, rel_tol=-1.0, engine="simt")
            handle.set_matrix_counts(counts, depth=depth, gene_scale=gs, method=method)
            info_c = handle.matrix_info()
            rc = handle.fit(5, W0, H0, max_iter=8, trace=1, rel_tol=-1.0, engine="simt")
            assert info_c["nnz"] == info_r["nnz"] == counts.nnz
            within("fused ScaleCounts(%s%s): sum of A" % (method, ", gsf" if gs is not None else ""), abs(info_c["sum"] - info_r["sum"]) / info_r["sum"], 1e-7)
            within("fused ScaleCounts(%s%s): loss trace" % (method, ", gsf" if gs is not None else ""), max_rel(rc["target_loss"], rr["target_loss"]), 1e-5)
            within("fused ScaleCounts(%s%s): H" % (method, ", gsf" if gs is not None else ""), rel_fro(rc["H"], rr["H"]), 1e-4)
    with pytest.raises(ValueError):
        handle.set_matrix_counts(counts, method="logrank")


def test_float32_and_device_resident_inputs(handle):
    import torch
    A = small_matrix(120, 200, 4, 8)
    W0, H0 = seeded_init(A, 4, 8)
    handle.set_matrix(A.astype(np.float32))
    r32 = handle.fit(4, W0, H0, max_iter=8, trace=1, rel_tol=-1.0, engine="simt")
    dA = torch.from_numpy(np.ascontiguousarray(A.astype(np.float32).T)).cuda()      # (cells, genes)
    handle.set_matrix(dA)
    rdev = handle.fit(4, W0, H0, max_iter=8, trace=1, rel_tol=-1.0, engine="simt")
    assert np.array_equal(r32["W"], rdev["W"]) and np.array_equal(r32["H"], rdev["H"])
    handle.set_matrix(A)       # release the borrowed tensor


def test_nnlm_vs_oracle(handle):
    from oracle import nnlm_ref
    g = np.load(os.path.join(GOLD, "aux.npz"))
    handle.set_matrix(g["A"])
    r = handle.nnlm(0, g["W"], init=np.full((5, 90), 0.005))
    within("nnlm coefficients vs golden", rel_fro(r["coefficients"], g["nnlm_coef"]), FACTOR_TOL)
    # W side (ProjectFeatures): solve genes given H
    A = g["A"]
    rng = np.random.default_rng(2)
    H = rng.gamma(0.5, 1.0, size=(5, 90))
    ref = nnlm_ref(H.T, A.T, init=np.full((5, 60), 0.005))        # coefficients k x genes
    r = handle.nnlm(1, H, init=np.full((60, 5), 0.005))
    within("nnlm (gene side) coefficients", rel_fro(r["coefficients"], ref["coefficients"].T), FACTOR_TOL)
    # mkl projection
    refk = nnlm_ref(g["W"], A, loss="mkl", init=np.full((5, 90), 0.005), max_iter=200)
    rk = handle.nnlm(0, g["W"], init=np.full((5, 90), 0.005), loss="mkl", max_iter=200)
    within("nnlm mkl coefficients", rel_fro(rk["coefficients"], refk["coefficients"]), FACTOR_TOL)


def test_project_features_and_samples(handle):
    from oracle import nnlm_ref
    from swne_b200 import ProjectFeatures, ProjectSamples
    A = small_matrix(80, 120, 4, 9)
    rng = np.random.default_rng(1)
    W = rng.gamma(0.5, 1.0, size=(80, 4)); H = rng.gamma(0.5, 1.0, size=(4, 120))
    Hs = ProjectSamples(A, W, handle=handle)
    Wf = ProjectFeatures(A, H, handle=handle)
    within("ProjectSamples", rel_fro(Hs, nnlm_ref(W, A)["coefficients"]), FACTOR_TOL)
    within("ProjectFeatures", rel_fro(Wf, nnlm_ref(H.T, A.T)["coefficients"].T), FACTOR_TOL)


def test_nnsvd_init_vs_lapack(handle):
    from oracle import nnsvd_init_ref
    A = small_matrix(300, 500, 6, 10, density_scale=2.0)
    Wr, Hr = nnsvd_init_ref(A, 6)
    handle.set_matrix(A)
    W, H, sv = handle.nnsvd_init(6, oversample=10, power_iters=6, seed=3)
    S = np.linalg.svd(A, compute_uv=False)[:6]
    within("singular values vs LAPACK, max rel", max_rel(sv, S), 1e-5)
    assert W.min() >= 0 and H.min() >= 0
    # leading (well separated) factors match the LAPACK-based NNDSVD
    for f in range(6):
        gap = (S[f] - S[f + 1]) / S[f] if f + 1 < len(S) else 1.0
        if f == 0 or gap > 0.05:
            within("NNDSVD column %d of W vs LAPACK" % f, rel_fro(W[:, f], Wr[:, f]), 2e-4)
            within("NNDSVD row %d of H vs LAPACK" % f, rel_fro(H[f], Hr[f]), 2e-4)
    # and it seeds an equally good fit
    from oracle import nnmf_ref
    ref = nnmf_ref(A, 6, *seeded_init(A, 6, 10), max_iter=30, rel_tol=-1.0)
    rng = np.random.default_rng(0)
    from oracle import zero_replace_ref
    W0, H0 = zero_replace_ref(W, H, A.mean(), rng)
    r = handle.fit(6, W0, H0, max_iter=30, rel_tol=-1.0)
    assert r["target_loss"][-1] < 1.02 * ref["target_loss"][-1]


def test_sample_coords_and_distances_vs_golden(handle):
    g = np.load(os.path.join(GOLD, "aux.npz"))
    sc = handle.sample_coords(g["H"], g["coords"], alpha=1.3, n_pull=3)
    ok = ~np.isnan(g["sample_coords"][:, 0])
    assert np.max(np.abs(sc[ok] - g["sample_coords"][ok])) < 1e-12
    assert np.all(np.isnan(sc[~ok]))                         # an all-zero cell is 0/0 in R too
    Hnz = g["H"][:, ok]
    from oracle import factor_distance_ref
    for name in ("cosine", "pearson", "euclidean"):
        assert np.max(np.abs(handle.factor_distance(Hnz, name) - factor_distance_ref(Hnz, name))) < 1e-7  # corrcoef leaves sqrt(2e-16) on the diagonal
    with pytest.raises(Exception):
        handle.factor_distance(Hnz, "IC")


def test_consumers_of_the_resident_h(handle):
    """get_sample_coords and the factor distance of get_factor_coords straight from the H the fit left on the
    device (H = NULL in the C-ABI) -- identical to passing the returned H back in (the host copy holds the same fp32 values),
    for the fp32 and the tcgen05 engine; any later use of the handle's factor buffers ends the residency, loudly."""
    from swne_b200._capi import SwneError
    A = small_matrix(400, 700, 6, 12)
    W0, H0 = seeded_init(A, 7, 12)
    handle.set_matrix(A)
    coords = np.random.default_rng(2).uniform(size=(7, 2))
    for engine in ("simt", "tc"):
        r = handle.fit(7, W0, H0, max_iter=15, engine=engine)
        sc_res = handle.sample_coords(None, coords, alpha=1.3, n_pull=4)
        for dist in ("cosine", "pearson", "euclidean"):
            assert np.array_equal(handle.factor_distance(None, dist, k=7), handle.factor_distance(r.H, dist))
        assert np.array_equal(sc_res, handle.sample_coords(r.H, coords, alpha=1.3, n_pull=4))
        with pytest.raises(SwneError):
            handle.sample_coords(None, coords[:5], alpha=1.3, n_pull=4)          # k of another fit
    handle.recon_error(r.W, r.H)                                              # reuses the factor buffers
    with pytest.raises(SwneError):
        handle.sample_coords(None, coords)


def test_embed_swne_vs_oracle(handle):
    from oracle import get_factor_coords_ref, get_sample_coords_ref
    from swne_b200 import EmbedSWNE
    rng = np.random.default_rng(6)
    H = rng.gamma(0.4, 1.0, size=(7, 400))
    e = EmbedSWNE(H, alpha_exp=1.5, n_pull=4, handle=handle)
    Hc = get_factor_coords_ref(H)
    assert np.max(np.abs(e["H_coords"] - Hc)) < 1e-6
    assert np.max(np.abs(e["sample_coords"] - get_sample_coords_ref(H, e["H_coords"], 1.5, 4))) < 1e-12


def test_find_num_factors_vs_oracle(handle):
    from oracle import find_num_factors_ref
    from swne_b200 import FindNumFactors
    A = small_matrix(120, 200, 4, 11, density_scale=2.0)
    rng = np.random.default_rng(1)
    A_rand = np.asfortranarray(rng.permutation(A.ravel()).reshape(A.shape))
    ks = [2, 3, 4, 5, 6]
    inits = {k: (rng.uniform(size=(120, k)) * 0.01, rng.uniform(size=(k, 200)) * 0.01,
                 rng.uniform(size=(120, k)) * 0.01, rng.uniform(size=(k, 200)) * 0.01) for k in ks}
    ref = find_num_factors_ref(A, ks, inits, A_rand, max_iter=40)
    r = FindNumFactors(A, ks, inits=inits, A_rand=A_rand, max_iter=40, handle=handle)
    within("FindNumFactors k_err (mse), max rel", max_rel(r["k_err"], ref["k_err"]), LOSS_TOL)
    assert r["k"] == ref["k"]
    rk = FindNumFactors(A, ks, inits=inits, A_rand=A_rand, max_iter=40, loss="mkl", handle=handle)
    refk = find_num_factors_ref(A, ks, inits, A_rand, max_iter=40, loss="mkl")
    within("FindNumFactors k_err (mkl), max rel", max_rel(rk["k_err"], refk["k_err"]), LOSS_TOL)


def test_find_num_factors_in_library_sweep(handle):
    """swne_find_num_factors: permutation, random starts, fits and errors on the device (what FindNumFactors calls when
    nothing is injected).  Its random numbers differ from numpy's, so it is compared with the host-orchestrated sweep
    (which the test above pins to the oracle) statistically, and with R/nmf.R:186-194's selection rule exactly."""
    from swne_b200 import FindNumFactors
    A = small_matrix(300, 500, 4, 21, density_scale=2.0)
    ks = [2, 3, 4, 5, 6, 8]
    for loss in ("mse", "mkl"):
        r = FindNumFactors(A, ks, seed=5, loss=loss, max_iter=60, handle=handle)
        rng = np.random.default_rng(7)
        inits = {k: (rng.uniform(size=(300, k)) * 0.01, rng.uniform(size=(k, 500)) * 0.01,
                     rng.uniform(size=(300, k)) * 0.01, rng.uniform(size=(k, 500)) * 0.01) for k in ks}
        A_rand = np.asfortranarray(rng.permutation(A.ravel()).reshape(A.shape))
        rh = FindNumFactors(A, ks, inits=inits, A_rand=A_rand, loss=loss, max_iter=60, handle=handle)
        within("in-library sweep vs host-orchestrated sweep, k_err on A (%s)" % loss, max_rel(r["k_err"][0], rh["k_err"][0]), 1e-2)
        within("in-library sweep vs host-orchestrated sweep, k_err on permuted A (%s)" % loss, max_rel(r["k_err"][1], rh["k_err"][1]), 1e-2)
        # more factors fit better (up to the noise of 60-iteration fits from random starts), structure beats the permutation
        e0 = r["k_err"][0]
        assert e0[-1] < e0[0] and np.all(np.diff(e0) < 5e-3 * e0[:-1]) and np.all(r["k_err"][1] > e0)
        d = (r["k_err"][0, :-1] - r["k_err"][0, 1:]) - (r["k_err"][1, :-1] - r["k_err"][1, 1:])
        assert np.allclose(r["err"], d, rtol=0, atol=1e-15)
        neg = np.nonzero(d < 0)[0]
        assert r["k"] == ks[int(neg[0]) if neg.size else int(np.argmin(d)) + 1]
        assert r["k"] == rh["k"]


def test_run_nmf_front_end(handle):
    from swne_b200 import RunNMF
    from swne_b200._capi import SwneError
    A = small_matrix(100, 150, 4, 12)
    for init in ("nnsvd", "ica", "random"):
        r = RunNMF(A, 4, init=init, max_iter=30, seed=1, handle=handle)
        assert r["W"].shape == (100, 4) and r["H"].shape == (4, 150) and r["factor_names"][0] == "factor_1"
        assert np.all(np.diff(r["target_loss"]) <= 1e-6 * r["target_loss"][:-1])
    with pytest.raises(ValueError, match="negative"):
        RunNMF(-A, 4, handle=handle)
    with pytest.raises(ValueError, match="k must be greater"):
        RunNMF(A, 2, handle=handle)
    with pytest.raises(ValueError, match="Invalid initialization"):
        RunNMF(A, 4, init="svd", handle=handle)
    B = A.copy(); B[3, 4] = np.nan
    with pytest.raises(SwneError):
        handle.set_matrix(B)
    handle.set_matrix(A)
    with pytest.raises(SwneError, match="init W"):
        handle.fit(4, -np.ones((100, 4)), np.ones((4, 150)))
    Hbad = np.ones((4, 150)); Hbad[2, 7] = np.nan          # validated on the device while the init is converted
    with pytest.raises(SwneError, match="init H"):
        handle.fit(4, np.ones((100, 4)), Hbad)
    ok = handle.fit(4, np.ones((100, 4)), np.ones((4, 150)), max_iter=2)     # the handle stays usable after the error
    assert np.all(np.isfinite(ok["W"]))
    with pytest.raises(SwneError):
        handle.fit(70, np.ones((100, 70)), np.ones((70, 150)))     # k > SWNE_MAX_K
    rk = RunNMF(A, 4, init="nnsvd", loss="mkl", max_iter=20, seed=1, handle=handle)
    assert np.all(np.diff(rk["target_loss"]) <= 1e-5 * np.abs(rk["target_loss"][:-1]))


def test_full_size_properties_cfg3(handle):
    """BASELINE config 3 shape (3000 x 100000, k=30): size-independent properties instead of an oracle run:
    monotone target loss, H/W nonnegative, KKT of the last H update, Gram-identity loss equals the direct
    residual."""
    import torch
    from swne_b200 import synth
    c = synth.CONFIGS["cfg3"]
    dA = synth.make_matrix_torch(c["n"], c["m"], c["kt"], c["scale"], c["seed"])
    handle.set_matrix(dA)
    info = handle.matrix_info()
    assert 0.05 < info["nnz"] / (c["n"] * c["m"]) < 0.4
    W0, H0 = synth.random_init(c["n"], c["m"], c["k"], 0)
    r = handle.fit(c["k"], W0, H0, max_iter=12, trace=1, rel_tol=-1.0)
    t = r["target_loss"]
    assert np.all(np.diff(t) <= 1e-6 * t[:-1]) and r["W"].min() >= 0 and r["H"].min() >= 0
    mse_direct, _ = handle.recon_error(r["W"], r["H"])
    assert abs(mse_direct - r["mse"][-1]) / mse_direct < 1e-4
    # KKT of the H update on a sample of cells: grad >= -tol where h == 0, |grad| small where h > 0
    W = torch.from_numpy(r["W"]).cuda()
    idx = torch.arange(0, c["m"], 997, device="cuda")
    a = dA[idx].double()                         # cells x genes
    Hs = torch.from_numpy(np.ascontiguousarray(r["H"][:, idx.cpu().numpy()])).cuda()
    grad = (W.T @ W) @ Hs - W.T @ a.T
    scale = (W.T @ a.T).abs().max()
    assert (grad[Hs > 0].abs().max() / scale) < 5e-3
    assert (grad[Hs == 0].min() / scale) > -5e-3
    handle.set_matrix(np.zeros((4, 4)) + 1.0)    # drop the borrowed tensor


def _cfg3_slice(cells, seed_shift=0):
    """the first `cells` cells of a cfg3-shaped matrix (same generator as bench.py), on the device and as host fp64"""
    from swne_b200 import synth
    c = synth.CONFIGS["cfg3"]
    dA = synth.make_matrix_torch(c["n"], cells, c["kt"], c["scale"], c["seed"] + seed_shift)
    A64 = np.asfortranarray(dA.cpu().numpy().T.astype(np.float64))
    return c, dA, A64


@pytest.mark.parametrize("cells,iters", [(20000, 10), (100000, 6)])
def test_cfg3_slice_vs_oracle(handle, cells, iters):
    """The headline configuration against the oracle itself: 3000 genes, k = 30, fused tcgen05 engine, NNLM's random
    start, every iteration traced.  20 000 cells = 157 tiles (one or two per CTA: the multi-tile queues, the waves of the
    second stream, partial last tiles); 100 000 cells = the full BASELINE config 3 (5-6 tiles per CTA)."""
    from oracle import nnmf_ref
    from swne_b200 import synth
    c, dA, A64 = _cfg3_slice(cells)
    W0, H0 = synth.random_init(c["n"], cells, c["k"], 0)
    ref = nnmf_ref(A64, c["k"], W0, H0, max_iter=iters, trace=1, rel_tol=-1.0)
    handle.set_matrix(dA)
    r = handle.fit(c["k"], W0, H0, max_iter=iters, trace=1, rel_tol=-1.0, engine="tc")
    assert r.options["engine"] == "tc"
    _check_against(r, ref)
    within("mse trace, max rel", max_rel(r["mse"], ref["mse"]), LOSS_TOL)
    # sweep counts: fp32 stops a column a sweep or two earlier than fp64 once its updates fall below fp32 resolution
    within("average_epochs, max rel", max_rel(r["average_epochs"], ref["average_epochs"]), 0.2)
    handle.set_matrix(np.zeros((4, 4)) + 1.0)    # drop the borrowed tensor


def test_cfg2_mkl_at_size_vs_oracle(handle):
    """BASELINE config 2 at size: 2000 x 3000, k = 16, loss = mkl with an L1 penalty on H (beta = (0, 0, 0.1)), and the
    scalar-alpha variant RunNMF can express (L2 on W)."""
    from oracle import nnmf_ref
    from swne_b200 import synth
    c = synth.CONFIGS["cfg1"]
    A = synth.make_matrix_np(c["n"], c["m"], c["kt"], c["scale"], c["seed"], dtype=np.float64)
    W0, H0 = seeded_init(A, c["k"], 1)
    handle.set_matrix(A)
    for kw in (dict(beta=[0, 0, 0.1]), dict(alpha=0.05)):
        kw.update(loss="mkl", max_iter=40, trace=1, rel_tol=-1.0)
        ref = nnmf_ref(A, c["k"], W0, H0, **kw)
        r = handle.fit(c["k"], W0, H0, **kw)
        tag = "beta L1" if "beta" in kw else "alpha L2"
        within("mkl target_loss, max rel (%s)" % tag, max_rel(r["target_loss"], ref["target_loss"]), LOSS_TOL)
        within("mkl trace, max rel (%s)" % tag, max_rel(r["mkl"], ref["mkl"]), LOSS_TOL)
        within("mse of the mkl fit, max rel (%s)" % tag, max_rel(r["mse"], ref["mse"]), LOSS_TOL)
        # the KL coordinate step amplifies rounding where a factor row collapses (40 iterations at this size:
        # measured 8e-4 with the L1 penalty, 2.4e-3 with the L2 one, i.e. fp32's 6e-8 amplified ~4e4 times -- the oracle's own
        # summation-order noise is amplified by the same factor, test_mkl_vs_oracle_with_self_noise); bar = 3x measured
        within("W rel Frobenius (%s)" % tag, rel_fro(r["W"], ref["W"]), 8e-3)
        within("H rel Frobenius (%s)" % tag, rel_fro(r["H"], ref["H"]), 8e-3)


def test_csc_ingest_at_size_equals_dense(handle):
    """dgCMatrix slots (int32 p / i, double x; src/swne_cpp_funcs.cpp:27-31) for a 3000 x 50000 cfg3-shaped matrix:
    same statistics, and the same fit as the dense upload of the same values."""
    from swne_b200 import synth
    c, dA, A64 = _cfg3_slice(50000, seed_shift=7)
    S = scipy.sparse.csc_matrix(A64)
    W0, H0 = synth.random_init(c["n"], 50000, c["k"], 3)
    handle.set_matrix(dA)
    info_d = handle.matrix_info()
    rd = handle.fit(c["k"], W0, H0, max_iter=6, trace=1, rel_tol=-1.0, engine="tc")
    handle.set_matrix(S)
    info_s = handle.matrix_info()
    rs = handle.fit(c["k"], W0, H0, max_iter=6, trace=1, rel_tol=-1.0, engine="tc")
    assert info_d["nnz"] == info_s["nnz"] == S.nnz
    assert abs(info_d["sumsq"] - info_s["sumsq"]) < 1e-9 * info_d["sumsq"]
    # same fp32 values on the device; the second stream accumulates with atomics, so agreement is to rounding
    within("CSC vs dense ingest: loss trace, max rel", max_rel(rs["target_loss"], rd["target_loss"]), 1e-6)
    within("CSC vs dense ingest: W", rel_fro(rs["W"], rd["W"]), 1e-4)
    within("CSC vs dense ingest: H", rel_fro(rs["H"], rd["H"]), 1e-4)
    handle.set_matrix(np.zeros((4, 4)) + 1.0)


def test_tcgen05_fp16_range_overflow_falls_back_to_fp32(handle):
    """H images are fp16 with a scale from the previous iteration's max(H); a > 4096x jump in one iteration must be
    detected and the fit redone on the fp32 engine (never a silent wrong answer)."""
    from oracle import nnmf_ref
    A = small_matrix(256, 384, 5, 21)
    W0, H0 = seeded_init(A, 6, 21)
    # a tiny H0 and a strong L2 penalty on W: the W update cannot absorb the scale, so the first H update grows H
    # by many orders of magnitude within one iteration
    H0 = H0 * 1e-6 + 1e-9
    kw = dict(alpha=1e3, max_iter=6, trace=1, rel_tol=-1.0)
    ref = nnmf_ref(A, 6, W0, H0, **kw)
    assert ref["H"].max() / H0.max() > 1e5
    handle.set_matrix(A)
    r = handle.fit(6, W0, H0, engine="tc", **kw)
    assert r.options["engine"] == "simt"
    # this deliberately badly scaled start is beyond fp32-vs-fp64 parity; the fallback must equal a plain fp32-engine fit
    rs = handle.fit(6, W0, H0, engine="simt", **kw)
    assert np.array_equal(r["W"], rs["W"]) and np.array_equal(r["H"], rs["H"])
    assert max_rel(r["target_loss"], rs["target_loss"]) < 1e-12 and np.all(np.isfinite(r["H"]))


def test_one_shot_entry_points_and_progress_callback(handle):
    """swne_nmf_fit_dense_f64 / swne_nmf_fit_csc_f64 (the NNLM_c_nnmf-shaped calls) and the interrupt callback."""
    import ctypes
    from swne_b200 import _capi as C
    lib = C.load()
    A = small_matrix(150, 260, 4, 22)
    W0, H0 = seeded_init(A, 5, 22)
    handle.set_matrix(A)
    base = handle.fit(5, W0, H0, max_iter=8, trace=2, rel_tol=-1.0, engine="simt")
    p = C.Params(); lib.swne_nmf_default_params(ctypes.byref(p), 5)
    p.max_iter = 8; p.trace = 2; p.rel_tol = -1.0; p.engine = C.ENGINE["simt"]
    L = 8 // 2 + 1
    for kind in ("dense", "csc"):
        W = np.asfortranarray(W0.copy()); H = np.asfortranarray(H0.copy())
        mse = np.zeros(L); mkl = np.zeros(L); terr = np.zeros(L); epo = np.zeros(L)
        res = C.Result()
        if kind == "dense":
            Af = np.asfortranarray(A, dtype=np.float64)
            code = lib.swne_nmf_fit_dense_f64(handle._h, ctypes.byref(p), C.dptr(Af), 150, 260, C.dptr(W), C.dptr(H), C.dptr(mse),
                                              C.dptr(mkl), C.dptr(terr), C.dptr(epo), L, ctypes.byref(res))
        else:
            S = scipy.sparse.csc_matrix(A)
            ip = np.ascontiguousarray(S.indptr, dtype=np.int32); ii = np.ascontiguousarray(S.indices, dtype=np.int32)
            xx = np.ascontiguousarray(S.data, dtype=np.float64)
            code = lib.swne_nmf_fit_csc_f64(handle._h, ctypes.byref(p), C.iptr(ip), C.iptr(ii), C.dptr(xx), 150, 260, C.dptr(W),
                                            C.dptr(H), C.dptr(mse), C.dptr(mkl), C.dptr(terr), C.dptr(epo), L, ctypes.byref(res))
        assert code in (0, 1) and res.n_iteration == 8 and res.n_trace == len(base.mse)
        assert np.array_equal(W, base.W) and np.array_equal(H, base.H)
        assert max_rel(terr[:res.n_trace], base.target_loss) < 1e-12
    # too-short trace arrays are rejected, not overrun
    res = C.Result()
    W = np.asfortranarray(W0.copy()); H = np.asfortranarray(H0.copy()); mse = np.zeros(2)
    assert lib.swne_nmf_fit(handle._h, ctypes.byref(p), C.dptr(W), C.dptr(H), C.dptr(mse), None, None, None, 2, ctypes.byref(res)) == C.SWNE_ERR_BAD_ARG
    # progress callback: called once per error block; a non-zero return interrupts (Rcpp::checkUserInterrupt hook)
    calls = []
    cb = C.PROGRESS_FN(lambda user, it, a, b, c, d: (calls.append(it), 1 if len(calls) == 2 else 0)[1])
    lib.swne_nmf_set_progress(handle._h, cb, None)
    try:
        with pytest.raises(C.SwneError) as e:
            handle.fit(5, W0, H0, max_iter=20, trace=2, rel_tol=-1.0)
        assert e.value.code == C.SWNE_ERR_INTERRUPTED and calls == [1, 3]
    finally:
        lib.swne_nmf_set_progress(handle._h, C.PROGRESS_FN(0), None)


def test_debug_variants_of_the_wide_engine_and_the_seed(handle):
    """SWNE_TCX_SOLVE_SIMT (wide contractions + CUDA-core solves) and SWNE_NNSVD_SIMT (fp32 sketch products) are read
    with getenv at call time, so they can be flipped in-process: both variants must meet the same bars."""
    from oracle import nnmf_ref
    A = small_matrix(400, 900, 12, 41)
    W0, H0 = seeded_init(A, 36, 41)
    ref = nnmf_ref(A, 36, W0, H0, max_iter=8, trace=1, rel_tol=-1.0)
    handle.set_matrix(A)
    os.environ["SWNE_TCX_SOLVE_SIMT"] = "1"
    try:
        r = handle.fit(36, W0, H0, max_iter=8, trace=1, rel_tol=-1.0, engine="tc")
    finally:
        del os.environ["SWNE_TCX_SOLVE_SIMT"]
    assert r.options["engine"] == "tc"
    _check_against(r, ref)
    # the slab pipeline of the wide engine (contractions on a high-priority stream, solves on a low-priority one, one
    # accumulator region in the second contraction) is chosen by size; forced here on 8 tiles cut into 3 and 8 slabs
    for slabs in ("3", "8"):
        os.environ["SWNE_TCX_SLABS"] = slabs
        try:
            rs = handle.fit(36, W0, H0, max_iter=8, trace=1, rel_tol=-1.0, engine="tc")
        finally:
            del os.environ["SWNE_TCX_SLABS"]
        _check_against(rs, ref)
    # the un-fused kernels also serve k <= 32 (their job when a shard is too long for the fused pass)
    W0s, H0s = seeded_init(A, 16, 42)
    refs = nnmf_ref(A, 16, W0s, H0s, max_iter=8, trace=1, rel_tol=-1.0)
    os.environ["SWNE_TC_FORCE_WIDE"] = "1"
    try:
        rw = handle.fit(16, W0s, H0s, max_iter=8, trace=1, rel_tol=-1.0, engine="tc")
    finally:
        del os.environ["SWNE_TC_FORCE_WIDE"]
    assert rw.options["engine"] == "tc"
    _check_against(rw, refs)
    # the seed: tensor-core sketch products against the fp32 kernels (same random test matrix -> same subspace)
    W1, H1, s1 = handle.nnsvd_init(4, seed=3)
    os.environ["SWNE_NNSVD_SIMT"] = "1"
    try:
        W2, H2, s2 = handle.nnsvd_init(4, seed=3)
    finally:
        del os.environ["SWNE_NNSVD_SIMT"]
    assert np.max(np.abs(s1 - s2) / s2) < 1e-4
    assert rel_fro(W1, W2) < 1e-2 and rel_fro(H1, H2) < 1e-2
