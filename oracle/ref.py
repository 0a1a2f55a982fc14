"""numpy / ctypes front of the CPU oracle.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every function cites the reference lines it restates.  Array conventions follow R: A is genes x cells
(n x m), W is n x k, H is k x m; everything fp64.
"""
import ctypes

import numpy as np

from .build import load_oracle

TINY = 1e-16


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _pad3(v):
    """NNLM: alpha/beta shorter than 3 are zero padded: c(as.double(alpha),0,0,0)[1:3]."""
    v = np.atleast_1d(np.asarray(v, dtype=np.float64)).ravel()
    return np.concatenate([v, np.zeros(3)])[:3].copy()


def nnmf_ref(A, k, W0, H0, alpha=0.0, beta=0.0, loss="mse", max_iter=500, rel_tol=1e-4,
             inner_max_iter=None, inner_rel_tol=1e-9, trace=None, n_threads=0, kl_l2_late=False):
    """NNLM::nnmf(A, k, init=list(W=W0,H=H0), method='scd', loss=...) restated (SURVEY Appendix A;
    call sites /root/reference/R/nmf.R:117,129,167,168).  W0/H0 must be given: NNLM's random init uses
    R's RNG, which cannot be reproduced here."""
    if loss not in ("mse", "mkl"):
        raise ValueError("loss must be mse or mkl")
    method = 1 if loss == "mse" else 3
    if inner_max_iter is None:
        inner_max_iter = 50 if loss == "mse" else 1
    if trace is None:
        trace = max(1, 100 // inner_max_iter)
    if trace <= 0:
        trace = 999999
    A = np.asfortranarray(A, dtype=np.float64)
    n, m = A.shape
    W = np.asfortranarray(np.array(W0, dtype=np.float64, copy=True))
    H = np.asfortranarray(np.array(H0, dtype=np.float64, copy=True))
    assert W.shape == (n, k) and H.shape == (k, m)
    al, be = _pad3(alpha), _pad3(beta)
    L = -(-max_iter // trace) + 1
    mse = np.zeros(L); mkl = np.zeros(L); terr = np.zeros(L); epo = np.zeros(L)
    nit = ctypes.c_int(0); ntr = ctypes.c_int(0)
    lib = load_oracle()
    lib.oracle_set_variant(1 if kl_l2_late else 0)      # the second reading of scd_kl_update's L2 step (nnmf_ref.c header)
    warn = lib.nnmf_ref(_dp(A), n, m, k, _dp(W), _dp(H), _dp(al), _dp(be), int(max_iter), float(rel_tol),
                        int(inner_max_iter), float(inner_rel_tol), method, int(trace), int(n_threads),
                        _dp(mse), _dp(mkl), _dp(terr), _dp(epo), ctypes.byref(nit), ctypes.byref(ntr))
    lib.oracle_set_variant(0)
    t = ntr.value
    return dict(W=np.array(W), H=np.array(H), mse=mse[:t].copy(), mkl=mkl[:t].copy(),
                target_loss=terr[:t].copy(), average_epochs=epo[:t].copy(), n_iteration=nit.value,
                not_converged=bool(warn))


def nnlm_ref(x, y, alpha=0.0, loss="mse", init=None, max_iter=10000, rel_tol=1e-12, n_threads=0):
    """NNLM::nnlm(x, y): argmin_{B>=0} loss(y, x B); x r x k, y r x c -> coefficients k x c
    (call sites /root/reference/R/nmf.R:223,247)."""
    method = 1 if loss == "mse" else 3
    x = np.asarray(x, dtype=np.float64); y = np.asfortranarray(y, dtype=np.float64)
    r, k = x.shape
    c = y.shape[1]
    xt = np.asfortranarray(x.T)  # k x r
    coef = np.asfortranarray(np.array(init, dtype=np.float64, copy=True)) if init is not None \
        else np.asfortranarray(np.full((k, c), 0.005))
    al = _pad3(alpha)
    lib = load_oracle()
    its = lib.nnlm_ref(_dp(coef), _dp(xt), _dp(y), k, r, c, _dp(al), int(max_iter), float(rel_tol),
                       int(n_threads), method)
    return dict(coefficients=np.array(coef), n_iteration=int(its))


def nnsvd_init_ref(A, k):
    """nnsvd_init (/root/reference/R/nmf.R:12-47), with a LAPACK thin SVD instead of the
    svd(..., LINPACK=T) call that modern R rejects."""
    A = np.asarray(A, dtype=np.float64)
    n, m = A.shape
    U, S, Vt = np.linalg.svd(A, full_matrices=False)
    U = U[:, :k]; S = S[:k]; V = Vt[:k, :].T
    W = np.zeros((n, k)); H = np.zeros((k, m))
    W[:, 0] = np.sqrt(S[0]) * np.abs(U[:, 0])          # :24
    H[0, :] = np.sqrt(S[0]) * np.abs(V[:, 0])          # :25
    for i in range(1, k):                               # :28-44
        uu, vv = U[:, i], V[:, i]
        uup, uun = np.where(uu >= 0, uu, 0.0), np.where(uu < 0, -uu, 0.0)
        vvp, vvn = np.where(vv >= 0, vv, 0.0), np.where(vv < 0, -vv, 0.0)
        n_uup, n_vvp = np.linalg.norm(uup), np.linalg.norm(vvp)
        n_uun, n_vvn = np.linalg.norm(uun), np.linalg.norm(vvn)
        termp, termn = n_uup * n_vvp, n_uun * n_vvn
        if termp >= termn:
            W[:, i] = np.sqrt(S[i] * termp) * uup / n_uup
            H[i, :] = np.sqrt(S[i] * termp) * vvp / n_vvp
        else:
            W[:, i] = np.sqrt(S[i] * termn) * uun / n_uun
            H[i, :] = np.sqrt(S[i] * termn) * vvn / n_vvn
    return W, H


def icafast_ref(X, nc, maxit=100, tol=1e-6):
    """ica::icafast(X, nc, center=TRUE, alg='par', fun='logcosh', alpha=1) restated from the published
    algorithm (call sites /root/reference/R/nmf.R:59,63).  X is nobs x nvars.
    Returns S (nobs x nc), M (nvars x nc)."""
    X = np.asarray(X, dtype=np.float64)
    nobs = X.shape[0]
    X = X - X.mean(axis=0, keepdims=True)
    vals, vecs = np.linalg.eigh(X.T @ X / nobs)
    order = np.argsort(vals)[::-1]
    vals, vecs = vals[order][:nc], vecs[:, order][:, :nc]
    Mprt = np.sqrt(vals)[:, None] * vecs.T              # nc x nvars
    Pmat = vecs / np.sqrt(vals)[None, :]                # nvars x nc
    Xw = X @ Pmat                                       # whitened
    R = np.eye(nc)
    it, vtol = 0, 1.0
    while vtol > tol and it < maxit:
        S = Xw @ R
        g = np.tanh(S)
        Rn = Xw.T @ g / nobs - R * np.mean(1.0 - g * g, axis=0)[None, :]
        u, _, vt = np.linalg.svd(Rn)
        Rn = u @ vt
        vtol = 1.0 - np.min(np.abs(np.sum(R * Rn, axis=0)))
        R = Rn
        it += 1
    M = R.T @ Mprt                                      # nc x nvars
    vafs = np.sum(M * M, axis=1)
    ix = np.argsort(-vafs, kind="stable")
    M = M[ix]; R = R[:, ix]
    return Xw @ R, M.T


def ica_init_ref(A, k, ica_fast=False):
    """ica_init (/root/reference/R/nmf.R:56-66).  The ica.fast=T branch calls irlba (randomised start),
    restated with an exact truncated SVD of the row-centred t(A)."""
    A = np.asarray(A, dtype=np.float64)
    if ica_fast:
        Ac = A - A.mean(axis=1, keepdims=True)
        u, _, _ = np.linalg.svd(Ac.T, full_matrices=False)
        S, _ = icafast_ref(u[:, :50], k, maxit=50, tol=1e-4)
        return Ac @ S, S.T
    S, M = icafast_ref(A.T, k, maxit=50, tol=1e-4)
    return M, S.T


def zero_replace_ref(W, H, A_mean, rng, zero_eps=1e-6):
    """/root/reference/R/nmf.R:121-127 with numpy's generator standing in for R's runif."""
    W = np.array(W, dtype=np.float64, copy=True); H = np.array(H, dtype=np.float64, copy=True)
    W[W < zero_eps] = 0.0; H[H < zero_eps] = 0.0
    zw = W == 0; zh = H == 0
    W[zw] = rng.uniform(0.0, A_mean / 100.0, size=int(zw.sum()))
    H[zh] = rng.uniform(0.0, A_mean / 100.0, size=int(zh.sum()))
    return W, H


def run_nmf_ref(A, k, alpha=0.0, init="nnsvd", loss="mse", max_iter=500, ica_fast=False, seed=0,
                n_threads=0, **kw):
    """RunNMF (/root/reference/R/nmf.R:96-134) on the oracle.  'random' init draws U(0,1)*0.01 from numpy."""
    A = np.asarray(A, dtype=np.float64)
    if np.any(A < 0):
        raise ValueError("The input matrix contains negative elements !")
    if k < 3:
        raise ValueError("k must be greater than or equal to 3 to create a viable SWNE plot")
    if init not in ("ica", "nnsvd", "random"):
        raise ValueError("Invalid initialization method")
    rng = np.random.default_rng(seed)
    n, m = A.shape
    if init == "random":
        W0 = rng.uniform(size=(n, k)) * 0.01; H0 = rng.uniform(size=(k, m)) * 0.01
    else:
        W0, H0 = ica_init_ref(A, k, ica_fast) if init == "ica" else nnsvd_init_ref(A, k)
        W0, H0 = zero_replace_ref(W0, H0, A.mean(), rng)
    res = nnmf_ref(A, k, W0, H0, alpha=alpha, loss=loss, max_iter=max_iter, n_threads=n_threads, **kw)
    res["W0"], res["H0"] = W0, H0
    return res


def kl_div_ref(x, y, pseudocount=1e-12):
    """kl_div (/root/reference/R/nmf.R:71-75)."""
    x = x + pseudocount; y = y + pseudocount
    return x * np.log(x / y) - x + y


def find_num_factors_ref(A, k_range, inits, A_rand, loss="mse", max_iter=250, n_threads=0):
    """FindNumFactors (/root/reference/R/nmf.R:157-205).  inits[k] = (W0, H0, W0r, H0r): the random
    inits NNLM would draw from R's RNG are injected; A_rand is the permuted matrix (:165).
    Fits always use mse; `loss` only picks the error metric."""
    if loss not in ("mse", "mkl"):
        raise ValueError("Invalid loss function")
    if len(k_range) < 3:
        raise ValueError("k.range sequence must have at least 3 values")
    A = np.asarray(A, dtype=np.float64); A_rand = np.asarray(A_rand, dtype=np.float64)
    k_err = np.zeros((2, len(k_range)))
    for c, k in enumerate(k_range):
        W0, H0, W0r, H0r = inits[k]
        z = nnmf_ref(A, k, W0, H0, max_iter=max_iter, n_threads=n_threads)
        zr = nnmf_ref(A_rand, k, W0r, H0r, max_iter=max_iter, n_threads=n_threads)
        Ah, Ahr = z["W"] @ z["H"], zr["W"] @ zr["H"]
        if loss == "mse":
            k_err[0, c] = np.mean((Ah - A) ** 2); k_err[1, c] = np.mean((Ahr - A_rand) ** 2)
        else:
            k_err[0, c] = np.mean(kl_div_ref(Ah, A)); k_err[1, c] = np.mean(kl_div_ref(Ahr, A_rand))
    return pick_num_factors_ref(k_err, k_range)


def pick_num_factors_ref(k_err, k_range):
    """/root/reference/R/nmf.R:186-198."""
    err_del = k_err[:, :-1] - k_err[:, 1:]
    diff = err_del[0] - err_del[1]
    neg = np.nonzero(diff < 0)[0]
    min_idx = int(neg[0]) if neg.size else int(np.argmin(diff)) + 1   # 0-based into k_range
    return dict(err=diff, k=k_range[min_idx], k_err=k_err)


def get_sample_coords_ref(H, H_coords, alpha=1.0, n_pull=None):
    """get_sample_coords (/root/reference/R/swne_plotting.R:70-84): per cell, the n_pull largest
    factors (R's order(decreasing=T): stable, ties keep the lower index first)."""
    H = np.asarray(H, dtype=np.float64); H_coords = np.asarray(H_coords, dtype=np.float64)
    k, m = H.shape
    if n_pull is not None and n_pull < 3:
        n_pull = 3
    if n_pull is None or n_pull > k:
        n_pull = k
    out = np.zeros((m, 2))
    for j in range(m):
        ix = np.argsort(-H[:, j], kind="stable")[:n_pull]
        w = H[ix, j] ** alpha
        s = w.sum()
        out[j, 0] = np.sum(H_coords[ix, 0] * (w / s))
        out[j, 1] = np.sum(H_coords[ix, 1] * (w / s))
    return out


def factor_distance_ref(H, distance="cosine"):
    """k x k factor distance of get_factor_coords (/root/reference/R/swne_plotting.R:43-52); 'IC' is
    excluded (randomised KDE mutual information, /root/reference/R/annotate_nmf.R:302-330)."""
    H = np.asarray(H, dtype=np.float64)
    d = distance.lower()
    if d in ("cor", "correlation"):
        d = "pearson"
    if d == "pearson":
        return np.sqrt(np.maximum(2.0 * (1.0 - np.corrcoef(H)), 0.0))
    if d == "cosine":
        nrm = np.linalg.norm(H, axis=1)
        return 1.0 - (H @ H.T) / np.outer(nrm, nrm)
    if d == "euclidean":
        sq = np.sum(H * H, axis=1)
        return np.sqrt(np.maximum(sq[:, None] + sq[None, :] - 2.0 * H @ H.T, 0.0))
    raise ValueError("Distance must be one of: pearson, IC, cosine, euclidean")


def _cmdscale(D, kdim=2):
    n = D.shape[0]
    J = np.eye(n) - np.ones((n, n)) / n
    B = -0.5 * J @ (D * D) @ J
    vals, vecs = np.linalg.eigh(B)
    order = np.argsort(vals)[::-1][:kdim]
    return vecs[:, order] * np.sqrt(np.maximum(vals[order], 0.0))[None, :]


def sammon_ref(D, kdim=2, niter=250, magic=0.2, tol=1e-4):
    """MASS::sammon(d, k=2, niter=250) (called at /root/reference/R/swne_plotting.R:53) restated from
    MASS's published VR_sammon: cmdscale start, diagonal-Newton steps with step halving."""
    D = np.asarray(D, dtype=np.float64)
    nn = D.shape[0]
    Y = _cmdscale(D, kdim)
    xu = Y.copy()

    def stress(P):
        e = tot = 0.0
        for j in range(1, nn):
            for q in range(j):
                d = D[j, q]; tot += d
                ee = np.sqrt(np.sum((P[j] - P[q]) ** 2))
                e += (d - ee) ** 2 / d
        return e / tot

    eprev = epast = stress(Y)
    i = 1
    while i <= niter:
        while True:
            for j in range(nn):
                xv1 = np.zeros(kdim); xv2 = np.zeros(kdim)
                for q in range(nn):
                    if q == j:
                        continue
                    dt = D[q, j]
                    xd = xu[j] - xu[q]
                    dpj = np.sqrt(np.sum(xd * xd))
                    dq = dt - dpj; dr = dt * dpj
                    xv1 += xd * dq / dr
                    xv2 += (dq - xd * xd * (1.0 + dq / dpj) / dpj) / dr
                Y[j] = xu[j] + magic * xv1 / np.abs(xv2)
            e = stress(Y)
            if e > eprev:
                e = eprev
                magic *= 0.2
                if magic > 1.0e-3:
                    continue
                return xu
            break
        magic = min(magic * 1.5, 0.5)
        eprev = e
        xu = Y - Y.mean(axis=0, keepdims=True)
        if i % 10 == 0:
            if e > epast - tol:
                break
            epast = e
        i += 1
    return xu


def get_factor_coords_ref(H, distance="cosine"):
    """get_factor_coords (/root/reference/R/swne_plotting.R:33-64): distance -> sammon -> min-max."""
    P = sammon_ref(factor_distance_ref(H, distance))
    return (P - P.min(axis=0)) / (P.max(axis=0) - P.min(axis=0))
