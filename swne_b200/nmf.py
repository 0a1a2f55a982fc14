"""Host-side mirror of SWNE's NMF front-end (/root/reference/R/nmf.R) over the C-ABI.

Same names, argument meaning and error behaviour as the R functions so the parity tests read like tests
of the reference: RunNMF (R/nmf.R:96-134), FindNumFactors (:157-205), nnsvd_init (:12-47), ica_init
(:56-66), kl_div (:71-75), ProjectFeatures (:222-225), ProjectSamples (:241-249), and nnmf / nnlm for the
NNLM calls themselves.  All numerics run in libswne_nmf.so on the GPU; this file
only validates, seeds and marshals.  Matrices follow R: A is genes x cells (n x m), W n x k, H k x m.
"""
import ctypes
import time

import numpy as np

from . import _capi as C


def _is_csc(A):
    return hasattr(A, "indptr") and hasattr(A, "indices") and hasattr(A, "data") and getattr(A, "format", "") == "csc"


class NMFHandle:
    """One GPU's resident matrix + workspaces (swne_nmf_handle_t).  Reuse it across fits (FindNumFactors)."""

    def __init__(self, device=0):
        self._lib = C.load()
        self._h = ctypes.c_void_p()
        C.check(self._lib.swne_nmf_create(ctypes.byref(self._h), int(device)))
        self.device = device
        self.n = self.m = 0
        self.rank, self.nranks = 0, 1
        self._keep = None

    def close(self):
        if self._h:
            self._lib.swne_nmf_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- communicator (cells sharded across ranks)
    @staticmethod
    def unique_id():
        buf = ctypes.create_string_buffer(128)
        C.check(C.load().swne_nmf_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id, rank, nranks):
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        C.check(self._lib.swne_nmf_comm_init(self._h, buf, int(rank), int(nranks)))
        self.rank, self.nranks = rank, nranks

    # ---- matrix
    def comm_info(self):
        """rank, number of ranks, and whether the fused engine's per-iteration sums go over NVLink peer memory (else NCCL)."""
        r, n, x = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        C.check(self._lib.swne_nmf_comm_info(self._h, ctypes.byref(r), ctypes.byref(n), ctypes.byref(x)))
        return dict(rank=r.value, nranks=n.value, peer_exchange=bool(x.value))

    def set_matrix(self, A):
        """A: numpy (n x m, any float dtype), scipy.sparse CSC (dgCMatrix analogue), or a torch CUDA
        float32 tensor of shape (m, n) (cell-major == R column-major n x m) that is used in place."""
        if not (hasattr(A, "data_ptr") and hasattr(A, "is_cuda")):
            self._keep = None       # a previously borrowed device tensor is released once the library owns a copy again
        if _is_csc(A):
            n, m = A.shape
            p = np.ascontiguousarray(A.indptr, dtype=np.int32)
            i = np.ascontiguousarray(A.indices, dtype=np.int32)
            if A.data.dtype == np.float32:
                x = np.ascontiguousarray(A.data)
                C.check(self._lib.swne_nmf_set_matrix_csc_f32(self._h, C.iptr(p), C.iptr(i), C.fptr(x), n, m))
            else:
                x = np.ascontiguousarray(A.data, dtype=np.float64)
                C.check(self._lib.swne_nmf_set_matrix_csc_f64(self._h, C.iptr(p), C.iptr(i), C.dptr(x), n, m))
        elif hasattr(A, "data_ptr") and hasattr(A, "is_cuda"):
            if not A.is_cuda or A.dim() != 2 or str(A.dtype) != "torch.float32" or not A.is_contiguous():
                raise ValueError("device matrix must be a contiguous CUDA float32 tensor of shape (cells, genes)")
            m, n = A.shape
            self._keep = A  # borrowed by the library
            C.check(self._lib.swne_nmf_set_matrix_dense_f32_device(self._h, ctypes.c_void_p(A.data_ptr()), n, m))
        else:
            A = np.asarray(A)
            if A.ndim != 2:
                raise ValueError("A must be a matrix")
            n, m = A.shape
            if A.dtype == np.float32:
                Af = np.asfortranarray(A)
                C.check(self._lib.swne_nmf_set_matrix_dense_f32(self._h, C.fptr(Af), n, m))
            else:
                Af = np.asfortranarray(A, dtype=np.float64)
                C.check(self._lib.swne_nmf_set_matrix_dense_f64(self._h, C.dptr(Af), n, m))
        self.n, self.m = n, m

    def set_matrix_counts(self, counts, depth=None, scale_factor=None, gene_scale=None, method="log"):
        """Raw counts (scipy CSC, genes x cells) -> ScaleCounts(method = "log" | "ft" | "none") fused into the upload
        (R/data_processing.R:275-307 without the batch correction; swne_nmf_set_matrix_csc_counts): the normalised matrix
        never exists on the host.  depth = column sums of the FULL count matrix (default: of `counts`), scale_factor =
        median depth by default, gene_scale = AdjustVariance's gsf per gene (None: adj.var = F)."""
        if not _is_csc(counts):
            raise ValueError("counts must be a scipy.sparse CSC matrix (dgCMatrix)")
        if method not in C.TRANSFORM:
            raise ValueError("method must be 'log', 'ft' or 'none'")
        n, m = counts.shape
        depth = np.asarray(counts.sum(axis=0)).ravel() if depth is None else np.asarray(depth, dtype=np.float64).ravel()
        if depth.shape != (m,):
            raise ValueError("depth must have one entry per cell")
        sf = float(np.median(depth)) if scale_factor is None else float(scale_factor)
        cs = np.ascontiguousarray(sf / depth, dtype=np.float64)
        gs = None if gene_scale is None else np.ascontiguousarray(gene_scale, dtype=np.float64)
        if gs is not None and gs.shape != (n,):
            raise ValueError("gene_scale must have one entry per gene")
        p = np.ascontiguousarray(counts.indptr, dtype=np.int32); i = np.ascontiguousarray(counts.indices, dtype=np.int32)
        x = np.ascontiguousarray(counts.data, dtype=np.float64)
        self._keep = None
        C.check(self._lib.swne_nmf_set_matrix_csc_counts(self._h, C.iptr(p), C.iptr(i), C.dptr(x), n, m, C.dptr(cs),
                                                         None if gs is None else C.dptr(gs), C.TRANSFORM[method]))
        self.n, self.m = n, m

    def matrix_info(self):
        n, m = ctypes.c_int(), ctypes.c_int()
        nnz = ctypes.c_longlong()
        s, s2 = ctypes.c_double(), ctypes.c_double()
        C.check(self._lib.swne_nmf_matrix_info(self._h, ctypes.byref(n), ctypes.byref(m), ctypes.byref(nnz),
                                               ctypes.byref(s), ctypes.byref(s2)))
        return dict(n=n.value, m=m.value, nnz=nnz.value, sum=s.value, sumsq=s2.value)

    # ---- NNLM::nnmf on the resident matrix
    def fit(self, k, W0=None, H0=None, alpha=0.0, beta=0.0, loss="mse", max_iter=500, rel_tol=1e-4, inner_max_iter=0,
            inner_rel_tol=1e-9, trace=0, verbose=0, engine="auto", full_traces=False, profile=False,
            show_warning=False, kl_l2_late=False, init="given", seed=0):
        """NNLM::nnmf on the resident matrix from the init (W0, H0), or -- init in {"random", "nnsvd", "ica", "ica_fast"} --
        RunNMF's seed computed on the device and handed to the fit without a host round trip (swne_run_nmf)."""
        if loss not in C.LOSS:
            raise ValueError("loss must be 'mse' or 'mkl'")
        if init not in C.INIT:
            raise ValueError("Invalid initialization method")
        if self.n == 0:
            raise C.SwneError(C.SWNE_ERR_NO_MATRIX, "set_matrix first")
        n, m = self.n, self.m
        if init == "given":
            # one copy each, straight into R's column-major layout (the library overwrites them with the result)
            W = np.array(W0, dtype=np.float64, order="F", copy=True)
            H = np.array(H0, dtype=np.float64, order="F", copy=True)
            if W.shape != (n, k) or H.shape != (k, m):
                raise ValueError("init shapes must be W %s and H %s" % ((n, k), (k, m)))
        else:
            W = np.empty((n, k), dtype=np.float64, order="F")
            H = np.empty((k, m), dtype=np.float64, order="F")
        p = C.Params()
        self._lib.swne_nmf_default_params(ctypes.byref(p), int(k))
        p.alpha[:] = list(C.pad3(alpha)); p.beta[:] = list(C.pad3(beta))
        p.loss = C.LOSS[loss]; p.max_iter = int(max_iter); p.rel_tol = float(rel_tol)
        p.inner_max_iter = int(inner_max_iter); p.inner_rel_tol = float(inner_rel_tol); p.trace = int(trace)
        p.verbose = int(verbose); p.engine = C.ENGINE[engine]; p.full_traces = int(full_traces)
        p.profile = int(bool(profile))
        p.nnlm_variant = C.VARIANT_KL_L2_LATE if kl_l2_late else 0
        imi = p.inner_max_iter if p.inner_max_iter > 0 else (50 if loss == "mse" else 1)
        tr = p.trace if p.trace > 0 else max(1, 100 // imi)
        L = -(-int(max_iter) // tr) + 1
        mse = np.zeros(L); mkl = np.zeros(L); terr = np.zeros(L); epo = np.zeros(L)
        res = C.Result()
        t0 = time.perf_counter()
        code = C.check(self._lib.swne_run_nmf(self._h, ctypes.byref(p), C.INIT[init], ctypes.c_uint64(int(seed) & (2**64 - 1)), C.dptr(W),
                                             C.dptr(H), C.dptr(mse), C.dptr(mkl), C.dptr(terr), C.dptr(epo), L, ctypes.byref(res)))
        wall = time.perf_counter() - t0
        if code == C.SWNE_WARN_NOT_CONVERGED and show_warning:
            import warnings
            warnings.warn("Target tolerance not reached. Try a larger max.iter.")
        t = res.n_trace
        out = NNMF(W=W, H=H, mse=mse[:t].copy(), mkl=mkl[:t].copy(), target_loss=terr[:t].copy(),
                   average_epochs=epo[:t].copy(), n_iteration=res.n_iteration, run_time=wall,
                   options=dict(method="scd", loss=loss, alpha=C.pad3(alpha), beta=C.pad3(beta), max_iter=max_iter,
                                rel_tol=rel_tol, inner_max_iter=imi, inner_rel_tol=inner_rel_tol, trace=tr,
                                engine=C.ENGINE_NAME[res.engine_used]),
                   not_converged=bool(res.not_converged),
                   stats=dict(loop_ms=res.loop_ms, upload_ms=res.upload_ms, download_ms=res.download_ms,
                              gpu_launches=res.gpu_launches, prof_ms=list(res.prof_ms),
                              prof_launches=list(res.prof_launches)))
        return out

    # ---- NNLM::nnlm on the resident matrix
    def nnlm(self, side, X, init=None, alpha=0.0, loss="mse", max_iter=10000, rel_tol=1e-12):
        """side 0: coefficients H (k x m) for fixed W = X (n x k); side 1: W (n x k) for fixed H = X (k x m)."""
        n, m = self.n, self.m
        X = np.asfortranarray(X, dtype=np.float64)
        k = X.shape[1] if side == 0 else X.shape[0]
        shape = (k, m) if side == 0 else (n, k)
        if X.shape != ((n, k) if side == 0 else (k, m)):
            raise ValueError("x has the wrong shape")
        # NNLM seeds nnlm with runif * 0.01 (R RNG); a deterministic constant start is used instead
        coef = np.asfortranarray(np.full(shape, 0.005) if init is None else np.array(init, dtype=np.float64, copy=True))
        al = C.pad3(alpha)
        sw = ctypes.c_longlong()
        C.check(self._lib.swne_nnlm(self._h, int(side), int(k), C.dptr(X), C.dptr(coef), C.dptr(al), C.LOSS[loss],
                                    int(max_iter), float(rel_tol), ctypes.byref(sw)))
        return dict(coefficients=np.array(coef), n_iteration=sw.value)

    def nnsvd_init(self, k, oversample=10, power_iters=4, seed=1):
        n, m = self.n, self.m
        W = np.zeros((n, k), order="F"); H = np.zeros((k, m), order="F"); sv = np.zeros(k)
        C.check(self._lib.swne_nnsvd_init(self._h, int(k), int(oversample), int(power_iters), int(seed), C.dptr(W),
                                          C.dptr(H), C.dptr(sv)))
        return W, H, sv

    def ica_init(self, k, ica_fast=False, seed=1):
        """ica_init (R/nmf.R:56-66) of the resident matrix on the device: signed W (n x k), H (k x m)."""
        n, m = self.n, self.m
        W = np.zeros((n, k), order="F"); H = np.zeros((k, m), order="F")
        C.check(self._lib.swne_ica_init(self._h, int(k), int(bool(ica_fast)), ctypes.c_uint64(int(seed)), C.dptr(W), C.dptr(H)))
        return W, H

    def find_num_factors(self, k_range, loss="mse", max_iter=250, seed=0, engine="auto", errors_only=False):
        """FindNumFactors' sweep (R/nmf.R:157-205) on the resident matrix in one library call: permutation, random starts,
        2 x len(k_range) fits and their errors all stay on the device (swne_find_num_factors).  The permutation depends on
        `seed` only and every fit's start on (seed, k) only, so a replica may sweep any subset of k.range (errors_only)."""
        ks = np.ascontiguousarray(k_range, dtype=np.int32)
        k_err = np.zeros((2, len(ks)))
        if errors_only:
            C.check(self._lib.swne_find_num_factors(self._h, C.iptr(ks), len(ks), C.LOSS[loss], int(max_iter), ctypes.c_uint64(int(seed)),
                                                    C.ENGINE[engine], C.dptr(k_err), None, None))
            return dict(k_err=k_err)
        diff = np.zeros(max(len(ks) - 1, 1)); pick = ctypes.c_int(0)
        C.check(self._lib.swne_find_num_factors(self._h, C.iptr(ks), len(ks), C.LOSS[loss], int(max_iter), ctypes.c_uint64(int(seed)),
                                                C.ENGINE[engine], C.dptr(k_err), ctypes.byref(pick), C.dptr(diff)))
        return dict(err=diff[:len(ks) - 1], k=int(pick.value), k_err=k_err)

    def recon_error(self, W, H):
        W = np.asfortranarray(W, dtype=np.float64); H = np.asfortranarray(H, dtype=np.float64)
        mse, kl = ctypes.c_double(), ctypes.c_double()
        C.check(self._lib.swne_recon_error(self._h, C.dptr(W), C.dptr(H), int(W.shape[1]), ctypes.byref(mse), ctypes.byref(kl)))
        return mse.value, kl.value

    def sample_coords(self, H, H_coords, alpha=1.0, n_pull=None):
        """get_sample_coords (R/swne_plotting.R:70-84).  H = None: the H the last fit left on the device."""
        Hc = np.asfortranarray(H_coords, dtype=np.float64)
        if H is None:
            k, m, hp = Hc.shape[0], self.m, None
        else:
            H = np.asfortranarray(H, dtype=np.float64)
            (k, m), hp = H.shape, C.dptr(H)
        out = np.zeros((m, 2), order="F")
        C.check(self._lib.swne_sample_coords(self._h, hp, k, m, C.dptr(Hc), float(alpha),
                                             0 if n_pull is None else int(n_pull), C.dptr(out)))
        return out

    def factor_distance(self, H, distance="cosine", k=None):
        """the k x k distance block of get_factor_coords (R/swne_plotting.R:43-52).  H = None (with k): the H the last fit left
        on the device."""
        d = distance.lower()
        if d in ("cor", "correlation"):
            d = "pearson"
        if d in ("mutual", "information", "mutual information", "ic"):
            raise C.SwneError(C.SWNE_ERR_UNSUPPORTED, "IC (randomised KDE mutual information) is not on the device path")
        if d not in ("pearson", "cosine", "euclidean"):
            raise ValueError("Distance must be one of: pearson, IC, cosine, euclidean")
        if H is None:
            if k is None:
                raise ValueError("k is needed with H = None")
            k, m, hp = int(k), self.m, None
        else:
            H = np.asfortranarray(H, dtype=np.float64)
            (k, m), hp = H.shape, C.dptr(H)
        out = np.zeros((k, k), order="F")
        C.check(self._lib.swne_factor_distance(self._h, hp, k, m, {"cosine": 0, "pearson": 1, "euclidean": 2}[d],
                                               C.dptr(out)))
        return out


class NNMF(dict):
    """The 'nnmf' list NNLM returns: W, H, mse, mkl, target.loss, average.epochs, n.iteration, run.time, options."""

    __getattr__ = dict.__getitem__


_default = {}


def default_handle(device=0):
    if device not in _default:
        _default[device] = NMFHandle(device)
    return _default[device]


# ------------------------------------------------------------------------------------------------
# seeding
# ------------------------------------------------------------------------------------------------
def nnsvd_init(A, k, handle=None, **kw):
    """nnsvd_init (R/nmf.R:12-47): NNDSVD of A from a GPU randomized truncated SVD (the R code's
    svd(A, k, k, LINPACK=T) is rejected by modern R; NNDSVD is invariant to the sign of singular pairs)."""
    h = handle or default_handle()
    if handle is None or h.n == 0:
        h.set_matrix(A)
    W, H, _ = h.nnsvd_init(k, **kw)
    return dict(W=W, H=H)


def ica_init(A, k, ica_fast=False, handle=None, seed=1):
    """ica_init (R/nmf.R:56-66) on the device: centred randomized SVD for the whitening + the FastICA fixed point
    (swne_ica_init).  Components are defined up to sign and order (as LAPACK's eigenvector signs are in the reference)."""
    h = handle or default_handle()
    if A is not None:
        h.set_matrix(A)
    W, H = h.ica_init(k, ica_fast=ica_fast, seed=seed)
    return dict(W=W, H=H)


def zero_replace(init, A_mean, rng, zero_eps=1e-6):
    """R/nmf.R:121-127: entries < 1e-6 -> 0, zeros -> runif(0, mean(A)/100).  The thresholding runs in the
    C-ABI (swne_zero_replace); the uniform draws come from the caller's generator, as R's come from R."""
    lib = C.load()
    out = {}
    for key in ("W", "H"):
        x = np.asfortranarray(np.array(init[key], dtype=np.float64, copy=True))
        nz = ctypes.c_size_t()
        C.check(lib.swne_zero_replace(C.dptr(x), x.size, float(zero_eps), None, 0, ctypes.byref(nz)))
        fill = np.ascontiguousarray(rng.uniform(0.0, A_mean / 100.0, size=nz.value))
        C.check(lib.swne_zero_replace(C.dptr(x), x.size, float(zero_eps), C.dptr(fill), fill.size, ctypes.byref(nz)))
        out[key] = x
    return out


# ------------------------------------------------------------------------------------------------
# the R front-end
# ------------------------------------------------------------------------------------------------
def nnmf(A, k, alpha=0.0, beta=0.0, init=None, loss="mse", max_iter=500, rel_tol=1e-4, n_threads=1, verbose=0,
         inner_max_iter=0, inner_rel_tol=1e-9, trace=0, seed=None, handle=None, **kw):
    """NNLM::nnmf(A, k, alpha, beta, method='scd', loss, init, max.iter, rel.tol, ...).  n_threads is
    accepted and ignored.  Without init, W and H start at U(0,1)*0.01 as in NNLM (numpy generator)."""
    h = handle or default_handle()
    if A is not None:
        h.set_matrix(A)
    n, m = h.n, h.m
    if init is None:
        rng = np.random.default_rng(seed)
        W0 = rng.uniform(size=(n, k)) * 0.01; H0 = rng.uniform(size=(k, m)) * 0.01
    else:
        W0, H0 = init["W"], init["H"]
    return h.fit(k, W0, H0, alpha=alpha, beta=beta, loss=loss, max_iter=max_iter, rel_tol=rel_tol,
                 inner_max_iter=inner_max_iter, inner_rel_tol=inner_rel_tol, trace=trace, verbose=verbose, **kw)


def RunNMF(A, k, alpha=0, init="ica", n_cores=1, loss="mse", max_iter=500, ica_fast=False, seed=None, handle=None,
           **kw):
    """RunNMF(A, k, alpha, init, n.cores, loss, max.iter, ica.fast) -- R/nmf.R:96-134.  Same checks and
    error messages; returns the nnmf list with factor names factor_1..factor_k."""
    # A: numpy (genes x cells), scipy CSC (dgCMatrix), or a CUDA float32 tensor of shape (cells, genes) -- the transpose of
    # the numpy orientation, because that is R's column-major genes x cells matrix as it sits in memory (set_matrix)
    if _is_csc(A):
        amin = A.data.min() if A.data.size else 0
    elif hasattr(A, "is_cuda"):
        amin = float(A.min()) if A.numel() else 0
    else:
        amin = np.asarray(A).min()
    if amin < 0:
        raise ValueError("The input matrix contains negative elements !")
    if k < 3:
        raise ValueError("k must be greater than or equal to 3 to create a viable SWNE plot")
    if init not in ("ica", "nnsvd", "random"):
        raise ValueError("Invalid initialization method")
    if loss not in ("mse", "mkl"):
        raise ValueError("loss must be 'mse' or 'mkl'")
    h = handle or default_handle()
    h.set_matrix(A)
    rng = np.random.default_rng(seed)
    # the seed (R/nmf.R:108-127: ica_init / nnsvd_init / NNLM's random start, then the zero replacement) is computed on the
    # device inside the same library call as the fit
    mode = {"random": "random", "nnsvd": "nnsvd", "ica": "ica_fast" if ica_fast else "ica"}[init]
    res = h.fit(k, alpha=alpha, loss=loss, max_iter=max_iter, init=mode, seed=int(rng.integers(1, 2**62)), **kw)
    res["factor_names"] = ["factor_%d" % (i + 1) for i in range(k)]
    return res


def kl_div(x, y, pseudocount=1e-12):
    """kl_div, R/nmf.R:71-75."""
    x = x + pseudocount; y = y + pseudocount
    return x * np.log(x / y) - x + y


def _permuted_entries(A, rng):
    """matrix(A[sample(length(A))], nrow(A)) of R/nmf.R:163.  Shuffling 6e7 entries takes numpy ~2 s, more than
    the 78 fits of the sweep; with torch + CUDA present the permutation is drawn and applied on the device and
    the result handed to set_matrix in place (cell-major float32)."""
    n, m = A.shape
    try:
        import torch
        on_gpu = torch.cuda.is_available()
    except ImportError:
        on_gpu = False
    if not on_gpu:
        return rng.permutation(A.ravel(order="F")).reshape((n, m), order="F")
    d = torch.from_numpy(np.ascontiguousarray(A.T, dtype=np.float32)).cuda()      # (cells, genes) == column-major A
    g = torch.Generator(device="cuda")
    g.manual_seed(int(rng.integers(0, 2**31 - 1)))
    return d.flatten()[torch.randperm(d.numel(), device="cuda", generator=g)].reshape(m, n).contiguous()


def FindNumFactors(A, k_range=(2, 4, 6, 8, 10, 12), n_cores=1, do_plot=False, seed=None, loss="mse", max_iter=250,
                   inits=None, A_rand=None, handle=None, **kw):
    """FindNumFactors -- R/nmf.R:157-205 including its quirks: the fits always use mse and random init
    (SURVEY fact 6); `loss` only selects the error metric.  `inits[k] = (W0, H0, W0r, H0r)` and `A_rand`
    may be injected (parity tests); otherwise numpy draws them (R's sample()/runif cannot be reproduced)."""
    if loss not in ("mse", "mkl"):
        raise ValueError("Invalid loss function")
    k_range = list(k_range)
    if len(k_range) < 3:
        raise ValueError("k.range sequence must have at least 3 values")
    A = np.asarray(A.toarray() if _is_csc(A) else A)
    n, m = A.shape
    if m > 20000:
        print("Warning: This function can be slow for very large datasets")
    rng = np.random.default_rng(seed)
    h = handle or default_handle()
    if inits is None and A_rand is None and not kw:
        # the whole sweep in the library: the permuted matrix, the random starts and the factors never leave the device
        h.set_matrix(A)
        return h.find_num_factors(k_range, loss=loss, max_iter=max_iter, seed=int(rng.integers(1 << 62)))
    if A_rand is None:
        A_rand = _permuted_entries(A, rng)
    k_err = np.zeros((2, len(k_range)))
    for which, mat in enumerate((A, A_rand)):
        h.set_matrix(mat)
        for c, k in enumerate(k_range):
            if inits is not None:
                W0, H0 = inits[k][2 * which], inits[k][2 * which + 1]
            else:
                W0 = rng.uniform(size=(n, k)) * 0.01; H0 = rng.uniform(size=(k, m)) * 0.01
            z = h.fit(k, W0, H0, loss="mse", max_iter=max_iter, **kw)
            mse, kl = h.recon_error(z.W, z.H)
            k_err[which, c] = mse if loss == "mse" else kl
    err_del = k_err[:, :-1] - k_err[:, 1:]
    diff = err_del[0] - err_del[1]
    neg = np.nonzero(diff < 0)[0]
    min_idx = int(neg[0]) if neg.size else int(np.argmin(diff)) + 1
    return dict(err=diff, k=k_range[min_idx], k_err=k_err)


def ProjectFeatures(newdata, H, alpha=(0, 0, 0), loss="mse", n_cores=1, handle=None):
    """ProjectFeatures (R/nmf.R:222-225): nnlm(t(H), t(newdata)) -> W (features x factors) for newdata."""
    h = handle or default_handle()
    h.set_matrix(newdata)
    return h.nnlm(1, H, alpha=alpha, loss=loss)["coefficients"]


def ProjectSamples(newdata, W, alpha=(0, 0, 0), loss="mse", n_cores=1, handle=None):
    """ProjectSamples (R/nmf.R:241-249) with features.use already applied by the caller."""
    h = handle or default_handle()
    h.set_matrix(newdata)
    return h.nnlm(0, W, alpha=alpha, loss=loss)["coefficients"]
