"""Host model (numpy + torch.distributed) of the cell-sharded mse schedule the CUDA path runs per rank
(SURVEY.md section 8e): used by the world_size-2 gloo test to show that sharding cells and all-reducing
only A H' (n x k), H H' (k x k) and the loss partial reproduces the unsharded iteration."""
import numpy as np
import torch
import torch.distributed as td

TINY = 1e-16


def _allreduce(x):
    t = torch.from_numpy(np.ascontiguousarray(x))
    td.all_reduce(t)
    return t.numpy()


def _scd(h, G, mu, max_iter=50, tol=1e-9):
    for _ in range(max_iter):
        rel = 0.0
        for kk in range(len(h)):
            tmp = max(h[kk] - mu[kk] / G[kk, kk], 0.0)
            if tmp != h[kk]:
                mu += (tmp - h[kk]) * G[:, kk]
                rel = max(rel, 2 * abs(h[kk] - tmp) / (tmp + h[kk] + TINY))
                h[kk] = tmp
        if rel <= tol:
            break


def sharded_nnmf_mse(A_loc, W0, H0_loc, iters, m_global):
    n = A_loc.shape[0]
    W, H = W0.copy(), H0_loc.copy()
    k = W.shape[1]
    normA = _allreduce(np.array([np.sum(A_loc * A_loc)]))[0]
    terr = []
    for _ in range(iters):
        B = _allreduce(A_loc @ H.T)                      # n x k   (the one big exchange)
        G = _allreduce(H @ H.T) + TINY * np.eye(k)
        for g in range(n):                               # replicated W update
            w = W[g].copy(); mu = G @ w - B[g]
            _scd(w, G, mu); W[g] = w
        G0 = W.T @ W
        Gh = G0 + TINY * np.eye(k)
        C = W.T @ A_loc
        part = 0.0
        for j in range(A_loc.shape[1]):                  # local H update
            h = H[:, j].copy(); mu = Gh @ h - C[:, j]
            _scd(h, Gh, mu); H[:, j] = h
            part += h @ G0 @ h - 2 * h @ C[:, j]
        part = _allreduce(np.array([part]))[0]
        terr.append(0.5 * (normA + part) / (n * m_global))
    return W, H, np.array(terr)


def _kl_cols_local(X, Y, A_cols, sumY, pen, inner, tol):
    """scd_kl_update on whole columns that live on this rank (the cell side): X k x cols in place, Y rows x k fixed."""
    k = X.shape[0]
    for j in range(A_cols.shape[1]):
        a = A_cols[:, j]; h = X[:, j].copy(); Ajt = Y @ h; sumH = h.sum()
        for _ in range(inner):
            rel = 0.0
            for kk in range(k):
                mu = Y[:, kk] / (Ajt + TINY)
                a2 = a @ (mu * mu) + pen[0]
                b = a @ mu - sumY[kk] + a2 * h[kk] - pen[2] - pen[1] * (sumH - h[kk])
                tmp = max(b / (a2 + TINY), 0.0)
                if tmp != h[kk]:
                    Ajt = Ajt + (tmp - h[kk]) * Y[:, kk]
                    rel = max(rel, 2 * abs(h[kk] - tmp) / (tmp + h[kk] + TINY))
                    sumH += tmp - h[kk]; h[kk] = tmp
            if rel <= tol:
                break
        X[:, j] = h


def sharded_nnmf_mkl(A_loc, W0, H0_loc, iters, m_global, alpha=(0, 0, 0), beta=(0, 0, 0), inner=1, tol=1e-9):
    """The cell-sharded mkl schedule of the CUDA path (swne_nmf.cu launch_kl_sharded): a gene's Newton sums run over all
    cells, so each coordinate step of the gene side is  partial sums on the rank's cells -> all-reduce of 2 n numbers ->
    the step, replicated; the A-hat update of a coordinate is applied lazily on the way into the next one; a gene leaves the
    sweeps when none of its coordinates moved by more than tol.  The cell side is local."""
    n, ml = A_loc.shape
    W, H = W0.copy(), H0_loc.copy()
    k = W.shape[1]
    klc = _allreduce(np.array([np.sum((A_loc + TINY) * np.log(A_loc + TINY) - A_loc)]))[0]
    N = n * m_global
    mkl = []
    for _ in range(iters):
        # ---- gene side (W), replicated; A-hat of the rank's cells only
        sumH_all = _allreduce(H.sum(axis=1))
        Ahat = W @ H                                     # n x ml
        sumW_row = W.sum(axis=1)
        dprev = np.zeros(n); active = np.ones(n, dtype=bool)
        for _s in range(inner):
            again = np.zeros(n, dtype=bool)
            for kk in range(k):
                kprev = (kk + k - 1) % k
                upd = active & (dprev != 0.0)
                Ahat[upd] = np.maximum(Ahat[upd] + dprev[upd, None] * H[kprev][None, :], 0.0)
                mu = H[kk][None, :] / (Ahat + TINY)
                part = np.stack([np.sum(A_loc * mu * mu, axis=1), np.sum(A_loc * mu, axis=1)], axis=1)
                part[~active] = 0.0
                part = _allreduce(part)
                for g in np.nonzero(active)[0]:
                    hk = W[g, kk]
                    a2 = part[g, 0] + alpha[0]
                    b = part[g, 1] - sumH_all[kk] + a2 * hk - alpha[2] - alpha[1] * (sumW_row[g] - hk)
                    tmp = max(b / (a2 + TINY), 0.0)
                    dprev[g] = tmp - hk
                    if tmp != hk:
                        if 2 * abs(tmp - hk) > tol * (tmp + hk + TINY):
                            again[g] = True
                        sumW_row[g] += tmp - hk; W[g, kk] = tmp
            active = active & again
            if not active.any():
                break
        # ---- cell side (H), local
        _kl_cols_local(H, W, A_loc, W.sum(axis=0), beta, inner, tol)
        Ah = W @ H
        part = _allreduce(np.array([np.sum(-A_loc * np.log(Ah + TINY) + Ah)]))[0]
        mkl.append((klc + part) / N)
    return W, H, np.array(mkl)
