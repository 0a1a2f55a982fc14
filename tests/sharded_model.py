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
