"""Consumers of H: get_factor_coords / get_sample_coords / EmbedSWNE (/root/reference/R/swne_plotting.R:33-145).

Device work: the k x k factor distance (a Gram over all cells) and the per-cell top-n_pull weighted mean
(the only O(cells) interpreted loop on the reference path).  The Sammon map of k points stays on the host.
"""
import numpy as np

from .nmf import default_handle


def _cmdscale(D, kdim=2):
    n = D.shape[0]
    J = np.eye(n) - 1.0 / n
    vals, vecs = np.linalg.eigh(-0.5 * J @ (D * D) @ J)
    o = np.argsort(vals)[::-1][:kdim]
    return vecs[:, o] * np.sqrt(np.maximum(vals[o], 0.0))[None, :]


def sammon(D, kdim=2, niter=250, magic=0.2, tol=1e-4):
    """MASS::sammon(d, k = 2, niter = 250) as called at R/swne_plotting.R:53 (vectorised over pairs)."""
    D = np.asarray(D, dtype=np.float64)
    nn = D.shape[0]
    iu = np.triu_indices(nn, 1)
    off = ~np.eye(nn, dtype=bool)
    tot = D[iu].sum()

    def stress(P):
        d = np.sqrt(((P[:, None, :] - P[None, :, :]) ** 2).sum(-1))
        return (((D[iu] - d[iu]) ** 2) / D[iu]).sum() / tot

    Y = _cmdscale(D, kdim)
    xu = Y.copy()
    eprev = epast = stress(Y)
    i = 1
    while i <= niter:
        while True:
            Y = xu.copy()
            for j in range(nn):                       # Gauss-Seidel over points, as MASS does (xu is fixed)
                xd = xu[j][None, :] - xu              # nn x kdim
                dpj = np.sqrt((xd * xd).sum(1))
                msk = off[j]
                dt = D[j][msk]; dp = dpj[msk]; x = xd[msk]
                dq = dt - dp; dr = dt * dp
                xv1 = (x * (dq / dr)[:, None]).sum(0)
                xv2 = ((dq[:, None] - x * x * ((1.0 + dq / dp) / dp)[:, None]) / dr[:, None]).sum(0)
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


def get_factor_coords(H, method="sammon", distance="cosine", handle=None):
    """get_factor_coords (R/swne_plotting.R:33-64): distance on the GPU, sammon on the host, min-max."""
    if method != "sammon":
        raise ValueError("Invalid factor projection method")
    h = handle or default_handle()
    D = h.factor_distance(H, distance)
    P = sammon(D)
    return (P - P.min(axis=0)) / (P.max(axis=0) - P.min(axis=0))


def get_sample_coords(H, H_coords, alpha=1.0, n_pull=None, handle=None):
    """get_sample_coords (R/swne_plotting.R:70-84) on the GPU."""
    h = handle or default_handle()
    return h.sample_coords(H, H_coords, alpha=alpha, n_pull=n_pull)


def EmbedSWNE(H, SNN=None, alpha_exp=1.0, snn_exp=1.0, n_pull=None, proj_method="sammon", dist_use="cosine",
              snn_factor_proj=True, handle=None):
    """EmbedSWNE (R/swne_plotting.R:102-145).  SNN: optional scipy.sparse cells x cells similarity."""
    H = np.asarray(H, dtype=np.float64)
    if SNN is not None and SNN.shape != (H.shape[1], H.shape[1]):
        raise ValueError("Check the dimensions of your SNN. nrow(H) must equal nrow(SNN) and ncol(SNN)")
    ix = H.sum(axis=0) > 0
    H = H[:, ix]
    if SNN is not None:
        SNN = SNN.tocsr()[ix][:, ix].astype(np.float64)
        SNN.data = SNN.data ** snn_exp
        rs = np.asarray(SNN.sum(axis=1)).ravel()
        SNN = SNN.multiply(1.0 / rs[:, None]).tocsr()
    Hf = (SNN @ H.T).T if (SNN is not None and snn_factor_proj) else H
    H_coords = get_factor_coords(Hf, method=proj_method, distance=dist_use, handle=handle)
    sample_coords = get_sample_coords(H, H_coords, alpha=alpha_exp, n_pull=n_pull, handle=handle)
    if SNN is not None:
        sample_coords = np.asarray(SNN @ sample_coords)
    return dict(H_coords=H_coords, sample_coords=sample_coords, kept=ix)
