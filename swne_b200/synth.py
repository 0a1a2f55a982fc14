"""Synthetic single-cell-like inputs: negative-binomial counts from a low-rank
Gamma model, depth-normalised and log1p-transformed like ScaleCounts' log path
(/root/reference/R/data_processing.R:277-307).  numpy for the small parity cases, torch on the GPU for
the bench sizes (no datasets can be fetched here)."""
import numpy as np

CONFIGS = {
    # name: n genes, m cells, k, kt (true rank), mean-count scale giving the target density, seed
    "cfg1": dict(n=2000, m=3000, k=16, kt=9, scale=0.12, seed=1),
    "cfg3": dict(n=3000, m=100000, k=30, kt=20, scale=0.19, seed=3),
    "cfg4": dict(n=3000, m=20000, k=None, kt=12, scale=0.19, seed=4),
    "cfg5": dict(n=3000, m=1300000, k=40, kt=30, scale=0.08, seed=5),
}


def make_matrix_np(n, m, kt, scale, seed, dtype=np.float32):
    """genes x cells (n x m) nonnegative matrix, R layout (Fortran order)."""
    rng = np.random.default_rng(seed)
    Wt = rng.gamma(0.3, 1.0, size=(n, kt))
    clusters = rng.integers(0, kt, size=m)
    Ht = rng.gamma(0.3, 1.0, size=(kt, m))
    Ht[clusters, np.arange(m)] += rng.gamma(2.0, 1.0, size=m)       # cluster structure
    lam = scale * (Wt @ Ht) / (0.3 * 0.3 * kt + 0.6) * 3.0
    size = 2.0
    X = rng.poisson(rng.gamma(size, lam / size))
    cs = X.sum(axis=0).astype(np.float64)
    cs[cs == 0] = 1.0
    A = np.log1p(X / cs[None, :] * np.median(cs))
    return np.asfortranarray(A.astype(dtype))


def make_matrix_torch(n, m, kt, scale, seed, device="cuda", chunk=20000):
    """same model on the GPU; returns a (m, n) float32 CUDA tensor (cell-major == R column-major n x m)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    Wt = torch._standard_gamma(torch.full((n, kt), 0.3, device=device), generator=g)
    out = torch.empty((m, n), dtype=torch.float32, device=device)
    counts_sum = torch.empty(m, dtype=torch.float32, device=device)
    for c0 in range(0, m, chunk):
        mc = min(chunk, m - c0)
        Ht = torch._standard_gamma(torch.full((mc, kt), 0.3, device=device), generator=g)
        cl = torch.randint(0, kt, (mc,), device=device, generator=g)
        Ht[torch.arange(mc, device=device), cl] += torch._standard_gamma(torch.full((mc,), 2.0, device=device), generator=g)
        lam = scale * (Ht @ Wt.T) / (0.3 * 0.3 * kt + 0.6) * 3.0
        X = torch.poisson(torch._standard_gamma(torch.full_like(lam, 2.0), generator=g) * (lam / 2.0), generator=g)
        counts_sum[c0:c0 + mc] = X.sum(dim=1)
        out[c0:c0 + mc] = X
    counts_sum[counts_sum == 0] = 1.0
    med = counts_sum.median()
    for c0 in range(0, m, chunk):
        mc = min(chunk, m - c0)
        out[c0:c0 + mc] = torch.log1p(out[c0:c0 + mc] / counts_sum[c0:c0 + mc, None] * med)
    return out


def random_init(n, m, k, seed, scale=0.01):
    """NNLM's default start: U(0,1) * 0.01."""
    rng = np.random.default_rng(seed)
    return np.asfortranarray(rng.uniform(size=(n, k)) * scale), np.asfortranarray(rng.uniform(size=(k, m)) * scale)
