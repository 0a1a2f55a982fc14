"""Multi-GPU plumbing: one process per GPU, cells (columns of A and H) sharded across ranks, the
ncclUniqueId of the library's communicator distributed through torch.distributed.
torch is used only for the rendezvous/broadcast; the data-path all-reduce is issued by libswne_nmf.so."""
import numpy as np


def shard_bounds(m, nranks):
    """contiguous, near-equal cell ranges: rank r owns [b[r], b[r+1])."""
    base, rem = divmod(m, nranks)
    b = [0]
    for r in range(nranks):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b


def shard_bounds_nnz(indptr, nranks):
    """equal-nnz contiguous column ranges for CSC input."""
    indptr = np.asarray(indptr, dtype=np.int64)
    m = indptr.size - 1
    total = indptr[-1]
    b = [0]
    for r in range(1, nranks):
        target = total * r // nranks
        b.append(int(min(max(np.searchsorted(indptr, target, side="left"), b[-1]), m)))
    b.append(m)
    return b


def init_comm(handle, group=None):
    """Creates the library's NCCL communicator over the default torch.distributed group."""
    import torch
    import torch.distributed as td
    rank, world = td.get_rank(group), td.get_world_size(group)
    if world == 1:
        return
    obj = [handle.unique_id() if rank == 0 else None]
    td.broadcast_object_list(obj, src=0, group=group)
    handle.comm_init(obj[0], rank, world)


def gather_H(H_local, group=None):
    """all ranks' H shards concatenated along cells (host side, via torch.distributed)."""
    import torch.distributed as td
    world = td.get_world_size(group)
    if world == 1:
        return H_local
    parts = [None] * world
    td.all_gather_object(parts, np.ascontiguousarray(H_local), group=group)
    return np.concatenate(parts, axis=1)


def find_num_factors_replicas(handle, A, k_range, loss="mse", max_iter=250, seed=0, group=None):
    """FindNumFactors (R/nmf.R:157-205) over the ranks as replicas: its 2 x len(k_range) fits are independent, so every
    rank holds the whole matrix on its own (communicator-free) handle and sweeps a strided subset of k_range; the permuted
    matrix and the random starts are functions of (seed, k) only, so the errors are those of the single-GPU sweep.  No data-path collective; k_err is gathered on the host."""
    import torch.distributed as td
    rank, world = td.get_rank(group), td.get_world_size(group)
    ks = [int(k) for k in k_range]
    if len(ks) < 3:
        raise ValueError("k.range sequence must have at least 3 values")
    handle.set_matrix(A)
    k_err = np.zeros((2, len(ks)))
    mine = list(range(rank, len(ks), world))
    if mine:
        r = handle.find_num_factors([ks[c] for c in mine], loss=loss, max_iter=max_iter, seed=seed, errors_only=True)
        k_err[:, mine] = r["k_err"]
    parts = [None] * world
    td.all_gather_object(parts, k_err, group=group)
    k_err = np.sum(parts, axis=0)
    err_del = k_err[:, :-1] - k_err[:, 1:]
    diff = err_del[0] - err_del[1]
    neg = np.nonzero(diff < 0)[0]
    return dict(err=diff, k=ks[int(neg[0]) if neg.size else int(np.argmin(diff)) + 1], k_err=k_err)
