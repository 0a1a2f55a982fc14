#!/usr/bin/env python
"""bench.py -- RunNMF iterations/s on BASELINE.json's configuration 3 (3 000 genes x 100 000 cells, k=30,
mse), synthetic NB-count data, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one outer NMF iteration (W update + H update + loss bookkeeping).  N > 1 shards the SAME
workload's cells across ranks (strong scaling; one NCCL all-reduce of A H', H H' and the loss partial per
iteration).  Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of NNLM::nnmf
(oracle/, "port": the real NNLM cannot be built here) on the host cores, on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "RunNMF iterations/s (3000 genes x 100000 cells, k=30, mse)"
UNIT = "iterations/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (one persistent nvidia-smi -lms process)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self._proc = None
        self._t = None
        self.t_start = None

    def _loop(self):
        for line in self._proc.stdout:
            parts = [x.strip() for x in line.strip().split(",")]
            if len(parts) >= 7:
                self.samples.append((time.perf_counter(), parts))

    def __enter__(self):
        try:
            self._proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                           "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                          stderr=subprocess.DEVNULL, text=True)
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        except Exception:
            self._proc = None
        return self

    def mark(self):
        """samples taken from now on count as 'under load'"""
        self.t_start = time.perf_counter()

    def __exit__(self, *exc):
        if self._proc is not None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=5)
            except Exception:
                self._proc.kill()

    def n_loaded(self):
        return sum(1 for t, _ in self.samples if self.t_start is not None and t >= self.t_start)

    def summary(self):
        rows = [r for t, r in self.samples if self.t_start is None or t >= self.t_start]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(rows)}


def workload_config(n, m, k):
    """the `config` both arms print: it names the workload only (what differs between the arms is said elsewhere)"""
    return {"workload": "cfg3: %d genes x %d cells, k=%d, mse, random init U(0,1)*0.01, timed from the state after the warm-up iterations" % (n, m, k),
            "genes": n, "cells": m, "k": k, "loss": "mse",
            "l2": "input (%.2f GB fp32) larger than L2 / last-level cache: no flush needed between iterations" % (4.0 * n * m / 1e9)}


def make_shards(cfg, m, world, device):
    """the synthetic matrix of the run, as the `world` cell shards the GPU arm uses (shard r: generator seed + 1000 r).
    Both arms call this, so they factorise the SAME matrix."""
    from swne_b200 import synth
    from swne_b200.dist import shard_bounds
    b = shard_bounds(m, world)
    return b, [synth.make_matrix_torch(cfg["n"], b[r + 1] - b[r], cfg["kt"], cfg["scale"], cfg["seed"] + 1000 * r, device=device)
               for r in range(world)]


def make_inits(n, b, k, world):
    """W0 (same on all ranks) and the H0 shards, as the GPU arm draws them"""
    from swne_b200 import synth
    W0 = synth.random_init(n, 1, k, 0)[0]
    H0 = [synth.random_init(n, b[r + 1] - b[r], k, 0 if world == 1 else 100 + r)[1] for r in range(world)]
    return W0, H0


def host_cores():
    """all the host threads the box offers, whatever OMP_NUM_THREADS says (torchrun sets it to 1)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_oracle_run(A64, k, W0, H0, warmup, iters, cores):
    """the CPU restatement of NNLM::nnmf (oracle/, fp64, OpenMP) on the full matrix: `warmup` untimed iterations, then
    `iters` timed ones continuing from that state (trace 2 = NNLM's default for mse, so the per-trace error block -- a
    dense W'H product in NNLM -- is part of the timed work, as in the reference)."""
    from oracle import nnmf_ref
    w = nnmf_ref(A64, k, W0, H0, max_iter=max(warmup, 1), rel_tol=-1.0, trace=2, n_threads=cores)
    t0 = time.perf_counter()
    r = nnmf_ref(A64, k, w["W"], w["H"], max_iter=iters, rel_tol=-1.0, trace=2, n_threads=cores)
    dt = time.perf_counter() - t0
    return r["n_iteration"] / dt, dt, float(r["target_loss"][-1])


def run_reference(args, cfg):
    """--impl reference: the reference's CPU implementation of the path (oracle port of NNLM::nnmf: the real NNLM is an
    un-vendored R package and cannot be built here) on ALL host cores, on the full workload and the same matrix and
    inits as the GPU arm at this N.  Rank 0 alone works; the other ranks exit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    t_all = time.perf_counter()
    world = max(1, args.gpus)
    n, m, k = cfg["n"], args.cells or cfg["m"], cfg["k"]
    device = "cuda" if torch.cuda.is_available() else "cpu"      # input generation only; the timed work is all CPU
    b, shards = make_shards(cfg, m, world, device)
    A64 = np.empty((n, m), dtype=np.float64, order="F")
    for r, sh in enumerate(shards):
        A64[:, b[r]:b[r + 1]] = sh.cpu().numpy().T
    del shards
    W0, H0s = make_inits(n, b, k, world)
    H0 = np.asfortranarray(np.concatenate(H0s, axis=1))
    cores = host_cores()
    iters = max(1, args.steps)
    val, dt, loss = cpu_oracle_run(A64, k, W0, H0, args.warmup, iters, cores)
    sample = "full workload (%d x %d, same matrix and inits as the GPU arm at N=%d), %d warm-up + %d timed iterations, %.1f s timed" % (
        n, m, world, max(args.warmup, 1), iters, dt)
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": iters, "warmup": max(args.warmup, 1),
            "ms_per_step": 1e3 * dt / iters, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference", "config": workload_config(n, m, k),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "final_target_loss": loss, "input_generated_on": device,
            "wall_s": time.perf_counter() - t_all}
    emit(line)


def run_ours(args, cfg):
    import torch
    import torch.distributed as td
    from swne_b200 import NMFHandle, synth
    from swne_b200.dist import init_comm, shard_bounds
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, m, k = cfg["n"], args.cells or cfg["m"], cfg["k"]
    b = shard_bounds(m, world)
    m_loc = b[rank + 1] - b[rank]
    # synthetic shard, generated on the device (same generator family on every rank, rank-dependent stream); the
    # reference arm builds the same shards (make_shards) and factorises their concatenation
    dA = synth.make_matrix_torch(n, m_loc, cfg["kt"], cfg["scale"], cfg["seed"] + 1000 * rank)
    torch.cuda.synchronize()
    h = NMFHandle(local)
    if world > 1:
        init_comm(h)
    h.set_matrix(dA)
    W0, H0s = make_inits(n, b, k, world)                 # W0 identical on all ranks, H0 per shard
    H0 = H0s[rank]
    del H0s

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up: W outer iterations from the random start; the timed run continues from that state
    # full_traces=-1: the mkl diagnostic of the mse fit (a sparse pass over the nonzeros, once per fit) stays out of
    # the per-iteration timing; the time-to-tolerance run below includes it, as RunNMF does.
    warm = h.fit(k, W0, H0, max_iter=max(args.warmup, 3), rel_tol=-1.0, engine=args.engine, full_traces=-1)
    barrier()
    with ClockSampler(local) as clk:
        time.sleep(0.2)
        clk.mark()
        t0 = time.perf_counter()
        r = h.fit(k, warm.W, warm.H, max_iter=args.steps, rel_tol=-1.0, engine=args.engine, full_traces=-1)
        barrier()
        wall = time.perf_counter() - t0
        # the timed region is a few tens of ms: keep the GPU under the same load (same step, untimed) until
        # nvidia-smi has caught enough samples of the clocks under this load
        # (every fit is a collective over the ranks: the decision to run one more is all-reduced, never local)
        t_load = time.perf_counter()
        while True:
            more = clk.n_loaded() < 8 and time.perf_counter() - t_load < 3.0
            if world > 1:
                flag = torch.tensor([1 if more else 0], dtype=torch.int32, device="cuda")
                td.all_reduce(flag, op=td.ReduceOp.MAX)
                more = bool(flag.item())
            if not more:
                break
            h.fit(k, r.W, r.H, max_iter=max(args.steps, 20), rel_tol=-1.0, engine=args.engine, full_traces=-1)
    loop_ms = maxr(r.stats["loop_ms"])
    ms_per_step = loop_ms / args.steps
    value = 1e3 / ms_per_step

    # ---- roofline of the dominant kernel, measured live with CUDA events on the library's stream
    pr = h.fit(k, r.W, r.H, max_iter=min(args.steps, 10), rel_tol=-1.0, engine=args.engine, profile=True, full_traces=-1)
    prof_ms, prof_n = pr.stats["prof_ms"], pr.stats["prof_launches"]
    engine = pr.options["engine"]
    peak, peak_src = peaks()
    a_bytes = 4.0 * n * m_loc
    if engine == "tc":
        dom = 0
        alg_bytes = a_bytes + 4.0 * k * (2 * m_loc + 2 * n)
        dom_name = "fused pass (W'A -> H solve -> A H', H H')"
    else:
        dom = 0 if prof_ms[0] >= prof_ms[2] else 2
        alg_bytes = a_bytes + 4.0 * k * (m_loc + n)
        dom_name = "contract_cells (W'A)" if dom == 0 else "contract_genes (A H')"
    # slot 0 of the tcgen05 engine holds the pass kernel plus three tiny operand-split kernels: per iteration
    dom_ms = prof_ms[dom] / max(pr.n_iteration if engine == "tc" else prof_n[dom], 1)
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    iters_prof = pr.n_iteration
    shares = {nm: round(ms / max(sum(prof_ms), 1e-9), 4) for nm, ms in zip(
        ("contract_H_or_pass", "solve_H", "contract_W", "solve_W", "gram_loss", "mkl_update", "allreduce", "other"), prof_ms)}
    # dram__bytes_read.sum + dram__bytes_write.sum of one tc_pass_kernel launch on this workload, from the round-2 ncu --set full
    # capture summarised in profiles/r2_iteration_kernels_ncu_full_key_metrics.txt (2.4887 GB + 0.0536 GB = 5.41 TB/s over the
    # 469.6 us launch): the pass reads A twice by design, so traffic is ~2.1x the algorithmic bytes
    traffic = 2.5423e9 if (engine == "tc" and world == 1 and m == 100000) else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": dom_name, "kernel_ms": dom_ms, "peak_source": peak_src,
                "iteration_frac_of_hbm_roofline": ((a_bytes + 4.0 * k * (2 * m_loc + 2 * n)) / (ms_per_step * 1e-3) / 1e9) / peak,
                "profiled_ms_per_iteration": sum(prof_ms) / max(iters_prof, 1), "family_share": shares}

    # ---- end to end through the C-ABI with HOST buffers: upload A (pinned fp32), W0/H0, K iterations, download
    e2e = None
    if not args.no_e2e:
        hostA = torch.empty((m_loc, n), dtype=torch.float32, pin_memory=True)
        hostA.copy_(dA)
        An = hostA.numpy().T          # (n, m) Fortran-ordered view of the pinned buffer: R's column-major layout
        h.set_matrix(An)              # untimed warm-up of the host path (first-touch of staging buffers, allocator)
        h.fit(k, warm.W, warm.H, max_iter=3, rel_tol=-1.0, engine=args.engine, full_traces=-1)
        barrier()
        t0 = time.perf_counter()
        h.set_matrix(An)
        re = h.fit(k, warm.W, warm.H, max_iter=args.steps, rel_tol=-1.0, engine=args.engine, full_traces=-1)
        barrier()
        e2e_s = maxr(time.perf_counter() - t0)
        h2d = An.nbytes + 8.0 * (n * k + k * m_loc)
        d2h = 8.0 * (n * k + k * m_loc)
        e2e = {"value": args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps,
               "d2h_bytes_per_step": d2h / args.steps, "wall_s": e2e_s,
               "note": "one RunNMF-style fit of K iterations: A (fp32, pinned host), W0, H0 uploaded once, W and H read back"}
        h.set_matrix(dA)

    # ---- time to tolerance with NNLM's stopping rule from the same random start (reported, not the headline)
    ttt = None
    if not args.no_tol:
        barrier()
        t0 = time.perf_counter()
        rt = h.fit(k, W0, H0, max_iter=500, rel_tol=1e-4, engine=args.engine)
        barrier()
        ttt = {"iterations": rt.n_iteration, "loop_s": maxr(rt.stats["loop_ms"]) / 1e3, "wall_s": maxr(time.perf_counter() - t0),
               "final_target_loss": float(rt.target_loss[-1]), "converged": not rt.not_converged}

    # ---- the same through RunNMF's own seed, init = "nnsvd" (R/nmf.R:109-112): randomized SVD, NNDSVD split and zero
    # replacement on the device, handed to the fit without a host round trip (swne_run_nmf); all ranks take part
    ttt_nnsvd = None
    if not args.no_tol:
        h.fit(k, init="nnsvd", seed=1, max_iter=2, rel_tol=-1.0, engine=args.engine, full_traces=-1)   # untimed: first-use allocations
        barrier()
        t0 = time.perf_counter()
        rn = h.fit(k, init="nnsvd", seed=1, max_iter=500, rel_tol=1e-4, engine=args.engine)
        barrier()
        wall_n = maxr(time.perf_counter() - t0)
        loop_n = maxr(rn.stats["loop_ms"]) / 1e3
        ttt_nnsvd = {"wall_s_seed_plus_fit": wall_n, "iterations": rn.n_iteration, "loop_s": loop_n, "seed_and_transfers_s": wall_n - loop_n,
                     "final_target_loss": float(rn.target_loss[-1]), "converged": not rn.not_converged}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # the same matrix (device copy -> host fp64) and inits, all host cores, a bounded number of iterations
        A64 = np.empty((n, m), dtype=np.float64, order="F")
        A64[:, :] = dA.cpu().numpy().T
        cores = host_cores()
        val, dt, _ = cpu_oracle_run(A64, k, W0, H0, 2, args.cpu_iters, cores)
        del A64
        cpu = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "full workload (%d x %d, the same matrix), 2 warm-up + %d timed iterations, %.1f s timed" % (n, m, args.cpu_iters, dt)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": workload_config(n, m, k),
                "run": {"cells_per_gpu": m_loc, "engine": engine, "parallelism": "cells sharded over %d GPU(s)" % world,
                        "exchange": (None if world == 1 else ("NVLink peer memory (one-shot all-reduce kernels)" if h.comm_info()["peer_exchange"]
                                                               else "NCCL all-reduce"))},
                "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(r.stats["gpu_launches"]),
                "roofline": roofline, "cpu_baseline": cpu, "time_to_tolerance": ttt, "time_to_tolerance_nnsvd_seed": ttt_nnsvd,
                "timed_region": {"wall_s_incl_upload_download": wall, "loop_ms_cuda_events_max_over_ranks": loop_ms,
                                 "final_target_loss": float(r.target_loss[-1])}}
        emit(line)
    h.close()
    if world > 1:
        td.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# --config cfg5: BASELINE.json configuration 5 as specified -- 3000 genes x 1.3 M cells, k = 40, mse, the matrix handed over
# as CSC (dgCMatrix slots: int32 p / i, float32 x), cells sharded over the ranks by EQUAL NNZ (dist.shard_bounds_nnz).
# The synthetic matrix is a concatenation of independently generated 20 000-cell chunks (chunk c: generator seed + c), so
# any rank can produce any cell range without holding the whole matrix.
# ------------------------------------------------------------------------------------------------
CFG5_CHUNK = 20000


def cfg5_chunk(cfg, c, m, device):
    from swne_b200 import synth
    mc = min(CFG5_CHUNK, m - c * CFG5_CHUNK)
    return synth.make_matrix_torch(cfg["n"], mc, cfg["kt"], cfg["scale"], cfg["seed"] + c, device=device)


def cfg5_csc_range(cfg, m, lo, hi, device):
    """CSC slots (p int32, i int32, x float32) of cells [lo, hi) of the synthetic matrix"""
    import torch
    ps, iis, xs = [np.zeros(1, dtype=np.int64)], [], []
    for c in range(lo // CFG5_CHUNK, (hi - 1) // CFG5_CHUNK + 1):
        D = cfg5_chunk(cfg, c, m, device)                           # (cells, genes)
        a, b = max(lo, c * CFG5_CHUNK) - c * CFG5_CHUNK, min(hi, (c + 1) * CFG5_CHUNK) - c * CFG5_CHUNK
        D = D[a:b]
        nz = D != 0
        ps.append(nz.sum(dim=1).cpu().numpy().astype(np.int64))
        iis.append(nz.nonzero()[:, 1].to(torch.int32).cpu().numpy())
        xs.append(D[nz].cpu().numpy())
        del D, nz
    p = np.cumsum(np.concatenate(ps))
    assert p[-1] < 2**31
    return p.astype(np.int32), np.concatenate(iis), np.concatenate(xs)


def run_cfg5(args):
    import scipy.sparse
    import torch
    import torch.distributed as td
    from swne_b200 import NMFHandle, synth
    from swne_b200.dist import init_comm, shard_bounds, shard_bounds_nnz
    cfg = dict(synth.CONFIGS["cfg5"])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n, m, k = cfg["n"], args.cells or cfg["m"], cfg["k"]
    metric = "RunNMF iterations/s (%d genes x %d cells, k=%d, mse, CSC input)" % (n, m, k)
    config = {"workload": "cfg5: %d genes x %d cells, k=%d, mse, sparse CSC input (int32 p/i, float32 x), cells sharded by equal nnz, "
                          "random init U(0,1)*0.01, timed from the state after the warm-up iterations" % (n, m, k),
              "genes": n, "cells": m, "k": k, "loss": "mse",
              "l2": "input (%.1f GB dense fp32 on the device) larger than L2" % (4.0 * n * m / 1e9)}
    if args.impl == "reference":
        if rank != 0:
            return
        # restated NNLM on a 100 000-cell slice (dense fp64: the full matrix would be 31 GB twice), scaled by cells: flagged
        cells = min(m, 100000)
        _, iis, xs = None, None, None
        p_, i_, x_ = cfg5_csc_range(cfg, m, 0, cells, "cuda" if torch.cuda.is_available() else "cpu")
        A64 = np.asfortranarray(scipy.sparse.csc_matrix((x_.astype(np.float64), i_, p_), shape=(n, cells)).toarray())
        W0 = synth.random_init(n, 1, k, 0)[0]
        H0 = synth.random_init(n, cells, k, 100)[1]
        cores = host_cores()
        val, dt, loss = cpu_oracle_run(A64, k, W0, H0, min(args.warmup, 1), min(max(args.steps, 1), 4), cores)
        val *= cells / m
        line = {"metric": metric, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": min(max(args.steps, 1), 4), "warmup": min(args.warmup, 1),
                "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": "EXTRAPOLATED: first %d of %d cells (dense fp64), value scaled by %d/%d; %.1f s timed" % (cells, m, cells, m, dt)},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    # ---- pass 1: nnz per cell of an equal-cell range, all-gathered -> the global column pointer -> equal-nnz bounds
    b0 = shard_bounds(m, world)
    cnt = torch.zeros(m, dtype=torch.int64, device="cuda")
    for c in range(b0[rank] // CFG5_CHUNK, (b0[rank + 1] - 1) // CFG5_CHUNK + 1):
        lo, hi = max(b0[rank], c * CFG5_CHUNK), min(b0[rank + 1], (c + 1) * CFG5_CHUNK)
        D = cfg5_chunk(cfg, c, m, "cuda")
        cnt[lo:hi] = (D[lo - c * CFG5_CHUNK:hi - c * CFG5_CHUNK] != 0).sum(dim=1)
        del D
    if world > 1:
        td.all_reduce(cnt)
    indptr = np.concatenate([[0], np.cumsum(cnt.cpu().numpy())])
    b = shard_bounds_nnz(indptr, world)
    lo, hi = b[rank], b[rank + 1]
    # ---- pass 2: this rank's cells as CSC on the host (pinned), handed to the library as dgCMatrix slots
    p_, i_, x_ = cfg5_csc_range(cfg, m, lo, hi, "cuda")
    m_loc = hi - lo
    S = scipy.sparse.csc_matrix((x_, i_, p_), shape=(n, m_loc))
    torch.cuda.empty_cache()
    h = NMFHandle(local)
    if world > 1:
        init_comm(h)
    barrier()
    t0 = time.perf_counter()
    h.set_matrix(S)
    barrier()
    ingest_s = maxr(time.perf_counter() - t0)
    info = h.matrix_info()
    W0 = synth.random_init(n, 1, k, 0)[0]
    H0 = synth.random_init(n, m_loc, k, 100 + rank)[1]
    warm = h.fit(k, W0, H0, max_iter=max(args.warmup, 3), rel_tol=-1.0, engine=args.engine, full_traces=-1)
    barrier()
    with ClockSampler(local) as clk:
        time.sleep(0.2)
        clk.mark()
        r = h.fit(k, warm.W, warm.H, max_iter=args.steps, rel_tol=-1.0, engine=args.engine, full_traces=-1)
        barrier()
        t_load = time.perf_counter()
        while True:
            more = clk.n_loaded() < 8 and time.perf_counter() - t_load < 3.0
            if world > 1:
                flag = torch.tensor([1 if more else 0], dtype=torch.int32, device="cuda")
                td.all_reduce(flag, op=td.ReduceOp.MAX)
                more = bool(flag.item())
            if not more:
                break
            h.fit(k, r.W, r.H, max_iter=max(args.steps, 10), rel_tol=-1.0, engine=args.engine, full_traces=-1)
    loop_ms = maxr(r.stats["loop_ms"])
    ms_per_step = loop_ms / args.steps
    pr = h.fit(k, r.W, r.H, max_iter=min(args.steps, 6), rel_tol=-1.0, engine=args.engine, profile=True, full_traces=-1)
    prof_ms = [x / max(pr.n_iteration, 1) for x in pr.stats["prof_ms"]]
    peak, peak_src = peaks()
    # dense algorithmic bytes per iteration on this rank: A once + W, H read and written
    alg = 4.0 * n * m_loc + 4.0 * k * (2 * m_loc + 2 * n)
    dom = int(np.argmax(prof_ms[:4]))
    names = ("contract W'A (tcx_cells)", "solve H", "contract A H' (tcx_genes)", "solve W")
    a_bytes = 4.0 * n * m_loc
    roofline = {"bound": "hbm", "achieved": a_bytes / (prof_ms[dom] * 1e-3) / 1e9 if dom in (0, 2) else None, "peak": peak, "unit": "GB/s",
                "frac": (a_bytes / (prof_ms[dom] * 1e-3) / 1e9) / peak if dom in (0, 2) else None, "traffic": None,
                "kernel": names[dom], "kernel_ms": prof_ms[dom], "peak_source": peak_src,
                "iteration_frac_of_hbm_roofline": (alg / (ms_per_step * 1e-3) / 1e9) / peak,
                "profiled_ms_per_iteration": {nm: round(v, 4) for nm, v in zip(("contract_H", "solve_H", "contract_W", "solve_W", "gram_loss", "mkl", "allreduce", "other"), prof_ms)}}
    # ---- end to end: CSC slots from host memory -> device (dense + split copy) -> K iterations -> W, H back
    barrier()
    t0 = time.perf_counter()
    h.set_matrix(S)
    re = h.fit(k, warm.W, warm.H, max_iter=args.steps, rel_tol=-1.0, engine=args.engine, full_traces=-1)
    barrier()
    e2e_s = maxr(time.perf_counter() - t0)
    h2d = p_.nbytes + i_.nbytes + x_.nbytes + 8.0 * (n * k + k * m_loc)
    e2e = {"value": args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": 8.0 * (n * k + k * m_loc) / args.steps,
           "wall_s": e2e_s, "note": "dgCMatrix slots (host) -> device ingest + K iterations + W, H read back"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cells = min(m_loc, 40000)
        A64 = np.asfortranarray(S[:, :cells].toarray().astype(np.float64))
        cores = host_cores()
        val, dt, _ = cpu_oracle_run(A64, k, W0, H0[:, :cells], 1, 3, cores)
        cpu = {"value": val * cells / m, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "EXTRAPOLATED: first %d of %d cells, 1 warm-up + 3 timed iterations (%.1f s), value scaled by %d/%d" % (cells, m, dt, cells, m)}
    if rank == 0:
        line = {"metric": metric, "value": 1e3 / ms_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "run": {"cells_this_rank": m_loc, "nnz_this_rank": int(info["nnz"]), "nnz_bounds": [int(x) for x in b], "engine": pr.options["engine"],
                        "csc_ingest_s": ingest_s, "parallelism": "cells sharded by equal nnz over %d GPU(s)" % world},
                "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(r.stats["gpu_launches"]), "roofline": roofline, "cpu_baseline": cpu}
        emit(line)
    h.close()
    if world > 1:
        td.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line, on the process's real stdout (everything else this process or its libraries print goes to stderr)"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # NCCL prints its version banner on stdout when the first communicator is created: keep stdout for the JSON line
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--cells", type=int, default=0, help="override the number of cells (debug / bounded CPU sample)")
    ap.add_argument("--cpu-iters", type=int, default=12)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-tol", action="store_true")
    ap.add_argument("--config", default="cfg3", choices=["cfg3", "cfg5"], help="cfg3: the headline (default); cfg5: 1.3 M cells, k=40, CSC input")
    args = ap.parse_args()
    from swne_b200 import synth
    if args.config == "cfg5":
        run_cfg5(args)
        return
    cfg = dict(synth.CONFIGS["cfg3"])
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
