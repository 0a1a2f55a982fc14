"""BASELINE cfg5: 3000 genes x 1,300,000 cells, k=40, mse, cells sharded over the ranks (run under torchrun):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/run_cfg5.py
Each rank generates its shard on the device (synthetic, rank-dependent stream), uploads nothing; timing = CUDA events
of the iteration loop, max over ranks.  CFG5_CELLS overrides the total cell count (weak/strong comparisons)."""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as td
from swne_b200 import NMFHandle, synth
from swne_b200.dist import init_comm, shard_bounds

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
c = synth.CONFIGS["cfg5"]
m = int(os.environ.get("CFG5_CELLS", c["m"]))
n, k = c["n"], c["k"]
b = shard_bounds(m, world)
m_loc = b[rank + 1] - b[rank]
dA = synth.make_matrix_torch(n, m_loc, c["kt"], c["scale"], c["seed"] + 1000 * rank)
h = NMFHandle(local)
if world > 1:
    init_comm(h)
h.set_matrix(dA)
W0 = synth.random_init(n, m_loc, k, 0)[0]
H0 = synth.random_init(n, m_loc, k, 100 + rank)[1]


def maxr(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


out = {}
for eng in os.environ.get("ENGINES", "auto").split(","):
    warm = h.fit(k, W0, H0, max_iter=5, rel_tol=-1.0, engine=eng, full_traces=-1)
    r = h.fit(k, warm.W, warm.H, max_iter=20, rel_tol=-1.0, engine=eng, full_traces=-1)
    ms = maxr(r.stats["loop_ms"]) / 20
    t0 = time.perf_counter()
    rt = h.fit(k, W0, H0, max_iter=500, rel_tol=1e-4, engine=eng, full_traces=0)
    wall = maxr(time.perf_counter() - t0)
    out[eng] = dict(engine=r.options["engine"], ms_per_iteration=ms, iterations_per_s=1e3 / ms,
                    hbm_GBps_per_gpu_alg=(4.0 * n * m_loc + 4.0 * k * 2 * (m_loc + n)) / (ms * 1e-3) / 1e9,
                    to_tol=dict(iterations=rt.n_iteration, loop_s=maxr(rt.stats["loop_ms"]) / 1e3, wall_s=wall,
                                target_loss=float(rt.target_loss[-1]), converged=not rt.not_converged))
if rank == 0:
    print(json.dumps(dict(config="cfg5: %d genes x %d cells, k=%d, mse, random init" % (n, m, k), n_gpus=world,
                          cells_per_gpu=m_loc, results=out)))
h.close()
if world > 1:
    td.barrier()
    td.destroy_process_group()
