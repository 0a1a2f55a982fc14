"""mkl on a cell-sharded matrix (3000 x CELLS, k = 30): per-iteration time at WORLD_SIZE ranks, max over ranks.
    python tools/exp_mkl_sharded.py                                      (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/exp_mkl_sharded.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as td
from swne_b200 import NMFHandle, synth
from swne_b200._capi import PROF_NAMES
from swne_b200.dist import init_comm, shard_bounds
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
c = synth.CONFIGS["cfg3"]
for m in [int(x) for x in os.environ.get("CELLS", "100000").split(",")]:
    b = shard_bounds(m, world)
    m_loc = b[rank + 1] - b[rank]
    dA = synth.make_matrix_torch(c["n"], m_loc, c["kt"], c["scale"], c["seed"] + 1000 * rank)
    h = NMFHandle(local)
    if world > 1:
        init_comm(h)
    h.set_matrix(dA)
    W0 = synth.random_init(c["n"], 1, c["k"], 0)[0]
    H0 = synth.random_init(c["n"], m_loc, c["k"], 100 + rank)[1]
    h.fit(c["k"], W0, H0, loss="mkl", max_iter=1, rel_tol=-1.0, trace=100)
    r = h.fit(c["k"], W0, H0, loss="mkl", max_iter=4, rel_tol=-1.0, trace=100, profile=True)
    t = torch.tensor([r.stats["loop_ms"] / 4] + [v / 4 for v in r.stats["prof_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    if rank == 0:
        print("mkl 3000 x %d k=30 on %d GPU(s): %.2f ms/iteration (max over ranks) %s, final mkl %.6f" % (
            m, world, float(t[0]), {nm: round(float(v), 2) for nm, v in zip(PROF_NAMES, t[1:]) if float(v) > 0}, r.mkl[-1]), flush=True)
    h.close()
if world > 1:
    td.barrier()
    td.destroy_process_group()
