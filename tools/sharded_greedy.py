"""BASELINE.json configs[4]: greedy CSA on a large synthetic tensor with the residual sharded by row panels over the GPUs of
one box.  Every rank holds its rows of the residual and runs the library's optimiser on its shard handle; the evaluations
exchange their partial vectors over NVLink peer memory inside the library (no host collective on the data path).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
           tools/sharded_greedy.py [N=152] [fragments=1] [max_iter=300] [lbfgs_memory=10]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import mambo_b200 as mb  # noqa: E402
import importlib  # noqa: E402

mg = importlib.import_module("quantummambo_jl_b200.multigpu")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 152
frags = int(sys.argv[2]) if len(sys.argv) > 2 else 1
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 300
mem = int(sys.argv[4]) if len(sys.argv) > 4 else 10
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

# planted tensor: one CSA fragment (lambda ~ N(0,1), theta ~ U[0, 2pi)) + the dense-sym background scaled down
rng = np.random.default_rng(7)
cL, uL = n * (n + 1) // 2, n * (n - 1) // 2
x_true = np.concatenate([rng.standard_normal(cL), 2 * np.pi * rng.random(uL)])
with mb.CSAEngine(np.zeros((n, n, n, n)), device=local) as e:
    T, _ = e.reconstruct(x_true)
T = np.asfortranarray(T)
eng = mg.ShardedEngine(T, device=local, rank=rank, world=world)
if world > 1:
    eng.setup_peers(rank, world)
else:
    eng.eng.peer_export()
eng.eng.optimize_reserve(mem)
c = eng.residual_cost()
if rank == 0:
    print(f"N={n} shards={world} rows/shard={eng.eng.info()['local_row_end'] - eng.eng.info()['local_row_begin']} initial cost {c:.6e}", flush=True)
xrng = np.random.default_rng(11)           # same starts on every rank
t0 = time.perf_counter()
for a in range(frags):
    x0 = np.concatenate([np.zeros(cL), 2 * np.pi * xrng.random(uL)])
    t1 = time.perf_counter()
    xm, res = eng.optimize(x0, g_tol=1e-8, max_iter=max_iter, lbfgs_memory=mem)
    c = eng.subtract_fragment(xm)
    if rank == 0:
        dt = time.perf_counter() - t1
        print(f"  fragment {a + 1}: cost {c:.6e}  iterations {res['iterations']}  evaluations {res['evaluations']}  {dt:.2f} s "
              f"({dt / max(res['evaluations'], 1) * 1e3:.3f} ms per evaluation incl. optimiser)", flush=True)
if rank == 0:
    print(f"total {time.perf_counter() - t0:.2f} s", flush=True)
eng.close()
if world > 1:
    dist.destroy_process_group()
