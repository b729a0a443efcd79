"""BASELINE.json configs[2]: synthetic dense-sym N=50, greedy CSA 1-norm workflow with a K-way multistart of every greedy step
spread over the GPUs of one box (one process per GPU, no data-path collective: every rank holds the residual, optimises its
share of the K random-theta starts over its evaluation lanes, the ranks exchange K scalars + the best minimiser, and every rank
applies the same residual update).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
           tools/multistart_greedy.py [N=50] [K=8] [fragments=4] [max_iter=300]
"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import mambo_b200 as mb  # noqa: E402

mg = importlib.import_module("quantummambo_jl_b200.multigpu")
syn = importlib.import_module("quantummambo_jl_b200.synthetic")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
frags = int(sys.argv[3]) if len(sys.argv) > 3 else 4
max_iter = int(sys.argv[4]) if len(sys.argv) > 4 else 300
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

T = syn.dense_sym_tbt(n, 0)
cL, uL = n * (n + 1) // 2, n * (n - 1) // 2
eng = mb.CSAEngine(T, device=local)
mine = mg.multistart_partition(K, world, rank)
lanes = [eng] + [eng.lane() for _ in range(max(len(mine), 1) - 1)]
for ln in lanes:
    ln.optimize_reserve(0)
c0 = c = eng.residual_cost()
rng = np.random.default_rng(3)                       # same starts on every rank
if rank == 0:
    print(f"N={n} dense-sym, {K}-way multistart over {world} GPU(s) ({len(mine)} start(s) per rank on {len(lanes)} lane(s)), "
          f"<= {max_iter} BFGS iterations per start, initial cost {c0:.6e}", flush=True)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
l1, evals = 0.0, 0
for a in range(frags):
    x0s = np.stack([np.concatenate([np.zeros(cL), 2 * np.pi * rng.random(uL)]) for _ in range(K)])
    t1 = time.perf_counter()
    local_res, xs = [], {}
    if mine:
        xm, res, _ = mb.multistart(lanes[:len(mine)], x0s[mine], max_iter=max_iter)
        for j, k in enumerate(mine):
            local_res.append((k, float(res[j]["minimum"])))
            xs[k] = xm[j]
            evals += res[j]["evaluations"]
    allr, best = mg.gather_best(local_res, world)
    kb = best[0]
    xb = torch.zeros(cL + uL, dtype=torch.float64, device="cuda")
    if kb in xs:
        xb.copy_(torch.from_numpy(xs[kb]))
    if world > 1:
        dist.all_reduce(xb)                          # only the owner contributed: a broadcast of the best minimiser
    xbest = xb.cpu().numpy()
    c = eng.subtract_fragment(xbest)
    l1 += float(mb.CSA_L1(mb.CSA_x_to_F_FRAG(xbest, n)))
    if rank == 0:
        print(f"  fragment {a + 1}: best start {kb} of {K} (minima {min(v for _, v in allr):.6e} .. {max(v for _, v in allr):.6e}), "
              f"residual cost {c:.6e}, {time.perf_counter() - t1:.3f} s", flush=True)
ev = torch.tensor([evals], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ev)
if rank == 0:
    dt = time.perf_counter() - t0
    print(f"total {dt:.3f} s, {int(ev.item())} evaluations over all ranks ({ev.item() / dt:.0f} evaluations/s incl. optimiser), "
          f"residual {c / c0:.4f} of the initial cost, sum of fragment 1-norms {l1:.10e}", flush=True)
for ln in lanes[1:]:
    ln.close()
eng.close()
if world > 1:
    dist.destroy_process_group()
