"""BASELINE.json configs[4]: greedy CSA on a large synthetic tensor with the residual sharded by row panels over the GPUs of
one box.  Every rank holds its rows of the residual and runs the library's optimiser on its shard handle; the evaluations
exchange their partial vectors over NVLink peer memory inside the library (no host collective on the data path).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
           tools/sharded_greedy.py [N=152] [planted R=1] [max_iter=3000] [lbfgs_memory=10] [alpha_max=2] [tol=1e-6]
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
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
mem = int(sys.argv[4]) if len(sys.argv) > 4 else 10
alpha_max = int(sys.argv[5]) if len(sys.argv) > 5 else 2
tol = float(sys.argv[6]) if len(sys.argv) > 6 else 1e-6
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

# planted tensor: sum of R CSA fragments (lambda ~ N(0,1), theta ~ U[0, 2pi)) + 1e-9 symmetric noise; the same on every rank
rng = np.random.default_rng(7)
cL, uL = n * (n + 1) // 2, n * (n - 1) // 2
T = np.zeros((n, n, n, n), order="F")
with mb.CSAEngine(np.zeros((n, n, n, n)), device=local) as e:
    for _ in range(R):
        W, _ = e.reconstruct(np.concatenate([rng.standard_normal(cL), 2 * np.pi * rng.random(uL)]))
        T += W
# the noise is drawn and symmetrised on the GPU (4.3 GB at N = 152: numpy would take longer than the whole decomposition)
gen = torch.Generator(device="cuda")
gen.manual_seed(1234)
g = torch.randn((n, n, n, n), dtype=torch.float64, device="cuda", generator=gen)
g = g + g.permute(1, 0, 2, 3)
g = g + g.permute(0, 1, 3, 2)
g = g + g.permute(2, 3, 0, 1)
T += (1e-9 / 8.0) * g.permute(3, 2, 1, 0).contiguous().cpu().numpy().T      # (any fixed layout: the noise is symmetric noise)
del g
torch.cuda.empty_cache()
T = np.asfortranarray(T)
# initial guess of the first step: tbt_svd_1st (guesses.jl:51-84) needs the whole matrix view -> an unsharded handle on this
# rank's GPU (deterministic, so every rank gets the same x0); later steps start from random theta like CSA_greedy_step
t_guess = time.perf_counter()
with mb.CSAEngine(T, device=local) as e:
    lam0, th0 = mb.tbt_svd_1st(e)
t_guess = time.perf_counter() - t_guess
eng = mg.ShardedEngine(T, device=local, rank=rank, world=world)
if world > 1:
    eng.setup_peers(rank, world)
else:
    eng.eng.peer_export()
eng.eng.optimize_reserve(mem)
c = c0 = eng.residual_cost()
if rank == 0:
    info = eng.eng.info()
    print(f"N={n} planted R={R} + 1e-9 noise, shards={world}, rows/shard={info['local_row_end'] - info['local_row_begin']}, "
          f"initial cost {c:.6e}, tolerance {tol:g} on the squared-norm residual (decompose.jl:45), L-BFGS memory {mem} "
          f"(0 = dense BFGS), <= {max_iter} iterations per fragment; SVD guess {t_guess:.2f} s (unsharded handle, incl. upload)", flush=True)
xrng = np.random.default_rng(11)           # same starts on every rank
t0 = time.perf_counter()
evals, l1 = 0, 0.0
for a in range(alpha_max):
    if c <= tol:
        break
    x0 = np.concatenate([lam0, th0]) if a == 0 else np.concatenate([np.zeros(cL), 2 * np.pi * xrng.random(uL)])
    t1 = time.perf_counter()
    xm, res = eng.optimize(x0, g_tol=1e-8, max_iter=max_iter, lbfgs_memory=mem)
    c = eng.subtract_fragment(xm)
    evals += res["evaluations"]
    l1 += float(mb.CSA_L1(mb.CSA_x_to_F_FRAG(xm, n)))
    if rank == 0:
        dt = time.perf_counter() - t1
        print(f"  fragment {a + 1}: cost {c:.6e}  iterations {res['iterations']}  evaluations {res['evaluations']}  converged {res['converged']}  "
              f"{dt:.2f} s ({dt / max(res['evaluations'], 1) * 1e3:.3f} ms per evaluation incl. optimiser)", flush=True)
if rank == 0:
    print(f"total {time.perf_counter() - t0:.2f} s, {evals} evaluations, final cost {c:.3e} ({'reached' if c <= tol else 'NOT reached'} {tol:g}), "
          f"sum of fragment 1-norms {l1:.10e}", flush=True)
eng.close()
if world > 1:
    dist.destroy_process_group()
