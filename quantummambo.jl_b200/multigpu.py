"""One-process-per-GPU host logic (torch.distributed plumbing; NCCL on GPUs, gloo in the CPU tests).

Two partitions of the CSA path (SURVEY.md 8e / BASELINE.json north_star):
  1. multistart replicas  -- K independent optimisations sharded over ranks, no data-path collective, a final
     gather + argmin of K scalars (precedent: orbitals.jl:39-67);
  2. row-panel sharding   -- each rank owns a contiguous block of 64-row panels of the residual matrix; per
     evaluation ONE allreduce(sum) of [cost_2body, S (N x N), G (N x N)] = 1 + 2 N^2 doubles, then every rank
     finishes (one-body terms, adjoint Frechet derivative, gradient assembly) redundantly.

The evaluator is duck-typed (`eval_partial(x) -> partial`, `eval_finish(partial) -> (cost, grad)`): on GPUs it is
`ShardedEngine` below (CUDA library, device buffers, NCCL); the CPU tests drive the same functions with an
oracle-backed stand-in over gloo.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np


def multistart_partition(K: int, world: int, rank: int) -> List[int]:
    """Start indices owned by `rank`: round-robin, so unequal run times average out."""
    return list(range(rank, K, world))


def panel_partition(n_panels: int, world: int, rank: int) -> Tuple[int, int]:
    """[begin, end) panels of `rank` -- the same split the library uses (mambo_csa_create_sharded)."""
    return (n_panels * rank) // world, (n_panels * (rank + 1)) // world


def allreduce_sum(t, dist_mod=None):
    """Sum `t` (torch tensor, in place) over the default process group; no-op in a single process."""
    import torch.distributed as dist
    d = dist if dist_mod is None else dist_mod
    if d.is_available() and d.is_initialized() and d.get_world_size() > 1:
        d.all_reduce(t, op=d.ReduceOp.SUM)
    return t


def sharded_cost_grad(evaluator, x):
    """Row-panel-sharded cost+gradient: partial on this rank's rows -> allreduce -> replicated finish."""
    part = evaluator.eval_partial(x)
    allreduce_sum(part)
    return evaluator.eval_finish(part)


def gather_best(local_results: Sequence[Tuple[int, float]], world: int):
    """All ranks learn (start index, minimum) of every start and the arg-min (multistart epilogue)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        allr = list(local_results)
    else:
        gathered = [None] * world
        dist.all_gather_object(gathered, list(local_results))
        allr = [r for part in gathered for r in part]
    allr.sort()
    best = min(allr, key=lambda kv: kv[1])
    return allr, best


def run_multistart(make_optimizer: Callable[[], Callable[[np.ndarray], Tuple[np.ndarray, float]]], x0s: np.ndarray,
                   rank: int, world: int):
    """Each rank optimises its share of the K starts with `optimizer(x0) -> (x_min, f_min)`; returns
    (sorted list of (k, f_k) for all K, (k_best, f_best), {k: x_min} of the local starts)."""
    opt = make_optimizer()
    mine = multistart_partition(len(x0s), world, rank)
    local, xs = [], {}
    for k in mine:
        xm, fm = opt(np.asarray(x0s[k], dtype=np.float64))
        local.append((k, float(fm)))
        xs[k] = xm
    allr, best = gather_best(local, world)
    return allr, best, xs


class ShardedEngine:
    """Row-panel shard of one residual tensor on this rank's GPU (CUDA library + torch device buffers)."""

    def __init__(self, tbt, obt=None, flavour="CSA", device: int = 0, rank: int = 0, world: int = 1, sym: int = 0):
        import torch
        from .engine import CSAEngine
        self.torch = torch
        self.dev = torch.device("cuda", device)
        self.eng = CSAEngine(tbt, obt, flavour, device=device, sym=sym, shard=rank, nshards=world)
        self.L = self.eng.L
        self.plen = self.eng.info()["partial_len"]
        self.stream = torch.cuda.Stream(self.dev)
        self.eng.set_stream(self.stream.cuda_stream)
        self.d_x = torch.empty(self.L, dtype=torch.float64, device=self.dev)
        self.d_part = torch.empty(self.plen, dtype=torch.float64, device=self.dev)
        self.d_out = torch.empty(self.L + 1, dtype=torch.float64, device=self.dev)

    def setup_peers(self, rank: int, world: int):
        """Exchange CUDA-IPC handles of the receive buffers over the default process group and map every peer's
        buffer (NVLink peer memory): afterwards `cost_grad_peer` needs no host-launched collective."""
        import torch.distributed as dist
        mine = self.eng.peer_export()
        handles = [None] * world
        dist.all_gather_object(handles, mine)
        for r, hd in enumerate(handles):
            if r != rank:
                self.eng.peer_attach(r, hd)
        dist.barrier()
        self.peers_ready = True

    def cost_grad_peer(self, x):
        """Sharded cost+gradient with the in-library peer exchange (push to every peer, wait, rank-ordered sum)."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            self.d_x.copy_(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)), non_blocking=False)
            self.eng.eval_sharded_device(self.d_x.data_ptr(), self.d_out.data_ptr())
        self.stream.synchronize()
        out = self.d_out.cpu().numpy()
        return float(out[0]), out[1:].copy()

    # -- sharded greedy step: the library's BFGS replicated on every rank, evaluations exchanged over peer memory --------
    def optimize(self, x0, **opt):
        """mambo_csa_optimize on this shard (call on every rank with the same x0 after setup_peers): returns the same
        (x_min, result) on every rank."""
        return self.eng.optimize(x0, **opt)

    def subtract_fragment(self, x) -> float:
        """F_rem -= to_OP(frag) on this rank's rows; returns the new residual cost summed over the shards."""
        c = self.torch.tensor([self.eng.subtract_fragment(x)], dtype=self.torch.float64, device=self.dev)
        allreduce_sum(c)
        return float(c.item())

    def residual_cost(self) -> float:
        c = self.torch.tensor([self.eng.residual_cost()], dtype=self.torch.float64, device=self.dev)
        allreduce_sum(c)
        return float(c.item())

    def eval_partial(self, x):
        torch = self.torch
        with torch.cuda.stream(self.stream):
            self.d_x.copy_(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)), non_blocking=False)
            self.eng.eval_partial_device(self.d_x.data_ptr(), self.d_part.data_ptr())
        self.stream.synchronize()
        return self.d_part

    def eval_finish(self, part):
        torch = self.torch
        with torch.cuda.stream(self.stream):
            self.eng.eval_finish_device(part.data_ptr(), self.d_out.data_ptr())
        self.stream.synchronize()
        out = self.d_out.cpu().numpy()
        return float(out[0]), out[1:].copy()

    def close(self):
        self.eng.close()
