"""world_size-2 gloo tests (CPU) of the one-process-per-GPU host logic: the row-panel allreduce protocol and the
multistart partition/gather.  The evaluator here is an oracle-backed stand-in with the engine's two-phase
interface; on GPUs the same functions drive the CUDA library (tests/test_gpu_parity.py, bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class OracleShard:
    """CPU stand-in for ShardedEngine: rows [begin,end) of the n^2 x n^2 residual via the oracle."""

    def __init__(self, T, rank, world):
        import csa_oracle as o
        self.o, self.T, self.n = o, T, T.shape[0]
        rows = self.n * self.n
        self.begin, self.end = rows * rank // world, rows * (rank + 1) // world

    def eval_partial(self, x):
        self.x = x
        return torch.from_numpy(self.o.partial_factored(x, self.T, self.begin, self.end))

    def eval_finish(self, part):
        return self.o.finish_factored(part.numpy(), self.x, self.n)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    import mambo_b200  # noqa: F401
    mg = importlib.import_module("quantummambo_jl_b200.multigpu")
    import csa_oracle as o
    n = 5
    T = o.pairsym_tbt(n, 1)
    x = o.random_x(n, 2)
    # --- row-panel protocol: partial -> allreduce -> finish equals the unsharded evaluation -----------------
    c, g = mg.sharded_cost_grad(OracleShard(T, rank, world), x)
    c0, g0 = o.cost_grad_factored(x, T)
    ok_rp = abs(c - c0) <= 1e-12 * abs(c0) and np.max(np.abs(g - g0)) <= 1e-11 * np.max(np.abs(g0))
    # --- multistart: 5 starts over 2 ranks, everyone learns the same arg-min -------------------------------------
    x0s = np.stack([o.random_x(n, 10 + k) for k in range(5)])

    def make_opt():
        return lambda x0: (x0, o.cost_grad_factored(x0, T, need_grad=False)[0])   # "optimiser" = evaluate the start

    allr, best, xs = mg.run_multistart(make_opt, x0s, rank, world)
    ref = [(k, o.cost_grad_factored(x0s[k], T, need_grad=False)[0]) for k in range(5)]
    ok_ms = [k for k, _ in allr] == [0, 1, 2, 3, 4] and all(abs(a[1] - b[1]) < 1e-14 for a, b in zip(allr, ref)) \
        and best[0] == min(ref, key=lambda kv: kv[1])[0] and sorted(xs) == mg.multistart_partition(5, world, rank)
    q.put((rank, bool(ok_rp), bool(ok_ms)))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_rowpanel_and_multistart_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), "row-panel allreduce protocol mismatch"
    assert all(r[2] for r in res), "multistart partition/gather mismatch"


def test_partitions():
    import importlib
    sys.path.insert(0, ROOT)
    import mambo_b200  # noqa: F401
    mg = importlib.import_module("quantummambo_jl_b200.multigpu")
    assert mg.multistart_partition(10, 4, 1) == [1, 5, 9]
    cover = []
    for r in range(8):
        b, e = mg.panel_partition(157, 8, r)
        cover += list(range(b, e))
    assert cover == list(range(157))
