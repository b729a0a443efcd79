"""Throughput of independent cost+grad evaluations versus the number of evaluation lanes (mambo_csa_create_lane).
  device leg : one host thread issues eval_device round-robin over the lanes' streams (CUDA events, device resident)
  host leg   : one host thread per lane calls the synchronous host closure cost_grad (H2D + D2H inside)
usage: python tools/lane_sweep.py N [max_lanes] [points]"""
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import mambo_b200 as mb  # noqa: E402
import importlib  # noqa: E402

syn = importlib.import_module("quantummambo_jl_b200.synthetic")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
max_lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
points = int(sys.argv[3]) if len(sys.argv) > 3 else 24
dev = torch.device("cuda", 0)
T = syn.dense_sym_tbt(n, 0)
xs = [syn.random_x(n, i) for i in range(points)]
parent = mb.CSAEngine(T)
L = parent.L
d_x = torch.from_numpy(np.stack(xs)).to(dev)
d_out = torch.empty(points, L + 1, dtype=torch.float64, device=dev)
group = [parent]
main = torch.cuda.Stream(dev)
for nl in range(1, max_lanes + 1):
    while len(group) < nl:
        group.append(parent.lane())
    streams = [torch.cuda.Stream(dev) for _ in range(nl)]
    for e, s in zip(group, streams):
        e.set_stream(s.cuda_stream)

    def sweep():
        for i in range(points):
            group[i % nl].eval_device(d_x[i].data_ptr(), d_out[i].data_ptr())

    def timed(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(main)
        for s in streams:
            s.wait_stream(main)
        for _ in range(reps):
            sweep()
        for s in streams:
            main.wait_stream(s)
        e1.record(main)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    timed(2)
    ms = timed(5)
    dev_rate = 5 * points / (ms * 1e-3)
    for e in group[:nl]:
        e.set_stream(None)

    def worker(k):
        for rep in range(3):
            for i in range(k, points, nl):
                xi = xs[i].copy()
                xi[0] += 1e-9 * (rep + 1)
                group[k].cost_grad(xi)

    for k in range(nl):
        group[k].cost_grad(xs[k])
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(k,)) for k in range(nl)]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    print(f"N={n} lanes={nl}: device {dev_rate:9.1f} evals/s ({1e3 / dev_rate:.4f} ms/eval)   host-threads {3 * points / dt:9.1f} evals/s", flush=True)
c, g = group[-1].cost_grad(xs[0])
c0, g0 = parent.cost_grad(xs[0])
print("lane vs parent bits equal:", c == c0 and np.array_equal(g, g0))
