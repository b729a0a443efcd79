"""Molecule-sized evaluations (BASELINE configs[0..1]: N = 6, 7, 10): the single-kernel path (tiny::k_eval) against the
11-launch pipeline (MAMBO_NO_TINY=1) -- host closure rate, device-resident rate (CUDA events), and the oracle on the host."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import mambo_b200 as mb
import csa_oracle as o

fx = np.load(os.path.join(ROOT, "tests", "golden", "molecules.npz"))
REPS = 2000
for name in ("lih", "h2o", "n2"):
    T, h = fx[f"{name}_tbt"], fx[f"{name}_obt"]
    n = T.shape[0]
    for flavour, obt in (("CSA", None), ("CSA_SD", h)):
        xs = [o.random_x(n, s, flavour) for s in range(8)]
        row = {}
        for label, env in (("single kernel", None), ("pipeline", "1")):
            if env: os.environ["MAMBO_NO_TINY"] = env
            else: os.environ.pop("MAMBO_NO_TINY", None)
            with mb.CSAEngine(T, obt, flavour) as e:
                for x in xs: e.cost_grad(x)
                t0 = time.perf_counter()
                for r in range(REPS): e.cost_grad(xs[r % 8])
                host = REPS / (time.perf_counter() - t0)
                d_x = [torch.tensor(x, dtype=torch.float64, device="cuda") for x in xs]
                d_o = torch.zeros(1 + len(xs[0]), dtype=torch.float64, device="cuda")
                st = torch.cuda.Stream()
                e.set_stream(st.cuda_stream)
                with torch.cuda.stream(st):
                    for r in range(16): e.eval_device(d_x[r % 8].data_ptr(), d_o.data_ptr())
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(st)
                    for r in range(REPS): e.eval_device(d_x[r % 8].data_ptr(), d_o.data_ptr())
                    b.record(st)
                    b.synchronize()
                dev = REPS / (a.elapsed_time(b) * 1e-3)
                e.set_stream(None)
                t0 = time.perf_counter()
                xm, res = e.optimize(xs[0], g_tol=1e-9, max_iter=500)
                opt = (time.perf_counter() - t0) / max(res["evaluations"], 1)
                row[label] = (host, dev, opt * 1e6, res["minimum"])
        os.environ.pop("MAMBO_NO_TINY", None)
        t0 = time.perf_counter()
        for r in range(50): o.cost_grad_factored(xs[r % 8], T, obt, flavour)
        cpu = 50 / (time.perf_counter() - t0)
        a, b = row["single kernel"], row["pipeline"]
        print(f"{name:4s} N={n:2d} {flavour:6s}: host closure {a[0]:8.0f}/s (pipeline {b[0]:7.0f}/s)  device-resident {a[1]:8.0f}/s (pipeline {b[1]:7.0f}/s)  "
              f"native BFGS {a[2]:6.1f} us/eval (pipeline {b[2]:6.1f})  minima {a[3]:.6e} / {b[3]:.6e}   oracle on the host {cpu:6.0f}/s", flush=True)
