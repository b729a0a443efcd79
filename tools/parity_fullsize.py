"""One-off parity check at a BASELINE size too large for the test-suite (default N = 152): CUDA engine vs the factored
oracle at one x, cost and norm-wise gradient error.  usage: python tools/parity_fullsize.py [N]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mambo_b200 as mb  # noqa: E402
import csa_oracle as o  # noqa: E402  (checker only)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 152
T = o.dense_sym_tbt(n, 0)
x = o.random_x(n, 1)
t0 = time.perf_counter()
c0, g0 = o.cost_grad_factored(x, T)
t_cpu = time.perf_counter() - t0
with mb.CSAEngine(T) as e:
    x1 = x.copy()
    x1[0] += 1e-9
    e.cost_grad(x1)                       # warm-up at another point (the closure memoises on x)
    t0 = time.perf_counter()
    c, g = e.cost_grad(x)
    t_gpu = time.perf_counter() - t0
    mode = e.info()["mode_name"]
print(f"N={n} mode={mode}: cost rel.err {abs(c - c0) / abs(c0):.2e}  grad max-norm rel.err {np.max(np.abs(g - g0)) / np.max(np.abs(g0)):.2e}  "
      f"(oracle {t_cpu:.2f} s, engine host call {t_gpu * 1e3:.2f} ms)")
