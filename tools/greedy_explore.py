"""Exploration: greedy CSA on a planted tensor with the native optimiser; how far does the cost fall per fragment for the
reference's random-theta start versus the SVD start (tbt_svd_1st), dense BFGS versus L-BFGS?
usage: greedy_explore.py N R alpha_max maxit"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mambo_b200 as mb
from greedy_bench import planted

n, R, amax, maxit = (int(a) for a in sys.argv[1:5])
T = planted(n, R, 0)
rngn = np.random.default_rng(5)
g = rngn.standard_normal((n, n, n, n)); g = g + g.transpose(1, 0, 2, 3); g = g + g.transpose(0, 1, 3, 2); g = g + g.transpose(2, 3, 0, 1)
T = T + 1e-9 * g / 8.0
cL, uL = n * (n + 1) // 2, n * (n - 1) // 2
for start in ("df", "svd", "random"):
    for lb in (0, 10):
        rng = np.random.default_rng(1)
        with mb.CSAEngine(T) as e:
            c = c0 = e.residual_cost()
            t0 = time.perf_counter(); ev = 0; trail = []
            for a in range(amax):
                if c <= 1e-6 * c0:
                    break
                if start == "df":
                    lam, th = mb.DF_based_greedy(e)
                    x0 = np.concatenate([lam, th])
                elif start == "svd":
                    lam, th = mb.tbt_svd_1st(e)
                    x0 = np.concatenate([lam, th])
                else:
                    x0 = np.concatenate([np.zeros(cL), 2 * np.pi * rng.random(uL)])
                xm, res = e.optimize(x0, g_tol=1e-8, max_iter=maxit, lbfgs_memory=lb)
                c = e.subtract_fragment(xm)
                ev += res["evaluations"]
                trail.append((c / c0, res["iterations"], res["converged"]))
            dt = time.perf_counter() - t0
        print(f"N={n} R={R} start={start} lbfgs={lb}: {dt:.2f}s {ev} evals; trail " + " ".join(f"{t[0]:.2e}/{t[1]}{'c' if t[2] else ''}" for t in trail), flush=True)
