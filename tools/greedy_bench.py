"""Greedy CSA decomposition to tolerance on a planted tensor with the library's device-resident BFGS
(mambo_csa_optimize) -- time, evaluations, cost trail.  usage: greedy_bench.py N R alpha_max [tol] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mambo_b200 as mb

def planted(n, R, seed):
    """T = sum_k W(x_k) built on the GPU with reconstruct (lambda ~ N(0,1), theta ~ U[0,2pi))."""
    rng = np.random.default_rng(2000 + seed)
    cL, uL = n * (n + 1) // 2, n * (n - 1) // 2
    T = np.zeros((n, n, n, n), order="F")
    with mb.CSAEngine(np.zeros((n, n, n, n))) as e:
        for _ in range(R):
            x = np.concatenate([rng.standard_normal(cL), 2 * np.pi * rng.random(uL)])
            W, _ = e.reconstruct(x)
            T += W
    return T

if __name__ == "__main__":
    n, R, amax = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    tol = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-6
    seed = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    maxit = int(sys.argv[6]) if len(sys.argv) > 6 else 1000
    T = planted(n, R, seed)
    rng = np.random.default_rng(seed)
    cL, uL = n * (n + 1) // 2, n * (n - 1) // 2
    with mb.CSAEngine(T) as e:
        c = e.residual_cost()
        print(f"N={n} R={R} mode={e.info()['mode_name']} initial cost {c:.6e}")
        t0 = time.perf_counter()
        tot_evals = 0
        for a in range(amax):
            if c <= tol:
                break
            x0 = np.concatenate([np.zeros(cL), 2 * np.pi * rng.random(uL)])
            t1 = time.perf_counter()
            xm, res = e.optimize(x0, g_tol=1e-8, max_iter=maxit, verbose=int(os.environ.get('VERB', '0')),
                                 lbfgs_memory=int(os.environ.get('LBFGS', '0')))
            c = e.subtract_fragment(xm)
            tot_evals += res["evaluations"]
            print(f"  frag {a + 1}: cost {c:.6e} iters {res['iterations']} evals {res['evaluations']} conv {res['converged']} "
                  f"{time.perf_counter() - t1:.2f}s ({(time.perf_counter() - t1) / max(res['evaluations'], 1) * 1e3:.3f} ms/eval)")
        dt = time.perf_counter() - t0
        print(f"total {dt:.2f}s, {tot_evals} evals, final cost {c:.3e}")
