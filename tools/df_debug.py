import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mambo_b200 as mb
from greedy_bench import planted
n, R = int(sys.argv[1]), int(sys.argv[2])
T = planted(n, R, 0)
rngn = np.random.default_rng(5)
g = rngn.standard_normal((n, n, n, n)); g = g + g.transpose(1, 0, 2, 3); g = g + g.transpose(0, 1, 3, 2); g = g + g.transpose(2, 3, 0, 1)
for sigma in (0.0, 1e-9):
    Tn = T + sigma * g / 8.0
    with mb.CSAEngine(Tn) as e:
        try:
            st = e.df_decompose(verify=True)
            print("sigma", sigma, {k: st[k] for k in ("krylov_dim", "restarts", "num_frags", "jacobi_sweeps_max", "ortho_dev_after", "max_rel_residual")})
        except Exception as exc:
            print("sigma", sigma, "FAILED", exc)
