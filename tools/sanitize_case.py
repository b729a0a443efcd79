"""Small end-to-end case touching every kernel family (evaluation in the three symmetry modes, CSA_SD, residual update,
lanes, leading eigenpair, optimiser) -- the target of compute-sanitizer runs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import mambo_b200 as mb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
syn = mb.synthetic
T = syn.dense_sym_tbt(n, 0)
x = syn.random_x(n, 1)
for sym in (mb.SYM_AUTO, mb.SYM_PAIR, mb.SYM_GENERAL):
    with mb.CSAEngine(T, sym=sym) as e:
        c, g = e.cost_grad(x)
        print("mode", e.info()["mode_name"], "cost", c)
rng = np.random.default_rng(0)
obt = rng.standard_normal((n, n))
obt = obt + obt.T
xs = np.concatenate([rng.standard_normal(n), x])
with mb.CSAEngine(T, obt, "CSA_SD") as e:
    print("SD cost", e.cost_grad(xs)[0], "after subtract", e.subtract_fragment(xs))
with mb.CSAEngine(T) as e:
    ln = e.lane()
    print("lane", ln.cost_grad(x)[0])
    lam, vec, st = e.leading_eig()
    print("eig", lam, st)
    xm, res = e.optimize(x, max_iter=5)
    xm, res = e.optimize(x, max_iter=5, lbfgs_memory=4)
    print("opt", res)
    print("sub", e.subtract_fragment(xm), e.residual_cost())
    ln.close()
