"""A few fused cost+grad evaluations at N=100 (dense-sym, packed mode): the smallest command that exercises
every kernel of the hot path; used as the ncu target."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mambo_b200 as mb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
sym = {"auto": mb.SYM_AUTO, "pair": mb.SYM_PAIR, "general": mb.SYM_GENERAL}[sys.argv[2] if len(sys.argv) > 2 else "auto"]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
T = mb.synthetic.dense_sym_tbt(n, 0)
with mb.CSAEngine(T, sym=sym) as e:
    for i in range(reps):
        c, g = e.cost_grad(mb.synthetic.random_x(n, i))
    print("cost", c, "mode", e.info()["mode_name"])
