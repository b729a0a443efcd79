"""Repeat one evaluation many times (fresh engines and repeated calls) and report any deviation from the first result."""
import os, sys, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mambo_b200 as mb
syn = importlib.import_module("quantummambo_jl_b200.synthetic")
n = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
T = syn.dense_sym_tbt(n, 5)
x = syn.random_x(n, 6)
ref = {}
bad = 0
for r in range(reps):
    with mb.CSAEngine(T) as e:
        for k in range(int(os.environ.get("CALLS", "6"))):
            xx = x.copy(); xx[0] += 1e-9 * k          # defeat the memo
            c, g = e.cost_grad(xx)
            if k not in ref:
                ref[k] = (c, g.copy())
            dc = abs(c - ref[k][0]) / abs(ref[k][0]); dg = np.max(np.abs(g - ref[k][1])) / np.max(np.abs(ref[k][1]))
            if dc > 0 or dg > 0:
                bad += 1
                print(f"engine {r} call {k}: cost dev {dc:.3e} grad dev {dg:.3e}", flush=True)
print(f"N={n}: {reps} engines x 6 calls, {bad} deviations")
