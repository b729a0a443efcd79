"""Time mambo_csa_leading_eig (device Lanczos, guesses.jl:51-63) on planted / dense synthetic tensors.
usage: python tools/eig_bench.py N [planted|dense] [max_iter]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mambo_b200 as mb  # noqa: E402
import importlib  # noqa: E402

syn = importlib.import_module("quantummambo_jl_b200.synthetic")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
kind = sys.argv[2] if len(sys.argv) > 2 else "planted"
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cL, uL = n * (n + 1) // 2, n * (n - 1) // 2
if kind == "dense":
    T = syn.dense_sym_tbt(n, 0)
else:
    rng = np.random.default_rng(5)
    with mb.CSAEngine(np.zeros((n, n, n, n))) as e:
        T = sum(e.reconstruct(np.concatenate([rng.standard_normal(cL), 2 * np.pi * rng.random(uL)]))[0] for _ in range(4))
with mb.CSAEngine(T) as e:
    info = e.info()
    e.leading_eig(max_iter=max_iter)            # warm-up: allocates the Lanczos workspace
    t0 = time.perf_counter()
    lam, vec, st = e.leading_eig(max_iter=max_iter)
    dt = time.perf_counter() - t0
    rows = info["rows"]
    print(f"N={n} {kind} mode={info['mode_name']} rows={rows}: lambda={lam:.12e} iterations={st['iterations']} "
          f"rel.residual={st['residual'] / abs(lam):.2e} wall={dt * 1e3:.2f} ms ({dt / st['iterations'] * 1e6:.1f} us/iteration; "
          f"matvec streams {info['tensor_bytes'] / 1e6:.0f} MB)")
