"""One double factorisation of the dense-sym tensor of size N on cuda:0 (profiling target).  usage: df_one.py N"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
import mambo_b200 as mb
syn = importlib.import_module("quantummambo_jl_b200.synthetic")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
with mb.CSAEngine(syn.dense_sym_tbt(n, 0)) as e:
    t0 = time.perf_counter()
    st = e.df_decompose()
    print(f"N={n}: {time.perf_counter() - t0:.3f} s", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items()})
