"""Small target for one `ncu --set full` capture of the round-2 kernels at N = 100: three evaluations (fused::k_ty), one residual
update (fused::k_fused STORE), three dense-BFGS iterations (k_sym_update_gemv), a few Lanczos steps (lanczos::k_symv) and a
double factorisation at N = 50 (df::k_dgemm, k_jacobi_batch, k_project_split)."""
import os, sys, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mambo_b200 as mb
syn = importlib.import_module("quantummambo_jl_b200.synthetic")
n = 100
T = syn.dense_sym_tbt(n, 0)
with mb.CSAEngine(T) as e:
    for k in range(3):
        e.cost_grad(syn.random_x(n, k))
    e.optimize(syn.random_x(n, 5), max_iter=4)
    e.leading_eig(max_iter=6)
    e.subtract_fragment(syn.random_x(n, 7))
with mb.CSAEngine(syn.dense_sym_tbt(50, 0)) as e:
    print(e.df_decompose()["num_frags"])
