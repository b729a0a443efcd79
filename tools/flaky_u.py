"""Determinism of the orbital-rotation chain alone: U diag(lambda1) U^T of a CSA_SD fragment (mambo_csa_reconstruct with
tbt_out = NULL: k_chain_u + operand build + k_sd_onebody) repeated many times at the same x, compared bitwise."""
import os, sys, importlib, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mambo_b200 as mb
syn = importlib.import_module("quantummambo_jl_b200.synthetic")
lib = mb.load()
for n in [int(a) for a in sys.argv[2:]]:
    reps = int(sys.argv[1])
    dT = torch.zeros(n ** 4, dtype=torch.float64, device="cuda")
    dO = torch.zeros(n * n, dtype=torch.float64, device="cuda")
    x = syn.random_x(n, 6, "CSA_SD")
    with mb.CSAEngine.from_device(dT.data_ptr(), n, dO.data_ptr(), flavour="CSA_SD") as e:
        ref, bad, worst = None, 0, 0.0
        out = np.empty((n, n), order="F")
        for r in range(reps):
            rc = lib.mambo_csa_reconstruct(e._h, x.ctypes.data, None, out.ctypes.data)
            assert rc == 0, lib.mambo_last_error()
            if ref is None:
                ref = out.copy()
            d = float(np.max(np.abs(out - ref)))
            if d > 0:
                bad += 1
                worst = max(worst, d)
        print(f"N={n}: {reps} chain evaluations, {bad} differ from the first (max abs dev {worst:.3e}, |h| max {np.abs(ref).max():.3e})", flush=True)
    del dT, dO
    torch.cuda.empty_cache()
