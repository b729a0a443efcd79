"""Bit-reproducibility of the device-resident evaluation: `reps` evaluations at one x, outputs compared on the device.
usage: flaky_out.py N reps  (env: MAMBO_TWO_GEMM, MAMBO_LOOKAHEAD, MAMBO_TY_RING ... as usual)"""
import os, sys, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mambo_b200 as mb
syn = importlib.import_module("quantummambo_jl_b200.synthetic")
n = int(sys.argv[1]); reps = int(sys.argv[2])
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(n)
G = torch.randn((n, n, n, n), generator=g, dtype=torch.float64, device=dev)
G = G + G.permute(1, 0, 2, 3); G = G + G.permute(0, 1, 3, 2); G = (G + G.permute(2, 3, 0, 1)).mul_(1.0 / (n * n)).contiguous()
x = syn.random_x(n, 6)
with mb.CSAEngine.from_device(G.data_ptr(), n) as e:
    del G
    d_x = torch.from_numpy(x).to(dev); d_out = torch.empty(e.L + 1, dtype=torch.float64, device=dev)
    e.eval_device(d_x.data_ptr(), d_out.data_ptr()); e.synchronize(); ref = d_out.clone()
    bad = 0; worst = 0.0
    for r in range(reps):
        e.eval_device(d_x.data_ptr(), d_out.data_ptr()); e.synchronize()
        if not torch.equal(d_out, ref):
            bad += 1
            worst = max(worst, float((d_out - ref).abs().max() / ref[1:].abs().max()))
    print(f"N={n}: {bad} of {reps} evaluations differ from the first (worst rel dev {worst:.2e})")
