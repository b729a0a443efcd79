"""Bit-reproducibility with several evaluation lanes in flight (kernels of different lanes overlap on the device)."""
import os, sys, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mambo_b200 as mb
syn = importlib.import_module("quantummambo_jl_b200.synthetic")
n = int(sys.argv[1]); rounds = int(sys.argv[2]); nl = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dev = torch.device("cuda", 0)
T = syn.dense_sym_tbt(n, 0)
xs = np.stack([syn.random_x(n, i) for i in range(nl)])
with mb.CSAEngine(T) as e:
    lanes = [e] + [e.lane() for _ in range(nl - 1)]
    streams = [torch.cuda.Stream(dev) for _ in range(nl)]
    for ln, st in zip(lanes, streams):
        ln.set_stream(st.cuda_stream)
    d_x = torch.from_numpy(xs).to(dev); d_out = torch.empty(nl, e.L + 1, dtype=torch.float64, device=dev)
    # reference: one lane at a time
    for i in range(nl):
        lanes[i].eval_device(d_x[i].data_ptr(), d_out[i].data_ptr()); torch.cuda.synchronize()
    ref = d_out.clone()
    bad = 0
    for r in range(rounds):
        for i in range(nl):
            lanes[i].eval_device(d_x[i].data_ptr(), d_out[i].data_ptr())
        torch.cuda.synchronize()
        if not torch.equal(d_out, ref):
            bad += 1
    print(f"N={n}: {nl} lanes x {rounds} rounds, {bad} rounds differ from the one-at-a-time reference")
    for ln in lanes:
        ln.set_stream(None)
    for ln in lanes[1:]:
        ln.close()
