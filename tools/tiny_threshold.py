"""Where the single-kernel evaluation stops paying: device-resident evaluations/s of tiny::k_eval (MAMBO_TINY_MAX=16) against
the pipeline for N = 8..16 (dense-sym tensors)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import mambo_b200 as mb
import csa_oracle as o
REPS = 1000
for n in (8, 10, 11, 12, 13, 14, 15, 16):
    T = o.dense_sym_tbt(n, 0)
    xs = [o.random_x(n, s) for s in range(8)]
    row = []
    for env in ("16", "0"):
        os.environ["MAMBO_TINY_MAX"] = env
        with mb.CSAEngine(T) as e:
            d_x = [torch.tensor(x, dtype=torch.float64, device="cuda") for x in xs]
            d_o = torch.zeros(1 + len(xs[0]), dtype=torch.float64, device="cuda")
            st = torch.cuda.Stream()
            e.set_stream(st.cuda_stream)
            with torch.cuda.stream(st):
                for r in range(16): e.eval_device(d_x[r % 8].data_ptr(), d_o.data_ptr())
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st)
                for r in range(REPS): e.eval_device(d_x[r % 8].data_ptr(), d_o.data_ptr())
                b.record(st); b.synchronize()
            row.append(a.elapsed_time(b) * 1e3 / REPS)
            e.set_stream(None)
    print(f"N={n:2d}: single kernel {row[0]:6.1f} us   pipeline {row[1]:6.1f} us", flush=True)
