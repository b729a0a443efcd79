"""Which internal buffer differs first when the same evaluation is repeated?  (development aid for a nondeterminism hunt)"""
import os, sys, importlib, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mambo_b200 as mb
syn = importlib.import_module("quantummambo_jl_b200.synthetic")
lib = mb.load()
lib.mambo_debug_fetch.restype = C.c_longlong
lib.mambo_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong]
NAMES = ["U", "A~", "Ypieces", "E~panels", "F", "Spart", "part", "out", "aux"]
n = int(sys.argv[1]); reps = int(sys.argv[2])
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(n)
G = torch.randn((n, n, n, n), generator=g, dtype=torch.float64, device=dev)
G = G + G.permute(1, 0, 2, 3); G = G + G.permute(0, 1, 3, 2); G = (G + G.permute(2, 3, 0, 1)).mul_(1.0 / (n * n)).contiguous()
x = syn.random_x(n, 6)
with mb.CSAEngine.from_device(G.data_ptr(), n) as e:
    del G
    d_x = torch.from_numpy(x).to(dev); d_out = torch.empty(e.L + 1, dtype=torch.float64, device=dev)
    def fetch():
        bufs = []
        for w in range(9):
            ln = lib.mambo_debug_fetch(e._h, w, None, 0)
            a = np.empty(max(ln, 1))
            if ln > 0:
                lib.mambo_debug_fetch(e._h, w, a.ctypes.data, ln)
            bufs.append(a[:max(ln, 0)])
        return bufs
    e.eval_device(d_x.data_ptr(), d_out.data_ptr()); ref = fetch()
    print("mode", e.info()["mode_name"], "sizes", [len(b) for b in ref])
    bad = 0
    for r in range(reps):
        e.eval_device(d_x.data_ptr(), d_out.data_ptr())
        cur = fetch()
        diffs = [(NAMES[w], int(np.sum(cur[w] != ref[w])), float(np.max(np.abs(cur[w] - ref[w]))) if len(ref[w]) else 0.0) for w in range(9)]
        diffs = [d for d in diffs if d[1] > 0]
        if diffs:
            bad += 1
            print(f"rep {r}:", diffs, flush=True)
            w = [k for k in range(9) if np.any(cur[k] != ref[k])][0]
            idx = np.nonzero(cur[w] != ref[w])[0]
            print(f"   first differing buffer {NAMES[w]}: {len(idx)} entries, index range {idx.min()}..{idx.max()}, first few {idx[:8]}")
    print(f"N={n}: {bad} of {reps} repeats differ")
