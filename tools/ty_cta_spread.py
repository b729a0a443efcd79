"""Is the per-CTA speed spread of fused::k_ty systematic (per SM) or random?  Runs three handles, each stamped (MAMBO_TY_DEBUG=1),
and prints the correlation of ns-per-step by blockIdx between runs, plus the smid of every CTA would be needed to go further."""
import ctypes as C, os, sys
import numpy as np
os.environ["MAMBO_TY_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mambo_b200 as mb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
T = mb.synthetic.dense_sym_tbt(n, 0)
lib = mb.load()
lib.mambo_csa_debug_ty_stamps.argtypes = [C.c_void_p, C.c_void_p]
runs = []
with mb.CSAEngine(T) as e:
    for rep in range(6):
        e.cost_grad(mb.synthetic.random_x(n, rep))
        buf = np.zeros(16 * 64 * 4 + 160 * 4, dtype=np.int64)
        assert lib.mambo_csa_debug_ty_stamps(e.handle, buf.ctypes.data) == 0
        cta = buf[16 * 64 * 4:].reshape(160, 4)[:148]
        dur = (cta[:, 2] - cta[:, 1]) / np.maximum(cta[:, 3], 1)
        runs.append(dur.copy())
        print(f"run {rep}: ns/step min {dur.min():.1f} median {np.median(dur):.1f} max {dur.max():.1f}; loop max {int((cta[:,2]-cta[:,1]).max())} ns, median {int(np.median(cta[:,2]-cta[:,1]))}; steps {int(cta[:,3].min())}-{int(cta[:,3].max())}")
R = np.array(runs[2:])
print("correlation between runs (by blockIdx):")
print(np.round(np.corrcoef(R), 2))
m = R.mean(axis=0)
print("mean ns/step by blockIdx: min", m.min(), "max", m.max(), "std", m.std(), " run-to-run std per CTA (mean)", R.std(axis=0).mean())
print("slow CTAs (blockIdx):", np.argsort(m)[-12:], "fast:", np.argsort(m)[:12])
