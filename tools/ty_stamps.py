"""Development aid: per-step clock64 stamps of CTA 0 of fused::k_ty (MAMBO_TY_DEBUG=1) -> where a step's cycles go."""
import ctypes as C
import os
import sys

import numpy as np

os.environ["MAMBO_TY_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mambo_b200 as mb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
T = mb.synthetic.dense_sym_tbt(n, 0)
lib = mb.load()
with mb.CSAEngine(T) as e:
    for i in range(3):
        e.cost_grad(mb.synthetic.random_x(n, i))
    buf = np.zeros(16 * 64 * 4 + 160 * 4, dtype=np.int64)
    lib.mambo_csa_debug_ty_stamps.argtypes = [C.c_void_p, C.c_void_p]
    assert lib.mambo_csa_debug_ty_stamps(e.handle, buf.ctypes.data) == 0
st = buf[:16 * 64 * 4].reshape(16, 64, 4)
cta = buf[16 * 64 * 4:].reshape(160, 4)[:148]
t0 = st[0, 0, 0]
nsteps = 40
print("warp step  top->wait  wait(full)  compute   step_total   (cycles)   top offset vs warp0")
for w in (0, 4, 1, 5):
    for k in range(2, 14):
        top, bw, aw, end = st[w, k]
        nxt = st[w, k + 1, 0]
        print(f"{w:3d} {k:4d} {bw - top:9d} {aw - bw:10d} {end - aw:9d} {nxt - top:11d}              {top - st[0, k, 0]:8d}")
    print()
for w in range(8):
    tot = st[w, nsteps, 0] - st[w, 2, 0]
    wait = sum(st[w, k, 2] - st[w, k, 1] for k in range(2, nsteps))
    pre = sum(st[w, k, 1] - st[w, k, 0] for k in range(2, nsteps))
    comp = sum(st[w, k, 3] - st[w, k, 2] for k in range(2, nsteps))
    print(f"warp {w}: cycles/step {tot / (nsteps - 2):8.0f}  top->wait {pre / (nsteps - 2):7.0f}  full-wait {wait / (nsteps - 2):7.0f}  compute {comp / (nsteps - 2):7.0f}")
print()
w = 0
nst = int(np.count_nonzero(st[w, :, 0]))
print("warp 0: steps stamped", nst, " first top -> last end (cycles):", st[w, nst - 1, 3] - st[w, 0, 0])
print("per-step totals:", [int(st[w, k + 1, 0] - st[w, k, 0]) for k in range(nst - 1)])
print("step 0: top->wait", st[w, 0, 1] - st[w, 0, 0], "full-wait", st[w, 0, 2] - st[w, 0, 1], "compute", st[w, 0, 3] - st[w, 0, 2])

t00 = cta[:, 0].min()
print("per CTA (ns, globaltimer): kernel entry spread", int(cta[:, 0].max() - t00), " loop begin - entry (median)", int(np.median(cta[:, 1] - cta[:, 0])),
      " loop duration min/median/max", int((cta[:, 2] - cta[:, 1]).min()), int(np.median(cta[:, 2] - cta[:, 1])), int((cta[:, 2] - cta[:, 1]).max()),
      " last exit - first entry", int(cta[:, 2].max() - t00))
dur = cta[:, 2] - cta[:, 1]
order = np.argsort(dur)
print("slowest CTAs:", [(int(b), int(dur[b]), int(cta[b, 3])) for b in order[-6:]], " fastest:", [(int(b), int(dur[b]), int(cta[b, 3])) for b in order[:4]])
print("ns per step: min/median/max", float((dur / cta[:, 3]).min()), float(np.median(dur / cta[:, 3])), float((dur / cta[:, 3]).max()))
np.save(os.environ.get("TY_SAVE", "/tmp/ty_cta.npy"), (dur / cta[:, 3]))
