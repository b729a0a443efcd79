import os, sys, ctypes as C
os.environ["MAMBO_CHAIN_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, mambo_b200 as mb
n = 100
T = mb.synthetic.dense_sym_tbt(n, 0)
with mb.CSAEngine(T) as e:
    for i in range(3):
        e.cost_grad(mb.synthetic.random_x(n, i))
    buf = (C.c_longlong * 96)()
    lib = mb.load()
    lib.mambo_csa_debug_stamps.argtypes = [C.c_void_p, C.c_void_p]
    assert lib.mambo_csa_debug_stamps(e.handle, buf) == 0
    a = np.array(buf[:]).reshape(16, 6)
    print("cycles per phase (issue loads, wait loads, dmma, combine+store, barrier):")
    for q in range(9):
        print(q, np.diff(a[q]), "total", a[q, 5] - a[q, 0])
    flat = np.array(buf[:])
    for q in range(8):
        b = flat[56 + 4 * q: 60 + 4 * q]
        print("barrier", q, "red issue", b[1] - b[0], "poll phase", b[2] - b[1], "polls", b[3])
    print("kernel start->squarings", flat[90] - flat[91], "squarings", a[8, 5] - flat[90], "tail", flat[92] - a[8, 5], "total", flat[92] - flat[91])
    print("poll loader (thread 0, last call): first pass", flat[56], "retry rounds", flat[57], "own total", flat[58], "incl. CTA sync", flat[59])
