"""Feasibility study for DESIGN.md 10 ("FP64 by INT8 slices on tcgen05", the Ozaki scheme): CPU emulation, exact integer arithmetic.

The hot product of one evaluation is Y = T~ A~ (M x M x N, M = N(N+1)/2), bound by the FP64 tensor pipe (37 TFLOP/s on a B200).  The
tensor cores run INT8 at a far higher rate, and a product of two int8 matrices with K <= 2^17 accumulates exactly in int32.  So:

    T~[i,j]  = 2^(e_i) * sum_k 2^(-b_k) t_k[i,j],   t_k int8      (one exponent per ROW i: constant along the contraction index j)
    A~[j,l]  = 2^(f_l) * sum_k 2^(-b_k) a_k[j,l],   a_k int8      (one exponent per COLUMN l)
    Y[i,l]  ~= 2^(e_i + f_l) * sum_{k + k' < D} 2^(-b_k - b_k') (t_k a_k')[i,l]        (every product exact; D = diagonals kept)

T~ is constant during a greedy step, so its slices are cut once per fragment (same HBM bytes as FP64 for 8 slices); A~ is re-cut every
evaluation (M x N, cheap).  This script measures what the truncation costs on the engine's own workload -- Y, cost and gradient
against the FP64 result, at a random point and at a nearly converged point of a planted tensor (where E = C - Y cancels) -- and prints
the number of int8 GEMMs per evaluation next to a simple time model.  Emulation only: nothing here runs on the GPU.

usage: proto_int8_slices.py [N ...]      (default 30 50)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import scipy.linalg as sla
import csa_oracle as o

FIRST_BITS, NEXT_BITS = 6, 7        # digits stay within [-64, 64]: int8


def slices(X, axis, S):
    """Signed-digit slices of X with one power-of-two scale per row (axis = 1) or column (axis = 0).  Returns (scale, [digits], [weights])."""
    amax = np.max(np.abs(X), axis=axis, keepdims=True)
    amax[amax == 0] = 1.0
    scale = 2.0 ** np.ceil(np.log2(amax))
    r = X / scale                                    # |r| <= 1, exact (power of two)
    digs, wts, shift = [], [], 0
    for k in range(S):
        bits = FIRST_BITS if k == 0 else NEXT_BITS
        shift += bits
        r = r * 2.0 ** bits
        q = np.rint(r)
        r = r - q                                    # |r| <= 1/2
        assert np.max(np.abs(q)) <= 64
        digs.append(q)
        wts.append(2.0 ** -shift)
    return scale, digs, wts


def y_from_slices(ts, as_, D):
    """sum over digit pairs k + k' < D of exact integer products (float64 holds them exactly: |entries| <= 64^2 * M < 2^53)."""
    (st, td, tw), (sa, ad, aw) = ts, as_
    acc = None
    pairs = 0
    for d in range(D - 1, -1, -1):                   # small terms first
        for k in range(min(d, len(td) - 1), -1, -1):
            kp = d - k
            if kp >= len(ad):
                continue
            term = (td[k] @ ad[kp]) * (tw[k] * aw[kp])
            acc = term if acc is None else acc + term
            pairs += 1
    return st * acc * sa, pairs


def cost_grad_from_y(x, packed, n, Y):
    """oracle.cost_grad_packed with the hot product supplied by the caller."""
    Tt, pa, pq, sw, tnorm2 = packed
    cL = o.cartan_2b_num_params(n)
    lam, theta = np.asarray(x[:cL]), np.asarray(x[cL:])
    K = o.construct_anti_symmetric(n, theta)
    U = sla.expm(K)
    L = o.get_cartan_matrix(n, lam)
    A = sw[:, None] * U[pa] * U[pq]
    Z = A.T @ Y
    Fp = Y @ L
    o_ = np.sum(U * U, axis=0)
    d = o_ * o_
    t = (L * L) @ d
    wn = float(d @ t)
    S = d[:, None] * L * d[None, :] - Z
    cost = tnorm2 + wn - 2.0 * float(np.sum(L * Z))
    c = np.where(pa == pq, 2.0, np.sqrt(2.0))[:, None]
    Gg = np.zeros((n, n))
    np.add.at(Gg, pa, c * U[pq] * Fp)
    off = pa != pq
    np.add.at(Gg, pq[off], c[off] * U[pa[off]] * Fp[off])
    G = 4.0 * U * (o_ * t)[None, :] - 2.0 * Gg
    gl = np.array([2.0 * S[i, j] * (2.0 if i != j else 1.0) for i in range(n) for j in range(i + 1)])
    Lf = o._frechet_adjoint(K, G)
    iu = np.triu_indices(n, 1)
    return cost, np.concatenate([gl, 2.0 * (Lf[iu] - Lf.T[iu])]), A


def study(n):
    M = n * (n + 1) // 2
    print(f"== N = {n}: M = {M}, one int8 GEMM = 2 M^2 N = {2 * M * M * n / 1e9:.3f} Gop; FP64 DMMA product at 30 TFLOP/s: {2 * M * M * n / 30e12 * 1e6:.1f} us")
    cases = []
    T = o.dense_sym_tbt(n, 0)
    cases.append(("dense-sym, random x", T, o.random_x(n, 1)))
    Tp, xs = o.planted_tbt(n, 1, 3, sigma=1e-9)
    cases.append(("planted + 1e-9 noise, x = planted + 1e-6", Tp, xs[0] + 1e-6 * np.random.default_rng(1).standard_normal(xs[0].shape)))
    for label, Tn, x in cases:
        packed = o.pack_tbt(Tn)
        Tt = packed[0]
        cL = o.cartan_2b_num_params(n)
        U = sla.expm(o.construct_anti_symmetric(n, x[cL:]))
        A = packed[3][:, None] * U[packed[1]] * U[packed[2]]
        Y64 = Tt @ A
        c64, g64, _ = cost_grad_from_y(x, packed, n, Y64)
        cref, gref = o.cost_grad_factored(x, Tn)                      # explicit D = W - T: no cancellation in the cost
        print(f"  -- {label}: cost {cref:.6e}; FP64 single-GEMM path vs explicit: grad rel {np.max(np.abs(g64 - gref)) / np.max(np.abs(gref)):.1e}")
        t0 = time.perf_counter()
        ts, as_ = slices(Tt, 1, 8), slices(A, 0, 8)
        Yfull, _ = y_from_slices(ts, as_, 15)
        print(f"     all 64 digit pairs vs FP64 product: max |dY| / max |Y| = {np.max(np.abs(Yfull - Y64)) / np.max(np.abs(Y64)):.1e}")
        for D in (4, 5, 6, 7, 8, 9, 10):
            Yd, pairs = y_from_slices(ts, as_, D)
            ey = np.max(np.abs(Yd - Yfull)) / np.max(np.abs(Yfull))
            cd, gd, _ = cost_grad_from_y(x, packed, n, Yd)
            eg = np.max(np.abs(gd - gref)) / np.max(np.abs(gref))
            ec = abs(cd - cref) / max(abs(cref), 1e-300)
            t_model = pairs * 2 * M * M * n / (0.6 * 4.5e15) * 1e6
            print(f"     D = {D:2d}: {pairs:2d} int8 GEMMs ({6 + 7 * (D - 1)} bits)  dY {ey:.1e}  cost rel {ec:.1e}  grad rel {eg:.1e}   "
                  f"model time at 60 % of 4.5 Pop/s: {t_model:6.1f} us")
        print(f"     ({time.perf_counter() - t0:.1f} s of emulation)")


if __name__ == "__main__":
    for n in [int(a) for a in sys.argv[1:]] or [30, 50]:
        study(n)
