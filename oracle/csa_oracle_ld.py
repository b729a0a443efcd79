"""CPU ORACLE, extended-precision ground truth (test infrastructure, not product code).

Two independent restatements that do NOT share arithmetic or libraries with ``oracle/csa_oracle.py``:

* ``cost_grad_factored_ld`` -- the factored O(N^5) form in ``numpy.longdouble`` (x87 80-bit, eps = 1.1e-19),
  with its own scaling-and-squaring Taylor ``expm`` and the block-triangular adjoint Frechet derivative -- no scipy, no
  LAPACK, no BLAS.  Usable to N ~ 20; it is the ground truth the FP64 oracle (and, through the golden vectors, the CUDA
  engine) must sit within 1e-12 of.
* ``cost_grad_literal_mp`` -- the reference's loop nests and formulas *literally* (unitaries.jl:31-46, 254-276, 314-320;
  fermionic.jl:20-32, 58-92; cost.jl:16-25; gradient.jl:5-41, 57-74, 76-146, 157-186, 218-284), evaluated with mpmath at
  40 digits including mpmath's own ``expm`` and complex ``eig`` for ``del_u_theta`` (gradient.jl:113-126).  N <= 4.

PARITY UNPINNED (see oracle/csa_oracle.py): the reference ships no vectors for this path and cannot run here; these
functions pin the restatement against itself through independent arithmetic, not against Julia output.
"""
from __future__ import annotations

import numpy as np

LD = np.longdouble


# ----------------------------------------------------------------------------------------------
# long-double factored form
# ----------------------------------------------------------------------------------------------

def _antisym(n, theta, dtype=LD):
    K = np.zeros((n, n), dtype=dtype)
    idx = 0
    for p in range(n):
        for q in range(p + 1, n):
            K[p, q] = theta[idx]
            K[q, p] = -theta[idx]
            idx += 1
    return K


def _cartan(n, lam, dtype=LD):
    L = np.zeros((n, n), dtype=dtype)
    idx = 0
    for i in range(n):
        for j in range(i + 1):
            L[i, j] = L[j, i] = lam[idx]
            idx += 1
    return L


def expm_ld(M):
    """exp(M) in long double: X = M / 2^s with ||X||_1 <= 1/2, Taylor to degree 32 (remainder < 1e-45), s squarings."""
    M = np.asarray(M, dtype=LD)
    n = M.shape[0]
    nrm = float(np.max(np.sum(np.abs(M), axis=0))) if n else 0.0
    s = 0
    while nrm > 0.5:
        nrm *= 0.5
        s += 1
    X = M / (LD(2) ** s)
    R = np.eye(n, dtype=LD)
    term = np.eye(n, dtype=LD)
    for k in range(1, 33):
        term = (term @ X) / LD(k)
        R = R + term
    for _ in range(s):
        R = R @ R
    return R


def cost_grad_factored_ld(x, tbt, obt=None, flavour: str = "CSA"):
    """Same quantities as csa_oracle.cost_grad_factored, every operation in long double.  Returns (cost, grad) as
    long-double scalars / arrays (round with float() / .astype(float) to compare)."""
    n = tbt.shape[0]
    cL = n * (n + 1) // 2
    x = np.asarray(x, dtype=LD)
    if flavour == "CSA":
        lam1, lam2, theta = None, x[:cL], x[cL:]
    else:
        lam1, lam2, theta = x[:n], x[n:n + cL], x[n + cL:]
    K = _antisym(n, theta)
    U = expm_ld(K)
    L = _cartan(n, lam2)
    A = (U[:, None, :] * U[None, :, :]).reshape(n * n, n)
    T = np.asarray(tbt, dtype=LD).reshape(n * n, n * n)
    D = (A @ L) @ A.T - T
    cost = np.sum(D * D)
    if flavour != "CSA":
        Dh = (U * lam1[None, :]) @ U.T - np.asarray(obt, dtype=LD)
        cost = cost + np.sum(Dh * Dh)
    E = D @ A
    Et = D.T @ A
    S = A.T @ E                                                   # one-sided, as gradient.jl:22-26
    gl = np.array([LD(2) * S[i, j] * (LD(2) if i != j else LD(1)) for i in range(n) for j in range(i + 1)], dtype=LD)
    F1 = (E @ L).reshape(n, n, n)
    F2 = (Et @ L).reshape(n, n, n)
    G = np.zeros((n, n), dtype=LD)
    for a in range(n):
        for b in range(n):
            G[a, b] = np.sum(U[:, b] * (F1[a, :, b] + F1[:, a, b] + F2[a, :, b] + F2[:, a, b]))
    g1 = None
    if flavour != "CSA":
        G = G + ((Dh + Dh.T) @ U) * lam1[None, :]
        g1 = np.array([LD(2) * np.sum(U[:, k][:, None] * Dh * U[:, k][None, :]) for k in range(n)], dtype=LD)
    M = np.zeros((2 * n, 2 * n), dtype=LD)
    M[:n, :n] = K.T
    M[n:, n:] = K.T
    M[:n, n:] = G
    Lf = expm_ld(M)[:n, n:]
    gt = np.array([LD(2) * (Lf[p, q] - Lf[q, p]) for p in range(n) for q in range(p + 1, n)], dtype=LD)
    parts = [gl, gt] if g1 is None else [g1, gl, gt]
    return cost, np.concatenate(parts) if parts[0].size + parts[-1].size else np.zeros(0, dtype=LD)


def reconstruct_ld(x, n: int, flavour: str = "CSA"):
    cL = n * (n + 1) // 2
    x = np.asarray(x, dtype=LD)
    off = 0 if flavour == "CSA" else n
    U = expm_ld(_antisym(n, x[off + cL:]))
    L = _cartan(n, x[off:off + cL])
    A = (U[:, None, :] * U[None, :, :]).reshape(n * n, n)
    W = ((A @ L) @ A.T).reshape(n, n, n, n)
    h = None if flavour == "CSA" else (U * x[:n][None, :]) @ U.T
    return h, W


# ----------------------------------------------------------------------------------------------
# mpmath literal form (N <= 4)
# ----------------------------------------------------------------------------------------------

def cost_grad_literal_mp(x, tbt, obt=None, flavour: str = "CSA", dps: int = 40):
    """The reference's formulas, loop for loop, in mpmath.  Returns (cost, grad) as Python floats rounded from `dps` digits."""
    import mpmath as mp
    mp.mp.dps = dps
    n = tbt.shape[0]
    cL, uL = n * (n + 1) // 2, n * (n - 1) // 2
    xs = [mp.mpf(float(v)) for v in x]
    if flavour == "CSA":
        lam1, lam2, theta = None, xs[:cL], xs[cL:]
    else:
        lam1, lam2, theta = xs[:n], xs[n:n + cL], xs[n + cL:]
    T = [[[[mp.mpf(float(tbt[p, q, r, s])) for s in range(n)] for r in range(n)] for q in range(n)] for p in range(n)]
    # unitaries.jl:31-46
    K = mp.zeros(n, n)
    idx = 0
    for i in range(n):
        for j in range(i + 1, n):
            K[i, j] = theta[idx]
            K[j, i] = -theta[idx]
            idx += 1
    U = mp.expm(K, method="taylor") if n > 1 else mp.matrix([[mp.mpf(1)]])
    # fermionic.jl:20-32 / gradient.jl:43-55
    Lm = mp.zeros(n, n)
    idx = 0
    for i in range(n):
        for j in range(i + 1):
            Lm[i, j] = lam2[idx]
            Lm[j, i] = lam2[idx]
            idx += 1
    # unitaries.jl:254-260 (identical values to the N^8 tbt_rotation of the CSA_SD representer, :266-272)
    rng = range(n)
    W = [[[[sum(U[a, l] * U[b, l] * U[c, m] * U[d, m] * Lm[l, m] for l in rng for m in rng) for d in rng] for c in rng] for b in rng] for a in rng]
    diff = [[[[W[p][q][r][s] - T[p][q][r][s] for s in rng] for r in rng] for q in rng] for p in rng]
    cost = sum(diff[p][q][r][s] ** 2 for p in rng for q in rng for r in rng for s in rng)             # cost.jl:16-25
    diff_ob = None
    if flavour != "CSA":
        h = [[sum(U[a, i] * U[b, i] * lam1[i] for i in rng) for b in rng] for a in rng]                  # unitaries.jl:314-320
        diff_ob = [[h[a][b] - mp.mpf(float(obt[a, b])) for b in rng] for a in rng]
        cost += sum(diff_ob[a][b] ** 2 for a in rng for b in rng)
    # gradient.jl:5-41
    gl = []
    for i in rng:
        for j in range(i + 1):
            cl = sum(2 * diff[p][q][r][s] * U[p, i] * U[q, i] * U[r, j] * U[s, j] for p in rng for q in rng for r in rng for s in rng)
            gl.append(cl * (2 if i != j else 1))
    # gradient.jl:57-74
    def wu(p, q, r, s, a, b):
        t_rs = sum(U[r, m] * U[s, m] * Lm[b, m] for m in rng)
        t_pq = sum(U[p, m] * U[q, m] * Lm[m, b] for m in rng)
        v = mp.mpf(0)
        if a == p: v += U[q, b] * t_rs
        if a == q: v += U[p, b] * t_rs
        if a == r: v += t_pq * U[s, b]
        if a == s: v += t_pq * U[r, b]
        return v
    # contraction of wu with diff first (exact reordering of the sums of gradient.jl:139-143): Gc[a][b] = sum diff . wu[..., a, b]
    Gc = [[sum(diff[p][q][r][s] * wu(p, q, r, s, a, b) for p in rng for q in rng for r in rng for s in rng) for b in rng] for a in rng]
    if diff_ob is not None:                                                                          # gradient.jl:218-231
        for a in rng:
            for b in rng:
                Gc[a][b] += sum(diff_ob[a][s] * U[s, b] * lam1[b] for s in rng) + sum(diff_ob[r][a] * U[r, b] * lam1[b] for r in rng)
    # gradient.jl:76-129 (exp branch :113-126): eigen of K, Phi with the hard 1e-8 switch
    gt = []
    if uL:
        D, O = mp.eig(mp.matrix(K))
        O = mp.matrix(O)
        OH = O.H
        # mpmath's eig does not normalise; the reference relies on unitary eigenvectors (O' as inverse): normalise columns
        for c in rng:
            nrm = mp.sqrt(sum(abs(O[r, c]) ** 2 for r in rng))
            for r in rng:
                O[r, c] = O[r, c] / nrm
        OH = O.H
        expD = mp.diag([mp.exp(d) for d in D])
        for i in rng:
            for j in range(i + 1, n):
                kappa = mp.zeros(n, n)
                kappa[i, j] = 1
                kappa[j, i] = -1
                I = OH * kappa * O
                for a in rng:
                    for b in rng:
                        dd = D[a] - D[b]
                        if abs(dd) > mp.mpf("1e-8"):
                            I[a, b] = I[a, b] * (mp.exp(dd) - 1) / dd
                ut = O * I * expD * OH
                gt.append(2 * sum(Gc[a][b] * mp.re(ut[a, b]) for a in rng for b in rng))
    g1 = []
    if flavour != "CSA":                                                                             # gradient.jl:157-186
        g1 = [sum(2 * diff_ob[r][s] * U[r, k] * U[s, k] for r in rng for s in rng) for k in rng]
    return float(cost), np.array([float(v) for v in (g1 + gl + gt)])
