"""CPU ORACLE (test infrastructure, not product code).

A numpy restatement of QuantumMAMBO.jl's CSA / CSA_SD fragment cost + analytic gradient, the greedy
drivers around them and the 1-norm post-processing.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this file; the product path
(``quantummambo.jl_b200``) never does.

PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or fixtures for this path
(``/root/reference/test/runtests.jl:1-5`` runs only Lanczos tests) and it cannot be executed here (no
``julia`` binary).  Substitute pins, all exercised by ``tests/test_oracle.py``:
  (1) the *literal* transcription below (loop nests and formulas 1:1 with the Julia source) agrees with
      the *factored* O(N^5) form to ~1e-13 for N in 3..8;
  (2) the analytic gradient agrees with central finite differences of the literal cost for
      pair-swap-symmetric targets;
  (3) analytic known answers: planted tensor => cost 0 / gradient 0; theta = 0 => U = I and
      W[i,i,j,j] = Lambda_ij.

Conventions (all arrays are numpy, indices 0-based; the Julia code is 1-based, column-major):
  * CSA     x = [lambda (cL = N(N+1)/2) ; theta (uL = N(N-1)/2)]
  * CSA_SD  x = [lambda1 (N) ; lambda2 (cL) ; theta (uL)]
  * lambda order: lower-triangular row-major (1,1),(2,1),(2,2),(3,1)...   fermionic.jl:23-30, gradient.jl:47-53
  * theta order : upper-triangular row-major (1,2),(1,3)..(1,N),(2,3)...  unitaries.jl:36-43
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

# ----------------------------------------------------------------------------------------------
# parameter counting / packing                                       (fermionic.jl:50-56, unitaries.jl:48-51)
# ----------------------------------------------------------------------------------------------

def cartan_2b_num_params(N: int) -> int:
    """fermionic.jl:50-52"""
    return N * (N + 1) // 2


def cartan_SD_num_params(N: int) -> int:
    """fermionic.jl:54-56"""
    return N + cartan_2b_num_params(N)


def real_orbital_rotation_num_params(N: int) -> int:
    """unitaries.jl:48-51"""
    return N * (N - 1) // 2


def get_rot_indices(n: int, idx: int):
    """unitaries.jl:418-424 -- 1-based idx -> 1-based (i, j), i < j.  Returned 0-based here."""
    i = int(np.ceil(((2 * n - 1) - np.sqrt((1 - 2 * n) ** 2 - 8 * idx)) / 2))
    i_prev = i - 1
    j = i + int(idx - i_prev * n + i_prev * (i_prev + 1) / 2)
    return i - 1, j - 1


def construct_anti_symmetric(n: int, params) -> np.ndarray:
    """unitaries.jl:391-415 (get_anti_symmetric + construct_anti_symmetric); same matrix as the Ulog
    built in unitaries.jl:31-45."""
    K = np.zeros((n, n))
    idx = 0
    for p in range(n):
        for q in range(p + 1, n):
            K[p, q] = params[idx]
            K[q, p] = -params[idx]
            idx += 1
    return K


def one_body_unitary(theta, n: int) -> np.ndarray:
    """unitaries.jl:31-46: U = exp(Ulog).  Julia's ``exp`` is Higham-2005 Pade scaling & squaring;
    scipy's expm (Al-Mohy/Higham 2009) agrees with it to a few ulp for these well-conditioned inputs."""
    return sla.expm(construct_anti_symmetric(n, theta))


def get_cartan_matrix(n: int, lam) -> np.ndarray:
    """gradient.jl:43-55 / fermionic.jl:20-32: full symmetric lambda matrix from lower-tri row-major."""
    L = np.zeros((n, n))
    idx = 0
    for i in range(n):
        for j in range(i + 1):
            L[i, j] = lam[idx]
            L[j, i] = lam[idx]
            idx += 1
    return L


def cartan_2b_to_tbt(lam, n: int) -> np.ndarray:
    """fermionic.jl:20-32"""
    tbt = np.zeros((n, n, n, n))
    idx = 0
    for i in range(n):
        for j in range(i + 1):
            tbt[i, i, j, j] = lam[idx]
            tbt[j, j, i, i] = lam[idx]
            idx += 1
    return tbt


def split_x(x, n: int, flavour: str):
    """fermionic.jl:413-427 (CSA_x_to_F_FRAG / CSA_SD_x_to_F_FRAG slicing)."""
    cL = cartan_2b_num_params(n)
    if flavour == "CSA":
        return None, np.asarray(x[:cL]), np.asarray(x[cL:])
    return np.asarray(x[:n]), np.asarray(x[n:n + cL]), np.asarray(x[n + cL:])


# ----------------------------------------------------------------------------------------------
# LITERAL restatement (N^6 / N^8 loops through einsum, N^6 scratch) -- small N only
# ----------------------------------------------------------------------------------------------

def cartan_tbt_rotation(U, tbt):
    """unitaries.jl:254-260: rotated[a,b,c,d] = U[a,l] U[b,l] U[c,m] U[d,m] tbt[l,l,m,m]."""
    n = U.shape[0]
    lam = np.einsum("llmm->lm", tbt)
    return np.einsum("al,bl,cm,dm,lm->abcd", U, U, U, U, lam, optimize=False) if n <= 5 else \
        np.einsum("al,bl,cm,dm,lm->abcd", U, U, U, U, lam, optimize=True)


def tbt_rotation(U, tbt):
    """unitaries.jl:266-272 (the general 4-index rotation used by the CSA_SD representer)."""
    return np.einsum("al,bk,cm,dp,lkmp->abcd", U, U, U, U, tbt, optimize=True)


def obt_rotation(U, obt):
    """unitaries.jl:314-320"""
    return np.einsum("ai,bj,ij->ab", U, U, obt)


def to_OP(x, n: int, flavour: str = "CSA"):
    """to_OP(CSA_x_to_F_FRAG(x)) -- fermionic.jl:1-8, 74-92, 58-72.  Returns (obt or None, tbt)."""
    lam1, lam2, theta = split_x(x, n, flavour)
    U = one_body_unitary(theta, n)
    tbt = cartan_2b_to_tbt(lam2, n)
    if flavour == "CSA":
        return None, cartan_tbt_rotation(U, tbt)
    obt = np.diag(lam1)
    return obt_rotation(U, obt), tbt_rotation(U, tbt)


def L2_partial_cost(tbt_a, tbt_b=None, obt_a=None, obt_b=None) -> float:
    """cost.jl:16-55 -- squared Frobenius norm over the filled (one- and) two-body tensors."""
    c = float(np.sum((tbt_a - (0.0 if tbt_b is None else tbt_b)) ** 2))
    if obt_a is not None:
        c += float(np.sum((obt_a - (0.0 if obt_b is None else obt_b)) ** 2))
    return c


def cost_literal(x, tbt, obt=None, flavour: str = "CSA") -> float:
    """decompose.jl:96-99 (CSA) / :221-224 (CSA_SD)."""
    n = tbt.shape[0]
    h, W = to_OP(x, n, flavour)
    if flavour == "CSA":
        return L2_partial_cost(tbt, W)
    return L2_partial_cost(tbt, W, obt, h)


def del_wr_lr(n, l, m, U):
    """gradient.jl:5-11"""
    return np.einsum("p,q,r,s->pqrs", U[:, l], U[:, l], U[:, m], U[:, m])


def grad_lr(n, U, diff):
    """gradient.jl:13-41 (grad_lr_ij folded in; the per-(i,j) expm recomputation is the same U)."""
    g = np.zeros(cartan_2b_num_params(n))
    idx = 0
    cw = 2 * diff
    for i in range(n):
        for j in range(i + 1):
            cl = np.sum(cw * del_wr_lr(n, i, j, U))
            if i != j:
                cl *= 2
            g[idx] = cl
            idx += 1
    return g


def del_w_u(n, U, L):
    """gradient.jl:57-74 -- dW[p,q,r,s]/dU[a,b] as a dense N^6 array (4 product-rule terms)."""
    wu = np.zeros((n,) * 6)
    UL = np.einsum("rm,sm,bm->rsb", U, U, L)          # sum_m U[r,m] U[s,m] L[b,m]
    ULl = np.einsum("pl,ql,lb->pqb", U, U, L)         # sum_l U[p,l] U[q,l] L[l,b]
    for p in range(n):
        wu[p, :, :, :, p, :] += np.einsum("qb,rsb->qrsb", U, UL)
    for q in range(n):
        wu[:, q, :, :, q, :] += np.einsum("pb,rsb->prsb", U, UL)
    for r in range(n):
        wu[:, :, r, :, r, :] += np.einsum("pqb,sb->pqsb", ULl, U)
    for s in range(n):
        wu[:, :, :, s, s, :] += np.einsum("pqb,rb->pqrb", ULl, U)
    return wu


def del_u_theta(n, u_params, idx0):
    """gradient.jl:113-126 (exp branch): dU/dtheta_idx through the eigendecomposition of K, with the
    hard 1e-8 switch of gradient.jl:119-121.  idx0 is 0-based."""
    i, j = get_rot_indices(n, idx0 + 1)
    kappa = np.zeros((n, n))
    kappa[i, j] = 1
    kappa[j, i] = -1
    K = construct_anti_symmetric(n, u_params)
    D, O = np.linalg.eig(K)
    I = O.conj().T @ kappa @ O
    for a in range(n):
        for b in range(n):
            if abs(D[a] - D[b]) > 1e-8:
                I[a, b] *= (np.exp(D[a] - D[b]) - 1) / (D[a] - D[b])
    expD = np.diag(np.exp(D))
    return np.real(O @ I @ expD @ O.conj().T)


def grad_theta(n, U, L, u_params, diff, diff_ob=None, lam1=None):
    """gradient.jl:131-146 (CSA) and :266-284 (CSA_SD adds the one-body term through del_h_u :218-231)."""
    wu = del_w_u(n, U, L)
    opnum = real_orbital_rotation_num_params(n)
    g = np.zeros(opnum)
    if diff_ob is not None:
        hu = np.zeros((n, n, n, n))
        for r in range(n):
            hu[r, :, r, :] += U * lam1[None, :]
        for s in range(n):
            hu[:, s, s, :] += U * lam1[None, :]
    for k in range(opnum):
        ut = del_u_theta(n, u_params, k)
        w_theta = np.einsum("pqrsab,ab->pqrs", wu, ut)
        val = np.sum(diff * w_theta)
        if diff_ob is not None:
            h_theta = np.einsum("pqab,ab->pq", hu, ut)
            val += np.sum(diff_ob * h_theta)
        g[k] = 2 * val
    return g


def gradient_literal(x, tbt, obt=None, flavour: str = "CSA"):
    """grad! closures: decompose.jl:101-105 -> gradient.jl:148-151 ; decompose.jl:226-231 -> gradient.jl:287-290."""
    n = tbt.shape[0]
    lam1, lam2, theta = split_x(x, n, flavour)
    U = one_body_unitary(theta, n)
    L = get_cartan_matrix(n, lam2)
    h, W = to_OP(x, n, flavour)
    diff = W - tbt
    if flavour == "CSA":
        return np.concatenate([grad_lr(n, U, diff), grad_theta(n, U, L, theta, diff)])
    diff_ob = h - obt
    g1 = np.array([np.sum(2 * diff_ob * np.outer(U[:, k], U[:, k])) for k in range(n)])   # gradient.jl:157-186
    return np.concatenate([g1, grad_lr(n, U, diff), grad_theta(n, U, L, theta, diff, diff_ob, lam1)])


# ----------------------------------------------------------------------------------------------
# FACTORED restatement (O(N^5), BLAS) -- the oracle for N = 50 / 100 / 152 and the CPU baseline
# ----------------------------------------------------------------------------------------------

def _frechet_adjoint(K, G):
    """Lf = top-right block of exp([[K^T, G],[0, K^T]]) = Dexp(K^T)[G]; <G, dexp_K(kappa)> = <Lf, kappa>."""
    n = K.shape[0]
    M = np.zeros((2 * n, 2 * n))
    M[:n, :n] = K.T
    M[n:, n:] = K.T
    M[:n, n:] = G
    return sla.expm(M)[:n, n:]


def cost_grad_factored(x, tbt, obt=None, flavour: str = "CSA", need_grad: bool = True, U=None):
    """Same numbers as cost_literal / gradient_literal via
         A = U (.) U (N^2 x N),  W = A L A^T,  D = W - T,  E = D A,  E' = D^T A,
         S = A^T E,  F = E L,  G = sum of the 4 product-rule terms,  grad_theta from the adjoint Frechet
         derivative of expm.
    The lambda gradient is the reference's one-sided expression (gradient.jl:22-26): S uses E only."""
    n = tbt.shape[0]
    lam1, lam2, theta = split_x(x, n, flavour)
    K = construct_anti_symmetric(n, theta)
    if U is None:
        U = sla.expm(K)
    L = get_cartan_matrix(n, lam2)
    # row (p,q) of A:  A[p,q,l] = U[p,l] U[q,l];  C-order (p slow) -- only the association with T matters.
    A = (U[:, None, :] * U[None, :, :]).reshape(n * n, n)
    T = np.ascontiguousarray(tbt).reshape(n * n, n * n)          # T[(p,q),(r,s)] in C order
    D = (A @ L) @ A.T - T
    cost = float(np.sum(D * D))
    if flavour != "CSA":
        h = (U * lam1[None, :]) @ U.T
        Dh = h - obt
        cost += float(np.sum(Dh * Dh))
    if not need_grad:
        return cost, None
    E = D @ A                                                    # E[(p,q), j]
    Et = D.T @ A                                                 # E'[(r,s), j]
    S = A.T @ E
    cL = cartan_2b_num_params(n)
    gl = np.zeros(cL)
    idx = 0
    for i in range(n):
        for j in range(i + 1):
            gl[idx] = 2.0 * S[i, j] * (2.0 if i != j else 1.0)
            idx += 1
    F1 = (E @ L).reshape(n, n, n)                                # [p,q,b]
    F2 = (Et @ L).reshape(n, n, n)
    G = np.einsum("qb,aqb->ab", U, F1) + np.einsum("qb,qab->ab", U, F1) \
        + np.einsum("sb,asb->ab", U, F2) + np.einsum("sb,sab->ab", U, F2)
    g1 = None
    if flavour != "CSA":
        G = G + ((Dh + Dh.T) @ U) * lam1[None, :]
        g1 = 2.0 * np.einsum("rk,rs,sk->k", U, Dh, U)
    Lf = _frechet_adjoint(K, G)
    gt = np.zeros(real_orbital_rotation_num_params(n))
    idx = 0
    for p in range(n):
        for q in range(p + 1, n):
            gt[idx] = 2.0 * (Lf[p, q] - Lf[q, p])
            idx += 1
    if flavour == "CSA":
        return cost, np.concatenate([gl, gt])
    return cost, np.concatenate([g1, gl, gt])


def pack_tbt(tbt):
    """The engine's packed residual for an 8-fold symmetric tensor: rows / columns are the pairs p <= q in the
    order r = q(q+1)/2 + p, scaled by sqrt(2) off the diagonal so that weighted sums over unique pairs become plain sums.
    Returns (Tt [M, M], pa, pq, sw, |Tt|^2)."""
    n = tbt.shape[0]
    pa, pq = [], []
    for q in range(n):
        for a in range(q + 1):
            pa.append(a)
            pq.append(q)
    pa, pq = np.array(pa), np.array(pq)
    sw = np.where(pa == pq, 1.0, np.sqrt(2.0))
    Tt = np.ascontiguousarray(tbt[pa, pq][:, pa, pq] * sw[:, None] * sw[None, :])
    return Tt, pa, pq, sw, float(np.sum(Tt * Tt))


def cost_grad_packed(x, packed, n: int):
    """CPU restatement of the ALGORITHM the CUDA engine runs for 8-fold symmetric tensors (CSA flavour): packed rows and the
    lean single-GEMM formulation -- Y = T~ A~ is the only M x M x N product, everything else is O(M N^2) or N x N closed
    forms.  Same numbers as cost_grad_factored to rounding; 8x fewer flops.  bench.py times it beside the
    factored port so that the GPU/CPU ratio can be split into "algorithm" and "hardware"."""
    Tt, pa, pq, sw, tnorm2 = packed
    cL = cartan_2b_num_params(n)
    lam, theta = np.asarray(x[:cL]), np.asarray(x[cL:])
    K = construct_anti_symmetric(n, theta)
    U = sla.expm(K)
    L = get_cartan_matrix(n, lam)
    A = sw[:, None] * U[pa] * U[pq]                               # A~ [M, n]
    Y = Tt @ A                                                    # the hot product
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
    Lf = _frechet_adjoint(K, G)
    iu = np.triu_indices(n, 1)
    gt = 2.0 * (Lf[iu] - Lf.T[iu])
    return cost, np.concatenate([gl, gt])


def partial_factored(x, tbt, row_begin: int, row_end: int):
    """Row-sharded first half of cost_grad_factored for a (pq)<->(rs)-symmetric tensor:
    [cost, S (n x n), G (n x n)] restricted to rows [row_begin, row_end) of the n^2 x n^2 residual, rows indexed
    as in the Julia memory layout (row (p,q) -> p + n q).  Summing the partials of a disjoint cover of the rows
    gives the unsharded vector."""
    n = tbt.shape[0]
    cL = cartan_2b_num_params(n)
    off = len(x) - cL - real_orbital_rotation_num_params(n)
    lam2, theta = np.asarray(x[off:off + cL]), np.asarray(x[off + cL:])
    U = one_body_unitary(theta, n)
    L = get_cartan_matrix(n, lam2)
    Af = (U[None, :, :] * U[:, None, :]).reshape(n * n, n)       # row index p + n q  (q slow)
    T = np.asarray(tbt).reshape(n * n, n * n, order="F")          # T[(pq),(rs)] with p fastest
    rows = np.arange(row_begin, min(row_end, n * n))
    D = (Af[rows] @ L) @ Af.T - T[rows]
    cost = float(np.sum(D * D))
    E = D @ Af
    S = Af[rows].T @ E
    F = E @ L
    G = np.zeros((n, n))
    for k, r in enumerate(rows):
        a, q = r % n, r // n
        G[a] += 2.0 * U[q] * F[k]
        G[q] += 2.0 * U[a] * F[k]
    return np.concatenate([[cost], S.ravel(), G.ravel()])


def finish_factored(part, x, n: int, flavour: str = "CSA", obt=None):
    """Second half: one-body terms, adjoint Frechet derivative, gradient assembly (replicated on every rank)."""
    lam1, lam2, theta = split_x(x, n, flavour)
    K = construct_anti_symmetric(n, theta)
    U = sla.expm(K)
    cost = float(part[0])
    S = np.asarray(part[1:1 + n * n]).reshape(n, n)
    G = np.asarray(part[1 + n * n:]).reshape(n, n).copy()
    g1 = None
    if flavour != "CSA":
        Dh = (U * lam1[None, :]) @ U.T - obt
        cost += float(np.sum(Dh * Dh))
        G += ((Dh + Dh.T) @ U) * lam1[None, :]
        g1 = 2.0 * np.einsum("rk,rs,sk->k", U, Dh, U)
    gl = np.array([2.0 * S[i, j] * (2.0 if i != j else 1.0) for i in range(n) for j in range(i + 1)])
    Lf = _frechet_adjoint(K, G)
    gt = np.array([2.0 * (Lf[p, q] - Lf[q, p]) for p in range(n) for q in range(p + 1, n)])
    return cost, (np.concatenate([gl, gt]) if g1 is None else np.concatenate([g1, gl, gt]))


def reconstruct_factored(x, n: int, flavour: str = "CSA"):
    """to_OP(frag) through the factored form (any N)."""
    lam1, lam2, theta = split_x(x, n, flavour)
    U = one_body_unitary(theta, n)
    L = get_cartan_matrix(n, lam2)
    A = (U[:, None, :] * U[None, :, :]).reshape(n * n, n)
    W = ((A @ L) @ A.T).reshape(n, n, n, n)
    if flavour == "CSA":
        return None, W
    return (U * lam1[None, :]) @ U.T, W


# ----------------------------------------------------------------------------------------------
# 1-norms (lcu.jl:139-208, 239-267, fermionic.jl:457-479, 527-538)
# ----------------------------------------------------------------------------------------------

def CSA_L1(lam, n: int, spin_orb: bool = False) -> float:
    """lcu.jl:139-182"""
    l1 = 0.0
    idx = 0
    for i in range(n):
        for j in range(i + 1):
            if spin_orb:
                if i != j:
                    l1 += 0.5 * abs(lam[idx])
            else:
                l1 += (2.0 if i != j else 0.5) * abs(lam[idx])
            idx += 1
    return l1


def ob_correction(tbt, spin_orb: bool = False):
    """fermionic.jl:457-479"""
    n = tbt.shape[0]
    s = sum(tbt[:, :, r, r] for r in range(n))
    return s if spin_orb else 2 * s


def one_body_L1(obt, tbt, spin_orb: bool = False) -> float:
    """lcu.jl:262-267 with to_OBF (fermionic.jl:527-538) and OBF_L1 (lcu.jl:239-260)."""
    M = obt + ob_correction(tbt, spin_orb)
    D = np.linalg.eigvalsh((M + M.T) / 2) if np.allclose(M, M.T) else np.real(np.linalg.eigvals(M))
    return float(np.sum(np.abs(D))) * (0.5 if spin_orb else 1.0)


# ----------------------------------------------------------------------------------------------
# initial guess (guesses.jl:3-84, unitaries.jl:130-147)
# ----------------------------------------------------------------------------------------------

def SOn_unitary_to_params(U):
    """unitaries.jl:130-147: upper-triangular entries of real(log U)."""
    n = U.shape[0]
    Ulog = np.real(sla.logm(U))
    return np.array([Ulog[i, j] for i in range(n) for j in range(i + 1, n)])


def tbt_svd_1st(tbt, tiny: float = 1e-12, fix_sign: bool = False):
    """guesses.jl:51-84 + SVD_to_CSA (:3-49).  Returns (lambda (cL), theta (uL)).  Note K1 of SURVEY.md:
    SVD_to_CSA fills lambda in *upper*-triangular order (guesses.jl:12-18) although every consumer reads it
    lower-triangular -- reproduced on purpose."""
    n = tbt.shape[0]
    N2 = n * n
    # Julia: Symmetric(reshape(tbt,(N,N))) uses the upper triangle of the column-major matrix view.
    M = np.reshape(tbt, (N2, N2), order="F")
    Msym = np.triu(M) + np.triu(M, 1).T
    lamv, Uv = np.linalg.eigh(Msym)
    ind = np.argsort(np.abs(lamv), kind="stable")[::-1]
    lamv = lamv[ind]
    Uv = Uv[:, ind]
    full_l = np.reshape(Uv[:, 0], (n, n), order="F")
    if fix_sign:
        # The sign of a LAPACK eigenvector is a convention, and through K1 (coefficients filled in upper- but read in
        # lower-triangular order) the resulting fragment depends on it.  fix_sign pins the convention the CUDA library
        # documents (largest-magnitude entry positive) so that the two can be compared entry by entry.
        big = np.argmax(np.abs(full_l.reshape(-1, order="F")))
        if full_l.reshape(-1, order="F")[big] < 0:
            full_l = -full_l
    cur_l = np.triu(full_l) + np.triu(full_l, 1).T
    lead = lamv[0]
    if np.sum(np.abs(cur_l - full_l)) > tiny:
        if np.sum(np.abs(full_l + full_l.T)) > tiny:
            raise ValueError("SVD operator is neither Hermitian or anti-Hermitian")
        cur = 1j * full_l
        cur = np.triu(cur) + np.triu(cur, 1).conj().T
        lead = -lead
        w, Ul = np.linalg.eigh(cur)
    else:
        w, Ul = np.linalg.eigh(cur_l)
    coeffs = np.array([w[i] * w[j] for i in range(n) for j in range(i, n)]) * lead
    return coeffs, SOn_unitary_to_params(Ul)


# ----------------------------------------------------------------------------------------------
# double factorisation (decompose.jl:297-436): SURVEY 8(f)3
# ----------------------------------------------------------------------------------------------

def _first_max_positive(v):
    """sign convention shared with the CUDA library: the largest-magnitude entry (the first one) is positive"""
    flat = v.reshape(-1, order="F")
    k = int(np.argmax(np.abs(flat)))
    return -v if flat[k] < 0 else v


def DF_decomposition(tbt, tol: float = 1e-8, tiny: float = 1e-12, fix_sign: bool = False):
    """decompose.jl:297-398 with do_Givens = DF_GIVENS = false (config.jl:14; SVD_tol = 1e-8, SVD_tiny = 1e-12, config.jl:21-22).

    Returns a list of fragments (coeff, omega (n), U (n x n), L (n x n)), ordered by decreasing |eigenvalue| of the
    N^2 x N^2 view: to_OP(frag) = coeff * sum_lm omega_l omega_m U[:,l] U[:,l]' (x) U[:,m] U[:,m]' = coeff * L (x) L with
    L = U diag(omega) U' (fermionic.jl:94-116).  The anti-Hermitian branch (:346-352: eigen of Hermitian(1im * L), coefficient
    negated) returns complex U, as the reference does.

    fix_sign: LAPACK leaves the sign of every eigenvector to chance, and DF_based_greedy depends on it (Vm = Um * U1dag contracts
    the EIGENVECTOR indices of both frames, decompose.jl:421).  fix_sign pins the convention the CUDA library documents (largest-
    magnitude entry of every N^2-eigenvector and of every column of U positive) so that the two can be compared entry by entry."""
    n = tbt.shape[0]
    N2 = n * n
    full = np.reshape(tbt, (N2, N2), order="F")                       # :317
    res = np.triu(full) + np.triu(full, 1).T                          # Symmetric(tbt_full): the upper triangle (:318)
    if np.sum(np.abs(full - res)) > tiny:                             # :319-322 (a general `eigen` in the reference)
        raise ValueError("non-symmetric two-body tensor: the reference falls back to a non-symmetric eigen here (not restated)")
    lam, U = np.linalg.eigh(res)                                      # :326-330
    ind = np.argsort(np.abs(lam), kind="stable")[::-1]                # :332
    lam, U = lam[ind], U[:, ind]
    num_ops = N2                                                      # :336-345
    for i in range(N2):
        if abs(lam[i]) < tol:
            num_ops = i
            break
    frags = []
    for i in range(num_ops):                                          # :352-381
        full_l = np.reshape(U[:, i], (n, n), order="F")
        if fix_sign:
            full_l = _first_max_positive(full_l)
        cur_l = np.triu(full_l) + np.triu(full_l, 1).T                # Symmetric(full_l)
        coeff = lam[i]
        if np.sum(np.abs(cur_l - full_l) ** 2) > tiny:
            if np.sum(np.abs(full_l + full_l.T)) > tiny:
                raise ValueError(f"DF fragment {i + 1} is neither Hermitian or anti-Hermitian!")
            cur = 1j * full_l
            cur_l = np.triu(cur) + np.triu(cur, 1).conj().T           # Hermitian(1im * full_l)
            coeff = -coeff
        w, Ul = np.linalg.eigh(cur_l)
        if fix_sign and not np.iscomplexobj(Ul):
            Ul = np.stack([_first_max_positive(Ul[:, k]) for k in range(n)], axis=1)
        frags.append((coeff, w, Ul, cur_l))
    return frags


def DF_to_tbt(frags, n: int):
    """sum of to_OP(frag).mbts[3] over DF fragments (fermionic.jl:94-116, the debug branch decompose.jl:383-387)."""
    tbt = np.zeros((n, n, n, n), dtype=complex if any(np.iscomplexobj(f[3]) for f in frags) else float)
    for coeff, w, Ul, _ in frags:
        Lm = (Ul * w) @ Ul.conj().T if np.iscomplexobj(Ul) else (Ul * w) @ Ul.T
        tbt += coeff * np.einsum("ab,cd->abcd", Lm, Lm)
    return tbt


def DF_based_greedy(tbt, tol: float = 1e-8, fix_sign: bool = False):
    """decompose.jl:400-436: CSA fragment in the orbital frame of the largest DF fragment; its Cartan coefficients collect
    every DF fragment in that frame.  Returns (lambda (cL, lower-triangular row-major), U1 (n x n), Lmat (n x n))."""
    frags = DF_decomposition(tbt, tol, fix_sign=fix_sign)
    n = tbt.shape[0]
    c1, w1, U1, _ = frags[0]
    Lmat = c1 * np.outer(w1, w1)                                      # :406-413
    for cm, wm, Um, _ in frags[1:]:                                   # :418-430
        Vm = Um @ U1.conj().T                                         # Vm = Um * U1dag
        v = (np.abs(Vm) ** 2) @ wm                                    # sum_p lambda_p |Vm[i,p]|^2
        Lmat = Lmat + cm * np.outer(v, v)
    lam = np.array([(Lmat[i, j] + Lmat[j, i]) / 2 for i in range(n) for j in range(i + 1)])   # cartan_mat_to_2b, fermionic.jl:34-48
    return lam, U1, Lmat


# ----------------------------------------------------------------------------------------------
# greedy drivers (decompose.jl:1-246) with scipy BFGS standing in for Optim.jl BFGS
# ----------------------------------------------------------------------------------------------

def greedy_step(tbt, obt=None, flavour: str = "CSA", x0=None, rng=None, do_svd=None, gtol: float = 1e-8,
                maxiter: int = 1000, use_literal: bool = False):
    """decompose.jl:82-120 / 207-246.  x0: explicit start (the reference draws it from the
    unseeded global RNG); otherwise theta ~ 2*pi*rand, lambda = 0 (CSA) or tbt_svd_1st (CSA_SD default)."""
    from scipy.optimize import minimize
    n = tbt.shape[0]
    cL, uL = cartan_2b_num_params(n), real_orbital_rotation_num_params(n)
    off = 0 if flavour == "CSA" else n
    if x0 is None:
        x0 = np.zeros(off + cL + uL)
        if do_svd is None:
            do_svd = flavour != "CSA"            # config.jl:9-10
        if do_svd:
            lam, th = tbt_svd_1st(tbt)
            x0[off:off + cL] = lam
            x0[off + cL:] = th
        else:
            rng = np.random.default_rng() if rng is None else rng
            x0[off + cL:] = 2 * np.pi * rng.random(uL)

    if use_literal:
        fun = lambda x: (cost_literal(x, tbt, obt, flavour), gradient_literal(x, tbt, obt, flavour))
    else:
        fun = lambda x: cost_grad_factored(x, tbt, obt, flavour)
    res = minimize(fun, x0, jac=True, method="BFGS", options={"gtol": gtol, "maxiter": maxiter})
    return res.x, float(res.fun), res


def greedy_decomposition(tbt, alpha_max: int, obt=None, flavour: str = "CSA", decomp_tol: float = 1e-5,
                         x0s=None, rng=None, **kw):
    """decompose.jl:1-80 / 122-205 (without the HDF5 checkpoint, which stays in Julia).
    Returns (list of x, list of costs, remaining tbt, remaining obt)."""
    n = tbt.shape[0]
    t_rem = np.array(tbt, dtype=float, copy=True)
    o_rem = None if obt is None else np.array(obt, dtype=float, copy=True)
    xs, costs = [], []
    curr = L2_partial_cost(t_rem, None, o_rem, None)
    a = 0
    while curr > decomp_tol and a < alpha_max:
        x0 = None if x0s is None else x0s[a]
        x, c, _ = greedy_step(t_rem, o_rem, flavour, x0=x0, rng=rng, **kw)
        h, W = reconstruct_factored(x, n, flavour)
        t_rem = t_rem - W
        if o_rem is not None:
            o_rem = o_rem - h
        xs.append(x)
        costs.append(c)
        curr = c
        a += 1
    return xs, costs, t_rem, o_rem


# ----------------------------------------------------------------------------------------------
# synthetic inputs -- deterministic
# ----------------------------------------------------------------------------------------------

def dense_sym_tbt(n: int, seed: int) -> np.ndarray:
    """N(0,1) tensor symmetrised over p<->q, r<->s, (pq)<->(rs), scaled 1/N^2."""
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((n, n, n, n))
    g = g + g.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    g *= 1.0 / (n * n)
    return g


def pairsym_tbt(n: int, seed: int) -> np.ndarray:
    """Only (pq)<->(rs) symmetric (exercises the un-packed N^2 x N^2 path)."""
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((n, n, n, n))
    return (g + g.transpose(2, 3, 0, 1)) / (2.0 * n * n)


def general_tbt(n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, n, n, n)) / (n * n)


def random_x(n: int, seed: int, flavour: str = "CSA") -> np.ndarray:
    """x = [lambda ~ N(0,1)/N ; theta ~ U[0, 2 pi)] (theta matches decompose.jl:93)."""
    rng = np.random.default_rng(1000 + seed)
    cL, uL = cartan_2b_num_params(n), real_orbital_rotation_num_params(n)
    parts = []
    if flavour != "CSA":
        parts.append(rng.standard_normal(n) / n)
    parts.append(rng.standard_normal(cL) / n)
    parts.append(2 * np.pi * rng.random(uL))
    return np.concatenate(parts)


def planted_tbt(n: int, R: int, seed: int, sigma: float = 0.0):
    """T = sum_{k<R} W(x_k), lambda ~ N(0,1), theta ~ U[0,2pi).  Returns (tbt, [x_k])."""
    rng = np.random.default_rng(2000 + seed)
    cL, uL = cartan_2b_num_params(n), real_orbital_rotation_num_params(n)
    t = np.zeros((n, n, n, n))
    xs = []
    for _ in range(R):
        x = np.concatenate([rng.standard_normal(cL), 2 * np.pi * rng.random(uL)])
        t += reconstruct_factored(x, n, "CSA")[1]
        xs.append(x)
    if sigma > 0:
        g = rng.standard_normal((n, n, n, n))
        g = g + g.transpose(1, 0, 2, 3)
        g = g + g.transpose(0, 1, 3, 2)
        g = g + g.transpose(2, 3, 0, 1)
        t += sigma * g / 8.0
    return t, xs
