"""CPU tests that pin the oracle (oracle/csa_oracle.py): literal transcription == factored form, finite
differences, analytic known answers, committed golden vectors.  The reference has no tests or fixtures for
this path, so these are the substitute pins."""
import math

import numpy as np
import pytest

import csa_oracle as o

KINDS = {"dense_sym": o.dense_sym_tbt, "pairsym": o.pairsym_tbt, "general": o.general_tbt}


def _obt(n, seed):
    return np.random.default_rng(100 + seed).standard_normal((n, n)) / n


@pytest.mark.parametrize("n", [3, 4, 5, 6, 7, 8, 9, 10, 11, 12])
@pytest.mark.parametrize("flavour", ["CSA", "CSA_SD"])
@pytest.mark.parametrize("kind", ["dense_sym", "pairsym", "general"])
def test_literal_equals_factored(n, flavour, kind):
    T = KINDS[kind](n, n)
    obt = _obt(n, n) if flavour == "CSA_SD" else None
    x = o.random_x(n, 7, flavour)
    cl, gl = o.cost_literal(x, T, obt, flavour), o.gradient_literal(x, T, obt, flavour)
    cf, gf = o.cost_grad_factored(x, T, obt, flavour)
    assert abs(cl - cf) <= 1e-12 * abs(cl)
    assert np.max(np.abs(gl - gf)) <= 1e-12 * np.max(np.abs(gl))


@pytest.mark.parametrize("flavour", ["CSA", "CSA_SD"])
def test_gradient_matches_finite_differences_on_symmetric_target(flavour):
    # Q2 of SURVEY.md: the reference's lambda gradient is the true derivative only for (pq)<->(rs)-symmetric D
    n = 4
    T = o.dense_sym_tbt(n, 3)
    obt = _obt(n, 3) if flavour == "CSA_SD" else None
    obt = None if obt is None else (obt + obt.T) / 2
    x = o.random_x(n, 2, flavour)
    g = o.gradient_literal(x, T, obt, flavour)
    h = 1e-6
    fd = np.zeros_like(x)
    for i in range(len(x)):
        xp, xm = x.copy(), x.copy()
        xp[i] += h
        xm[i] -= h
        fd[i] = (o.cost_literal(xp, T, obt, flavour) - o.cost_literal(xm, T, obt, flavour)) / (2 * h)
    assert np.max(np.abs(fd - g)) <= 1e-7 * np.max(np.abs(g))


def test_one_sided_lambda_gradient_differs_from_fd_for_general_tensor():
    # the oracle must reproduce the reference, not calculus, for non-symmetric tensors (gradient.jl:22-26)
    n = 3
    T = o.general_tbt(n, 1)
    x = o.random_x(n, 5)
    g = o.gradient_literal(x, T)
    h = 1e-6
    i = 1  # an off-diagonal lambda
    xp, xm = x.copy(), x.copy()
    xp[i] += h
    xm[i] -= h
    fd = (o.cost_literal(xp, T) - o.cost_literal(xm, T)) / (2 * h)
    assert abs(fd - g[i]) > 1e-4 * abs(g[i])


def test_planted_tensor_has_zero_cost_and_gradient():
    n = 5
    T, xs = o.planted_tbt(n, 1, 0)
    c, g = o.cost_grad_factored(xs[0], T)
    assert c < 1e-24
    assert np.max(np.abs(g)) < 1e-11


def test_theta_zero_gives_identity_rotation():
    n = 4
    lam = np.random.default_rng(0).standard_normal(o.cartan_2b_num_params(n))
    x = np.concatenate([lam, np.zeros(o.real_orbital_rotation_num_params(n))])
    _, W = o.to_OP(x, n, "CSA")
    L = o.get_cartan_matrix(n, lam)
    for i in range(n):
        for j in range(n):
            assert W[i, i, j, j] == pytest.approx(L[i, j], abs=1e-15)
    mask = np.ones_like(W, dtype=bool)
    for i in range(n):
        for j in range(n):
            mask[i, i, j, j] = False
    assert np.max(np.abs(W[mask])) < 1e-15


def test_parameter_orderings():
    # lambda: lower-triangular row-major; theta: upper-triangular row-major (unitaries.jl:418-424)
    n = 6
    idx = 1
    for i in range(n):
        for j in range(i + 1, n):
            assert o.get_rot_indices(n, idx) == (i, j)
            idx += 1
    K = o.construct_anti_symmetric(n, np.arange(1, 16, dtype=float))
    assert K[0, 1] == 1 and K[0, 5] == 5 and K[1, 2] == 6 and K[2, 1] == -6
    L = o.get_cartan_matrix(3, np.arange(1, 7, dtype=float))
    assert L[0, 0] == 1 and L[1, 0] == 2 and L[1, 1] == 3 and L[2, 0] == 4 and L[0, 2] == 4


def test_golden_vectors(golden):
    for name, c in golden.items():
        n, sd = int(c["meta"][0]), int(c["meta"][1])
        fl = "CSA_SD" if sd else "CSA"
        obt = c.get("obt")
        cf, gf = o.cost_grad_factored(c["x"], c["tbt"], obt, fl)
        assert abs(cf - float(c["cost"])) <= 1e-12 * abs(float(c["cost"])), name
        assert np.max(np.abs(gf - c["grad"])) <= 1e-12 * np.max(np.abs(c["grad"])), name
        h, W = o.reconstruct_factored(c["x"], n, fl)
        assert np.max(np.abs(W - c["W"])) <= 1e-13 * max(np.max(np.abs(c["W"])), 1.0), name
        if sd:
            assert np.max(np.abs(h - c["h"])) <= 1e-13, name


def test_two_phase_partials_sum_to_the_unsharded_result():
    n = 5
    for fl in ("CSA", "CSA_SD"):
        T = o.pairsym_tbt(n, 1)
        obt = _obt(n, 0) if fl != "CSA" else None
        x = o.random_x(n, 2, fl)
        c0, g0 = o.cost_grad_factored(x, T, obt, fl)
        p = o.partial_factored(x, T, 0, 7) + o.partial_factored(x, T, 7, 19) + o.partial_factored(x, T, 19, 25)
        c1, g1 = o.finish_factored(p, x, n, fl, obt)
        assert abs(c1 - c0) <= 1e-13 * abs(c0)
        assert np.max(np.abs(g1 - g0)) <= 1e-12 * np.max(np.abs(g0))


def test_csa_l1_and_one_body_l1():
    n = 3
    lam = np.array([1.0, -2.0, 3.0, 0.5, -0.25, 4.0])
    # i>=j order: (0,0)=1 (1,0)=-2 (1,1)=3 (2,0)=.5 (2,1)=-.25 (2,2)=4   lcu.jl:162-178
    assert o.CSA_L1(lam, n) == pytest.approx(0.5 * (1 + 3 + 4) + 2 * (2 + 0.5 + 0.25))
    assert o.CSA_L1(lam, n, spin_orb=True) == pytest.approx(0.5 * (2 + 0.5 + 0.25))
    T = o.dense_sym_tbt(n, 0)
    obt = np.diag([1.0, -2.0, 0.5])
    M = obt + 2 * sum(T[:, :, r, r] for r in range(n))
    assert o.one_body_L1(obt, T) == pytest.approx(np.sum(np.abs(np.linalg.eigvalsh(M))))


def test_greedy_on_planted_tensor_reduces_cost():
    n = 4
    T, xs = o.planted_tbt(n, 1, 3)
    c0 = o.L2_partial_cost(T)
    rng = np.random.default_rng(0)
    x0 = xs[0] + 0.05 * rng.standard_normal(xs[0].shape)      # start near the planted fragment
    out_xs, costs, t_rem, _ = o.greedy_decomposition(T, 1, x0s=[x0], decomp_tol=1e-10)
    assert costs[0] < 1e-10 * c0
    assert o.L2_partial_cost(t_rem) == pytest.approx(costs[0], abs=1e-12)


def test_svd_initial_guess_reproduces_leading_fragment():
    # guesses.jl:51-84: for a rank-one tensor A (x) A with symmetric A the SVD guess is exact -- whenever the
    # eigenvector matrix is a proper rotation.  (K1 of SURVEY.md: the reference takes real(log(U)) even when
    # det U = -1, in which case exp(log) does not give U back; reproduced, so those draws are skipped.)
    n = 4
    checked = 0
    for seed in range(8):
        rng = np.random.default_rng(seed)
        A = rng.standard_normal((n, n))
        A = A + A.T
        T = np.einsum("pq,rs->pqrs", A, A)
        _, Ul = np.linalg.eigh(A / np.linalg.norm(A))
        lam, th = o.tbt_svd_1st(T)
        U = o.one_body_unitary(th, n)
        w = np.zeros((n, n))     # K1: coefficients come in upper-triangular order
        idx = 0
        for i in range(n):
            for j in range(i, n):
                w[i, j] = w[j, i] = lam[idx]
                idx += 1
        W = np.einsum("al,bl,cm,dm,lm->abcd", U, U, U, U, w)
        err = np.max(np.abs(W - T)) / np.max(np.abs(T))
        if np.max(np.abs(np.abs(U) - np.abs(Ul))) < 1e-8:    # log/exp round trip reproduced the eigenvectors
            assert err < 1e-9
            checked += 1
    assert checked >= 1


@pytest.mark.parametrize("n", [3, 5, 7, 9])
@pytest.mark.parametrize("flavour", ["CSA", "CSA_SD"])
@pytest.mark.parametrize("kind", ["dense_sym", "pairsym", "general"])
def test_c_literal_restatement_matches_numpy_oracle(n, flavour, kind):
    """oracle/csa_literal.c: the reference's loop nests as plain C loops (N^6 cartan rotation, N^8 general rotation for
    CSA_SD, dense N^6 del_w_u, uL x N^6 theta contraction) against the factored numpy form -- all tensor kinds, so the
    one-sided lambda gradient (Q2) and the full product rule (Q3) are pinned on non-symmetric inputs too."""
    import csa_literal as cl
    T = KINDS[kind](n, n)
    obt = _obt(n, n) if flavour == "CSA_SD" else None
    x = o.random_x(n, 1, flavour)
    c, g = cl.cost_grad_literal_c(x, T, obt, flavour)
    c0, g0 = o.cost_grad_factored(x, T, obt, flavour)
    assert abs(c - c0) <= 1e-12 * abs(c0)
    assert np.max(np.abs(g - g0)) <= 1e-11 * np.max(np.abs(g0))


# ---- extended-precision ground truth (oracle/csa_oracle_ld.py): independent arithmetic, no scipy / LAPACK / BLAS --------

def _ld_tol(c_ld, tt):
    # cost: 1e-12 relative, down to what FP64 resolves (every element of D = W - T carries eps |T|)
    return 1e-12 * abs(c_ld) + 1e-14 * np.sqrt(abs(c_ld) * tt) + 1e-26 * tt


@pytest.mark.parametrize("n", [3, 4, 5])
@pytest.mark.parametrize("flavour", ["CSA", "CSA_SD"])
@pytest.mark.parametrize("kind", ["dense_sym", "general"])
def test_mpmath_literal_reference_formulas_equal_long_double_factored_form(n, flavour, kind):
    """The reference's formulas loop for loop at 40 digits (mpmath expm, mpmath eig for del_u_theta with the 1e-8 switch,
    gradient.jl:113-126) against the factored form in long double: two restatements that share neither arithmetic nor
    libraries agree to 1e-15, so the factorisation (and the adjoint-Frechet theta gradient) is the reference's function."""
    import csa_oracle_ld as ld
    T = KINDS[kind](n, 10 + n)
    obt = _obt(n, 3) if flavour == "CSA_SD" else None
    x = o.random_x(n, 11, flavour)
    cm, gm = ld.cost_grad_literal_mp(x, T, obt, flavour)
    cl_, gl_ = ld.cost_grad_factored_ld(x, T, obt, flavour)
    assert abs(float(cl_) - cm) <= 1e-15 * abs(cm)
    assert np.max(np.abs(gl_.astype(float) - gm)) <= 1e-15 * np.max(np.abs(gm))


@pytest.mark.parametrize("n", [3, 6, 9, 12, 16])
@pytest.mark.parametrize("flavour", ["CSA", "CSA_SD"])
@pytest.mark.parametrize("kind", ["dense_sym", "pairsym", "general"])
def test_fp64_oracle_within_1e12_of_extended_precision(n, flavour, kind):
    import csa_oracle_ld as ld
    T = KINDS[kind](n, n + 1)
    obt = _obt(n, n) if flavour == "CSA_SD" else None
    x = o.random_x(n, 5, flavour)
    c, g = o.cost_grad_factored(x, T, obt, flavour)
    cl_, gl_ = ld.cost_grad_factored_ld(x, T, obt, flavour)
    assert abs(c - float(cl_)) <= 1e-12 * abs(float(cl_))
    assert np.max(np.abs(g - gl_.astype(float))) <= 1e-12 * np.max(np.abs(gl_.astype(float)))
    if n <= 9:
        cq, gq = o.cost_literal(x, T, obt, flavour), o.gradient_literal(x, T, obt, flavour)
        assert abs(cq - float(cl_)) <= 1e-12 * abs(float(cl_))
        assert np.max(np.abs(gq - gl_.astype(float))) <= 1e-12 * np.max(np.abs(gl_.astype(float)))


@pytest.mark.parametrize("n", [5, 9, 13])
def test_fp64_oracle_in_the_cancellation_regime_against_extended_precision(n):
    """x at / near a planted fragment: the residual is tiny against the tensor, so cost and gradient are differences of
    O(|T|) quantities.  The FP64 oracle must stay within FP64's floor of the long-double value (the same tolerance model
    the CUDA engine is held to in tests/test_gpu_parity.py)."""
    import csa_oracle_ld as ld
    T, xs = o.planted_tbt(n, 1, n)
    tt = float(np.sum(T * T))
    rng = np.random.default_rng(0)
    for eps in (0.0, 1e-7, 1e-3):
        x = xs[0] + eps * rng.standard_normal(xs[0].shape)
        c, g = o.cost_grad_factored(x, T)
        cl_, gl_ = ld.cost_grad_factored_ld(x, T)
        assert abs(c - float(cl_)) <= _ld_tol(float(cl_), tt)
        assert np.max(np.abs(g - gl_.astype(float))) <= 1e-12 * np.max(np.abs(gl_.astype(float))) + 1e-13 * np.sqrt(tt)


def test_golden_vectors_carry_the_extended_precision_values(golden):
    """tests/golden/csa_golden.npz stores, next to the literal FP64 values, cost / gradient from the long-double factored
    form (rounded to FP64): the CUDA engine is compared with both (tests/test_gpu_parity.py::test_golden_vectors)."""
    import csa_oracle_ld as ld
    for name, c in golden.items():
        n, sd = int(c["meta"][0]), int(c["meta"][1])
        fl = "CSA_SD" if sd else "CSA"
        cl_, gl_ = ld.cost_grad_factored_ld(c["x"], c["tbt"], c.get("obt"), fl)
        assert float(cl_) == float(c["cost_ld"]), name
        assert np.array_equal(gl_.astype(float), c["grad_ld"]), name
        assert abs(float(c["cost"]) - float(c["cost_ld"])) <= 1e-12 * abs(float(c["cost_ld"])), name
        assert np.max(np.abs(c["grad"] - c["grad_ld"])) <= 1e-12 * np.max(np.abs(c["grad_ld"])), name


@pytest.mark.parametrize("n", [2, 5, 12, 23])
def test_packed_single_gemm_restatement_equals_factored(n):
    """oracle.cost_grad_packed restates the ALGORITHM of the CUDA engine (packed rows, lean single-GEMM formulation with
    the N x N closed forms) on the CPU; bench.py times it as `cpu_baseline_packed`.  Same numbers as the factored form."""
    T = o.dense_sym_tbt(n, 3)
    x = o.random_x(n, 4)
    c0, g0 = o.cost_grad_factored(x, T)
    c, g = o.cost_grad_packed(x, o.pack_tbt(T), n)
    assert abs(c - c0) <= 1e-12 * abs(c0)
    assert np.max(np.abs(g - g0)) <= 1e-12 * np.max(np.abs(g0))


# ---- double factorisation (decompose.jl:297-436) ---------------------------------------------------------------------------
@pytest.mark.parametrize("n", [3, 5, 8])
def test_df_oracle_reconstructs_the_tensor_and_orders_fragments(n):
    T = o.dense_sym_tbt(n, n)
    fr = o.DF_decomposition(T)
    assert len(fr) == n * (n + 1) // 2                       # rank of an 8-fold-symmetric tensor's matrix view
    c = np.array([f[0] for f in fr])
    assert np.all(np.diff(np.abs(c)) <= 0)                   # decreasing |eigenvalue| (decompose.jl:332)
    assert np.max(np.abs(o.DF_to_tbt(fr, n) - T)) <= 1e-13 * np.max(np.abs(T)) * n * n
    for coeff, w, U, L in fr:
        assert np.max(np.abs((U * w) @ U.T - L)) <= 1e-13 and abs(np.linalg.norm(L) - 1) <= 1e-13


def test_df_oracle_truncates_at_tol_and_takes_the_antihermitian_branch():
    n = 4
    T = o.dense_sym_tbt(n, 0)
    fr = o.DF_decomposition(T, tol=abs(o.DF_decomposition(T)[3][0]) * 0.999)
    assert len(fr) == 4
    # pair-symmetric tensor with an anti-symmetric eigenvector: coefficient negated, complex frame (decompose.jl:355-361)
    rng = np.random.default_rng(1)
    A = rng.standard_normal((n, n)); A = A - A.T; A /= np.linalg.norm(A)
    Tanti = 0.3 * np.einsum("ab,cd->abcd", A, A)
    fr = o.DF_decomposition(Tanti)
    assert len(fr) == 1 and abs(fr[0][0] + 0.3) <= 1e-14 and np.iscomplexobj(fr[0][2])
    assert np.max(np.abs(o.DF_to_tbt(fr, n) - Tanti)) <= 1e-14


def test_df_based_greedy_oracle_known_answer():
    """a single planted DF fragment c * L (x) L: the greedy CSA fragment in its own frame is lambda_ij = c w_i w_j"""
    n = 5
    rng = np.random.default_rng(3)
    A = rng.standard_normal((n, n)); A = A + A.T; A /= np.linalg.norm(A)
    T = 1.7 * np.einsum("ab,cd->abcd", A, A)
    lam, U1, Lm = o.DF_based_greedy(T)
    w = np.linalg.eigvalsh(A)
    ref = 1.7 * np.outer(w, w)
    assert np.max(np.abs(Lm - ref)) <= 1e-13 or np.max(np.abs(Lm - ref[::-1, ::-1])) <= 1e-13
    # and the CSA fragment it defines reproduces the tensor: W = sum_ij Lm_ij U1[:,i]U1[:,i]' (x) U1[:,j]U1[:,j]'
    W = np.einsum("ij,ai,bi,cj,dj->abcd", Lm, U1, U1, U1, U1)
    assert np.max(np.abs(W - T)) <= 1e-13


def test_ne2_known_answer():
    """SURVEY 8c pin (3): the Ne^2 symmetry operator (symmetries.jl:8-13: tbt[i,i,j,j] = 1) is a CSA fragment at theta = 0.
    Through cartan_2b_to_tbt (fermionic.jl:20-32) its coefficients are lambda = 1 everywhere: cost 0, gradient 0.  The vector
    of Ne2_lambda_builder (symmetries.jl:28-45: 2 off the diagonal) is NOT that fragment under cartan_2b_to_tbt -- the
    reconstruction carries 2 at [i,i,j,j], i != j -- so its cost against Ne^2 is N (N - 1); both facts are pinned here."""
    for n in (3, 6, 9):
        ne2 = np.zeros((n, n, n, n))
        for i in range(n):
            for j in range(n):
                ne2[i, i, j, j] = 1.0
        cL, uL = o.cartan_2b_num_params(n), o.real_orbital_rotation_num_params(n)
        x = np.concatenate([np.ones(cL), np.zeros(uL)])
        c, g = o.cost_grad_factored(x, ne2)
        assert c < 1e-24 and np.max(np.abs(g)) < 1e-12
        assert o.cost_literal(x, ne2) < 1e-24
        lam_b = np.array([1.0 if i == j else 2.0 for i in range(n) for j in range(i + 1)])       # Ne2_lambda_builder
        xb = np.concatenate([lam_b, np.zeros(uL)])
        cb, _ = o.cost_grad_factored(xb, ne2)
        assert abs(cb - n * (n - 1)) < 1e-10
        assert abs(o.cost_literal(xb, ne2) - n * (n - 1)) < 1e-10
        # Ne^2 is invariant under every orbital rotation: the cost does not depend on theta, the theta gradient vanishes
        xr = np.concatenate([np.ones(cL), 2 * np.pi * np.random.default_rng(n).random(uL)])
        cr, gr = o.cost_grad_factored(xr, ne2)
        assert cr < 1e-20 and np.max(np.abs(gr[cL:])) < 1e-9


@pytest.mark.parametrize("n", [4, 6])
@pytest.mark.parametrize("flavour", ["CSA", "CSA_SD"])
@pytest.mark.parametrize("kind", ["dense_sym", "general"])
def test_gradient_against_reverse_mode_autodiff(n, flavour, kind):
    """An independent derivative: the cost written as plain tensor algebra in torch (float64, matrix_exp) and differentiated by
    reverse-mode autodiff, against the oracle's analytic gradient (gradient.jl:1-290).  On pair-symmetric targets every component
    must agree; on general tensors the theta (and lambda_1) parts are exact (Q3) while the reference's one-sided lambda expression
    (Q2) is not the derivative -- both facts are asserted."""
    import torch
    T = KINDS[kind](n, 2)
    obt = _obt(n, 2) if flavour == "CSA_SD" else None
    x = o.random_x(n, 4, flavour)
    c0, g0 = o.cost_grad_factored(x, T, obt, flavour)
    cL, uL = o.cartan_2b_num_params(n), o.real_orbital_rotation_num_params(n)
    off = n if flavour == "CSA_SD" else 0
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    lam, th = xt[off:off + cL], xt[off + cL:]
    iu = torch.triu_indices(n, n, 1)
    K = torch.zeros((n, n), dtype=torch.float64)
    K = K.index_put((iu[0], iu[1]), th)                     # upper triangle row-major (unitaries.jl:406-415)
    K = K - K.T
    U = torch.matrix_exp(K)
    il = torch.tril_indices(n, n, 0)
    L = torch.zeros((n, n), dtype=torch.float64).index_put((il[0], il[1]), lam)   # lower triangle row-major (fermionic.jl:20-32)
    L = L + L.T - torch.diag(torch.diagonal(L))
    W = torch.einsum("lm,pl,ql,rm,sm->pqrs", L, U, U, U, U)
    cost = torch.sum((W - torch.tensor(T)) ** 2)
    if flavour == "CSA_SD":
        h = (U * xt[:n]) @ U.T
        cost = cost + torch.sum((h - torch.tensor(obt)) ** 2)
    cost.backward()
    ga = xt.grad.numpy()
    assert abs(float(cost.detach()) - c0) <= 1e-12 * abs(c0)
    scale = np.max(np.abs(ga))
    assert np.max(np.abs(ga[off + cL:] - g0[off + cL:])) <= 1e-10 * scale          # theta: exact for any tensor
    if flavour == "CSA_SD":
        assert np.max(np.abs(ga[:n] - g0[:n])) <= 1e-10 * scale                    # lambda_1
    dl = np.max(np.abs(ga[off:off + cL] - g0[off:off + cL]))
    if kind == "dense_sym":
        assert dl <= 1e-10 * scale
    else:
        assert dl > 1e-3 * scale                                                   # the reference's one-sided expression (Q2)
