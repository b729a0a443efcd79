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


@pytest.mark.parametrize("n", [3, 4, 5])
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


@pytest.mark.parametrize("n", [3, 5, 7])
def test_c_literal_restatement_matches_numpy_oracle(n):
    import csa_literal as cl
    T = o.pairsym_tbt(n, n)
    x = o.random_x(n, 1)
    c, g = cl.cost_grad_literal_c(x, T)
    c0, g0 = o.cost_grad_factored(x, T)
    assert abs(c - c0) <= 1e-12 * abs(c0)
    assert np.max(np.abs(g - g0)) <= 1e-11 * np.max(np.abs(g0))
