"""Host half of the initial guess (quantummambo.jl_b200/guesses.py: SVD_to_CSA, SOn_unitary_to_params -- guesses.jl:3-24,
unitaries.jl:130-147) against the oracle's restatement; the device half (leading eigenpair) is covered by the -m gpu tests."""
import numpy as np
import scipy.linalg as sla

import csa_oracle as o


def test_son_unitary_to_params_inverts_the_orbital_rotation(mb):
    for n in (2, 5, 9):
        theta = 0.3 * np.random.default_rng(n).standard_normal(n * (n - 1) // 2)      # inside the principal branch of log
        U = o.one_body_unitary(theta, n)
        got = mb.SOn_unitary_to_params(U)
        assert np.allclose(got, theta, atol=1e-12)
        assert np.allclose(got, o.SOn_unitary_to_params(U), atol=1e-14)


def test_svd_to_csa_fills_upper_triangular_order_like_the_reference(mb):
    """guesses.jl:12-18 loops `for i in 1:N, j in i:N` (K1 of SURVEY 8a): coefficient k belongs to the k-th pair of the
    upper triangle, row-major, scaled by the eigenvalue."""
    n = 4
    rng = np.random.default_rng(0)
    w = rng.standard_normal(n)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    if np.linalg.det(Q) < 0:
        Q[:, 0] = -Q[:, 0]
    lam = -1.7
    coeffs, theta = mb.SVD_to_CSA(lam, w, Q)
    want = [lam * w[i] * w[j] for i in range(n) for j in range(i, n)]
    assert np.allclose(coeffs, want, rtol=0, atol=1e-15)
    assert theta.shape == (n * (n - 1) // 2,)
    assert np.allclose(sla.expm(o.construct_anti_symmetric(n, theta)), Q, atol=1e-10)


def test_one_body_l1_and_correction_match_the_oracle(mb):
    """lcu.jl:262-267 / fermionic.jl:457-479 host mirrors (post-processing of the final 1-norms)."""
    n = 6
    T = o.dense_sym_tbt(n, 3)
    h = np.random.default_rng(1).standard_normal((n, n))
    h = h + h.T
    H = mb.F_OP(([0.0], h, T))
    assert np.allclose(mb.ob_correction(T), o.ob_correction(T), atol=1e-15)
    assert abs(mb.one_body_L1(H) - o.one_body_L1(h, T)) <= 1e-12 * o.one_body_L1(h, T)
