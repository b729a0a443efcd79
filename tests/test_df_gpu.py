"""GPU parity tests of the device double factorisation (mambo_csa_df_decompose / _df_get / _df_based_greedy, through the
C-ABI) against the oracle's restatement of DF_decomposition / DF_based_greedy (decompose.jl:297-436).

Eigenvectors are only defined up to a sign (and up to a rotation inside a degenerate eigenspace), so the comparison is on
what the reference's consumers see: the number of fragments, their coefficients in the reference's order, every fragment's
operator L (x) L (sign-free), the sum of all fragments, and DF_based_greedy's Cartan coefficients.  Tolerance 1e-10
relative (FP64), as for the cost and gradient."""
import numpy as np
import pytest

import csa_oracle as o

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


def _planted(n, R, seed):
    return o.planted_tbt(n, R, seed)[0]


def _check_against_oracle(mb, T, tol=1e-8, sym=None, per_fragment=True):
    n = T.shape[0]
    kw = {} if sym is None else {"sym": sym}
    with mb.CSAEngine(T, **kw) as e:
        st = e.df_decompose(tol=tol, verify=True)
        nf = st["num_frags"]
        coeff, omega, U, L, kind = e.df_get(0, nf)
        lam, U1, Lm = e.df_based_greedy()
    fr = o.DF_decomposition(T, tol)
    assert nf == len(fr)
    assert st["jacobi_sweeps_max"] > 0 and st["anti_frags"] == 0 and not kind.any()
    assert st["ortho_dev_after"] <= 1e-12 and st["max_rel_residual"] <= 1e-11
    c_ref = np.array([f[0] for f in fr])
    assert rel(coeff, c_ref) <= RTOL
    rec = np.zeros((n, n, n, n))
    for k in range(nf):
        Lk = (U[k] * omega[k]) @ U[k].T                                   # U diag(omega) U'
        assert np.max(np.abs(Lk - L[k])) <= 1e-12 * max(1.0, np.max(np.abs(L[k])))
        assert np.max(np.abs(U[k].T @ U[k] - np.eye(n))) <= 1e-12
        assert np.all(np.diff(omega[k]) >= 0)                             # ascending, like LAPACK
        assert abs(np.linalg.norm(L[k]) - 1.0) <= 1e-12                   # unit eigenvector of the N^2 x N^2 view
        rec += coeff[k] * np.einsum("ab,cd->abcd", Lk, Lk)
    assert rel(rec, o.DF_to_tbt(fr, n)) <= RTOL
    if per_fragment:                                                      # well-separated spectrum: fragment by fragment
        for k in range(nf):
            Lr = fr[k][3]
            assert rel(np.einsum("ab,cd->abcd", L[k], L[k]), np.einsum("ab,cd->abcd", Lr, Lr)) <= 1e-8
            assert rel(omega[k] if np.dot(L[k].ravel(), Lr.ravel()) > 0 else -omega[k][::-1], fr[k][1]) <= 1e-8
    lam0, _, Lm0 = o.DF_based_greedy(T, tol, fix_sign=True)
    assert rel(lam, lam0) <= 1e-9 and rel(Lm, Lm0) <= 1e-9
    assert np.max(np.abs(np.abs(U1.T @ U[0]) - np.eye(n))) <= 1e-10       # U1 is the frame of the leading fragment
    return st


@pytest.mark.parametrize("n", [2, 3, 6, 7, 10])
def test_df_decomposition_matches_oracle(mb, n):
    st = _check_against_oracle(mb, o.dense_sym_tbt(n, n))
    assert st["num_frags"] == n * (n + 1) // 2                            # the symmetric subspace; the rest is annihilated


def test_df_in_the_unpacked_mode(mb):
    st = _check_against_oracle(mb, o.dense_sym_tbt(6, 1), sym=mb.SYM_PAIR)
    assert st["restarts"] >= 1                                            # the anti-symmetric null space is only reached by a restart


def test_df_low_rank_tensor_stops_early(mb):
    n, R = 8, 2
    st = _check_against_oracle(mb, _planted(n, R, 3), per_fragment=False)
    assert st["num_frags"] <= R * n and st["krylov_dim"] <= R * n + 3


def test_df_degenerate_eigenvalues(mb):
    n = 6
    rng = np.random.default_rng(5)
    A = rng.standard_normal((n, n)); A = A + A.T
    B = rng.standard_normal((n, n)); B = B + B.T
    A /= np.linalg.norm(A)
    B -= A * np.sum(A * B)
    B /= np.linalg.norm(B)
    T = 0.7 * (np.einsum("ab,cd->abcd", A, A) + np.einsum("ab,cd->abcd", B, B))
    n_ = T.shape[0]
    with mb.CSAEngine(T) as e:
        st = e.df_decompose(verify=True)
        coeff, omega, U, L, kind = e.df_get(0, st["num_frags"])
    assert st["num_frags"] == 2 and st["restarts"] >= 1
    assert rel(coeff, [0.7, 0.7]) <= RTOL
    rec = sum(coeff[k] * np.einsum("ab,cd->abcd", L[k], L[k]) for k in range(2))
    assert rel(rec, T) <= RTOL


def test_df_truncation_and_max_frags(mb):
    T = o.dense_sym_tbt(7, 0)
    fr = o.DF_decomposition(T, 1e-8)
    cut = abs(fr[9][0]) * 0.999                                           # keeps exactly the ten leading fragments
    with mb.CSAEngine(T) as e:
        assert e.df_decompose(tol=cut)["num_frags"] == 10
        assert e.df_decompose(max_frags=4)["num_frags"] == 4
        coeff, *_ = e.df_get(0, 4)
        with pytest.raises(mb.MamboError):
            e.df_get(3, 2)
    assert rel(coeff, [f[0] for f in fr[:4]]) <= RTOL


def test_df_zero_tensor_and_errors(mb):
    with mb.CSAEngine(np.zeros((4, 4, 4, 4))) as e:
        with pytest.raises(mb.MamboError):
            e.df_get(0, 1)                                                # nothing decomposed yet
        assert e.df_decompose()["num_frags"] == 0
        with pytest.raises(mb.MamboError):
            e.df_based_greedy()


def test_df_refuses_a_nonsymmetric_matrix_view(mb):
    """decompose.jl:319-322: the reference leaves the symmetric eigenproblem for such input; the library says so"""
    with mb.CSAEngine(o.general_tbt(4, 1)) as e:
        with pytest.raises(mb.MamboError):
            e.df_decompose()
    with mb.CSAEngine(o.dense_sym_tbt(4, 1), sym=mb.SYM_GENERAL) as e:      # forced mode, symmetric tensor: fine
        assert e.df_decompose()["num_frags"] == 10


def test_df_follows_the_greedy_residual(mb):
    """the decomposition is of the RESIDENT residual: after subtract_fragment it sees T - W(x)"""
    n = 6
    T = o.dense_sym_tbt(n, 2)
    x = o.random_x(n, 4)
    W = o.to_OP(x, n)[1]                                                  # (obt, tbt)
    with mb.CSAEngine(T) as e:
        e.subtract_fragment(x)
        st = e.df_decompose()
        coeff, *_ = e.df_get(0, st["num_frags"])
    c_ref = np.array([f[0] for f in o.DF_decomposition(T - W)])
    assert rel(coeff, c_ref) <= RTOL


def test_df_medium_size_properties(mb):
    """N = 24 (300 x 300 packed): against numpy's dense eigensolver on the matrix view, and the completeness of the sum"""
    n = 24
    T = o.dense_sym_tbt(n, 4)
    with mb.CSAEngine(T) as e:
        st = e.df_decompose(verify=True)
        nf = st["num_frags"]
        coeff, omega, U, L, kind = e.df_get(0, nf)
    assert nf == n * (n + 1) // 2 and st["max_rel_residual"] <= 1e-11 and st["ortho_dev_after"] <= 1e-12
    ev = np.linalg.eigvalsh(T.reshape(n * n, n * n, order="F"))
    ev = ev[np.abs(ev) >= 1e-8]
    ev = ev[np.argsort(-np.abs(ev), kind="stable")]
    assert rel(coeff, ev) <= RTOL
    rec = np.einsum("k,kab,kcd->abcd", coeff, L, L)
    assert rel(rec, T) <= RTOL


def test_dgemm_device_matches_numpy(mb):
    import torch
    lib = mb.load()
    rng = np.random.default_rng(0)
    for ta, M, N, K in [(1, 70, 50, 33), (0, 70, 50, 33), (1, 300, 257, 129), (0, 129, 300, 257)]:
        lda = ((M if ta else K) + 7) // 2 * 2
        ldb, ldc = (N + 5) // 2 * 2, (N + 3) // 2 * 2
        A = rng.standard_normal((K, lda) if ta else (M, lda))
        B = rng.standard_normal((K, ldb))
        dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
        dC = torch.zeros(M, ldc, dtype=torch.float64, device="cuda")
        assert lib.mambo_dgemm_device(0, None, ta, M, N, K, dA.data_ptr(), lda, dB.data_ptr(), ldb, dC.data_ptr(), ldc) == 0
        torch.cuda.synchronize()
        ref = (A[:, :M].T if ta else A[:, :K]) @ B[:, :N]
        assert rel(dC.cpu().numpy()[:, :N], ref) <= 1e-13
        assert float(dC[:, N:].abs().max()) == 0.0 if ldc > N else True    # nothing written past column N
    assert lib.mambo_dgemm_device(0, None, 1, 4, 4, 4, dA.data_ptr(), 3, dB.data_ptr(), 4, dC.data_ptr(), 4) != 0   # odd lda
