"""Initial guess of a greedy step from the largest "SVD" component of the residual two-body tensor.

Mirrors /root/reference/src/UTILS/guesses.jl: `tbt_svd_1st` (:51-84) and `SVD_to_CSA` (:3-49) with
`SOn_unitary_to_params` (unitaries.jl:130-147).  What changes relative to the reference: the full
`eigen(Symmetric(reshape(tbt,(N^2,N^2))))` of guesses.jl:55-63 -- O(N^6), of which only the eigenpair of largest
|eigenvalue| is kept -- is replaced by the library's device Lanczos on the residual that already lives in HBM
(`mambo_csa_leading_eig`); the O(N^3) rest (eigen of the N x N operator, log of the rotation) stays on the host, as it
stays in Julia for the ccall shim.
"""
from __future__ import annotations

import numpy as np

from .engine import CSAEngine
from .structures import cartan_2b_num_params

SVD_tiny = 1e-12   # config.jl:21


def SOn_unitary_to_params(U) -> np.ndarray:
    """unitaries.jl:130-147: upper-triangular entries (row-major) of real(log(U))."""
    from scipy.linalg import logm
    n = U.shape[0]
    Ulog = np.real(logm(np.asarray(U)))
    iu = np.triu_indices(n, 1)
    return np.ascontiguousarray(Ulog[iu])


def SVD_to_CSA(lam: float, omega, U):
    """guesses.jl:3-24 (do_givens = false): Cartan coefficients lam * omega_i * omega_j filled in the reference's own
    (upper-triangular, i <= j) order -- SURVEY 8a K1 -- and the rotation parameters of U.  Returns (coeffs, theta)."""
    n = U.shape[0]
    iu = np.triu_indices(n)
    coeffs = lam * np.asarray(omega)[iu[0]] * np.asarray(omega)[iu[1]]
    assert coeffs.size == cartan_2b_num_params(n)
    return coeffs, SOn_unitary_to_params(U)


def tbt_svd_1st(engine: CSAEngine, tiny: float = SVD_tiny, tol: float = 0.0, max_iter: int = 0):
    """guesses.jl:51-84 on the engine's device-resident residual.  Returns (lambda (cL), theta (uL))."""
    lead, full_l, _ = engine.leading_eig(tol=tol, max_iter=max_iter)
    cur_l = np.triu(full_l) + np.triu(full_l, 1).T                  # Symmetric(full_l): upper triangle
    if np.sum(np.abs(cur_l - full_l)) > tiny:
        if np.sum(np.abs(full_l + full_l.T)) > tiny:
            raise ValueError("SVD operator is neither Hermitian or anti-Hermitian, cannot do double factorization into "
                             "Hermitian fragment!")
        cur = 1j * full_l
        cur = np.triu(cur) + np.triu(cur, 1).conj().T                # Hermitian(1im * full_l)
        lead = -lead
        omega, Ul = np.linalg.eigh(cur)
    else:
        omega, Ul = np.linalg.eigh(cur_l)
    return SVD_to_CSA(lead, omega, Ul)


def DF_based_greedy(engine: CSAEngine, tol: float = 0.0, as_x: bool = True):
    """decompose.jl:400-436 on the engine's device-resident residual: the CSA fragment in the orbital frame of the largest
    double-factorisation fragment, its Cartan coefficients collecting every DF fragment in that frame (mambo_csa_df_decompose
    + mambo_csa_df_based_greedy; the reference runs DF_decomposition's full `eigen` first, decompose.jl:402).

    as_x: return (lambda (cL), theta (uL)), i.e. the fragment as a start vector x0 = [lambda; theta] of CSA_greedy_step -- the
    frame goes through SOn_unitary_to_params like SVD_to_CSA's (guesses.jl:20); a column of U1 is negated first when
    det U1 = -1 (an eigenvector's sign is free and the fragment does not depend on it, but log needs a proper rotation).
    Otherwise return (lambda, U1) with the frame as a matrix, like the reference's f_matrix_rotation."""
    engine.df_decompose(tol=tol)
    lam, U1, _ = engine.df_based_greedy()
    if not as_x:
        return lam, U1
    U1 = np.array(U1)
    if np.linalg.det(U1) < 0:
        U1[:, 0] = -U1[:, 0]
    return lam, SOn_unitary_to_params(U1)
