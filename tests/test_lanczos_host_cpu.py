"""Host half of the Lanczos driver (mambo_debug_tridiag_leading: bisection + inverse iteration on the Ritz tridiagonal),
pinned against LAPACK on the CPU; the device half is covered by the -m gpu parity tests."""
import ctypes as C

import numpy as np
import pytest


@pytest.mark.parametrize("m", [1, 2, 3, 8, 33, 150, 400])
def test_tridiagonal_leading_pair_matches_lapack(mb, m):
    lib = mb.load()
    rng = np.random.default_rng(m)
    for trial in range(4):
        a = rng.standard_normal(m)
        b = np.abs(rng.standard_normal(m)) + 1e-3
        if trial == 3:                      # Toeplitz: clustered extremal eigenvalues
            a[:] = 1.0
            b[:] = 0.5
        T = np.diag(a) + np.diag(b[:m - 1], 1) + np.diag(b[:m - 1], -1)
        w, v = np.linalg.eigh(T)
        k = int(np.argmax(np.abs(w)))
        th = C.c_double()
        z = np.empty(m)
        assert lib.mambo_debug_tridiag_leading(a.ctypes.data, b.ctypes.data, m, C.byref(th), z.ctypes.data) == 0
        assert abs(th.value - w[k]) <= 1e-14 * max(1.0, abs(w[k])) * 8
        gap = np.min(np.abs(np.delete(w, k) - w[k])) if m > 1 else 1.0
        err = min(np.abs(z - v[:, k]).max(), np.abs(z + v[:, k]).max())
        assert err <= 1e-13 * max(1.0, abs(w[k]) / gap)
        assert np.linalg.norm(T @ z - th.value * z) <= 1e-13 * max(1.0, abs(w[k]))


def _glued(block_a, block_b, glue, copies):
    """`copies` identical unreduced blocks coupled by `glue`: every eigenvalue appears `copies` times to ~glue^2."""
    a = np.tile(block_a, copies)
    b = np.concatenate([np.concatenate([block_b, [glue]]) for _ in range(copies)])[:-1]
    return a, b


@pytest.mark.parametrize("nb,copies,glue", [(6, 2, 1e-13), (21, 2, 1e-9), (10, 3, 1e-12), (40, 4, 1e-14), (1, 2, 1e-15)])
def test_tridiagonal_cluster_vectors_span_the_invariant_subspace(mb, nb, copies, glue):
    """Host half of the double factorisation (mambo_debug_tridiag_cluster): eigenvalues that coincide to rounding inside one
    unreduced block -- what a degenerate eigenvalue of a molecular tensor becomes when the Lanczos block runs on past its
    convergence -- get orthonormal vectors with small residuals (a single twisted-factorisation step returns the same vector
    for all of them)."""
    lib = mb.load()
    rng = np.random.default_rng(nb * 10 + copies)
    a, b = _glued(rng.standard_normal(nb), np.abs(rng.standard_normal(nb - 1)) + 0.1, glue, copies)
    m = len(a)
    T = np.diag(a) + np.diag(b, 1) + np.diag(b, -1)
    w = np.linalg.eigvalsh(T)
    bb = np.concatenate([b, [0.0]])
    tn = np.max(np.abs(T).sum(axis=1))
    for c in range(0, m, copies):                                     # every cluster of `copies` eigenvalues
        th = np.ascontiguousarray(w[c:c + copies])
        assert th[-1] - th[0] < 1e-8
        Z = np.empty((copies, m))
        assert lib.mambo_debug_tridiag_cluster(a.ctypes.data, bb.ctypes.data, m, th.ctypes.data, copies, Z.ctypes.data) == 0
        assert np.max(np.abs(Z @ Z.T - np.eye(copies))) < 1e-13
        res = T @ Z.T - Z.T * th[None, :]
        assert np.max(np.linalg.norm(res, axis=0)) < 1e-13 * tn + 2 * (th[-1] - th[0])
