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
