"""Development check of the device double factorisation against the oracle (run on a B200)."""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mambo_b200 as mb
import csa_oracle as oracle
import ctypes as C
import torch

def gemm_check():
    lib = mb.load()
    rng = np.random.default_rng(0)
    for (ta, M, N, K) in [(1, 70, 50, 33), (0, 70, 50, 33), (1, 300, 257, 129), (0, 129, 300, 257), (1, 1000, 1000, 1000)]:
        lda = ((M if ta else K) + 7) // 2 * 2; ldb = (N + 5) // 2 * 2; ldc = (N + 3) // 2 * 2
        A = rng.standard_normal((K, lda) if ta else (M, lda)); B = rng.standard_normal((K, ldb))
        dA = torch.from_numpy(A).cuda(); dB = torch.from_numpy(B).cuda(); dC = torch.zeros(M, ldc, dtype=torch.float64, device="cuda")
        rc = lib.mambo_dgemm_device(0, None, ta, M, N, K, dA.data_ptr(), lda, dB.data_ptr(), ldb, dC.data_ptr(), ldc)
        assert rc == 0, lib.mambo_last_error()
        torch.cuda.synchronize()
        ref = (A[:, :M].T if ta else A[:, :K]) @ B[:, :N]
        err = np.abs(dC.cpu().numpy()[:, :N] - ref).max() / np.abs(ref).max()
        print(f"gemm ta={ta} {M}x{N}x{K}: rel err {err:.2e}")
        assert err < 1e-13
    # speed
    M = N = K = 5056
    dA = torch.randn(K, M, dtype=torch.float64, device="cuda"); dB = torch.randn(K, N, dtype=torch.float64, device="cuda"); dC = torch.empty(M, N, dtype=torch.float64, device="cuda")
    for ta in (1, 0):
        for _ in range(2):
            lib.mambo_dgemm_device(0, None, ta, M, N, K, dA.data_ptr(), M, dB.data_ptr(), N, dC.data_ptr(), N)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            lib.mambo_dgemm_device(0, None, ta, M, N, K, dA.data_ptr(), M, dB.data_ptr(), N, dC.data_ptr(), N)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        print(f"gemm ta={ta} {M}^3: {dt*1e3:.2f} ms, {2*M*N*K/dt/1e12:.2f} TFLOP/s")

def df_case(name, T, tol=1e-8, sym=mb.SYM_AUTO, full_check=True):
    n = T.shape[0]
    with mb.CSAEngine(T, device=0, sym=sym) as e:
        t0 = time.perf_counter()
        st = e.df_decompose(tol=tol, verify=True)
        dt = time.perf_counter() - t0
        nf = st["num_frags"]
        coeff, omega, U, L, kind = e.df_get(0, nf)
        lam, U1, Lm = e.df_based_greedy() if (nf and not kind.any()) else (None, None, None)
    print(f"{name}: N={n} frags={nf} m={st['krylov_dim']} restarts={st['restarts']} ortho {st['ortho_dev_before']:.1e}->{st['ortho_dev_after']:.1e} "
          f"resid {st['max_rel_residual']:.1e} sweeps {st['jacobi_sweeps_max']} | {dt:.3f}s lanczos {st['seconds_lanczos']:.3f} tri {st['seconds_tridiag']:.3f} "
          f"back {st['seconds_backtransform']:.3f} frag {st['seconds_fragments']:.3f}")
    if not full_check or nf == 0:
        return
    t0 = time.perf_counter()
    fr = oracle.DF_decomposition(T, tol)
    t_cpu = time.perf_counter() - t0
    assert len(fr) == nf, (len(fr), nf)
    c_ref = np.array([f[0] for f in fr])
    print(f"   oracle {t_cpu:.3f}s; coeff max rel err {np.abs(coeff - c_ref).max() / np.abs(c_ref).max():.2e}")
    # reconstruction from the device fragments
    rec = np.zeros((n, n, n, n))
    for k in range(nf):
        Lk = (U[k] * omega[k]) @ U[k].T
        assert np.abs(Lk - L[k]).max() < 1e-12 * max(1.0, np.abs(L[k]).max()), np.abs(Lk - L[k]).max()
        rec += coeff[k] * np.einsum("ab,cd->abcd", Lk, Lk)
    ref = oracle.DF_to_tbt(fr, n)
    print(f"   sum of fragments vs oracle's: {np.abs(rec - ref).max() / np.abs(ref).max():.2e}; vs tensor: {np.abs(rec - T).max() / np.abs(T).max():.2e}")
    orth = max(np.abs(U[k].T @ U[k] - np.eye(n)).max() for k in range(nf))
    print(f"   max |U^T U - I| = {orth:.2e}")
    if lam is not None:
        lam0, U10, Lm0 = oracle.DF_based_greedy(T, tol, fix_sign=True)
        # the frame of the leading fragment is fixed up to the signs of its columns; Lmat is invariant under those
        print(f"   DF_based_greedy: lambda rel err {np.abs(lam - lam0).max() / np.abs(lam0).max():.2e}")

if __name__ == "__main__":
    gemm_check()
    df_case("dense-sym", oracle.dense_sym_tbt(6, 0))
    df_case("dense-sym", oracle.dense_sym_tbt(10, 1))
    df_case("dense-sym forced pair", oracle.dense_sym_tbt(7, 2), sym=mb.SYM_PAIR)
    pl = oracle.planted_tbt(8, 2, 3)[0]
    df_case("planted R=2 (low rank)", pl)
    df_case("zero-ish", 1e-12 * oracle.dense_sym_tbt(5, 0))
    # degenerate eigenvalues: T = sum of two fragments L (x) L with equal coefficients, orthogonal L's
    n = 6
    rng = np.random.default_rng(5)
    A = rng.standard_normal((n, n)); A = A + A.T; B = rng.standard_normal((n, n)); B = B + B.T
    A /= np.linalg.norm(A); B -= A * np.sum(A * B); B /= np.linalg.norm(B)
    Tdeg = 0.7 * (np.einsum("ab,cd->abcd", A, A) + np.einsum("ab,cd->abcd", B, B))
    df_case("degenerate pair", Tdeg)
    df_case("dense-sym", oracle.dense_sym_tbt(24, 4))
    df_case("dense-sym N=50", oracle.dense_sym_tbt(50, 0), full_check=False)
    if len(sys.argv) > 1:
        df_case("dense-sym N=100", oracle.dense_sym_tbt(100, 0), full_check=False)
