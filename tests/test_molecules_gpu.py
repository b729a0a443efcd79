"""BASELINE configs[0..1] on molecular integrals (tests/golden/molecules.npz: LiH N=6, H2O N=7, N2 N=10, STO-3G, 1 Angstrom,
generated and pinned by tools/molecular_integrals.py / tests/test_molecules_cpu.py): the CUDA path through the C-ABI
against the oracle.  Tolerances (north_star): cost / gradient 1e-10 relative, final 1-norms 1e-8."""
import os

import numpy as np
import pytest

import csa_oracle as o

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOLS = {"lih": 6, "h2o": 7, "n2": 10}


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(ROOT, "tests", "golden", "molecules.npz"))


@pytest.fixture(scope="module")
def lib(mb):
    if not os.path.exists(os.path.join(os.path.dirname(mb.__file__), "quantummambo.jl_b200", "csrc", "libmambo_b200.so")):
        mb.build()
    return mb.load()


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


@pytest.mark.parametrize("name", MOLS)
@pytest.mark.parametrize("shift", ["", "_bliss"])
@pytest.mark.parametrize("flavour", ["CSA", "CSA_SD"])
def test_cost_and_gradient_on_molecular_tensors(mb, lib, fx, name, shift, flavour):
    """gradient.jl / cost.jl on the molecules' tensors, plain and BLISS-shifted (bliss.jl:2-28), at random points and at the
    SVD start of a greedy step (guesses.jl:51-84)."""
    n = MOLS[name]
    T, h = fx[f"{name}{shift}_tbt"], fx[f"{name}{shift}_obt"]
    obt = h if flavour == "CSA_SD" else None
    off = n if flavour == "CSA_SD" else 0
    cL = o.cartan_2b_num_params(n)
    lam, th = o.tbt_svd_1st(T)
    x_svd = np.zeros(off + cL + len(th))
    x_svd[off:off + cL] = lam
    x_svd[off + cL:] = th
    with mb.CSAEngine(T, obt, flavour) as e:
        assert e.info()["mode_name"] == "fullsym-packed"
        for x in [o.random_x(n, s, flavour) for s in (0, 1)] + [x_svd]:
            c, g = e.cost_grad(x)
            c0, g0 = o.cost_grad_factored(x, T, obt, flavour)
            assert abs(c - c0) <= 1e-10 * abs(c0)
            assert rel(g, g0) <= 1e-10


def test_lih_greedy_csa_step_by_step_against_the_oracle(mb, lib, fx):
    """configs[0] (`julia L1.jl lih`, wrappers.jl:419-425: CSA_greedy_decomposition, then lambda = one_body_L1(H) + sum
    CSA_L1(fragments)).  Greedy trajectories on a real molecule are chaotic in the last digits of x0 (two oracle runs whose
    starts differ by 1e-12 end 1e-5 apart after 20 fragments), so the comparison is made step by step on a shared residual:
    every step starts from the same x0 on the oracle's residual; the step's minimum agrees to 1e-8, every fragment's 1-norm to 5e-8,
    the total 1-norm to 1e-8 (north_star), and the engine is left with the oracle's residual."""
    n, alpha = 6, 8
    T, h = fx["lih_tbt"], fx["lih_obt"]
    cL = o.cartan_2b_num_params(n)
    x0s = [o.random_x(n, s) for s in range(alpha)]
    oxs, ocosts, t_rem, _ = o.greedy_decomposition(T, alpha, x0s=x0s, decomp_tol=1e-6, gtol=1e-10, maxiter=3000)
    assert len(oxs) == alpha
    l1, l1_o = 0.0, 0.0
    with mb.CSAEngine(T, None, "CSA") as e:
        for a in range(alpha):
            sol = mb.CSA_greedy_step(None, x0=x0s[a], engine=e, g_tol=1e-10, maxiter=3000)
            assert abs(sol.minimum - ocosts[a]) <= 1e-8 * ocosts[a] + 1e-14
            fa = mb.CSA_L1(mb.CSA_x_to_F_FRAG(sol.minimizer, n, False, cL))
            fo = o.CSA_L1(oxs[a][:cL], n)
            # a single fragment's 1-norm is conditioned worse than the total (flat directions of the minimum: two ORACLE runs whose
            # starts differ by 1e-13 differ by up to 4e-9 in one fragment's 1-norm); the north_star bound applies to the total below
            assert abs(fa - fo) <= 5e-8 * fo
            l1 += fa
            l1_o += fo
            e.subtract_fragment(oxs[a])                  # both sides continue from the oracle's residual
        Tr, _ = e.get_residual()
        assert abs(e.residual_cost() - float(np.sum(t_rem ** 2))) <= 1e-10 * float(np.sum(t_rem ** 2))
    assert np.max(np.abs(Tr - t_rem)) <= 1e-12 * np.max(np.abs(T))
    lam1 = mb.one_body_L1(mb.F_OP(([0.0], h, T)))
    assert abs(lam1 - o.one_body_L1(h, T)) <= 1e-12 * lam1
    assert abs((lam1 + l1) - (lam1 + l1_o)) <= 1e-8 * (lam1 + l1_o)


def test_lih_free_running_greedy_csa_reaches_the_tolerance(mb, lib, fx):
    """The same workflow free-running on the engine (native device optimiser, SVD starts) down to decomp_tol = 1e-6 on the
    squared norm (decompose.jl:45): the fragments plus the residual reproduce the tensor, and the 1-norm lands where the
    oracle's free-running greedy lands (1e-2: different optimisers and starts pick different local minima)."""
    n = 6
    T, h = fx["lih_tbt"], fx["lih_obt"]
    cL = o.cartan_2b_num_params(n)
    H = mb.F_OP(([0.0], h, T))
    frags = mb.CSA_greedy_decomposition(H, 100, decomp_tol=1e-6, do_svd=True, native=True, g_tol=1e-9, maxiter=2000)
    assert 5 <= len(frags) <= 100
    W = sum(o.reconstruct_factored(f.x(), n)[1] for f in frags)
    assert float(np.sum((T - W) ** 2)) <= 1e-6
    l1 = sum(mb.CSA_L1(f) for f in frags)
    oxs, _, _, _ = o.greedy_decomposition(T, 100, decomp_tol=1e-6, do_svd=True, gtol=1e-9, maxiter=2000)
    l1_o = sum(o.CSA_L1(x[:cL], n) for x in oxs)
    assert abs(l1 - l1_o) <= 5e-2 * l1_o


@pytest.mark.parametrize("name", ["h2o", "n2"])
def test_interaction_picture_step_on_bliss_shifted_molecules(mb, lib, fx, name):
    """configs[1]: the CSA_SD step of the INTERACTION routine (wrappers.jl:250-255: H0 = CSA_SD_greedy_decomposition(H, 1)[1],
    HR = H - H0) on the BLISS-shifted H2O / N2 operators, SVD start (config.jl:10).  Same start, scipy BFGS on both sides:
    the minimum, the fragment's 1-norm and the 1-norms of HR agree to 1e-8."""
    n = MOLS[name]
    T, h = fx[f"{name}_bliss_tbt"], fx[f"{name}_bliss_obt"]
    cL = o.cartan_2b_num_params(n)
    xo, co, _ = o.greedy_step(T, h, "CSA_SD", gtol=1e-10, maxiter=3000)
    H = mb.F_OP(([0.0], h, T))
    with mb.CSAEngine(T, h, "CSA_SD") as e:
        sol = mb.CSA_SD_greedy_step(None, engine=e, g_tol=1e-10, maxiter=3000)
        assert sol.minimum < 0.8 * e.residual_cost()
        assert abs(sol.minimum - co) <= 1e-8 * co
        fa = mb.CSA_L1(mb.CSA_SD_x_to_F_FRAG(sol.minimizer, n, False, cL))
        fo = o.CSA_L1(xo[n:n + cL], n)
        assert abs(fa - fo) <= 1e-8 * fo
        e.subtract_fragment(xo)
        Tr, hr = e.get_residual()
    ho, Wo = o.reconstruct_factored(xo, n, "CSA_SD")
    assert np.max(np.abs(Tr - (T - Wo))) <= 1e-12 * np.max(np.abs(T))
    ob = mb.one_body_L1(mb.F_OP(([0.0], hr, Tr)))
    assert abs(ob - o.one_body_L1(h - ho, T - Wo)) <= 1e-8 * ob


@pytest.mark.parametrize("name", MOLS)
@pytest.mark.parametrize("shift", ["", "_bliss"])
def test_double_factorisation_of_molecular_tensors(mb, lib, fx, name, shift):
    """DF_decomposition (decompose.jl:297-398) on molecular integrals: rank-deficient Coulomb matrices with exactly degenerate
    eigenvalues (the pi shells of LiH / N2) -- Lanczos restarts, and eigenvalues that coincide to rounding inside one block
    (their tridiagonal eigenvectors are redone on the host, `cluster_vectors`).  Coefficients against LAPACK to 1e-10, eigen-
    residuals, the sum of the fragments, and DF_L1 (lcu.jl:210-221) fragment by fragment wherever the eigenvalue is isolated
    (inside a degenerate eigenspace the basis, and with it sum |omega|, is LAPACK's or Lanczos' free choice)."""
    n = MOLS[name]
    T = fx[f"{name}{shift}_tbt"]
    fr = o.DF_decomposition(T)
    co = np.array([f[0] for f in fr])
    with mb.CSAEngine(T) as e:
        st = e.df_decompose(verify=True)
        nf = st["num_frags"]
        c, w, U, L, kind = e.df_get(0, nf)
    scale = np.max(np.abs(co))
    strict = np.abs(np.abs(co) - 1e-8) > 1e-11                    # eigenvalues sitting on the truncation threshold may fall either way
    assert abs(nf - len(fr)) <= int(np.sum(~strict))
    k = min(nf, len(fr))
    assert np.max(np.abs(np.abs(c[:k]) - np.abs(co[:k]))) <= 1e-10 * scale
    assert st["max_rel_residual"] < 1e-12 and not kind.any()
    rec = np.einsum("e,epq,ers->pqrs", c, L, L)
    assert np.max(np.abs(rec - T)) <= 5e-8
    for i in range(k):
        gap = min(abs(abs(co[i]) - abs(co[j])) for j in range(len(co)) if j != i)
        if gap > 1e-6 * scale and abs(c[i] - co[i]) <= 1e-10 * scale:
            l1, l1_o = 0.5 * abs(c[i]) * np.sum(np.abs(w[i])) ** 2, 0.5 * abs(co[i]) * np.sum(np.abs(fr[i][1])) ** 2
            assert abs(l1 - l1_o) <= 1e-8 * max(l1_o, 1e-6 * scale)
            assert np.max(np.abs((U[i] * w[i]) @ U[i].T - L[i])) <= 1e-12
