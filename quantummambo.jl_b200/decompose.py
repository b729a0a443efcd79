"""Greedy CSA / CSA_SD drivers with the cost+gradient hot path on the B200 engine.

Mirrors /root/reference/src/UTILS/decompose.jl:1-246 name for name: CSA_greedy_decomposition,
CSA_greedy_step, CSA_SD_greedy_decomposition, CSA_SD_greedy_step.  What changes relative to the reference:
  * the closures `cost(x)` / `grad!(g, x)` (decompose.jl:96-105, 221-231) call the CUDA library;
  * F_rem lives on the device: `F_rem = F_rem - to_OP(frag)` (decompose.jl:56,182) is `subtract_fragment`;
  * Optim.jl's BFGS is replaced here by scipy's BFGS (g_tol 1e-8, 1000 iterations: Optim's defaults) or by
    the library's native optimiser; in Julia, Optim.jl keeps driving the closures.
  * the HDF5 checkpoint (decompose.jl:17-38, 58-64) stays in Julia; this host keeps the same on-disk layout
    ("CSA"/"x" = tot_L x alpha) in an .npz because neither libhdf5 nor h5py exist in this image.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from .engine import CSAEngine, write_x_block, CSA as _CSA, CSA_SD as _CSA_SD
from .guesses import tbt_svd_1st
from .structures import (F_OP, F_FRAG, CSA_x_to_F_FRAG, CSA_SD_x_to_F_FRAG, cartan_2b_num_params,
                         real_orbital_rotation_num_params)

# config.jl
EPS = 1e-5                 # ϵ, config.jl:19 (tolerance on the *squared* norm, decompose.jl:45)
SVD_for_CSA = False        # config.jl:9
SVD_for_CSA_SD = True      # config.jl:10


@dataclass
class OptimResult:
    """The two fields of Optim's result the drivers read (decompose.jl:54,57)."""
    minimizer: np.ndarray
    minimum: float
    iterations: int = 0
    f_calls: int = 0
    converged: bool = False


def _npz_name(name: str) -> str:
    """np.savez appends '.npz' to names without it; the resume check must look for the file that was written."""
    return name if name.endswith(".npz") else name + ".npz"


def _bfgs(engine: CSAEngine, x0, g_tol: float = 1e-8, maxiter: int = 1000, native: bool = False) -> OptimResult:
    if native:
        x, res = engine.optimize(x0, g_tol=g_tol, max_iter=maxiter)
        return OptimResult(x, res["minimum"], res["iterations"], res["evaluations"], bool(res["converged"]))
    from scipy.optimize import minimize
    r = minimize(lambda x: engine.cost_grad(x), np.asarray(x0, dtype=np.float64), jac=True, method="BFGS",
                 options={"gtol": g_tol, "maxiter": maxiter})
    return OptimResult(np.asarray(r.x), float(r.fun), int(r.nit), int(r.nfev), bool(r.success))


def _x0_svd(engine: CSAEngine, offset: int) -> np.ndarray:
    """decompose.jl:87-92 / 212-217: x0 from the largest SVD fragment of the (device-resident) residual."""
    N = engine.N
    cL, uL = cartan_2b_num_params(N), real_orbital_rotation_num_params(N)
    lam, theta = tbt_svd_1st(engine)
    x0 = np.zeros(offset + cL + uL)
    x0[offset:offset + cL] = lam
    x0[offset + cL:] = theta
    return x0


def _x0_random(N: int, offset: int, rng) -> np.ndarray:
    """decompose.jl:86-94: lambda = 0, theta = 2*pi*rand (the reference draws from the unseeded global RNG)."""
    cL, uL = cartan_2b_num_params(N), real_orbital_rotation_num_params(N)
    x0 = np.zeros(offset + cL + uL)
    x0[offset + cL:] = 2 * np.pi * rng.random(uL)
    return x0


def CSA_greedy_step(F, do_svd: bool = SVD_for_CSA, x0=None, rng=None, engine: Optional[CSAEngine] = None,
                    native: bool = False, **opt) -> OptimResult:
    """decompose.jl:82-120.  `F` is an F_OP (a fresh engine is created) or ignored when `engine` holds F_rem."""
    own = engine is None
    if own:
        engine = CSAEngine(F.tbt, flavour="CSA")
    try:
        if x0 is None:
            if do_svd:
                x0 = _x0_svd(engine, 0)
            else:
                x0 = _x0_random(engine.N, 0, np.random.default_rng() if rng is None else rng)
        return _bfgs(engine, x0, native=native, **opt)
    finally:
        if own:
            engine.close()


def CSA_SD_greedy_step(F, do_svd: bool = SVD_for_CSA_SD, x0=None, rng=None, engine: Optional[CSAEngine] = None,
                       native: bool = False, **opt) -> OptimResult:
    """decompose.jl:207-246."""
    own = engine is None
    if own:
        engine = CSAEngine(F.tbt, F.obt, flavour="CSA_SD")
    try:
        if x0 is None:
            if do_svd:
                x0 = _x0_svd(engine, engine.N)
            else:
                x0 = _x0_random(engine.N, engine.N, np.random.default_rng() if rng is None else rng)
        return _bfgs(engine, x0, native=native, **opt)
    finally:
        if own:
            engine.close()


def _greedy(H: F_OP, alpha_max: int, flavour: str, decomp_tol: float, verbose: bool, SAVELOAD: bool, SAVENAME: str,
            x0s, rng, native: bool, device: int, do_svd: Optional[bool] = None, **opt):
    N = H.N
    cL, uL = cartan_2b_num_params(N), real_orbital_rotation_num_params(N)
    sd = flavour == "CSA_SD"
    if sd and H.spin_orb:
        raise ValueError("CSA_SD decomposition should be done with Hamiltonian represented in orbitals, not spin-orbitals!")
    tot_L = cL + uL + (N if sd else 0)
    group = flavour
    to_frag = CSA_SD_x_to_F_FRAG if sd else CSA_x_to_F_FRAG
    step = CSA_SD_greedy_step if sd else CSA_greedy_step
    rng = np.random.default_rng() if rng is None else rng
    if do_svd is None:
        do_svd = SVD_for_CSA_SD if sd else SVD_for_CSA          # config.jl:9-10, the steps' defaults

    Farr: list[F_FRAG] = []
    X = np.zeros((tot_L, alpha_max))
    a_curr = 0
    SAVENAME = _npz_name(SAVENAME)
    with CSAEngine(H.tbt, H.obt if sd else None, flavour=flavour, device=device) as eng:   # F_rem = copy(H) on the device
        if SAVELOAD and os.path.exists(SAVENAME):
            with np.load(SAVENAME) as z:
                if f"{group}/x" in z:
                    x = z[f"{group}/x"]
                    if x.shape[0] != tot_L:
                        raise ValueError(f"Trying to load from {SAVENAME}, saved x has wrong dimensions for {group} parameters of H!")
                    a_curr = x.shape[1]
                    print(f"Found saved x under filename {SAVENAME} for {group} decomposition, loaded {a_curr} fragments...")
                    for i in range(a_curr):                       # decompose.jl:28-32: replay onto the residual
                        Farr.append(to_frag(x[:, i], N, H.spin_orb, cL))
                        eng.subtract_fragment(x[:, i])
                    X[:, :a_curr] = x
        curr_cost = eng.residual_cost()                           # decompose.jl:40
        if verbose:
            print(f"Initial L2 cost is {curr_cost}")
        while curr_cost > decomp_tol and a_curr < alpha_max:
            x0 = None if x0s is None else x0s[a_curr]
            sol = step(None, do_svd=do_svd, x0=x0, rng=rng, engine=eng, native=native, **opt)
            a_curr += 1
            if verbose:
                print(f"Current L2 cost after {a_curr} fragments is {sol.minimum}")
            Farr.append(to_frag(sol.minimizer, N, H.spin_orb, cL))
            eng.subtract_fragment(sol.minimizer)                  # decompose.jl:56
            curr_cost = sol.minimum
            X[:, a_curr - 1] = sol.minimizer
            if SAVELOAD:                                          # decompose.jl:58-64: only this group's "x" is rewritten
                keep = {}
                if os.path.exists(SAVENAME):
                    with np.load(SAVENAME) as z:
                        keep = {k: z[k] for k in z.files if k != f"{group}/x"}
                np.savez(SAVENAME, **keep, **{f"{group}/x": X[:, :a_curr]})
                # the same dataset as the library's raw block, for the Julia side to h5write
                write_x_block(SAVENAME[:-4] + f".{group}.xblk", group, X[:, :a_curr])
        if verbose:
            print(f"Finished {group} decomposition, total number of fragments is {a_curr}, remainder L2-norm is {curr_cost}")
        if curr_cost > decomp_tol and (verbose or not sd):
            import warnings
            warnings.warn(f"{group} decomposition did not converge, remaining L2-norm is {curr_cost}")
    return Farr


def CSA_greedy_decomposition(H: F_OP, alpha_max: int, decomp_tol: float = EPS, verbose: bool = False, SAVELOAD: bool = False,
                             SAVENAME: str = "CSA.npz", x0s=None, rng=None, native: bool = False, device: int = 0,
                             do_svd: Optional[bool] = None, **opt):
    """decompose.jl:1-80: only the two-body tensor is optimised (filled[1:2] = false)."""
    return _greedy(H, alpha_max, "CSA", decomp_tol, verbose, SAVELOAD, SAVENAME, x0s, rng, native, device, do_svd, **opt)


def CSA_SD_greedy_decomposition(H: F_OP, alpha_max: int, decomp_tol: float = EPS, verbose: bool = False, SAVELOAD: bool = False,
                                SAVENAME: str = "CSA_SD.npz", x0s=None, rng=None, native: bool = False, device: int = 0,
                                do_svd: Optional[bool] = None, **opt):
    """decompose.jl:122-205: one- and two-body tensors are optimised together; every step starts from the largest SVD
    fragment of the residual (SVD_for_CSA_SD = true, config.jl:10)."""
    return _greedy(H, alpha_max, "CSA_SD", decomp_tol, verbose, SAVELOAD, SAVENAME, x0s, rng, native, device, do_svd, **opt)


def to_OP(frag: F_FRAG, device: int = 0) -> F_OP:
    """fermionic.jl:1-8 / 74-92: the fragment as an operator (reconstruction runs on the GPU)."""
    n = frag.N
    sd = not hasattr(frag.C, "lam")
    zero = np.zeros((n, n, n, n))
    with CSAEngine(zero, np.zeros((n, n)) if sd else None, flavour="CSA_SD" if sd else "CSA", device=device) as eng:
        t, o = eng.reconstruct(frag.x())
    c = frag.coeff if frag.has_coeff else 1.0
    if sd:
        return F_OP(([0], c * o, c * t), frag.spin_orb)
    return F_OP(([0.0], [0.0], c * t), frag.spin_orb)
