"""Host-side mirrors of the reference's containers on the CSA path (src/UTILS/structures.jl).

Only what crosses the engine boundary is mirrored: F_OP (:23-47), cartan_2b (:117-124), cartan_SD (:137-142),
real_orbital_rotation (:165-171), F_FRAG (:258-279) and the CSA / CSA_SD tags (:228-236).  They are plain
containers; numerics live in the CUDA library.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np


class CSA:      # structures.jl:228-231
    name = "CSA"


class CSA_SD:   # structures.jl:233-236
    name = "CSA_SD"


@dataclass
class F_OP:
    """structures.jl:23-47.  mbts[i] is the i-body tensor (mbts[0] the constant); filled[i] False => zero."""
    mbts: Tuple
    spin_orb: bool = False
    Nbods: int = field(init=False)
    filled: list = field(init=False)
    N: int = field(init=False)

    def __post_init__(self):
        self.Nbods = len(self.mbts) - 1
        self.filled = [False] * (self.Nbods + 1)
        self.N = 0
        for i, t in enumerate(self.mbts):
            a = np.asarray(t)
            if not (a.size == 1 and float(a.ravel()[0]) == 0.0):
                self.filled[i] = True
                self.N = a.shape[0]

    @property
    def obt(self) -> Optional[np.ndarray]:
        return np.asarray(self.mbts[1], dtype=np.float64) if self.Nbods >= 1 and self.filled[1] else None

    @property
    def tbt(self) -> Optional[np.ndarray]:
        return np.asarray(self.mbts[2], dtype=np.float64) if self.Nbods >= 2 and self.filled[2] else None


@dataclass
class cartan_2b:            # structures.jl:117-124
    spin_orb: bool
    lam: np.ndarray         # N(N+1)/2, lower-triangular row-major (fermionic.jl:23-30)
    N: int


@dataclass
class cartan_SD:            # structures.jl:137-142
    spin_orb: bool
    lam1: np.ndarray
    lam2: np.ndarray
    N: int


@dataclass
class real_orbital_rotation:  # structures.jl:165-171
    N: int
    thetas: np.ndarray        # N(N-1)/2, upper-triangular row-major (unitaries.jl:36-43)


@dataclass
class F_FRAG:               # structures.jl:258-279
    nUs: int
    U: tuple
    TECH: object
    C: object
    N: int
    spin_orb: bool
    coeff: float = 1.0
    has_coeff: bool = False

    def x(self) -> np.ndarray:
        """The optimiser vector this fragment was built from (inverse of CSA(_SD)_x_to_F_FRAG)."""
        if isinstance(self.TECH, CSA) or self.TECH is CSA:
            return np.concatenate([self.C.lam, self.U[0].thetas])
        return np.concatenate([self.C.lam1, self.C.lam2, self.U[0].thetas])


def cartan_2b_num_params(N: int) -> int:        # fermionic.jl:50-52
    return N * (N + 1) // 2


def cartan_SD_num_params(N: int) -> int:        # fermionic.jl:54-56
    return N + cartan_2b_num_params(N)


def real_orbital_rotation_num_params(N: int) -> int:   # unitaries.jl:48-51
    return N * (N - 1) // 2


def CSA_x_to_F_FRAG(x, N: int, spin_orb: bool = False, cartan_L: Optional[int] = None) -> F_FRAG:
    """fermionic.jl:413-419 (CSA_GIVENS = false)."""
    cL = cartan_2b_num_params(N) if cartan_L is None else cartan_L
    x = np.asarray(x, dtype=np.float64)
    return F_FRAG(1, (real_orbital_rotation(N, x[cL:].copy()),), CSA(), cartan_2b(spin_orb, x[:cL].copy(), N), N, spin_orb)


def CSA_SD_x_to_F_FRAG(x, N: int, spin_orb: bool = False, cartan_L: Optional[int] = None) -> F_FRAG:
    """fermionic.jl:421-427 (CSA_GIVENS = false)."""
    cL = cartan_2b_num_params(N) if cartan_L is None else cartan_L
    x = np.asarray(x, dtype=np.float64)
    return F_FRAG(1, (real_orbital_rotation(N, x[cL + N:].copy()),), CSA_SD(),
                  cartan_SD(spin_orb, x[:N].copy(), x[N:N + cL].copy(), N), N, spin_orb)


def CSA_L1(F: F_FRAG) -> float:
    """lcu.jl:139-182 -- O(N^2) post-processing on the fragment's lambda (no tensor work)."""
    lam = F.C.lam if hasattr(F.C, "lam") else F.C.lam2
    l1, idx = 0.0, 0
    for i in range(F.N):
        for j in range(i + 1):
            if F.spin_orb:
                if i != j:
                    l1 += 0.5 * abs(lam[idx])
            else:
                l1 += (2.0 if i != j else 0.5) * abs(lam[idx])
            idx += 1
    return l1


def DF_L1(coeff, omega, spin_orb: bool = False) -> float:
    """lcu.jl:210-221: 1-norm of a double-factorisation fragment, |coeff| (sum |lambda|)^2 / 2 (orbitals only, like the reference);
    `coeff`, `omega` as returned per fragment by CSAEngine.df_get."""
    if spin_orb:
        raise ValueError("1-norm DF calculation not implemented for spin-orb=true")
    return 0.5 * abs(float(coeff)) * float(np.sum(np.abs(omega))) ** 2


def ob_correction(tbt, spin_orb: bool = False) -> np.ndarray:
    """fermionic.jl:457-479: one-body correction carried by the two-body tensor, (2x) sum_r tbt[:, :, r, r]."""
    tbt = np.asarray(tbt, dtype=np.float64)
    obt = np.einsum("pqrr->pq", tbt)
    return obt if spin_orb else 2.0 * obt


def one_body_L1(H: F_OP) -> float:
    """lcu.jl:262-267: 1-norm of the corrected one-body term, sum |eigenvalues| of obt + ob_correction (to_OBF,
    fermionic.jl:527-538; OBF_L1, lcu.jl:239-260).  O(N^3) post-processing, stays on the host as in the reference."""
    M = np.asarray(H.obt if H.obt is not None else np.zeros((H.N, H.N))) + ob_correction(H.tbt, H.spin_orb)
    D = np.linalg.eigvalsh(M) if np.allclose(M, M.T) else np.real(np.linalg.eigvals(M))
    return float(np.sum(np.abs(D))) * (0.5 if H.spin_orb else 1.0)
