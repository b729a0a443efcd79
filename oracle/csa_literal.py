"""ctypes wrapper of oracle/csa_literal.c (CPU ORACLE, test infrastructure): the reference's loop nests in C,
single-threaded like the Einsum.jl loops they restate.  expm / eigen stay with numpy (oracle/csa_oracle.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

import csa_oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "libcsa_literal.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-C", HERE], check=True)
        L = C.CDLL(_SO)
        L.csa_lit_cost.restype = C.c_double
        L.csa_lit_del_w_u.restype = C.c_void_p
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def cost_grad_literal_c(x, tbt, obt=None, flavour="CSA"):
    """(cost, grad) of the closures (decompose.jl:96-105 CSA, :221-231 CSA_SD) through the C loop nests."""
    L = lib()
    n = tbt.shape[0]
    cL, uL = o.cartan_2b_num_params(n), o.real_orbital_rotation_num_params(n)
    lam1, lam, th = o.split_x(x, n, flavour)
    U = np.ascontiguousarray(o.one_body_unitary(th, n))
    Lam = np.ascontiguousarray(o.get_cartan_matrix(n, lam))
    T = np.ascontiguousarray(tbt, dtype=np.float64)
    W = np.empty_like(T)
    sd = flavour != "CSA"
    if sd:
        # fermionic.jl:58-72, 85-92: the sparse cartan tensor goes through the general N^8 rotation
        cart = np.ascontiguousarray(o.cartan_2b_to_tbt(lam, n))
        L.csa_lit_tbt_rotation(n, _p(U), _p(cart), _p(W))
    else:
        L.csa_lit_reconstruct(n, _p(U), _p(Lam), _p(W))
    cost = L.csa_lit_cost(n, _p(T), _p(W))
    diff = np.ascontiguousarray(W - T)
    gl = np.empty(cL)
    L.csa_lit_grad_lambda(n, _p(U), _p(diff), _p(gl))
    wu = L.csa_lit_del_w_u(n, _p(U), _p(Lam))
    if not wu:
        raise MemoryError("n^6 scratch")
    dU = np.ascontiguousarray(np.stack([o.del_u_theta(n, th, k) for k in range(uL)])) if uL else np.zeros((0, n, n))
    gt = np.empty(uL)
    if sd:
        lam1 = np.ascontiguousarray(lam1, dtype=np.float64)
        hT = np.ascontiguousarray(obt, dtype=np.float64)
        h = np.empty_like(hT)
        L.csa_lit_obt_rotation(n, _p(U), _p(np.ascontiguousarray(np.diag(lam1))), _p(h))
        L.csa_lit_cost_ob.restype = C.c_double
        cost += L.csa_lit_cost_ob(n, _p(hT), _p(h))
        diff_ob = np.ascontiguousarray(h - hT)
        g1 = np.empty(n)
        L.csa_lit_grad_l_ob(n, _p(U), _p(diff_ob), _p(g1))
        hu = np.empty((n, n, n, n))
        L.csa_lit_del_h_u(n, _p(U), _p(lam1), _p(hu))
        L.csa_lit_grad_theta_sd(n, C.c_void_p(wu), _p(hu), _p(dU), uL, _p(diff), _p(diff_ob), _p(gt))
        L.csa_lit_free(C.c_void_p(wu))
        return cost, np.concatenate([g1, gl, gt])
    L.csa_lit_grad_theta(n, C.c_void_p(wu), _p(dU), uL, _p(diff), _p(gt))
    L.csa_lit_free(C.c_void_p(wu))
    return cost, np.concatenate([gl, gt])
