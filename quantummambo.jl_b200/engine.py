"""CSAEngine: object wrapper over one libmambo_b200 handle (one GPU, one residual tensor).

Host-side mirror of the closures the reference builds in CSA_greedy_step / CSA_SD_greedy_step
(/root/reference/src/UTILS/decompose.jl:96-105, 221-231): `cost(x)`, `grad(x)`, plus the residual update of
decompose.jl:56,182.  All numerics run in the CUDA library; numpy only carries buffers across the C-ABI.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib

CSA, CSA_SD = 0, 1
SYM_AUTO, SYM_GENERAL, SYM_PAIR, SYM_FULL = 0, 1, 2, 3
TBT_ON_DEVICE = 4
MODE_NAMES = {SYM_GENERAL: "general", SYM_PAIR: "pairsym", SYM_FULL: "fullsym-packed"}


def _flavour_id(flavour) -> int:
    if flavour in (CSA, "CSA"):
        return CSA
    if flavour in (CSA_SD, "CSA_SD"):
        return CSA_SD
    raise ValueError(f"unknown flavour {flavour!r}")


def _f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} values, got {a.size}")
    return a


class CSAEngine:
    """Device-resident CSA / CSA_SD cost+gradient evaluator for one two-body tensor.

    tbt : (N,N,N,N) array with Julia semantics tbt[p,q,r,s]; it is handed to the library in column-major
          order exactly as a Julia Array{Float64,4} lies in memory.
    obt : (N,N) one-body tensor (CSA_SD only).
    """

    def __init__(self, tbt, obt=None, flavour="CSA", device: int = 0, sym: int = SYM_AUTO,
                 shard: int = 0, nshards: int = 1):
        self._h = C.c_void_p()
        self._lib = _lib.load()
        tbt = np.asarray(tbt, dtype=np.float64)
        if tbt.ndim != 4 or len(set(tbt.shape)) != 1:
            raise ValueError("tbt must have shape (N,N,N,N)")
        self.N = int(tbt.shape[0])
        self.flavour = _flavour_id(flavour)
        tf = np.asfortranarray(tbt)                     # memory order of the Julia array
        of = None
        if self.flavour == CSA_SD:
            if obt is None:
                raise ValueError("CSA_SD needs obt")
            of = np.asfortranarray(np.asarray(obt, dtype=np.float64))
        _lib.check(self._lib.mambo_csa_create_sharded(
            C.byref(self._h), self.N, self.flavour, tf.ctypes.data, None if of is None else of.ctypes.data,
            device, sym, shard, nshards))
        self.device = device
        self.L = self._lib.mambo_csa_num_params(self._h)

    @classmethod
    def from_device(cls, d_tbt_ptr: int, N: int, d_obt_ptr: int = 0, flavour="CSA", device: int = 0, sym: int = SYM_AUTO,
                    shard: int = 0, nshards: int = 1) -> "CSAEngine":
        """The same handle from a tensor that already lies on `device` (MAMBO_TBT_ON_DEVICE): d_tbt_ptr points to N^4 doubles in
        the Julia memory order (element (p,q,r,s) at p + N q + N^2 r + N^3 s), d_obt_ptr to N^2 (CSA_SD).  The library packs
        its own copy; the caller's buffers may be freed afterwards."""
        self = object.__new__(cls)
        self._h = C.c_void_p()
        self._lib = _lib.load()
        self.N = int(N)
        self.flavour = _flavour_id(flavour)
        if self.flavour == CSA_SD and not d_obt_ptr:
            raise ValueError("CSA_SD needs obt")
        _lib.check(self._lib.mambo_csa_create_sharded(
            C.byref(self._h), self.N, self.flavour, C.c_void_p(d_tbt_ptr), C.c_void_p(d_obt_ptr) if d_obt_ptr else None,
            device, sym | TBT_ON_DEVICE, shard, nshards))
        self.device = device
        self.L = self._lib.mambo_csa_num_params(self._h)
        return self

    # -- lifetime -----------------------------------------------------------------------------------------
    def lane(self) -> "CSAEngine":
        """Another evaluation context on the same GPU sharing this engine's device-resident residual
        (mambo_csa_create_lane): drive it from its own host thread / stream to overlap independent evaluations."""
        other = object.__new__(CSAEngine)
        other._h = C.c_void_p()
        other._lib = self._lib
        other.N, other.flavour, other.device, other.L = self.N, self.flavour, self.device, self.L
        _lib.check(self._lib.mambo_csa_create_lane(self._h, C.byref(other._h)))
        return other

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.mambo_csa_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def handle(self):
        return self._h

    def info(self) -> dict:
        s = _lib.MamboInfo()
        _lib.check(self._lib.mambo_csa_get_info(self._h, C.byref(s)))
        d = {k: getattr(s, k) for k, _ in s._fields_}
        d["mode_name"] = MODE_NAMES.get(d["mode"], "?")
        return d

    # -- closures (host buffers) ---------------------------------------------------------------------------
    def cost(self, x) -> float:
        x = _f64(x, self.L)
        c = C.c_double()
        _lib.check(self._lib.mambo_csa_cost(self._h, x.ctypes.data, C.byref(c)))
        return c.value

    def grad(self, x) -> np.ndarray:
        x = _f64(x, self.L)
        g = np.empty(self.L)
        _lib.check(self._lib.mambo_csa_grad(self._h, x.ctypes.data, g.ctypes.data))
        return g

    def cost_grad(self, x):
        x = _f64(x, self.L)
        g = np.empty(self.L)
        c = C.c_double()
        _lib.check(self._lib.mambo_csa_cost_grad(self._h, x.ctypes.data, C.byref(c), g.ctypes.data))
        return c.value, g

    # -- greedy state ---------------------------------------------------------------------------------------
    def subtract_fragment(self, x) -> float:
        x = _f64(x, self.L)
        c = C.c_double()
        _lib.check(self._lib.mambo_csa_subtract_fragment(self._h, x.ctypes.data, C.byref(c)))
        return c.value

    def residual_cost(self) -> float:
        c = C.c_double()
        _lib.check(self._lib.mambo_csa_residual_cost(self._h, C.byref(c)))
        return c.value

    def get_residual(self):
        n = self.N
        t = np.empty((n, n, n, n), order="F")
        o = np.empty((n, n), order="F") if self.flavour == CSA_SD else None
        _lib.check(self._lib.mambo_csa_get_residual(self._h, t.ctypes.data, None if o is None else o.ctypes.data))
        return t, o

    def one_body_matrix(self, spin_orb: bool = False, obt=None) -> np.ndarray:
        """obt + ob_correction of the RESIDENT residual (mambo_csa_one_body_matrix; lcu.jl:262-267, fermionic.jl:457-479): the
        matrix whose eigenvalues give one_body_L1.  CSA_SD engines use their own one-body residual, CSA engines add `obt`."""
        n = self.N
        out = np.empty((n, n), order="F")
        oin = None if obt is None else np.asfortranarray(np.asarray(obt, dtype=np.float64))
        _lib.check(self._lib.mambo_csa_one_body_matrix(self._h, 1 if spin_orb else 0, None if oin is None else oin.ctypes.data,
                                                       out.ctypes.data))
        return out

    def one_body_L1(self, spin_orb: bool = False, obt=None) -> float:
        """lcu.jl:262-267 for the resident residual: sum |eigenvalues| of one_body_matrix (the N x N eigen stays on the host)."""
        M = self.one_body_matrix(spin_orb, obt)
        D = np.linalg.eigvalsh(M) if np.allclose(M, M.T) else np.real(np.linalg.eigvals(M))
        return float(np.sum(np.abs(D))) * (0.5 if spin_orb else 1.0)

    def reconstruct(self, x):
        x = _f64(x, self.L)
        n = self.N
        t = np.empty((n, n, n, n), order="F")
        o = np.empty((n, n), order="F") if self.flavour == CSA_SD else None
        _lib.check(self._lib.mambo_csa_reconstruct(self._h, x.ctypes.data, t.ctypes.data, None if o is None else o.ctypes.data))
        return t, o

    def leading_eig(self, tol: float = 0.0, max_iter: int = 0):
        """(eigenvalue, N x N operator, info) of largest |eigenvalue| of the residual's N^2 x N^2 matrix view
        (mambo_csa_leading_eig: device Lanczos; replaces the full `eigen` of guesses.jl:55-63)."""
        n = self.N
        lam, it, res = C.c_double(), C.c_int(), C.c_double()
        vec = np.empty((n, n), order="F")
        _lib.check(self._lib.mambo_csa_leading_eig(self._h, tol, max_iter, C.byref(lam), vec.ctypes.data, C.byref(it), C.byref(res)))
        return lam.value, vec, {"iterations": it.value, "residual": res.value}

    # -- double factorisation (decompose.jl:297-436) ------------------------------------------------------------
    def df_decompose(self, tol: float = 0.0, max_frags: int = 0, verify: bool = False) -> dict:
        """DF_decomposition of the resident residual on the device (mambo_csa_df_decompose).  Returns the statistics; the
        fragments stay on the device until fetched with `df_get`."""
        nf = C.c_int()
        st = _lib.DfStats()
        _lib.check(self._lib.mambo_csa_df_decompose(self._h, tol, max_frags, 1 if verify else 0, C.byref(nf), C.byref(st)))
        out = {k: getattr(st, k) for k, _ in st._fields_}
        out["num_frags"] = nf.value
        return out

    def df_get(self, first: int = 0, count: int | None = None, want_L: bool = True, want_U: bool = True):
        """(coeff [k], omega [k, N], U [k, N, N], L [k, N, N], kind [k]) of fragments [first, first + count): U[e] and L[e] are
        ordinary numpy matrices (U[e][:, l] = eigenvector l)."""
        n = self.N
        if count is None:
            raise ValueError("count is required")
        coeff = np.empty(count)
        omega = np.empty((count, n))
        kind = np.empty(count, dtype=np.int32)
        U = np.empty((count, n, n)) if want_U else None
        L = np.empty((count, n, n)) if want_L else None
        _lib.check(self._lib.mambo_csa_df_get(self._h, first, count, coeff.ctypes.data, omega.ctypes.data,
                                              U.ctypes.data if want_U else None, L.ctypes.data if want_L else None, kind.ctypes.data))
        # the library writes column-major N x N blocks: as C-ordered numpy blocks they are the transposes
        if want_U:
            U = U.transpose(0, 2, 1)
        if want_L:
            L = L.transpose(0, 2, 1)
        return coeff, omega, U, L, kind

    def df_based_greedy(self):
        """DF_based_greedy (decompose.jl:400-436) from the last `df_decompose`: (lambda (cL), U1 (N x N), Lmat (N x N))."""
        n = self.N
        lam = np.empty(n * (n + 1) // 2)
        U1 = np.empty((n, n), order="F")
        Lm = np.empty((n, n), order="F")
        _lib.check(self._lib.mambo_csa_df_based_greedy(self._h, lam.ctypes.data, U1.ctypes.data, Lm.ctypes.data))
        return lam, U1, Lm

    def symv_device(self, d_v_ptr: int, d_y_ptr: int):
        """y = T~ v on the device (rows_padded doubles each): one HBM pass over the resident residual."""
        _lib.check(self._lib.mambo_csa_symv_device(self._h, d_v_ptr, d_y_ptr))

    # -- native optimiser / multistart ---------------------------------------------------------------------------
    def optimize_reserve(self, lbfgs_memory: int = 0):
        """Pre-allocate the optimiser's workspaces (mambo_csa_optimize_reserve): call on every lane / shard sharing a
        GPU before concurrent `optimize` runs start."""
        _lib.check(self._lib.mambo_csa_optimize_reserve(self._h, lbfgs_memory))

    def optimize(self, x0, g_tol: float = 1e-8, max_iter: int = 1000, lbfgs_memory: int = 0, verbose: int = 0):
        """Minimise cost(x) from x0 with the library's BFGS (mambo_csa_optimize)."""
        x0 = _f64(x0, self.L)
        xm = np.empty(self.L)
        opt = _lib.OptOptions(max_iter, g_tol, lbfgs_memory, verbose)
        res = _lib.OptResult()
        _lib.check(self._lib.mambo_csa_optimize(self._h, x0.ctypes.data, xm.ctypes.data, C.byref(opt), C.byref(res)))
        return xm, {k: getattr(res, k) for k, _ in res._fields_}

    def set_wait_mode(self, blocking: bool):
        """mambo_csa_set_wait_mode: spin (False, default) or sleep on a blocking-sync event (True) inside the host closures."""
        _lib.check(self._lib.mambo_csa_set_wait_mode(self._h, 1 if blocking else 0))

    # -- device-resident entry points (torch tensors carry the device memory; no host sync) -------------------
    def set_stream(self, cuda_stream_ptr: int | None):
        _lib.check(self._lib.mambo_csa_set_stream(self._h, cuda_stream_ptr))

    def eval_device(self, d_x_ptr: int, d_out_ptr: int):
        _lib.check(self._lib.mambo_csa_eval_device(self._h, d_x_ptr, d_out_ptr))

    def cost_device(self, d_x_ptr: int, d_cost_ptr: int):
        _lib.check(self._lib.mambo_csa_cost_device(self._h, d_x_ptr, d_cost_ptr))

    def eval_partial_device(self, d_x_ptr: int, d_partial_ptr: int):
        _lib.check(self._lib.mambo_csa_eval_partial_device(self._h, d_x_ptr, d_partial_ptr))

    def eval_finish_device(self, d_partial_ptr: int, d_out_ptr: int):
        _lib.check(self._lib.mambo_csa_eval_finish_device(self._h, d_partial_ptr, d_out_ptr))

    # -- peer exchange of row-panel shards (NVLink peer memory, no host collective) ----------------------------------
    def peer_export(self) -> bytes:
        buf = (C.c_ubyte * 64)()
        _lib.check(self._lib.mambo_csa_peer_export(self._h, buf))
        return bytes(buf)

    def peer_attach(self, rank: int, handle: bytes):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        _lib.check(self._lib.mambo_csa_peer_attach(self._h, rank, buf))

    def peer_attach_local(self, rank: int, peer: "CSAEngine"):
        _lib.check(self._lib.mambo_csa_peer_attach_local(self._h, rank, peer._h))

    def set_peer_wait(self, stream_ordered: bool):
        """mambo_csa_set_peer_wait: let the stream (not a spinning kernel) wait for the peers' flags -- needed for lanes on shards."""
        _lib.check(self._lib.mambo_csa_set_peer_wait(self._h, 1 if stream_ordered else 0))

    def eval_sharded_push_device(self, d_x_ptr: int):
        _lib.check(self._lib.mambo_csa_eval_sharded_push_device(self._h, d_x_ptr))

    def eval_sharded_finish_device(self, d_out_ptr: int):
        _lib.check(self._lib.mambo_csa_eval_sharded_finish_device(self._h, d_out_ptr))

    def eval_sharded_device(self, d_x_ptr: int, d_out_ptr: int):
        _lib.check(self._lib.mambo_csa_eval_sharded_device(self._h, d_x_ptr, d_out_ptr))

    def set_profiling(self, enable: bool):
        _lib.check(self._lib.mambo_csa_set_profiling(self._h, 1 if enable else 0))

    def get_segments(self) -> dict:
        seg = (C.c_double * 3)()
        ne = C.c_int()
        _lib.check(self._lib.mambo_csa_get_segments(self._h, seg, C.byref(ne)))
        k = max(ne.value, 1)
        return {"evals": ne.value, "pre_ms": seg[0] / k, "fused_post_ms": seg[1] / k, "finish_ms": seg[2] / k}

    def get_profile(self) -> dict:
        ms, nl, tot = C.c_double(), C.c_int(), C.c_longlong()
        _lib.check(self._lib.mambo_csa_get_profile(self._h, C.byref(ms), C.byref(nl), C.byref(tot)))
        return {"fused_ms": ms.value, "fused_launches": nl.value, "total_launches": tot.value}

    def synchronize(self):
        _lib.check(self._lib.mambo_csa_synchronize(self._h))


def cost_grad_batch(engines, xs):
    """cost and gradient at the K rows of `xs` in one library call (mambo_csa_cost_grad_batch): point k runs on
    engines[k % len(engines)] (an engine and lanes created from it); returns (costs [K], grads [K, L])."""
    engines = list(engines)
    lib = engines[0]._lib
    L = engines[0].L
    xs = np.ascontiguousarray(xs, dtype=np.float64).reshape(-1, L)
    K = xs.shape[0]
    costs, grads = np.empty(K), np.empty((K, L))
    hs = (C.c_void_p * len(engines))(*[e._h.value for e in engines])
    _lib.check(lib.mambo_csa_cost_grad_batch(hs, len(engines), K, xs.ctypes.data, costs.ctypes.data, grads.ctypes.data))
    return costs, grads


def multistart(engines, x0s, g_tol: float = 1e-8, max_iter: int = 1000, lbfgs_memory: int = 0, verbose: int = 0):
    """K independent BFGS runs from the rows of `x0s`, scheduled over `engines` by the library
    (mambo_multistart_run: one worker thread per handle pulling starts from a shared counter; precedent
    orbitals.jl:39-67).  `engines` may be handles on different GPUs and/or lanes of one GPU (CSAEngine.lane()).
    Returns (x_min [K, L], list of K result dicts, index of the best start)."""
    engines = list(engines)
    lib = engines[0]._lib
    L = engines[0].L
    x0s = np.ascontiguousarray(x0s, dtype=np.float64).reshape(-1, L)
    K = x0s.shape[0]
    xm = np.empty((K, L))
    hs = (C.c_void_p * len(engines))(*[e._h.value for e in engines])
    opt = _lib.OptOptions(max_iter, g_tol, lbfgs_memory, verbose)
    res = (_lib.OptResult * K)()
    best = C.c_int(-1)
    _lib.check(lib.mambo_multistart_run(hs, len(engines), K, x0s.ctypes.data, xm.ctypes.data, C.byref(opt), res, C.byref(best)))
    return xm, [{k: getattr(r, k) for k, _ in r._fields_} for r in res], best.value


def frag_l1(x, N: int, flavour="CSA", spin_orb: bool = False) -> float:
    """CSA_L1 of the fragment with parameter vector x (mambo_csa_frag_l1; lcu.jl:139-182)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = C.c_double()
    _lib.check(_lib.load().mambo_csa_frag_l1(int(N), _flavour_id(flavour), 1 if spin_orb else 0, x.ctypes.data, C.byref(out)))
    return out.value


def write_x_block(path: str, group: str, X) -> None:
    """The checkpoint dataset "<group>/x" (decompose.jl:58-64; tot_L x alpha, column-major) as the library's raw block file."""
    X = np.asfortranarray(np.asarray(X, dtype=np.float64))
    if X.ndim != 2:
        raise ValueError("X must be tot_L x alpha")
    _lib.check(_lib.load().mambo_csa_write_x_block(os.fsencode(path), group.encode(), X.shape[0], X.shape[1], X.ctypes.data))


def read_x_block(path: str):
    """(group, X [tot_L, alpha]) from a file written by write_x_block / mambo_csa_write_x_block."""
    lib = _lib.load()
    name = C.create_string_buffer(16)
    tot_L, alpha = C.c_longlong(), C.c_longlong()
    _lib.check(lib.mambo_csa_read_x_block(os.fsencode(path), name, C.byref(tot_L), C.byref(alpha), None, 0))
    X = np.empty((tot_L.value, alpha.value), order="F")
    _lib.check(lib.mambo_csa_read_x_block(os.fsencode(path), name, C.byref(tot_L), C.byref(alpha), X.ctypes.data, X.size))
    return name.value.decode(), X
