"""ctypes binding of libmambo_b200.so (the C-ABI declared in include/mambo_csa.h).

This is the same binding a Julia `ccall` shim performs; Python is only the host
language available in this image.  The library is built in-tree by `build()` (nvcc, sm_100a) and there is
no fallback: if it is missing or no B200 is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = CSRC / "libmambo_b200.so"
INCLUDE = PKG_DIR.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static"]
SOURCES = ["mambo_csa.cu", "optimizer.cu", "df.cu"]


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the CUDA engine for sm_100a into csrc/libmambo_b200.so (no GPU needed: nvcc cross-compiles).
    Every translation unit is compiled to its own object (in parallel, only when it or a header changed), then linked."""
    from concurrent.futures import ThreadPoolExecutor
    import re
    srcs = [CSRC / s for s in SOURCES]
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

    def deps_of(path: Path, seen=None):
        """the file and, transitively, every header it includes with quotes"""
        seen = set() if seen is None else seen
        path = path.resolve()
        if path in seen or not path.exists():
            return seen
        seen.add(path)
        for inc in re.findall(r'^\s*#include\s+"([^"]+)"', path.read_text(), flags=re.M):
            deps_of(path.parent / inc, seen)
        return seen

    def compile_one(src: Path):
        obj = src.with_suffix(".o")
        if not force and obj.exists() and obj.stat().st_mtime >= max(d.stat().st_mtime for d in deps_of(src)):
            return obj, False
        cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", str(obj), str(src)]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True, cwd=str(CSRC))
        return obj, True

    with ThreadPoolExecutor(max_workers=len(srcs)) as pool:
        done = list(pool.map(compile_one, srcs))
    objs = [o for o, _ in done]
    if force or not LIB_PATH.exists() or any(ch for _, ch in done) or any(LIB_PATH.stat().st_mtime < o.stat().st_mtime for o in objs):
        cmd = [nvcc, *LINK_FLAGS, "-o", str(LIB_PATH), *map(str, objs)]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True, cwd=str(CSRC))
    return LIB_PATH


class MamboInfo(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("flavour", C.c_int), ("mode", C.c_int), ("num_params", C.c_int),
        ("rows", C.c_int), ("rows_padded", C.c_int), ("shard", C.c_int), ("nshards", C.c_int),
        ("local_row_begin", C.c_int), ("local_row_end", C.c_int), ("partial_len", C.c_int),
        ("num_sms", C.c_int), ("launches_per_eval", C.c_int),
        ("flops_per_eval", C.c_double), ("tensor_bytes", C.c_double),
        ("asym_pair", C.c_double), ("asym_pq", C.c_double),
        ("hot_gemms", C.c_int), ("lanes_hint", C.c_int),
    ]


class OptOptions(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("g_tol", C.c_double), ("lbfgs_memory", C.c_int), ("verbose", C.c_int)]


class OptResult(C.Structure):
    _fields_ = [("minimum", C.c_double), ("iterations", C.c_int), ("evaluations", C.c_int),
                ("converged", C.c_int), ("device", C.c_int)]


class DfStats(C.Structure):
    _fields_ = [("krylov_dim", C.c_int), ("restarts", C.c_int), ("num_frags", C.c_int), ("anti_frags", C.c_int),
                ("jacobi_sweeps_max", C.c_int), ("cluster_vectors", C.c_int), ("ortho_dev_before", C.c_double), ("ortho_dev_after", C.c_double),
                ("max_rel_residual", C.c_double), ("seconds_lanczos", C.c_double), ("seconds_tridiag", C.c_double),
                ("seconds_backtransform", C.c_double), ("seconds_fragments", C.c_double)]


_P = C.c_void_p
_D = C.POINTER(C.c_double)

# name -> (restype, argtypes); every symbol include/mambo_csa.h declares
SIGNATURES = {
    "mambo_csa_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, _P, _P, C.c_int, C.c_uint]),
    "mambo_csa_create_sharded": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, _P, _P, C.c_int, C.c_uint, C.c_int, C.c_int]),
    "mambo_csa_create_lane": (C.c_int, [_P, C.POINTER(_P)]),
    "mambo_csa_destroy": (None, [_P]),
    "mambo_csa_get_info": (C.c_int, [_P, C.POINTER(MamboInfo)]),
    "mambo_csa_num_params": (C.c_int, [_P]),
    "mambo_last_error": (C.c_char_p, []),
    "mambo_csa_cost": (C.c_int, [_P, _P, _D]),
    "mambo_csa_grad": (C.c_int, [_P, _P, _P]),
    "mambo_csa_cost_grad": (C.c_int, [_P, _P, _D, _P]),
    "mambo_csa_cost_grad_batch": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, _P, _P, _P]),
    "mambo_csa_set_wait_mode": (C.c_int, [_P, C.c_int]),
    "mambo_csa_one_body_matrix": (C.c_int, [_P, C.c_int, _P, _P]),
    "mambo_csa_frag_l1": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _D]),
    "mambo_csa_write_x_block": (C.c_int, [C.c_char_p, C.c_char_p, C.c_longlong, C.c_longlong, _P]),
    "mambo_csa_read_x_block": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), _P, C.c_longlong]),
    "mambo_csa_subtract_fragment": (C.c_int, [_P, _P, _D]),
    "mambo_csa_residual_cost": (C.c_int, [_P, _D]),
    "mambo_csa_get_residual": (C.c_int, [_P, _P, _P]),
    "mambo_csa_reconstruct": (C.c_int, [_P, _P, _P, _P]),
    "mambo_csa_leading_eig": (C.c_int, [_P, C.c_double, C.c_int, _D, _P, C.POINTER(C.c_int), _D]),
    "mambo_csa_symv_device": (C.c_int, [_P, _P, _P]),
    "mambo_debug_tridiag_leading": (C.c_int, [_P, _P, C.c_int, _D, _P]),
    "mambo_debug_tridiag_cluster": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P]),
    "mambo_csa_set_stream": (C.c_int, [_P, _P]),
    "mambo_csa_get_stream": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int)]),
    "mambo_csa_eval_device": (C.c_int, [_P, _P, _P]),
    "mambo_csa_eval_partial_device": (C.c_int, [_P, _P, _P]),
    "mambo_csa_eval_finish_device": (C.c_int, [_P, _P, _P]),
    "mambo_csa_peer_export": (C.c_int, [_P, _P]),
    "mambo_csa_peer_attach": (C.c_int, [_P, C.c_int, _P]),
    "mambo_csa_peer_attach_local": (C.c_int, [_P, C.c_int, _P]),
    "mambo_csa_set_peer_wait": (C.c_int, [_P, C.c_int]),
    "mambo_csa_eval_sharded_push_device": (C.c_int, [_P, _P]),
    "mambo_csa_eval_sharded_finish_device": (C.c_int, [_P, _P]),
    "mambo_csa_eval_sharded_device": (C.c_int, [_P, _P, _P]),
    "mambo_csa_cost_device": (C.c_int, [_P, _P, _P]),
    "mambo_csa_synchronize": (C.c_int, [_P]),
    "mambo_csa_set_profiling": (C.c_int, [_P, C.c_int]),
    "mambo_csa_get_segments": (C.c_int, [_P, _D, C.POINTER(C.c_int)]),
    "mambo_csa_get_profile": (C.c_int, [_P, _D, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    "mambo_csa_df_decompose": (C.c_int, [_P, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(DfStats)]),
    "mambo_csa_df_get": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "mambo_csa_df_based_greedy": (C.c_int, [_P, _P, _P, _P]),
    "mambo_dgemm_device": (C.c_int, [C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_longlong, _P, C.c_longlong, _P, C.c_longlong]),
    "mambo_csa_optimize": (C.c_int, [_P, _P, _P, C.POINTER(OptOptions), C.POINTER(OptResult)]),
    "mambo_csa_optimize_reserve": (C.c_int, [_P, C.c_int]),
    "mambo_debug_bfgs_pass": (C.c_int, [C.c_int, C.c_int, C.c_int, _D, _D]),
    "mambo_multistart_run": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, _P, _P, C.POINTER(OptOptions),
                                       C.POINTER(OptResult), C.POINTER(C.c_int)]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the engine; fails loudly when the shared library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the CSA engine has no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class MamboError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmambo_b200 error {code}: {msg}")
        self.code = code


def check(rc: int) -> None:
    if rc != 0:
        raise MamboError(rc, load().mambo_last_error().decode(errors="replace"))
