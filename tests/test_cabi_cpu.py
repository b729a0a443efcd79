"""CPU tests of the boundary: the shared library builds for sm_100a, loads, exports every symbol that
include/mambo_csa.h declares, refuses to run without a GPU (no CPU fallback), and the product package never
touches the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mambo_csa.h")
PKG = os.path.join(ROOT, "quantummambo.jl_b200")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(mambo_[a-z_0-9]+)\s*\(", src)
    return sorted(set(names))


@pytest.fixture(scope="module")
def lib(mb):
    mb.build()
    return mb.load()


def test_library_exports_every_declared_symbol(lib):
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/mambo_csa.h but not exported"


def test_binding_table_covers_the_header(mb):
    import importlib
    _lib = importlib.import_module("quantummambo_jl_b200._lib")
    assert sorted(_lib.SIGNATURES) == _declared_functions()


def test_library_is_sm100a_and_has_dmma_and_tma(mb):
    so = os.path.join(PKG, "csrc", "libmambo_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in out
    assert "DMMA.8x8x4" in out          # FP64 tensor-core path of the fused kernel
    assert "UBLKCP" in out              # cp.async.bulk (TMA engine) operand ring
    assert "SYNCS" in out               # mbarrier hand-off


def test_no_cpu_fallback(mb, lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the refusal path is for GPU-less hosts")
    with pytest.raises(mb.MamboError) as ei:
        mb.CSAEngine(np.zeros((3, 3, 3, 3)))
    assert ei.value.code == -2          # MAMBO_ECUDA
    assert "no CPU fallback" in str(ei.value)


def test_argument_validation_without_gpu(lib):
    h = C.c_void_p()
    t = np.zeros(16)
    assert lib.mambo_csa_create(C.byref(h), 0, 0, t.ctypes.data, None, 0, 0) == -1       # N < 1
    assert lib.mambo_csa_create(C.byref(h), 2, 7, t.ctypes.data, None, 0, 0) == -1       # bad flavour
    assert lib.mambo_csa_create_sharded(C.byref(h), 2, 0, t.ctypes.data, None, 0, 0, 3, 2) == -1
    assert b"shard" in lib.mambo_last_error()
    assert lib.mambo_csa_num_params(None) == -1
    lib.mambo_csa_destroy(None)          # must be a no-op


def test_product_path_never_imports_the_oracle():
    offenders = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(import|from)\s+.*oracle", txt, flags=re.M) or "csa_oracle" in txt:
                    offenders.append(f)
    assert not offenders, offenders


def test_structures_round_trip(mb):
    n = 5
    x = np.arange(n * n, dtype=float)
    f = mb.CSA_x_to_F_FRAG(x, n)
    assert np.array_equal(f.x(), x) and f.C.lam.size == 15 and f.U[0].thetas.size == 10
    xs = np.arange(n * n + n, dtype=float)
    g = mb.CSA_SD_x_to_F_FRAG(xs, n)
    assert np.array_equal(g.x(), xs) and g.C.lam1.size == n
    H = mb.F_OP(([0], np.eye(n), np.ones((n, n, n, n))))
    assert H.N == n and H.filled == [False, True, True] and H.Nbods == 2
    assert mb.cartan_SD_num_params(n) == 20 and mb.real_orbital_rotation_num_params(n) == 10


def test_synthetic_inputs_are_deterministic_and_symmetric(mb):
    syn = mb.synthetic
    a, b = syn.dense_sym_tbt(4, 3), syn.dense_sym_tbt(4, 3)
    assert np.array_equal(a, b)
    assert np.array_equal(a, a.transpose(1, 0, 2, 3)) and np.array_equal(a, a.transpose(2, 3, 0, 1))
    import csa_oracle as o
    assert np.array_equal(a, o.dense_sym_tbt(4, 3))
    assert np.array_equal(syn.random_x(4, 1, "CSA_SD"), o.random_x(4, 1, "CSA_SD"))


def test_frag_l1_and_x_block_round_trip(tmp_path):
    """The host-only entries of the C-ABI (no GPU involved): CSA_L1 of a fragment (lcu.jl:139-182) against the oracle, and the
    raw checkpoint block (decompose.jl:58-64 layout: tot_L x alpha, column-major) written / read back bit for bit."""
    import numpy as np
    import csa_oracle as o
    import mambo_b200 as mb
    n = 5
    cL = o.cartan_2b_num_params(n)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(cL + n * (n - 1) // 2)
    assert mb.frag_l1(x, n) == o.CSA_L1(x[:cL], n)
    assert mb.frag_l1(x, n, spin_orb=True) == o.CSA_L1(x[:cL], n, spin_orb=True)
    xsd = np.concatenate([rng.standard_normal(n), x])
    assert mb.frag_l1(xsd, n, "CSA_SD") == o.CSA_L1(x[:cL], n)
    X = rng.standard_normal((len(x), 3))
    path = str(tmp_path / "frags.xblk")
    mb.write_x_block(path, "CSA", X)
    group, Y = mb.read_x_block(path)
    assert group == "CSA" and np.array_equal(X, Y) and Y.flags["F_CONTIGUOUS"]
    raw = open(path, "rb").read()
    assert raw[:8] == b"MAMBOX01" and len(raw) == 8 + 16 + 16 + X.size * 8
    assert np.array_equal(np.frombuffer(raw[40:], dtype="<f8").reshape(X.shape, order="F"), X)   # Julia: reshape(reinterpret(Float64, ...), tot_L, alpha)
    mb.write_x_block(path, "CSA_SD", X[:, :0])                                                      # empty block is valid
    group, Y = mb.read_x_block(path)
    assert group == "CSA_SD" and Y.shape == (len(x), 0)
