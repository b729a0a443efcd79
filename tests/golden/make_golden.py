"""Generates tests/golden/csa_golden.npz from the oracle's *literal* transcription of the reference
(oracle/csa_oracle.py: cost_literal / gradient_literal follow decompose.jl:96-105, 221-231 and gradient.jl:1-290
loop for loop).  The reference itself cannot be executed in this image (no julia), so these fixtures pin the
restatement against regressions -- they are not outputs of the Julia code ("parity unpinned").

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import csa_oracle as o  # noqa: E402
import csa_oracle_ld as ld  # noqa: E402

CASES = [  # (name, N, flavour, tensor kind, seed)
    ("csa_n3_sym", 3, "CSA", "dense_sym", 0),
    ("csa_n4_pair", 4, "CSA", "pairsym", 1),
    ("csa_n5_general", 5, "CSA", "general", 2),
    ("csa_n6_sym", 6, "CSA", "dense_sym", 3),
    ("sd_n3_sym", 3, "CSA_SD", "dense_sym", 4),
    ("sd_n4_general", 4, "CSA_SD", "general", 5),
    ("sd_n5_pair", 5, "CSA_SD", "pairsym", 6),
    ("sd_n7_sym", 7, "CSA_SD", "dense_sym", 7),
    ("csa_n12_sym", 12, "CSA", "dense_sym", 8),
    ("sd_n10_pair", 10, "CSA_SD", "pairsym", 9),
    ("csa_n9_general", 9, "CSA", "general", 10),
]
KINDS = {"dense_sym": o.dense_sym_tbt, "pairsym": o.pairsym_tbt, "general": o.general_tbt}


def main():
    out = {}
    for name, n, fl, kind, seed in CASES:
        T = KINDS[kind](n, seed)
        obt = None
        if fl == "CSA_SD":
            obt = np.random.default_rng(100 + seed).standard_normal((n, n)) / n
        x = o.random_x(n, seed, fl)
        cost = o.cost_literal(x, T, obt, fl)
        grad = o.gradient_literal(x, T, obt, fl)
        h, W = o.to_OP(x, n, fl)
        out[f"{name}/tbt"] = T
        out[f"{name}/x"] = x
        out[f"{name}/cost"] = np.array(cost)
        out[f"{name}/grad"] = grad
        # extended-precision ground truth (long double factored form, own expm; oracle/csa_oracle_ld.py), rounded to FP64
        c_ld, g_ld = ld.cost_grad_factored_ld(x, T, obt, fl)
        out[f"{name}/cost_ld"] = np.array(float(c_ld))
        out[f"{name}/grad_ld"] = g_ld.astype(np.float64)
        out[f"{name}/W"] = W
        if obt is not None:
            out[f"{name}/obt"] = obt
            out[f"{name}/h"] = h
        out[f"{name}/meta"] = np.array([n, 0 if fl == "CSA" else 1])
        print(name, "cost", cost, "|grad|inf", np.abs(grad).max())
    np.savez_compressed(os.path.join(HERE, "csa_golden.npz"), **out)


if __name__ == "__main__":
    main()
