"""quantummambo.jl_b200 -- B200-native engine for QuantumMAMBO.jl's CSA / CSA_SD fragment cost + gradient.

The directory name carries a dot (it mirrors the reference repository's name), so it is imported through
`mambo_b200.py` at the repository root, which registers it as the module `quantummambo_jl_b200`.
"""
from ._lib import build, load, MamboError  # noqa: F401
from .engine import (CSAEngine, CSA, CSA_SD, SYM_AUTO, SYM_GENERAL, SYM_PAIR, SYM_FULL)  # noqa: F401
