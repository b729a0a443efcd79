"""quantummambo.jl_b200 -- B200-native engine for QuantumMAMBO.jl's CSA / CSA_SD fragment cost + gradient.

The directory name carries a dot (it mirrors the reference repository's name), so it is imported through
`mambo_b200.py` at the repository root, which registers it as the module `quantummambo_jl_b200`.
"""
from ._lib import build, load, MamboError  # noqa: F401
from .engine import (CSAEngine, multistart, cost_grad_batch, frag_l1, write_x_block, read_x_block, CSA, CSA_SD, SYM_AUTO,  # noqa: F401
                     SYM_GENERAL, SYM_PAIR, SYM_FULL)
from .structures import (F_OP, F_FRAG, cartan_2b, cartan_SD, real_orbital_rotation, CSA_x_to_F_FRAG,  # noqa: F401
                         CSA_SD_x_to_F_FRAG, CSA_L1, DF_L1, ob_correction, one_body_L1, cartan_2b_num_params, cartan_SD_num_params,
                         real_orbital_rotation_num_params)
from .decompose import (CSA_greedy_decomposition, CSA_greedy_step, CSA_SD_greedy_decomposition,  # noqa: F401
                        CSA_SD_greedy_step, to_OP, OptimResult)
from . import synthetic  # noqa: F401
from .guesses import tbt_svd_1st, SVD_to_CSA, SOn_unitary_to_params, DF_based_greedy  # noqa: F401
