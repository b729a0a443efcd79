"""Import shim: loads the package directory `quantummambo.jl_b200/` (whose name is not a valid Python
identifier) as the module `quantummambo_jl_b200` and re-exports it."""
import importlib.util
import sys
from pathlib import Path

_NAME = "quantummambo_jl_b200"
_DIR = Path(__file__).resolve().parent / "quantummambo.jl_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, _DIR / "__init__.py", submodule_search_locations=[str(_DIR)])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

pkg = sys.modules[_NAME]
globals().update({k: v for k, v in vars(pkg).items() if not k.startswith("__")})
