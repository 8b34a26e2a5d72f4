"""Import alias for the package directory `linearmixingmodels.jl_b200/` (its name is not a valid
Python identifier).  `import lmm_b200` gives the package; `lmm_b200.api`, `lmm_b200._lib`,
`lmm_b200.dist` are its submodules."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "linearmixingmodels.jl_b200")
_spec = importlib.util.spec_from_file_location("lmm_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["lmm_b200"] = _mod
_spec.loader.exec_module(_mod)
