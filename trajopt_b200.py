"""Import shim: the package directory name mandated for this repo contains characters Python cannot
import directly, so `import trajopt_b200` loads it from its path under this module name."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "trajectoryoptimization.jl-c79d492b-0548-5874-b488-5a62c1d9d0ca_b200")
_spec = importlib.util.spec_from_file_location("trajopt_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["trajopt_b200"] = _mod
_spec.loader.exec_module(_mod)
