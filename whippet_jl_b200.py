"""Import shim: the package directory is `whippet.jl_b200/` (a dotted name Python cannot
import directly); `import whippet_jl_b200` loads it under this module name."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "whippet.jl_b200")
_spec = _u.spec_from_file_location("whippet_jl_b200", _os.path.join(_dir, "__init__.py"),
                                   submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["whippet_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
