"""Import shim: the package directory is named ``arm-spmv_b200`` (after the reference repo),
which is not a valid Python identifier.  ``import arm_spmv_b200`` lands here and is redirected
to that directory."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "arm-spmv_b200")
_spec = _u.spec_from_file_location("arm_spmv_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["arm_spmv_b200"] = _mod
_spec.loader.exec_module(_mod)
