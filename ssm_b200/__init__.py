"""Import alias: `import ssm_b200` loads the package that lives in the directory
`superslomo-videointerpolation-pytorch_b200/` (whose name is not a valid Python identifier)."""
import importlib.util
import os
import sys

_real = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                     "superslomo-videointerpolation-pytorch_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_real, "__init__.py"),
                                               submodule_search_locations=[_real])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
