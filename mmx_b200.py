"""Import shim: the product package lives in the directory ``transformer-mm-explainability_b200/`` (the name the
project brief mandates, not a valid Python identifier); this module registers it as package ``mmx_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "transformer-mm-explainability_b200")
_spec = importlib.util.spec_from_file_location("mmx_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mmx_b200"] = _mod
_spec.loader.exec_module(_mod)
