"""Import shim: the package directory name required by the project layout
(`probabilistic-self-update-line-vector-set-based-point-cloud-registration_b200/`) is not a valid
Python identifier, so it is loaded by path and published as `psulvsb_b200`."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "probabilistic-self-update-line-vector-set-based-point-cloud-registration_b200")
_spec = importlib.util.spec_from_file_location("psulvsb_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["psulvsb_b200"] = _mod
_spec.loader.exec_module(_mod)
