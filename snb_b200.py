"""Import shim: registers the directory `segmentation-networks-benchmark_b200/` as the package `snb_b200`.

The hyphens in the reference's name cannot appear in a Python identifier, so `import snb_b200` loads that
directory's __init__.py under this name and replaces this module in sys.modules.
"""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "segmentation-networks-benchmark_b200")
_spec = importlib.util.spec_from_file_location(
    "snb_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["snb_b200"] = _mod
_spec.loader.exec_module(_mod)
