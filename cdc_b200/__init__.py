"""Importable alias of the package directory `conditional-diffusion-model-for-compression_b200/`
(its hyphenated name is the layout the build contract asks for but is not a Python identifier)."""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                     "conditional-diffusion-model-for-compression_b200")
_spec = importlib.util.spec_from_file_location("cdc_b200", os.path.join(_root, "__init__.py"),
                                               submodule_search_locations=[_root])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cdc_b200"] = _mod
_spec.loader.exec_module(_mod)
