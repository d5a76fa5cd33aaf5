"""Import alias: the package directory is ``tc-resnet_b200/`` (not a valid Python identifier), so
``import tcresnet_b200`` loads it under this name."""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tc-resnet_b200")
_spec = importlib.util.spec_from_file_location("tcresnet_b200", os.path.join(_root, "__init__.py"),
                                               submodule_search_locations=[_root])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["tcresnet_b200"] = _mod
_spec.loader.exec_module(_mod)
