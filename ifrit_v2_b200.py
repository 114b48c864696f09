"""Import shim: the package directory is named ``ifrit-v2_b200`` (a hyphen, as the
repository layout prescribes), which Python cannot import by name.  Importing
``ifrit_v2_b200`` loads that directory as the package ``ifrit_v2_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ifrit-v2_b200")
_spec = importlib.util.spec_from_file_location(
    "ifrit_v2_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ifrit_v2_b200"] = _mod
_spec.loader.exec_module(_mod)
