"""Loader for the package directory `a-lego-loam_b200/` (its name is not a valid Python identifier).

    import alego_pkg
    alego = alego_pkg.load()        # module object, also importable afterwards as `alego_b200`
"""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_ROOT, "a-lego-loam_b200")


def load():
    if "alego_b200" in sys.modules:
        return sys.modules["alego_b200"]
    spec = importlib.util.spec_from_file_location("alego_b200", os.path.join(_PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["alego_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
