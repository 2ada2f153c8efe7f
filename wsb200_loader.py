"""Import helper for the package directory `rustronomy-watershed_b200/`.

A hyphen is not a valid character in a Python module name, so the package is
loaded from its path and registered as `rustronomy_watershed_b200`.
"""
import importlib.util
import os
import sys

_NAME = "rustronomy_watershed_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rustronomy-watershed_b200")


def load():
    mod = sys.modules.get(_NAME)
    if mod is not None:
        return mod
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                  submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


def build(force: bool = False, verbose: bool = False) -> str:
    spec = importlib.util.spec_from_file_location("_wsb200_build", os.path.join(_DIR, "build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.build(force=force, verbose=verbose)
