"""Imports the package directory `flashweave.jl_b200/` (its name contains a dot, so the normal
import statement cannot reach it) and registers it as module `flashweave_jl_b200`."""
import importlib.util
import os
import sys

_NAME = "flashweave_jl_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "flashweave.jl_b200")


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


def load_sub(name):
    load()
    full = _NAME + "." + name
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, os.path.join(_DIR, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod
