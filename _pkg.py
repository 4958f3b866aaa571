"""Loads the product package `alphazero.jl_b200/` (the directory name has a dot, so plain `import` cannot)."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_NAME = "alphazero_jl_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    d = os.path.join(_ROOT, "alphazero.jl_b200")
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


def build(force=False, verbose=False):
    d = os.path.join(_ROOT, "alphazero.jl_b200")
    spec = importlib.util.spec_from_file_location(_NAME + "_build", os.path.join(d, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force, verbose=verbose)
