"""Python >= 3.12 removed the stdlib ``imp`` module, which the reference's
adjoint.py:9,1681-1688 still uses to load its freshly built ``kernels`` extension.
This is a two-function stand-in (test infrastructure; never imported by the product)."""
import importlib.machinery
import importlib.util
import os
import sys


def find_module(name, path):
    for directory in path:
        for suffix in list(importlib.machinery.EXTENSION_SUFFIXES) + [".so"]:
            candidate = os.path.join(directory, name + suffix)
            if os.path.isfile(candidate):
                return open(candidate, "rb"), candidate, None
    raise ImportError("imp shim: no extension module %r under %r" % (name, path))


def load_module(name, file, filename, description):
    spec = importlib.util.spec_from_file_location(name, filename)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    sys.modules[name] = module
    return module
