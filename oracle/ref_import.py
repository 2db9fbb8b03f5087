"""Import the UNMODIFIED reference (NVlabs/DiffRL) in this container, as the parity oracle.

TEST INFRASTRUCTURE ONLY.  Nothing under ``diffrl_b200/`` may import this module; only
``tests/``, ``oracle/make_golden.py`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may.  ``/root/reference`` exists only in the build container, never on the GPU box.

The reference sources are *not* copied anywhere.  They are imported from where they lie
(``$DIFFRL_REFERENCE``, default ``/root/reference``) through a meta-path hook that applies, in
memory, the three non-numerical fixes needed on Python 3.12 / a read-only tree:

1. ``dflex/dflex/adjoint.py:1108-1115``  ``node.slice.value`` -> ``node.slice``
   (``ast.Index`` no longer wraps subscripts since Python 3.9);
2. ``dflex/dflex/adjoint.py:1814``  the JIT build directory (``<package>/kernels``) is redirected to
   ``oracle/_ref/kernels`` because the reference tree is read-only -- the compiled
   ``kernels*.so`` (the reference's own generated CPU kernels, g++ -O2) is the only thing
   that lands in the repo tree, and ``oracle/_ref/`` is git-ignored;
3. stand-ins for modules missing from the image: ``imp`` (stdlib, removed in 3.12), ``gym.spaces``,
   ``urdfpy``, ``tensorboardX`` (``oracle/refshim``), and ``numpy.Inf`` (removed in NumPy 2,
   used by ``envs/dflex_env.py:48``).

None of these touches the arithmetic of the hot path (``dflex/dflex/sim.py:1076-2601`` and the
``*.h`` headers), which is compiled exactly as shipped.
"""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("DIFFRL_REFERENCE", "/root/reference")
REF_BUILD_DIR = os.path.join(_HERE, "_ref", "kernels")
_SHIM_DIR = os.path.join(_HERE, "refshim")

_PATCHES = {
    "adjoint.py": [
        ("node.slice.value", "node.slice"),
        (
            'build_path = os.path.dirname(os.path.realpath(__file__)) + "/kernels"',
            "build_path = os.environ['DFLEX_REF_BUILD_PATH']",
        ),
    ],
}


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "dflex", "dflex", "sim.py"))


class _PatchedLoader(importlib.machinery.SourceFileLoader):
    """SourceFileLoader that rewrites a few literal substrings of the file it loads."""

    def get_data(self, path):
        data = super().get_data(path)
        edits = _PATCHES.get(os.path.basename(path))
        if edits and path.endswith(".py"):
            text = data.decode("utf-8")
            for old, new in edits:
                if old not in text:
                    raise ImportError("oracle patch target %r not found in %s" % (old, path))
                text = text.replace(old, new)
            data = text.encode("utf-8")
        return data

    def get_code(self, fullname):
        # never trust / write byte-code caches for patched sources
        source = self.get_data(self.get_filename(fullname))
        return self.source_to_code(source, self.get_filename(fullname))


class _ReferenceDflexFinder(importlib.abc.MetaPathFinder):
    """Resolves ``dflex`` and ``dflex.<sub>`` to ``$DIFFRL_REFERENCE/dflex/dflex``."""

    def __init__(self, package_dir):
        self.package_dir = package_dir

    def find_spec(self, fullname, path=None, target=None):
        if fullname == "dflex":
            init = os.path.join(self.package_dir, "__init__.py")
            return importlib.util.spec_from_file_location(
                fullname, init, loader=_PatchedLoader(fullname, init),
                submodule_search_locations=[self.package_dir])
        if fullname.startswith("dflex."):
            leaf = fullname.split(".", 1)[1]
            if "." in leaf:
                return None
            filename = os.path.join(self.package_dir, leaf + ".py")
            if os.path.isfile(filename):
                return importlib.util.spec_from_file_location(
                    fullname, filename, loader=_PatchedLoader(fullname, filename))
        return None


_installed = False


def install():
    """Make ``import dflex``, ``import envs``, ``import utils`` ... resolve to the reference."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise ImportError("reference tree not found at %s (it only exists in the build container)"
                          % REFERENCE_ROOT)
    if "dflex" in sys.modules and not getattr(sys.modules["dflex"], "__file__", "").startswith(REFERENCE_ROOT):
        raise ImportError("a non-reference 'dflex' is already imported in this process; "
                          "run the oracle in its own interpreter")
    import numpy as np
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    os.makedirs(REF_BUILD_DIR, exist_ok=True)
    os.environ["DFLEX_REF_BUILD_PATH"] = REF_BUILD_DIR
    # the reference keys its CUDA code path off torch.cuda.is_available(); its CUDA flags
    # (compute_35) do not build with nvcc 12.9, so the oracle is always the CPU path.
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
    sys.meta_path.insert(0, _ReferenceDflexFinder(os.path.join(REFERENCE_ROOT, "dflex", "dflex")))
    sys.path.insert(0, _SHIM_DIR)
    sys.path.insert(1, REFERENCE_ROOT)
    _installed = True


def make_env(name, num_envs, **overrides):
    """Instantiate a reference env on the CPU path with the benchmark's deterministic settings
    (SURVEY.md section 8d): stochastic_init=False, no_grad=False, seed=0, YAML MM caching."""
    install()
    import envs  # noqa: the reference package
    mm_freq = {"AntEnv": 16, "HumanoidEnv": 48, "SNUHumanoidEnv": 8, "CartPoleSwingUpEnv": 4,
               "HopperEnv": 16, "CheetahEnv": 16}[name]
    kwargs = dict(num_envs=num_envs, device="cpu", render=False, seed=0, stochastic_init=False,
                  no_grad=False, MM_caching_frequency=mm_freq)
    kwargs.update(overrides)
    return getattr(envs, name)(**kwargs)
