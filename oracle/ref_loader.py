"""
ORACLE support (test infrastructure): import the UNMODIFIED reference modules from /root/reference when that
tree exists (build container), else from the byte-for-byte copy oracle/make_ref.py placed under oracle/_ref/
(git-ignored; it travels to the GPU box, where /root/reference is absent).

tsu/gibbs.py and tsu/core.py import only numpy + stdlib, so they are loaded by file path.
tsu/models/ising.py imports matplotlib at module scope (ising.py:20-21); matplotlib is not
installed here, so empty stand-in modules are registered in sys.modules first (nothing in the
hot path touches them).

Used by the CPU tests, by `bench.py --impl reference` / the cpu_baseline leg (the reference timed on the host
cores) and by tests/test_reference_suite.py; never by the product package.
"""

import importlib.util
import os
import sys
import types
from contextlib import contextmanager
from unittest import mock

import numpy as np

from .make_ref import ref_root

REFERENCE_ROOT = os.environ.get("TSU_REFERENCE_ROOT") or ref_root() or "/root/reference"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "tsu", "gibbs.py"))


def _load(name, relpath):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def load_reference():
    """returns (gibbs_module, core_module, ising_module) of the reference."""
    if "mods" in _cache:
        return _cache["mods"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    # stand-ins for the plotting imports of ising.py
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        fig = types.ModuleType("matplotlib.figure")
        fig.Figure = type("Figure", (), {})
        mpl.pyplot = plt
        mpl.figure = fig
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        sys.modules["matplotlib.figure"] = fig
    # build a private package "tsu_reference" so that ising.py's relative import works
    pkg = types.ModuleType("tsu_reference")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "tsu")]
    sys.modules["tsu_reference"] = pkg
    gibbs = _load("tsu_reference.gibbs", "tsu/gibbs.py")
    core = _load("tsu_reference.core", "tsu/core.py")
    models = types.ModuleType("tsu_reference.models")
    models.__path__ = [os.path.join(REFERENCE_ROOT, "tsu", "models")]
    sys.modules["tsu_reference.models"] = models
    ising = _load("tsu_reference.models.ising", "tsu/models/ising.py")
    _cache["mods"] = (gibbs, core, ising)
    return _cache["mods"]


@contextmanager
def injected_numpy_random(uniforms=None, order=None, normals=None, randint=None):
    """patch the global numpy stream the reference draws from.

    uniforms: iterable consumed by np.random.rand() (scalar calls, gibbs.py:126,320)
    order:    array returned by np.random.permutation(n) (gibbs.py:157)
    normals:  iterable of arrays consumed by np.random.randn(*shape) (core.py:78,143)
    randint:  array returned by np.random.randint(0, 2, size=n) (gibbs.py:201)
    """
    patches = []
    if uniforms is not None:
        it = iter(uniforms)
        patches.append(mock.patch("numpy.random.rand", lambda *a: next(it)))
    if order is not None:
        patches.append(mock.patch("numpy.random.permutation", lambda n: np.asarray(order)))
    if normals is not None:
        itn = iter(normals)

        def _randn(*shape):
            v = np.asarray(next(itn), dtype=np.float64)
            return v.reshape(shape) if shape else float(v)

        patches.append(mock.patch("numpy.random.randn", _randn))
    if randint is not None:
        patches.append(mock.patch("numpy.random.randint", lambda *a, **k: np.asarray(randint).copy()))
    for p in patches:
        p.start()
    try:
        yield
    finally:
        for p in reversed(patches):
            p.stop()
