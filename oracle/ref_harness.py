"""Loader for the UNMODIFIED reference modules (container-only; test infrastructure).

``/root/reference`` exists only in the build container, never on the GPU box, so
everything here is used by (a) the CPU tests that pin ``oracle/`` against the real
reference and (b) ``oracle/make_golden.py`` which writes the committed fixtures in
``tests/golden/``.  Nothing in the product, the ``-m gpu`` tests, ``smoke()`` or
``bench.py`` imports this file.

The reference has no package: each ``PPOV*/`` folder holds flat scripts that do
``from config import ...`` / ``import gym`` / ``from netCDF4 import Dataset``.  We load
them by path under private module names, with

* a stub ``gym`` (``Env``, ``spaces.Discrete``, ``spaces.Box``) -- the reference only
  uses the base class and the two space constructors (PPOV2.1/environment.py:19-29),
* a stub ``netCDF4`` (``Dataset``) -- imported by PPOV2.1/model.py:6 but never used on
  the hot path,
* an ``np`` proxy installed into the loaded ``environment`` module whose
  ``.random.rand/.randn`` replay an injected stream (the reference is unseeded and
  draws from the global MT19937: environment.py:44,58,60,108).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("PLUME_REFERENCE_ROOT", "/root/reference")

_TRAIN_SCRIPT = {"1.1": "train_ppo1.1.py", "2.0": "train_ppo2.0.py", "2.1": "train_ppo2.0.py"}


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "PPOV2.1"))


def _stub_gym() -> dict:
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Env:  # noqa: D401 - stub
        pass

    class Discrete:
        def __init__(self, n):
            self.n = n

    class Box:
        def __init__(self, low, high, dtype=None, shape=None):
            self.low, self.high, self.dtype = low, high, dtype

    gym.Env = Env
    spaces.Discrete = Discrete
    spaces.Box = Box
    gym.spaces = spaces
    return {"gym": gym, "gym.spaces": spaces}


def _stub_netcdf() -> dict:
    nc = types.ModuleType("netCDF4")

    class Dataset:  # pragma: no cover - never called on the hot path
        def __init__(self, *a, **k):
            raise RuntimeError("netCDF4 is stubbed in the oracle harness")

    nc.Dataset = Dataset
    return {"netCDF4": nc}


class InjectedRandom:
    """Replays injected draws in the order the reference consumes them.

    ``feed`` is any object with ``rand(*shape)`` and ``randn(*shape)``.
    """

    def __init__(self, feed):
        self.feed = feed

    def rand(self, *shape):
        return self.feed.rand(*shape)

    def randn(self, *shape):
        return self.feed.randn(*shape)


class NpProxy:
    """Forwards everything to numpy except ``.random``."""

    def __init__(self, random_obj):
        self.random = random_obj

    def __getattr__(self, name):
        return getattr(np, name)


class QueueFeed:
    """A feed that pops pre-computed arrays: ``rand2`` (source), ``randn_field``,
    ``rand_field`` per reset, ``randn2`` per step."""

    def __init__(self):
        self.rand_q = []
        self.randn_q = []

    def push_reset(self, u_src, z_field, u_field):
        self.rand_q.append(np.asarray(u_src, dtype=np.float64))
        self.randn_q.append(np.asarray(z_field, dtype=np.float64))
        self.rand_q.append(np.asarray(u_field, dtype=np.float64))

    def push_step(self, z2):
        self.randn_q.append(np.asarray(z2, dtype=np.float64))

    def rand(self, *shape):
        a = self.rand_q.pop(0)
        assert a.shape == tuple(shape), (a.shape, shape)
        return a

    def randn(self, *shape):
        a = self.randn_q.pop(0)
        assert a.shape == tuple(shape), (a.shape, shape)
        return a


_CACHE: dict = {}


def load_reference(version: str = "2.1") -> types.SimpleNamespace:
    """Returns a namespace with the reference's ``config``, ``environment``, ``model``
    and training-script modules of ``PPOV<version>`` loaded unmodified."""
    if version in _CACHE:
        return _CACHE[version]
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    root = os.path.join(REFERENCE_ROOT, f"PPOV{version}")
    shadow = ["config", "environment", "model", "netcdf_writer", "gym", "gym.spaces", "netCDF4"]
    saved = {k: sys.modules.get(k) for k in shadow}
    for k, v in {**_stub_gym(), **_stub_netcdf()}.items():
        sys.modules[k] = v
    loaded = {}

    def _load(fname: str, public: str):
        spec = importlib.util.spec_from_file_location(f"_plume_ref_{version.replace('.', '_')}_{public}",
                                                      os.path.join(root, fname))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[public] = mod  # so that `from config import ...` resolves to this version
        spec.loader.exec_module(mod)
        loaded[public] = mod
        return mod

    try:
        _load("config.py", "config")
        _load("environment.py", "environment")
        _load("model.py", "model")
        if version == "2.0":
            _load("netcdf_writer.py", "netcdf_writer")
        train = _load(_TRAIN_SCRIPT[version], "train")
        extras = {}
        for fname, key in (("evaluate_with_lstm.py", "evaluate_with_lstm"), ("evaluate_model.py", "evaluate_model"),
                           ("train_lstm.py", "train_lstm")):
            if os.path.exists(os.path.join(root, fname)):
                try:
                    if key == "evaluate_with_lstm" and version == "2.1":
                        sys.modules["check_gaussian"] = types.ModuleType("check_gaussian")
                    extras[key] = _load(fname, key)
                except Exception as exc:  # pragma: no cover - optional pieces
                    extras[key] = exc
    finally:
        for k in shadow + ["train", "evaluate_with_lstm", "evaluate_model", "check_gaussian", "train_lstm"]:
            if saved.get(k) is not None:
                sys.modules[k] = saved[k]
            else:
                sys.modules.pop(k, None)
    ns = types.SimpleNamespace(version=version, config=loaded["config"], environment=loaded["environment"],
                               model=loaded["model"], train=train, **extras)
    _CACHE[version] = ns
    return ns


def make_reference_env(version: str, feed) -> object:
    """Constructs the reference ``MethaneEnv`` with its RNG replaced by ``feed``.  The
    constructor itself calls ``reset()`` (environment.py:40) and therefore consumes one
    reset worth of draws."""
    ref = load_reference(version)
    ref.environment.np = NpProxy(InjectedRandom(feed))
    try:
        env = ref.environment.MethaneEnv()
    finally:
        pass  # the proxy stays installed: later reset()/step() calls keep drawing from `feed`
    return env


class FakeNC:
    """Stands in for ``netCDF4.Dataset`` (absent in this image) over a dict of numpy arrays with the
    training_data.nc variable names: supports ``with``, ``nc[name]`` and ``name in nc.variables``, which is all
    ``load_trajectory_segments`` (PPOV2.1/model.py:67-91) uses."""

    def __init__(self, arrays: dict):
        self.variables = dict(arrays)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __getitem__(self, name):
        return self.variables[name]


class reference_modules:
    """Context manager: while active, ``import config`` / ``import model`` inside reference code (e.g. the
    function-local ``from config import GRID_SIZE`` of train_lstm.py:13) resolve to the loaded reference modules."""

    def __init__(self, version: str = "2.1"):
        self.ref = load_reference(version)
        self.saved = {}

    def __enter__(self):
        for k, m in (("config", self.ref.config), ("model", self.ref.model)):
            self.saved[k] = sys.modules.get(k)
            sys.modules[k] = m
        return self.ref

    def __exit__(self, *exc):
        for k, v in self.saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
        return False


class _MemVariable:
    """One variable of ``MemDataset``: a numpy array with item assignment and free attributes."""

    def __init__(self, dtype, shape, fill_value):
        self.data = np.full(shape, fill_value if fill_value is not None else 0, dtype=dtype)

    def __setitem__(self, key, value):
        self.data[key] = value

    def __getitem__(self, key):
        return self.data[key]


class MemDataset:
    """Writable in-memory stand-in for ``netCDF4.Dataset`` (absent in this image): exactly what
    ``NetCDFWriter`` (PPOV2.1/model.py:351-422) calls -- createDimension, createVariable(name, dtype, dims,
    fill_value=, zlib=), attribute assignment, close."""

    def __init__(self, filename=None, mode="w", format=None):
        self.dimensions, self.variables = {}, {}

    def createDimension(self, name, size):
        self.dimensions[name] = size

    def createVariable(self, name, dtype, dims, fill_value=None, zlib=False):
        var = _MemVariable(dtype, tuple(self.dimensions[d] for d in dims), fill_value)
        self.variables[name] = var
        return var

    def close(self):
        pass


def run_reference_lines(relpath: str, first: str, last: str, namespace: dict) -> dict:
    """Executes a block of an UNMODIFIED reference script in ``namespace``: the lines from the first one containing
    ``first`` to the first later one containing ``last`` (inclusive), dedented.  The reference's older drivers keep
    their GAE loops inline in ``train_ppo()`` (PPOV1.1/train_ppo1.0.py:72-89, PPOV1.0/ppo0.0.py:337-353,
    PPOV1.2's driver :369-382), so running them means running those source lines; the text is read from
    ``/root/reference`` at test time and never stored in this repository."""
    import textwrap
    path = os.path.join(REFERENCE_ROOT, relpath)
    with open(path, encoding="utf-8") as f:
        lines = f.read().splitlines()
    start = next(i for i, l in enumerate(lines) if first in l)
    end = next(i for i in range(start, len(lines)) if last in lines[i])
    exec(compile(textwrap.dedent("\n".join(lines[start:end + 1])), path, "exec"), namespace)
    return namespace
