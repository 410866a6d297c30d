"""TEST INFRASTRUCTURE — import shim that lets the UNMODIFIED reference BuildingEnv run here.

The reference (``/root/reference``, chrisyeh96/sustaingym @ eac2a4d) needs ``gymnasium``,
``pettingzoo`` and ``pvlib`` which are not installed in this image and cannot be installed
(no network).  BuildingEnv only touches a sliver of those packages, so this
module registers ~40 lines of stand-ins in ``sys.modules`` and then imports the reference's
own ``sustaingym.envs.building`` code from where it lies.  Nothing from the reference is
copied into this repo; this file only *imports* it.

It is used by ``tests/golden/gen_building_golden.py`` (run in the build container, where
``/root/reference`` exists) to produce the committed fixtures under ``tests/golden/``.
It is never imported by the product package, and it is useless on the GPU box (no
``/root/reference`` there) — the fixtures travel instead.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("SUSTAINGYM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sustaingym"))


class _Space:
    def __class_getitem__(cls, item):
        return cls


class _Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()


class _Generic(_Space):
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs


class _Env:
    """gymnasium.Env stand-in: only ``reset(seed=)`` seeding (gymnasium/core.py semantics)."""

    def __class_getitem__(cls, item):
        return cls

    def reset(self, *, seed=None, options=None):
        if seed is not None or not hasattr(self, "np_random"):
            self.np_random = np.random.default_rng(seed)


class _ParallelEnv:
    def __class_getitem__(cls, item):
        return cls


def _parse_epw(f):
    """pvlib.iotools.parse_epw stand-in: EPW data rows start after 8 header lines;
    column 6 = dry-bulb temperature (temp_air), column 13 = global horizontal irradiance."""
    import pandas as pd

    lines = f.read().splitlines() if hasattr(f, "read") else open(f).read().splitlines()
    rows = [ln.split(",") for ln in lines[8:] if ln.strip()]
    df = pd.DataFrame({
        "temp_air": [float(r[6]) for r in rows],
        "ghi": [float(r[13]) for r in rows],
    })
    return df, {}


def install() -> None:
    """Registers the stand-ins and puts the reference on ``sys.path``."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")
        spaces = types.ModuleType("gymnasium.spaces")
        spaces.Box = _Box
        spaces.Space = _Space
        spaces.MultiDiscrete = type("MultiDiscrete", (_Generic,), {})
        spaces.Discrete = type("Discrete", (_Generic,), {})
        spaces.Dict = type("Dict", (_Generic,), {})
        gym.spaces = spaces
        gym.Env = _Env
        gym.ActionWrapper = type("ActionWrapper", (_Env,), {})
        core = types.ModuleType("gymnasium.core")
        core.ObsType = object
        envs = types.ModuleType("gymnasium.envs")
        reg = types.ModuleType("gymnasium.envs.registration")
        reg.register = lambda *a, **k: None
        envs.registration = reg
        gym.core, gym.envs = core, envs
        sys.modules.update({
            "gymnasium": gym, "gymnasium.spaces": spaces, "gymnasium.core": core,
            "gymnasium.envs": envs, "gymnasium.envs.registration": reg,
        })
    if "pettingzoo" not in sys.modules:
        pz = types.ModuleType("pettingzoo")
        pz.ParallelEnv = _ParallelEnv
        pzu = types.ModuleType("pettingzoo.utils")
        pzue = types.ModuleType("pettingzoo.utils.env")
        pzue.ParallelEnv = _ParallelEnv
        pzu.env = pzue
        pz.utils = pzu
        sys.modules.update({"pettingzoo": pz, "pettingzoo.utils": pzu, "pettingzoo.utils.env": pzue})
    if "pvlib" not in sys.modules:
        pv = types.ModuleType("pvlib")
        io_ = types.ModuleType("pvlib.iotools")
        io_.parse_epw = _parse_epw
        io_.read_epw = lambda p: _parse_epw(open(p))
        pv.iotools = io_
        sys.modules.update({"pvlib": pv, "pvlib.iotools": io_})
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def reference_building(building: str = "OfficeSmall", weather: str = "Hot_Dry",
                       location: str = "Tucson", **kwargs):
    """Returns ``(BuildingEnv instance, parameters dict)`` built by the reference's own code."""
    install()
    import contextlib
    import io

    from sustaingym.envs.building.env import BuildingEnv
    from sustaingym.envs.building.utils import ParameterGenerator

    with contextlib.redirect_stdout(io.StringIO()):
        params = ParameterGenerator(building=building, weather=weather, location=location, **kwargs)
    return BuildingEnv(params), params
