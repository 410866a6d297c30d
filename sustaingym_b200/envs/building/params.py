"""``parameters`` dicts of BuildingEnv on disk.

The reference builds the dict with ``ParameterGenerator`` (sustaingym/envs/building/utils.py:462-779) from
EnergyPlus HTM tables and EPW weather files at construction time.  The engine only needs its numeric
content — the RC matrices, per-zone vectors and the four 5-minute weather tables — so an offline export
(``tests/golden/export_building_params.py``, which runs the reference's generator) is enough to run any shipped
building / climate without the reference tree.
"""
from __future__ import annotations

from typing import Any

import numpy as np

_ARRAYS = ("A", "B", "D", "target", "ac_map", "max_power", "out_temp", "ground_temp", "ghi", "metabolism")
_SCALARS = ("n", "time_resolution", "reward_beta", "reward_pnorm", "episode_len", "is_continuous_action")


def save_parameters(path: str, p: dict[str, Any]) -> None:
    out = {k: np.asarray(p[k]) for k in _ARRAYS if k in p}
    for k in ("out_temp", "ground_temp", "ghi", "metabolism"):
        out[k] = out[k].astype(np.float32) if np.array_equal(out[k].astype(np.float32).astype(np.float64), out[k]) else out[k]
    for k in _SCALARS:
        out[k] = np.asarray(p[k])
    out["temp_range"] = np.asarray(p.get("temp_range", (-40, 40)), np.float64)
    np.savez_compressed(path, **out)


def load_parameters(path: str) -> dict[str, Any]:
    """The dict ``BuildingEnv(parameters)`` / ``BatchedBuildingEnv(parameters, num_envs)`` takes."""
    with np.load(path) as z:
        p: dict[str, Any] = {k: z[k].astype(np.float64) for k in _ARRAYS if k in z}
        p["n"] = int(z["n"])
        p["time_resolution"] = int(z["time_resolution"])
        p["reward_beta"] = float(z["reward_beta"])
        p["reward_pnorm"] = float(z["reward_pnorm"])
        p["episode_len"] = int(z["episode_len"])
        p["is_continuous_action"] = bool(z["is_continuous_action"])
        p["temp_range"] = tuple(float(v) for v in z["temp_range"])
        p["zones"] = None
    if p["max_power"].ndim == 0:
        p["max_power"] = float(p["max_power"])
    return p
