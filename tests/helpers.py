"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

from conftest import GOLDEN


def building_fixtures():
    return sorted(glob.glob(os.path.join(GOLDEN, "building_*.npz")))


def load_building_fixture(path):
    g = np.load(path)
    params = dict(
        n=int(g["n"]), A_d=g["A_d"], BD_d=g["BD_d"], A=g["A"], B=g["B"], D=g["D"], target=g["target"],
        ac_map=g["ac_map"], out_temp=g["out_temp"], ground_temp=g["ground_temp"], ghi=g["ghi"],
        metabolism=g["metabolism"], reward_beta=float(g["error_rate"]), reward_pnorm=float(g["pnorm"]),
        episode_len=int(g["episode_len"]), time_resolution=int(g["time_res"]), max_power=float(g["max_power"]),
        is_continuous_action=bool(g["is_continuous_action"]), temp_range=(-40, 40), zones=None)
    return g, params


def oracle_building_params(g):
    from oracle.building import BuildingParams
    return BuildingParams(
        n=int(g["n"]), A_d=g["A_d"], BD_d=g["BD_d"], target=g["target"], ac_map=g["ac_map"],
        out_temp=g["out_temp"], ground_temp=g["ground_temp"], ghi=g["ghi"], metabolism=g["metabolism"],
        q_rate=float(g["q_rate"]), error_rate=float(g["error_rate"]), pnorm=float(g["pnorm"]),
        episode_len=int(g["episode_len"]), time_res=int(g["time_res"]), max_power=float(g["max_power"]),
        is_continuous_action=bool(g["is_continuous_action"]))


def rel_err(a, b, floor=1e-3):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)) if a.size else 0.0


def tie_distance(net, x):
    """Distance (in A) of the oracle's pre-rounding pilots to the nearest rounding discontinuity
    of env.py:373-378 — used to exclude genuine ties from bit-exactness checks of the schedule."""
    amps = np.asarray(x, np.float64) * 32
    is_av = (net.min_pilot == 6)[None, :]
    frac_av = np.abs(amps - np.floor(amps) - 0.5)
    d_av = np.where(amps >= 6, frac_av, np.abs(amps - 6.0))
    d_av = np.minimum(d_av, np.abs(amps - 6.0))
    y = amps / 8
    d_cc = np.abs(y - np.floor(y) - 0.5) * 8
    return np.where(is_av, d_av, d_cc)
