"""Ambient-condition tables for the batched CogenEnv.

The reference builds ``ambients_dfs``: a list of per-day DataFrames with 96 fifteen-minute rows and
the columns ``Ambient Temperature, Ambient Pressure, Ambient rel. Humidity, Target Net Power, Target
Process Steam, Energy Price, Gas Price`` (``sustaingym/data/cogen/load_ambients.py:38-132``).  The
engine consumes the same content as one dense ``[n_days, 96, 7]`` float64 array (column order =
``AMBIENT_COLS``).  ``from_dataframes`` converts the reference's list; ``synthetic_ambients`` makes a
table inside the observation-space bounds (cogen/env.py:136-142) for tests and benchmarks.
"""
from __future__ import annotations

import os

import numpy as np

AMBIENT_COLS = ("TAMB", "PAMB", "RHAMB", "Target_Power", "Target_Steam", "Energy_Price", "Gas_Price")
REFERENCE_COLUMNS = ("Ambient Temperature", "Ambient Pressure", "Ambient rel. Humidity", "Target Net Power",
                     "Target Process Steam", "Energy Price", "Gas Price")
STEPS_PER_DAY = 96


def from_dataframes(dfs) -> np.ndarray:
    """``load_ambients.construct_df()`` output -> [n_days, 96, 7] float64."""
    return np.stack([np.asarray(df[list(REFERENCE_COLUMNS)], np.float64) for df in dfs])


_REAL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cogen_ambients_real.npz")


def real_ambients(renewables_magnitude: float = 0.0) -> np.ndarray:
    """The reference's own ambient table (253 days from 2021-05 on; built once from its raw xlsx / csv files by
    ``python -m sustaingym_b200.data.load_ambients``, which restates ``load_ambients.construct_df``), for any
    ``renewables_magnitude``: target power = max(raw target - magnitude * wind capacity factor, 0)
    (load_ambients.py:120-126)."""
    with np.load(_REAL) as z:
        tab, cf = z["ambients"].copy(), z["wind_cf"]
    if renewables_magnitude:
        tab[..., 3] = np.maximum(tab[..., 3] - float(renewables_magnitude) * cf, 0.0)
    return tab


def synthetic_ambients(n_days: int = 64, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n_days * STEPS_PER_DAY) / STEPS_PER_DAY
    diurnal = np.sin(2 * np.pi * (t - 0.3))
    tamb = 70 + 18 * diurnal + 8 * np.sin(2 * np.pi * t / 30) + rng.normal(0, 1.0, t.size)
    pamb = 14.5 + 0.2 * np.sin(2 * np.pi * t / 7) + rng.normal(0, 0.02, t.size)
    rh = np.clip(0.55 - 0.25 * diurnal + rng.normal(0, 0.05, t.size), 0.02, 0.98)
    power = np.clip(420 + 120 * np.sin(2 * np.pi * (t - 0.45)) + rng.normal(0, 15, t.size), 50, 690)
    steam = np.clip(900 + 150 * np.sin(2 * np.pi * (t - 0.2)) + rng.normal(0, 20, t.size), 100, 1290)
    price = np.clip(45 + 30 * np.sin(2 * np.pi * (t - 0.55)) + rng.gamma(1.5, 6, t.size), 1, 1400)
    gas = np.clip(4.0 + 0.8 * np.sin(2 * np.pi * t / 60) + rng.normal(0, 0.05, t.size), 0.5, 6.9)
    tab = np.stack([np.clip(tamb, 33, 114), np.clip(pamb, 14.01, 14.99), rh, power, steam, price, gas], axis=1)
    return tab.reshape(n_days, STEPS_PER_DAY, 7)
