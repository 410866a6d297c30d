"""Trace generators for the batched EVChargingEnv.

Mirrors the role of ``sustaingym/envs/evcharging/event_generation.py``: an object the env asks,
at ``reset()``, for the day's charging sessions and MOER table.  The reference builds a fresh
acnsim ``EventQueue`` of Python objects per reset (:149-207); here a generator owns a *pool* of
day traces and MOER tables already packed for the device (``sustaingym_b200.data.ev_traces``), and
a reset only selects indices into the pool.

Day selection follows the reference:
  * sequential generators: ``day = seed % num_days`` and each further reset advances by one,
    cycling (event_generation.py:273-291);
  * the MOER table the env sees belongs to the day *after* the events' day, because
    ``get_event_queue()`` advances ``self.day`` before ``get_moer()`` is called
    (event_generation.py:206, env.py:321-323).  ``moer_day_offset = 1`` reproduces that.
"""
from __future__ import annotations

import numpy as np

from ...data.ev_traces import TraceTables, pack_sessions, synthetic_days, synthetic_moer
from .network import ChargingNetworkSpec, site_str_to_site


class AbstractTraceGenerator:
    """Pool of day traces + MOER tables. Subclasses fill ``tables`` and ``moer``."""

    TIME_STEP_DURATION = 5      # event_generation.py:56
    MAX_STEPS_OF_TRACE = 288    # event_generation.py:58

    def __init__(self, site: str | ChargingNetworkSpec = "caltech", requested_energy_cap: float = 100,
                 sequential: bool = True, moer_day_offset: int = 1, seed: int | None = None):
        self.network = site_str_to_site(site) if isinstance(site, str) else site
        self.site = self.network.site
        self.station_ids = list(self.network.station_ids)
        self.num_stations = len(self.station_ids)
        self.requested_energy_cap = requested_energy_cap
        self.sequential = sequential
        self.moer_day_offset = moer_day_offset
        self.rng = np.random.default_rng(seed)
        self.tables: TraceTables | None = None
        self.moer: np.ndarray | None = None

    @property
    def num_days(self) -> int:
        return self.tables.num_days

    def set_seed(self, seed: int | None) -> None:
        self.rng = np.random.default_rng(seed)

    def first_days(self, seeds: np.ndarray | None, num_envs: int) -> np.ndarray:
        """Trace index of the episode that starts at this reset, for each env."""
        D = self.num_days
        if self.sequential:
            s = np.zeros(num_envs, np.int64) if seeds is None else np.asarray(seeds, np.int64)
            return (s % D).astype(np.int32)
        if seeds is None:
            return self.rng.integers(0, D, num_envs).astype(np.int32)
        return np.array([np.random.default_rng(int(s)).integers(0, D) for s in np.asarray(seeds)], np.int32)

    def moer_days(self, days: np.ndarray) -> np.ndarray:
        return ((np.asarray(days, np.int64) + self.moer_day_offset) % self.moer.shape[0]).astype(np.int32)

    def __repr__(self) -> str:
        return f"{type(self).__name__}(site={self.site!r}, days={self.num_days})"


class ArrayTraceGenerator(AbstractTraceGenerator):
    """Generator over explicitly provided days: ``days`` is a list of per-day lists of
    ``(arrival, departure, estimated_departure, requested_kwh, station_index)`` tuples — the columns
    of the reference's events DataFrame (event_generation.py:125-147); ``moer`` is ``[Dm, 289, 37]``."""

    def __init__(self, days: list, moer: np.ndarray, site="caltech", requested_energy_cap: float = 100,
                 max_sessions: int | None = None, **kw):
        super().__init__(site, requested_energy_cap, **kw)
        self.days = days
        self.tables = pack_sessions(days, self.num_stations, max_sessions, requested_energy_cap)
        self.moer = np.ascontiguousarray(moer, dtype=np.float32)
        if self.moer.ndim != 3 or self.moer.shape[1:] != (289, 37):
            raise ValueError(f"moer must have shape [days, 289, 37], got {self.moer.shape}")


class SyntheticTraceGenerator(ArrayTraceGenerator):
    """Synthetic pool with the marginals of the packaged Caltech data."""

    def __init__(self, site="caltech", num_days: int = 64, mean_sessions: float = 25.0, seed: int = 0,
                 requested_energy_cap: float = 100, **kw):
        net = site_str_to_site(site) if isinstance(site, str) else site
        days = synthetic_days(num_days, net.num_stations, mean_sessions, seed)
        moer = synthetic_moer(num_days, seed + 1)
        super().__init__(days, moer, net, requested_energy_cap, seed=seed, **kw)
