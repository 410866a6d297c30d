"""Day traces and MOER tables in the device-resident layout the EV kernels read.

In the reference a day's sessions are produced at ``reset()`` by ``RealTraceGenerator`` /
``GMMsTraceGenerator`` (``sustaingym/envs/evcharging/event_generation.py:149-207, 293-328,
416-515``) as a DataFrame with columns arrival, departure, estimated_departure,
requested_energy, station_id, then turned into acnsim ``PluginEvent``s.  Here a pool of D days
is packed once into dense per-station slot tables so that ``reset()`` on the GPU is an index
bump:

    arrival / departure / est_departure : int16 [D, K, n]   (slot k of station i on day d)
    energy                              : float64 [D, K, n] (requested kWh after the cap)

Slots of a station are ordered by arrival; unused slots hold arrival = departure = 32767.
Sessions of one station must not overlap (acnsim raises StationOccupiedError otherwise).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

EMPTY = 32767
STEPS_PER_DAY = 288


@dataclass
class TraceTables:
    arrival: np.ndarray        # int16 [D, K, n]
    departure: np.ndarray      # int16 [D, K, n]
    est_departure: np.ndarray  # int16 [D, K, n]
    energy: np.ndarray         # float64 [D, K, n]

    @property
    def num_days(self) -> int:
        return self.arrival.shape[0]

    @property
    def max_sessions(self) -> int:
        return self.arrival.shape[1]

    @property
    def num_stations(self) -> int:
        return self.arrival.shape[2]

    def sessions_of_day(self, d: int) -> list:
        """[(arrival, departure, est_departure, energy, station)] of day d, sorted by arrival."""
        out = []
        K, n = self.max_sessions, self.num_stations
        for i in range(n):
            for k in range(K):
                if self.arrival[d, k, i] != EMPTY:
                    out.append((int(self.arrival[d, k, i]), int(self.departure[d, k, i]),
                                int(self.est_departure[d, k, i]), float(self.energy[d, k, i]), i))
        out.sort(key=lambda s: (s[0], s[4]))
        return out

    def arrival_table(self) -> np.ndarray:
        """[D, Smax + 1] uint32: the day's sessions sorted by (arrival, station), packed as
        ``arrival | station << 16 | slot << 24`` and terminated by 0x7fff entries — the event list the
        event-driven kernel walks instead of scanning all stations (the reference's EventQueue of
        PluginEvents, event_generation.py:188-196)."""
        D, K, n = self.arrival.shape
        rows = []
        for d in range(D):
            k, i = np.nonzero(self.arrival[d] != EMPTY)
            arr = self.arrival[d, k, i].astype(np.int64)
            order = np.lexsort((i, arr))
            rows.append((arr[order] | (i[order].astype(np.int64) << 16) | (k[order].astype(np.int64) << 24)).astype(np.uint32))
        smax = max(1, max(len(r) for r in rows))
        tab = np.full((D, smax + 1), 0x7FFF, np.uint32)
        for d, r in enumerate(rows):
            tab[d, :len(r)] = r
        return tab

    def max_profit(self, marginal_profit_per_kwh: float = 0.03, a_pers_to_kwh: float = 5 / 60 * 0.208,
                   scale: float = 32.0) -> np.ndarray:
        """``EVChargingEnv._calculate_max_profit`` (env.py:422-429) for every day: [D]."""
        valid = self.arrival != EMPTY
        dur = (self.departure.astype(np.int64) - self.arrival.astype(np.int64)) * valid
        kwh = np.minimum(self.energy, dur * scale * a_pers_to_kwh) * valid
        return (kwh * marginal_profit_per_kwh).sum(axis=(1, 2))


def pack_sessions(days: list, num_stations: int, max_sessions: int | None = None,
                  requested_energy_cap: float = 100.0) -> TraceTables:
    """``days``: list of per-day lists of (arrival, departure, est_departure, energy_kwh, station).
    Applies the requested-energy cap (event_generation.py:169-170) and validates occupancy."""
    per = [[[] for _ in range(num_stations)] for _ in days]
    for d, ss in enumerate(days):
        for (arr, dep, est, kwh, stn) in ss:
            if not (0 <= stn < num_stations):
                raise ValueError(f"day {d}: station index {stn} out of range")
            if not (0 <= arr <= dep <= STEPS_PER_DAY - 1):
                raise ValueError(f"day {d}: need 0 <= arrival <= departure <= 287, got {(arr, dep)}")
            per[d][stn].append((int(arr), int(dep), int(est), min(float(kwh), requested_energy_cap)))
    K = max(1, max(len(s) for day in per for s in day))
    if max_sessions is not None:
        if K > max_sessions:
            raise ValueError(f"a station has {K} sessions in one day but max_sessions={max_sessions}")
        K = max_sessions
    if K > 254:
        raise ValueError("more than 254 sessions per station per day (255 is the kernels' empty hint)")
    D = len(days)
    arrival = np.full((D, K, num_stations), EMPTY, np.int16)
    departure = np.full((D, K, num_stations), EMPTY, np.int16)
    est = np.full((D, K, num_stations), EMPTY, np.int16)
    energy = np.zeros((D, K, num_stations), np.float64)
    for d in range(D):
        for i in range(num_stations):
            ss = sorted(per[d][i])
            for k, (arr, dep, e, kwh) in enumerate(ss):
                if k > 0 and arr < ss[k - 1][1]:
                    raise ValueError(f"day {d} station {i}: sessions overlap (StationOccupiedError in acnsim)")
                arrival[d, k, i], departure[d, k, i], est[d, k, i], energy[d, k, i] = arr, dep, e, kwh
    return TraceTables(arrival, departure, est, energy)


def synthetic_days(num_days: int, num_stations: int = 54, mean_sessions: float = 25.0, seed: int = 0,
                   station_weights: np.ndarray | None = None) -> list:
    """Synthetic day traces with the marginals of the packaged Caltech data:
    arrivals U{0..240}, durations U{12..120} steps, estimated departure = departure + U{-12..12}
    clipped to (arrival, 287], requested energy clip(LogNormal(2.4, 0.8), 0.5, 100) kWh, station
    drawn among those free at arrival (the GMM generator's rule, event_generation.py:493-511)."""
    rng = np.random.default_rng(seed)
    w = np.ones(num_stations) if station_weights is None else np.asarray(station_weights, float)
    days = []
    for _ in range(num_days):
        m = int(rng.poisson(mean_sessions))
        arr = np.sort(rng.integers(0, 241, m))
        dur = rng.integers(12, 121, m)
        dep = np.minimum(arr + dur, STEPS_PER_DAY - 1)
        est = np.clip(dep + rng.integers(-12, 13, m), arr + 1, STEPS_PER_DAY - 1)
        kwh = np.clip(rng.lognormal(2.4, 0.8, m), 0.5, 100.0)
        free_at = np.full(num_stations, -1)
        ss = []
        for j in range(m):
            avail = np.nonzero(free_at < arr[j])[0]
            if len(avail) == 0:
                continue
            p = w[avail] / w[avail].sum()
            i = int(rng.choice(avail, p=p))
            free_at[i] = dep[j]
            ss.append((int(arr[j]), int(dep[j]), int(est[j]), float(kwh[j]), i))
        days.append(ss)
    return days


def synthetic_moer(num_days: int, seed: int = 0) -> np.ndarray:
    """[Dm, 289, 37] float32 tables shaped like ``MOERLoader.retrieve`` output
    (sustaingym/data/load_moer.py:364-377): column 0 the realised marginal emission rate
    (kg CO2/kWh, a smooth random walk in (0, 0.66)), columns 1..36 noisy forecasts of the next
    36 five-minute values."""
    rng = np.random.default_rng(seed)
    T = 289 + 36
    walk = np.cumsum(rng.normal(0, 0.01, (num_days, T)), axis=1) + rng.uniform(0.15, 0.5, (num_days, 1))
    walk = np.abs(walk) % 1.32
    walk = np.where(walk > 0.66, 1.32 - walk, walk)
    out = np.empty((num_days, 289, 37), np.float32)
    out[:, :, 0] = walk[:, :289]
    for j in range(1, 37):
        out[:, :, j] = np.clip(walk[:, j:289 + j] + rng.normal(0, 0.004 * np.sqrt(j), (num_days, 289)), 0, 1)
    return out
