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

import os

from ...data.ev_traces import EMPTY, TraceTables, pack_sessions, synthetic_days, synthetic_moer
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
        """MOER table index that goes with each trace index.  Sequential generators: the day after the events' day
        (see the module docstring).  Non-sequential generators: the reference draws the events' day at random and
        fetches THAT day's MOER (event_generation.py:479-480, env.py:321-323) — with an i.i.d. pool of synthetic
        days the equivalent is an independent random calendar day of the period."""
        if not self.sequential and getattr(self, "moer_source", "") == "real":
            return self.rng.integers(0, self.moer.shape[0], len(np.asarray(days))).astype(np.int32)
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


# ---- the reference's default periods (sustaingym/envs/evcharging/utils.py:48-64) ----------------------
DEFAULT_DATE_RANGES = (("2019-05-01", "2019-08-31"), ("2019-09-01", "2019-12-31"),
                       ("2020-02-01", "2020-05-31"), ("2021-05-01", "2021-08-31"))
DEFAULT_PERIOD_TO_RANGE = {
    "Summer 2019": DEFAULT_DATE_RANGES[0], "Pre-COVID-19 Summer": DEFAULT_DATE_RANGES[0],
    "Fall 2019": DEFAULT_DATE_RANGES[1], "Pre-COVID-19 Fall": DEFAULT_DATE_RANGES[1],
    "Spring 2020": DEFAULT_DATE_RANGES[2], "In-COVID-19": DEFAULT_DATE_RANGES[2],
    "Summer 2021": DEFAULT_DATE_RANGES[3], "Post-COVID-19": DEFAULT_DATE_RANGES[3],
}
_GMM_FILE = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "data", "ev_gmms.npz")


def _resolve_moer(moer, data_dir: str | None, date_period, synthetic_seed: int):
    """MOER tables of a generator: an explicit ``[days, 289, 37]`` array; ``"synthetic"`` (explicit opt-in: smooth
    random-walk tables, NOT the reference's data); or ``None`` = the reference's SGIP tables of the period's calendar
    days, read from ``data_dir/moer`` (``$SUSTAINGYM_DATA_DIR``) through ``data/load_moer.py`` — and an error when
    they are not there, because the carbon term of the reward and two observation fields depend on them
    (``MOERLoader.retrieve(self.day)``, event_generation.py:209-218).  Returns (array f32, source tag)."""
    if isinstance(moer, str):
        if moer != "synthetic":
            raise ValueError("moer must be an array, None (real SGIP tables) or 'synthetic'")
        import pandas as pd
        begin, end = _date_range(date_period)
        n_days = (pd.Timestamp(end) - pd.Timestamp(begin)).days + 1
        return np.ascontiguousarray(synthetic_moer(n_days + 1, synthetic_seed), dtype=np.float32), "synthetic"
    if moer is not None:
        return np.ascontiguousarray(moer, dtype=np.float32), "array"
    data_dir = data_dir or os.environ.get("SUSTAINGYM_DATA_DIR")
    moer_dir = os.path.join(data_dir, "moer") if data_dir else None
    if moer_dir is None or not os.path.isdir(moer_dir):
        raise FileNotFoundError(
            "no MOER tables: pass data_dir (or set $SUSTAINGYM_DATA_DIR) to the reference's sustaingym/data so that "
            "data_dir/moer/SGIP_CAISO_*.csv.gz can be read, pass moer=<[days,289,37] array>, or opt in to "
            "moer='synthetic' (random-walk tables that do NOT match the reference's carbon costs)")
    import pandas as pd
    from ...data.load_moer import load_moer_days
    begin, end = _date_range(date_period)
    n_days = (pd.Timestamp(end) - pd.Timestamp(begin)).days + 1
    return np.ascontiguousarray(load_moer_days(moer_dir, begin, n_days), dtype=np.float32), "real"


def _date_range(date_period) -> tuple[str, str]:
    if isinstance(date_period, str):
        return DEFAULT_PERIOD_TO_RANGE[date_period]
    return tuple(date_period)


def load_gmm(site: str, date_period, n_components: int = 30) -> dict:
    """The packaged mixture of a site / period as flat arrays (``tools/export_ev_gmms.py``; the reference's
    ``load_gmm_model``, utils.py:248-279): cum_weights [nc], means [nc,4], chol [nc,4,4], counts, station_p [n]."""
    begin, end = _date_range(date_period)
    key = f"{site}|{begin}|{end}|{n_components}|"
    with np.load(_GMM_FILE) as z:
        if key + "weights" not in z:
            raise FileNotFoundError(f"no packaged {n_components}-component GMM for {site} {begin}..{end}")
        usage = z[key + "station_usage"].astype(np.float64)
        return {"cum_weights": np.cumsum(z[key + "weights"]), "means": z[key + "means"].copy(),
                "chol": z[key + "chol"].copy(), "counts": z[key + "counts"].astype(np.int32),
                "station_p": usage / usage.sum()}


class GmmDevice:
    """A mixture resident on the device + the launch of ``sg_ev_sample_gmm_days`` (include/sustaingym_b200.h)."""

    def __init__(self, gmm: dict, num_stations: int, device):
        import ctypes as C

        import torch

        from ... import _lib
        self._lib, self._C, self.device, self.num_stations = _lib, C, torch.device(device), num_stations
        self._t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(self.device) for k, v in gmm.items()}
        self.struct = _lib.EVGmm(len(gmm["cum_weights"]), len(gmm["counts"]), num_stations, 0,
                                 self._t["cum_weights"].data_ptr(), self._t["means"].data_ptr(),
                                 self._t["chol"].data_ptr(), self._t["counts"].data_ptr(),
                                 self._t["station_p"].data_ptr())

    def alloc_pool(self, num_days: int, max_sessions: int, max_day_sessions: int) -> dict:
        import torch
        dev, n = self.device, self.num_stations
        return {"sessions": torch.zeros(num_days * max_sessions * n * 16, dtype=torch.uint8, device=dev),
                "arrivals": torch.zeros(num_days, max_day_sessions + 1, dtype=torch.int32, device=dev),
                "dropped": torch.zeros(num_days, dtype=torch.int32, device=dev),
                "max_kwh": torch.zeros(num_days, dtype=torch.float64, device=dev),
                "max_sessions": max_sessions}

    def sample(self, pool: dict, first_day: int, seed: int, requested_energy_cap: float, kwh_per_step: float) -> None:
        """Fills ``pool`` with days first_day .. first_day + D - 1 of stream ``seed`` (asynchronous)."""
        import torch
        D, smax = pool["arrivals"].shape[0], pool["arrivals"].shape[1] - 1
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.load().sg_ev_sample_gmm_days(
                self._C.byref(self.struct), D, first_day, seed, requested_energy_cap, kwh_per_step,
                pool["max_sessions"], smax, pool["sessions"].data_ptr(), pool["arrivals"].data_ptr(),
                pool["dropped"].data_ptr(), pool["max_kwh"].data_ptr(),
                torch.cuda.current_stream(self.device).cuda_stream), "sg_ev_sample_gmm_days")


class GMMsTraceGenerator(AbstractTraceGenerator):
    """``GMMsTraceGenerator`` of the reference (event_generation.py:330-515) with the sampling on the GPU.

    The reference samples ONE day per ``reset()`` on the host.  Here a pool of ``num_days`` days is sampled
    by one kernel launch (``sg_ev_sample_gmm_days``: one thread per day) directly into the device tables
    the step kernels read, so a reset stays an index bump; ``resample()`` refills the same buffers with the
    next ``num_days`` days of the stream (pointers do not change: captured CUDA graphs stay valid).
    Day d of a stream is a pure function of ``(seed, d)``.  ``gmm`` may be passed as a dict of arrays
    (``load_gmm`` layout) for custom mixtures.
    """

    A_PERS_TO_KWH = 5 / 60 * 0.208      # env.py:107-110 (one 5-minute period at 1 A, 208 V)

    def __init__(self, site="caltech", date_period="Summer 2019", n_components: int = 30,
                 requested_energy_cap: float = 100, seed: int | None = None, num_days: int = 1024,
                 max_sessions: int = 8, max_day_sessions: int = 128, moer: np.ndarray | str | None = None,
                 gmm: dict | None = None, device="cuda", data_dir: str | None = None, **kw):
        import torch
        kw.setdefault("sequential", False)           # the reference draws i.i.d. days (:479-480)
        super().__init__(site, requested_energy_cap, seed=seed, **kw)
        self.date_period = date_period
        self.n_components = n_components
        self.gmm = gmm if gmm is not None else load_gmm(self.site, date_period, n_components)
        if len(self.gmm["station_p"]) != self.num_stations:
            raise ValueError("the mixture's station usage does not match the charging network")
        self.seed = 0 if seed is None else int(seed)
        self._num_days, self._K, self._smax = int(num_days), int(max_sessions), int(max_day_sessions)
        self.device = torch.device(device)
        self._gmm_dev = GmmDevice(self.gmm, self.num_stations, self.device)
        self.device_pool = self._gmm_dev.alloc_pool(self._num_days, self._K, self._smax)
        self.moer, self.moer_source = _resolve_moer(moer, data_dir, date_period, self.seed + 1)
        self._next_day = 0
        self.resample()

    @property
    def num_days(self) -> int:
        return self._num_days

    def resample(self) -> None:
        """Refills the pool with the next ``num_days`` days of the stream (asynchronous, current stream)."""
        self._gmm_dev.sample(self.device_pool, self._next_day, self.seed, float(self.requested_energy_cap),
                             32 * self.A_PERS_TO_KWH)
        self._next_day += self._num_days
        self._tables = None
        self.pool_version = getattr(self, "pool_version", 0) + 1

    @property
    def tables(self) -> TraceTables:
        """Host copy of the pool in the ``TraceTables`` form (for inspection / the oracle)."""
        if self._tables is None:
            rec = self.device_pool["sessions"].cpu().numpy().view(
                np.dtype([("arr", "<i2"), ("dep", "<i2"), ("est", "<i2"), ("res", "<i2"), ("energy", "<f8")])
            ).reshape(self._num_days, self._K, self.num_stations)
            self._tables = TraceTables(rec["arr"].copy(), rec["dep"].copy(), rec["est"].copy(), rec["energy"].copy())
        return self._tables

    @tables.setter
    def tables(self, value) -> None:      # AbstractTraceGenerator.__init__ assigns None
        self._tables = value

    def __repr__(self) -> str:
        begin, end = _date_range(self.date_period)
        return (f"{self.n_components}-component GMM-based trace generator for {self.site} ({begin} to {end}), "
                f"pool of {self._num_days} days on {self.device}, MOER: {self.moer_source}")


class RealTraceGenerator(ArrayTraceGenerator):
    """``RealTraceGenerator`` of the reference (event_generation.py:210-328): the real ACN-Data sessions of a
    site / period, one trace per calendar day, read from the reference's packaged
    ``data/evcharging/acn_data/<site>/<begin> <end>.csv.gz`` files under ``data_dir`` (a checkout of the
    reference's ``sustaingym/data``; the files are data, not shipped here).  Filters follow
    ``_create_events`` (:293-328): claimed sessions only unless ``use_unclaimed``, stations of the network,
    departure and estimated departure on the arrival's calendar day, estimated departure after arrival.  MOER tables
    come from ``data_dir/moer`` (``data/load_moer.py``; day d+1's table goes with day d's sessions and the last day
    wraps to the first, as the reference's day cursor does).  Not reproduced: the reference's pytz LMT day stamp makes
    its session window run 00:53-00:53 local, which drops the rare sessions that arrive before 00:53."""

    def __init__(self, site="caltech", date_period="Summer 2019", data_dir: str | None = None,
                 sequential: bool = True, use_unclaimed: bool = False, requested_energy_cap: float = 100,
                 moer: np.ndarray | str | None = None, seed: int | None = None, **kw):
        import pandas as pd
        net = site_str_to_site(site) if isinstance(site, str) else site
        begin, end = _date_range(date_period)
        data_dir = data_dir or os.environ.get("SUSTAINGYM_DATA_DIR")
        if data_dir is None:
            raise FileNotFoundError("RealTraceGenerator needs data_dir (or $SUSTAINGYM_DATA_DIR): the reference's sustaingym/data")
        path = None
        for b, e in DEFAULT_DATE_RANGES:
            if b <= begin and end <= e:
                path = os.path.join(data_dir, "evcharging", "acn_data", net.site, f"{b} {e}.csv.gz")
        if path is None or not os.path.exists(path):
            raise FileNotFoundError(f"no packaged ACN-Data file covers {net.site} {begin}..{end} under {data_dir}")
        df = pd.read_csv(path, compression="gzip").rename(columns={"requested_energy (kWh)": "kwh"})
        for col in ("arrival", "departure", "estimated_departure"):
            df[col] = pd.to_datetime(df[col], utc=True).dt.tz_convert("America/Los_Angeles")
        if not use_unclaimed:
            df = df[df["claimed"]]
        index = {sid: i for i, sid in enumerate(net.station_ids)}
        df = df[df["station_id"].isin(index)]
        first = pd.Timestamp(begin)                    # calendar dates (naive): a period may span a DST change
        n_days = (pd.Timestamp(end) - first).days + 1
        days = [[] for _ in range(n_days)]
        step = self.TIME_STEP_DURATION
        for r in df.itertuples(index=False):
            d = (r.arrival.tz_localize(None).normalize() - first).days
            if not (0 <= d < n_days):
                continue
            if max(r.departure, r.estimated_departure).day != r.arrival.day:      # :312-315
                continue
            arr, dep, est = ((x.hour * 60 + x.minute) // step for x in (r.arrival, r.departure, r.estimated_departure))
            if est <= arr:                                                        # :326
                continue
            days[d].append((arr, dep, est, float(r.kwh), index[r.station_id]))
        self.date_period, self.use_unclaimed = date_period, use_unclaimed
        # the reference's MOERLoader tables of the same days (get_moer, :209-218); synthetic tables only on request
        moer, self.moer_source = _resolve_moer(moer, data_dir, date_period, 1)
        super().__init__(days, moer, net, requested_energy_cap, sequential=sequential, seed=seed, **kw)

    def __repr__(self) -> str:
        begin, end = _date_range(self.date_period)
        return f"Real trace generator for {self.site} ({begin} to {end}), {self.num_days} days, MOER: {self.moer_source}"
