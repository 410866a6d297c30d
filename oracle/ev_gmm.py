"""TEST INFRASTRUCTURE — CPU oracle for the GMM day-trace sampler.

Restates ``GMMsTraceGenerator._sample`` and ``._create_events``
(/root/reference/sustaingym/envs/evcharging/event_generation.py:416-515):

* number of sessions of the day drawn from the empirical per-day counts (:479-480);
* ``int(n * 1.2)`` candidates per batch from the Gaussian mixture, GROUPED BY COMPONENT like
  ``sklearn.mixture.GaussianMixture.sample`` returns them, filtered (:441-444, :452-455), rescaled to
  5-minute periods (:447-449) and to kWh (:458), batches concatenated and cut to the first n (:460-463)
  — the cut therefore favours low component indices, as in the reference;
* energy clipped to ``[0, requested_energy_cap]`` (:486), events sorted by arrival (:490);
* stations drawn sequentially among those free at the arrival, with the empirical usage probabilities,
  uniformly when their mass is <= 1e-5; sessions that find no free station are dropped (:492-515).

What is NOT reproducible from the reference: its random streams (NumPy ``Generator`` + scikit-learn's
``RandomState``).  The sampler therefore defines its own counter-based stream — Philox4x32-10 keyed by
the seed, one counter lane per day — and every floating-point step is written with individually rounded
IEEE-754 double operations (no fused multiply-add, an explicit polynomial logarithm), so that the CUDA
kernel ``sg_ev_sample_gmm_days`` and this oracle agree BIT FOR BIT.  PARITY with the reference is
distributional only and is pinned by ``tests/test_ev_gmm.py`` against the mixture's analytic moments.

Only ``tests/`` may import this module.
"""
from __future__ import annotations

import math

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
EMPTY = 32767


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for r in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


class Stream:
    """Random stream of one day: key = seed, counter = (block index, day lo, day hi, 0)."""

    def __init__(self, seed: int, day: int):
        self.key = (seed & MASK, (seed >> 32) & MASK)
        self.day = (day & MASK, (day >> 32) & MASK)
        self.block = 0
        self.buf: list[int] = []

    def u32(self) -> int:
        if not self.buf:
            self.buf = list(philox4x32_10((self.block, self.day[0], self.day[1], 0), self.key))
            self.block = (self.block + 1) & MASK
        return self.buf.pop(0)

    def uniform(self) -> float:
        """53-bit uniform in [0, 1)."""
        a, b = self.u32() >> 5, self.u32() >> 6
        return (a * 67108864.0 + b) / 9007199254740992.0


LN2 = 0.6931471805599453
SQRT_HALF = 0.7071067811865476
LOG_COEF = [2.0 / (2 * k + 1) for k in range(1, 12)]     # atanh series: 2/3, 2/5, ... 2/23


def det_log(x: float) -> float:
    """ln(x), x > 0 finite and normal: argument reduction to [sqrt(1/2), sqrt(2)) and the atanh series,
    every operation rounded once (the kernel uses the same sequence with __dmul_rn/__dadd_rn/__ddiv_rn)."""
    m, e = math.frexp(x)
    if m < SQRT_HALF:
        m = m * 2.0
        e -= 1
    s = (m - 1.0) / (m + 1.0)
    z = s * s
    p = LOG_COEF[10]
    for k in range(9, -1, -1):
        p = p * z + LOG_COEF[k]
    r = 2.0 * s + s * (z * p)
    return float(e) * LN2 + r


def normal_pair(rs: Stream):
    """Marsaglia polar method."""
    while True:
        u1 = 2.0 * rs.uniform() - 1.0
        u2 = 2.0 * rs.uniform() - 1.0
        s = u1 * u1 + u2 * u2
        if 0.0 < s < 1.0:
            break
    f = math.sqrt((-2.0 * det_log(s)) / s)
    return u1 * f, u2 * f


def sample_batch(rs: Stream, gmm: dict, m: int, requested_energy_cap: float) -> list:
    """One ``gmm.sample(m)`` call of the reference (:436) followed by its filters and rescaling (:441-458):
    the accepted (arrival, departure, est_departure, kWh) candidates, grouped by component."""
    cum_w, means, chol = gmm["cum_weights"], gmm["means"], gmm["chol"]
    nc = len(cum_w)
    per_comp = [0] * nc
    for _ in range(m):
        u = rs.uniform()
        c = 0
        while c < nc - 1 and not (cum_w[c] > u):
            c += 1
        per_comp[c] += 1
    out = []
    for c in range(nc):
        for _ in range(per_comp[c]):
            z = [0.0] * 4
            z[0], z[1] = normal_pair(rs)
            z[2], z[3] = normal_pair(rs)
            x = [0.0] * 4
            for i in range(4):
                acc = float(means[c][i])
                for j in range(i + 1):
                    acc = acc + float(chol[c][i][j]) * z[j]
                x[i] = acc
            if not (0.0 <= x[0] and x[1] < 1.0 and x[2] < 1.0 and x[3] >= 0.0):
                continue
            arr = math.floor((1440.0 * x[0]) / 5.0)
            dep = math.floor((1440.0 * x[1]) / 5.0)
            est = math.floor((1440.0 * x[2]) / 5.0)
            if not (arr < dep and arr < est):
                continue
            out.append((int(arr), int(dep), int(est), min(max(x[3] * 100.0, 0.0), requested_energy_cap)))
    return out


def sample_day(gmm: dict, seed: int, day: int, requested_energy_cap: float, max_sessions: int,
               max_day_sessions: int, kwh_per_step: float):
    """Returns (events, dropped, max_kwh): events = [(arrival, departure, est_departure, kWh, station,
    slot)] in arrival order (ties in generation order)."""
    counts, p_station = gmm["counts"], gmm["station_p"]
    n_st = len(p_station)
    rs = Stream(seed, day)
    n = int(counts[min(int(rs.uniform() * len(counts)), len(counts) - 1)])
    n = min(n, 256)                                       # the kernel's candidate buffer
    m = int(n * 1.2)
    cand = []
    batches = 0
    while len(cand) < n and batches < 64:
        batches += 1
        cand.extend(sample_batch(rs, gmm, m, requested_energy_cap))
    cand = cand[:n]                                       # :463
    # stable sort by arrival
    cand.sort(key=lambda s: s[0])
    station_dep = [-1] * n_st
    used = [0] * n_st
    events, dropped, max_kwh = [], 0, 0.0
    for (arr, dep, est, kwh) in cand:
        avail = [i for i in range(n_st) if station_dep[i] < arr]
        u = rs.uniform()
        if not avail:
            dropped += 1
            continue
        total = 0.0
        for i in avail:
            total = total + float(p_station[i])
        if total <= 1e-5:
            idx = avail[min(int(u * len(avail)), len(avail) - 1)]
        else:
            target = u * total
            cum, idx = 0.0, avail[-1]
            for i in avail:
                cum = cum + float(p_station[i])
                if cum > target:
                    idx = i
                    break
        if used[idx] >= max_sessions or len(events) >= max_day_sessions:
            dropped += 1
            continue
        station_dep[idx] = max(dep, station_dep[idx])
        events.append((arr, dep, est, kwh, idx, used[idx]))
        used[idx] += 1
        max_kwh = max_kwh + min(kwh, float(dep - arr) * kwh_per_step)
    return events, dropped, max_kwh


def sample_days(gmm: dict, seed: int, num_days: int, n_stations: int, requested_energy_cap: float = 100.0,
                max_sessions: int = 8, max_day_sessions: int = 128, kwh_per_step: float = 32 * 5 / 60 * 0.208,
                first_day: int = 0):
    """Device-layout tables of ``num_days`` sampled days: (records [D,K,n] structured, arrivals [D,Smax+1]
    uint32, dropped [D] int32, max_kwh [D] float64)."""
    rec = np.zeros((num_days, max_sessions, n_stations),
                   dtype=np.dtype([("arr", "<i2"), ("dep", "<i2"), ("est", "<i2"), ("res", "<i2"), ("energy", "<f8")]))
    rec["arr"] = rec["dep"] = rec["est"] = EMPTY
    arrivals = np.full((num_days, max_day_sessions + 1), 0x7FFF, np.uint32)
    dropped = np.zeros(num_days, np.int32)
    max_kwh = np.zeros(num_days, np.float64)
    for d in range(num_days):
        ev, dr, mk = sample_day(gmm, seed, first_day + d, requested_energy_cap, max_sessions, max_day_sessions, kwh_per_step)
        dropped[d], max_kwh[d] = dr, mk
        keys = []
        for (arr, dep, est, kwh, stn, slot) in ev:
            rec[d, slot, stn] = (arr, dep, est, 0, kwh)
            keys.append((arr, stn, slot))
        keys.sort()
        for s, (arr, stn, slot) in enumerate(keys):
            arrivals[d, s] = arr | (stn << 16) | (slot << 24)
    return rec, arrivals, dropped, max_kwh


def mixture_moments(gmm: dict):
    """Mean and covariance of the unfiltered mixture (for the distributional pin)."""
    w = np.diff(np.concatenate([[0.0], gmm["cum_weights"]]))
    mu = (w[:, None] * gmm["means"]).sum(0)
    cov = np.zeros((4, 4))
    for c in range(len(w)):
        L = gmm["chol"][c]
        dm = gmm["means"][c] - mu
        cov += w[c] * (L @ L.T + np.outer(dm, dm))
    return mu, cov
