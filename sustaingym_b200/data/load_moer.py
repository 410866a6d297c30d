"""Marginal-emission-rate tables from the reference's packaged monthly files.

Restates ``MOERLoader`` / ``load_moer`` (sustaingym/data/load_moer.py:306-377): the files
``data/moer/<BA>_<YYYY>-<MM>.csv.gz`` hold one row per 5 minutes (UTC index; column ``moer`` = realised
rate, ``f1..f36`` = forecasts for the next 36 steps, kg CO2 / kWh); months are concatenated with rows
that overlap the previous file dropped (:326-330); ``retrieve(day)`` returns the 289 rows from the
day's local midnight to the next midnight inclusive (:364-377).  The engine wants the days of a period
stacked once: ``[n_days, 289, 37]`` float32, ready for the device.
"""
from __future__ import annotations

import datetime as dt
import os

import numpy as np

BA_CALTECH_JPL = "SGIP_CAISO_SCE"      # event_generation.py: balancing authority of both garages


def load_moer_days(moer_dir: str, first_day: str | dt.date, num_days: int, ba: str = BA_CALTECH_JPL,
                   reference_day_start: bool = True) -> np.ndarray:
    """MOER tables of ``num_days`` consecutive days starting at ``first_day``: [num_days, 289, 37] f32; row t of
    day d is ``start(d) + 5 t`` minutes (row 288 closes the day), like ``MOERLoader.retrieve(day)``.

    ``reference_day_start``: the reference builds its day stamps with ``datetime.replace(tzinfo=pytz.timezone(
    'America/Los_Angeles'))`` (evcharging/utils.py:78-80), which attaches pytz's LMT offset of -7:53 instead of
    PST/PDT, so its ``retrieve(day)`` window starts at 07:53 UTC (first row 07:55 UTC = 00:55 PDT / 23:55 PST).
    True reproduces that; False starts at the local civil midnight."""
    import pandas as pd
    d0 = pd.Timestamp(str(first_day))
    if reference_day_start:
        first = pd.Timestamp(year=d0.year, month=d0.month, day=d0.day, hour=7, minute=53, tz="UTC")
        starts = [first + pd.Timedelta(days=d) for d in range(num_days)]
    else:
        starts = [(d0 + pd.Timedelta(days=d)).tz_localize("America/Los_Angeles").tz_convert("UTC") for d in range(num_days)]
    end_utc = starts[-1] + pd.Timedelta(days=1, minutes=5)
    y, m = starts[0].year, starts[0].month
    frames = []
    while (y, m) <= (end_utc.year, end_utc.month):
        path = os.path.join(moer_dir, f"{ba}_{y}-{m:02d}.csv.gz")
        if os.path.exists(path):                    # (a file runs ~2 days past its month: the last one may be dispensable)
            df = pd.read_csv(path, compression="gzip", index_col="time")
            df.index = pd.to_datetime(pd.DatetimeIndex(df.index))
            df.sort_index(inplace=True)
            if frames:                              # load_moer.py:326-330: drop rows the previous file already holds
                df = df[df.index > frames[-1].index[-1]]
            frames.append(df)
        y, m = (y + 1, 1) if m == 12 else (y, m + 1)
    if not frames:
        raise FileNotFoundError(f"no {ba}_YYYY-MM.csv.gz files for {first_day} under {moer_dir}")
    df = pd.concat(frames, axis=0)
    out = np.empty((num_days, 289, 37), np.float32)
    for d, t0 in enumerate(starts):
        rows = df[(t0 <= df.index) & (df.index < t0 + pd.Timedelta(days=1, minutes=5))].values
        if rows.shape != (289, 37):
            raise ValueError(f"MOER table of day {d} ({t0}) has shape {rows.shape}, expected (289, 37)")
        out[d] = rows
    return out
