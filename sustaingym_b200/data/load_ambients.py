"""CogenEnv ambient tables from the reference's packaged raw data.

Restates ``sustaingym/data/cogen/load_ambients.py:18-132`` (``load_wind_data`` + ``construct_df``) without
pandas' Excel stack: operating data (15-minute plant targets and ambient conditions, ``operating_data.xlsx``)
joined with the ERCOT day-ahead HB_HOUSTON settlement price (two yearly workbooks, hourly, forward-filled to
15 minutes; days that do not have 24 hourly rows — daylight-saving switches — are dropped before the
fill, :69-76), the Henry Hub daily gas spot price (missing days forward-filled, :79-92) and an IEC class 2
wind-turbine curve applied to a 15-minute wind-speed series that is subtracted from the power target
(:18-36, :120-126).  Output: ``[n_days, 96, 7]`` float64 in ``AMBIENT_COLS`` order, first and last calendar
day dropped and only days with exactly 96 rows kept (:128-134) — what ``BatchedCogenEnv(ambients=...)`` takes.

    python -m sustaingym_b200.data.load_ambients DATA_DIR OUT.npz

``DATA_DIR`` is the reference's ``sustaingym/data/cogen/ambients_data``.
"""
from __future__ import annotations

import bisect
import csv
import datetime as dt
import os
import sys

import numpy as np

from .xlsx_reader import excel_serial_to_datetime, parse_us_date, read_xlsx

WIND_CURVE_PTS = [0, 0, 0, 0.0052, 0.0423, 0.1031, 0.1909, 0.3127, 0.4731, 0.6693, 0.8554, 0.9641, 0.9942, 0.9994,
                  1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0]      # load_ambients.py:23-25


def load_wind_data(data_dir: str, n_mw: float) -> np.ndarray:
    """Wind generation [MW] per 15-minute row of the wind-speed file (load_ambients.py:18-36)."""
    speeds = []
    with open(os.path.join(data_dir, "0_39.97_-128.77_2019_15min.csv"), newline="") as f:
        rd = csv.reader(f)
        next(rd)                                   # site metadata line (header=1 in the reference)
        col = next(rd).index("wind speed at 100m (m/s)")
        for row in rd:
            if row:
                speeds.append(float(row[col]))
    return n_mw * np.interp(np.asarray(speeds), np.arange(32), WIND_CURVE_PTS)


def _energy_prices(data_dir: str):
    """Sorted (hour-beginning datetimes, HB_HOUSTON prices) after dropping the days without 24 rows."""
    rows = []
    for year in (2021, 2022):
        wb = read_xlsx(os.path.join(data_dir, f"rpt.00013060.0000000000000000.DAMLZHBSPP_{year}.xlsx"))
        for sheet in wb.values():
            hdr = sheet[0]
            i_date, i_hour = hdr.index("Delivery Date"), hdr.index("Hour Ending")
            i_pt, i_price = hdr.index("Settlement Point"), hdr.index("Settlement Point Price")
            for r in sheet[1:]:
                if len(r) > i_price and r[i_pt] == "HB_HOUSTON":
                    d = parse_us_date(r[i_date]) if isinstance(r[i_date], str) else excel_serial_to_datetime(r[i_date]).date()
                    hb = int(str(r[i_hour])[:2]) - 1                       # Hour Beginning = Hour Ending - 1 h (:62-64)
                    rows.append((dt.datetime(d.year, d.month, d.day) + dt.timedelta(hours=hb), float(r[i_price])))
    per_day: dict = {}
    for t, _ in rows:
        per_day[t.date()] = per_day.get(t.date(), 0) + 1
    rows = sorted((t, p) for t, p in rows if per_day[t.date()] == 24)       # :69-72
    return [t for t, _ in rows], [p for _, p in rows]


def _gas_prices(data_dir: str):
    days, prices = [], []
    with open(os.path.join(data_dir, "Henry_Hub_Natural_Gas_Spot_Price.csv"), newline="") as f:
        rd = csv.reader(f)
        for _ in range(4):
            next(rd)                               # header=4 in the reference
        next(rd)
        for row in rd:
            if len(row) >= 2 and row[0].strip():
                days.append(parse_us_date(row[0]))
                prices.append(float(row[1]) if row[1].strip() else float("nan"))
    order = np.argsort(days)
    days = [days[i] for i in order]
    prices = [prices[i] for i in order]
    for i in range(1, len(prices)):                # ffill of NaNs (:86-87); missing days are covered by the look-up
        if prices[i] != prices[i]:
            prices[i] = prices[i - 1]
    return days, prices


def construct_ambients(data_dir: str, renewables_magnitude: float = 0.0, return_wind_cf: bool = False):
    """``construct_df`` (load_ambients.py:38-134) as one ``[n_days, 96, 7]`` float64 array (and, on request, the
    wind capacity factor ``[n_days, 96]`` of every kept row: target power = max(raw target - n_mw * cf, 0))."""
    e_t, e_p = _energy_prices(data_dir)
    g_d, g_p = _gas_prices(data_dir)
    sheet = next(iter(read_xlsx(os.path.join(data_dir, "operating_data.xlsx")).values()))
    hdr = sheet[3]                                 # header=3 (:97)
    cols = [hdr.index(c) for c in ("Timestamp", "Target Net Power", "Target Process Steam", "Ambient Temperature",
                                   "Ambient Pressure", "Ambient rel. Humidity")]
    recs = []
    for r in sheet[4:]:
        if len(r) <= max(cols) or r[cols[0]] is None:
            continue
        ts = excel_serial_to_datetime(r[cols[0]])
        recs.append((ts, *[float(r[c]) for c in cols[1:]]))
    cf = load_wind_data(data_dir, 1.0)[:len(recs)]
    wind = float(renewables_magnitude) * cf
    table, stamps = [], []
    for k, (ts, power, steam, tamb, pamb, rh) in enumerate(recs):
        ie = bisect.bisect_right(e_t, ts) - 1      # resample('15min').ffill(): the latest hourly price at or before ts
        ig = bisect.bisect_right(g_d, ts.date()) - 1
        if ie < 0 or ig < 0 or ts > e_t[-1] + dt.timedelta(minutes=45) or ts.date() > g_d[-1]:
            raise KeyError(f"{ts} is outside the price tables")      # the reference's .loc raises as well
        power = max(power - (wind[k] if k < len(wind) else 0.0), 0.0)  # :125-126
        table.append((tamb, pamb, rh / 100.0, power, steam, e_p[ie], g_p[ig], cf[k] if k < len(cf) else 0.0))
        stamps.append(ts)
    by_day: dict = {}
    for ts, row in zip(stamps, table):
        by_day.setdefault(ts.date(), []).append(row)
    days = [by_day[d] for d in sorted(by_day)][1:-1]                 # :130
    days = [d for d in days if len(d) == 96]                         # :132
    arr = np.asarray(days, np.float64)
    return (arr[..., :7].copy(), arr[..., 7].copy()) if return_wind_cf else arr[..., :7].copy()


if __name__ == "__main__":
    out, cf = construct_ambients(sys.argv[1], 0.0, return_wind_cf=True)
    np.savez_compressed(sys.argv[2], ambients=out, wind_cf=cf)      # wind = 0 table + capacity factors: any magnitude follows
    print(out.shape, "->", sys.argv[2])
