"""Fixtures that pin the EV trace path to things the REFERENCE holds:

  * per Caltech period, the number of claimed sessions per station id in the reference's packaged ACN-Data CSVs
    (producer of the pickled `station_usage`: sustaingym/envs/evcharging/train_gmm_model.py:86-106);
  * the sessions of 2020-02-03 (Caltech, period 2020-02-01..2020-05-31, seed 2 of a sequential RealTraceGenerator),
    the day behind the one number the reference's notebooks kept: `max_profit 14.45262`
    (examples/evcharging/env_validation.ipynb, cell 24).

    python tests/golden/gen_ev_pins.py /root/reference/sustaingym/data   ->  tests/golden/ev_pins.npz
"""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sustaingym_b200.envs.evcharging.event_generation import DEFAULT_DATE_RANGES, RealTraceGenerator  # noqa: E402


def main(data_dir: str) -> None:
    out = {}
    for begin, end in DEFAULT_DATE_RANGES:
        df = pd.read_csv(os.path.join(data_dir, "evcharging", "acn_data", "caltech", f"{begin} {end}.csv.gz"), compression="gzip")
        vc = df[df["claimed"]]["station_id"].value_counts()
        ids = np.array(sorted(vc.index))
        out[f"usage_ids|{begin}|{end}"] = ids
        out[f"usage_counts|{begin}|{end}"] = np.array([vc[i] for i in ids], np.int64)
    gen = RealTraceGenerator("caltech", ("2020-02-01", "2020-05-31"), data_dir=data_dir, moer="synthetic")
    day = gen.tables.sessions_of_day(2)
    out["spring2020_day2_sessions"] = np.array(day, np.float64)          # arrival, departure, est. departure, kWh, station
    out["spring2020_day2_max_profit_reference"] = np.array(14.45262)     # the notebook's number
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ev_pins.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/sustaingym/data")
