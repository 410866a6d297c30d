"""CPU: the pieces of the EV path that CAN be pinned to something the reference holds.

  * station order: the recalled registration order of acnportal's `caltech_acn()` (network.py) indexes the claimed-session
    counts of the reference's packaged ACN-Data CSVs into exactly the `station_usage` arrays of the reference's pickled
    GMMs (producer: sustaingym/envs/evcharging/train_gmm_model.py:86-106), for all four periods;
  * trace filters + energy cap + `_calculate_max_profit` (env.py:422-429): Caltech 2020-02-01..05-31, seed 2 gives
    max_profit 14.45262, the one number the reference's notebooks kept (examples/evcharging/env_validation.ipynb cell 24);
    checked on the committed sessions of that day and, when the reference data tree is present, through
    RealTraceGenerator itself (which is how a DST bug in its day index was found).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

PINS = os.path.join(GOLDEN, "ev_pins.npz")
REF_DATA = "/root/reference/sustaingym/data"


def test_station_order_reproduces_reference_station_usage():
    from sustaingym_b200.envs.evcharging.event_generation import DEFAULT_DATE_RANGES
    from sustaingym_b200.envs.evcharging.network import caltech_acn
    ids = caltech_acn().station_ids
    pins = np.load(PINS)
    gmms = np.load(os.path.join(ROOT, "sustaingym_b200", "data", "ev_gmms.npz"))
    for begin, end in DEFAULT_DATE_RANGES:
        counts = dict(zip(pins[f"usage_ids|{begin}|{end}"], pins[f"usage_counts|{begin}|{end}"]))
        ours = np.array([counts.get(s, 0) for s in ids], np.float64)            # station_id_cnts(), :86-106
        usage = gmms[f"caltech|{begin}|{end}|30|station_usage"]
        np.testing.assert_array_equal(ours, usage, err_msg=f"{begin}..{end}")
        assert ours.sum() > 1000 and (ours > 0).sum() >= 30                      # (ids outside the network are skipped, :100-101)


def test_max_profit_of_the_notebook_day_from_committed_sessions():
    from sustaingym_b200.data.ev_traces import pack_sessions
    pins = np.load(PINS)
    day = [(int(a), int(d), int(e), float(k), int(s)) for a, d, e, k, s in pins["spring2020_day2_sessions"]]
    tables = pack_sessions([day], 54)
    assert abs(tables.max_profit()[0] - float(pins["spring2020_day2_max_profit_reference"])) < 5e-6
    assert abs(tables.max_profit()[0] - 14.45262) < 5e-6


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference data tree not present")
def test_real_trace_generator_reproduces_notebook_max_profit():
    from sustaingym_b200.envs.evcharging.event_generation import RealTraceGenerator
    gen = RealTraceGenerator("caltech", ("2020-02-01", "2020-05-31"), data_dir=REF_DATA)
    assert gen.num_days == 121 and gen.moer_source == "real"
    day = int(gen.first_days(np.array([2]), 1)[0])                                # seed 2 -> 2020-02-03 (event_generation.py:273-282)
    assert abs(gen.tables.max_profit()[day] - 14.45262) < 5e-6
    pins = np.load(PINS)
    got = np.array(gen.tables.sessions_of_day(day), np.float64)
    np.testing.assert_array_equal(got, pins["spring2020_day2_sessions"])
    # the period spans the switch to daylight-saving time (2020-03-08): calendar days stay one trace each
    n = [(gen.tables.arrival[d] != 32767).sum() for d in range(gen.num_days)]
    assert n[36] > 0 and n[37] > 0 and sum(n) > 500
