"""Golden statistics for the GMM day-trace sampler, from the reference's own pickled scikit-learn models.

Run in the build container (needs scikit-learn + /root/reference):

    python tests/golden/gen_ev_gmm_golden.py

For each packaged mixture: 400 000 samples drawn with ``GaussianMixture.sample`` (the call the reference
makes, event_generation.py:436), pushed through the reference's filters and rescaling (:441-458), and
summarised: acceptance rate, mean / std of arrival, departure, estimated departure (5-minute periods),
requested energy (kWh, before the cap), and the mean of ``count`` (sessions per day).  The oracle
(oracle/ev_gmm.py) and the kernel must reproduce these within sampling error.
"""
import os
import pickle
import warnings

import numpy as np

SRC = "/root/reference/sustaingym/data/evcharging/gmms"
out = {}
for site in ("caltech", "jpl"):
    for fn in sorted(os.listdir(os.path.join(SRC, site))):
        begin, end, n = fn[:-4].split(" ")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            d = pickle.load(open(os.path.join(SRC, site, fn), "rb"))
        g = d["gmm"]
        g.set_params(random_state=12345)
        x = g.sample(400_000)[0]
        keep = (0 <= x[:, 0]) & (x[:, 1] < 1) & (x[:, 2] < 1) & (x[:, 3] >= 0)
        y = x[keep].copy()
        y[:, :3] = 1440 * y[:, :3] // 5
        keep2 = (y[:, 0] < y[:, 1]) & (y[:, 0] < y[:, 2])
        y = y[keep2]
        y[:, 3] *= 100
        key = f"{site}|{begin}|{end}|"
        out[key + "accept"] = np.float64(len(y) / len(x))
        out[key + "mean"] = y.mean(0)
        out[key + "std"] = y.std(0)
        out[key + "count_mean"] = np.float64(np.mean(d["count"]))
        print(key, out[key + "accept"], out[key + "mean"], out[key + "std"], out[key + "count_mean"])
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ev_gmm_stats.npz"), **out)
