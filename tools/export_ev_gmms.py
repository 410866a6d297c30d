"""Converts the reference's pickled scikit-learn mixtures into the flat arrays the GPU sampler reads.

Run once in the build container (needs scikit-learn to unpickle and the reference tree):

    python tools/export_ev_gmms.py /root/reference/sustaingym/data/evcharging/gmms sustaingym_b200/data/ev_gmms.npz

Per ``<site>/<begin> <end> <n_components>.pkl`` (sustaingym/envs/evcharging/utils.py:207-279: keys
'gmm', 'count', 'station_usage') the archive holds ``<site>|<begin>|<end>|<n>|{weights,means,chol,counts,
station_usage}``: mixture weights, component means [nc,4] (arrival, departure, estimated departure as
fractions of a day, requested energy / 100 kWh), lower Cholesky factors of the full covariances [nc,4,4]
(what GaussianMixture.sample multiplies standard normals with), the per-day session counts and the
per-station usage counts of the period.
"""
import os
import pickle
import sys
import warnings

import numpy as np


def main(src: str, dst: str) -> None:
    out = {}
    for site in sorted(os.listdir(src)):
        for fn in sorted(os.listdir(os.path.join(src, site))):
            if not fn.endswith(".pkl"):
                continue
            begin, end, n = fn[:-4].split(" ")
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                with open(os.path.join(src, site, fn), "rb") as f:
                    d = pickle.load(f)
            g = d["gmm"]
            assert g.covariance_type == "full"
            key = f"{site}|{begin}|{end}|{n}|"
            out[key + "weights"] = np.asarray(g.weights_, np.float64)
            out[key + "means"] = np.asarray(g.means_, np.float64)
            out[key + "chol"] = np.linalg.cholesky(np.asarray(g.covariances_, np.float64))
            out[key + "counts"] = np.asarray(d["count"], np.int32)
            out[key + "station_usage"] = np.asarray(d["station_usage"], np.float64)
            print(key, g.means_.shape, len(d["count"]), len(d["station_usage"]))
    np.savez_compressed(dst, **out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
