"""Charging-network description consumed by the EV kernels.

The reference builds its network with ``acnportal.acnsim.sites.caltech_acn()``
(``sustaingym/envs/evcharging/utils.py:83-88``) and reads ``station_ids``, ``min_pilot_signals``,
``constraint_matrix``, ``magnitudes``, ``_phase_angles`` from it (``env.py:132-134, 373, 450-451,
485-486``).  acnportal is not vendored in the reference tree; the Caltech garage below is written
out from its published topology (the station order is verified against the
reference's GMM pickles).

What the kernels need is much smaller than the 8x54 complex matrix: stations whose columns of
``A~ = constraint_matrix * exp(j*phase)`` are identical form a *class*, and every constraint
current is a complex combination of the per-class pilot sums.  ``class_reduce`` derives that
form for any network (5 classes for Caltech).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

EVSE_AV = 0   # AeroVironment: allowable pilots {0} U {6..32} A  (min_pilot_signals == 6, env.py:373)
EVSE_CC = 1   # ClipperCreek:  allowable pilots {0, 8, 16, 24, 32} A


@dataclass(frozen=True)
class ChargingNetworkSpec:
    site: str
    station_ids: tuple
    phase_deg: np.ndarray          # [n]
    evse_type: np.ndarray          # [n] uint8 (EVSE_AV / EVSE_CC)
    constraint_matrix: np.ndarray  # [J, n] real coefficients
    magnitudes: np.ndarray         # [J] limits in A
    constraint_names: tuple

    @property
    def num_stations(self) -> int:
        return len(self.station_ids)

    @property
    def min_pilot_signals(self) -> np.ndarray:
        return np.where(self.evse_type == EVSE_AV, 6, 8)

    def class_reduce(self):
        """Returns (station_class [n] u8, coef [J, C] complex128) with
        ``A~ @ s == coef @ class_sums(s)`` for every schedule ``s``."""
        At = self.constraint_matrix * np.exp(1j * np.deg2rad(self.phase_deg))[None, :]
        cols = np.round(np.stack([At.real, At.imag]), 12)
        keys, station_class, coef = {}, np.zeros(self.num_stations, np.uint8), []
        for i in range(self.num_stations):
            key = cols[:, :, i].tobytes()
            if key not in keys:
                keys[key] = len(keys)
                coef.append(At[:, i])
            station_class[i] = keys[key]
        return station_class, np.stack(coef, axis=1)


def _ids(nums):
    return [f"CA-{i}" for i in nums]


def caltech_acn(voltage: float = 208.0, transformer_cap: float = 150.0) -> ChargingNetworkSpec:
    """The 54-EVSE Caltech garage: three line-to-line phase groups, two 80 A pods on phase AB,
    a 150 kVA delta-wye transformer constrained on both its secondary (120 V) and primary
    (277 V) sides."""
    pods = {"CC Pod": _ids([322, 493, 496, 320, 495, 321, 323, 494]),
            "AV Pod": _ids([324, 325, 326, 327, 489, 490, 491, 492])}
    groups = [  # (name, phase angle in degrees, members)
        ("AB", 30.0, _ids([308, 508, 303, 513, 310, 506, 316, 500, 318, 498]) + pods["AV Pod"] + pods["CC Pod"]),
        ("BC", -90.0, _ids([304, 512, 305, 511, 313, 503, 311, 505, 317, 499, 148, 149, 212, 213])),
        ("CA", 150.0, _ids([307, 509, 309, 507, 306, 510, 315, 501, 319, 497, 312, 504, 314, 502])),
    ]
    station_ids = tuple(s for _, _, members in groups for s in members)
    n = len(station_ids)
    phase = np.concatenate([np.full(len(m), ang) for _, ang, m in groups])
    evse = np.array([EVSE_CC if s in pods["CC Pod"] else EVSE_AV for s in station_ids], np.uint8)
    ind = {name: np.array([1.0 if s in set(m) else 0.0 for s in station_ids]) for name, _, m in groups}
    sec = {"a": ind["AB"] - ind["CA"], "b": ind["BC"] - ind["AB"], "c": ind["CA"] - ind["BC"]}
    pri = {"a": 0.25 * (sec["a"] - sec["c"]), "b": 0.25 * (sec["b"] - sec["a"]), "c": 0.25 * (sec["c"] - sec["b"])}
    sec_lim = transformer_cap * 1000 / 3 / 120
    pri_lim = transformer_cap * 1000 / 3 / 277
    rows, lims, names = [], [], []
    for ph in "abc":
        rows.append(sec[ph]); lims.append(sec_lim); names.append(f"Secondary {ph.upper()}")
    for ph in "abc":
        rows.append(pri[ph]); lims.append(pri_lim); names.append(f"Primary {ph.upper()}")
    for pod, members in pods.items():
        rows.append(np.array([1.0 if s in set(members) else 0.0 for s in station_ids]))
        lims.append(80.0); names.append(pod)
    assert n == 54
    return ChargingNetworkSpec("caltech", station_ids, phase, evse, np.stack(rows), np.array(lims), tuple(names))


def site_str_to_site(site: str) -> ChargingNetworkSpec:
    """Mirror of ``sustaingym/envs/evcharging/utils.py:83-88``."""
    if site == "caltech":
        return caltech_acn()
    if site == "jpl":
        raise NotImplementedError(
            "the JPL garage (acnportal jpl_acn, 52 EVSEs) is not in the reference tree and is not "
            "reconstructed here; pass a ChargingNetworkSpec to use another network")
    raise ValueError(f"unknown site {site!r}")
