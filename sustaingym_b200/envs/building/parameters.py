"""``ParameterGenerator`` for BuildingEnv without the reference tree.

Restates ``sustaingym/envs/building/utils.py``: ``get_zones`` (:179-291, the "Zone Information" table of an EnergyPlus
HTM report), ``Nfind_neighbor`` (:308-428, zone adjacency -> thermal conductances / capacities), the weather
interpolation and the ``A`` / ``B`` / ``D`` construction of ``ParameterGenerator`` (:462-779).  It is an offline CPU tool —
its outputs (``A_d``, ``BD_d`` after discretisation, the four 5-minute weather tables) are the constants of the building
kernels — pinned bit for bit to the reference's generator run under the import shim, for all 16 shipped buildings
(``tests/test_building_parameters.py``).

Inputs come from one of two places:
  * ``data_dir`` (or ``$SUSTAINGYM_DATA_DIR``): the reference's ``sustaingym/data`` tree — the raw ``*.table.htm`` and
    ``*.epw`` files are parsed here;
  * otherwise the packaged archive ``sustaingym_b200/data/building_models.npz`` (``tools/export_building_models.py``): the
    zone tables of the 16 prototype buildings, their U-factors, the monthly ground temperatures of the 19 locations and
    the hourly (air temperature, GHI) series of the 16 climates, so that ``BuildingEnv("OfficeSmall")`` works standalone.

Not restated: ``stochastic_summer_percentage`` (block-MVN resampling of the weather, ``stochastic_generator.py``).
"""
from __future__ import annotations

import io
import os
import re
from typing import Any, NamedTuple, Sequence

import numpy as np

_ARCHIVE = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "data", "building_models.npz")
# columns of the "Zone Information" table that the model uses, in the order of ``Zone`` below.  The reference picks them by
# line offset (utils.py:213-252): Zone Name, Centroid Z (its "Zaxis"), Minimum/Maximum X, Y, Z, Floor Area, Exterior NET
# Wall Area (which it calls ExteriorGrossArea) and Exterior Window Area.
ZONE_COLUMNS = ("Centroid Z-Coordinate {m}", "Minimum X {m}", "Maximum X {m}", "Minimum Y {m}", "Maximum Y {m}",
                "Minimum Z {m}", "Maximum Z {m}", "Floor Area {m2}", "Exterior Net Wall Area {m2}", "Exterior Window Area {m2}")
OCCU_COEF = 7.139322          # W/K, utils.py:657 (coefficient c4 of the BEAR paper)
AIR_DENSITY = 1.225           # kg/m^3, utils.py:335
SPECIFIC_HEAT_AIR = 1000      # J/kg-K, utils.py:611
DAYS_PER_MONTH = (31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31)


class Zone(NamedTuple):
    name: str
    Zaxis: float
    Xmin: float
    Xmax: float
    Ymin: float
    Ymax: float
    Zmin: float
    Zmax: float
    FloorArea: float
    ExteriorGrossArea: float      # (the table's NET wall area, see ZONE_COLUMNS)
    ExteriorWindowArea: float
    ind: int


class Ufactor(NamedTuple):
    intwall: float
    floor: float
    outwall: float
    roof: float
    ceiling: float
    groundfloor: float
    window: float


# ---- parsing ----------------------------------------------------------------------------------------------------------
_CELL = re.compile(r"<td[^>]*>(.*?)</td>", re.S)


def parse_zone_table(text: str) -> tuple[list[str], np.ndarray]:
    """(zone names, [n, 10] float64 columns ZONE_COLUMNS) of the HTM report's "Zone Information" table, in file order."""
    start = text.rfind("Zone Information")        # the heading and the FullName comment both carry the words: take the table after the last
    if start < 0:
        raise ValueError("no 'Zone Information' table in the building report")
    end = text.find("Zone Internal Gains Nominal", start)
    table = text[start:end if end > 0 else len(text)]
    table = table[:table.find("</table>")] if "</table>" in table else table
    rows = [[c.strip() for c in _CELL.findall(r)] for r in table.split("<tr>")[1:]]
    header, body = rows[0], [r for r in rows[1:] if len(r) == len(rows[0])]
    col = {name: header.index(name) for name in ("Zone Name",) + ZONE_COLUMNS}
    names = [r[col["Zone Name"]] for r in body]
    num = np.array([[float(r[col[c]]) for c in ZONE_COLUMNS] for r in body], np.float64)
    return names, num


def zones_from_table(names: Sequence[str], num: np.ndarray) -> tuple[list[list[Zone]], int, list[Zone]]:
    """``get_zones`` after the parsing: zones sorted (stably) by their centroid height, indexed, grouped into layers."""
    order = sorted(range(len(names)), key=lambda i: num[i, 0])
    all_zones = [Zone(str(names[i]), *(float(v) for v in num[i]), ind=k) for k, i in enumerate(order)]
    layers: list[list[Zone]] = []
    for z in all_zones:
        if layers and layers[-1][0].Zaxis == z.Zaxis:
            layers[-1].append(z)
        else:
            layers.append([z])
    return layers, len(all_zones), all_zones


def get_zones(path_or_file) -> tuple[list[list[Zone]], int, list[Zone]]:
    """utils.py:179-291 for a path or a text file object."""
    text = path_or_file.read() if isinstance(path_or_file, io.TextIOBase) else open(path_or_file, "r").read()
    return zones_from_table(*parse_zone_table(text))


def parse_epw(path: str) -> tuple[np.ndarray, np.ndarray]:
    """Hourly (dry-bulb temperature in C, global horizontal irradiance in Wh/m2) of an EnergyPlus weather file: columns
    6 and 13 of the data records after the 8 header lines (what ``pvlib.iotools.parse_epw`` names temp_air / ghi)."""
    temp, ghi = [], []
    with open(path, "r", errors="replace") as f:
        for i, line in enumerate(f):
            if i < 8 or not line.strip():
                continue
            parts = line.split(",")
            temp.append(float(parts[6])); ghi.append(float(parts[13]))
    return np.array(temp, np.float64), np.array(ghi, np.float64)


# ---- geometry -> RC network (utils.py:294-428) ---------------------------------------------------------------------------
def _inside(z1: Zone, z2: Zone, strict: bool) -> bool:
    """One corner of z1's footprint lies in z2's footprint (``checkconnect`` / ``checkconnect_layer``)."""
    if strict:      # zones of adjacent layers: half-open boxes
        lo = z2.Xmin <= z1.Xmin < z2.Xmax and z2.Ymin <= z1.Ymin < z2.Ymax
        hi = z2.Xmin < z1.Xmax <= z2.Xmax and z2.Ymin < z1.Ymax <= z2.Ymax
    else:
        lo = z2.Xmin <= z1.Xmin <= z2.Xmax and z2.Ymin <= z1.Ymin <= z2.Ymax
        hi = z2.Xmin <= z1.Xmax <= z2.Xmax and z2.Ymin <= z1.Ymax <= z2.Ymax
    return lo or hi


def rc_network(n: int, layers: Sequence[Sequence[Zone]], u: Ufactor, specific_heat: float = SPECIFIC_HEAT_AIR):
    """``Nfind_neighbor``: (neighbours per zone name, conductances [n, n+1] in W/K with the outside in the last column,
    heat capacities [n] in J/K, exterior window areas [n] in m2).  The reference's overlap expressions are kept as they are
    (its y-overlap subtracts ``max(z1.Ymin, z1.Ymin)``)."""
    R, C, W = np.zeros((n, n + 1)), np.zeros(n), np.zeros(n)
    neighbors: dict[str, list[int]] = {}
    link = lambda a, b: (neighbors.setdefault(a.name, []).append(b.ind), neighbors.setdefault(b.name, []).append(a.ind))  # noqa: E731
    for k, layer in enumerate(layers):
        if k + 1 < len(layers):                       # floor / ceiling in series between stacked zones
            for z1 in layer:
                for z2 in layers[k + 1]:
                    if _inside(z1, z2, True) or _inside(z2, z1, True):
                        x_overlap = min(z1.Xmax, z2.Xmax) - max(z1.Xmin, z2.Xmin)
                        y_overlap = min(z1.Ymax, z2.Ymax) - max(z1.Ymin, z1.Ymin)
                        U = (x_overlap * y_overlap) * (u.floor * u.ceiling / (u.floor + u.ceiling))
                        R[z2.ind, z1.ind] = R[z1.ind, z2.ind] = U
                        link(z1, z2)
        for i, z1 in enumerate(layer):
            height, xlen, ylen = z1.Zmax - z1.Zmin, z1.Xmax - z1.Xmin, z1.Ymax - z1.Ymin
            C[z1.ind] = specific_heat * height * xlen * ylen * AIR_DENSITY
            W[z1.ind] = z1.ExteriorWindowArea
            last = i == len(layer) - 1                 # the last zone of a layer also gets the roof term
            if z1.ExteriorGrossArea > 0 or last:
                if last:
                    R[z1.ind, -1] = z1.ExteriorGrossArea * u.outwall + xlen * ylen * u.roof + z1.ExteriorWindowArea * u.window
                else:
                    R[z1.ind, -1] = z1.ExteriorGrossArea * u.outwall + z1.ExteriorWindowArea * u.window
                neighbors.setdefault(z1.name, []).append(n)
            for z2 in layer[i + 1:]:                    # interior walls within the layer
                if _inside(z1, z2, False) or _inside(z2, z1, False):
                    x_overlap = min(z1.Xmax, z2.Xmax) - max(z1.Xmin, z2.Xmin)
                    y_overlap = min(z1.Ymax, z2.Ymax) - max(z1.Ymin, z1.Ymin)
                    U = height * np.sqrt(x_overlap ** 2 + y_overlap ** 2) * u.intwall
                    R[z2.ind, z1.ind] = R[z1.ind, z2.ind] = U
                    link(z1, z2)
    return neighbors, R, C, W


# ---- tables ------------------------------------------------------------------------------------------------------------------
class _Archive:
    _z = None

    @classmethod
    def get(cls):
        if cls._z is None:
            if not os.path.exists(_ARCHIVE):
                raise FileNotFoundError(f"{_ARCHIVE} is missing (tools/export_building_models.py writes it) and no data_dir was given")
            cls._z = np.load(_ARCHIVE, allow_pickle=False)
        return cls._z


def building_names() -> list[str]:
    return [str(b) for b in _Archive.get()["buildings"]]


def _resample(hourly: np.ndarray, time_res: int) -> np.ndarray:
    """Linear interpolation of an hourly series to the step length (utils.py:578-603)."""
    from scipy import interpolate
    x = np.arange(len(hourly))
    return interpolate.interp1d(x, np.array(hourly))(np.arange(0, len(hourly) - 1, 1 / 3600 * time_res))


def ParameterGenerator(building: str, weather: str, location: str, U_Wall: Sequence[float] = (0,) * 7,
                       ground_temp: Sequence[float] = (0,) * 12, shgc: float = 0.252, shgc_weight: float = 0.01,
                       ground_weight: float = 0.5, full_occ: np.ndarray | float = 0, max_power: np.ndarray | int = 8000,
                       ac_map: np.ndarray | int = 1, time_res: int = 300, reward_beta: float = 0.999,
                       reward_pnorm: float = 2, target: np.ndarray | float = 22, activity_sch: np.ndarray | float = 120,
                       temp_range: tuple[float, float] = (-40, 40), is_continuous_action: bool = True, root: str = "",
                       stochastic_summer_percentage: float | None = None, episode_len: int = 288,
                       block_size: int | None = None, data_dir: str | None = None, verbose: bool = False) -> dict[str, Any]:
    """The reference's signature (utils.py:462-485) plus ``data_dir`` (raw HTM / EPW files instead of the packaged archive) and
    ``verbose`` (the reference prints the zone list; off by default here).  ``building`` / ``weather`` that are not names of
    the packaged tables are paths relative to ``root`` to an HTM report / an EPW file, with ``U_Wall`` / ``ground_temp``."""
    if episode_len * time_res % (24 * 60 * 60) != 0:
        raise ValueError("Episode must be a multiple of 1 day")
    if stochastic_summer_percentage is not None:
        raise NotImplementedError("stochastic weather (stochastic_generator.py) is not part of this package")
    data_dir = data_dir or os.environ.get("SUSTAINGYM_DATA_DIR")
    z = _Archive.get() if os.path.exists(_ARCHIVE) else None
    known_building = z is not None and building in set(str(b) for b in z["buildings"])
    known_weather = z is not None and f"weather_file|{weather}" in z
    # ---- zones ----
    if known_building:
        ufac = Ufactor(*(float(v) for v in z[f"ufactor|{building}"]))
        if data_dir:
            layers, n, all_zones = get_zones(os.path.join(data_dir, "building", str(z[f"htm_file|{building}"])))
        else:
            layers, n, all_zones = zones_from_table([str(s) for s in z[f"zone_names|{building}"]], z[f"zone_table|{building}"])
    else:
        ufac = Ufactor(*U_Wall)
        layers, n, all_zones = get_zones(os.path.join(root, building))
    if verbose:
        for zone in all_zones:
            print(zone.name, " [Zone index]: ", zone.ind)
    # ---- weather ----
    if known_weather:
        if data_dir:
            temp_h, ghi_h = parse_epw(os.path.join(data_dir, "building", str(z[f"weather_file|{weather}"])))
        else:
            temp_h = z[f"weather_temp10|{weather}"].astype(np.float64) / 10.0       # (exactly the parsed decimals)
            ghi_h = z[f"weather_ghi|{weather}"].astype(np.float64)
    else:
        temp_h, ghi_h = parse_epw(os.path.join(root, weather))
    monthly = [float(v) for v in z[f"ground_temp|{location}"]] if (z is not None and f"ground_temp|{location}" in z) else list(ground_temp)
    ground_h = np.concatenate([np.ones(d * 24) * monthly[m] for m, d in enumerate(DAYS_PER_MONTH)])
    all_ground_temp = _resample(ground_h, time_res)
    out_temp = _resample(temp_h, time_res)
    solar = _resample(ghi_h, time_res)
    ghi_max = max(ghi_h)
    SHGC = shgc * shgc_weight * (ghi_max / (1 / 3600 * time_res))                   # GHI from Wh to W
    # ---- RC network -> A, B, D (utils.py:611-670, 701-779) ----
    neighbors, Rtable, Ctable, Windowtable = rc_network(n, layers, ufac, SPECIFIC_HEAT_AIR)
    RCtable = Rtable / np.array([Ctable]).T
    connectmap = np.zeros((n, n + 1))
    for i, zone in enumerate(all_zones):
        connectmap[i, neighbors.get(zone.name, [])] = 1
    ground_connect = np.zeros((n, 1))
    for room in layers[0]:                                 # the lowest layer stands on the ground
        ground_connect[room.ind] = room.FloorArea * ufac.groundfloor * ground_weight
    people_full = (np.zeros(n) + full_occ).reshape(n, 1)
    ACweight = np.diag(np.zeros(n) + ac_map) * max_power
    weightcmap = np.concatenate((people_full, ground_connect, np.zeros((n, 1)), ACweight, (Windowtable * SHGC).reshape(n, 1)),
                                axis=-1) / Ctable[:, None]
    A = RCtable[:, :-1].copy()
    np.fill_diagonal(A, -np.diag(RCtable @ connectmap.T) - weightcmap[:, 1])
    A += weightcmap[:, 0] * OCCU_COEF / n
    BD = weightcmap.copy()
    BD[:, 2] = connectmap[:, -1] * RCtable[:, -1]
    return {
        "n": n, "zones": all_zones, "target": np.zeros(n) + target, "out_temp": out_temp, "ground_temp": all_ground_temp,
        "ghi": solar / (1 / 3600 * time_res) / (ghi_max / (1 / 3600 * time_res)),
        "metabolism": activity_sch * np.ones(len(out_temp)), "reward_beta": reward_beta, "reward_pnorm": reward_pnorm,
        "ac_map": np.zeros(n) + ac_map, "max_power": max_power, "temp_range": temp_range,
        "is_continuous_action": is_continuous_action, "time_resolution": time_res, "A": A, "B": BD[:, 1:], "D": BD[:, 0],
        "episode_len": episode_len,
    }
