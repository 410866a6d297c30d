"""Packs what ``sustaingym_b200.envs.building.parameters.ParameterGenerator`` needs to run without the reference tree:

    python tools/export_building_models.py /root/reference sustaingym_b200/data/building_models.npz

* the "Zone Information" table of the 16 prototype-building HTM reports (parsed by OUR parser), with the HTM file names and
  the U-factor tuples of the reference's ``BUILDINGS`` table (sustaingym/envs/building/utils.py:57-120);
* the monthly ground temperatures of ``GROUND_TEMP`` (:122-142);
* per climate of ``WEATHER`` (:144-161) the EPW file name and the hourly dry-bulb temperature (stored x10 as int16: the
  files carry one decimal) and global horizontal irradiance (int32) — checked here to reconstruct the parsed values exactly.

The three tables are read from the reference module itself (imported under oracle/ref_shim.py), not retyped.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(ref_root: str, dst: str) -> None:
    os.environ["SUSTAINGYM_REFERENCE_ROOT"] = ref_root
    from oracle import ref_shim
    ref_shim.install()
    from sustaingym.envs.building import utils as R
    from sustaingym_b200.envs.building.parameters import parse_epw, parse_zone_table
    data = os.path.join(ref_root, "sustaingym", "data", "building")
    out = {"buildings": np.array(sorted(R.BUILDINGS))}
    for name, (htm, ufac) in R.BUILDINGS.items():
        names, num = parse_zone_table(open(os.path.join(data, htm)).read())
        out[f"htm_file|{name}"] = np.array(htm)
        out[f"ufactor|{name}"] = np.array(tuple(ufac), np.float64)
        out[f"zone_names|{name}"] = np.array(names)
        out[f"zone_table|{name}"] = num
        print(f"{name:22s} {len(names):4d} zones")
    for loc, temps in R.GROUND_TEMP.items():
        out[f"ground_temp|{loc}"] = np.array(temps, np.float64)
    for climate, epw in R.WEATHER.items():
        temp, ghi = parse_epw(os.path.join(data, epw))
        t10, g = np.rint(temp * 10).astype(np.int16), np.rint(ghi).astype(np.int32)
        assert np.array_equal(t10.astype(np.float64) / 10.0, temp) and np.array_equal(g.astype(np.float64), ghi), climate
        out[f"weather_file|{climate}"] = np.array(epw)
        out[f"weather_temp10|{climate}"] = t10
        out[f"weather_ghi|{climate}"] = g
        print(f"{climate:18s} {len(temp)} hours, T in [{temp.min()}, {temp.max()}] C, max GHI {ghi.max()}")
    np.savez_compressed(dst, **out)
    print(dst, os.path.getsize(dst) // 1024, "KB")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference",
         sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "sustaingym_b200", "data", "building_models.npz"))
