"""Offline tool: runs the reference's ParameterGenerator (sustaingym/envs/building/utils.py:462-779) for the
shipped buildings / climates and writes one ``.npz`` per setting that ``load_parameters()``
(sustaingym_b200/envs/building/params.py) turns back into the ``parameters`` dict BuildingEnv takes —
so every shipped building runs on the engine where the reference tree is not present (the GPU box).

    python tests/golden/export_building_params.py OUT_DIR [--buildings OfficeSmall,Hospital] [--weather Hot_Dry] [--location Tucson]

Needs /root/reference (imported under oracle/ref_shim.py: gymnasium / pvlib stubs).  Without --buildings all 16
ASHRAE 90.1 prototypes are exported (n = 3 ... 118 zones).  Weather tables are stored as float32 (they are
EPW values with one decimal) to keep a file near 2 MB.
"""
import argparse
import contextlib
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

BUILDINGS = ["ApartmentHighRise", "ApartmentMidRise", "Hospital", "HotelLarge", "HotelSmall", "OfficeLarge",
             "OfficeMedium", "OfficeSmall", "OutPatientHealthCare", "RestaurantFastFood", "RestaurantSitDown",
             "RetailStandalone", "RetailStripmall", "SchoolPrimary", "SchoolSecondary", "Warehouse"]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("out_dir")
    ap.add_argument("--buildings", default=",".join(BUILDINGS))
    ap.add_argument("--weather", default="Hot_Dry")
    ap.add_argument("--location", default="Tucson")
    args = ap.parse_args()
    from oracle import ref_shim
    ref_shim.install()
    from sustaingym.envs.building.utils import ParameterGenerator

    from sustaingym_b200.envs.building.params import save_parameters
    os.makedirs(args.out_dir, exist_ok=True)
    for b in args.buildings.split(","):
        with contextlib.redirect_stdout(io.StringIO()):
            p = ParameterGenerator(building=b, weather=args.weather, location=args.location)
        path = os.path.join(args.out_dir, f"{b}_{args.weather}_{args.location}.npz")
        save_parameters(path, p)
        print(f"{b}: n={p['n']} T={len(p['out_temp'])} -> {path} ({os.path.getsize(path)} B)")


if __name__ == "__main__":
    main()
