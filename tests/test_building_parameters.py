"""CPU: ``sustaingym_b200.envs.building.parameters.ParameterGenerator`` (restatement of the reference's
``sustaingym/envs/building/utils.py:179-291, 308-428, 462-779``) against

  * the reference's own generator run under the import shim (when ``/root/reference`` is present): all 16 shipped
    buildings, A / B / D, the 5-minute weather tables, zone names and order — bit for bit, from the packaged archive AND
    from the raw HTM / EPW files; other climates / locations / keyword arguments on a few buildings;
  * the committed reference fixtures (``tests/golden/building_*.npz``, produced by the reference's generator), anywhere.
"""
import os

import numpy as np
import pytest

from helpers import building_fixtures

REF = "/root/reference"
HAVE_REF = os.path.isdir(os.path.join(REF, "sustaingym"))
GOLDEN_WINDOW = (1200, 3200)      # tests/golden/gen_building_golden.py WINDOW
KEYS = ("A", "B", "D", "out_temp", "ground_temp", "ghi", "metabolism", "target", "ac_map")


def same(p, q):
    assert p["n"] == q["n"] and [z.name for z in p["zones"]] == [z.name for z in q["zones"]]
    for k in KEYS:
        np.testing.assert_array_equal(np.asarray(p[k]), np.asarray(q[k]), err_msg=k)
    for k in ("reward_beta", "reward_pnorm", "max_power", "time_resolution", "episode_len", "is_continuous_action"):
        assert p[k] == q[k], k
    assert tuple(p["temp_range"]) == tuple(q["temp_range"])


@pytest.fixture(scope="module")
def ref_generator():
    if not HAVE_REF:
        pytest.skip("reference tree not present")
    import contextlib
    import io
    from oracle import ref_shim
    ref_shim.install()
    from sustaingym.envs.building.utils import ParameterGenerator

    def gen(*a, **kw):
        with contextlib.redirect_stdout(io.StringIO()):      # (the reference prints the zone list)
            return ParameterGenerator(*a, **kw)
    return gen


def test_all_16_buildings_match_the_reference_generator(ref_generator):
    from sustaingym_b200.envs.building.parameters import ParameterGenerator, building_names
    names = building_names()
    assert len(names) == 16
    sizes = {}
    for b in names:
        ref = ref_generator(b, "Hot_Dry", "Tucson")
        same(ParameterGenerator(b, "Hot_Dry", "Tucson"), ref)                                        # packaged archive
        same(ParameterGenerator(b, "Hot_Dry", "Tucson", data_dir=os.path.join(REF, "sustaingym", "data")), ref)   # raw files
        sizes[b] = ref["n"]
    assert sizes["OfficeSmall"] == 6 and sizes["RetailStandalone"] == 5 and sizes["OutPatientHealthCare"] == 118


@pytest.mark.parametrize("building,weather,location,kw", [
    ("OfficeSmall", "Cold_Humid", "Rochester", {}),
    ("RetailStandalone", "Very_Hot_Humid", "Honolulu", dict(time_res=900, episode_len=96, reward_beta=0.9, target=21.5)),
    ("Hospital", "Subarctic/Arctic", "Fairbanks", dict(full_occ=3.0, max_power=6000, shgc=0.3, ground_weight=0.7)),
    ("Warehouse", "Mixed_Marine", "Seattle", dict(ac_map=np.array([1, 0, 1]), is_continuous_action=False, activity_sch=90.0)),
])
def test_other_climates_and_arguments_match(ref_generator, building, weather, location, kw):
    from sustaingym_b200.envs.building.parameters import ParameterGenerator
    same(ParameterGenerator(building, weather, location, **kw), ref_generator(building, weather, location, **kw))


def test_packaged_archive_reproduces_the_committed_reference_fixtures():
    """Anywhere (no reference tree): the fixtures under tests/golden were written by the reference's generator."""
    from sustaingym_b200.envs.building.env import discretize
    from sustaingym_b200.envs.building.parameters import ParameterGenerator
    fixtures = {"officesmall_s5": "OfficeSmall", "retail_s5": "RetailStandalone", "officemedium_s9": "OfficeMedium",
                "hospital_s6": "Hospital", "outpatient_s3": "OutPatientHealthCare", "apartmenthighrise_s21": "ApartmentHighRise"}
    seen = 0
    for path in building_fixtures():
        tag = os.path.basename(path)[len("building_"):-4]
        if tag not in fixtures:
            continue
        g = np.load(path)
        p = ParameterGenerator(fixtures[tag], "Hot_Dry", "Tucson")
        for k in ("A", "B", "D"):
            np.testing.assert_array_equal(p[k], g[k], err_msg=f"{tag} {k}")
        lo, hi = GOLDEN_WINDOW                              # the fixtures keep a 2 000-row window of the weather tables
        np.testing.assert_array_equal(p["out_temp"][lo:hi], g["out_temp"]); np.testing.assert_array_equal(p["ghi"][lo:hi], g["ghi"])
        np.testing.assert_array_equal(p["ground_temp"][lo:hi], g["ground_temp"])
        A_d, BD_d = discretize(p["A"], p["B"], p["D"], p["time_resolution"])
        np.testing.assert_array_equal(A_d, g["A_d"]); np.testing.assert_array_equal(BD_d, g["BD_d"])
        seen += 1
    assert seen >= 5


def test_stochastic_weather_is_declined():
    from sustaingym_b200.envs.building.parameters import ParameterGenerator
    with pytest.raises(NotImplementedError):
        ParameterGenerator("OfficeSmall", "Hot_Dry", "Tucson", stochastic_summer_percentage=0.5)
    with pytest.raises(ValueError):
        ParameterGenerator("OfficeSmall", "Hot_Dry", "Tucson", episode_len=100)
