"""CPU: the Cogen oracle's plant model reproduces the reference ONNX graph (golden fixture generated
from the reference's model.onnx by tests/golden/gen_cogen_pack.py), incl. the SURVEY.md D.2 vector."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.cogen import CogenOracle, PlantModel, flatten_obs

D2_Y = [52.027393, 53.164238, 52.73687, 4.4654775, 11.712152, 1.9682132, 56.49287, 64.87639, 54.705082, 130.90273,
        138.19304, 66.16172, 135.78653, 138.5116, 146.0293, 433.47562, 755.5165, 401.17264, 711.41895, 512.43866,
        845.78284, 176.07434, -375.72876, -385.58398, 91.69225, 121.10405, 13.481275, 360.4708, 1104.8684]


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "cogen_model_io.npz"))


def test_plant_model_matches_onnx_graph(golden):
    pm = PlantModel()
    Y = pm(golden["X"], np.float32)
    np.testing.assert_array_equal(Y, golden["Y32"])           # same op order -> bit-exact
    np.testing.assert_allclose(Y[0], D2_Y, rtol=5e-7)         # the vector quoted in SURVEY.md D.2
    Y64 = pm(golden["X"], np.float64)
    np.testing.assert_allclose(Y64, golden["Y64"], rtol=1e-12, atol=1e-12)


def test_plant_model_against_reference_file_when_present(golden):
    ref = "/root/reference/sustaingym/data/cogen/onnx_model/model.onnx"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present (GPU box)")
    from oracle import onnx_reader as R
    g = R.load(ref)
    with np.errstate(over="ignore"):
        Y = R.run(g, golden["X"][:64], np.float32)
    np.testing.assert_array_equal(Y, golden["Y32"][:64])


def test_gates_switch_units_off(golden):
    pm = PlantModel()
    X = golden["X"][1:9].copy()
    X[:, 5] = 0                                                # GT1 off
    Y = pm(X)
    assert np.all(Y[:, 0] == 0)                                # GT1_NG_M gated by g(GT1_PWR)
    X[:, 12] = 0
    assert np.all(pm(X)[:, 3] == 0)                            # HR1_DBNG_M gated by g(gGT1 + gHR1)


def test_step_identities(golden):
    rng = np.random.default_rng(0)
    amb = rng.uniform(0, 1, (4, 96, 7)) * np.array([80, 1, 1, 700, 1300, 100, 7]) + np.array([32, 14, 0, 0, 0, 0, 0])
    B = 8
    o = CogenOracle(amb, B, forecast_horizon=12)
    acts = golden["X"][1:1 + B, 3:]
    obs = o.reset(np.arange(B) % 3, acts)
    assert flatten_obs(obs).shape == (B, 124)
    for t in range(96):
        obs, r, term, trunc, info = o.step(golden["X"][1 + t:1 + t + B, 3:])
        np.testing.assert_allclose(info["agent_rewards"].sum(1), r, rtol=1e-6)
        assert np.all(term == (t == 95)) and not trunc.any()
        f = flatten_obs(obs)
        assert np.allclose(f[:, -1], (t + 1) / 96)
        np.testing.assert_array_equal(obs["Prev_Action"], golden["X"][1 + t:1 + t + B, 3:])
    # same action twice -> no ramp cost
    o2 = CogenOracle(amb, 1, forecast_horizon=3)
    a = golden["X"][5:6, 3:]
    o2.reset(np.array([0]), a)
    _, _, _, _, info = o2.step(a)
    assert np.all(info["ramp"] == 0)
    with pytest.raises(IndexError):                            # last day has no next day to spill into
        CogenOracle(amb, 1, forecast_horizon=3).reset(np.array([3]), a) and None
        o3 = CogenOracle(amb, 1, forecast_horizon=3); o3.reset(np.array([3]), a)
        for _ in range(96):
            o3.step(a)
