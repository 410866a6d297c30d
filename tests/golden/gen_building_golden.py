"""Generates ``tests/golden/building_*.npz`` by running the UNMODIFIED reference BuildingEnv.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/gen_building_golden.py

Each fixture holds the reference's own A/B/D-derived ``A_d``/``BD_d`` (env.py:196-203), a real
weather window cut from the Tucson EPW tables the reference's ParameterGenerator produced
(utils.py:606-630, truncated to T = 2000 rows so the cursor wrap at ``len(weather)``
env.py:306-307 is exercised), the action stream, and the (obs, reward, done) stream the
reference's ``step()`` returned.  NumPy version is recorded: scalar promotion differs between
NumPy 1.x and 2.x.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402

WINDOW = (1200, 3200)   # rows of the 5-min tables kept in the fixture
STEPS = 300             # > episode_len so `done` flips inside the trajectory


def run_case(name, building, seed, *, discrete=False, steps=STEPS, **pg_kwargs):
    ref_shim.install()
    import contextlib
    import io

    from sustaingym.envs.building.env import BuildingEnv
    from sustaingym.envs.building.utils import ParameterGenerator

    with contextlib.redirect_stdout(io.StringIO()):
        params = ParameterGenerator(building=building, weather="Hot_Dry", location="Tucson",
                                    is_continuous_action=not discrete, **pg_kwargs)
    for k in ("out_temp", "ground_temp", "ghi", "metabolism"):
        params[k] = np.ascontiguousarray(params[k][WINDOW[0]:WINDOW[1]])
    env = BuildingEnv(params)
    n = env.n
    obs0, _ = env.reset(seed=seed)
    epoch0 = int(env.epoch)
    rng = np.random.default_rng(1234 + seed)
    actions, obs, rewards, dones = [], [], [], []
    for _ in range(steps):
        if discrete:
            hi = (env.Qhigh * env.DISCRETE_LENGTH - env.Qlow * env.DISCRETE_LENGTH).astype(np.int64)
            a = rng.integers(0, np.maximum(hi, 1)).astype(np.int64)
        else:
            a = (rng.uniform(-1, 1, n) * env.ac_map).astype(np.float32)
        o, r, d, tr, _info = env.step(a)
        assert d == tr
        actions.append(a); obs.append(o.copy()); rewards.append(float(r)); dones.append(bool(d))
    out = dict(
        n=n, A_d=env.A_d, BD_d=env.BD_d, target=np.asarray(params["target"], float),
        ac_map=np.asarray(params["ac_map"], float), out_temp=params["out_temp"],
        ground_temp=params["ground_temp"], ghi=params["ghi"], metabolism=params["metabolism"],
        q_rate=env.q_rate, error_rate=env.error_rate, pnorm=float(params["reward_pnorm"]),
        episode_len=env.episode_len, time_res=env.timestep, max_power=float(np.max(params["max_power"])),
        is_continuous_action=not discrete, seed=seed, epoch0=epoch0, obs0=obs0,
        actions=np.stack(actions), obs=np.stack(obs), rewards=np.array(rewards, np.float64),
        dones=np.array(dones), numpy_version=np.__version__,
        A=params["A"], B=params["B"], D=params["D"],
    )
    path = os.path.join(HERE, f"building_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: n={n} epoch0={epoch0} T={len(params['out_temp'])} "
          f"reward[0]={rewards[0]!r} -> {os.path.relpath(path)} ({os.path.getsize(path)} B)")


if __name__ == "__main__":
    only = sys.argv[1:]
    if only:
        _run_case = run_case
        run_case = lambda name, *a, **k: _run_case(name, *a, **k) if name in only else None  # noqa: E731
    # seed -> epoch = int(seed/365 * 2000); 5 -> 27
    run_case("officesmall_s5", "OfficeSmall", 5)
    run_case("retail_s5", "RetailStandalone", 5)
    # cursor wrap: seed 330 -> epoch 1808; wraps to 0 after 192 steps
    run_case("officesmall_wrap", "OfficeSmall", 330)
    # D != 0 (occupancy feeds the dynamics), partial AC map, per-zone targets, p-norm 2
    run_case("officesmall_occ", "OfficeSmall", 40, full_occ=4.0,
             ac_map=np.array([1, 1, 0, 1, 1, 0]), target=np.array([21., 22., 23., 22., 21., 20.]),
             activity_sch=140.0, reward_beta=0.9)
    # MultiDiscrete actions (env.py:234-235)
    run_case("retail_discrete", "RetailStandalone", 77, discrete=True)
    # a mid-size building (n = 18) to exercise the 8-accumulator f32 pairwise sum of avg_temp
    run_case("officemedium_s9", "OfficeMedium", 9, steps=120)
    # the generic-n kernel path on the reference's largest / smallest buildings (every
    # shipped building runs on the engine): n = 118, 90, 55 and 3
    run_case("outpatient_s3", "OutPatientHealthCare", 3, steps=100)
    run_case("apartmenthighrise_s21", "ApartmentHighRise", 21, steps=100)
    run_case("hospital_s6", "Hospital", 6, steps=100)
    run_case("restaurantfastfood_s2", "RestaurantFastFood", 2, steps=100)
