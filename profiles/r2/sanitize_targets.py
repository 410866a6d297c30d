"""Small invocations of the three big kernels for compute-sanitizer (racecheck / memcheck), SURVEY.md section 5.
usage: python profiles/r2/sanitize_targets.py ev|building|building_step|market|cogen"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

which = sys.argv[1]
if which == "ev":
    # ev_step_main_kernel<1>: tile CTAs (several tiles per warp: SG_EV_MAX_CTAS=2) + helper CTA, busy garage so that the
    # inline pod projection, the interior-point helper path, arrivals, departures and the auto-reset all run
    os.environ["SG_EV_MAX_CTAS"] = "2"
    from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
    from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator
    gen = SyntheticTraceGenerator("caltech", num_days=4, mean_sessions=90.0, seed=3)
    env = BatchedEVChargingEnv(gen, 300, project_action_in_env=True, auto_reset=True, collect_stats=True)
    env.reset(seed=0)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for t in range(int(os.environ.get("SAN_STEPS", 60))):
        env.step(torch.rand(300, 54, generator=g, device="cuda"))
    env.t[:] = 286                                      # ... and two steps across the end of the day
    for t in range(3):
        env.step(torch.rand(300, 54, generator=g, device="cuda"))
    torch.cuda.synchronize()
    print("ev ok", float(env.stats[0]), "steps,", float(env.stats[6]), "projected")
elif which == "building":
    from helpers import building_fixtures, load_building_fixture
    from sustaingym_b200.envs.building.env import BatchedBuildingEnv
    g_, params = load_building_fixture([p for p in building_fixtures() if "officemedium_s9" in p][0])
    env = BatchedBuildingEnv(dict(params, episode_len=8), 96, auto_reset=True, collect_stats=True)
    env.reset(epochs=np.arange(96))
    acts = torch.rand(12, 96, params["n"], device="cuda") * 2 - 1
    env.rollout(acts, keep_obs=True)
    env.step(acts[0])
    torch.cuda.synchronize()
    print("building ok (building_mma_kernel, n =", params["n"], ")")
elif which == "building_step":
    # building_step1_kernel<5>: persistent CTAs over a two-slot tile ring — capped to 2 CTAs' worth of tiles per CTA and a
    # ragged last tile; plus the one-tile-per-CTA rollout kernel (cp.async start-up, action ring)
    from helpers import building_fixtures, load_building_fixture
    from sustaingym_b200.envs.building.env import BatchedBuildingEnv
    g_, params = load_building_fixture([p for p in building_fixtures() if "retail_s5" in p][0])
    B = 148 * 4 * 128 * 2 + 77                          # every CTA walks over >= 2 tiles, the last tile is partial
    env = BatchedBuildingEnv(dict(params, episode_len=3), B, auto_reset=True, collect_stats=True)
    env.reset(epochs=np.arange(B) % 1000)
    a = torch.rand(B, params["n"], device="cuda") * 2 - 1
    for t in range(4):
        env.step(a)
    small = BatchedBuildingEnv(dict(params, episode_len=3), 200, auto_reset=True)
    small.reset(epochs=np.arange(200))
    small.rollout(torch.rand(5, 200, params["n"], device="cuda") * 2 - 1, keep_obs=True)
    torch.cuda.synchronize()
    print("building_step ok (building_step1_kernel + building_step_kernel, n =", params["n"], ")", float(env.stats[0]), "steps")
elif which == "market":
    from sustaingym_b200.envs.electricitymarket.env import BatchedElectricityMarketEnv
    env = BatchedElectricityMarketEnv(1000)
    env.reset(seed=0)
    for t in range(4):
        env.step(torch.rand(1000, 2, device="cuda") * 100)
    torch.cuda.synchronize()
    print("market ok")
elif which == "cogen":
    from sustaingym_b200.envs.cogen.env import BatchedCogenEnv
    env = BatchedCogenEnv(256, forecast_horizon=12)
    rng = np.random.default_rng(0)
    env.reset(days=rng.integers(0, 100, 256), init_actions=env.sample_actions(rng))
    for t in range(2):
        env.step(torch.from_numpy(env.sample_actions(rng)).cuda())
    torch.cuda.synchronize()
    print("cogen ok (cogen_mlp_tc_kernel)")
