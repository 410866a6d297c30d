"""Reference trajectories of EVChargingEnv — for the day acnportal + cvxpy (+ MOSEK or CLARABEL) are importable.

The EV oracle (oracle/ev.py) restates un-vendored third-party code from memory; this script is
the pin waiting for the dependency: it runs the UNMODIFIED reference env (sustaingym/envs/evcharging/env.py) on the
real Caltech traces with seeded random actions and stores observations / rewards / schedules, in the layout
tests/test_oracle_ev.py::test_oracle_matches_reference_trajectory compares the oracle (and through it the kernels) with.

    python tests/golden/gen_ev_golden.py /root/reference  ->  tests/golden/ev_reference_trajectory.npz   (or exits 3)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(ref_root: str) -> int:
    try:
        import acnportal  # noqa: F401
        import cvxpy  # noqa: F401
        import gymnasium  # noqa: F401
    except ImportError as e:
        print(f"reference dependencies missing ({e}); no EV golden trajectory can be produced here")
        return 3
    sys.path.insert(0, ref_root)
    # the package __init__ of this commit imports a module that does not exist: import the files directly
    from sustaingym.envs.evcharging.env import EVChargingEnv
    from sustaingym.envs.evcharging.event_generation import RealTraceGenerator
    out = {}
    for seed, period in ((2, ("2020-02-01", "2020-05-31")), (10, ("2019-05-01", "2019-08-31"))):
        gen = RealTraceGenerator("caltech", period, sequential=True)
        env = EVChargingEnv(gen, project_action_in_env=True)
        obs, info = env.reset(seed=seed)
        rng = np.random.default_rng(seed)
        acts, rows, rews, scheds = [], [], [], []
        done = False
        while not done:
            a = rng.uniform(0, 1, env.num_stations).astype(np.float32)
            acts.append(a.copy())
            obs, r, done, _, info = env.step(a.copy())
            rows.append(np.concatenate([obs[k].ravel() for k in sorted(obs)]))
            rews.append(r)
            scheds.append(np.array([env._simulator.pilot_signals[i, env.t - 1] for i in range(env.num_stations)]))
        key = f"{period[0]}|{seed}|"
        out[key + "actions"], out[key + "obs"] = np.array(acts), np.array(rows, np.float32)
        out[key + "rewards"], out[key + "schedule"] = np.array(rews), np.array(scheds)
        out[key + "max_profit"] = np.array(info["max_profit"])
        out[key + "moer"] = np.asarray(env.moer)
        out[key + "sessions"] = np.array([(ev.arrival, ev.departure, ev.estimated_departure, ev.requested_energy,
                                           env.cn.station_ids.index(ev.station_id)) for ev in env._evs], np.float64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ev_reference_trajectory.npz"), **out)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
