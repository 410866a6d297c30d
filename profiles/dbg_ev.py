"""Debug helper: first obs/schedule mismatch between the EV kernel and the oracle on a test configuration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import ev as O
from helpers import tie_distance
from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator

B, mean, seed_gen, seed = int(os.environ.get("DB", 200)), float(os.environ.get("DMEAN", 80.0)), 5, 9
gen = SyntheticTraceGenerator("caltech", num_days=5, mean_sessions=mean, seed=seed_gen)
env = BatchedEVChargingEnv(gen, B, project_action_in_env=True)
net = O.caltech_acn()
orc = O.EVOracle(net, [[O.Session(*x) for x in d] for d in gen.days], gen.moer.astype(np.float64), B, 36, True)
env.reset(seed=seed)
days = gen.first_days(seed + np.arange(B), B)
orc.reset(days, gen.moer_days(days))
rng = np.random.default_rng(seed)
n = 54
clean = np.ones(B, bool)
for t in range(288):
    a = rng.uniform(0, 1, (B, 54)).astype(np.float32)
    ob, r, d, _, info = env.step(torch.from_numpy(a).cuda())
    o_ob, o_r, o_d, _, o_info = orc.step(a)
    clean &= tie_distance(net, o_info["x"]).min(axis=1) > 1e-6
    ob = ob.cpu().numpy(); f = O.flatten_obs(o_ob)
    sched = env.pilots.cpu().numpy().astype(np.float64)
    st = env.status.cpu().numpy()
    # occupancy: oracle plugged set vs the kernel's mask
    occ_g = env.occ.cpu().numpy().astype(np.uint64)
    occ_o = np.zeros(B, np.uint64)
    bi, sidx = np.nonzero(orc.plugged)
    np.bitwise_or.at(occ_o, bi, np.uint64(1) << orc._stn[orc.day][bi, sidx].astype(np.uint64))
    if (occ_g != occ_o).any() and t < 287:
        w = np.nonzero(occ_g != occ_o)[0]
        print("step", t, "occupancy mismatch envs", w[:8].tolist(), "status", st[w[:8]].tolist() if 'st' in dir() else None,
              "status now", env.status.cpu().numpy()[w[:8]].tolist(), [hex(int(x)) for x in (occ_g[w[:4]] ^ occ_o[w[:4]])])
    bad = np.argwhere((ob[:, n:2 * n] != f[:, n:2 * n]) & clean[:, None])
    bads = np.argwhere((sched != o_info["schedule"]) & clean[:, None])
    if len(bad) or len(bads):
        print("step", t, "est mismatches", bad[:5].tolist(), "sched mismatches", bads[:5].tolist())
        for (b, i) in bad[:3]:
            print(f" env {b} station {i}: gpu est {ob[b, n + i]} dem {ob[b, i]} | oracle est {f[b, n + i]} dem {f[b, i]} | status {st[b]} slow {b in o_info['slow']} pilot {sched[b, i]} vs {o_info['schedule'][b, i]}")
            dl, si = env.export_state()
            print("   exported delivered", float(dl[b, i]), "slot", int(si[b, i]), "occ bit", (int(env.occ[b]) >> int(i)) & 1, "tile", b // 16, "lane", b % 16)
            print("   day sessions at station:", [s for s in gen.tables.sessions_of_day(int(days[b])) if s[4] == i])
        break
else:
    print("no mismatch in 288 steps; clean fraction", clean.mean())
