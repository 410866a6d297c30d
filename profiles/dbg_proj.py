import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from oracle import ev as O
from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator
env = BatchedEVChargingEnv(SyntheticTraceGenerator(num_days=1), 1)
net = O.caltech_acn()
rng = np.random.default_rng(0)
M = 2000
a = rng.uniform(0, 1, (M, 54)).astype(np.float32)
occ = rng.uniform(0, 1, (M, 54)) < rng.uniform(0.05, 1.0, (M, 1))
dem = np.where(occ, np.clip(rng.lognormal(1.5, 1.2, (M, 54)), 0.002, 100), 0).astype(np.float32)
a[0] = 0; dem[1] = 0; a[2] = 1; dem[2] = 100
x, st, it = env.project(torch.from_numpy(a).cuda(), torch.from_numpy(dem).cuda())
t0=time.time(); xo, info = O.project_action(net, a, dem); print('oracle time', time.time()-t0)
x, st, it = x.cpu().numpy(), st.cpu().numpy(), it.cpu().numpy()
slow = np.zeros(M, bool); slow[info["slow"]] = True
print('slow', slow.sum(), 'gpu projected', (st&1).sum(), 'cap', (st&2).sum())
print('iters hist gpu', np.bincount(it[slow]))
print('iters hist oracle', np.bincount(info['iters'][slow]))
err = np.abs(x-xo).max(axis=1)
print('max err', err.max(), 'err of capped', err[(st&2)>0][:10], 'idx', np.nonzero(st&2)[0][:10])
u = O.projection_bounds(dem)
for m in np.nonzero(st&2)[0][:5]:
    print(m, 'n_free', (u[m]>0).sum(), 'min u', u[m][u[m]>0].min(), 'oracle iters', info['iters'][m], 'kkt', info['kkt'][m], 'err', err[m])
