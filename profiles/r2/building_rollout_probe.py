"""BuildingEnv n=5 fused 288-step rollout at 262 144 envs (the bench's config-4 rollout leg), CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from sustaingym_b200.envs.building.env import BatchedBuildingEnv
g = np.load(os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden", "building_retail_s5.npz"))
n, T, B, S = 5, 105108, 262144, 288
t = np.arange(T) / 288.0
params = dict(n=n, A_d=g["A_d"], BD_d=g["BD_d"], target=g["target"], ac_map=g["ac_map"],
              out_temp=20.5 + 21.5 * np.sin(2 * np.pi * (t - 0.3)), ground_temp=np.full(T, 20.9), ghi=np.clip(np.sin(2 * np.pi * (t - 0.25)), 0, 1),
              metabolism=np.full(T, 120.0), reward_beta=0.999, reward_pnorm=2, episode_len=288, time_resolution=300, max_power=8000,
              is_continuous_action=True, temp_range=(-40, 40))
sets = []
for i in range(2):
    env = BatchedBuildingEnv(params, B, auto_reset=True)
    env.reset(epochs=np.random.default_rng(i).integers(0, T - 289, B))
    sets.append((env, torch.rand(S, B, n, device="cuda") * 2 - 1))
for i in range(2): sets[i % 2][0].rollout(sets[i % 2][1], keep_obs=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(6): sets[i % 2][0].rollout(sets[i % 2][1], keep_obs=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 6
print(f"rollout {ms:.3f} ms per launch = {B * S / ms / 1e6:.2f}e9 env-steps/s, frac of 6535 GB/s {105 * B * S / (ms * 1e-3) / 1e9 / 6535.1:.3f}")
