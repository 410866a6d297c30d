"""BuildingEnv n=5 single step at 262 144 and 32 768 envs: direct launches vs CUDA-graph replays (CUDA events)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from sustaingym_b200.envs.building.env import BatchedBuildingEnv
g = np.load(os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden", "building_retail_s5.npz"))
n, T = 5, 105108
t = np.arange(T) / 288.0
params = dict(n=n, A_d=g["A_d"], BD_d=g["BD_d"], target=g["target"], ac_map=g["ac_map"],
              out_temp=20.5 + 21.5 * np.sin(2 * np.pi * (t - 0.3)), ground_temp=np.full(T, 20.9), ghi=np.clip(np.sin(2 * np.pi * (t - 0.25)), 0, 1),
              metabolism=np.full(T, 120.0), reward_beta=0.999, reward_pnorm=2, episode_len=288, time_resolution=300, max_power=8000,
              is_continuous_action=True, temp_range=(-40, 40))
for B in (262144, 32768):
    envs = []
    for i in range(12):
        env = BatchedBuildingEnv(params, B, auto_reset=True)
        env.reset(epochs=np.random.default_rng(i).integers(0, T - 289, B))
        a = torch.rand(B, n, device="cuda") * 2 - 1
        envs.append((env, a, env.capture_graph(a)))
    def timed(fn, steps=600):
        for i in range(50): fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for i in range(steps): fn(i)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps * 1e3, (time.perf_counter() - t0) / steps * 1e6
    print(B, "graph replay   us/step (gpu, wall):", timed(lambda i: envs[i % 12][2].replay()))
    print(B, "env.step       us/step (gpu, wall):", timed(lambda i: envs[i % 12][0].step(envs[i % 12][1])))
    def raw(i):
        env, a, _ = envs[i % 12]
        env.lib.sg_building_step(__import__("ctypes").byref(env._model), B, a.data_ptr(), env.state.data_ptr(), env.epoch.data_ptr(),
                                 env.runs.data_ptr(), env.reward.data_ptr(), env.done.data_ptr(), env._flags, None, env._stream())
    print(B, "raw C call     us/step (gpu, wall):", timed(raw))
