"""Quick Cogen step timing: tensor-core vs FFMA plant kernel (4096 envs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sustaingym_b200.envs.cogen.env import BatchedCogenEnv
B = int(os.environ.get("PROF_B", 4096))
for tcf in (True, False):
    env = BatchedCogenEnv(B, forecast_horizon=12, tensor_cores=tcf)
    rng = np.random.default_rng(0)
    env.reset(days=rng.integers(0, 60, B), init_actions=env.sample_actions(rng))
    acts = [torch.from_numpy(env.sample_actions(rng)).cuda() for _ in range(4)]
    for i in range(10): env.step(acts[i % 4]); env.day_t[:, 1] = 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for i in range(n):
        env.step(acts[i % 4])
        if i % 40 == 39: env.day_t[:, 1] = 0
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"tensor_cores={tcf}: {ms*1e3:.1f} us/step  {B/(ms*1e-3):.3e} env-steps/s  {3.684096e6*B/(ms*1e-3)/1e12:.1f} TFLOP/s (algorithmic)", flush=True)
