"""Small driver for ncu: EV step at the bench configuration (65 536 envs, projection on)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator

B = int(os.environ.get("PROF_B", 65536))
steps = int(os.environ.get("PROF_STEPS", 130))
project = os.environ.get("PROF_PROJECT", "1") == "1"
gen = SyntheticTraceGenerator("caltech", num_days=256, mean_sessions=25.0, seed=1234)
env = BatchedEVChargingEnv(gen, B, project_action_in_env=project, auto_reset=True)
env.reset(seed=0)
g = torch.Generator(device="cuda"); g.manual_seed(0)
acts = [torch.rand(B, 54, generator=g, device="cuda") for _ in range(4)]
for i in range(steps):
    env.launch_step(acts[i % 4])
torch.cuda.synchronize()
print("done", int(env.t[0]))
