"""Quick EV step timing (bench configuration, CUDA graph replay, 4 env sets); prints us/step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator

B = int(os.environ.get("PROF_B", 65536))
gen = SyntheticTraceGenerator("caltech", num_days=256, mean_sessions=25.0, seed=1234)

def run(project, scale, stats, nsets=4, steps=3000, label=""):
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    envs, graphs = [], []
    for i in range(nsets):
        env = BatchedEVChargingEnv(gen, B, project_action_in_env=project, auto_reset=True, collect_stats=stats)
        env.reset(seed=i * B)
        a = torch.rand(B, 54, generator=g, device="cuda") * scale
        env.launch_step(a)
        envs.append((env, a)); graphs.append(env.capture_graph(a))
    torch.cuda.synchronize()
    for i in range(600): graphs[i % nsets].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): graphs[i % nsets].replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{os.environ.get('SG_LIB_PATH','default')[-16:]:16s} {label:24s} {e0.elapsed_time(e1)/steps*1e3:8.2f} us/step", flush=True)

run(True, 1.0, True, label="bench config")
run(True, 1.0, False, label="bench config, no stats")
run(True, 0.3, False, label="no violations, no stats")
