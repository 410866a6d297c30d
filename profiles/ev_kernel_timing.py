"""How long does one EV step's kernel pair take on the GPU?  Three ways of timing it (diagnostic for bench.py's
roofline.kernel_ms): per-launch event pairs, one event pair around L back-to-back stream launches (both queued
behind a filler so the host is ahead), and CUDA-graph replays."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator

B, N_SETS, L = 65536, 4, 120
dev = torch.device("cuda")
gen = SyntheticTraceGenerator("caltech", num_days=256, mean_sessions=25.0, seed=1234)
envs, acts, graphs = [], [], []
for i in range(N_SETS):
    env = BatchedEVChargingEnv(gen, B, moer_forecast_steps=36, project_action_in_env=True, device=dev, auto_reset=True, collect_stats=True)
    env.reset(seed=i * B)
    a = torch.rand(B, 54, device=dev)
    env.launch_step(a)
    envs.append(env); acts.append(a); graphs.append(env.capture_graph(a))
for i in range(1200):
    graphs[i % N_SETS].replay()
torch.cuda.synchronize()
filler = torch.empty(1 << 28, dtype=torch.float32, device=dev)

def fill(ms=60.0):
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(); filler.normal_(); f1.record(); torch.cuda.synchronize()
    for _ in range(int(np.ceil(ms / max(f0.elapsed_time(f1), 0.05)))):
        filler.normal_()

for rep in range(2):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(L)]
    fill()
    for i in range(L):
        evs[i][0].record(); envs[i % N_SETS].launch_step(acts[i % N_SETS]); evs[i][1].record()
    torch.cuda.synchronize()
    d = sorted(a.elapsed_time(b) for a, b in evs)
    print(f"per-launch event pairs : trimmed mean {1e3*np.mean(d[L//10:L-L//10]):.1f} us  min {1e3*d[0]:.1f} max {1e3*d[-1]:.1f}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fill()
    e0.record()
    for i in range(L):
        envs[i % N_SETS].launch_step(acts[i % N_SETS])
    e1.record(); torch.cuda.synchronize()
    print(f"back-to-back launches  : {1e3*e0.elapsed_time(e1)/L:.1f} us per step")
    e0.record()
    for i in range(2000):
        graphs[i % N_SETS].replay()
    e1.record(); torch.cuda.synchronize()
    print(f"graph replays          : {1e3*e0.elapsed_time(e1)/2000:.1f} us per step")

# one graph stepping all N_SETS env sets (8 kernel nodes): amortises the per-replay launch latency
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(dev)
s.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        for env, a in zip(envs, acts):
            env.launch_step(a)
torch.cuda.current_stream(dev).wait_stream(s)
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(288 * 2):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"one graph per {N_SETS} steps : {1e3*e0.elapsed_time(e1)/(288*2*N_SETS):.1f} us per step (2 episodes)")
    e0.record()
    for i in range(288 * 2 * N_SETS):
        graphs[i % N_SETS].replay()
    e1.record(); torch.cuda.synchronize()
    print(f"one graph per step     : {1e3*e0.elapsed_time(e1)/(288*2*N_SETS):.1f} us per step (2 episodes)")
