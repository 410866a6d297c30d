"""EV step timing under a few action distributions / launch modes (diagnostic, not the bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator

B = int(os.environ.get("PROF_B", 65536))
gen = SyntheticTraceGenerator("caltech", num_days=256, mean_sessions=25.0, seed=1234)

def run(project, scale, nsets=4, steps=2000, graph=True, label=""):
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    envs, acts, graphs = [], [], []
    for i in range(nsets):
        env = BatchedEVChargingEnv(gen, B, project_action_in_env=project, auto_reset=True, collect_stats=True)
        env.reset(seed=i * B)
        a = torch.rand(B, 54, generator=g, device="cuda") * scale
        env.launch_step(a)
        envs.append(env); acts.append(a)
        if graph:
            graphs.append(env.capture_graph(a))
    torch.cuda.synchronize()
    def step(i):
        if graph: graphs[i % nsets].replay()
        else: envs[i % nsets].launch_step(acts[i % nsets])
    for i in range(400): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps): step(i)
    e1.record(); torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    nproj = sum(float(e.stats[6]) for e in envs)
    cnt = sum(float(e.stats[0]) for e in envs)
    print(f"{label:40s} project={project} scale={scale} sets={nsets} graph={graph}: {e0.elapsed_time(e1)/steps*1e3:8.2f} us/step (wall {wall/steps*1e6:7.2f}) proj/step={nproj/cnt*B:7.1f}", flush=True)

run(True, 1.0, label="bench config")
run(True, 0.3, label="no violations")
run(False, 1.0, label="projection off")
run(True, 1.0, nsets=1, label="one env set (L2-resident)")
run(True, 1.0, graph=False, label="direct launches")
run(True, 0.3, graph=False, label="direct launches, no violations")
