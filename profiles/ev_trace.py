"""Phase timeline of ev_step_main_kernel from a -DSG_EV_TRACE build (diagnostic)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sustaingym_b200 import _lib
from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator

B = 65536
KE = int(os.environ.get("KE", 16))
scale = float(os.environ.get("SCALE", 1.0))
gen = SyntheticTraceGenerator("caltech", num_days=256, mean_sessions=25.0, seed=1234)
env = BatchedEVChargingEnv(gen, B, project_action_in_env=True, auto_reset=True, collect_stats=False)
env.reset(seed=0)
g = torch.Generator(device="cuda"); g.manual_seed(0)
acts = [torch.rand(B, 54, generator=g, device="cuda") * scale for _ in range(4)]
for i in range(130): env.launch_step(acts[i % 4])
torch.cuda.synchronize()
ntiles = (B + KE - 1) // KE
trace = torch.zeros(ntiles, 16, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.sg_ev_set_trace.argtypes = [C.c_void_p]; lib.sg_ev_set_trace.restype = C.c_int
assert lib.sg_ev_set_trace(trace.data_ptr()) == 0
env.launch_step(acts[2]); torch.cuda.synchronize()
lib.sg_ev_set_trace(None)
tr = trace.cpu().numpy().astype(np.float64)
tr = tr[tr[:, 10] > 0]                       # first tile of each warp only (persistent grids stamp entry once)
t0 = tr[:, 10].min()
names = {10: "cta entry", 0: "tables + prologue done", 2: "fill done", 11: "scan done", 12: "keys written", 3: "keys synced (last group)",
         13: "loads issued", 14: "loads arrived",
         4: "items done (last chunk)", 5: "owner done", 6: "checks done", 7: "arrivals done", 8: "epilogue done", 9: "fallback + stats done"}
print(f"kernel span (first entry -> last exit): {(tr[:, 9].max() - t0) / 1e3:.2f} us; tiles {ntiles}")
order = [10, 0, 2, 11, 12, 3, 13, 14, 4, 6, 7, 8, 9]
for a, b_ in zip(order[:-1], order[1:]):
    d = (tr[:, b_] - tr[:, a]) / 1e3
    print(f"  {names[a]:30s} -> {names[b_]:28s} median {np.median(d):7.2f}  p90 {np.percentile(d, 90):7.2f}  max {d.max():7.2f} us")
print(f"  entry time of tiles: median {np.median(tr[:, 10] - t0) / 1e3:.2f} us, p90 {np.percentile(tr[:, 10] - t0, 90) / 1e3:.2f}, max {(tr[:, 10].max() - t0) / 1e3:.2f}")
d = (tr[:, 9] - tr[:, 10]) / 1e3
print(f"  per-warp total: median {np.median(d):.2f} p90 {np.percentile(d, 90):.2f} max {d.max():.2f} us")
