"""Phase timeline of ev_step_main_kernel over ALL tiles of one mid-day launch, from a -DSG_EV_TRACE build (diagnostic)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from sustaingym_b200 import _lib
from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator

B = 65536
KE = int(os.environ.get("KE", 16))
T0 = int(os.environ.get("TSTEP", 150))
gen = SyntheticTraceGenerator("caltech", num_days=256, mean_sessions=25.0, seed=1234)
env = BatchedEVChargingEnv(gen, B, project_action_in_env=True, auto_reset=True, collect_stats=True)
env.reset(seed=0)
if os.environ.get("STATIC"): env._flags |= 32          # SG_EV_FLAG_STATIC_TILES
g = torch.Generator(device="cuda"); g.manual_seed(0)
acts = [torch.rand(B, 54, generator=g, device="cuda") for _ in range(4)]
for i in range(T0): env.launch_step(acts[i % 4])
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); flush.zero_(); torch.cuda.synchronize()
ntiles = (B + KE - 1) // KE
trace = torch.zeros(ntiles + 4096, 16, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.sg_ev_set_trace.argtypes = [C.c_void_p]; lib.sg_ev_set_trace.restype = C.c_int
assert lib.sg_ev_set_trace(trace.data_ptr()) == 0
env.launch_step(acts[2]); torch.cuda.synchronize()
lib.sg_ev_set_trace(None)
tr_all = trace.cpu().numpy().astype(np.float64)
tr = tr_all[:ntiles]
hl = tr_all[ntiles:]
t0 = tr[:, 5].min()
end = tr[:, 15].max()
print(f"t={T0} tiles {ntiles}; kernel span first tile start -> last tile end: {(end - t0) / 1e3:.2f} us")
names = {5: "tile start", 0: "header+prologue", 2: "stage fill", 11: "scan", 12: "keys (last grp)", 3: "keys synced", 13: "loads issued",
         14: "loads arrived", 4: "items done", 6: "checks/rate/wb", 7: "arrivals", 8: "epilogue", 15: "bulk store + wait"}
order = [5, 0, 2, 11, 12, 3, 13, 14, 4, 6, 7, 8, 15]
if os.environ.get("KEYS"):      # -DSG_EV_TRACE_KEYS build: stamps 14 / 13 sit inside the scan -> keys interval
    names.update({14: "j0 known", 13: "group bounds"})
    order = [5, 0, 2, 11, 14, 13, 12, 3, 4, 6, 7, 8, 15]
tot = (tr[:, 15] - tr[:, 5]) / 1e3
print(f"per-tile total: median {np.median(tot):.2f} p10 {np.percentile(tot,10):.2f} p90 {np.percentile(tot, 90):.2f} max {tot.max():.2f} us")
for a, b_ in zip(order[:-1], order[1:]):
    d = (tr[:, b_] - tr[:, a]) / 1e3
    print(f"  {names[a]:18s} -> {names[b_]:18s} median {np.median(d):7.2f}  p90 {np.percentile(d, 90):7.2f}  max {d.max():7.2f} us   share {d.sum() / tot.sum() * 100:5.1f} %")
start = (tr[:, 5] - t0) / 1e3
fin = (tr[:, 15] - t0) / 1e3
print("tile start times: ", np.percentile(start, [0, 25, 50, 75, 90, 99, 100]).round(1))
print("tile end times:   ", np.percentile(fin, [0, 25, 50, 75, 90, 99, 100]).round(1))
sm = tr[:, 1].astype(int)
per_sm_end = np.array([fin[sm == s_].max() for s_ in np.unique(sm)])
per_sm_n = np.array([(sm == s_).sum() for s_ in np.unique(sm)])
print(f"SMs {len(per_sm_end)}: last tile end per SM min {per_sm_end.min():.1f} median {np.median(per_sm_end):.1f} max {per_sm_end.max():.1f} us; tiles per SM min {per_sm_n.min()} max {per_sm_n.max()}")
# tiles by start quartile: do late tiles run faster (fewer warps active)?
for lo, hi in [(0, 25), (25, 50), (50, 75), (75, 100)]:
    a, b_ = np.percentile(start, [lo, hi])
    sel = (start >= a) & (start <= b_)
    print(f"  tiles starting in [{a:5.1f},{b_:5.1f}] us: median duration {np.median(tot[sel]):.2f} us")
ph = np.stack([(tr[:, b_] - tr[:, a]) / 1e3 for a, b_ in zip(order[:-1], order[1:])], 1)
print("p99 per phase:", np.percentile(ph, 99, axis=0).round(2))
print("count of tiles > 20 us:", int((tot > 20).sum()), " > 25 us:", int((tot > 25).sum()))
occn = None
for title, idx in (("10 longest tiles", np.argsort(-tot)[:10]), ("10 last-ending tiles", np.argsort(-fin)[:10])):
    print(title)
    for i in idx:
        print(f"  tile {i:5d} sm {sm[i]:3d} start {start[i]:6.1f} end {fin[i]:6.1f} total {tot[i]:6.2f} | " + " ".join(f"{x:5.2f}" for x in ph[i]))

hl = hl[hl[:, 0] > 0]
if len(hl):
    hs, he = (hl[:, 0] - t0) / 1e3, (hl[:, 1] - t0) / 1e3
    print(f"helper envs {len(hl)}: start min {hs.min():.1f} median {np.median(hs):.1f} max {hs.max():.1f}; end max {he.max():.1f}; duration median {np.median(he - hs):.2f} p90 {np.percentile(he - hs, 90):.2f} max {(he - hs).max():.2f} us; helper warps used {len(np.unique(hl[:, 2]))}")
