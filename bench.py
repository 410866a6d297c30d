#!/usr/bin/env python
"""Headline benchmark: EVChargingEnv (Caltech, 54 stations, k = 36, action projection on),
batch 65 536 envs per GPU, continuous actions (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-secondary]

A *step* is one ``env.step()`` of the whole batch (one pass of the hot path: projection ->
pilot rounding -> acnsim period -> observation -> reward).  Reported on ONE JSON line:

  value      env-steps/s, whole job, actions already resident in HBM, CUDA-event timed, max over ranks
  e2e        the same through the public host API (pinned host actions in, host obs/reward/done out) with the
             H2D copy of the actions and the D2H copy of the results of every step inside the timed region
  roofline   algorithmic bytes per launch (1709 B/env-step) / the step time OF THE SAME TIMED
             WINDOW as ``value`` (ONE kernel per step, launched as a CUDA graph), against
             MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the NumPy oracle port timed on this box's host cores on a bounded sample
  secondary  BASELINE configs 3-5 (Market 65 536; Building n=5 262 144 envs split over the ranks, single step
             and fused rollout; Cogen 4 agents k=12, 32 768 envs split over the ranks), each with value,
             roofline and (N = 1) cpu_baseline, timed under the same launch as the headline

Phase of the simulated day.  The cost of an EV step follows the garage occupancy (empty at night, ~12 EVs
around noon), so the four env sets are pre-rolled — outside the timed region — to t = 36, 108, 180, 252 (the
mid-points of the four quarters of the 288-step day) and stepped round-robin: ANY timed window, however
short, samples the day-average workload.  The CPU arms pre-roll the same phases (one quarter of each worker's
envs per phase).

L2 policy: the four env sets (each 65 536 envs, ~85 MB touched per step) are stepped round-robin, so
consecutive timed steps touch disjoint ~340 MB of HBM (> 126 MB L2).

Multi-GPU (torchrun): one process per GPU, envs are independent so the batch shards with no data-path
collective ("weak": 65 536 envs per GPU); the one collective is the all-reduce of the 16-double
reward-statistics vector (SUM + MIN + MAX, ``sustaingym_b200.dist.all_reduce_stats``) every 16 steps, issued
on a side stream inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 65536
N_STATIONS = 54
MOER_STEPS = 36
OBS_DIM = 2 * N_STATIONS + MOER_STEPS + 2
ALGO_BYTES_PER_ENV_STEP = 1709          # SURVEY.md 8(d), EV n=54 k=36
N_SETS = 4
PHASES = (36, 108, 180, 252)            # day phase (t) each env set / CPU quarter is pre-rolled to
STATS_EVERY = int(os.environ.get("SG_BENCH_STATS_EVERY", 16))   # steps between two exchanges of the statistics vector
POOL_DAYS = 256
MEAN_SESSIONS = 25.0
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
WORKLOAD = "EVChargingEnv caltech-54 continuous batch65536 project_action_in_env=True moer_steps=36"
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r2", "ev_step_traffic.json")   # written from the ncu --set full capture


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the step kernel from the committed ncu --set full capture (None if absent)."""
    if os.path.exists(TRAFFIC_FILE):
        d = json.load(open(TRAFFIC_FILE))
        return d.get("dram_bytes_per_launch"), d.get("source")
    return None, None


class ClockSampler(threading.Thread):
    """Samples SM clock and clock-event reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self._stop_evt = index, period_s, [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def finish(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "samples": len(s), "reasons": sorted(self.reasons)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except (ValueError, IndexError):
            pass
    return local_rank


# ----------------------------------------------------------------------------------------------
# CPU legs (the oracle port): the only place outside tests/smoke that executes oracle/
# ----------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One EVOracle of B envs; quarter q of the envs is pre-rolled to day phase PHASES[q] (untimed), exactly like
    the GPU arm's four env sets, then `steps` steps are timed; an env that finishes its day starts the next one
    (the GPU arm auto-resets in-kernel)."""
    seed, B, steps = args
    import numpy as np
    from oracle import ev as O
    from sustaingym_b200.data.ev_traces import synthetic_days, synthetic_moer
    net = O.caltech_acn()
    D = 32
    days = synthetic_days(D, N_STATIONS, MEAN_SESSIONS, seed)
    moer = synthetic_moer(D, seed + 1).astype(np.float64)
    orc = O.EVOracle(net, [[O.Session(*x) for x in d] for d in days], moer, B, MOER_STEPS, True)
    day = np.arange(B) % D
    orc.reset(day, (day + 1) % D)
    rng = np.random.default_rng(seed)
    acts = rng.uniform(0, 1, (8, B, N_STATIONS)).astype(np.float32)
    quarter = (np.arange(B) * len(PHASES)) // B
    # pre-roll: quarter q is (re)started max(PHASES) - PHASES[q] steps into the pre-roll, so that after max(PHASES)
    # steps it stands at t = PHASES[q]
    for t in range(max(PHASES)):
        for q, ph in enumerate(PHASES):
            if t == max(PHASES) - ph and t > 0:
                orc.reset(day, (day + 1) % D, mask=quarter == q)
        orc.step(acts[t % 8])

    def next_day():
        fin = orc.t >= 288
        if fin.any():
            day[fin] = (day[fin] + 1) % D
            orc.reset(day, (day + 1) % D, mask=fin)
    t0 = time.perf_counter()
    for t in range(steps):
        next_day()
        orc.step(acts[t % 8])
    return B * steps, time.perf_counter() - t0


def cpu_oracle_rate(procs: int, B: int, steps: int):
    """env-steps/s of the vectorised NumPy oracle on `procs` processes (one EVOracle each)."""
    if procs == 1:
        n, dt = _cpu_worker((1234, B, steps))
        return n / dt
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(1234 + i, B, steps) for i in range(procs)])
    # each worker times its own loop (set-up and pre-roll excluded); aggregate = sum of per-worker rates
    return sum(n / dt for n, dt in res)


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU path.  The reference itself (acnportal + cvxpy + MOSEK)
    cannot be installed here, so this times the oracle port — a vectorised NumPy restatement that
    is much FASTER than the real per-env Python/MOSEK path — on all host cores, on the same day phases
    as the GPU arm."""
    if rank != 0:
        return
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    cores = os.cpu_count() or 1
    B = 1024
    steps = max(1, args.steps)
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_rate(cores, 64, 2)
    t0 = time.perf_counter()
    rate = cpu_oracle_rate(cores, B, steps)
    wall = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * B * cores / rate, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": B * cores, "day_phases": list(PHASES)},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{cores} processes x {B} envs x {steps} steps of the NumPy oracle "
                                   f"(oracle/ev.py EVOracle, projection on), a quarter of each worker's envs pre-rolled "
                                   f"to t = {list(PHASES)} (untimed); wall incl. pre-roll {wall:.1f}s"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# shared helpers of the GPU arm
# ----------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, rank, local_rank, world, dev):
        self.rank, self.local_rank, self.world, self.dev = rank, local_rank, world, dev

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier(device_ids=[self.local_rank])
        torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps: int, warmup: int) -> float:
        """ms per call of fn: `warmup` untimed calls, then exactly `steps` calls between a barrier + synchronize on
        both sides, CUDA events on the launching stream, max over ranks."""
        import torch
        for i in range(warmup):
            fn(i)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps


def shard(total: int, ctx: Ctx) -> int:
    from sustaingym_b200.dist import shard_range
    lo, hi = shard_range(total, ctx.rank, ctx.world)
    return hi - lo


# ----------------------------------------------------------------------------------------------
# secondary configs (BASELINE.json configs[2..4])
# ----------------------------------------------------------------------------------------------
def sec_building(ctx: Ctx, hbm: float, cpu: bool) -> dict:
    """Config 4: BuildingEnv 5-zone (RetailStandalone), 262 144 envs split over the ranks: the drop-in single
    step (direct launch or CUDA-graph replay, whichever is faster at this shard size) and the 288-step fused rollout."""
    import numpy as np
    import torch
    from sustaingym_b200.envs.building.env import BatchedBuildingEnv
    TOTAL = 262144
    B = shard(TOTAL, ctx)
    g = np.load(os.path.join(ROOT, "tests", "golden", "building_retail_s5.npz"))
    n, T = 5, 105108
    t = np.arange(T) / 288.0
    params = dict(n=n, A_d=g["A_d"], BD_d=g["BD_d"], target=g["target"], ac_map=g["ac_map"],
                  out_temp=20.5 + 21.5 * np.sin(2 * np.pi * (t - 0.3)) * np.sin(np.pi * t / 365) ** 2,
                  ground_temp=np.full(T, 20.9), ghi=np.clip(np.sin(2 * np.pi * (t - 0.25)), 0, 1),
                  metabolism=np.full(T, 120.0), reward_beta=0.999, reward_pnorm=2, episode_len=288,
                  time_resolution=300, max_power=8000, is_continuous_action=True, temp_range=(-40, 40))
    algo_step = 105                                            # SURVEY.md 8(d), n = 5
    # ---- single step: enough env sets that consecutive replays touch > 2 x L2 of distinct memory -------------
    n_sets = int(min(128, max(4, math.ceil(300e6 / (B * 120.0)))))
    rng = np.random.default_rng(100 + ctx.rank)
    sets = []
    for i in range(n_sets):
        env = BatchedBuildingEnv(params, B, auto_reset=True)
        env.reset(epochs=rng.integers(0, T - 289, B))
        a = torch.rand(B, n, device=ctx.dev) * 2 - 1
        sets.append((env, a, env.capture_graph(a)))
    steps1 = 2000
    ms_graph = ctx.timed(lambda i: sets[i % n_sets][2].replay(), steps1, 100)
    ms_direct = ctx.timed(lambda i: sets[i % n_sets][0].step(sets[i % n_sets][1]), steps1, 100)
    # both are the public step: `env.step(actions)` launches the kernel directly, `env.capture_graph(actions).replay()`
    # replays it as a one-node CUDA graph (cheaper on the host, a little dearer on the device); the better one counts
    ms1, how1 = (ms_graph, "one CUDA-graph replay per step") if ms_graph <= ms_direct else (ms_direct, "env.step(): one direct launch per step")
    v1 = TOTAL / (ms1 * 1e-3)
    # ---- fused rollout: S steps per launch, actions pre-staged [S, B, n], observations kept ------------------
    S = 288
    rsets = []
    for i in range(2):
        env = sets[i][0]
        rsets.append((env, torch.rand(S, B, n, device=ctx.dev) * 2 - 1))
    msr = ctx.timed(lambda i: rsets[i % 2][0].rollout(rsets[i % 2][1], keep_obs=True), 6, 2)
    vr = TOTAL * S / (msr * 1e-3)
    out = {
        "config": {"workload": f"BuildingEnv RetailStandalone n=5, 262144 envs over {ctx.world} GPU(s) = {B} per GPU",
                   "l2_policy": f"single step: {n_sets} env sets round-robin ({n_sets * B * 120 / 1e6:.0f} MB touched between "
                                f"reuses); rollout: 2 sets x {S * B * (n * 4 + (n + 4) * 4 + 9) / 1e6:.0f} MB of actions+obs per launch"},
        "metric": METRIC, "unit": UNIT, "scaling": "strong", "dtype": "f64",
        "single_step": {"value": v1, "ms_per_step": ms1, "ms_per_step_graph_replay": ms_graph, "ms_per_step_direct_launch": ms_direct,
                        "steps": steps1, "gpu_launches": steps1,
                        "roofline": {"bound": "hbm", "achieved": algo_step * B / (ms1 * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                     "frac": algo_step * B / (ms1 * 1e-3) / 1e9 / hbm, "traffic": None,
                                     "kernel": f"building_step1_kernel<5> (persistent CTAs, two-slot tile ring; {how1})"}},
        "fused_rollout": {"value": vr, "ms_per_launch": msr, "steps_per_launch": S, "launches": 6, "gpu_launches": 6,
                          "roofline": {"bound": "hbm", "achieved": algo_step * B * S / (msr * 1e-3) / 1e9, "peak": hbm,
                                       "unit": "GB/s", "frac": algo_step * B * S / (msr * 1e-3) / 1e9 / hbm, "traffic": None,
                                       "kernel": "building_step_kernel<5> (288 fused steps, actions pre-staged)"}},
        "value": v1,
    }
    if cpu:
        from oracle.building import BuildingOracle, BuildingParams
        p = BuildingParams(n=n, A_d=g["A_d"], BD_d=g["BD_d"], target=g["target"], ac_map=g["ac_map"],
                           out_temp=params["out_temp"], ground_temp=params["ground_temp"], ghi=params["ghi"],
                           metabolism=params["metabolism"], q_rate=0.024, error_rate=0.999)
        orc = BuildingOracle(p, 4096); orc.reset(np.arange(4096))
        a = np.random.default_rng(0).uniform(-1, 1, (4096, n)).astype(np.float32)
        t0 = time.perf_counter()
        for _ in range(150):
            orc.step(a)
        out["cpu_baseline"] = {"value": 4096 * 150 / (time.perf_counter() - t0), "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": "NumPy oracle (oracle/building.py), 4096 envs x 150 steps, 1 process; the reference class "
                                         "itself under the import shim does 14 557 steps/s per process (BASELINE.md)"}
    del sets, rsets
    torch.cuda.empty_cache()
    return out


def sec_cogen(ctx: Ctx, pk: dict, cpu: bool) -> dict:
    """Config 5: CogenEnv multi-agent (4 units) k = 12, 32 768 envs split over the ranks."""
    import numpy as np
    import torch
    from sustaingym_b200.data.cogen_ambients import synthetic_ambients
    from sustaingym_b200.envs.cogen.env import BatchedCogenEnv
    TOTAL = 32768
    B = shard(TOTAL, ctx)
    amb = synthetic_ambients(256, seed=0)
    rng = np.random.default_rng(200 + ctx.rank)
    n_sets = int(min(16, max(2, math.ceil(300e6 / (B * 2700.0)))))      # scratch 1564 B + obs/action/reward ~1.1 KB per env
    sets = []
    for i in range(n_sets):
        env = BatchedCogenEnv(B, amb, forecast_horizon=12, forecast_noise_std=0.0)
        env.reset(days=rng.integers(0, 200, B), init_actions=env.sample_actions(rng))
        sets.append((env, torch.from_numpy(env.sample_actions(rng)).to(ctx.dev)))
    cnt = [0]

    def step(i):
        env, a = sets[i % n_sets]
        env.step(a)
        cnt[0] += 1
        if cnt[0] % (90 * n_sets) == 0:                 # keep the clocks inside the day
            for e, _ in sets:
                e.day_t[:, 1] = 0
    steps = 240 if B <= 8192 else 80
    ms = ctx.timed(step, steps, 16)
    flops = 3684096 * B
    tf32_peak = float(pk.get("bf16_tflops", 1590.0)) / 2.0
    out = {
        "config": {"workload": f"CogenEnv 4 agents k=12, 32768 envs over {ctx.world} GPU(s) = {B} per GPU; one step = inputs + plant "
                               f"model (tcgen05 3xTF32) + constraints/reward/obs kernels",
                   "l2_policy": f"{n_sets} env sets round-robin (~{n_sets * B * 2700 / 1e6:.0f} MB of per-env data between reuses); "
                                f"the 22 MB of weights are shared and L2-resident by design"},
        "metric": METRIC, "unit": UNIT, "scaling": "strong", "dtype": "tf32x3 (f32 accumulate)",
        "value": TOTAL / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "gpu_launches": 3 * steps,
        "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                     "frac": flops / (ms * 1e-3) / 1e12 / tf32_peak, "traffic": None, "kernel": "cogen_mlp_tc_kernel",
                     "note": "algorithmic flops (3.684 MFLOP per env-step); the 3xTF32 split executes 3x that on the tensor pipe; "
                             "peak = measured dense bf16 / 2"},
    }
    if cpu:
        from oracle.cogen import CogenOracle
        orc = CogenOracle(amb, 1024, forecast_horizon=12)
        a = sets[0][0].sample_actions(rng)[:1024]
        orc.reset(rng.integers(0, 200, 1024), a)
        t0 = time.perf_counter()
        for _ in range(5):
            orc.step(a)
        out["cpu_baseline"] = {"value": 1024 * 5 / (time.perf_counter() - t0), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                               "sample": "NumPy oracle (reference ONNX weights, oracle/cogen.py), 1024 envs x 5 steps, BLAS threads"}
    del sets
    torch.cuda.empty_cache()
    return out


def sec_market(ctx: Ctx, hbm: float, cpu: bool) -> dict:
    """Config 3: ElectricityMarketEnv (spec-defined), 65 536 envs per GPU."""
    import numpy as np
    import torch
    from sustaingym_b200.envs.electricitymarket.env import BatchedElectricityMarketEnv, default_fleet, synthetic_load_moer
    B = 65536
    cost, cap = default_fleet()
    load, moer = synthetic_load_moer(64, 1, seed=1)
    n_sets = 24                                           # 24 x ~11 MB per step: beyond L2
    sets = []
    for i in range(n_sets):
        env = BatchedElectricityMarketEnv(B, load, moer, cost, cap)
        env.reset(seed=i + 100 * ctx.rank)
        a = torch.rand(B, 2, device=ctx.dev) * 120
        sets.append((env, a, env.capture_graph(a)))
    cnt = [0]

    def step(i):
        sets[i % n_sets][2].replay()
        cnt[0] += 1
        if cnt[0] % (n_sets * 280) == 0:
            for e, _, _ in sets:
                e.day_t[:, 1] = 0
    steps = 3000
    ms = ctx.timed(step, steps, 100)
    algo = (8 + 4 * 8 * 2 + 8 + 8 + 40 + 8 + 1 + 24) * B      # action, state r/w, day_t, prev_a, obs, reward, done, info
    out = {
        "config": {"workload": "ElectricityMarketEnv spec v1 (envs/electricitymarket/SPEC.md; no reference source exists), 33 "
                               "generators, 1 battery, k=1, 65536 envs per GPU, one CUDA-graph replay per step",
                   "l2_policy": f"{n_sets} env sets round-robin (~{n_sets * algo / 1e6:.0f} MB between reuses)"},
        "metric": METRIC, "unit": UNIT, "scaling": "weak", "dtype": "f64",
        "value": ctx.world * B / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "gpu_launches": steps,
        "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": algo / (ms * 1e-3) / 1e9 / hbm, "traffic": None, "kernel": "market_step_kernel"},
    }
    if cpu:
        from oracle.market import MarketOracle, MarketParams
        orc = MarketOracle(MarketParams(cost, cap), load, moer, 16384)
        orc.reset(np.arange(16384) % 64)
        a = np.random.default_rng(0).uniform(0, 120, (16384, 2)).astype(np.float32)
        t0 = time.perf_counter()
        for _ in range(50):
            orc.step(a)
        out["cpu_baseline"] = {"value": 16384 * 50 / (time.perf_counter() - t0), "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": "NumPy oracle (oracle/market.py), 16384 envs x 50 steps, 1 process"}
    del sets
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="envs per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        args.steps = 400 if args.steps is None else args.steps    # ~10 s of CPU work on all cores
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference(args, rank)
        return
    args.steps = 20000 if args.steps is None else args.steps
    args.warmup = 1200 if args.warmup is None else max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist

    from sustaingym_b200.dist import all_reduce_stats
    from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
    from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator

    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = Ctx(rank, local_rank, world, dev)
    pk, peak_src = peaks()
    hbm = float(pk["hbm_gbs"])
    B = args.batch
    gen = SyntheticTraceGenerator("caltech", num_days=POOL_DAYS, mean_sessions=MEAN_SESSIONS, seed=1234 + rank)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    envs, acts, graphs = [], [], []
    shared_stats = None
    for i in range(N_SETS):
        env = BatchedEVChargingEnv(gen, B, moer_forecast_steps=MOER_STEPS, project_action_in_env=True,
                                   device=dev, auto_reset=True, collect_stats=True)
        if shared_stats is None:
            shared_stats = env.stats
        env.stats = shared_stats                # every env set accumulates into one statistics vector
        env.reset(seed=i * B + rank * N_SETS * B)
        a = torch.rand(B, N_STATIONS, generator=g, device=dev, dtype=torch.float32)
        env.launch_step(a)                      # first launch outside capture (module load, smem carve-out)
        gr = env.capture_graph(a)
        for _ in range(PHASES[i] - 1):          # pre-roll (untimed) to this set's day phase
            gr.replay()
        envs.append(env); acts.append(a); graphs.append(gr)
    torch.cuda.synchronize()
    assert [int(e.t[0]) for e in envs] == list(PHASES)

    def occupancy() -> float:
        """mean number of plugged-in EVs per env over the four env sets (popcount of the occupancy masks)"""
        tot = 0.0
        sh = torch.arange(64, device=dev, dtype=torch.int64)
        for e in envs:
            tot += float(((e.occ.unsqueeze(1) >> sh) & 1).sum(1).double().mean())
        return tot / len(envs)

    for i in range(args.warmup):
        graphs[i % N_SETS].replay()
    envs[0].reset_stats()                       # statistics of the timed window only
    occ0 = occupancy()
    side = torch.cuda.Stream(dev)
    main_stream = torch.cuda.current_stream(dev)
    tick = torch.cuda.Event()
    stats_buf = torch.zeros((world + 1) * shared_stats.numel(), dtype=torch.float64, device=dev)   # no allocation in the timed region
    job_stats = [None]
    for _ in range(2):                                      # warms the side stream, the event and the NCCL communicator up
        tick.record(main_stream)
        side.wait_event(tick)
        with torch.cuda.stream(side):
            job_stats[0] = all_reduce_stats(shared_stats, out=stats_buf)
        main_stream.wait_stream(side)
    # (NVML is initialised BEFORE the barrier: it takes tens of milliseconds on an 8-GPU box and a different time on every
    # rank; after the barrier that skew would sit inside the timed region of every rank that has to wait for the slowest
    # one at the first exchange — at --steps 20 on 8 GPUs it was 18 ms of a 1 ms region)
    sampler = ClockSampler(physical_gpu_index(local_rank))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_collectives = 0
    torch.cuda.synchronize(dev)
    ctx.barrier()
    sampler.start()
    e0.record()
    for i in range(args.steps):
        graphs[(args.warmup + i) % N_SETS].replay()
        if i % STATS_EVERY == STATS_EVERY - 1:
            # the one collective of the path, on a side stream so the step kernels keep running
            tick.record(main_stream)
            side.wait_event(tick)
            with torch.cuda.stream(side):
                job_stats[0] = all_reduce_stats(shared_stats, out=stats_buf)
            n_collectives += 1
    main_stream.wait_stream(side)               # the periodic exchanges are complete inside the timed region
    e1.record()
    ctx.barrier()
    job_stats[0] = all_reduce_stats(shared_stats, out=stats_buf)   # the window's totals for the report (not timed)
    clocks = sampler.finish()
    ms_total = ctx.max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    launches = args.steps                       # one ev_step_main_kernel<1> per step (per rank)
    occ1 = occupancy()
    st = job_stats[0].cpu().numpy()
    achieved = ALGO_BYTES_PER_ENV_STEP * B / (ms_step * 1e-3) / 1e9

    # ---- extra: the two step kernels alone, CUDA-event pairs over ONE WHOLE EPISODE of every env set -------------
    # (chunks queued behind ~20 ms of filler work so the host is always ahead and no launch gap leaks into a pair)
    CH = 96
    L = 288 * N_SETS
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(L)]
    filler = torch.empty(1 << 28, dtype=torch.float32, device=dev)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(); filler.normal_(); f1.record()
    torch.cuda.synchronize()
    n_fill = int(np.ceil(20.0 / max(f0.elapsed_time(f1), 0.05)))
    for c0 in range(0, L, CH):
        for _ in range(n_fill):
            filler.normal_()
        for i in range(c0, c0 + CH):
            env, a = envs[i % N_SETS], acts[i % N_SETS]
            evs[i][0].record()
            env.launch_step(a)
            evs[i][1].record()
        torch.cuda.synchronize()
    del filler
    durs = sorted(a_.elapsed_time(b_) for a_, b_ in evs)
    kern_events_ms = float(np.mean(durs[L // 20: L - L // 20]))     # 5 % trimmed mean over one episode

    # ---- e2e: the public host API (the SB3 VecEnv adapter INTEGRATION.md tells users to use), pinned host actions in,
    # host obs/reward/done out, every step.  step_async / step_wait on two env batches: while one batch's results are
    # copied to the host and expanded, the other batch's actions are copied to the device and its kernel runs.
    from sustaingym_b200.envs.vector import SB3VecEnv
    E = 240
    D = N_SETS                                   # env batches in flight (their day phases average the day, like `value`)
    pipe = [SB3VecEnv(e, dict_obs=False, copy=False) for e in envs]
    h_act = torch.rand(4, B, N_STATIONS, dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    for i in range(2 * D):
        pipe[i % D].step_async(h_act[i % 4]); pipe[i % D].step_wait()
    ctx.barrier()
    stamps = np.zeros(E + 1)
    for v in pipe:
        v.env.host_timing.update(sync=0.0, expand=0.0, steps=0)
    t_async = 0.0
    t0 = time.perf_counter()
    for i in range(D - 1):
        pipe[i].step_async(h_act[i % 4])
    for i in range(E):
        ta = time.perf_counter()
        if i + D - 1 < E:
            pipe[(i + D - 1) % D].step_async(h_act[(i + D - 1) % 4])
        t_async += time.perf_counter() - ta
        obs_h, rew_h, done_h, _ = pipe[i % D].step_wait()
        stamps[i + 1] = time.perf_counter() - t0
    torch.cuda.synchronize()
    step_ms = np.diff(stamps) * 1e3
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * B * E / dt
    assert obs_h.shape == (B, OBS_DIM) and rew_h.shape == (B,)
    h2d = pipe[0].env.h2d_bytes_per_step
    d2h = pipe[0].env.d2h_bytes_per_step
    e2e_threads = pipe[0].env.host_threads
    e2e_step_ms = {"p50": float(np.median(step_ms)), "p90": float(np.percentile(step_ms, 90)), "max": float(step_ms.max()),
                   "n_over_2ms": int((step_ms > 2.0).sum()), "step_async": 1e3 * t_async / E,
                   "stream_sync": 1e3 * sum(v.env.host_timing["sync"] for v in pipe) / E,
                   "after_sync": 1e3 * sum(v.env.host_timing["expand"] for v in pipe) / E}

    cpu_base = None
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    if want_cpu:
        Bc, Sc = 4096, 1200                      # ~10 s of single-core CPU work
        t0 = time.perf_counter()
        rate = cpu_oracle_rate(1, Bc, Sc)
        cpu_base = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                    "sample": f"NumPy oracle (oracle/ev.py EVOracle, projection on), {Bc} envs x {Sc} steps, quarters "
                              f"pre-rolled to t = {list(PHASES)}, 1 process, wall incl. pre-roll {time.perf_counter() - t0:.1f}s; "
                              f"host has {os.cpu_count()} cores"}

    del graphs, envs, pipe
    torch.cuda.empty_cache()
    secondary = None
    if not args.no_secondary:
        secondary = {"market_config3": sec_market(ctx, hbm, want_cpu),
                     "building_config4": sec_building(ctx, hbm, want_cpu),
                     "cogen_config5": sec_cogen(ctx, pk, want_cpu)}

    if rank == 0:
        traffic, traffic_src = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"env-sharded x{world}", "trace_pool_days": POOL_DAYS,
                       "mean_sessions_per_day": MEAN_SESSIONS,
                       "day_phases_at_start": list(PHASES),
                       "mean_occupied_stations": 0.5 * (occ0 + occ1),
                       "n_projected": st[6], "n_projected_per_step": st[6] / max(args.steps, 1),
                       "l2_policy": f"{N_SETS} env sets stepped round-robin: ~{N_SETS * 85} MB touched between reuses (> 126 MB L2)",
                       "launch": "one CUDA graph per step holding ONE kernel, ev_step_main_kernel<1>: tile CTAs + 4 helper CTAs (the helpers "
                                 "run the interior-point projections the tile warps hand over through a work list)",
                       "collective": f"all_reduce_stats (one all-gather of 16 doubles per rank, reduced locally: SUM / MIN / MAX) every {STATS_EVERY} steps on a side stream, "
                                     f"{n_collectives} inside the timed region"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": achieved / hbm, "traffic": traffic, "kernel": "ev_step_main_kernel<1>",
                         "traffic_source": traffic_src,
                         "kernel_ms": ms_step,
                         "kernel_ms_note": "step time of the timed window itself (kernel + graph launch), the same window as `value`",
                         "kernel_ms_event_pairs_whole_episode": kern_events_ms,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_ENV_STEP * B,
                         "peak_source": peak_src},
            "cpu_baseline": cpu_base,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": E, "batches_in_flight": D, "host_threads_per_rank": e2e_threads,
                    "step_ms": e2e_step_ms,
                    "api": "SB3VecEnv(BatchedEVChargingEnv).step_async / step_wait -> " + pipe_api_note()},
            "gpu_launches": launches, "clocks": clocks,
            "reward_stats": {"count": st[0], "mean": st[1] / max(st[0], 1), "min": st[3], "max": st[4], "n_done": st[5],
                             "n_projected": st[6], "n_maxiter": st[7]},
            "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def pipe_api_note() -> str:
    from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
    return BatchedEVChargingEnv.HOST_PATH_NOTE


if __name__ == "__main__":
    main()
