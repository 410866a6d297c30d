"""Secondary benchmarks (BASELINE.json configs 3-5): BuildingEnv, CogenEnv, ElectricityMarketEnv on one
GPU at their per-GPU batch, CUDA-event timed, with the roofline each kernel is bound by and the NumPy
oracle timed beside it.  The headline (config 2, EV) is bench.py.  One JSON line per env.

    python profiles/bench_envs.py [building] [cogen] [cogen32k] [market] [ev_gmm]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = float(PEAKS["hbm_gbs"])
FFMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12        # fp32 FMA peak at the max SM clock


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps          # ms per call


def bench_building():
    from oracle.building import BuildingOracle, BuildingParams, discretize
    from sustaingym_b200.envs.building.env import BatchedBuildingEnv
    g = np.load(os.path.join(ROOT, "tests", "golden", "building_retail_s5.npz"))
    n, T = 5, 105108
    t = np.arange(T) / 288.0
    params = dict(n=n, A_d=g["A_d"], BD_d=g["BD_d"], target=g["target"], ac_map=g["ac_map"],
                  out_temp=20.5 + 21.5 * np.sin(2 * np.pi * (t - 0.3)) * np.sin(np.pi * t / 365) ** 2,
                  ground_temp=np.full(T, 20.9), ghi=np.clip(np.sin(2 * np.pi * (t - 0.25)), 0, 1),
                  metabolism=np.full(T, 120.0), reward_beta=0.999, reward_pnorm=2, episode_len=288,
                  time_resolution=300, max_power=8000, is_continuous_action=True, temp_range=(-40, 40))
    B, S = 32768, 288                     # config 4: 262144 envs over 8 GPUs
    sets = []
    for i in range(4):                    # 4 x (B*S*n*4 = 189 MB of actions): far beyond L2
        env = BatchedBuildingEnv(params, B, auto_reset=True)
        env.reset(epochs=np.random.default_rng(i).integers(0, T - 289, B))
        acts = torch.rand(S, B, n, device="cuda") * 2 - 1
        sets.append((env, acts))
    it = [0]

    def roll():
        env, acts = sets[it[0] % 4]; it[0] += 1
        env.rollout(acts, keep_obs=True)
    ms = timed(roll, 12, 4)
    steps_per_s = B * S / (ms * 1e-3)
    algo = 105 * B * S
    # single-step launches for comparison (launch-latency bound)
    env, acts = sets[0]
    g1 = env.capture_graph(acts[0])
    ms1 = timed(g1.replay, 400, 20)
    # CPU oracle
    p = BuildingParams(n=n, A_d=g["A_d"], BD_d=g["BD_d"], target=g["target"], ac_map=g["ac_map"], out_temp=params["out_temp"],
                       ground_temp=params["ground_temp"], ghi=params["ghi"], metabolism=params["metabolism"],
                       q_rate=0.024, error_rate=0.999)
    orc = BuildingOracle(p, 4096); orc.reset(np.arange(4096))
    a = np.random.default_rng(0).uniform(-1, 1, (4096, n)).astype(np.float32)
    t0 = time.perf_counter()
    for _ in range(100):
        orc.step(a)
    cpu = 4096 * 100 / (time.perf_counter() - t0)
    return {"env": "BuildingEnv", "metric": "env_steps_per_sec", "value": steps_per_s, "unit": "env-steps/s",
            "config": {"workload": "BuildingEnv RetailStandalone n=5, 32768 envs/GPU (262144 over 8), 288-step fused rollout per launch, obs kept",
                       "l2_policy": "4 env sets round-robin, 189 MB of actions + 340 MB of obs per launch"},
            "ms_per_launch": ms, "single_step_launch_ms": ms1, "single_step_env_steps_per_sec": B / (ms1 * 1e-3),
            "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": HBM, "unit": "GB/s",
                         "frac": algo / (ms * 1e-3) / 1e9 / HBM, "traffic": None, "kernel": "building_step_kernel<5,false>",
                         "algorithmic_bytes_per_launch": algo},
            "cpu_baseline": {"value": cpu, "unit": "env-steps/s", "cores": 1, "kind": "port",
                             "sample": "NumPy oracle, 4096 envs x 100 steps (reference class under shim measured 14557 steps/s in the build container, BASELINE.md)"}}


def bench_cogen(B=4096):
    from oracle.cogen import CogenOracle
    from sustaingym_b200.data.cogen_ambients import synthetic_ambients
    from sustaingym_b200.envs.cogen.env import BatchedCogenEnv
    # config 5: 32768 envs over 8 GPUs = 4096 per GPU; 'cogen32k' runs the whole config on one GPU
    amb = synthetic_ambients(256, seed=0)
    env = BatchedCogenEnv(B, amb, forecast_horizon=12, forecast_noise_std=0.0)
    rng = np.random.default_rng(0)
    env.reset(days=rng.integers(0, 200, B), init_actions=env.sample_actions(rng))
    acts = [torch.from_numpy(env.sample_actions(rng)).cuda() for _ in range(4)]
    it = [0]

    def step():
        env.step(acts[it[0] % 4]); it[0] += 1
        if it[0] % 96 == 0:
            env.day_t[:, 1] = 0
    ms = timed(step, 300 if B <= 4096 else 60, 30 if B <= 4096 else 10)
    flops = 3684096 * B
    # the FFMA kernel on the same workload, for the record
    env_f = BatchedCogenEnv(B, amb, forecast_horizon=12, forecast_noise_std=0.0, tensor_cores=False)
    env_f.reset(days=rng.integers(0, 200, B), init_actions=env_f.sample_actions(rng))

    def step_f():
        env_f.step(acts[it[0] % 4]); it[0] += 1
        if it[0] % 96 == 0:
            env_f.day_t[:, 1] = 0
    ms_f = timed(step_f, 100 if B <= 4096 else 20, 10 if B <= 4096 else 4)
    tf32_peak = float(PEAKS.get("bf16_tflops", 1590.0)) / 2.0        # dense TF32 = half the measured bf16 rate
    orc = CogenOracle(amb, 1024, forecast_horizon=12)
    a = env.sample_actions(rng)[:1024]
    orc.reset(rng.integers(0, 200, 1024), a)
    t0 = time.perf_counter()
    for _ in range(5):
        orc.step(a)
    cpu = 1024 * 5 / (time.perf_counter() - t0)
    return {"env": "CogenEnv", "metric": "env_steps_per_sec", "value": B / (ms * 1e-3), "unit": "env-steps/s",
            "config": {"workload": f"CogenEnv 4 agents k=12, {B} envs on one GPU (config 5: 32768 over 8 GPUs), one step = inputs + plant (tcgen05) + finish kernels"},
            "ms_per_step": ms,
            "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                         "frac": flops / (ms * 1e-3) / 1e12 / tf32_peak, "traffic": None, "kernel": "cogen_mlp_tc_kernel",
                         "algorithmic_flops_per_launch": flops,
                         "note": "tcgen05 kind::tf32, 3xTF32 split (the tensor pipe executes 3x the algorithmic FLOPs); "
                                 "peak = measured cuBLAS bf16 / 2 (MEASURED_PEAKS.json); layer 1 + activations run on the FFMA/SFU pipes",
                         "ffma_kernel_ms_per_step": ms_f, "speedup_over_ffma_kernel": ms_f / ms},
            "cpu_baseline": {"value": cpu, "unit": "env-steps/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": "NumPy oracle (reference ONNX weights), 1024 envs x 5 steps, BLAS threads"}}


def bench_market():
    from oracle.market import MarketOracle, MarketParams
    from sustaingym_b200.envs.electricitymarket.env import BatchedElectricityMarketEnv, default_fleet, synthetic_load_moer
    B = 65536
    cost, cap = default_fleet()
    load, moer = synthetic_load_moer(64, 1, seed=1)
    envs = []
    for i in range(8):
        env = BatchedElectricityMarketEnv(B, load, moer, cost, cap)
        env.reset(seed=i)
        a = torch.rand(B, 2, device="cuda") * 120
        envs.append((env, a, env.capture_graph(a)))
    it = [0]

    def step():
        env, a, g = envs[it[0] % 8]; it[0] += 1
        g.replay()
        if it[0] % (8 * 280) == 0:
            for e, _, _ in envs:
                e.day_t[:, 1] = 0
    ms = timed(step, 2000, 100)
    algo = (8 + 4 * 8 * 2 + 8 + 8 + 40 + 8 + 1 + 24) * B      # action, state r/w, day_t, prev_a, obs, reward, done, info
    orc = MarketOracle(MarketParams(cost, cap), load, moer, 16384)
    orc.reset(np.arange(16384) % 64)
    a = np.random.default_rng(0).uniform(0, 120, (16384, 2)).astype(np.float32)
    t0 = time.perf_counter()
    for _ in range(50):
        orc.step(a)
    cpu = 16384 * 50 / (time.perf_counter() - t0)
    return {"env": "ElectricityMarketEnv (spec-defined)", "metric": "env_steps_per_sec", "value": B / (ms * 1e-3), "unit": "env-steps/s",
            "config": {"workload": "ElectricityMarketEnv spec v1, 33 generators, 1 battery, k=1, 65536 envs, one step per CUDA-graph replay",
                       "l2_policy": "8 env sets round-robin"},
            "ms_per_step": ms,
            "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": HBM, "unit": "GB/s",
                         "frac": algo / (ms * 1e-3) / 1e9 / HBM, "traffic": None, "kernel": "market_step_kernel",
                         "algorithmic_bytes_per_launch": algo},
            "cpu_baseline": {"value": cpu, "unit": "env-steps/s", "cores": 1, "kind": "port", "sample": "NumPy oracle, 16384 envs x 50 steps"}}


def bench_ev_gmm():
    """Reset-path operator: GMM day-trace sampler, one pool of days per launch."""
    from oracle import ev_gmm as G
    from sustaingym_b200.envs.evcharging.event_generation import GmmDevice, load_gmm
    gmm = load_gmm("caltech", "Summer 2019")
    D, K, smax = 65536, 8, 128
    dev = GmmDevice(gmm, 54, "cuda")
    pools = [dev.alloc_pool(D, K, smax) for _ in range(2)]      # 2 x 485 MB of tables: beyond L2
    it = [0]
    kwh = 32 * (5 / 60 * 0.208)

    def sample():
        dev.sample(pools[it[0] % 2], it[0] * D, 1, 100.0, kwh); it[0] += 1
    ms = timed(sample, 10, 3)
    algo = D * (K * 54 * 16 + (smax + 1) * 4 + 12)             # tables written per launch
    sessions = int((pools[0]["arrivals"] != 0x7FFF).sum().item())
    t0 = time.perf_counter()
    G.sample_days(gmm, 1, 64, 54, 100.0, K, smax, kwh)
    cpu = 64 / (time.perf_counter() - t0)
    return {"env": "EVChargingEnv reset path (GMM trace sampler)", "metric": "days_per_sec", "value": D / (ms * 1e-3), "unit": "days/s",
            "config": {"workload": "GMMsTraceGenerator caltech Summer 2019 (30 components), 65536 days per launch, K=8 slots, 128 events/day",
                       "l2_policy": "2 pools of 485 MB round-robin"},
            "ms_per_launch": ms, "sessions_per_launch": sessions,
            "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": HBM, "unit": "GB/s",
                         "frac": algo / (ms * 1e-3) / 1e9 / HBM, "traffic": None, "kernel": "ev_gmm_days_kernel",
                         "algorithmic_bytes_per_launch": algo,
                         "note": "thread per day, serial rejection sampling + station draw: latency bound, not a streaming kernel"},
            "cpu_baseline": {"value": cpu, "unit": "days/s", "cores": 1, "kind": "port",
                             "sample": "pure-Python oracle (bit-exact restatement), 64 days; the reference samples one day per reset with scikit-learn + pandas"}}


def bench_building_large():
    """The reference's largest building (OutPatientHealthCare, n = 118 zones): f64 tensor-core (DMMA) cluster kernel."""
    from sustaingym_b200.envs.building.env import BatchedBuildingEnv
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import load_building_fixture
    g, params = load_building_fixture(os.path.join(ROOT, "tests", "golden", "building_outpatient_s3.npz"))
    n, B, S = int(params["n"]), 8192, 96
    sets = []
    for i in range(3):
        env = BatchedBuildingEnv(params, B, auto_reset=True)
        env.reset(epochs=np.random.default_rng(i).integers(0, 1500, B))
        sets.append((env, torch.rand(S, B, n, device="cuda") * 2 - 1))
    it = [0]

    def roll():
        env, acts = sets[it[0] % 3]; it[0] += 1
        env.rollout(acts, keep_obs=True)
    ms = timed(roll, 6, 2)
    algo = (4 * n + 4 * n + 4 * (n + 4) + 16 + 8 + 4 + 1) * B * S
    flops = 2 * n * (2 * n + 4) * B * S
    return {"env": "BuildingEnv (n=118)", "metric": "env_steps_per_sec", "value": B * S / (ms * 1e-3), "unit": "env-steps/s",
            "config": {"workload": f"BuildingEnv OutPatientHealthCare n={n}, {B} envs, {S}-step fused rollout per launch, obs kept",
                       "l2_policy": "3 env sets round-robin (371 MB of actions + 383 MB of obs per launch)"},
            "ms_per_launch": ms,
            "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": 40.0, "unit": "TFLOP/s",
                         "frac": flops / (ms * 1e-3) / 1e12 / 40.0, "traffic": None, "kernel": "building_mma_kernel<4>",
                         "algorithmic_flops_per_launch": flops, "hbm_gbs": algo / (ms * 1e-3) / 1e9,
                         "note": "f64 DMMA (mma.sync m8n8k4) on 2-CTA clusters; peak = NVIDIA's nominal 40 TFLOP/s f64 tensor rate for B200 "
                                 "(MEASURED_PEAKS.json holds no f64 figure); thread-per-env generic kernel on the same workload: 94.2 ms per launch"},
            "cpu_baseline": None}


if __name__ == "__main__":
    which = sys.argv[1:] or ["building", "cogen", "market"]
    for w in which:
        print(json.dumps({"building": bench_building, "cogen": bench_cogen, "cogen32k": lambda: bench_cogen(32768),
                          "market": bench_market, "ev_gmm": bench_ev_gmm, "building118": bench_building_large}[w]()), flush=True)
