"""Where the end-to-end (host buffers in / out) step time goes: enqueue, stream wait, host expansion — by host
thread count and for the dense path.  One GPU."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from sustaingym_b200 import _lib
from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator

B = 65536
gen = SyntheticTraceGenerator("caltech", num_days=256, mean_sessions=25.0, seed=1234)
h_act = torch.rand(4, B, 54).pin_memory()


def run(host_obs, threads, E=120, pre=120):
    envs = [BatchedEVChargingEnv(gen, B, auto_reset=True, host_obs=host_obs, host_threads=threads) for _ in range(2)]
    d_act = h_act[0].cuda()
    for e in envs:
        e.reset(seed=0)
        for _ in range(pre):
            e.launch_step(d_act)
    torch.cuda.synchronize()
    T = {"enqueue": 0.0, "wait": 0.0, "expand": 0.0}
    lib = _lib.load()
    orig = lib.sg_ev_expand_obs

    for i in range(4):
        envs[i % 2].step_async(h_act[i % 4]); envs[i % 2].step_wait()
    t_all = time.perf_counter()
    t0 = time.perf_counter(); envs[0].step_async(h_act[0]); T["enqueue"] += time.perf_counter() - t0
    for i in range(E):
        t0 = time.perf_counter(); envs[(i + 1) % 2].step_async(h_act[(i + 1) % 4]); T["enqueue"] += time.perf_counter() - t0
        e = envs[i % 2]
        t0 = time.perf_counter(); e._async_stream.synchronize(); T["wait"] += time.perf_counter() - t0
        t0 = time.perf_counter(); e.step_wait(); T["expand"] += time.perf_counter() - t0
    envs[E % 2].step_wait()
    dt = time.perf_counter() - t_all
    print(f"{host_obs:8s} threads={threads:2d}: {B * (E + 1) / dt:.3e} env-steps/s  {dt / (E + 1) * 1e3:.3f} ms/step | per step: "
          + "  ".join(f"{k} {v / E * 1e3:.3f} ms" for k, v in T.items()), flush=True)


print("cpus", os.cpu_count(), "torch threads", torch.get_num_threads())
run("dense", 1)
for inc in ("0", "1"):
    os.environ["SG_HOST_INCREMENTAL"] = inc
    print("SG_HOST_INCREMENTAL =", inc, "(0: full rows, streaming stores; 1: changed entries, cached stores)")
    for th in (4, 8, 12, 16):
        run("compact", th)


def run_adapter(threads, E=200):
    """the bench's e2e loop: SB3VecEnv adapters over two env batches"""
    from sustaingym_b200.envs.vector import SB3VecEnv
    envs = [BatchedEVChargingEnv(gen, B, auto_reset=True, host_threads=threads, collect_stats=True) for _ in range(2)]
    d_act = h_act[0].cuda()
    for e in envs:
        e.reset(seed=0)
        for _ in range(120):
            e.launch_step(d_act)
    torch.cuda.synchronize()
    pipe = [SB3VecEnv(e, dict_obs=False, copy=False) for e in envs]
    inner = {"env.step_wait": 0.0}
    for e in envs:                                   # time the engine call inside the adapter
        orig = e.step_wait
        def timed(orig=orig):
            t0 = time.perf_counter(); r = orig(); inner["env.step_wait"] += time.perf_counter() - t0; return r
        e.step_wait = timed
    for i in range(4):
        pipe[i % 2].step_async(h_act[i % 4]); pipe[i % 2].step_wait()
    T = {"async": 0.0, "wait": 0.0}
    t_all = time.perf_counter()
    pipe[0].step_async(h_act[0])
    for i in range(E):
        t0 = time.perf_counter(); pipe[(i + 1) % 2].step_async(h_act[(i + 1) % 4]); T["async"] += time.perf_counter() - t0
        t0 = time.perf_counter(); pipe[i % 2].step_wait(); T["wait"] += time.perf_counter() - t0
    pipe[E % 2].step_wait()
    dt = time.perf_counter() - t_all
    print(f"adapter  threads={threads:2d}: {B * (E + 1) / dt:.3e} env-steps/s  {dt / (E + 1) * 1e3:.3f} ms/step | per step: "
          + "  ".join(f"{k} {v / E * 1e3:.3f} ms" for k, v in T.items()) + f"  env.step_wait {inner['env.step_wait'] / (E + 5) * 1e3:.3f} ms", flush=True)


os.environ["SG_HOST_INCREMENTAL"] = "0"
run_adapter(16)
run_adapter(12)
