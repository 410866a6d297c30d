"""Concurrent pinned host<->device copy bandwidth on N GPUs of one node (torchrun), with the CPU / NUMA placement of each
rank: names the limiter of the end-to-end (host-buffer) path at N > 1.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 profiles/r2/pcie_probe.py
"""
import glob
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)


def gpu_numa(idx: int):
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[idx]) if vis else idx
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        path = f"/sys/bus/pci/devices/{bus[-12:]}/numa_node"
        return bus, int(open(path).read()) if os.path.exists(path) else None
    except Exception as e:  # noqa: BLE001
        return str(e), None


def node_cpus(node):
    p = f"/sys/devices/system/node/node{node}/cpulist"
    if node is None or node < 0 or not os.path.exists(p):
        return None
    out = set()
    for part in open(p).read().strip().split(","):
        a, _, b = part.partition("-")
        out |= set(range(int(a), int(b or a) + 1))
    return out


MB = 1 << 20
SZ = 38 * MB                      # one dense observation batch of 65 536 envs
h_in = torch.empty(SZ, dtype=torch.uint8).pin_memory()
h_out = torch.empty(SZ, dtype=torch.uint8).pin_memory()
d_a, d_b = torch.empty(SZ, dtype=torch.uint8, device=dev), torch.empty(SZ, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def barrier():
    if world > 1:
        dist.barrier(device_ids=[local])
    torch.cuda.synchronize()


def run(mode: str, secs: float = 0.8) -> float:
    barrier()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < secs:
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s1):
                h_out.copy_(d_a, non_blocking=True)
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s2):
                d_b.copy_(h_in, non_blocking=True)
        s1.synchronize(); s2.synchronize()
        n += 1
    dt = time.perf_counter() - t0
    per = n * SZ / dt / 1e9 * (2 if mode == "both" else 1)
    t = torch.tensor([per], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t)
    return per, float(t.item())


bus, numa = gpu_numa(local)
aff = sorted(os.sched_getaffinity(0))
nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
info = {"rank": rank, "gpu_bus": bus, "gpu_numa_node": numa, "cpu_affinity": f"{aff[0]}-{aff[-1]} ({len(aff)} cpus)", "numa_nodes": nodes}
res = {}
for mode in ("d2h", "h2d", "both"):
    res[mode] = run(mode)
# same with the rank pinned to the CPUs of its GPU's NUMA node (when the cpuset allows it)
local_cpus = node_cpus(numa)
pinned = None
if local_cpus and (local_cpus & set(aff)):
    try:
        os.sched_setaffinity(0, local_cpus & set(aff))
        pinned = {m: run(m) for m in ("d2h", "both")}
    except OSError as e:
        pinned = str(e)
gathered = [None] * world
if world > 1:
    dist.all_gather_object(gathered, info)
else:
    gathered = [info]
if rank == 0:
    print(json.dumps({"n_gpus": world, "ranks": gathered,
                      "GBps_per_rank0_and_aggregate": {m: {"rank0": round(v[0], 1), "aggregate": round(v[1], 1)} for m, v in res.items()},
                      "pinned_to_gpu_numa_node": ({m: {"rank0": round(v[0], 1), "aggregate": round(v[1], 1)} for m, v in pinned.items()}
                                                  if isinstance(pinned, dict) else pinned)}), flush=True)
if world > 1:
    dist.destroy_process_group()
