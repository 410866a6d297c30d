"""Host-side microbenchmark of sg_ev_expand_obs (compact records -> dense [B,146] f32 observation) and of plain
memory writes, by thread count.  No GPU needed."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sustaingym_b200 import _lib

REC = np.dtype([("active", "<u8"), ("reward", "<f8"), ("t", "<i2"), ("moer_day", "<i2"), ("done", "u1"), ("flags", "u1"),
                ("n_active", "u1"), ("status", "u1"), ("demand", "<f4", 16), ("est", "<i2", 16)])
lib = _lib.load()
rng = np.random.default_rng(0)
B, n, k, Dm = 65536, 54, 36, 64
moer = rng.uniform(0, 1, (Dm, 289, 37)).astype(np.float32)


def fast(B):
    rec = np.zeros(B, REC)
    occ = rng.uniform(size=(B, n)) < 0.11
    rec["active"] = (occ * (1 << np.arange(n, dtype=np.uint64))).sum(1).astype(np.uint64)
    rec["t"] = rng.integers(0, 289, B); rec["moer_day"] = rng.integers(0, Dm, B)
    rec["demand"] = rng.uniform(1, 50, (B, 16)); rec["est"] = rng.integers(0, 200, (B, 16))
    rec["flags"] = occ.sum(1) > 16
    return rec


bufs = []
for _ in range(4):
    b = np.zeros(16 + B * 120, np.uint8); b[16:] = fast(B).view(np.uint8); bufs.append(b)
obs = np.zeros((B, 146), np.float32); rew = np.zeros(B); done = np.zeros(B, np.uint8); prev = np.zeros(B, np.uint64)
p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for thr in [1, 2, 4, 8, 16, 32]:
    if thr > (os.cpu_count() or 1):
        break
    for inc in [0, 1]:
        lib.sg_ev_expand_obs(p(bufs[0]), B, n, k, p(moer), Dm, p(prev), 0, None, None, 0, p(obs), p(rew), p(done), None, thr)
        t0 = time.perf_counter()
        for i in range(30):
            lib.sg_ev_expand_obs(p(bufs[i % 4]), B, n, k, p(moer), Dm, p(prev), inc, None, None, 0, p(obs), p(rew), p(done), None, thr)
        dt = (time.perf_counter() - t0) / 30
        print(f"expand threads={thr:2d} incremental={inc}: {dt * 1e3:7.3f} ms/step  {B / dt:.3e} env-steps/s", flush=True)
