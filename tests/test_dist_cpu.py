"""CPU: the multi-process host logic (env sharding + the statistics all-reduce) on gloo, world_size 2."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sustaingym_b200.dist import all_reduce_stats, shard_range, summarize


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(1001, rank, world)
    rng = np.random.default_rng(7)
    rewards = rng.normal(size=1001)[lo:hi]                       # this rank's envs
    st = torch.zeros(16, dtype=torch.float64)
    st[0], st[1], st[2] = len(rewards), rewards.sum(), (rewards ** 2).sum()
    st[3], st[4], st[5] = rewards.min(), rewards.max(), rank + 1
    tot = all_reduce_stats(st)
    q.put((rank, lo, hi, tot.numpy(), st.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition():
    for n, w in [(65536, 8), (1001, 2), (5, 8), (262144, 8)]:
        r = [shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_stats_all_reduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rewards = np.random.default_rng(7).normal(size=1001)
    for rank, lo, hi, tot, local in res:
        assert tot[0] == 1001 and abs(tot[1] - rewards.sum()) < 1e-9 and abs(tot[2] - (rewards ** 2).sum()) < 1e-9
        assert tot[3] == rewards.min() and tot[4] == rewards.max() and tot[5] == 3
        assert local[0] == hi - lo                                # input untouched
    s = summarize(torch.from_numpy(res[0][3]))
    assert abs(s["mean"] - rewards.mean()) < 1e-12 and abs(s["std"] - rewards.std()) < 1e-9


def test_single_process_passthrough():
    st = torch.arange(16, dtype=torch.float64)
    assert torch.equal(all_reduce_stats(st), st)
