"""CUDA-graph capture of one batched ``step()`` (device path): the Python / ctypes launch path costs
~25 us per call, a graph replay ~3 us, which matters for the envs whose step kernel is shorter than that
(BuildingEnv, ElectricityMarketEnv at BASELINE batch sizes)."""
from __future__ import annotations

import torch


def capture_step_graph(env, actions: torch.Tensor) -> "torch.cuda.CUDAGraph":
    """Captures ``env.step(actions)`` (reading ``actions`` in place, writing the env's own output tensors)
    into a CUDA graph; ``graph.replay()`` then advances all envs with a single graph launch."""
    env.step(actions)                                   # first launch outside capture (module load, attributes)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(env.device)
    s.wait_stream(torch.cuda.current_stream(env.device))
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            env.step(actions)
    torch.cuda.current_stream(env.device).wait_stream(s)
    return g
