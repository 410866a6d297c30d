"""Checkpoint / resume for the batched envs (the reference's envs are plain Python objects that
users pickle; here the state lives in device tensors).

``state_dict()`` returns host copies of every tensor that defines the simulation state (plus the host-side counters an
env lists), ``load_state_dict()`` writes them back in place — the tensors' addresses do not change, so captured CUDA
graphs and compact-observation buffers stay valid.  A checkpoint restores the envs bit for bit: the same actions after
``load_state_dict`` give the same observations, rewards and internal state as the run the checkpoint was taken from
(``tests/test_state_dict_gpu.py``).  Static tables (trace pools, MOER, weather, model weights) are NOT part of it: a
checkpoint is loaded into an env built with the same arguments; shapes and dtypes are checked.

``SG_NVTX=1`` wraps every ``step`` / ``reset`` / ``rollout`` of the batched envs in an NVTX range (``sg:<env>.<call>``) for
timeline profilers."""
from __future__ import annotations

import contextlib
import os

import torch

_NVTX = os.environ.get("SG_NVTX", "0") not in ("", "0")


@contextlib.contextmanager
def nvtx_range(name: str):
    if not _NVTX:
        yield
        return
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


def nvtx_method(name: str):
    """Decorator: the method runs inside an NVTX range when SG_NVTX is set; otherwise the method is left untouched."""
    def deco(fn):
        if not _NVTX:
            return fn

        def wrapped(*a, **k):
            with nvtx_range(name):
                return fn(*a, **k)
        wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
        return wrapped
    return deco


class StateDictMixin:
    """Batched envs list their state in ``_STATE_TENSORS`` (attribute names of device tensors, ``None`` entries are
    skipped) and ``_STATE_ATTRS`` (plain host attributes)."""

    _STATE_TENSORS: tuple = ()
    _STATE_ATTRS: tuple = ()

    def state_dict(self) -> dict:
        torch.cuda.synchronize(self.device)
        sd = {"__class__": type(self).__name__, "num_envs": int(self.num_envs)}
        for name in self._STATE_TENSORS:
            t = getattr(self, name, None)
            if t is not None:
                sd[name] = t.detach().to("cpu", copy=True)
        for name in self._STATE_ATTRS:
            sd["attr:" + name] = getattr(self, name)
        gen = getattr(self, "_gen", None)
        if isinstance(gen, torch.Generator):
            sd["rng:_gen"] = gen.get_state()
        return sd

    def load_state_dict(self, sd: dict) -> None:
        if sd.get("__class__") != type(self).__name__ or int(sd.get("num_envs", -1)) != int(self.num_envs):
            raise ValueError(f"checkpoint of {sd.get('__class__')} x {sd.get('num_envs')} envs loaded into "
                             f"{type(self).__name__} x {self.num_envs} envs")
        for name in self._STATE_TENSORS:
            t = getattr(self, name, None)
            if t is None:
                continue
            if name not in sd:
                raise KeyError(f"checkpoint has no '{name}'")
            src = sd[name]
            if tuple(src.shape) != tuple(t.shape) or src.dtype != t.dtype:
                raise ValueError(f"'{name}': checkpoint {tuple(src.shape)} {src.dtype}, env {tuple(t.shape)} {t.dtype}")
            t.copy_(src)                                   # in place: device addresses (graphs, C structs) stay valid
        for name in self._STATE_ATTRS:
            if "attr:" + name in sd:
                setattr(self, name, sd["attr:" + name])
        gen = getattr(self, "_gen", None)
        if isinstance(gen, torch.Generator) and "rng:_gen" in sd:
            gen.set_state(sd["rng:_gen"])
        torch.cuda.synchronize(self.device)
