"""Batched EVChargingEnv on the B200 engine.

``BatchedEVChargingEnv`` advances ``num_envs`` independent copies of the reference's
``EVChargingEnv`` (``sustaingym/envs/evcharging/env.py``) in lock-step with one kernel launch per
``step()``; ``EVChargingEnv`` is the ``num_envs = 1`` view with the reference's exact Gymnasium
signature (``reset(*, seed, options) -> (obs, info)``, ``step(action) -> 5-tuple``).

All compute happens in ``libsustaingym_b200.so`` (``sg_ev_step`` / ``sg_ev_reset``); this module is
host plumbing: it owns the device tensors and mirrors the reference API.
"""
from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np
import torch

from ... import _lib, spaces
from .._state import StateDictMixin, nvtx_method
from .event_generation import AbstractTraceGenerator
from .network import ChargingNetworkSpec

OBS_KEYS = ("demands", "est_departures", "forecasted_moer", "prev_moer", "timestep")  # sorted (flatten order)


class BatchedEVChargingEnv(StateDictMixin):
    """See module docstring.  Mirrors the constructor of the reference (env.py:116-119) plus
    ``num_envs``/``device``/``auto_reset``.

    Observations are returned flat, ``[num_envs, 2n+k+2]`` f32 in gymnasium's flatten order
    (``demands | est_departures | forecasted_moer | prev_moer | timestep``); ``obs_dict()`` gives
    the Dict view (slices of the same storage).
    """

    _STATE_TENSORS = ('slots', 'header', 'obs', 'final_obs', 'reward', 'done', 'status', 'pilots', 'stats')          # checkpointed by state_dict() (envs/_state.py)

    def load_state_dict(self, sd: dict) -> None:
        StateDictMixin.load_state_dict(self, sd)
        self.work.zero_()                              # scheduler / work-list scratch: at rest between steps
        if self._compact is not None:                  # incremental host expansion: forget what the host rows held
            self._compact["prev"].fill_((1 << self.num_stations) - 1)


    # env.py:99-114
    TIMESTEP_DURATION = 5
    ACTION_SCALE_FACTOR = 32
    VOLTAGE = 208
    MARGINAL_REVENUE_PER_KWH = 0.15
    OPERATING_MARGIN = 0.20
    MARGINAL_PROFIT_PER_KWH = MARGINAL_REVENUE_PER_KWH * OPERATING_MARGIN
    CO2_COST_PER_METRIC_TON = 30.85
    A_MINS_TO_KWH = (1 / 60) * (VOLTAGE / 1000)
    VIOLATION_WEIGHT = 0.001
    A_PERS_TO_KWH = A_MINS_TO_KWH * TIMESTEP_DURATION
    PROFIT_FACTOR = A_PERS_TO_KWH * MARGINAL_PROFIT_PER_KWH
    VIOLATION_FACTOR = A_PERS_TO_KWH * VIOLATION_WEIGHT
    CARBON_COST_FACTOR = A_PERS_TO_KWH * (CO2_COST_PER_METRIC_TON / 1000)

    HOST_PATH_NOTE = ("BatchedEVChargingEnv.step_async(pinned host actions) / step_wait(): sg_ev_step_host_compact_stream "
                      "(H2D actions -> step -> pack 120-byte records -> D2H -> stream-ordered host function: a thread pool "
                      "rebuilds the dense [B,146] f32 observation, bit-identical to a dense copy)")

    @property
    def h2d_bytes_per_step(self) -> int:
        return self.num_envs * self.num_stations * (8 if self.discrete_bins > 0 else 4)

    @property
    def d2h_bytes_per_step(self) -> int:
        if self.host_obs == "compact":
            dn = self.host_dense_envs
            return (_lib.SG_EV_COMPACT_HEADER_BYTES + (self.num_envs - dn) * _lib.SG_EV_COMPACT_RECORD_BYTES
                    + dn * (self.obs_dim * 4 + 8 + 1 + 1))
        return self.num_envs * (self.obs_dim * 4 + 8 + 1)

    def __init__(self, data_generator: AbstractTraceGenerator, num_envs: int = 1,
                 moer_forecast_steps: int = 36, project_action_in_env: bool = True, verbose: int = 0,
                 device: str | torch.device | None = None, auto_reset: bool = False,
                 discrete_bins: int = -1, collect_stats: bool = False, host_obs: str = "compact",
                 host_threads: int | None = None):
        assert 1 <= moer_forecast_steps <= 36                                   # env.py:120
        self.lib = _lib.load()                                                  # fails loudly if not built
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedEVChargingEnv needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.data_generator = data_generator
        self.num_envs = int(num_envs)
        self.max_timestep = 288
        self.moer_forecast_steps = moer_forecast_steps
        self.project_action_in_env = project_action_in_env
        self.verbose = verbose
        self.auto_reset = auto_reset
        self.discrete_bins = discrete_bins
        self.cn: ChargingNetworkSpec = data_generator.network
        self.num_stations = n = self.cn.num_stations
        self.obs_dim = 2 * n + moer_forecast_steps + 2
        self._async_stream = None
        if host_obs not in ("compact", "dense"):
            raise ValueError("host_obs must be 'compact' (120-byte records over PCIe, expanded on the host) or 'dense'")
        self.host_obs = host_obs
        import os
        if host_threads is None:
            host_threads = int(os.environ.get("SG_HOST_THREADS", 0)) or max(
                1, min(16, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
        self.host_threads = int(host_threads)
        # host expansion mode: False = every row rewritten with streaming stores (default: no read-for-ownership of the
        # 38 MB observation batch); True = only the entries that changed, with ordinary stores (cache-resident batches)
        self.host_incremental = os.environ.get("SG_HOST_INCREMENTAL", "0") == "1"
        # compact host path: the first host_dense_envs envs can cross PCIe as dense rows (DMA) while the others travel
        # as records expanded by the host threads.  Measured on the B200 box (profiles/r2/e2e_notes.md): any split
        # gives the same ~0.6 ms per 65 536-env step — the host memory system (~115 GB/s of total traffic: actions
        # read by the DMA engine, observations written, records written and read) is the limit, not PCIe or the
        # cores — so the default is 0 (all records: least PCIe traffic, result buffers need no pinning)
        frac = float(os.environ.get("SG_HOST_DENSE_FRACTION", "0.0"))
        self.host_dense_envs = int(round(min(max(frac, 0.0), 1.0) * num_envs)) if host_obs == "compact" else 0

        # ---- spaces (env.py:143-150, 171-172; wrappers.py:36-41 for the discrete variant) ---------
        k = moer_forecast_steps
        self.single_observation_space = spaces.Dict({
            "timestep": spaces.Box(0, 1, shape=(1,), dtype=np.float32),
            "est_departures": spaces.Box(-288, 288, shape=(n,), dtype=np.float32),
            "demands": spaces.Box(0, data_generator.requested_energy_cap, shape=(n,), dtype=np.float32),
            "prev_moer": spaces.Box(0, 1, shape=(1,), dtype=np.float32),
            "forecasted_moer": spaces.Box(0, 1, shape=(k,), dtype=np.float32),
        })
        if discrete_bins > 0:
            self.single_action_space = spaces.MultiDiscrete(np.ones(n, dtype=np.int64) * discrete_bins)
        else:
            self.single_action_space = spaces.Box(low=0, high=1.0, shape=(n,), dtype=np.float32)
        self.obs_slices = {"demands": slice(0, n), "est_departures": slice(n, 2 * n),
                           "forecasted_moer": slice(2 * n, 2 * n + k),
                           "prev_moer": slice(2 * n + k, 2 * n + k + 1),
                           "timestep": slice(2 * n + k + 1, 2 * n + k + 2)}

        # ---- device-resident constants ----------------------------------------------------------
        dev = self.device
        cls, coef = self.cn.class_reduce()
        if coef.shape[1] > _lib.SG_EV_MAX_CLASSES or coef.shape[0] > _lib.SG_EV_MAX_CONSTRAINTS:
            raise ValueError("network has more station classes/constraints than the kernels support")
        pool = getattr(data_generator, "device_pool", None)     # GMMsTraceGenerator: tables sampled on the GPU
        if pool is not None:
            if pool["sessions"].device != dev:
                raise ValueError("the generator's device pool lives on another device than the env")
            sessions_dev, arrivals_dev = pool["sessions"], pool["arrivals"]
            K_slots, n_trace_days = data_generator._K, data_generator.num_days
        else:
            t = data_generator.tables
            rec = np.zeros(t.arrival.shape, dtype=np.dtype([("arr", "<i2"), ("dep", "<i2"), ("est", "<i2"),
                                                             ("res", "<i2"), ("energy", "<f8")]))
            rec["arr"], rec["dep"], rec["est"], rec["energy"] = t.arrival, t.departure, t.est_departure, t.energy
            sessions_dev = torch.from_numpy(rec.view(np.uint8).reshape(-1)).to(dev)
            arrivals_dev = torch.from_numpy(t.arrival_table().view(np.int32)).to(dev)
            K_slots, n_trace_days = t.max_sessions, t.num_days
        J, Cn = coef.shape
        re = np.zeros((_lib.SG_EV_MAX_CONSTRAINTS, _lib.SG_EV_MAX_CLASSES)); re[:J, :Cn] = coef.real
        im = np.zeros((_lib.SG_EV_MAX_CONSTRAINTS, _lib.SG_EV_MAX_CLASSES)); im[:J, :Cn] = coef.imag
        mag = np.zeros(_lib.SG_EV_MAX_CONSTRAINTS); mag[:J] = self.cn.magnitudes
        self._c = {
            "cls": torch.from_numpy(cls.astype(np.uint8)).to(dev),
            "type": torch.from_numpy(self.cn.evse_type.astype(np.uint8)).to(dev),
            "re": torch.from_numpy(re).to(dev),
            "im": torch.from_numpy(im).to(dev),
            "mag": torch.from_numpy(mag).to(dev),
            "sessions": sessions_dev,
            "moer": torch.from_numpy(data_generator.moer).to(dev),
            "arrivals": arrivals_dev,
        }
        self._pool = pool
        self._max_profit_version = -1
        self._max_profit_host = None if pool is not None else t.max_profit(
            self.MARGINAL_PROFIT_PER_KWH, self.A_PERS_TO_KWH, self.ACTION_SCALE_FACTOR)
        self._model = _lib.EVModel(
            n, coef.shape[1], coef.shape[0], k, K_slots, n_trace_days, data_generator.moer.shape[0],
            int(project_action_in_env), self._c["cls"].data_ptr(), self._c["type"].data_ptr(),
            self._c["re"].data_ptr(), self._c["im"].data_ptr(), self._c["mag"].data_ptr(),
            self._c["sessions"].data_ptr(), self._c["moer"].data_ptr(), self._c["arrivals"].data_ptr(),
            self._c["arrivals"].shape[1] - 1, 0, None, None)
        self._c["tables"] = torch.zeros(_lib.SG_EV_TABLE_BYTES, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(self.lib.sg_ev_prepare_tables(C.byref(self._model), self._c["tables"].data_ptr(),
                                                     torch.cuda.current_stream(dev).cuda_stream), "sg_ev_prepare_tables")
        self._model.tables = self._c["tables"].data_ptr()
        # observation-tail table (forecasted_moer | prev_moer | timestep of every (MOER day, t)), see sg_ev_model.obs_tail
        tail_floats = (k + 3) & ~1
        self._c["obs_tail"] = torch.zeros(data_generator.moer.shape[0], _lib.SG_EV_MOER_ROWS, tail_floats,
                                          dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(self.lib.sg_ev_prepare_obs_tail(C.byref(self._model), self._c["obs_tail"].data_ptr(),
                                                       torch.cuda.current_stream(dev).cuda_stream), "sg_ev_prepare_obs_tail")
        self._model.obs_tail = self._c["obs_tail"].data_ptr()

        # ---- per-env state and outputs ---------------------------------------------------------------
        B = self.num_envs
        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=dev)  # noqa: E731
        # session records of the plugged-in EVs: [B, n] x 32 bytes, and one 64-byte header per env
        # {t, day, moer_day, arrival cursor | occupancy mask | pad || profit, carbon_cost, excess_charge | pad}
        # (see sg_ev_state); torch allocations are 512-byte aligned, which the vector accesses need
        self.slots = z(B, n, 4, dtype=torch.float64)
        self.header = z(B, 8, dtype=torch.float64)
        self.clock = self.header.view(torch.int32)[:, 0:4]        # columns: t, day, moer_day, arrival cursor
        self.t, self.day, self.moer_day = self.clock[:, 0], self.clock[:, 1], self.clock[:, 2]
        self.occ = self.header.view(torch.int64)[:, 2]            # bit i: station i has a plugged-in session
        self.breakdown = self.header[:, 4:7]
        self._state = _lib.EVState(self.slots.data_ptr(), self.header.data_ptr())
        self.obs = z(B, self.obs_dim, dtype=torch.float32)
        self.final_obs = z(B, self.obs_dim, dtype=torch.float32) if auto_reset else None
        self.reward = z(B, dtype=torch.float64)
        self.done = z(B, dtype=torch.uint8)
        self._d_trunc = z(B, dtype=torch.bool)               # truncated is always False (env.py:291)
        self.status = z(B, dtype=torch.uint8)
        self.pilots = z(B, n, dtype=torch.uint8)
        self.work = z(B + _lib.SG_EV_WORK_HEADER_INTS, dtype=torch.int32)   # tile scheduler + work list of the exact projection
        self.stats = z(_lib.SG_STATS_LEN, dtype=torch.float64) if collect_stats else None
        if self.stats is not None:
            self.reset_stats()
        self._action_dtype = torch.int64 if discrete_bins > 0 else torch.float32
        self._d_action = z(B, n, dtype=self._action_dtype)
        # host staging for the host-buffer path.  Pinned where the DMA engines touch it: the actions always, the
        # results only on the dense path — on the compact path they are written by the expansion threads, and
        # ordinary (first-touch, NUMA-local) memory is ~2.5x faster for them than cudaHostAlloc memory (measured)
        pin = (lambda t: t.pin_memory()) if (host_obs == "dense" or self.host_dense_envs > 0) else (lambda t: t)
        self._h_action = torch.zeros(B, n, dtype=self._action_dtype).pin_memory()
        self._h_obs = pin(torch.zeros(B, self.obs_dim, dtype=torch.float32))
        self._h_reward = pin(torch.zeros(B, dtype=torch.float64))
        self._h_done = pin(torch.zeros(B, dtype=torch.uint8))
        self._h_trunc = torch.zeros(B, dtype=torch.bool)                       # truncated is always False (env.py:291)
        self._flags = (_lib.SG_EV_ACTION_I64 if discrete_bins > 0 else _lib.SG_EV_ACTION_F32) | \
            (_lib.SG_EV_FLAG_AUTORESET if auto_reset else 0)
        self.kernel_launches = 0
        self._needs_reset = True
        # compact host path (allocated on first use): records, overflow area, host MOER table, activity masks
        self._compact = None
        self._host_obs_valid = False          # _h_obs holds the expansion of the previous host-path step
        self._pending = None                  # "compact" / "dense": what step_async enqueued
        self.host_timing = {"sync": 0.0, "expand": 0.0, "steps": 0}

    # ------------------------------------------------------------------------------------------
    def __repr__(self) -> str:
        return (f"EVChargingGym (action projection = {self.project_action_in_env}, "
                f"moer forecast steps = {self.moer_forecast_steps}) using {self.data_generator!r} "
                f"x {self.num_envs} envs on {self.device}")

    def _stream(self) -> int:
        cur = torch.cuda.current_stream(self.device)
        if self._async_stream is not None:      # work queued by step_async() touches the same state: order after it
            cur.wait_stream(self._async_stream)
        return cur.cuda_stream

    def reset_stats(self) -> None:
        _lib.check(self.lib.sg_stats_reset(self.stats.data_ptr(), self._stream()), "sg_stats_reset")

    def obs_dict(self, obs=None) -> dict:
        obs = self.obs if obs is None else obs
        return {k: obs[..., s] for k, s in self.obs_slices.items()}

    @nvtx_method("sg:ev.reset")
    def reset(self, *, seed: int | np.ndarray | None = None, options: dict | None = None,
              mask: torch.Tensor | np.ndarray | None = None):
        """env.py:293-338 for every env (or those in ``mask``).  ``seed`` may be one int (env i gets
        ``seed + i``) or an array of per-env seeds; it selects the day exactly as the reference's
        sequential ``RealTraceGenerator`` does."""
        if options is not None and "verbose" in options:
            self.verbose = options["verbose"]
        B = self.num_envs
        if seed is None:
            seeds = None
        elif np.isscalar(seed):
            seeds = int(seed) + np.arange(B)
        else:
            seeds = np.asarray(seed)
        self.data_generator.set_seed(None if seeds is None else int(seeds[0]))
        days = self.data_generator.first_days(seeds, B)
        mdays = self.data_generator.moer_days(days)
        d_days = torch.from_numpy(days).to(self.device)
        d_mdays = torch.from_numpy(mdays).to(self.device)
        d_mask = None
        if mask is not None:
            d_mask = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sg_ev_reset(C.byref(self._model), C.byref(self._state), B,
                                            _lib.ptr(d_mask), d_days.data_ptr(), d_mdays.data_ptr(),
                                            self.obs.data_ptr(), self._stream()), "sg_ev_reset")
        self.kernel_launches += 1
        self._needs_reset = False
        self._host_obs_valid = False
        return self.obs, self._infos()

    @nvtx_method("sg:ev.step")
    def step(self, actions):
        """env.py:229-291 for every env.  ``actions``: ``[num_envs, n]`` CUDA tensor (device path,
        returns CUDA tensors, asynchronous) or NumPy array (host path: pinned H2D copy, step, D2H
        copy of obs/reward/done, synchronised; returns NumPy arrays)."""
        if self._needs_reset:
            raise RuntimeError("call reset() before step()")      # env.py:283 NOTE
        B, n = self.num_envs, self.num_stations
        host_np = isinstance(actions, np.ndarray)
        if host_np or actions.device.type == "cpu":
            if tuple(actions.shape) != (B, n):
                raise ValueError(f"actions must have shape {(B, n)}, got {tuple(actions.shape)}")
            if host_np:
                self._h_action.numpy()[...] = actions
                h_act = self._h_action
            elif actions.is_pinned() and actions.dtype == self._action_dtype and actions.is_contiguous():
                h_act = actions                                   # zero-copy: DMA straight from the caller's buffer
            else:
                self._h_action.copy_(actions)
                h_act = self._h_action
            if self.host_obs == "compact":
                self._enqueue_host(h_act, self._stream())
                obs, rew, done, trunc, infos = self._finish_host(torch.cuda.current_stream(self.device))
            else:
                with torch.cuda.device(self.device):
                    _lib.check(self.lib.sg_ev_step_host(
                        C.byref(self._model), C.byref(self._state), B, h_act.data_ptr(), self._flags,
                        max(self.discrete_bins, 0), self._h_obs.data_ptr(), self._h_reward.data_ptr(),
                        self._h_done.data_ptr(), self._d_action.data_ptr(), self.obs.data_ptr(),
                        self.reward.data_ptr(), self.done.data_ptr(), self.status.data_ptr(),
                        _lib.ptr(self.stats), self.work.data_ptr(), self._stream()), "sg_ev_step_host")
                self.kernel_launches += 1
            if host_np:
                done = self._h_done.numpy().astype(bool)
                return (self._h_obs.numpy().copy(), self._h_reward.numpy().copy(), done,
                        np.zeros(B, bool), self._infos())
            # torch host path: views of the pinned result buffers (valid until the next step)
            return self._h_obs, self._h_reward, self._h_done.view(torch.bool), self._h_trunc, self._infos()
        if actions.device != self.device or actions.dtype != self._action_dtype or tuple(actions.shape) != (B, n):
            raise ValueError(f"actions must be a {self._action_dtype} tensor of shape {(B, n)} on {self.device}")
        actions = actions.contiguous()
        self.launch_step(actions)
        return self.obs, self.reward, self.done.view(torch.bool), self._d_trunc, self._infos()

    def step_async(self, actions: torch.Tensor) -> None:
        """SB3 ``VecEnv.step_async``: enqueue H2D(actions) -> step -> D2H(obs, reward, done) on this env's
        own stream and return immediately.  ``actions`` must be a pinned, contiguous host tensor of the
        action dtype and must stay untouched until ``step_wait``.  Several batched envs stepped this way
        overlap their copies in both PCIe directions."""
        if self._needs_reset:
            raise RuntimeError("call reset() before step()")
        if not (actions.device.type == "cpu" and actions.is_pinned() and actions.is_contiguous()
                and actions.dtype == self._action_dtype and tuple(actions.shape) == (self.num_envs, self.num_stations)):
            raise ValueError("step_async needs a pinned contiguous host tensor of shape (num_envs, n)")
        if self._async_stream is None:
            self._async_stream = torch.cuda.Stream(self.device)
        # order this step after everything queued on the caller's stream (reset(), step(), graph replays ...)
        self._async_stream.wait_stream(torch.cuda.current_stream(self.device))
        self._enqueue_host(actions, self._async_stream.cuda_stream)

    def step_wait(self):
        """SB3 ``VecEnv.step_wait``: wait for the step enqueued by ``step_async``; returns views of the
        pinned result buffers (obs f32 [B, obs_dim], reward f64 [B], done bool [B], truncated, infos)."""
        return self._finish_host(self._async_stream)

    # ---- host-buffer path ---------------------------------------------------------------------------------------
    def _ensure_compact(self) -> dict:
        if self._compact is None:
            B, n, dev = self.num_envs, self.num_stations, self.device
            nbytes = _lib.SG_EV_COMPACT_HEADER_BYTES + B * _lib.SG_EV_COMPACT_RECORD_BYTES
            cap = max(256, B // 64)                     # envs with > 16 active sessions per step before the dense fallback
            self._compact = {
                "cap": cap,
                "d": torch.zeros(nbytes, dtype=torch.uint8, device=dev),
                "h": torch.zeros(nbytes, dtype=torch.uint8).pin_memory(),
                "d_idx": torch.zeros(cap, dtype=torch.int32, device=dev),
                "d_rows": torch.zeros(cap, 2 * n, dtype=torch.float32, device=dev),
                "h_idx": torch.zeros(cap, dtype=torch.int32).pin_memory(),
                "h_rows": torch.zeros(cap, 2 * n, dtype=torch.float32).pin_memory(),
                "prev": torch.zeros(B, dtype=torch.int64),
                "moer": np.ascontiguousarray(self.data_generator.moer, dtype=np.float32),
                "h_status": torch.zeros(B, dtype=torch.uint8).pin_memory(),
            }
            c = self._compact
            c["x"] = _lib.EVHostExpand(
                c["moer"].ctypes.data, c["moer"].shape[0], 0, c["prev"].data_ptr(), c["h_idx"].data_ptr(),
                c["h_rows"].data_ptr(), self._h_obs.data_ptr(), self._h_reward.data_ptr(), self._h_done.data_ptr(),
                c["h_status"].data_ptr(), self.host_threads, 0)
        return self._compact

    def _enqueue_host(self, h_act: torch.Tensor, stream: int) -> None:
        """H2D(actions) -> step -> [pack ->] D2H on ``stream``, no synchronisation."""
        args = (C.byref(self._model), C.byref(self._state), self.num_envs, h_act.data_ptr(), self._flags,
                max(self.discrete_bins, 0))
        dev_out = (self._d_action.data_ptr(), self.obs.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(),
                   self.status.data_ptr(), _lib.ptr(self.stats), self.work.data_ptr())
        with torch.cuda.device(self.device):
            if self.host_obs == "compact":
                c = self._ensure_compact()
                x = c["x"]
                x.incremental = int(self._host_obs_valid and self.host_incremental)
                _lib.check(self.lib.sg_ev_step_host_compact_stream(
                    *args, c["h"].data_ptr(), *dev_out, c["d"].data_ptr(), c["d_idx"].data_ptr(),
                    c["d_rows"].data_ptr(), c["cap"], _lib.ptr(self.final_obs), C.byref(x), self.host_dense_envs, stream),
                    "sg_ev_step_host_compact_stream")
                self._host_obs_valid = True       # (the expansion is stream-ordered: the next one may build on it)
                self.kernel_launches += 1
            else:
                _lib.check(self.lib.sg_ev_step_host_async(
                    *args, self._h_obs.data_ptr(), self._h_reward.data_ptr(), self._h_done.data_ptr(), *dev_out,
                    stream), "sg_ev_step_host_async")
        self._pending = self.host_obs
        self.kernel_launches += 1

    def _finish_host(self, stream: "torch.cuda.Stream"):
        import time
        t0 = time.perf_counter()
        stream.synchronize()
        t1 = time.perf_counter()
        B = self.num_envs
        if self._pending == "compact":
            # the expansion ran on the stream (host function); only a step whose overflow area was too small for the
            # envs with more than 16 active sessions needs the dense observation
            c = self._compact
            if int(c["h"][:4].view(torch.int32)[0]) > c["cap"]:
                self._h_obs.copy_(self.obs)
                torch.cuda.current_stream(self.device).synchronize()
        self._pending = None
        ht = self.host_timing                 # where the host-path time goes (seconds, cumulative)
        ht["sync"] += t1 - t0; ht["expand"] += time.perf_counter() - t1; ht["steps"] += 1
        # zero-copy results: no torch CPU op (and none of its OpenMP threads) on the per-step path
        return (self._h_obs, self._h_reward, self._h_done.view(torch.bool), self._h_trunc, self._infos())

    @property
    def _max_profit_per_day(self) -> np.ndarray:
        """``_calculate_max_profit`` (env.py:422-429) of every day of the pool: [D] f64 on the host (cached; a
        device-sampled pool is read back once per ``resample()``)."""
        if self._pool is None:
            return self._max_profit_host
        version = getattr(self.data_generator, "pool_version", 0)
        if self._max_profit_host is None or self._max_profit_version != version:
            self._max_profit_host = self._pool["max_kwh"].cpu().numpy() * self.MARGINAL_PROFIT_PER_KWH
            self._max_profit_version = version
        return self._max_profit_host

    def episode_info(self) -> dict:
        """Per-env ``info`` of the reference at the end of an episode (env.py:403-406), as columns:
        ``max_profit`` [B] and ``reward_breakdown`` (list of dicts) — what ``BaseAlgorithm.run`` tabulates."""
        bd = self.breakdown.cpu().numpy()
        day = self.day.cpu().numpy()
        mp = self._max_profit_per_day
        return {"max_profit": [float(mp[d]) for d in day],
                "reward_breakdown": [{"profit": float(r[0]), "carbon_cost": float(r[1]), "excess_charge": float(r[2])}
                                     for r in bd]}

    def launch_step(self, actions: torch.Tensor) -> None:
        """Enqueues the step kernel on the current stream without building return values (used by
        ``step``, CUDA-graph capture and the benchmark)."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sg_ev_step(
                C.byref(self._model), C.byref(self._state), self.num_envs, actions.data_ptr(), self._flags,
                max(self.discrete_bins, 0), self.obs.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(),
                self.pilots.data_ptr(), _lib.ptr(self.final_obs), self.status.data_ptr(), _lib.ptr(self.stats),
                self.work.data_ptr(), self._stream()), "sg_ev_step")
        self.kernel_launches += 1
        self._host_obs_valid = False          # the next host-path step rewrites every observation row

    def export_state(self):
        """Dense views of the sparse per-station state: (delivered kWh f64 [B,n], trace slot u8 [B,n],
        255 where the station is empty)."""
        B, n = self.num_envs, self.num_stations
        delivered = torch.empty(B, n, dtype=torch.float64, device=self.device)
        sess_idx = torch.empty(B, n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sg_ev_export_state(C.byref(self._model), C.byref(self._state), B,
                                                   delivered.data_ptr(), sess_idx.data_ptr(), self._stream()),
                       "sg_ev_export_state")
        return delivered, sess_idx

    def capture_graph(self, actions: torch.Tensor) -> "torch.cuda.CUDAGraph":
        """Captures one step (reading ``actions`` in place) into a CUDA graph; ``graph.replay()``
        then advances all envs with a single graph launch."""
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                self.launch_step(actions)
        torch.cuda.current_stream(self.device).wait_stream(s)
        return g

    def project(self, actions: torch.Tensor, demands: torch.Tensor):
        """The action projection alone (env.py:200-221): returns (x f64 [B,n], status u8 [B], iters i32 [B])."""
        B = actions.shape[0]
        x = torch.empty(B, self.num_stations, dtype=torch.float64, device=self.device)
        st = torch.empty(B, dtype=torch.uint8, device=self.device)
        it = torch.empty(B, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sg_ev_project(C.byref(self._model), B, actions.contiguous().data_ptr(),
                                              demands.contiguous().data_ptr(), x.data_ptr(), st.data_ptr(),
                                              it.data_ptr(), self._stream()), "sg_ev_project")
        return x, st, it

    def _infos(self) -> dict:
        """Lazy batched info (env.py:403-406): tensors, not per-env dicts — no sync on the hot path."""
        return {"reward_breakdown": self.breakdown, "day": self.day, "status": self.status,
                "max_profit_per_day": self._max_profit_per_day}

    def close(self) -> None:
        pass


class EVChargingEnv:
    """The reference's single-env Gymnasium interface (env.py:20-470) on top of a 1-env batch.

    ``reset(*, seed, options) -> (obs dict, info)``; ``step(action) -> (obs dict, float reward,
    bool terminated, False, info)``.  NumPy in, NumPy out.  Unlike the reference, the caller's
    action array is not scaled in place (env.py:366 mutates it when projection is off)."""

    metadata: dict = {}
    render_mode = None
    spec = None

    def __init__(self, data_generator: AbstractTraceGenerator, moer_forecast_steps: int = 36,
                 project_action_in_env: bool = True, verbose: int = 0, device=None, discrete_bins: int = -1):
        self._b = BatchedEVChargingEnv(data_generator, 1, moer_forecast_steps, project_action_in_env, verbose,
                                       device, auto_reset=False, discrete_bins=discrete_bins)
        self.data_generator = data_generator
        self.max_timestep = 288
        self.moer_forecast_steps = moer_forecast_steps
        self.project_action_in_env = project_action_in_env
        self.verbose = verbose
        self.cn = self._b.cn
        self.num_stations = self._b.num_stations
        self.observation_space = self._b.single_observation_space
        self.action_space = self._b.single_action_space
        self.t = 0
        self._episodes = 0
        self._seed0 = 0

    def __repr__(self) -> str:
        return (f"EVChargingGym (action projection = {self.project_action_in_env}, "
                f"moer forecast steps = {self.moer_forecast_steps}) using {self.data_generator!r}")

    def _obs(self, flat: np.ndarray) -> dict:
        return {k: flat[0, s].copy() for k, s in self._b.obs_slices.items()}

    def _info(self) -> dict[str, Any]:
        bd = self._b.breakdown[0].cpu().numpy()
        day = int(self._b.day[0])
        return {"max_profit": float(self._b._max_profit_per_day[day]),
                "reward_breakdown": {"profit": float(bd[0]), "carbon_cost": float(bd[1]),
                                     "excess_charge": float(bd[2])}}

    @property
    def moer(self) -> np.ndarray:
        """The current day's MOER table ``[289, 37]`` (column 0: historical value, 1..36: forecasts) — the array the
        reference keeps in ``env.moer`` after ``reset`` (env.py:317-318) and its offline-optimal controller reads."""
        return np.asarray(self.data_generator.moer[int(self._b.moer_day[0])], dtype=np.float64)

    @property
    def _evs(self) -> list:
        """The current day's charging sessions as records with ``station_id``, ``arrival``, ``departure``,
        ``estimated_departure`` and ``requested_energy`` (kWh) — the attributes of the acnportal ``EV`` objects the
        reference holds in ``env._evs`` (env.py:313-316) that its controllers use."""
        from types import SimpleNamespace
        ids = self.cn.station_ids
        return [SimpleNamespace(station_id=ids[st], arrival=int(arr), departure=int(dep), estimated_departure=int(est),
                                requested_energy=float(kwh))
                for arr, dep, est, kwh, st in self.data_generator.days[int(self._b.day[0])]]

    def reset(self, *, seed: int | None = None, options: dict | None = None):
        if seed is None:
            # the reference's sequential generator simply moves on to the next day (event_generation.py:284-291)
            seed_eff = self._seed0 + self._episodes
        else:
            self._seed0, self._episodes, seed_eff = int(seed), 0, int(seed)
        self._episodes += 1
        self.action_space.seed(seed)
        obs, _ = self._b.reset(seed=seed_eff, options=options)
        self.t = 0
        return self._obs(obs.cpu().numpy()), self._info()

    def step(self, action):
        a = np.asarray(action).reshape(1, self.num_stations)
        a = a.astype(np.int64 if self._b.discrete_bins > 0 else np.float32)
        obs, rew, done, trunc, _ = self._b.step(a)
        self.t += 1
        return self._obs(obs), float(rew[0]), bool(done[0]), False, self._info()

    def close(self) -> None:
        self._b.close()
