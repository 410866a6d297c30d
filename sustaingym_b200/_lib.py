"""ctypes binding of ``libsustaingym_b200.so`` (the C ABI declared in ``include/sustaingym_b200.h``).

There is deliberately NO fallback: if the shared library is missing or a call fails, the caller
gets an exception.  The product path never routes through ``oracle/`` or any CPU code.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SG_LIB_PATH", os.path.join(_PKG_DIR, "libsustaingym_b200.so"))

ABI_VERSION = 2
SG_STATS_LEN = 16
SG_EV_MOER_COLS = 37
SG_EV_MOER_ROWS = 289
SG_EV_EPISODE_STEPS = 288
SG_EV_TABLE_BYTES = 4096


def cogen_scratch_floats(B: int) -> int:
    """SG_COGEN_SCRATCH_FLOATS(B) of include/sustaingym_b200.h."""
    return int(B) * (23 + 23 * 16) + 4

SG_EV_MAX_STATIONS = 64
SG_EV_MAX_CLASSES = 8
SG_EV_MAX_CONSTRAINTS = 16
SG_EV_STATUS_PROJECTED = 1
SG_EV_STATUS_MAXITER = 2
SG_EV_ACTION_F32 = 0
SG_EV_ACTION_I64 = 1
SG_EV_FLAG_AUTORESET = 16
SG_EV_FLAG_STATIC_TILES = 32
SG_EV_WORK_HEADER_INTS = 96
SG_EV_COMPACT_ENTRIES = 16
SG_EV_COMPACT_HEADER_BYTES = 16
SG_EV_COMPACT_RECORD_BYTES = 120
SG_BUILDING_FLAG_AUTORESET = 16
SG_BUILDING_FLAG_NO_MMA = 32

_vp = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f64 = C.c_double


class BuildingModel(C.Structure):
    _fields_ = [
        ("n", _i32), ("T", _i32), ("episode_len", _i32), ("discrete", _i32),
        ("q_rate", _f64), ("error_rate", _f64), ("pnorm", _f64),
        ("A_d", _vp), ("BD_d", _vp), ("target", _vp), ("ac_map", _vp), ("weather", _vp),
    ]


class EVModel(C.Structure):
    _fields_ = [
        ("n_stations", _i32), ("n_classes", _i32), ("n_constraints", _i32), ("moer_steps", _i32),
        ("max_sessions", _i32), ("n_days", _i32), ("n_moer_days", _i32), ("project", _i32),
        ("station_class", _vp), ("station_type", _vp), ("con_re", _vp), ("con_im", _vp),
        ("magnitudes", _vp), ("sessions", _vp), ("moer", _vp), ("arrivals", _vp),
        ("max_day_sessions", _i32), ("reserved", _i32), ("tables", _vp), ("obs_tail", _vp),
    ]


class EVGmm(C.Structure):
    _fields_ = [
        ("n_components", _i32), ("n_counts", _i32), ("n_stations", _i32), ("reserved", _i32),
        ("cum_weights", _vp), ("means", _vp), ("chol", _vp), ("counts", _vp), ("station_p", _vp),
    ]


class EVHostExpand(C.Structure):
    _fields_ = [
        ("h_moer", _vp), ("n_moer_days", _i32), ("incremental", _i32), ("h_prev_active", _vp), ("h_ovf_idx", _vp),
        ("h_ovf_rows", _vp), ("h_obs", _vp), ("h_reward", _vp), ("h_done", _vp), ("h_status", _vp),
        ("n_threads", _i32), ("reserved", _i32),
    ]


class EVState(C.Structure):
    _fields_ = [
        ("slots", _vp), ("header", _vp),
    ]


class CogenModel(C.Structure):
    _fields_ = [
        ("n_sub", _i32), ("forecast_horizon", _i32), ("n_days", _i32), ("reserved", _i32),
        ("ramp_penalty", _f64), ("supply_imbalance_penalty", _f64), ("constraint_violation_penalty", _f64),
        ("blob", _vp), ("offsets", _vp), ("meta", _vp), ("feats", _vp), ("mu", _vp), ("sinv", _vp), ("b3", _vp),
        ("ambient", _vp), ("tc_blob", _vp), ("tc_offsets", _vp),
    ]


class MarketModel(C.Structure):
    _fields_ = [
        ("n_gen", _i32), ("k", _i32), ("n_days", _i32), ("use_intermediate_rewards", _i32),
        ("e_max", _f64), ("e_init", _f64), ("p_max", _f64), ("eta_c", _f64), ("eta_d", _f64),
        ("price_cap", _f64), ("p_co2", _f64), ("terminal_price", _f64),
        ("gen_cost", _vp), ("gen_cap", _vp), ("load", _vp), ("moer", _vp),
    ]


# name -> (restype, argtypes); every symbol include/sustaingym_b200.h declares
SIGNATURES = {
    "sg_abi_version": (C.c_int, []),
    "sg_last_error": (C.c_char_p, []),
    "sg_device_sm_count": (C.c_int, []),
    "sg_stats_reset": (C.c_int, [_vp, _vp]),
    "sg_building_reset": (C.c_int, [C.POINTER(BuildingModel), _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sg_building_step": (C.c_int, [C.POINTER(BuildingModel), _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "sg_building_rollout": (C.c_int, [C.POINTER(BuildingModel), _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                      _i32, _vp, _vp]),
    "sg_building_step_host": (C.c_int, [C.POINTER(BuildingModel), _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _i32, _vp, _vp]),
    "sg_ev_prepare_tables": (C.c_int, [C.POINTER(EVModel), _vp, _vp]),
    "sg_ev_prepare_obs_tail": (C.c_int, [C.POINTER(EVModel), _vp, _vp]),
    "sg_ev_reset": (C.c_int, [C.POINTER(EVModel), C.POINTER(EVState), _i64, _vp, _vp, _vp, _vp, _vp]),
    "sg_ev_step": (C.c_int, [C.POINTER(EVModel), C.POINTER(EVState), _i64, _vp, _i32, _i32, _vp, _vp, _vp, _vp,
                             _vp, _vp, _vp, _vp, _vp]),
    "sg_ev_step_host": (C.c_int, [C.POINTER(EVModel), C.POINTER(EVState), _i64, _vp, _i32, _i32, _vp, _vp, _vp,
                                  _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sg_ev_step_host_async": (C.c_int, [C.POINTER(EVModel), C.POINTER(EVState), _i64, _vp, _i32, _i32, _vp, _vp, _vp,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sg_ev_pack_obs": (C.c_int, [_i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "sg_ev_step_host_compact_async": (C.c_int, [C.POINTER(EVModel), C.POINTER(EVState), _i64, _vp, _i32, _i32, _vp, _vp,
                                                _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "sg_ev_step_host_compact_stream": (C.c_int, [C.POINTER(EVModel), C.POINTER(EVState), _i64, _vp, _i32, _i32, _vp, _vp,
                                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp,
                                                 C.POINTER(EVHostExpand), _i64, _vp]),
    "sg_ev_expand_obs": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32]),
    "sg_ev_export_state": (C.c_int, [C.POINTER(EVModel), C.POINTER(EVState), _i64, _vp, _vp, _vp]),
    "sg_ev_agent_obs": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "sg_ev_sample_gmm_days": (C.c_int, [C.POINTER(EVGmm), _i64, _i64, C.c_uint64, _f64, _f64, _i32, _i32, _vp, _vp,
                                        _vp, _vp, _vp]),
    "sg_cogen_plant": (C.c_int, [C.POINTER(CogenModel), _i64, _vp, _vp, _vp, _vp]),
    "sg_cogen_prepare_tc": (C.c_int, [C.POINTER(CogenModel), _vp, _vp]),
    "sg_cogen_reset": (C.c_int, [C.POINTER(CogenModel), _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sg_cogen_step": (C.c_int, [C.POINTER(CogenModel), _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                _i32, _vp, _vp]),
    "sg_tc_gemm_selftest": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp]),
    "sg_market_reset": (C.c_int, [C.POINTER(MarketModel), _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sg_market_step": (C.c_int, [C.POINTER(MarketModel), _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sg_ev_project": (C.c_int, [C.POINTER(EVModel), _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Loads the shared library (once). Raises ``NativeLibraryError`` if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m sustaingym_b200._build` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    if lib.sg_abi_version() != ABI_VERSION:
        raise NativeLibraryError("ABI version mismatch between _lib.py and the shared library")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().sg_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def ptr(t) -> int | None:
    """Device/host pointer of a torch tensor (``None`` -> NULL)."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError("tensor passed to the native library must be contiguous")
    return t.data_ptr()


def current_stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream
