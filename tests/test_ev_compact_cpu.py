"""CPU: sg_ev_expand_obs (host function of the C-ABI, no CUDA call) against a NumPy restatement of the record
format in include/sustaingym_b200.h — full and incremental expansion, overflow rows, argument checks."""
import ctypes as C

import numpy as np
import pytest

REC = np.dtype([("active", "<u8"), ("reward", "<f8"), ("t", "<i2"), ("moer_day", "<i2"), ("done", "u1"), ("flags", "u1"),
                ("n_active", "u1"), ("status", "u1"), ("demand", "<f4", 16), ("est", "<i2", 16)])


def make_records(rng, B, n, Dm, p_busy):
    rec = np.zeros(B, REC)
    dense = np.zeros((B, 2 * n), np.float32)
    ovf_idx, ovf_rows = [], []
    for b in range(B):
        occ = rng.uniform(size=n) < rng.choice([0.05, 0.15, p_busy])
        idx = np.nonzero(occ)[0]
        dense[b, idx] = rng.uniform(0.01, 90, len(idx)).astype(np.float32)
        dense[b, n + idx] = rng.integers(-20, 250, len(idx))
        rec["active"][b] = sum(1 << int(i) for i in idx)
        rec["n_active"][b] = len(idx)
        k = min(len(idx), 16)
        rec["demand"][b, :k] = dense[b, idx[:k]]
        rec["est"][b, :k] = dense[b, n + idx[:k]]
        if len(idx) > 16:
            rec["flags"][b] = 1
            ovf_idx.append(b); ovf_rows.append(dense[b].copy())
    rec["reward"] = rng.normal(size=B)
    rec["t"] = rng.integers(0, 289, B)
    rec["moer_day"] = rng.integers(0, Dm, B)
    rec["done"] = rec["t"] == 288
    rec["status"] = rng.integers(0, 2, B)
    return rec, dense, np.array(ovf_idx, np.int32), np.array(ovf_rows, np.float32).reshape(-1, 2 * n)


def expected_obs(rec, dense, moer, n, k):
    B = len(rec)
    out = np.zeros((B, 2 * n + k + 2), np.float32)
    out[:, :2 * n] = dense
    rows = moer[rec["moer_day"], rec["t"]]
    out[:, 2 * n:2 * n + k] = rows[:, 1:k + 1]
    out[:, 2 * n + k] = rows[:, 0]
    out[:, 2 * n + k + 1] = (rec["t"].astype(np.float64) * (1.0 / 288)).astype(np.float32)
    return out


@pytest.mark.parametrize("n,k,threads", [(54, 36, 1), (54, 36, 5), (52, 7, 3), (64, 1, 2)])
def test_expand_obs_matches_numpy(native_lib, n, k, threads):
    assert REC.itemsize == 120
    rng = np.random.default_rng(n + k)
    B, Dm = 1000, 5
    moer = rng.uniform(0, 1, (Dm, 289, 37)).astype(np.float32)
    obs = np.zeros((B, 2 * n + k + 2), np.float32)
    rew, done, status = np.zeros(B), np.zeros(B, np.uint8), np.zeros(B, np.uint8)
    prev = np.zeros(B, np.uint64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    for step in range(6):
        rec, dense, oi, orows = make_records(rng, B, n, Dm, 0.5)
        buf = np.zeros(16 + B * 120, np.uint8)
        buf[:4] = np.array([len(oi)], np.int32).view(np.uint8)
        buf[16:] = rec.view(np.uint8)
        incremental = int(step not in (0, 3))          # steps 0 and 3 rewrite every row
        if step == 3:
            obs[:] = 7.0                                # garbage that a full expansion must overwrite
        rc = native_lib.sg_ev_expand_obs(p(buf), B, n, k, p(moer), Dm, p(prev), incremental, p(oi), p(orows), len(oi),
                                         p(obs), p(rew), p(done), p(status), threads)
        assert rc == 0, native_lib.sg_last_error()
        np.testing.assert_array_equal(obs, expected_obs(rec, dense, moer, n, k))
        np.testing.assert_array_equal(rew, rec["reward"])
        np.testing.assert_array_equal(done, rec["done"])
        np.testing.assert_array_equal(status, rec["status"])
        np.testing.assert_array_equal(prev, rec["active"])
        assert len(oi) > 0


def test_expand_obs_rejects_bad_records(native_lib):
    B, n, k = 4, 54, 36
    moer = np.zeros((2, 289, 37), np.float32)
    rec = np.zeros(B, REC)
    rec["moer_day"][2] = 9                               # outside the MOER table
    buf = np.zeros(16 + B * 120, np.uint8); buf[16:] = rec.view(np.uint8)
    obs, rew, done = np.zeros((B, 146), np.float32), np.zeros(B), np.zeros(B, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    assert native_lib.sg_ev_expand_obs(p(buf), B, n, k, p(moer), 2, None, 0, None, None, 0, p(obs), p(rew), p(done), None, 2) != 0
    assert b"out-of-range" in native_lib.sg_last_error()
    assert native_lib.sg_ev_expand_obs(p(buf), B, n, k, p(moer), 2, None, 1, None, None, 0, p(obs), p(rew), p(done), None, 2) != 0
