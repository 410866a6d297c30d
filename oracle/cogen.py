"""TEST INFRASTRUCTURE — CPU oracle for the CogenEnv hot path (NumPy, vectorised over envs).

Restates (reference = chrisyeh96/sustaingym @ eac2a4d):

* ``plant_model``       the ONNX graph ``sustaingym/data/cogen/onnx_model/model.onnx`` (SURVEY.md
                        Appendix B) evaluated from the packed weights ``sustaingym_b200/data/cogen_plant.npz``
                        (produced from the reference file by ``tests/golden/gen_cogen_pack.py``)
* ``dyn_constraint_violation``  ``CogenEnv._dyn_constraint_volation``  cogen/env.py:232-274
* ``compute_reward``    ``CogenEnv._compute_reward``                  cogen/env.py:276-353
* ``forecast`` / obs    ``_forecast_from_time`` / ``_get_obs``        cogen/env.py:145-184
* ``step``              ``CogenEnv.step``                             cogen/env.py:355-387
* ``agent_rewards``     ``MultiAgentCogenEnv.step``                   cogen/multiagent_env.py:97-101
* ``flatten_obs``       ``gymnasium.spaces.flatten`` of the Dict obs (sorted keys, Discrete -> one-hot)

Pins: ``plant_model`` is checked against ``tests/golden/cogen_model_io.npz`` — outputs of the
reference ONNX graph evaluated node by node (``oracle/onnx_reader.py``) on 4097 inputs including the
SURVEY.md D.2 vector.  The reference runs the graph with ONNX Runtime (not installable here), whose
sgemm sums in another order: PARITY with ORT itself is UNPINNED (expected ~1e-6 relative); the
reward/obs glue is pinned only by transcription (the reference's tests hold no numeric vectors).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
from __future__ import annotations

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLANT_NPZ = os.path.join(ROOT, "sustaingym_b200", "data", "cogen_plant.npz")

# model input order (model.json) minus the three ambient inputs = the 15 action entries
ACTION_KEYS = ["GT1_PAC_FFU", "GT1_EVC_FFU", "GT1_PWR", "GT2_PAC_FFU", "GT2_EVC_FFU", "GT2_PWR",
               "GT3_PAC_FFU", "GT3_EVC_FFU", "GT3_PWR", "HR1_HPIP_M_PROC", "HR2_HPIP_M_PROC",
               "HR3_HPIP_M_PROC", "ST_PWR", "IPPROC_M", "CT_NrBays"]
AMBIENT_COLS = ["TAMB", "PAMB", "RHAMB", "Target_Power", "Target_Steam", "Energy_Price", "Gas_Price"]
STEPS_PER_DAY = 96


class PlantModel:
    def __init__(self, path: str = PLANT_NPZ):
        z = np.load(path)
        self.z = z
        self.meta, self.feats, self.mu, self.sinv = z["meta"], z["feats"], z["mu"], z["sinv"]
        self.b3, self.offs, self.blob = z["b3"], z["offsets"], z["blob"]

    def _arr(self, j, c, shape):
        o = int(self.offs[j, c])
        return self.blob[o:o + int(np.prod(shape))].reshape(shape)

    def __call__(self, X: np.ndarray, dtype=np.float32) -> np.ndarray:
        """X [B,18] -> Y [B,29] (Appendix B), evaluated in ``dtype``."""
        X = np.asarray(X, dtype)
        B = len(X)
        g = lambda v: np.abs(np.sign(v))                                  # noqa: E731
        gGT = [g(X[:, 5]), g(X[:, 8]), g(X[:, 11])]
        gHR = [g(X[:, 12]), g(X[:, 13]), g(X[:, 13])]                     # sic: HR3's gate reads HR2 (node 46)
        gGH = [g(gGT[i] + gHR[i]) for i in range(3)]
        gST = g(X[:, 15])
        E = np.concatenate([X, np.stack([X[:, 3] * gGT[0], X[:, 6] * gGT[1], X[:, 9] * gGT[2],
                                         X[:, 4] * gGT[0], X[:, 7] * gGT[1], X[:, 10] * gGT[2],
                                         X[:, 16] * gST], axis=1)], axis=1).astype(dtype)
        gates = gGT + gGH
        Y = np.zeros((B, 29), dtype)
        with np.errstate(over="ignore"):
            for j in range(23):
                d, H1, H2, act, yi, gate = (int(v) for v in self.meta[j])
                x = (E[:, self.feats[j, :d]] - self.mu[j, :d].astype(dtype)) * self.sinv[j, :d].astype(dtype)
                f = (lambda v: 1 / (1 + np.exp(-v))) if act == 0 else (lambda v: np.maximum(v, 0))
                h = f(x @ self._arr(j, 0, (d, H1)).astype(dtype) + self._arr(j, 1, (H1,)).astype(dtype)).astype(dtype)
                h = f(h @ self._arr(j, 2, (H1, H2)).astype(dtype) + self._arr(j, 3, (H2,)).astype(dtype)).astype(dtype)
                y = h @ self._arr(j, 4, (H2,)).astype(dtype) + dtype(self.b3[j])
                Y[:, yi] = y * gates[gate] if gate >= 0 else y
        Y[:, 6:9] = Y[:, 0:3] + Y[:, 3:6]                                 # GTi_HRi_NG_M
        Y[:, 21] = (Y[:, 6] + Y[:, 7]) + Y[:, 8]                          # PLANT_NG_M
        Y[:, 27] = (((X[:, 5] + X[:, 8]) + X[:, 11]) + X[:, 15]) - Y[:, 26]   # PLANT_NetPWR
        Y[:, 28] = ((X[:, 12] + X[:, 13]) + X[:, 14]) + X[:, 16]          # PLANT_PROCSTEAM
        return Y


def dyn_constraint_violation(X: np.ndarray, Y: np.ndarray) -> np.ndarray:
    """cogen/env.py:232-274 (index quirks kept: cv[14] uses Y[22], cv[15] uses Y[23], both 'max')."""
    f = np.float32
    X, Y = X.astype(f), Y.astype(f)
    cv = np.zeros((len(X), 16))
    pairs = [(9, 5, 10), (15, 12, 16), (11, 8, 12), (17, 13, 18), (13, 11, 14), (19, 14, 20), (24, 15, 25)]
    for i, (lo, xin, hi) in enumerate(pairs):
        cv[:, 2 * i] = np.maximum(f(0), Y[:, lo] - X[:, xin])
        cv[:, 2 * i + 1] = np.maximum(f(0), X[:, xin] - Y[:, hi])
    cv[:, 14] = np.maximum(f(0), X[:, 16] - Y[:, 22])
    cv[:, 15] = np.maximum(f(0), X[:, 16] - Y[:, 23])
    return cv


class CogenOracle:
    """B independent CogenEnv instances in lock-step.

    ``ambient``: [D, 96, 7] f64/f32 table (columns AMBIENT_COLS) — the content of the reference's
    ``ambients_dfs`` (load_ambients.construct_df).  Actions/prev-actions are [B, 15] f32 rows in
    ACTION_KEYS order."""

    def __init__(self, ambient: np.ndarray, num_envs: int, ramp_penalty: float = 2.0,
                 supply_imbalance_penalty: float = 1000, constraint_violation_penalty: float = 1000,
                 forecast_horizon: int = 3, plant: PlantModel | None = None):
        self.amb = np.asarray(ambient, np.float64)
        self.B, self.k = num_envs, forecast_horizon
        self.rp, self.ip, self.cp = ramp_penalty, supply_imbalance_penalty, constraint_violation_penalty
        self.plant = plant or PlantModel()
        self.n_days = self.amb.shape[0]
        assert 0 <= forecast_horizon < STEPS_PER_DAY - 1

    def forecast(self, day, t, noise=None):
        """cogen/env.py:145-162: rows [t, t+k] of the day, spilling into day+1; noise on rows >= 1."""
        k = self.k
        tt = t[:, None] + np.arange(k + 1)[None, :]
        dd = day[:, None] + (tt >= STEPS_PER_DAY)
        tt = tt % STEPS_PER_DAY
        if (dd >= self.n_days).any():
            raise IndexError("forecast window runs past the last day (cogen/env.py:152-155)")
        f = self.amb[dd, tt]                                     # [B, k+1, 7] f64
        if noise is not None:
            f = f.copy()
            f[:, 1:, :] += noise
        return f.astype(np.float32)

    def reset(self, day: np.ndarray, prev_action: np.ndarray):
        self.day = np.asarray(day, np.int64).copy()
        self.t = np.zeros(self.B, np.int64)
        self.prev = np.asarray(prev_action, np.float32).copy()
        self.fc = self.forecast(self.day, self.t)
        return self.obs()

    def obs(self):
        o = {c: self.fc[:, :, i] for i, c in enumerate(AMBIENT_COLS)}
        o["Time"] = (self.t / STEPS_PER_DAY).astype(np.float32)[:, None]
        o["Prev_Action"] = self.prev.copy()
        return o

    def step(self, action: np.ndarray, noise: np.ndarray | None = None):
        f = np.float32
        a = np.asarray(action, f)
        X = np.concatenate([self.fc[:, 0, :3], a], axis=1).astype(f)          # env.py:294-302
        Y = self.plant(X, np.float32)
        # fuel: env.py:310-316 ; ramp: :319-326 (f32 |diff| times Python float stays f32)
        fuel = np.stack([Y[:, 6], Y[:, 7], Y[:, 8], np.zeros(self.B, f)], 1)
        pw = [2, 5, 8, 12]
        ramp = (f(self.rp) * np.abs(a[:, pw] - self.prev[:, pw])).astype(f)
        total_ramp = ((ramp[:, 0] + ramp[:, 1]) + ramp[:, 2]) + ramp[:, 3]      # sum(dict.values()) of f32
        cv = dyn_constraint_violation(X, Y)
        dyn = self.cp * np.stack([cv[:, 0:4].sum(1), cv[:, 4:8].sum(1), cv[:, 8:12].sum(1), cv[:, 12:16].sum(1)], 1)
        total_dyn = ((dyn[:, 0] + dyn[:, 1]) + dyn[:, 2]) + dyn[:, 3]
        steam_pen = np.maximum(f(0), self.fc[:, 0, 4] - Y[:, 28])
        energy_pen = np.maximum(f(0), self.fc[:, 0, 3] - Y[:, 27])
        non_delivery = (f(self.ip) * (steam_pen + energy_pen)).astype(f)
        reward = -(((Y[:, 21] + total_ramp) + non_delivery).astype(np.float64) + total_dyn)
        agent_rewards = -(fuel.astype(np.float64) + ramp + dyn + (non_delivery / f(4))[:, None].astype(np.float64))
        self.prev = a.copy()
        self.t = self.t + 1
        self.fc = self.forecast(self.day, self.t, noise)
        term = self.t >= STEPS_PER_DAY
        info = {"Y": Y, "X": X, "fuel": fuel, "ramp": ramp, "dyn_cv": dyn, "non_delivery": non_delivery,
                "agent_rewards": agent_rewards}
        return self.obs(), reward, term, np.zeros(self.B, bool), info


def flatten_obs(obs: dict) -> np.ndarray:
    """gymnasium.spaces.flatten of the observation Dict (keys sorted; Discrete -> one-hot), as f32.
    (The reference's flat obs is float64 because int64 one-hots meet f32 boxes; values are equal.)"""
    B = len(obs["Time"])
    pa = obs["Prev_Action"]
    idx = {k: i for i, k in enumerate(ACTION_KEYS)}
    parts = []
    for key in sorted(["Energy_Price", "Gas_Price", "PAMB", "Prev_Action", "RHAMB", "TAMB", "Target_Power",
                       "Target_Steam", "Time"]):
        if key != "Prev_Action":
            parts.append(np.asarray(obs[key], np.float32).reshape(B, -1))
            continue
        for ak in sorted(ACTION_KEYS):
            v = pa[:, idx[ak]]
            if ak == "CT_NrBays":
                oh = np.zeros((B, 12), np.float32); oh[np.arange(B), v.astype(int) - 1] = 1; parts.append(oh)
            elif ak.endswith("_FFU"):
                oh = np.zeros((B, 2), np.float32); oh[np.arange(B), v.astype(int)] = 1; parts.append(oh)
            else:
                parts.append(v[:, None].astype(np.float32))
    return np.concatenate(parts, axis=1)
