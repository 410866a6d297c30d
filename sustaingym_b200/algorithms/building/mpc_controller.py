"""Model-predictive controller for BuildingEnv (sustaingym/algorithms/building/mpc_controller.py:9-116).

The reference states the plan as a cvxpy program solved by ECOS_BB: over ``planning_steps`` periods, with the weather and
occupant heat frozen at the current epoch,

    minimise   sum_t  beta * || ac_map * (x_t - target) ||_p  +  (1 - beta) * 24 * || u_t ||_p        t = 0 .. P-1
    subject to x_0 = state,  x_{t+1} = A_d x_t + B_d[:, 3:-1] u_t + B_d[:, 2] T_out + B_d[:, 1] T_ground
                                        + B_d[:, 0] occupower + B_d[:, -1] ghi,        -1 <= u_t <= 1

(so the state reached by the LAST action is not priced, and with ``planning_steps = 1`` the plan is u = 0: that is the
reference's formulation, kept).  Neither cvxpy nor ECOS exists in this image.  Here the program is solved by HiGHS
(SciPy): for p = 1 it IS a linear program (absolute values by epigraph variables); for p = 2 every norm gets an epigraph
variable bounded below by supporting hyperplanes that are added where the current plan violates them (Kelley's cutting
planes).  Each round yields a lower bound (the LP value) and an upper bound (the true objective of the LP's plan, which
is always feasible), and the loop ends when they agree to ``gap_tol`` — the returned plan carries that certificate in
``self.gap``.  Parity with ECOS: unpinned (optimal value within the gap; the optimal plan is unique for p = 2 when the
norms are non-zero).  The reference's ``MPCAgent_DataDriven`` (a different, learnt B_d layout) is not restated: no env of
the reference produces that layout.
"""
from __future__ import annotations

import numpy as np


class MPCAgent:
    """``MPCAgent(env, beta, pnorm, safety_margin=0.9, planning_steps=1).predict() -> (action, predicted next state)``."""

    def __init__(self, env, beta: float, pnorm: float, safety_margin: float = 0.9, planning_steps: int = 1,
                 gap_tol: float = 1e-7, max_rounds: int = 400):
        assert pnorm in (1, 2), "pnorm 1 (linear program) or 2 (cutting planes)"
        self.env = env
        self.beta = beta
        self.pnorm = pnorm
        self.safety_margin = safety_margin           # (unused by the reference's program as well)
        self.planning_steps = planning_steps
        self.gap_tol, self.max_rounds = gap_tol, max_rounds
        self.action_dim: int = env.action_space.shape[0]
        self.temp = float(env.out_temp[env.epoch])
        self.ground_temp = float(env.ground_temp[env.epoch])
        self.metabolism = float(env.metabolism[env.epoch])
        self.ghi = float(env.ghi[env.epoch])
        self.gap = 0.0
        self.rounds = 0

    def predict(self):
        from scipy.optimize import linprog
        env, P = self.env, self.planning_steps
        A_d, B_d, n, m = np.asarray(env.A_d), np.asarray(env.B_d), env.n, self.action_dim
        self.temp = float(env.out_temp[env.epoch])
        self.ground_temp = float(env.ground_temp[env.epoch])
        self.metabolism = float(env.metabolism[env.epoch])
        self.ghi = float(env.ghi[env.epoch])
        x0 = np.asarray(env.state[:n], dtype=np.float64)
        avg_temp, Meta = float(np.sum(x0) / n), self.metabolism
        self.Occupower = (6.461927 + 0.946892 * Meta + 0.0000255737 * Meta**2 - 0.0627909 * avg_temp * Meta
                          + 0.0000589172 * avg_temp * Meta**2 - 0.19855 * avg_temp**2
                          + 0.000940018 * avg_temp**2 * Meta - 0.00000149532 * avg_temp**2 * Meta**2)
        drift = (B_d[:, 2] * self.temp + B_d[:, 1] * self.ground_temp + B_d[:, 0] * self.Occupower + B_d[:, -1] * self.ghi)
        Bu = B_d[:, 3:-1]
        ac, target = np.asarray(env.ac_map, np.float64), np.asarray(env.target, np.float64)
        wx, wu = self.beta, (1 - self.beta) * 24
        # variables: x_1..x_P | u_0..u_{P-1} | then epigraph variables
        ox, ou = 0, n * P
        nxu = n * P + m * P
        Aeq = np.zeros((n * P, nxu)); beq = np.zeros(n * P)
        for t in range(P):                              # x_{t+1} - A_d x_t - Bu u_t = drift  (x_0 is data)
            r = slice(n * t, n * (t + 1))
            Aeq[r, ox + n * t: ox + n * (t + 1)] = np.eye(n)
            Aeq[r, ou + m * t: ou + m * (t + 1)] = -Bu
            if t == 0:
                beq[r] = drift + A_d @ x0
            else:
                Aeq[r, ox + n * (t - 1): ox + n * t] = -A_d
                beq[r] = drift
        const = wx * np.linalg.norm(ac * (x0 - target), self.pnorm)          # the t = 0 state term
        lo = np.concatenate([np.full(n * P, -np.inf), -np.ones(m * P)])
        hi = np.concatenate([np.full(n * P, np.inf), np.ones(m * P)])

        def residuals(z):          # the vectors under the norms: priced states x_1..x_{P-1}, actions u_0..u_{P-1}
            xs = [ac * (z[ox + n * (t - 1): ox + n * t] - target) for t in range(1, P)]
            us = [z[ou + m * t: ou + m * (t + 1)] for t in range(P)]
            return xs, us

        if self.pnorm == 1:
            # |v_i| <= e_i for every priced component
            ne = n * (P - 1) + m * P
            nv = nxu + ne
            rows, rhs = [], []
            k = 0
            for t in range(1, P):
                for i in range(n):
                    for sgn in (1.0, -1.0):
                        row = np.zeros(nv); row[ox + n * (t - 1) + i] = sgn * ac[i]; row[nxu + k] = -1.0
                        rows.append(row); rhs.append(sgn * ac[i] * target[i])
                    k += 1
            for t in range(P):
                for i in range(m):
                    for sgn in (1.0, -1.0):
                        row = np.zeros(nv); row[ou + m * t + i] = sgn; row[nxu + k] = -1.0
                        rows.append(row); rhs.append(0.0)
                    k += 1
            c = np.concatenate([np.zeros(nxu), np.full(n * (P - 1), wx), np.full(m * P, wu)])
            res = linprog(c, A_ub=np.array(rows), b_ub=np.array(rhs), A_eq=np.hstack([Aeq, np.zeros((n * P, ne))]), b_eq=beq,
                          bounds=np.stack([np.concatenate([lo, np.zeros(ne)]), np.concatenate([hi, np.full(ne, np.inf)])], 1),
                          method="highs")
            if res.status != 0:
                raise RuntimeError(f"MPC LP failed: {res.message}")
            z, self.gap, self.rounds = res.x[:nxu], 0.0, 1
            self.objective = res.fun + const
        else:
            ne = (P - 1) + P                           # one epigraph variable per norm
            nv = nxu + ne
            c = np.concatenate([np.zeros(nxu), np.full(P - 1, wx), np.full(P, wu)])
            rows, rhs = [], []

            def cut(kind, t, g):                        # g . v <= s   for the norm (kind, t), |g|_2 = 1
                row = np.zeros(nv)
                if kind == 0:
                    row[ox + n * (t - 1): ox + n * t] = g * ac
                    row[nxu + (t - 1)] = -1.0
                    rows.append(row); rhs.append(float(g @ (ac * target)))
                else:
                    row[ou + m * t: ou + m * (t + 1)] = g
                    row[nxu + (P - 1) + t] = -1.0
                    rows.append(row); rhs.append(0.0)

            bounds = np.stack([np.concatenate([lo, np.zeros(ne)]), np.concatenate([hi, np.full(ne, np.inf)])], 1)
            Aeq_f = np.hstack([Aeq, np.zeros((n * P, ne))])
            z = None
            for rnd in range(self.max_rounds):
                res = linprog(c, A_ub=np.array(rows) if rows else None, b_ub=np.array(rhs) if rows else None, A_eq=Aeq_f,
                              b_eq=beq, bounds=bounds, method="highs")
                if res.status != 0:
                    raise RuntimeError(f"MPC LP failed: {res.message}")
                z = res.x[:nxu]
                xs, us = residuals(z)
                upper = wx * sum(np.linalg.norm(v) for v in xs) + wu * sum(np.linalg.norm(v) for v in us)
                self.gap, self.rounds = upper - res.fun, rnd + 1
                if self.gap <= self.gap_tol * max(1.0, abs(upper)):
                    break
                for t, v in enumerate(xs, start=1):
                    nv_ = np.linalg.norm(v)
                    if nv_ > res.x[nxu + (t - 1)] + 1e-12:
                        cut(0, t, v / nv_)
                for t, v in enumerate(us):
                    nv_ = np.linalg.norm(v)
                    if nv_ > res.x[nxu + (P - 1) + t] + 1e-12:
                        cut(1, t, v / nv_)
            self.objective = upper + const
        u0 = z[ou: ou + m]
        state = z[ox: ox + n]
        if env.is_continuous_action:
            action = u0
        else:
            action = (u0 * 100 - env.Qlow * 100).astype(int)
        return action, state
