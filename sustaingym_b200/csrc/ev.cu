// EVChargingEnv batched step()/reset() and action projection for sm_100a.
//
// Replaces, per env and per step, the reference's
//   EVChargingEnv.step            sustaingym/envs/evcharging/env.py:229-291
//   _project_action (cvxpy+MOSEK) env.py:178-221, 473-500, sustaingym/envs/utils.py:6-24
//   _to_schedule                  env.py:340-379
//   acnsim Simulator.step -> ChargingNetwork.update_pilots -> EV.charge ->
//   Linear2StageBattery._charge_stepwise, event queue (Unplug < Plugin)   [acnportal 0.3.x]
//   _get_observation              env.py:381-394
//   _get_reward                   env.py:431-464
//   DiscreteActionWrapper.action  sustaingym/envs/wrappers.py:43-45
//
// One kernel launch per step (ev_step_main_kernel, documented at its definition): a warp owns a tile
// of 32 consecutive envs, the observation/pilot rows of the tile are written with coalesced vector
// stores by the whole warp, and one thread advances one env by walking only its OCCUPIED stations
// (64-bit occupancy mask + 32-byte session records, double-buffered).  The network constraints
// depend on the schedule only through <= 8 per-class sums, kept per thread.
//
// The projection has a fast path (box-clipped action already satisfies the network: it IS the
// projection), decided per thread with an f32 margin test, and a slow path run warp-cooperatively
// inside the same kernel for the rare envs that need it: exact test, exact single-class (pod)
// solve, else a primal-dual interior point on the QCQP whose Newton systems are reduced to
// n_classes x n_classes + n_constraints x n_constraints Cholesky factorisations.
//
// All divisions by constants use the FMA-corrected reciprocal (Markstein) which reproduces the
// IEEE quotient the reference's Python arithmetic produces (checked on 2e7 samples per constant,
// tests/test_div_const.py), so results stay bit-comparable with the oracle.
//
// Bound: HBM.  Algorithmic bytes per env-step: SURVEY.md 8d (1709 B at n=54, k=36).
#include <algorithm>

#include "common.cuh"

namespace sg {
namespace {

// ---- env.py:99-114 (evaluated in IEEE double exactly like the Python class attributes) --------
constexpr double kVoltage = 208.0;
constexpr double kPeriod = 5.0;                                   // minutes
constexpr double kPeriodH = kPeriod / 60.0;                       // hours
constexpr double kActionScale = 32.0;
constexpr double kMarginalProfit = 0.15 * 0.20;
constexpr double kAMinsToKwh = (1.0 / 60.0) * (kVoltage / 1000.0);
constexpr double kAPersToKwh = kAMinsToKwh * 5.0;
constexpr double kProfitFactor = kAPersToKwh * kMarginalProfit;
constexpr double kViolationFactor = kAPersToKwh * 0.001;
constexpr double kCarbonFactor = kAPersToKwh * (30.85 / 1000.0);
// event_generation.py:60-62 and acnportal defaults
constexpr double kCapacity = 100.0, kMaxPower = 100.0, kTransitionSoc = 0.8, kFullEps = 1e-3;
constexpr double kOneMinusTrans = 1.0 - kTransitionSoc;
constexpr int kEmpty = 32767;
constexpr int kEpisode = SG_EV_EPISODE_STEPS;
// correctly rounded reciprocals for div_c
constexpr double kInv1000 = 1.0 / 1000.0, kInv208 = 1.0 / kVoltage, kInv100 = 1.0 / kCapacity;
constexpr double kInvOMT = 1.0 / kOneMinusTrans, kInvPeriodH = 1.0 / kPeriodH, kInvAPers = 1.0 / kAPersToKwh;
// fixed-point scale and safety margin of the fast-path feasibility test (normalised units)
constexpr double kFixScale = 1048576.0, kFixInv = 1.0 / 1048576.0, kFeasMargin = 1e-4;

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kMC = SG_EV_MAX_CLASSES;

// a / b for a compile-time constant b with inv = RN(1/b): q' = RN(q + RN(a - q b) * inv)
__device__ __forceinline__ double div_c(double a, double b, double inv) {
    const double q = a * inv;
    const double r = fma(-q, b, a);
    return fma(r, inv, q);
}
// ~1e-15-accurate reciprocal / reciprocal square root for the interior-point iteration, which is
// self-correcting (residuals are recomputed exactly every iteration)
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = r * fma(-x, r, 2.0);
    r = r * fma(-x, r, 2.0);
    return r * fma(-x, r, 2.0);
}
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = y * fma(-0.5 * x * y, y, 1.5);
    y = y * fma(-0.5 * x * y, y, 1.5);
    return y * fma(-0.5 * x * y, y, 1.5);
}

struct EvShared {
    double con_re[SG_EV_MAX_CONSTRAINTS * kMC];
    double con_im[SG_EV_MAX_CONSTRAINTS * kMC];
    double lim[SG_EV_MAX_CONSTRAINTS];     // magnitudes in A
    uint8_t cls[SG_EV_MAX_STATIONS];
    uint8_t type[SG_EV_MAX_STATIONS];
    double stat[11 * kWarpsPerBlock];
};

// The host pads con_re / con_im to [SG_EV_MAX_CONSTRAINTS][SG_EV_MAX_CLASSES] and magnitudes to
// [SG_EV_MAX_CONSTRAINTS] with zeros, so staging is a straight copy (no index arithmetic).
__device__ __forceinline__ void load_shared(const sg_ev_model& m, EvShared& sh) {
    for (int i = threadIdx.x; i < SG_EV_MAX_CONSTRAINTS * kMC; i += blockDim.x) {
        sh.con_re[i] = m.con_re[i];
        sh.con_im[i] = m.con_im[i];
    }
    for (int i = threadIdx.x; i < SG_EV_MAX_CONSTRAINTS; i += blockDim.x) sh.lim[i] = m.magnitudes[i];
    for (int i = threadIdx.x; i < SG_EV_MAX_STATIONS; i += blockDim.x) {
        const bool in = i < m.n_stations;
        sh.cls[i] = in ? m.station_class[i] : 255;
        sh.type[i] = in ? m.station_type[i] : 0;
    }
    __syncthreads();
}

// class sums of a per-station f64 value (two stations per lane) -> every lane gets all sums
template <int MAXC>
__device__ __forceinline__ void class_sums(const double (&v)[2], const int (&cls)[2], int C, double (&S)[MAXC]) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        if (c < C) S[c] = warp_sum((cls[0] == c ? v[0] : 0.0) + (cls[1] == c ? v[1] : 0.0));
        else S[c] = 0.0;
    }
}

// class sums of integer pilots (<= 32 A, <= 64 stations => < 2^16): two classes per REDUX
__device__ __forceinline__ void class_sums_pilots(const unsigned (&v)[2], const int (&cls)[2], int C, double (&S)[kMC]) {
#pragma unroll
    for (int w = 0; w < kMC / 2; ++w) {
        if (2 * w < C) {
            unsigned part = 0;
#pragma unroll
            for (int p = 0; p < 2; ++p)
                if ((cls[p] >> 1) == w) part += v[p] << ((cls[p] & 1) * 16);
            const unsigned tot = __reduce_add_sync(kFull, part);
            S[2 * w] = (double)(tot & 0xffffu);
            S[2 * w + 1] = (double)(tot >> 16);
        } else { S[2 * w] = 0.0; S[2 * w + 1] = 0.0; }
    }
}

// env.py:373-378: pilot in A from the normalised action
__device__ __forceinline__ double to_pilot(double x, int type) {
    const double amps = x * kActionScale;
    if (type == 0) return amps >= 6.0 ? rint(amps) : 0.0;     // np.round = half-to-even = rint
    return rint(amps * 0.125) * 8.0;                          // /8 is exact
}

// env.py:188-189: min(1, demands / A_PERS_TO_KWH / 32)
__device__ __forceinline__ double upper_bound(float demand_f32) {
    return fmin(1.0, div_c((double)demand_f32, kAPersToKwh, kInvAPers) * (1.0 / kActionScale));
}

// =================================================================================================
// Interior-point projection (slow path).  Same iteration as oracle/ev.py::_ipm, but the Newton
// system (D + M^T K M) dx = rho is solved through the C x C SPD system (I + H K H) v = H^-1 g
// (Woodbury on the class-indicator structure), Cholesky with reciprocal square roots.
// Returns the iteration count; *hit_cap is set when the cap was reached.
// =================================================================================================
template <int MAXC, int MAXJ>
__device__ __noinline__ int ipm_project(const EvShared& sh, const int C, const int J, const double (&a)[2],
                                         const double (&u)[2], const int (&cls)[2], const bool (&valid)[2],
                                         double (&x)[2], bool* hit_cap) {
    const double tol = 1e-13, tau = 0.995;
    const int max_iter = 60;
    bool fr[2];
    double uf[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) { fr[p] = valid[p] && u[p] > 0.0; uf[p] = fr[p] ? u[p] : 1.0; }
    const int nfree = __popc(__ballot_sync(kFull, fr[0])) + __popc(__ballot_sync(kFull, fr[1]));
    const double inv_ncomp = 1.0 / (2.0 * nfree + J);
    double r2[MAXJ];
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) { const double r = (j < J) ? sh.lim[j] * (1.0 / kActionScale) : 1.0; r2[j] = r * r; }

    double S[MAXC];
    // strictly interior start: x = theta*u with every |A~ x| <= r/2
    {
        double uu[2] = {fr[0] ? u[0] : 0.0, fr[1] ? u[1] : 0.0};
        class_sums<MAXC>(uu, cls, C, S);
        double theta = 0.5;
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            if (j < J) {
                double wr = 0, wi = 0;
#pragma unroll
                for (int c = 0; c < MAXC; ++c) { wr = fma(sh.con_re[j * kMC + c], S[c], wr); wi = fma(sh.con_im[j * kMC + c], S[c], wi); }
                const double mag = sqrt(wr * wr + wi * wi);
                theta = fmin(theta, 0.5 * sqrt(r2[j]) / fmax(mag, 1e-300));
            }
        }
        x[0] = fr[0] ? theta * uf[0] : 0.0;
        x[1] = fr[1] ? theta * uf[1] : 0.0;
    }
    double mu = 0.1;
    double zl[2], zu[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) { zl[p] = fr[p] ? mu / x[p] : 0.0; zu[p] = fr[p] ? mu / (uf[p] - x[p]) : 0.0; }
    double lam[MAXJ], wre[MAXJ], wim[MAXJ], q[MAXJ];
    class_sums<MAXC>(x, cls, C, S);
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
        double wr = 0, wi = 0;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) { wr = fma(sh.con_re[j * kMC + c], S[c], wr); wi = fma(sh.con_im[j * kMC + c], S[c], wi); }
        lam[j] = (j < J) ? mu / (r2[j] - (wr * wr + wi * wi)) : 0.0;
    }

    int it = 0;
    *hit_cap = true;
    for (; it < max_iter; ++it) {
        class_sums<MAXC>(x, cls, C, S);
        double nu[MAXC];
#pragma unroll
        for (int c = 0; c < MAXC; ++c) nu[c] = 0.0;
        double lamq = 0.0;
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            if (j < J) {
                double wr = 0, wi = 0;
#pragma unroll
                for (int c = 0; c < MAXC; ++c) { wr = fma(sh.con_re[j * kMC + c], S[c], wr); wi = fma(sh.con_im[j * kMC + c], S[c], wi); }
                wre[j] = wr; wim[j] = wi;
                q[j] = r2[j] - (wr * wr + wi * wi);
                lamq = fma(lam[j], q[j], lamq);
#pragma unroll
                for (int c = 0; c < MAXC; ++c)
                    nu[c] = fma(2.0 * lam[j], wr * sh.con_re[j * kMC + c] + wi * sh.con_im[j * kMC + c], nu[c]);
            } else { wre[j] = wim[j] = 0.0; q[j] = 1.0; }
        }
        double xs[2], us[2], part = 0.0, rmax = 0.0;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            xs[p] = fr[p] ? x[p] : 1.0;
            us[p] = fr[p] ? uf[p] - x[p] : 1.0;
            double nuc = 0.0;
#pragma unroll
            for (int c = 0; c < MAXC; ++c) nuc = (cls[p] == c) ? nu[c] : nuc;
            const double rd = fr[p] ? (x[p] - a[p]) - zl[p] + zu[p] + nuc : 0.0;
            part += fr[p] ? zl[p] * xs[p] + zu[p] * us[p] : 0.0;
            rmax = fmax(rmax, fabs(rd));
        }
        const double gap = (warp_sum(part) + lamq) * inv_ncomp;
        const double resid = warp_max(rmax);
        if (gap < tol && resid < 100.0 * tol) { *hit_cap = false; break; }
        mu = 0.1 * gap;
        // ---- reduced Newton system -----------------------------------------------------------
        // Unknowns: class-sum step ds = H v (H = E^(1/2), E_c = sum over free stations of 1/D_i) and
        // the constraint multipliers' step dlam.  With K0 = 2 sum_j lam_j Q_j, B_j = 2 H G_j,
        // T_j = q_j / lam_j, t_j = mu / lam_j - q_j:
        //      (I + H K0 H) v + B^T dlam = H^-1 g          (stationarity, box multipliers eliminated)
        //              -B v + T dlam   = t                 (linearised complementarity)
        // dlam is NOT eliminated through 1/q_j (that puts 4 lam/q G G^T ~ 1e16 into the matrix and
        // the Cholesky pivots cancel catastrophically near convergence); instead v is eliminated:
        //      (T + B A0^-1 B^T) dlam = t + B A0^-1 ghat,   A0 = I + H K0 H   (both SPD, well scaled)
        double G[MAXJ][MAXC], K0[MAXC][MAXC];
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
#pragma unroll
            for (int d = 0; d < MAXC; ++d) K0[c][d] = 0.0;
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
#pragma unroll
            for (int c = 0; c < MAXC; ++c) G[j][c] = (j < J) ? wre[j] * sh.con_re[j * kMC + c] + wim[j] * sh.con_im[j * kMC + c] : 0.0;
            if (j < J) {
                const double l2 = 2.0 * lam[j];
#pragma unroll
                for (int c = 0; c < MAXC; ++c)
#pragma unroll
                    for (int d = 0; d <= c; ++d) {
                        const double Q = sh.con_re[j * kMC + c] * sh.con_re[j * kMC + d] + sh.con_im[j * kMC + c] * sh.con_im[j * kMC + d];
                        K0[c][d] = fma(l2, Q, K0[c][d]);
                    }
            }
        }
        double ixs[2], ius[2], rho[2], inv[2], rhoD[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            ixs[p] = fast_rcp(xs[p]);
            ius[p] = fast_rcp(us[p]);
            const double Dd = fr[p] ? 1.0 + zl[p] * ixs[p] + zu[p] * ius[p] : 1.0;
            double nuc = 0.0;
#pragma unroll
            for (int c = 0; c < MAXC; ++c) nuc = (cls[p] == c) ? nu[c] : nuc;
            rho[p] = fr[p] ? -(x[p] - a[p]) - nuc + mu * ixs[p] - mu * ius[p] : 0.0;
            inv[p] = fr[p] ? fast_rcp(Dd) : 0.0;
            rhoD[p] = rho[p] * inv[p];
        }
        double E[MAXC], g[MAXC];
        class_sums<MAXC>(inv, cls, C, E);
        class_sums<MAXC>(rhoD, cls, C, g);
        double h[MAXC], L[MAXC][MAXC], what[MAXC];
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            const bool on = (c < C) && E[c] > 0.0;
            const double ih = on ? fast_rsqrt(E[c]) : 0.0;
            h[c] = on ? E[c] * ih : 0.0;
            what[c] = g[c] * ih;
        }
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
#pragma unroll
            for (int d = 0; d <= c; ++d) L[c][d] = h[c] * K0[c][d] * h[d] + (c == d ? 1.0 : 0.0);
        // Cholesky of A0 in place; the diagonal stores 1/L[d][d]
#pragma unroll
        for (int d = 0; d < MAXC; ++d) {
            double sum = L[d][d];
#pragma unroll
            for (int e = 0; e < d; ++e) sum = fma(-L[d][e], L[d][e], sum);
            const double il = fast_rsqrt(sum);
            L[d][d] = il;
#pragma unroll
            for (int c = d + 1; c < MAXC; ++c) {
                double s2 = L[c][d];
#pragma unroll
                for (int e = 0; e < d; ++e) s2 = fma(-L[c][e], L[d][e], s2);
                L[c][d] = s2 * il;
            }
        }
        // what = L^-1 ghat ;  W_j = L^-1 B_j
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            double sum = what[c];
#pragma unroll
            for (int e = 0; e < c; ++e) sum = fma(-L[c][e], what[e], sum);
            what[c] = sum * L[c][c];
        }
        double W[MAXJ][MAXC], Sm[MAXJ][MAXJ], rhs[MAXJ];
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
#pragma unroll
            for (int c = 0; c < MAXC; ++c) {
                double sum = 2.0 * G[j][c] * h[c];
#pragma unroll
                for (int e = 0; e < c; ++e) sum = fma(-L[c][e], W[j][e], sum);
                W[j][c] = sum * L[c][c];
            }
            const double il = (j < J) ? fast_rcp(lam[j]) : 1.0;
            double r = (j < J) ? mu * il - q[j] : 0.0;
#pragma unroll
            for (int c = 0; c < MAXC; ++c) r = fma(W[j][c], what[c], r);
            rhs[j] = r;
#pragma unroll
            for (int k2 = 0; k2 <= j; ++k2) {
                double sum = (k2 == j) ? ((j < J) ? q[j] * il : 1.0) : 0.0;
#pragma unroll
                for (int c = 0; c < MAXC; ++c) sum = fma(W[j][c], W[k2][c], sum);
                Sm[j][k2] = sum;
            }
        }
        // Cholesky of S (J x J) and solve for dlam
#pragma unroll
        for (int d = 0; d < MAXJ; ++d) {
            double sum = Sm[d][d];
#pragma unroll
            for (int e = 0; e < d; ++e) sum = fma(-Sm[d][e], Sm[d][e], sum);
            const double il = fast_rsqrt(sum);
            Sm[d][d] = il;
#pragma unroll
            for (int c = d + 1; c < MAXJ; ++c) {
                double s2 = Sm[c][d];
#pragma unroll
                for (int e = 0; e < d; ++e) s2 = fma(-Sm[c][e], Sm[d][e], s2);
                Sm[c][d] = s2 * il;
            }
        }
        double dlam[MAXJ];
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            double sum = rhs[j];
#pragma unroll
            for (int e = 0; e < j; ++e) sum = fma(-Sm[j][e], dlam[e], sum);
            dlam[j] = sum * Sm[j][j];
        }
#pragma unroll
        for (int j = MAXJ - 1; j >= 0; --j) {
            double sum = dlam[j];
#pragma unroll
            for (int e = j + 1; e < MAXJ; ++e) sum = fma(-Sm[e][j], dlam[e], sum);
            dlam[j] = sum * Sm[j][j];
        }
        // v = L^-T (what - W^T dlam) ; ds = H v ; y = K0 ds + 2 sum_j dlam_j G_j
        double v[MAXC], ds[MAXC], y[MAXC];
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            double sum = what[c];
#pragma unroll
            for (int j = 0; j < MAXJ; ++j) sum = fma(-W[j][c], dlam[j], sum);
            v[c] = sum;
        }
#pragma unroll
        for (int c = MAXC - 1; c >= 0; --c) {
            double sum = v[c];
#pragma unroll
            for (int e = c + 1; e < MAXC; ++e) sum = fma(-L[e][c], v[e], sum);
            v[c] = sum * L[c][c];
        }
#pragma unroll
        for (int c = 0; c < MAXC; ++c) ds[c] = h[c] * v[c];
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            double sum = 0.0;
#pragma unroll
            for (int d = 0; d < MAXC; ++d) sum = fma(d <= c ? K0[c][d] : K0[d][c], ds[d], sum);
#pragma unroll
            for (int j = 0; j < MAXJ; ++j) sum = fma(2.0 * dlam[j], G[j][c], sum);
            y[c] = sum;
        }
        // ---- directions; fraction-to-boundary step as alpha = tau / max(tau, worst ratio) ----------
        double dx[2], dzl[2], dzu[2], worst = 0.0;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            double yc = 0.0;
#pragma unroll
            for (int c = 0; c < MAXC; ++c) yc = (cls[p] == c) ? y[c] : yc;
            dx[p] = fr[p] ? (rho[p] - yc) * inv[p] : 0.0;
            dzl[p] = fr[p] ? mu * ixs[p] - zl[p] - zl[p] * ixs[p] * dx[p] : 0.0;
            dzu[p] = fr[p] ? mu * ius[p] - zu[p] + zu[p] * ius[p] * dx[p] : 0.0;
            if (fr[p]) {
                worst = fmax(worst, -dx[p] * ixs[p]);
                worst = fmax(worst, dx[p] * ius[p]);
                worst = fmax(worst, -dzl[p] * fast_rcp(zl[p]));
                worst = fmax(worst, -dzu[p] * fast_rcp(zu[p]));
            }
        }
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            if (j < J) {
                double dwr = 0, dwi = 0;
#pragma unroll
                for (int c = 0; c < MAXC; ++c) {
                    dwr = fma(sh.con_re[j * kMC + c], ds[c], dwr);
                    dwi = fma(sh.con_im[j * kMC + c], ds[c], dwi);
                }
                const double bq = wre[j] * dwr + wim[j] * dwi;
                const double c2 = dwr * dwr + dwi * dwi;
                const double den = bq + sqrt(fma(bq, bq, c2 * tau * q[j]));
                worst = fmax(worst, den * fast_rcp(q[j]));                 // tau*q/den >= alpha
                worst = fmax(worst, -dlam[j] * fast_rcp(lam[j]));
            }
        }
        worst = warp_max(worst);
        const double alpha = worst > tau ? tau * fast_rcp(worst) : 1.0;
        if (!(alpha > 0.0)) break;                       // NaN / breakdown guard: report as cap hit
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            x[p] = fma(alpha, dx[p], x[p]);
            zl[p] = fma(alpha, dzl[p], zl[p]);
            zu[p] = fma(alpha, dzu[p], zu[p]);
        }
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) lam[j] = fma(alpha, dlam[j], lam[j]);
    }
    return it;
}

__device__ __forceinline__ int ipm_dispatch(const EvShared& sh, int C, int J, const double (&a)[2],
                                            const double (&u)[2], const int (&cls)[2], const bool (&valid)[2],
                                            double (&x)[2], bool* hit_cap) {
    if (C <= 5 && J <= 8) return ipm_project<5, 8>(sh, C, J, a, u, cls, valid, x, hit_cap);
    return ipm_project<SG_EV_MAX_CLASSES, SG_EV_MAX_CONSTRAINTS>(sh, C, J, a, u, cls, valid, x, hit_cap);
}

// Exact projection when only "simple" constraints (a single class with a non-zero coefficient,
// e.g. the 80 A pods) are violated: x_i = clip(a_i - nu, 0, u_i) on that class with nu >= 0 such
// that the class sum meets the limit (KKT of the QCQP with every other multiplier zero).  nu is
// the root of a decreasing piecewise-linear function: safeguarded Newton, which terminates on
// the exact linear piece.  The caller re-checks all constraints afterwards and falls back to the
// interior-point solver if another constraint is (or became) violated.
__device__ __forceinline__ bool simple_project(const EvShared& sh, const int C, const int J, const double (&a)[2],
                                               const double (&u)[2], const int (&cls)[2], const bool (&valid)[2],
                                               double (&x)[2]) {
    double S[kMC];
    class_sums<kMC>(x, cls, C, S);
    bool any_fixed = false;
    for (int j = 0; j < J; ++j) {
        int cstar = -1, nz = 0;
        double wr = 0, wi = 0;
#pragma unroll
        for (int c = 0; c < kMC; ++c) {
            const double re = sh.con_re[j * kMC + c], im = sh.con_im[j * kMC + c];
            if (re != 0.0 || im != 0.0) { nz++; cstar = c; }
            wr = fma(re, S[c], wr); wi = fma(im, S[c], wi);
        }
        const double r = sh.lim[j] * (1.0 / kActionScale);
        if (!(sqrt(wr * wr + wi * wi) > r)) continue;        // satisfied
        if (nz != 1) return false;                           // a coupled constraint is violated -> IPM
        const double kappa = sqrt(sh.con_re[j * kMC + cstar] * sh.con_re[j * kMC + cstar] +
                                  sh.con_im[j * kMC + cstar] * sh.con_im[j * kMC + cstar]);
        const double R = r / kappa;                           // limit on the class sum
        bool in[2];
        double amax = 0.0;
#pragma unroll
        for (int p = 0; p < 2; ++p) { in[p] = valid[p] && cls[p] == cstar && u[p] > 0.0; if (in[p]) amax = fmax(amax, a[p]); }
        double lo = 0.0, hi = warp_max(amax), nu = 0.0;
        unsigned sig_prev = 0xffffffffu;
        for (int it = 0; it < 100; ++it) {
            double s_mid = 0.0, s_up = 0.0;
            unsigned code = 0;
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                if (in[p]) {
                    const double v = a[p] - nu;
                    if (v >= u[p]) { s_up += u[p]; code |= 2u << (2 * p); }
                    else if (v > 0.0) { s_mid += a[p]; code |= 1u << (2 * p); }
                }
            }
            const int n_mid = __popc(__ballot_sync(kFull, code & 1u)) + __popc(__ballot_sync(kFull, code & 4u));
            // classification signature (order-independent hash is enough to detect "no change")
            const unsigned sig = __reduce_xor_sync(kFull, code * 2654435761u + (unsigned)(threadIdx.x & 31) * 40503u * code);
            const double sum_mid = warp_sum(s_mid), sum_up = warp_sum(s_up);
            const double phi = sum_mid - n_mid * nu + sum_up;
            if (sig == sig_prev && n_mid > 0) break;          // same linear piece twice: nu is its exact root
            sig_prev = sig;
            if (phi > R) lo = nu; else hi = nu;
            double nu_new = n_mid > 0 ? (sum_mid + sum_up - R) / n_mid : 0.5 * (lo + hi);
            if (!(nu_new > lo && nu_new < hi)) { nu_new = 0.5 * (lo + hi); sig_prev = 0xffffffffu; }
            if (nu_new == nu) break;
            nu = nu_new;
        }
#pragma unroll
        for (int p = 0; p < 2; ++p)
            if (in[p]) x[p] = fmin(fmax(a[p] - nu, 0.0), u[p]);
        class_sums<kMC>(x, cls, C, S);
        any_fixed = true;
    }
    return any_fixed;
}

// Exact test: is the (box-clipped) action inside every network constraint?  x normalised.
__device__ __forceinline__ bool network_feasible(const EvShared& sh, int C, int J, const double (&x)[2],
                                                 const int (&cls)[2]) {
    double S[kMC];
    class_sums<kMC>(x, cls, C, S);
    bool ok = true;
    for (int j = 0; j < J; ++j) {
        double wr = 0, wi = 0;
#pragma unroll
        for (int c = 0; c < kMC; ++c) { wr = fma(sh.con_re[j * kMC + c], S[c], wr); wi = fma(sh.con_im[j * kMC + c], S[c], wi); }
        const double r = sh.lim[j] * (1.0 / kActionScale);
        ok = ok && !(sqrt(wr * wr + wi * wi) > r);
    }
    return ok;
}

// Same with a relative slack of 1e-12 (a point placed exactly on a constraint boundary by the
// simple projection may exceed it by a few ulps).
__device__ __forceinline__ bool network_feasible_tol(const EvShared& sh, int C, int J, const double (&x)[2],
                                                     const int (&cls)[2]) {
    double S[kMC];
    class_sums<kMC>(x, cls, C, S);
    bool ok = true;
    for (int j = 0; j < J; ++j) {
        double wr = 0, wi = 0;
#pragma unroll
        for (int c = 0; c < kMC; ++c) { wr = fma(sh.con_re[j * kMC + c], S[c], wr); wi = fma(sh.con_im[j * kMC + c], S[c], wi); }
        const double r = sh.lim[j] * (1.0 / kActionScale);
        ok = ok && !(sqrt(wr * wr + wi * wi) > r * (1.0 + 1e-12));
    }
    return ok;
}


// =================================================================================================
// Per-env state (include/sustaingym_b200.h, sg_ev_state).  A plugged-in session is one 32-byte
// record = one DRAM/L2 sector, read and written with single 256-bit accesses:
//     { f64 requested kWh, f64 delivered kWh, i32 departure, i32 est_departure, i32 arrival, i32 slot }
// kept per (env, station) in two copies: step t reads copy (t & 1) and writes copy (1 - t & 1).
// Only the occupied stations (bit set in the env's 64-bit occupancy mask) hold live records, so one
// step touches ~occupancy x 64 B of state instead of a dense [n] row, and because a step never
// overwrites its own inputs it can be re-executed for an env (the projection fallback below does
// that) without an undo log.
// =================================================================================================
struct __align__(32) Slot {
    double energy, delivered;
    int dep, est, arr, k;
};

__device__ __forceinline__ Slot ld_slot(const Slot* p) {
    unsigned long long a, b, c, d;
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    Slot s;
    s.energy = __longlong_as_double((long long)a);
    s.delivered = __longlong_as_double((long long)b);
    s.dep = (int)(unsigned)(c & 0xffffffffull); s.est = (int)(unsigned)(c >> 32);
    s.arr = (int)(unsigned)(d & 0xffffffffull); s.k = (int)(unsigned)(d >> 32);
    return s;
}
__device__ __forceinline__ void st_slot(Slot* p, const Slot& s) {
    const unsigned long long a = (unsigned long long)__double_as_longlong(s.energy);
    const unsigned long long b = (unsigned long long)__double_as_longlong(s.delivered);
    const unsigned long long c = (unsigned long long)(unsigned)s.dep | ((unsigned long long)(unsigned)s.est << 32);
    const unsigned long long d = (unsigned long long)(unsigned)s.arr | ((unsigned long long)(unsigned)s.k << 32);
    asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

constexpr int kMainWarps = 2;                        // warps per CTA of the main kernel
constexpr int kMainThreads = kMainWarps * 32;
constexpr int kItemCap = 256;                        // (env, station) work items staged per warp and chunk
constexpr float kPreMargin = 2e-3f;                  // normalised units; covers the f32 accumulation error

// result of one (env, occupied station) work item, consumed by the thread that owns the env
struct __align__(16) ItemRec {
    double rate;                                     // charging rate in A
    float x;                                         // normalised pilot before rounding (class sums of the projection test)
    unsigned meta;                                   // pilot | class << 8 | departed << 16 | station << 24
};

struct WarpShared {
    ItemRec items[kItemCap];                         // the key (env lane << 8 | station) of item j is staged in items[j].meta
    float S[kMC][32];                                // per-env class sums of the clipped action
    int P[kMC][32];                                  // per-env class sums of the integer pilots
    double xs[SG_EV_MAX_STATIONS];                   // projected action of the env being re-done
};

struct MainShared {
    float re[SG_EV_MAX_CONSTRAINTS * kMC], im[SG_EV_MAX_CONSTRAINTS * kMC];   // f32 copies for the pre-checks
    float r_in[SG_EV_MAX_CONSTRAINTS];               // (limit/32 - margin)^2 in normalised action units
    float lim2_lo[SG_EV_MAX_CONSTRAINTS];            // 0.999 * limit^2 (A^2): below it the rounded schedule has no excess
    double ppilot[40];                               // pilot (A) -> kW at 208 V, the reference's IEEE operation sequence
    WarpShared w[kMainWarps];
    double stat[11 * kMainWarps];
};

// One period of Linear2StageBattery._charge_stepwise + EV.charge for a plugged-in EV (capacity
// 100 kWh, max_power 100 kW, transition SoC 0.8), branch-free: returns the charging rate in A and
// advances `delivered`.  p_pilot = pilot * 208 / 1000 kW comes from a 33-entry table.
__device__ __forceinline__ double charge_period(const double energy, double& delivered, const double p_pilot) {
    const double charge = fmax(0.0, kCapacity - energy) + delivered;
    const double rate_to_full = div_c(kCapacity - charge, kPeriodH, kInvPeriodH);
    const double soc = div_c(charge, kCapacity, kInv100);
    const double cap2 = soc < kTransitionSoc ? kMaxPower : div_c(1.0 - soc, kOneMinusTrans, kInvOMT) * kMaxPower;
    const double power = fmin(fmin(p_pilot, cap2), rate_to_full);
    const double rate = div_c(power * 1000.0, kVoltage, kInv208);                     // A
    delivered += div_c(rate * kVoltage, 1000.0, kInv1000) * kPeriodH;                 // EV.charge
    return rate;
}

// env.py:373-378 on an f32 action: amps = a * 32 is exact in f32, so rintf == np.round.
__device__ __forceinline__ int pilot_f32(const float a, const int type) {
    const float amps = a * 32.f;
    const float pf = type == 0 ? (amps >= 6.f ? rintf(amps) : 0.f) : rintf(amps * 0.125f) * 8.f;
    return min(max(__float2int_rn(pf), 0), 32);      // table index safety for out-of-space actions
}

// env.py:188-189 + 363-378 for one occupied station, branch-free and exact: with q = dem / A_PERS_TO_KWH,
// 32 * clip(a, 0, min(1, q / 32)) == min(32 * max(a, 0), min(32, q)) because scaling by 32 is exact.
__device__ __forceinline__ int pilot_clipped(const float af, const float dem, const int type, float* x_out) {
    const double q = div_c((double)dem, kAPersToKwh, kInvAPers);
    const double amps = fmin((double)(fmaxf(af, 0.f) * 32.f), fmin(32.0, q));
    *x_out = (float)amps * 0.03125f;
    const double p = type == 0 ? (amps >= 6.0 ? rint(amps) : 0.0) : rint(amps * 0.125) * 8.0;
    return __double2int_rn(p);
}

__device__ __forceinline__ float timestep_frac(const int t) {     // env.py:392, t / 288 as f32 (exhaustively checked)
    return (float)((double)t * (1.0 / (double)kEpisode));
}

// Warp-cooperative exact projection of ONE env (lane l owns stations 2l, 2l+1): exact feasibility
// test of the box-clipped action, then the single-class (pod) projection, then the interior-point
// solver.  Writes the projected normalised action to xs[0..n) and returns the status bits (0: the
// clipped action is feasible, nothing to redo).
__device__ __noinline__ int project_env(const EvShared& sh, const int n, const int C, const int J, const long long b,
                                         const int lane, const void* __restrict__ action, const bool i64,
                                         const int bins, const Slot* __restrict__ cur,
                                         const unsigned long long occ_b, double* xs) {
    const int i0 = 2 * lane;
    const bool valid[2] = {i0 < n, i0 + 1 < n};
    double a[2], u[2], x[2];
    int cls[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int i = i0 + p;
        a[p] = 0.0; u[p] = 0.0; cls[p] = 255;
        if (valid[p]) {
            const float af = i64 ? __fdiv_rn((float)reinterpret_cast<const long long*>(action)[b * n + i], (float)(bins - 1))
                                 : reinterpret_cast<const float*>(action)[b * n + i];
            a[p] = (double)af;
            cls[p] = sh.cls[i];
            if ((occ_b >> i) & 1ull) {
                const Slot s = ld_slot(cur + i);
                const double rem = s.energy - s.delivered;
                const float dem = rem > kFullEps ? (float)rem : 0.f;          // obs 'demands', env.py:388
                u[p] = upper_bound(dem);
            }
        }
        x[p] = valid[p] ? fmin(fmax(a[p], 0.0), u[p]) : 0.0;
    }
    int st = 0;
    if (!network_feasible(sh, C, J, x, cls)) {
        bool cap = false;
        if (!(simple_project(sh, C, J, a, u, cls, valid, x) && network_feasible_tol(sh, C, J, x, cls))) {
#pragma unroll
            for (int p = 0; p < 2; ++p) x[p] = valid[p] ? fmin(fmax(a[p], 0.0), u[p]) : 0.0;
            ipm_dispatch(sh, C, J, a, u, cls, valid, x, &cap);
        }
        st = SG_EV_STATUS_PROJECTED | (cap ? SG_EV_STATUS_MAXITER : 0);
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) if (valid[p]) xs[i0 + p] = x[p];
    __syncwarp();
    return st;
}

// Re-execution of one env's occupied stations with the projected action xs (one thread, rare path).
// Inputs are read from the state copy of the current step (untouched by the first pass), outputs
// overwrite what the first pass wrote.  P: this env's column of the pilot class sums (stride 32).
struct RedoOut { double rate_sum; unsigned long long occ_new; };
__device__ __noinline__ RedoOut redo_env(const uint8_t* __restrict__ s_type, const uint8_t* __restrict__ s_cls,
                                          const double* __restrict__ ppilot, int* P, const double* xs, const int n,
                                          const int tn, const unsigned long long occ, const Slot* __restrict__ cur,
                                          Slot* __restrict__ nxt, float* __restrict__ row,
                                          uint8_t* __restrict__ pilot_row) {
    RedoOut out;
    out.rate_sum = 0.0;
    out.occ_new = occ;
    for (int c = 0; c < kMC; ++c) P[c * 32] = 0;
    for (unsigned long long mask = occ; mask; mask &= mask - 1) {
        const int i = __ffsll((long long)mask) - 1;
        Slot q = ld_slot(cur + i);
        const int pi = min(max(__double2int_rn(to_pilot(xs[i], s_type[i])), 0), 32);
        P[s_cls[i] * 32] += pi;
        out.rate_sum += charge_period(q.energy, q.delivered, ppilot[pi]);
        if (pilot_row != nullptr) pilot_row[i] = (uint8_t)pi;
        bool active = false;
        if (q.dep <= tn) {                                            // Unplug event (timestamp = departure)
            out.occ_new &= ~(1ull << i);
        } else {
            st_slot(nxt + i, q);
            active = q.energy - q.delivered > kFullEps;               // Interface.active_sessions()
        }
        if (row != nullptr) {
            row[i] = active ? (float)(q.energy - q.delivered) : 0.f;
            row[n + i] = active ? (float)(q.est - tn) : 0.f;
        }
    }
    return out;
}

// =================================================================================================
// Main kernel: one launch = one step() of all envs.  A warp owns a tile of 32 consecutive envs.
//
//   fill     the tile's observation rows (contiguous in memory) are written by the whole warp with
//            coalesced 8-byte stores: zeros for demands / est_departures (an empty station reads 0)
//            and the MOER row + timestep tail; pilot rows are zeroed (projection on: an empty station
//            has upper bound 0) or computed densely with their class sums (projection off);
//   items    the occupied stations of the 32 envs (bits of the per-env 64-bit masks) are enumerated
//            into one work list and processed one (env, station) item per lane, so the lanes stay
//            balanced whatever the per-env occupancy: one 32-byte session record (single 256-bit
//            load) + one action value in, battery arithmetic (branch-free, exact), the record goes
//            to the other state copy, the two observation entries and the pilot byte are patched
//            into the pre-filled rows, the item's rate / pilot / clipped action go to a record in
//            shared memory;
//   owner    thread e sums the records of env e in station order (deterministic f64 sums) and tests
//            the class sums of the clipped action against the network constraints in f32 with a
//            margin; an env that is not certainly feasible is handed to the warp-cooperative exact
//            projection (project_env) and its stations are re-done with the projected action (its
//            inputs are still intact in the current state copy);
//   arrivals Plugin events with timestamp <= t+1 from the day's arrival-sorted list;
//   epilogue excess current of the rounded schedule, reward, cumulative breakdown, clock, mask.
// =================================================================================================
template <int MODE>
__global__ void __launch_bounds__(kMainThreads, 8)
ev_step_main_kernel(const sg_ev_model m, const sg_ev_state s, const long long B, const void* __restrict__ action,
                    const int flags, const int bins, float* __restrict__ obs, double* __restrict__ reward,
                    uint8_t* __restrict__ done, uint8_t* __restrict__ pilots, float* __restrict__ final_obs,
                    uint8_t* __restrict__ status, double* __restrict__ stats) {
    __shared__ EvShared sh;
    __shared__ MainShared ms;
    const int n = m.n_stations, kf = m.moer_steps, OBS = 2 * n + kf + 2, J = m.n_constraints, C = m.n_classes;
    load_shared(m, sh);
    for (int i = threadIdx.x; i < SG_EV_MAX_CONSTRAINTS * kMC; i += blockDim.x) {
        ms.re[i] = (float)sh.con_re[i];
        ms.im[i] = (float)sh.con_im[i];
    }
    for (int i = threadIdx.x; i < SG_EV_MAX_CONSTRAINTS; i += blockDim.x) {
        const double lim = sh.lim[i];
        const double r = lim * (1.0 / kActionScale) - (double)kPreMargin;
        ms.r_in[i] = r > 0.0 ? (float)(r * r * 0.99999) : -1.f;
        ms.lim2_lo[i] = (float)(lim * lim * 0.999);
    }
    for (int i = threadIdx.x; i <= 32; i += blockDim.x) ms.ppilot[i] = div_c((double)i * kVoltage, 1000.0, kInv1000);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpShared& ws = ms.w[warp];
    const bool i64 = (flags & 1) == SG_EV_ACTION_I64;
    const float* __restrict__ moer = m.moer;
    const uint4* __restrict__ recs = reinterpret_cast<const uint4*>(m.sessions);
    Slot* slots = reinterpret_cast<Slot*>(s.slots);
    const long long ntiles = (B + 31) / 32;
    const long long wstride = (long long)gridDim.x * kMainWarps;
    const bool pair_rows = ((kf & 1) == 0);            // OBS even: rows are 8-byte aligned, tail pairs never straddle
    const int tp = (kf + 2) >> 1;                      // float2 pairs in a row's MOER + timestep tail
    StatsAcc acc;

    auto act_at = [&](const long long o) -> float {
        return i64 ? __fdiv_rn((float)reinterpret_cast<const long long*>(action)[o], (float)(bins - 1))   // wrappers.py:45
                   : __ldg(reinterpret_cast<const float*>(action) + o);
    };

    for (long long tile = (long long)blockIdx.x * kMainWarps + warp; tile < ntiles; tile += wstride) {
        const long long b0 = tile * 32;
        const int nv = (int)min(32ll, B - b0);
        const long long b = b0 + lane;
        const bool env = lane < nv;
        int4 clk = make_int4(kEpisode, 0, 0, 0);
        if (env) clk = *reinterpret_cast<const int4*>(s.clock + b * 4);
        const int t = clk.x, tn = t + 1;
        int day = clk.y, nidx = clk.w;
        const int md = clk.z;
        const bool live = env && t < kEpisode;
        unsigned long long occ = 0ull;
        double bd0 = 0.0, bd1 = 0.0, bd2 = 0.0;
        if (live) {
            occ = s.occ[b];
            if (t != 0) { const double* bd = s.breakdown + b * 3; bd0 = bd[0]; bd1 = bd[1]; bd2 = bd[2]; }
        }
        const bool dn = tn >= kEpisode;
        const bool autoreset = live && dn && (flags & SG_EV_FLAG_AUTORESET);
        const int md_out = autoreset ? (md + 1) % m.n_moer_days : md;
        const unsigned live_mask = __ballot_sync(kFull, live);
        const unsigned ar_mask = __ballot_sync(kFull, autoreset);
        const int moff = (md * SG_EV_MOER_ROWS + min(tn, SG_EV_MOER_ROWS - 1)) * SG_EV_MOER_COLS;
        const float tsf = timestep_frac(tn);
        float moer0f = 0.f;
        if (live) moer0f = __ldg(moer + moff);
        // the day's next two pending arrivals (prefetched: the arrivals phase comes last)
        const unsigned* arow = m.arrivals + (long long)day * (m.max_day_sessions + 1);
        unsigned ent = 0x7fffu, ent2 = 0x7fffu;
        if (live && !autoreset) {
            ent = __ldg(arow + nidx);
            if (nidx < m.max_day_sessions) ent2 = __ldg(arow + nidx + 1);
        }
        const int par = t & 1;
        const int info = tn | (par << 16) | ((int)autoreset << 17);
#pragma unroll
        for (int c = 0; c < kMC; ++c) { ws.S[c][lane] = 0.f; ws.P[c][lane] = 0; }

        // ---------------- fill: coalesced observation (and pilot) rows ------------------------------
        if (nv == 32 && live_mask == kFull && ar_mask == 0u && pair_rows) {
            float2* o2 = reinterpret_cast<float2*>(obs + b0 * OBS);
            const int rp = OBS >> 1;                                   // float2 pairs per row
            const float2 z2 = make_float2(0.f, 0.f);
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
                if (lane < n) stg_stream_f2(o2 + r * rp + lane, z2);
                if (lane + 32 < n) stg_stream_f2(o2 + r * rp + lane + 32, z2);
            }
            // tails: 32 rows x tp pairs, one pair per lane and iteration
            int row = lane / tp, col = lane - row * tp;
            const int drow = 32 / tp, dcol = 32 - drow * tp;
            for (int it = 0; it < tp; ++it) {
                const int moff_r = __shfl_sync(kFull, moff, row);
                const float ts_r = __shfl_sync(kFull, tsf, row);
                const int e = 2 * col;
                const float* p = moer + moff_r;
                float2 v;
                v.x = e < kf ? __ldg(p + 1 + e) : __ldg(p);
                v.y = e + 1 < kf ? __ldg(p + 2 + e) : (e + 1 == kf ? __ldg(p) : ts_r);
                stg_stream_f2(o2 + row * rp + n + col, v);
                row += drow; col += dcol;
                if (col >= tp) { col -= tp; row += 1; }
            }
            if (MODE != 0 && pilots != nullptr) {
                uint4* p4 = reinterpret_cast<uint4*>(pilots + b0 * n);  // 32 n bytes, 16-byte aligned
                for (int c = lane; c < 2 * n; c += 32) p4[c] = make_uint4(0u, 0u, 0u, 0u);
            }
        } else {
            for (int r = 0; r < nv; ++r) {
                if (!((live_mask >> r) & 1u)) continue;
                const bool ar = (ar_mask >> r) & 1u;
                const int moff_r = __shfl_sync(kFull, moff, r);
                const float ts_r = __shfl_sync(kFull, tsf, r);
                float* orow = ar ? (final_obs != nullptr ? final_obs + (b0 + r) * OBS : nullptr) : obs + (b0 + r) * OBS;
                auto fill_row = [&](float* dst, const float* row, const float tsv) {
                    for (int e = lane; e < OBS; e += 32) {
                        const int q = e - 2 * n;
                        dst[e] = q < 0 ? 0.f : (q < kf ? __ldg(row + 1 + q) : (q == kf ? __ldg(row) : tsv));
                    }
                };
                if (orow != nullptr) fill_row(orow, moer + moff_r, ts_r);
                if (ar) {                                  // reset observation of the next episode (env.py:293-338)
                    const int mdo_r = __shfl_sync(kFull, md_out, r);
                    fill_row(obs + (b0 + r) * OBS, moer + (long long)mdo_r * SG_EV_MOER_ROWS * SG_EV_MOER_COLS, 0.f);
                }
                if (MODE != 0 && pilots != nullptr)
                    for (int i = lane; i < n; i += 32) pilots[(b0 + r) * n + i] = 0;
            }
        }
        if constexpr (MODE == 0) {
            // every EVSE receives a pilot, occupied or not (env.py:368-378); class sums for the excess current
            for (int r = 0; r < nv; ++r) {
                if (!((live_mask >> r) & 1u)) continue;
                const int i0 = 2 * lane;
                unsigned pv[2] = {0u, 0u};
                int cl[2] = {255, 255};
#pragma unroll
                for (int p = 0; p < 2; ++p)
                    if (i0 + p < n) {
                        cl[p] = sh.cls[i0 + p];
                        pv[p] = (unsigned)pilot_f32(act_at((b0 + r) * n + i0 + p), sh.type[i0 + p]);
                    }
                if (pilots != nullptr) {
#pragma unroll
                    for (int p = 0; p < 2; ++p) if (i0 + p < n) pilots[(b0 + r) * n + i0 + p] = (uint8_t)pv[p];
                }
#pragma unroll
                for (int w = 0; w < kMC / 2; ++w) {
                    unsigned part = 0;
#pragma unroll
                    for (int p = 0; p < 2; ++p)
                        if ((cl[p] >> 1) == w) part += pv[p] << ((cl[p] & 1) * 16);
                    const unsigned tot = __reduce_add_sync(kFull, part);
                    if (lane == r) { ws.P[2 * w][lane] = (int)(tot & 0xffffu); ws.P[2 * w + 1][lane] = (int)(tot >> 16); }
                }
            }
        }
        __syncwarp();

        // ---------------- items: one (env, occupied station) per lane ------------------------------------
        const int cnt = __popcll(occ);
        int off = cnt;                                                 // inclusive scan over the lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(kFull, off, d);
            if (lane >= d) off += v;
        }
        const int total = __shfl_sync(kFull, off, 31);
        off -= cnt;
        double rate_sum = 0.0;
        unsigned long long occ_new = occ;
        for (int c0 = 0; c0 < total; c0 += kItemCap) {
            const int nitems = min(kItemCap, total - c0);
            {   // keys of this chunk
                int j = off - c0;
                for (unsigned long long mask = occ; mask; mask &= mask - 1, ++j)
                    if (j >= 0 && j < kItemCap) ws.items[j].meta = ((unsigned)lane << 8) | (unsigned)(__ffsll((long long)mask) - 1);
            }
            __syncwarp();
            constexpr int V = 2;
            for (int base = 0; base < nitems; base += 32 * V) {
                int e[V], si[V], inf[V];
                Slot q[V];
                float af[V];
                bool on[V];
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    const int j = base + u * 32 + lane;
                    on[u] = j < nitems;
                    const unsigned key = on[u] ? ws.items[j].meta : 0u;
                    e[u] = (int)(key >> 8); si[u] = (int)(key & 0xffu);
                    inf[u] = __shfl_sync(kFull, info, e[u]);
                }
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    af[u] = 0.f;
                    q[u].energy = 0.0; q[u].delivered = 0.0; q[u].dep = q[u].est = q[u].arr = q[u].k = 0;
                    if (on[u]) {
                        const long long be = b0 + e[u];
                        q[u] = ld_slot(slots + ((long long)((inf[u] >> 16) & 1) * B + be) * n + si[u]);
                        af[u] = act_at(be * n + si[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    if (!on[u]) continue;
                    const int i = si[u], tn_e = inf[u] & 0xffff, par_e = (inf[u] >> 16) & 1;
                    const bool ar_e = (inf[u] >> 17) & 1;
                    const long long be = b0 + e[u];
                    const int type = sh.type[i], cls = sh.cls[i];
                    int pi;
                    float xf = 0.f;
                    if (MODE == 0) {
                        pi = pilot_f32(af[u], type);
                    } else {
                        const double rem = q[u].energy - q[u].delivered;
                        const float dem = rem > kFullEps ? (float)rem : 0.f;              // obs 'demands', env.py:388
                        pi = pilot_clipped(af[u], dem, type, &xf);
                        if (pilots != nullptr && pi != 0) pilots[be * n + i] = (uint8_t)pi;
                    }
                    ItemRec rec;
                    rec.rate = charge_period(q[u].energy, q[u].delivered, ms.ppilot[pi]);
                    rec.x = xf;
                    const bool departed = q[u].dep <= tn_e;                               // Unplug event (timestamp = departure)
                    rec.meta = (unsigned)pi | ((unsigned)cls << 8) | ((unsigned)departed << 16) | ((unsigned)i << 24);
                    ws.items[base + u * 32 + lane] = rec;
                    if (!departed) {
                        st_slot(slots + ((long long)(par_e ^ 1) * B + be) * n + i, q[u]);
                        const double rem = q[u].energy - q[u].delivered;
                        float* row = ar_e ? final_obs : obs;                              // terminal observation only if asked for
                        if (rem > kFullEps && row != nullptr) {                           // Interface.active_sessions()
                            row[be * OBS + i] = (float)rem;
                            row[be * OBS + n + i] = (float)(q[u].est - tn_e);
                        }
                    }
                }
            }
            __syncwarp();
            // owner: thread e sums the records of env e in station order
            {
                const int lo = max(off - c0, 0), hi = min(off + cnt - c0, nitems);
                for (int j = lo; __any_sync(kFull, j < hi); ++j) {
                    if (j < hi) {
                        const ItemRec rec = ws.items[j];
                        const int cls = (rec.meta >> 8) & 0xff;
                        rate_sum += rec.rate;
                        if (MODE != 0) {
                            ws.S[cls][lane] += rec.x;
                            ws.P[cls][lane] += (int)(rec.meta & 0xffu);
                        }
                        if ((rec.meta >> 16) & 1u) occ_new &= ~(1ull << (rec.meta >> 24));
                    }
                }
            }
            __syncwarp();
        }

        // ---------------- network constraints on the clipped action (projection only) -------------------
        int st = 0;
        if constexpr (MODE != 0) {
            bool ok = true;
            if (live && cnt > 0) {
                float Sf[kMC];
#pragma unroll
                for (int c = 0; c < kMC; ++c) Sf[c] = ws.S[c][lane];
                for (int j = 0; j < J; ++j) {
                    float wr = 0.f, wi = 0.f;
#pragma unroll
                    for (int c = 0; c < kMC; ++c) { wr = fmaf(ms.re[j * kMC + c], Sf[c], wr); wi = fmaf(ms.im[j * kMC + c], Sf[c], wi); }
                    ok = ok && (fmaf(wr, wr, wi * wi) <= ms.r_in[j]);
                }
            }
            unsigned bad = __ballot_sync(kFull, !ok);
            while (bad) {
                const int r = __ffs((int)bad) - 1;
                bad &= bad - 1;
                const unsigned long long occ_r = __shfl_sync(kFull, occ, r);
                const int par_r = (__shfl_sync(kFull, info, r) >> 16) & 1;
                const Slot* cur_r = slots + ((long long)par_r * B + (b0 + r)) * n;
                const int st_r = project_env(sh, n, C, J, b0 + r, lane, action, i64, bins, cur_r, occ_r, ws.xs);
                if (lane == r) {
                    st = st_r;
                    if (st_r != 0) {
                        float* row = autoreset ? final_obs : obs;
                        const RedoOut ro = redo_env(sh.type, sh.cls, ms.ppilot, &ws.P[0][lane], ws.xs, n, tn, occ, cur_r,
                                                    slots + ((long long)(par_r ^ 1) * B + b) * n,
                                                    row != nullptr ? row + b * OBS : nullptr,
                                                    pilots != nullptr ? pilots + b * n : nullptr);
                        rate_sum = ro.rate_sum;
                        occ_new = ro.occ_new;
                    }
                }
                __syncwarp();
            }
        }

        // ---------------- arrivals: Plugin events with timestamp <= t+1 (after the Unplug events) -------
        if ((int)(ent & 0x7fffu) <= tn) {
            float* row = autoreset ? final_obs : obs;
            Slot* nxt = slots + ((long long)(par ^ 1) * B + b) * n;
            do {
                const int i = (ent >> 16) & 0xff, k = ent >> 24;
                const uint4 rr = __ldg(recs + ((long long)day * m.max_sessions + k) * n + i);
                Slot ns;
                ns.arr = (int)(short)(rr.x & 0xffffu);
                ns.dep = (int)(short)(rr.x >> 16);
                ns.est = (int)(short)(rr.y & 0xffffu);
                ns.k = k;
                ns.energy = __hiloint2double((int)rr.w, (int)rr.z);
                ns.delivered = 0.0;
                if (ns.dep > tn) {                                        // (zero-length sessions never charge)
                    occ_new |= 1ull << i;
                    st_slot(nxt + i, ns);
                    if (row != nullptr && ns.energy > kFullEps) {
                        row[b * OBS + i] = (float)ns.energy;
                        row[b * OBS + n + i] = (float)(ns.est - tn);
                    }
                }
                ++nidx;
                ent = ent2;
                ent2 = (int)(ent & 0x7fffu) <= tn ? __ldg(arow + nidx + 1) : 0x7fffu;
            } while ((int)(ent & 0x7fffu) <= tn);
        }

        // ---------------- per-env epilogue: excess current, reward, clock ---------------------------------
        if (live) {
            double excess = 0.0;
            {
                float Pf[kMC];
#pragma unroll
                for (int c = 0; c < kMC; ++c) Pf[c] = (float)ws.P[c][lane];
                bool near = false;
                for (int j = 0; j < J; ++j) {
                    float wr = 0.f, wi = 0.f;
#pragma unroll
                    for (int c = 0; c < kMC; ++c) { wr = fmaf(ms.re[j * kMC + c], Pf[c], wr); wi = fmaf(ms.im[j * kMC + c], Pf[c], wi); }
                    near = near || (fmaf(wr, wr, wi * wi) > ms.lim2_lo[j]);
                }
                if (near) {                                               // exact (env.py:449-452)
                    for (int j = 0; j < J; ++j) {
                        double wr = 0, wi = 0;
#pragma unroll
                        for (int c = 0; c < kMC; ++c) {
                            wr = fma(sh.con_re[j * kMC + c], (double)ws.P[c][lane], wr);
                            wi = fma(sh.con_im[j * kMC + c], (double)ws.P[c][lane], wi);
                        }
                        const double lim = sh.lim[j], sq = fma(wr, wr, wi * wi);
                        if (sq > lim * lim) excess += fmax(0.0, sqrt(sq) - lim);
                    }
                }
            }
            const double profit = kProfitFactor * rate_sum;
            const double excess_charge = excess * kViolationFactor;
            const double carbon = kCarbonFactor * rate_sum * (double)moer0f;
            const double rew = profit - carbon - excess_charge;
            // cumulative reward breakdown (env.py:460-462); restarted on the first step of an episode so that
            // after an auto-reset the finished episode's totals stay readable until the next step
            double* bd = s.breakdown + b * 3;
            bd[0] = bd0 + profit; bd[1] = bd1 + carbon; bd[2] = bd2 + excess_charge;
            int t_out = tn;
            if (autoreset) {
                day = (day + 1) % m.n_days;                               // sequential days, event_generation.py:284-291
                t_out = 0; nidx = 0; occ_new = 0ull;
            }
            s.occ[b] = occ_new;
            *reinterpret_cast<int4*>(s.clock + b * 4) = make_int4(t_out, day, md_out, nidx);
            reward[b] = rew;
            done[b] = dn ? 1 : 0;
            if (status != nullptr) status[b] = (uint8_t)st;
            if (stats != nullptr) {
                acc.add(rew, dn, profit, carbon, excess_charge);
                acc.n_proj += (st & SG_EV_STATUS_PROJECTED) ? 1.0 : 0.0;
                acc.n_maxiter += (st & SG_EV_STATUS_MAXITER) ? 1.0 : 0.0;
            }
        } else if (env) {                                                 // finished episode, not reset: stays terminal
            reward[b] = 0.0; done[b] = 1;
            if (status != nullptr) status[b] = 0;
        }
        __syncwarp();
    }
    if (stats != nullptr) {
        __syncthreads();
        stats_flush_block(acc, stats, ms.stat);
    }
}


__global__ void ev_reset_kernel(const sg_ev_model m, const sg_ev_state s, const long long B,
                                const uint8_t* __restrict__ mask, const int32_t* __restrict__ day_in,
                                const int32_t* __restrict__ moer_day_in, float* __restrict__ obs) {
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    if (mask != nullptr && mask[b] == 0) return;
    const int n = m.n_stations, kf = m.moer_steps, OBS = 2 * n + kf + 2;
    int day = day_in ? day_in[b] : s.clock[b * 4 + 1];
    int md = moer_day_in ? moer_day_in[b] : s.clock[b * 4 + 2];
    day = ((day % m.n_days) + m.n_days) % m.n_days;
    md = ((md % m.n_moer_days) + m.n_moer_days) % m.n_moer_days;
    float* orow = obs + b * OBS;
    for (int i = lane; i < 2 * n; i += 32) orow[i] = 0.f;
    const float* mrow = m.moer + ((long long)md * SG_EV_MOER_ROWS) * SG_EV_MOER_COLS;
    for (int e = lane; e < kf + 2; e += 32) orow[2 * n + e] = e < kf ? mrow[1 + e] : (e == kf ? mrow[0] : 0.f);
    if (lane == 0) {
        // empty garage: no occupied station, so no session record is live (the records need no clearing)
        s.clock[b * 4 + 0] = 0; s.clock[b * 4 + 1] = day; s.clock[b * 4 + 2] = md; s.clock[b * 4 + 3] = 0;
        s.occ[b] = 0ull;
        s.breakdown[b * 3] = s.breakdown[b * 3 + 1] = s.breakdown[b * 3 + 2] = 0.0;
    }
}

// Dense per-station views of the sparse state (debugging, checkpoint inspection, multi-agent info):
// delivered kWh and trace slot of the session plugged in at each station, 0 / 255 where empty.
__global__ void ev_export_kernel(const sg_ev_model m, const sg_ev_state s, const long long B,
                                 double* __restrict__ delivered, uint8_t* __restrict__ sess_idx) {
    const int n = m.n_stations;
    const long long total = B * n;
    const Slot* slots = reinterpret_cast<const Slot*>(s.slots);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / n;
        const int i = (int)(idx % n);
        const int t = s.clock[b * 4];
        const bool on = (s.occ[b] >> i) & 1ull;
        double d = 0.0;
        int k = 255;
        if (on) { const Slot q = ld_slot(slots + ((long long)(t & 1) * B + b) * n + i); d = q.delivered; k = q.k; }
        if (delivered != nullptr) delivered[idx] = d;
        if (sess_idx != nullptr) sess_idx[idx] = (uint8_t)k;
    }
}

// Stand-alone projection operator (env.py:200-221): one warp per problem.
__global__ void __launch_bounds__(128)
ev_project_kernel(const sg_ev_model m, const long long B, const float* __restrict__ action,
                  const float* __restrict__ demands, double* __restrict__ x_out, uint8_t* __restrict__ status,
                  int32_t* __restrict__ iters) {
    __shared__ EvShared sh;
    load_shared(m, sh);
    const int lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int n = m.n_stations, i0 = 2 * lane;
    const bool valid[2] = {i0 < n, i0 + 1 < n};
    double a[2], u[2], x[2];
    int cls[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        a[p] = valid[p] ? (double)action[b * n + i0 + p] : 0.0;
        u[p] = valid[p] ? upper_bound(demands[b * n + i0 + p]) : 0.0;
        cls[p] = valid[p] ? sh.cls[i0 + p] : 255;
        x[p] = valid[p] ? fmin(fmax(a[p], 0.0), u[p]) : 0.0;
    }
    int st = 0, it = 0;
    if (!network_feasible(sh, m.n_classes, m.n_constraints, x, cls)) {
        bool cap = false;
        if (!(simple_project(sh, m.n_classes, m.n_constraints, a, u, cls, valid, x) &&
              network_feasible_tol(sh, m.n_classes, m.n_constraints, x, cls))) {
#pragma unroll
            for (int p = 0; p < 2; ++p) x[p] = valid[p] ? fmin(fmax(a[p], 0.0), u[p]) : 0.0;
            it = ipm_dispatch(sh, m.n_classes, m.n_constraints, a, u, cls, valid, x, &cap);
        }
        st = SG_EV_STATUS_PROJECTED | (cap ? SG_EV_STATUS_MAXITER : 0);
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) if (valid[p]) x_out[b * n + i0 + p] = x[p];
    if (lane == 0) { if (status) status[b] = (uint8_t)st; if (iters) iters[b] = it; }
}


// Per-agent observations of MultiAgentEVChargingEnv (multiagent_env.py:102-148): agent i sees the
// aggregate observation of `periods_delay` steps ago, except its own station's est_departure and
// demand entries, which are current.  out[b, i, :] for all 54 agents.
__global__ void ev_agent_obs_kernel(const float* __restrict__ cur, const float* __restrict__ past, const long long B,
                                    const int n, const int obs_dim, float* __restrict__ out) {
    const long long total = B * n * obs_dim;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(idx % obs_dim);
        const long long bi = idx / obs_dim;
        const int i = (int)(bi % n);
        const long long b = bi / n;
        const bool own = (e == i) || (e == n + i);
        out[idx] = own ? cur[b * obs_dim + e] : past[b * obs_dim + e];
    }
}

int check_model(const sg_ev_model* m) {
    SG_CHECK_ARG(m != nullptr, "ev: model is NULL");
    SG_CHECK_ARG(m->n_stations >= 1 && m->n_stations <= SG_EV_MAX_STATIONS, "ev: n_stations=%d outside [1,64]", m->n_stations);
    SG_CHECK_ARG(m->n_classes >= 1 && m->n_classes <= SG_EV_MAX_CLASSES, "ev: n_classes=%d outside [1,8]", m->n_classes);
    SG_CHECK_ARG(m->n_constraints >= 0 && m->n_constraints <= SG_EV_MAX_CONSTRAINTS, "ev: n_constraints=%d outside [0,16]", m->n_constraints);
    SG_CHECK_ARG(m->moer_steps >= 1 && m->moer_steps <= 36, "ev: moer_forecast_steps=%d outside [1,36]", m->moer_steps);
    SG_CHECK_ARG(m->station_class && m->station_type && m->con_re && m->con_im && m->magnitudes, "ev: NULL network table");
    return 0;
}

int check_state(const sg_ev_state* s) {
    SG_CHECK_ARG(s != nullptr && s->slots && s->clock && s->breakdown && s->occ, "ev: NULL state array");
    SG_CHECK_ARG((reinterpret_cast<uintptr_t>(s->slots) & 31u) == 0, "ev: state.slots must be 32-byte aligned");
    return 0;
}

}  // namespace
}  // namespace sg

extern "C" {

int sg_ev_reset(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs, const uint8_t* reset_mask,
                const int32_t* day_in, const int32_t* moer_day_in, float* obs, void* stream) {
    if (int rc = sg::check_model(m)) return rc;
    if (int rc = sg::check_state(s)) return rc;
    SG_CHECK_ARG(obs && m->moer, "ev_reset: NULL argument");
    SG_CHECK_ARG(m->n_days >= 1 && m->n_moer_days >= 1, "ev_reset: empty trace/MOER pool");
    if (num_envs <= 0) return 0;
    const unsigned grid = (unsigned)((num_envs + 3) / 4);
    sg::ev_reset_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*m, *s, num_envs, reset_mask,
                                                                                day_in, moer_day_in, obs);
    return sg::launch_status("ev_reset_kernel");
}

int sg_ev_step(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs, const void* action, int32_t flags,
               int32_t discrete_bins, float* obs, double* reward, uint8_t* done, uint8_t* pilots,
               float* final_obs, uint8_t* status, double* stats, void* stream) {
    if (int rc = sg::check_model(m)) return rc;
    if (int rc = sg::check_state(s)) return rc;
    SG_CHECK_ARG(action && obs && reward && done, "ev_step: NULL argument");
    SG_CHECK_ARG(m->sessions && m->moer && m->arrivals && m->n_days >= 1 && m->n_moer_days >= 1 && m->max_sessions >= 1,
                 "ev_step: empty trace/MOER pool");
    SG_CHECK_ARG(m->max_sessions <= 256, "ev_step: more than 256 session slots per station per day");
    SG_CHECK_ARG(!((flags & 1) == SG_EV_ACTION_I64) || discrete_bins >= 2, "ev_step: discrete_bins must be >= 2");
    if (num_envs <= 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int sms = sg_device_sm_count();
    if (sms <= 0) sms = 148;
    // one warp per 32-env tile; at most 8 resident CTAs per SM, larger batches loop (persistent)
    const long long ntiles = (num_envs + 31) / 32;
    const unsigned grid = (unsigned)std::min<long long>((ntiles + sg::kMainWarps - 1) / sg::kMainWarps, (long long)sms * 8);
    if (m->project)
        sg::ev_step_main_kernel<1><<<grid, sg::kMainThreads, 0, st>>>(*m, *s, num_envs, action, flags, discrete_bins, obs,
                                                                    reward, done, pilots, final_obs, status, stats);
    else
        sg::ev_step_main_kernel<0><<<grid, sg::kMainThreads, 0, st>>>(*m, *s, num_envs, action, flags, discrete_bins, obs,
                                                                    reward, done, pilots, final_obs, status, stats);
    return sg::launch_status("ev_step_main_kernel");
}

int sg_ev_step_host(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs, const void* h_action,
                    int32_t flags, int32_t discrete_bins, float* h_obs, double* h_reward, uint8_t* h_done,
                    void* d_action, float* d_obs, double* d_reward, uint8_t* d_done, uint8_t* d_status,
                    double* d_stats, void* stream) {
    SG_CHECK_ARG(m && h_action && h_obs && h_reward && h_done && d_action && d_obs && d_reward && d_done,
                 "ev_step_host: NULL argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t B = (size_t)num_envs, n = (size_t)m->n_stations, OBS = 2 * n + m->moer_steps + 2;
    const size_t abytes = B * n * (((flags & 1) == SG_EV_ACTION_I64) ? 8 : 4);
    SG_CUDA(cudaMemcpyAsync(d_action, h_action, abytes, cudaMemcpyHostToDevice, st));
    if (int rc = sg_ev_step(m, s, num_envs, d_action, flags, discrete_bins, d_obs, d_reward, d_done, nullptr,
                            nullptr, d_status, d_stats, stream)) return rc;
    SG_CUDA(cudaMemcpyAsync(h_obs, d_obs, B * OBS * 4, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(h_reward, d_reward, B * 8, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(h_done, d_done, B, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int sg_ev_export_state(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs, double* delivered,
                       uint8_t* sess_idx, void* stream) {
    if (int rc = sg::check_model(m)) return rc;
    if (int rc = sg::check_state(s)) return rc;
    if (num_envs <= 0) return 0;
    const long long total = (long long)num_envs * m->n_stations;
    const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148ll * 16);
    sg::ev_export_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*m, *s, num_envs, delivered, sess_idx);
    return sg::launch_status("ev_export_kernel");
}

int sg_ev_agent_obs(const float* cur_obs, const float* past_obs, int64_t num_envs, int32_t n_stations,
                    int32_t obs_dim, float* out, void* stream) {
    SG_CHECK_ARG(cur_obs && past_obs && out, "ev_agent_obs: NULL argument");
    SG_CHECK_ARG(n_stations >= 1 && obs_dim >= 2 * n_stations, "ev_agent_obs: bad dimensions");
    if (num_envs <= 0) return 0;
    const long long total = (long long)num_envs * n_stations * obs_dim;
    const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148ll * 16);
    sg::ev_agent_obs_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(cur_obs, past_obs, num_envs,
                                                                                   n_stations, obs_dim, out);
    return sg::launch_status("ev_agent_obs_kernel");
}

int sg_ev_project(const sg_ev_model* m, int64_t num_envs, const float* action, const float* demands,
                  double* x_out, uint8_t* status, int32_t* iters, void* stream) {
    if (int rc = sg::check_model(m)) return rc;
    SG_CHECK_ARG(action && demands && x_out, "ev_project: NULL argument");
    if (num_envs <= 0) return 0;
    const unsigned grid = (unsigned)((num_envs + 3) / 4);
    sg::ev_project_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*m, num_envs, action, demands,
                                                                                  x_out, status, iters);
    return sg::launch_status("ev_project_kernel");
}

}  // extern "C"
