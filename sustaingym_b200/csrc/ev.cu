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
// Mapping: one warp advances one env; lane l owns stations 2l and 2l+1 (54 stations -> 27
// lanes), so the env-major rows action[B,n] (f32), delivered[B,n] (f64), sess_idx[B,n] (u8) and
// obs[B,2n+k+2] (f32) are read/written as 8/16/2-byte vectors that are contiguous across the
// warp and across consecutive warps (fully coalesced, each byte touched once).  The network
// constraints depend on the schedule only through <= 8 per-class sums:
//   * the integer pilots are summed with packed 16-bit fields through REDUX (2 classes / word);
//   * the fast-path feasibility test of the projection sums 2^-20 fixed-point actions the same
//     way, with a safety margin; anything not certainly feasible is queued for the solver kernel.
// The small constraint tables are staged in shared memory once per CTA.  Trace (16-byte session
// records) and MOER tables are shared by all envs of a day and stay L2-resident.
//
// All divisions by constants use the FMA-corrected reciprocal (Markstein) which reproduces the
// IEEE quotient the reference's Python arithmetic produces (checked on 2e7 samples per constant,
// tests/test_div_const.py), so results stay bit-comparable with the oracle.
//
// The projection has a fast path (box-clipped action already satisfies the network: it IS the
// projection) executed inline, and a slow path (primal-dual interior point on the QCQP, Newton
// systems reduced to n_classes x n_classes by Woodbury) executed by a second, register-heavy
// kernel over a device-side work list — so the common path keeps a small register footprint.
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
    if (threadIdx.x < SG_EV_MAX_CONSTRAINTS) sh.lim[threadIdx.x] = m.magnitudes[threadIdx.x];
    if (threadIdx.x < SG_EV_MAX_STATIONS) {
        const bool in = (int)threadIdx.x < m.n_stations;
        sh.cls[threadIdx.x] = in ? m.station_class[threadIdx.x] : 255;
        sh.type[threadIdx.x] = in ? m.station_type[threadIdx.x] : 0;
    }
    __syncthreads();
}

struct Session { int arr, dep, est; double energy; };

__device__ __forceinline__ Session load_session(const sg_ev_model& m, int day, int k, int i) {
    Session s;
    if (k >= m.max_sessions) { s.arr = s.dep = s.est = kEmpty; s.energy = 0.0; return s; }
    const uint4* recs = reinterpret_cast<const uint4*>(m.sessions);
    const uint4 r = __ldg(recs + ((long long)day * m.max_sessions + k) * m.n_stations + i);
    s.arr = (int)(short)(r.x & 0xffffu);
    s.dep = (int)(short)(r.x >> 16);
    s.est = (int)(short)(r.y & 0xffffu);
    s.energy = __longlong_as_double((long long)(((unsigned long long)r.w << 32) | r.z));
    return s;
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

// class sums of small non-negative integers (< 2^26 / n per station): one REDUX per class
__device__ __forceinline__ void class_sums_u32(const unsigned (&v)[2], const int (&cls)[2], int C, double (&S)[kMC],
                                               double scale) {
#pragma unroll
    for (int c = 0; c < kMC; ++c) {
        if (c < C) S[c] = (double)__reduce_add_sync(kFull, (cls[0] == c ? v[0] : 0u) + (cls[1] == c ? v[1] : 0u)) * scale;
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

// Conservative test (fast path): fixed-point class sums through REDUX; true only when every
// constraint holds with a margin that covers the quantisation error.
__device__ __forceinline__ bool network_certainly_feasible(const EvShared& sh, int C, int J, int lane,
                                                           const double (&x)[2], const int (&cls)[2]) {
    const unsigned xi[2] = {__double2uint_rn(x[0] * kFixScale), __double2uint_rn(x[1] * kFixScale)};
    double S[kMC];
    class_sums_u32(xi, cls, C, S, kFixInv);
    bool bad = false;
    if (lane < J) {
        double wr = 0, wi = 0;
#pragma unroll
        for (int c = 0; c < kMC; ++c) { wr = fma(sh.con_re[lane * kMC + c], S[c], wr); wi = fma(sh.con_im[lane * kMC + c], S[c], wi); }
        const double r = sh.lim[lane] * (1.0 / kActionScale) - kFeasMargin;
        bad = !(r > 0.0 && fma(wr, wr, wi * wi) <= r * r);
    }
    return !__any_sync(kFull, bad);
}

// =================================================================================================
// One env, one step.  MODE 0: no projection.  MODE 1: projection, fast path inline, otherwise
// defer to the work list.  MODE 2: projection with the solver inline (work-list kernel).
// =================================================================================================
template <int MODE>
__device__ __forceinline__ void ev_env_step(const sg_ev_model& m, const sg_ev_state& s, const EvShared& sh,
                                            const long long b, const int lane, const void* __restrict__ action,
                                            const int flags, const int bins, float* __restrict__ obs,
                                            double* __restrict__ reward, uint8_t* __restrict__ done,
                                            uint8_t* __restrict__ pilots, float* __restrict__ final_obs,
                                            uint8_t* __restrict__ status, int32_t* __restrict__ todo,
                                            StatsAcc& acc) {
    const int n = m.n_stations, kf = m.moer_steps, OBS = 2 * n + kf + 2;
    const int C = m.n_classes, J = m.n_constraints;
    const int4 clk = *reinterpret_cast<const int4*>(s.clock + b * 4);
    const int t = clk.x;
    if (t >= kEpisode) {            // finished episode that was not reset: no-op, stays terminal
        if (lane == 0) { reward[b] = 0.0; done[b] = 1; if (status) status[b] = 0; }
        return;
    }
    int day = clk.y, md = clk.z;
    const int i0 = 2 * lane;
    const bool valid[2] = {i0 < n, i0 + 1 < n};
    const bool vec = ((n & 1) == 0);     // rows are 8/16-byte aligned when n is even

    // ---- coalesced row loads -------------------------------------------------------------------
    float af[2] = {0.f, 0.f};
    if ((flags & 1) == SG_EV_ACTION_I64) {
        const long long* ap = reinterpret_cast<const long long*>(action) + b * n;
#pragma unroll
        for (int p = 0; p < 2; ++p)
            if (valid[p]) af[p] = __fdiv_rn((float)ap[i0 + p], (float)(bins - 1));   // wrappers.py:45
    } else {
        const float* ap = reinterpret_cast<const float*>(action) + b * n;
        if (vec) {
            if (valid[0]) { const float2 v = ldg_stream_f2(reinterpret_cast<const float2*>(ap + i0)); af[0] = v.x; af[1] = v.y; }
        } else {
#pragma unroll
            for (int p = 0; p < 2; ++p) if (valid[p]) af[p] = __ldg(ap + i0 + p);
        }
    }
    double del[2] = {0.0, 0.0};
    int k[2] = {0, 0};
    {
        const double* dp = s.delivered + b * n;
        const uint8_t* kp = s.sess_idx + b * n;
        if (vec) {
            if (valid[0]) {
                const double2 v = *reinterpret_cast<const double2*>(dp + i0);
                del[0] = v.x; del[1] = v.y;
                const uchar2 kk = *reinterpret_cast<const uchar2*>(kp + i0);
                k[0] = kk.x; k[1] = kk.y;
            }
        } else {
#pragma unroll
            for (int p = 0; p < 2; ++p) if (valid[p]) { del[p] = dp[i0 + p]; k[p] = kp[i0 + p]; }
        }
    }
    int cls[2], type[2];
    Session ses[2];
    bool plugged[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        cls[p] = valid[p] ? sh.cls[i0 + p] : 255;
        type[p] = valid[p] ? sh.type[i0 + p] : 0;
        if (valid[p]) ses[p] = load_session(m, day, k[p], i0 + p);
        else { ses[p].arr = ses[p].dep = ses[p].est = kEmpty; ses[p].energy = 0.0; }
        // plugged in during period t: its Plugin event (timestamp <= t) has been processed (events
        // are only drained after the first charge, so nothing is plugged at t = 0) and its Unplug
        // event (timestamp = departure) has not.
        plugged[p] = valid[p] && t >= 1 && ses[p].arr <= t && t < ses[p].dep;
    }

    // ---- action -> normalised pilot x (env.py:363-366) -------------------------------------------
    double x[2];
    int st = 0;
    if constexpr (MODE == 0) {
        x[0] = (double)af[0]; x[1] = (double)af[1];
    } else {
        double a[2], u[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const double rem = ses[p].energy - del[p];
            const float dem = (plugged[p] && rem > kFullEps) ? (float)rem : 0.f;     // obs 'demands', env.py:388
            a[p] = (double)af[p];
            u[p] = upper_bound(dem);
            x[p] = valid[p] ? fmin(fmax(a[p], 0.0), u[p]) : 0.0;
        }
        if constexpr (MODE == 1) {
            if (!network_certainly_feasible(sh, C, J, lane, x, cls)) {
                if (lane == 0) todo[2 + atomicAdd(todo, 1)] = (int)b;
                return;                                     // the work-list kernel redoes this env
            }
        } else {
            if (!network_feasible(sh, C, J, x, cls)) {
                bool cap = false;
                if (!(simple_project(sh, C, J, a, u, cls, valid, x) && network_feasible_tol(sh, C, J, x, cls))) {
#pragma unroll
                    for (int p = 0; p < 2; ++p) x[p] = valid[p] ? fmin(fmax(a[p], 0.0), u[p]) : 0.0;
                    ipm_dispatch(sh, C, J, a, u, cls, valid, x, &cap);
                }
                st = SG_EV_STATUS_PROJECTED | (cap ? SG_EV_STATUS_MAXITER : 0);
            }
        }
    }

    // ---- schedule (env.py:368-378), charging (acnsim update_pilots), rates ------------------------
    double pilot[2], rate_sum = 0.0;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        pilot[p] = valid[p] ? to_pilot(x[p], type[p]) : 0.0;
        if (plugged[p]) {
            // Linear2StageBattery._charge_stepwise (capacity 100 kWh, max_power 100 kW, transition 0.8)
            const double charge = fmax(0.0, kCapacity - ses[p].energy) + del[p];
            const double rate_to_full = div_c(kCapacity - charge, kPeriodH, kInvPeriodH);
            const double soc = div_c(charge, kCapacity, kInv100);
            const double p_pilot = div_c(pilot[p] * kVoltage, 1000.0, kInv1000);
            const double cap2 = soc < kTransitionSoc ? kMaxPower
                                                     : div_c(1.0 - soc, kOneMinusTrans, kInvOMT) * kMaxPower;
            const double power = fmin(fmin(p_pilot, cap2), rate_to_full);
            const double rate = div_c(power * 1000.0, kVoltage, kInv208);                 // A
            del[p] += div_c(rate * kVoltage, 1000.0, kInv1000) * kPeriodH;               // EV.charge
            rate_sum += rate;
        }
    }
    const double total_rate = warp_sum(rate_sum);

    // ---- excess current from the rounded schedule (env.py:449-452) ---------------------------------
    double excess = 0.0;
    {
        const unsigned pi[2] = {(unsigned)pilot[0], (unsigned)pilot[1]};
        double S[kMC];
        class_sums_pilots(pi, cls, C, S);
        double sq = 0.0, lim = 0.0;
        if (lane < J) {
            double wr = 0, wi = 0;
#pragma unroll
            for (int c = 0; c < kMC; ++c) { wr = fma(sh.con_re[lane * kMC + c], S[c], wr); wi = fma(sh.con_im[lane * kMC + c], S[c], wi); }
            sq = fma(wr, wr, wi * wi);
            lim = sh.lim[lane];
        }
        const bool over = lane < J && sq > lim * lim;
        if (__any_sync(kFull, over)) excess = warp_sum(over ? fmax(0.0, sqrt(sq) - lim) : 0.0);
    }

    // ---- advance time, drain events with timestamp <= t+1: Unplug (0) before Plugin (10) ------------
    const int tn = t + 1;
    float dem_o[2], est_o[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        while (valid[p] && ses[p].dep <= tn && ses[p].arr != kEmpty) {   // session over (or zero-length)
            k[p] += 1; del[p] = 0.0;
            ses[p] = load_session(m, day, k[p], i0 + p);
        }
        const bool pl = valid[p] && ses[p].arr <= tn && tn < ses[p].dep;
        const double rem = ses[p].energy - del[p];
        const bool active = pl && rem > kFullEps;                         // Interface.active_sessions()
        dem_o[p] = active ? (float)rem : 0.f;
        est_o[p] = active ? (float)(ses[p].est - tn) : 0.f;
    }
    const bool dn = tn >= kEpisode;

    // ---- reward (env.py:444-462) ---------------------------------------------------------------------
    const float* mrow = m.moer + ((long long)md * SG_EV_MOER_ROWS + tn) * SG_EV_MOER_COLS;
    const double moer0 = (double)__ldg(mrow);
    const double profit = kProfitFactor * total_rate;
    const double excess_charge = excess * kViolationFactor;
    const double carbon = kCarbonFactor * total_rate * moer0;
    const double rew = profit - carbon - excess_charge;

    const bool autoreset = dn && (flags & SG_EV_FLAG_AUTORESET);

    // ---- observation row: demands | est_departures | forecasted_moer | prev_moer | timestep ----------
    auto write_obs = [&](float* orow, const float* row, const int tstep) {
        if (vec) {
            if (valid[0]) {
                stg_stream_f2(reinterpret_cast<float2*>(orow + i0), make_float2(dem_o[0], dem_o[1]));
                stg_stream_f2(reinterpret_cast<float2*>(orow + n + i0), make_float2(est_o[0], est_o[1]));
            }
        } else {
#pragma unroll
            for (int p = 0; p < 2; ++p) if (valid[p]) { orow[i0 + p] = dem_o[p]; orow[n + i0 + p] = est_o[p]; }
        }
        for (int e = lane; e < kf + 2; e += 32) {
            float v;
            if (e < kf) v = __ldg(row + 1 + e);
            else if (e == kf) v = __ldg(row);
            else v = (float)((double)tstep / (double)kEpisode);     // env.py:392
            orow[2 * n + e] = v;
        }
    };
    if (!autoreset) {
        write_obs(obs + b * OBS, mrow, tn);
    } else {
        if (final_obs != nullptr) write_obs(final_obs + b * OBS, mrow, tn);   // terminal observation
        // reset observation of the next episode (env.py:293-338): next day, empty garage, t = 0
        day = (day + 1) % m.n_days;                      // sequential days, event_generation.py:284-291
        md = (md + 1) % m.n_moer_days;
        dem_o[0] = dem_o[1] = est_o[0] = est_o[1] = 0.f;
        k[0] = k[1] = 0; del[0] = del[1] = 0.0;
        write_obs(obs + b * OBS, m.moer + ((long long)md * SG_EV_MOER_ROWS) * SG_EV_MOER_COLS, 0);
    }

    // ---- state write-back ------------------------------------------------------------------------------
    {
        double* dp = s.delivered + b * n;
        uint8_t* kp = s.sess_idx + b * n;
        if (vec) {
            if (valid[0]) {
                *reinterpret_cast<double2*>(dp + i0) = make_double2(del[0], del[1]);
                *reinterpret_cast<uchar2*>(kp + i0) = make_uchar2((unsigned char)k[0], (unsigned char)k[1]);
            }
        } else {
#pragma unroll
            for (int p = 0; p < 2; ++p) if (valid[p]) { dp[i0 + p] = del[p]; kp[i0 + p] = (uint8_t)k[p]; }
        }
        if (pilots != nullptr) {
            if (vec) {
                if (valid[0]) *reinterpret_cast<uchar2*>(pilots + b * n + i0) =
                    make_uchar2((unsigned char)pilot[0], (unsigned char)pilot[1]);
            } else {
#pragma unroll
                for (int p = 0; p < 2; ++p) if (valid[p]) pilots[b * n + i0 + p] = (uint8_t)pilot[p];
            }
        }
    }
    // occupancy mask (bit i: station i has a plugged-in session at the new time) and arrival cursor
    // of the event-driven kernel — derived state every kernel keeps consistent
    unsigned occ_lo, occ_hi;
    {
        bool pln[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) pln[p] = !autoreset && valid[p] && ses[p].arr <= tn && tn < ses[p].dep;
        const unsigned long long mine = ((unsigned long long)pln[0] << (i0 & 63)) | ((unsigned long long)pln[1] << ((i0 + 1) & 63));
        occ_lo = __reduce_or_sync(kFull, (unsigned)(mine & 0xffffffffull));
        occ_hi = __reduce_or_sync(kFull, (unsigned)(mine >> 32));
    }
    if (lane == 0) {
        // cumulative reward breakdown (env.py:460-462); cleared on the first step of an episode so
        // that after an auto-reset the finished episode's totals stay readable until the next step
        double* bd = s.breakdown + b * 3;
        if (t == 0) { bd[0] = profit; bd[1] = carbon; bd[2] = excess_charge; }
        else { bd[0] += profit; bd[1] += carbon; bd[2] += excess_charge; }
        int nidx = autoreset ? 0 : clk.w;
        if (!autoreset && m.arrivals != nullptr) {
            const unsigned* arow = m.arrivals + (long long)clk.y * (m.max_day_sessions + 1);
            while ((int)(__ldg(arow + nidx) & 0x7fffu) <= tn) ++nidx;
        }
        if (s.occ != nullptr) s.occ[b] = ((unsigned long long)occ_hi << 32) | occ_lo;
        *reinterpret_cast<int4*>(s.clock + b * 4) = make_int4(autoreset ? 0 : tn, day, md, nidx);
        reward[b] = rew;
        done[b] = dn ? 1 : 0;
        if (status != nullptr) status[b] = (uint8_t)st;
        acc.add(rew, dn, profit, carbon, excess_charge);
        acc.n_proj += (st & SG_EV_STATUS_PROJECTED) ? 1.0 : 0.0;
        acc.n_maxiter += (st & SG_EV_STATUS_MAXITER) ? 1.0 : 0.0;
    }
}

// =================================================================================================
// Tile kernel (dense main kernel: default without projection, optional with it).  The warp-per-env kernels above spend ~800 warp instructions
// per env because every per-env scalar operation is replicated over 32 lanes and both stations of
// a lane run the f64 battery path although ~85 % of the stations are empty.  Here one warp owns a
// TILE of 32 consecutive envs and one THREAD advances one env:
//   phase 0  the tile's action rows and sess_idx bytes are copied global->shared with coalesced
//            8/16-byte accesses (the only place the env-major layout is touched), clock rows are
//            one coalesced 16-byte load per lane;
//   phase 1  each thread scans its env's stations (cheap: one byte per station; stations known to
//            have no current-or-next session carry the hint sess_idx = 255) and appends the
//            stations whose session has started or starts now to a compact list in shared memory;
//            the f64 battery arithmetic then runs over the LISTS (trip count ~ the largest list,
//            mostly-active lanes); class sums are per-thread accumulators — no shuffles at all;
//            `delivered` is read/written only for listed stations (sparse 8-byte accesses);
//   phase 2  observation rows are assembled from the shared tiles and stored row by row with
//            8-byte vectors; sess_idx / pilots tiles go back with coalesced copies.
// Envs whose clipped action is not inside the network constraints are queued for the solver
// kernel exactly as before and leave no trace in global memory.
// =================================================================================================
constexpr int kTileE = 32;

struct TileLayout {
    int RS, off_est, off_k, off_pil, off_lst, off_S, off_P, off_rate, off_energy, off_con, bytes;
    __host__ __device__ explicit TileLayout(int n) {
        RS = n | 1;                                        // odd row stride: conflict-free column walks
        int o = kTileE * RS * 4;                           // act/dem tile  f32 [32][RS]
        off_est = o; o += kTileE * RS * 2;                 // est tile      i16 [32][RS]
        o = (o + 15) & ~15;
        off_k = o; o += kTileE * n;                        // sess_idx tile u8  [32*n] (global layout)
        o = (o + 15) & ~15;
        off_pil = o; o += kTileE * n;                      // pilots tile   u8  [32*n]
        o = (o + 15) & ~15;
        off_lst = o; o += kTileE * n;                      // station lists u8  [32][n]
        o = (o + 15) & ~15;
        off_S = o; o += kMC * kTileE * 8;                  // class sums of x        f64 [8][32]
        off_P = o; o += kMC * kTileE * 4;                  // class sums of pilots   i32 [8][32]
        o = (o + 15) & ~15;
        off_rate = o; o += 40 * 8;
        off_energy = o; o += 40 * 8;
        off_con = o; o += (2 * SG_EV_MAX_CONSTRAINTS * kMC + SG_EV_MAX_CONSTRAINTS) * 8 + 2 * SG_EV_MAX_STATIONS;
        o = (o + 15) & ~15;
        bytes = o;
    }
};

template <int MODE>
__global__ void __launch_bounds__(kTileE)
ev_step_tile_kernel(const sg_ev_model m, const sg_ev_state s, const long long B, const void* __restrict__ action,
                    const int flags, const int bins, float* __restrict__ obs, double* __restrict__ reward,
                    uint8_t* __restrict__ done, uint8_t* __restrict__ pilots, float* __restrict__ final_obs,
                    uint8_t* __restrict__ status, double* __restrict__ stats, int32_t* __restrict__ todo) {
    extern __shared__ __align__(16) unsigned char tsm[];
    const int n = m.n_stations, kf = m.moer_steps, OBS = 2 * n + kf + 2, C = m.n_classes, J = m.n_constraints;
    const int K = m.max_sessions;
    const TileLayout L(n);
    const int RS = L.RS;
    float* tact = reinterpret_cast<float*>(tsm);
    short* test = reinterpret_cast<short*>(tsm + L.off_est);
    uint8_t* tk = tsm + L.off_k;
    uint8_t* tpil = tsm + L.off_pil;
    uint8_t* tlst = tsm + L.off_lst;
    double* sS = reinterpret_cast<double*>(tsm + L.off_S);
    int* sP = reinterpret_cast<int*>(tsm + L.off_P);
    double* srate = reinterpret_cast<double*>(tsm + L.off_rate);
    double* senergy = reinterpret_cast<double*>(tsm + L.off_energy);
    double* scre = reinterpret_cast<double*>(tsm + L.off_con);          // [16][8]
    double* scim = scre + SG_EV_MAX_CONSTRAINTS * kMC;                  // [16][8]
    double* slim = scim + SG_EV_MAX_CONSTRAINTS * kMC;                  // [16]
    uint8_t* scls = reinterpret_cast<uint8_t*>(slim + SG_EV_MAX_CONSTRAINTS);   // [64]
    uint8_t* stype = scls + SG_EV_MAX_STATIONS;                          // [64]
    const int lane = threadIdx.x;
    for (int pI = lane; pI <= 32; pI += 32) {
        const double pp = div_c((double)pI * kVoltage, 1000.0, kInv1000);
        const double rate = div_c(pp * 1000.0, kVoltage, kInv208);
        srate[pI] = rate;
        senergy[pI] = div_c(rate * kVoltage, 1000.0, kInv1000) * kPeriodH;
    }
    for (int i = lane; i < SG_EV_MAX_CONSTRAINTS * kMC; i += 32) { scre[i] = m.con_re[i]; scim[i] = m.con_im[i]; }
    if (lane < SG_EV_MAX_CONSTRAINTS) slim[lane] = m.magnitudes[lane];
    for (int i = lane; i < SG_EV_MAX_STATIONS; i += 32) {
        scls[i] = i < n ? m.station_class[i] : 0;
        stype[i] = i < n ? m.station_type[i] : 0;
    }
    __syncwarp();
    const bool i64 = (flags & 1) == SG_EV_ACTION_I64;
    const uint4* recs = reinterpret_cast<const uint4*>(m.sessions);
    const long long ntiles = (B + kTileE - 1) / kTileE;
    StatsAcc acc;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long b0 = tile * kTileE;
        const int nv = (int)min((long long)kTileE, B - b0);
        const long long b = b0 + lane;
        const bool env = lane < nv;

        // ---------------- phase 0: coalesced tile loads (8 rows of loads in flight) -----------------
        int4 clk = make_int4(kEpisode, 0, 0, 0);
        if (env) clk = *reinterpret_cast<const int4*>(s.clock + b * 4);
        {
            const uint8_t* gk = s.sess_idx + b0 * n;
            const int nbytes = nv * n, nvec = nbytes >> 4;
            for (int c = lane; c < nvec; c += 32) reinterpret_cast<uint4*>(tk)[c] = reinterpret_cast<const uint4*>(gk)[c];
            for (int c = (nvec << 4) + lane; c < nbytes; c += 32) tk[c] = gk[c];
        }
        const bool col = 2 * lane < n;
        for (int r0 = 0; r0 < nv; r0 += 8) {
            float2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                v[u] = make_float2(0.f, 0.f);
                if (col && r0 + u < nv) {
                    if (i64) {
                        const long long* ap = reinterpret_cast<const long long*>(action) + (b0 + r0 + u) * n + 2 * lane;
                        v[u].x = __fdiv_rn((float)ap[0], (float)(bins - 1));            // wrappers.py:45
                        v[u].y = __fdiv_rn((float)ap[1], (float)(bins - 1));
                    } else {
                        v[u] = ldg_stream_f2(reinterpret_cast<const float2*>(reinterpret_cast<const float*>(action) + (b0 + r0 + u) * n) + lane);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (col && r0 + u < nv) { tact[(r0 + u) * RS + 2 * lane] = v[u].x; tact[(r0 + u) * RS + 2 * lane + 1] = v[u].y; }
        }
        __syncwarp();

        // ---------------- phase 1: one thread = one env ------------------------------------------
        const int t = clk.x, tn = t + 1;
        int day = clk.y, md = clk.z;
        const bool live = env && t < kEpisode;
        bool commit = live;
        const int sbase = day * K * n;
        for (int c = 0; c < C; ++c) { sS[c * kTileE + lane] = 0.0; sP[c * kTileE + lane] = 0; }
        int nl = 0;
        float* myact = tact + lane * RS;
        short* myest = test + lane * RS;
        uint8_t* myk = tk + lane * n;
        uint8_t* mypil = tpil + lane * n;
        uint8_t* mylst = tlst + lane * n;
        // moer0 of the row the reward needs: issue early
        float moer0f = 0.f;
        if (live) moer0f = __ldg(m.moer + ((long long)md * SG_EV_MOER_ROWS + tn) * SG_EV_MOER_COLS);
        if (live) {
            constexpr int U = 9;
            for (int i0 = 0; i0 < n; i0 += U) {
                int kk[U], arrv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) kk[u] = (i0 + u < n) ? myk[i0 + u] : 255;
#pragma unroll
                for (int u = 0; u < U; ++u) {                         // U independent L2 look-ups in flight
                    arrv[u] = kEmpty;
                    if (kk[u] < K) arrv[u] = (int)(short)(__ldg(reinterpret_cast<const unsigned*>(recs + (sbase + kk[u] * n + i0 + u))) & 0xffffu);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 + u;
                    if (i < n) {
                        bool listed = false;
                        if (kk[u] != 255) {
                            if (arrv[u] == kEmpty) myk[i] = 255;     // nothing (more) at this station today
                            else if (arrv[u] <= tn) { mylst[nl++] = (uint8_t)i; listed = true; }
                        }
                        myest[i] = 0;
                        if constexpr (MODE == 0) {
                            // every EVSE receives a pilot, occupied or not (it counts in the constraint currents)
                            const float amps = myact[i] * 32.f;
                            const float pf = stype[i] == 0 ? (amps >= 6.f ? rintf(amps) : 0.f) : rintf(amps * 0.125f) * 8.f;
                            const int pi = (int)pf;
                            mypil[i] = (uint8_t)pi;
                            sP[scls[i] * kTileE + lane] += pi;
                        } else {
                            mypil[i] = 0;
                        }
                        if (!listed) myact[i] = 0.f;                  // 'demands' of an empty station
                    }
                }
            }
        }
        constexpr int V = 4;                                          // listed stations processed per batch
        struct Rec { int i, k, arr, dep, est; double energy, del; bool on, plugged; };
        auto load_batch = [&](int e0, bool enable, Rec (&q)[V]) {
#pragma unroll
            for (int u = 0; u < V; ++u) {
                q[u].on = enable && e0 + u < nl;
                q[u].i = q[u].on ? mylst[e0 + u] : 0;
                q[u].k = q[u].on ? myk[q[u].i] : 255;
            }
            uint4 raw[V];
#pragma unroll
            for (int u = 0; u < V; ++u) {
                raw[u] = make_uint4(0x7fff7fffu, 0x7fffu, 0u, 0u);
                if (q[u].on && q[u].k < K) raw[u] = __ldg(recs + (sbase + q[u].k * n + q[u].i));
            }
#pragma unroll
            for (int u = 0; u < V; ++u) {
                q[u].arr = (int)(short)(raw[u].x & 0xffffu);
                q[u].dep = (int)(short)(raw[u].x >> 16);
                q[u].est = (int)(short)(raw[u].y & 0xffffu);
                q[u].energy = __hiloint2double((int)raw[u].w, (int)raw[u].z);
                q[u].plugged = q[u].on && t >= 1 && q[u].arr <= t && t < q[u].dep;
            }
#pragma unroll
            for (int u = 0; u < V; ++u) q[u].del = q[u].plugged ? s.delivered[b * n + q[u].i] : 0.0;
        };
        // pass A (projection only): clipped action and its class sums -> feasibility
        if constexpr (MODE == 1) {
            const int nmax = __reduce_max_sync(kFull, live ? nl : 0);
            for (int e0 = 0; e0 < nmax; e0 += V) {
                Rec q[V];
                load_batch(e0, live, q);
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    if (q[u].plugged) {
                        const double rem = q[u].energy - q[u].del;
                        const float dem = rem > kFullEps ? (float)rem : 0.f;
                        const double x = fmin(fmax((double)myact[q[u].i], 0.0), upper_bound(dem));
                        sS[scls[q[u].i] * kTileE + lane] += x;
                    }
                }
            }
            if (live) {
                bool ok = true;
                for (int j = 0; j < J; ++j) {
                    double wr = 0, wi = 0;
                    for (int c = 0; c < C; ++c) {
                        const double S = sS[c * kTileE + lane];
                        wr = fma(scre[j * kMC + c], S, wr);
                        wi = fma(scim[j * kMC + c], S, wi);
                    }
                    const double r = slim[j] * (1.0 / kActionScale);
                    ok = ok && (fma(wr, wr, wi * wi) <= r * r);
                }
                if (!ok) {
                    commit = false;
                    todo[2 + atomicAdd(todo, 1)] = (int)b;        // the work-list kernel redoes this env
                }
            }
        }
        // pass B: pilots, charging, events, new demands / departures for the listed stations
        double rate_sum = 0.0;
        unsigned long long occ_new = 0ull;
        {
            const int nmax = __reduce_max_sync(kFull, commit ? nl : 0);
            for (int e0 = 0; e0 < nmax; e0 += V) {
                Rec q[V];
                load_batch(e0, commit, q);
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    if (!q[u].on) continue;
                    const int i = q[u].i;
                    int k = q[u].k, arr = q[u].arr, dep = q[u].dep, est = q[u].est;
                    double energy = q[u].energy, del = q[u].del;
                    const bool plugged = q[u].plugged;
                    if (plugged) {
                        int pi;
                        double pilot;
                        if constexpr (MODE == 1) {
                            const double rem = energy - del;
                            const float dem = rem > kFullEps ? (float)rem : 0.f;
                            const double x = fmin(fmax((double)myact[i], 0.0), upper_bound(dem));
                            pilot = to_pilot(x, stype[i]);
                            pi = __double2int_rn(pilot);
                            mypil[i] = (uint8_t)pi;
                            sP[scls[i] * kTileE + lane] += pi;
                        } else {
                            pi = mypil[i];
                            pilot = (double)pi;
                        }
                        const double charge = fmax(0.0, kCapacity - energy) + del;
                        if (kCapacity - charge > 2.0) {              // neither battery limit can bind
                            del += senergy[pi];
                            rate_sum += srate[pi];
                        } else {
                            const double rate_to_full = div_c(kCapacity - charge, kPeriodH, kInvPeriodH);
                            const double soc = div_c(charge, kCapacity, kInv100);
                            const double p_pilot = div_c(pilot * kVoltage, 1000.0, kInv1000);
                            const double cap2 = soc < kTransitionSoc ? kMaxPower
                                                                     : div_c(1.0 - soc, kOneMinusTrans, kInvOMT) * kMaxPower;
                            const double power = fmin(fmin(p_pilot, cap2), rate_to_full);
                            const double rate = div_c(power * 1000.0, kVoltage, kInv208);
                            del += div_c(rate * kVoltage, 1000.0, kInv1000) * kPeriodH;
                            rate_sum += rate;
                        }
                    }
                    bool moved = false;
                    while (dep <= tn && arr != kEmpty) {              // session over (or zero-length): next slot
                        k += 1; del = 0.0; moved = true;
                        arr = dep = est = kEmpty; energy = 0.0;
                        if (k < K) {
                            const uint4 r = __ldg(recs + (sbase + k * n + i));
                            arr = (int)(short)(r.x & 0xffffu); dep = (int)(short)(r.x >> 16);
                            est = (int)(short)(r.y & 0xffffu); energy = __hiloint2double((int)r.w, (int)r.z);
                        }
                    }
                    if (moved) myk[i] = (arr == kEmpty) ? 255 : (uint8_t)k;
                    if (plugged || moved) s.delivered[b * n + i] = del;
                    const double rem = energy - del;
                    const bool pln = arr <= tn && tn < dep;
                    const bool active = pln && rem > kFullEps;                     // Interface.active_sessions()
                    if (pln) occ_new |= 1ull << i;
                    myact[i] = active ? (float)rem : 0.f;
                    myest[i] = active ? (short)(est - tn) : (short)0;
                }
            }
        }
        // ---------------- per-env epilogue: reward, clock, statistics -----------------------------
        const bool dn = tn >= kEpisode;
        const bool autoreset = commit && dn && (flags & SG_EV_FLAG_AUTORESET);
        int md_out = md, t_out = tn;
        if (commit) {
            double excess = 0.0;
            for (int j = 0; j < J; ++j) {
                double wr = 0, wi = 0;
                for (int c = 0; c < C; ++c) {
                    const double S = (double)sP[c * kTileE + lane];
                    wr = fma(scre[j * kMC + c], S, wr);
                    wi = fma(scim[j * kMC + c], S, wi);
                }
                const double lim = slim[j], sq = fma(wr, wr, wi * wi);
                if (sq > lim * lim) excess += fmax(0.0, sqrt(sq) - lim);
            }
            const double moer0 = (double)moer0f;
            const double profit = kProfitFactor * rate_sum;
            const double excess_charge = excess * kViolationFactor;
            const double carbon = kCarbonFactor * rate_sum * moer0;
            const double rew = profit - carbon - excess_charge;
            double* bd = s.breakdown + b * 3;
            if (t == 0) { bd[0] = profit; bd[1] = carbon; bd[2] = excess_charge; }
            else { bd[0] += profit; bd[1] += carbon; bd[2] += excess_charge; }
            int nidx = clk.w;
            if (m.arrivals != nullptr) {
                const unsigned* arow = m.arrivals + (long long)day * (m.max_day_sessions + 1);
                while ((int)(__ldg(arow + nidx) & 0x7fffu) <= tn) ++nidx;
            }
            if (autoreset) {
                day = (day + 1) % m.n_days;                           // sequential days, event_generation.py:284-291
                md_out = (md + 1) % m.n_moer_days;
                t_out = 0; nidx = 0; occ_new = 0ull;
                for (int i = 0; i < n; ++i) myk[i] = 0;               // delivered is already 0: every session ended
            }
            if (s.occ != nullptr) s.occ[b] = occ_new;
            *reinterpret_cast<int4*>(s.clock + b * 4) = make_int4(t_out, day, md_out, nidx);
            reward[b] = rew;
            done[b] = dn ? 1 : 0;
            if (status != nullptr) status[b] = 0;
            if (stats != nullptr) acc.add(rew, dn, profit, carbon, excess_charge);
        } else if (env && t >= kEpisode) {                            // finished episode, not reset: stays terminal
            reward[b] = 0.0; done[b] = 1;
            if (status != nullptr) status[b] = 0;
        }
        __syncwarp();

        // ---------------- phase 2: coalesced write-back (tails of 8 rows in flight) ----------------
        const unsigned commit_mask = __ballot_sync(kFull, commit);
        {
            uint8_t* gk = s.sess_idx + b0 * n;
            const int nbytes = nv * n, nvec = nbytes >> 4;
            for (int c = lane; c < nvec; c += 32) reinterpret_cast<uint4*>(gk)[c] = reinterpret_cast<const uint4*>(tk)[c];
            for (int c = (nvec << 4) + lane; c < nbytes; c += 32) gk[c] = tk[c];
        }
        const float ts_new = (float)((double)tn / (double)kEpisode);            // env.py:392 (same t for a lock-step tile row)
        for (int r0 = 0; r0 < nv; r0 += 8) {
            float tail0[8], tail1[8];
            int arr8 = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = min(r0 + u, kTileE - 1);
                const int md_r = __shfl_sync(kFull, md, r), tn_r = __shfl_sync(kFull, tn, r);
                arr8 |= __shfl_sync(kFull, (int)autoreset, r) << u;
                tail0[u] = tail1[u] = 0.f;
                if (r0 + u < nv && ((commit_mask >> r) & 1u)) {
                    const float* row = m.moer + ((long long)md_r * SG_EV_MOER_ROWS + tn_r) * SG_EV_MOER_COLS;
                    if (lane < kf) tail0[u] = __ldg(row + 1 + lane); else if (lane == kf) tail0[u] = __ldg(row);
                    const int e = lane + 32;
                    if (e < kf) tail1[u] = __ldg(row + 1 + e); else if (e == kf) tail1[u] = __ldg(row);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + u;
                if (r >= nv || !((commit_mask >> r) & 1u)) continue;
                const bool ar = (arr8 >> u) & 1;
                const float ts_r = __shfl_sync(kFull, ts_new, r);
                float* orow = (ar && final_obs != nullptr) ? final_obs + (b0 + r) * OBS : obs + (b0 + r) * OBS;
                if (!(ar && final_obs == nullptr)) {                  // the step's own observation (terminal one if auto-reset)
                    if (col) {
                        stg_stream_f2(reinterpret_cast<float2*>(orow) + lane,
                                      make_float2(tact[r * RS + 2 * lane], tact[r * RS + 2 * lane + 1]));
                        stg_stream_f2(reinterpret_cast<float2*>(orow + n) + lane,
                                      make_float2((float)test[r * RS + 2 * lane], (float)test[r * RS + 2 * lane + 1]));
                    }
                    if (lane <= kf) orow[2 * n + lane] = tail0[u]; else if (lane == kf + 1) orow[2 * n + lane] = ts_r;
                    if (lane + 32 <= kf) orow[2 * n + lane + 32] = tail1[u]; else if (lane + 32 == kf + 1) orow[2 * n + lane + 32] = ts_r;
                }
                if (ar) {                                             // reset observation of the next episode
                    const int mdo_r = __shfl_sync(kFull, md_out, r);
                    const float* row = m.moer + ((long long)mdo_r * SG_EV_MOER_ROWS) * SG_EV_MOER_COLS;
                    float* o2 = obs + (b0 + r) * OBS;
                    if (col) {
                        stg_stream_f2(reinterpret_cast<float2*>(o2) + lane, make_float2(0.f, 0.f));
                        stg_stream_f2(reinterpret_cast<float2*>(o2 + n) + lane, make_float2(0.f, 0.f));
                    }
#pragma unroll
                    for (int pass = 0; pass < 2; ++pass) {
                        const int e = lane + 32 * pass;
                        if (e < kf + 2) o2[2 * n + e] = e < kf ? __ldg(row + 1 + e) : (e == kf ? __ldg(row) : 0.f);
                    }
                }
                if (pilots != nullptr && col)
                    reinterpret_cast<uchar2*>(pilots + (b0 + r) * n)[lane] = make_uchar2(tpil[r * n + 2 * lane], tpil[r * n + 2 * lane + 1]);
            }
        }
        __syncwarp();
    }
    if (stats != nullptr) {
        // one warp per CTA: plain warp reduction + one set of atomics
        double v[11] = {acc.count, acc.sum, acc.sumsq, acc.mn, acc.mx, acc.n_done, acc.n_proj, acc.n_maxiter, acc.c0, acc.c1, acc.c2};
#pragma unroll
        for (int i = 0; i < 11; ++i) {
            const double r = (i == 3) ? warp_min(v[i]) : (i == 4) ? warp_max(v[i]) : warp_sum(v[i]);
            if (lane == 0) {
                if (i == 3) atomic_min_double(stats + 3, r);
                else if (i == 4) atomic_max_double(stats + 4, r);
                else if (r != 0.0) atomicAdd(stats + i, r);
            }
        }
    }
}

// =================================================================================================
// Event-driven kernel (default main kernel when the action projection is on).
//
// With the projection on, a station without a plugged-in EV has upper bound 0 (env.py:188-189): its
// pilot is 0, it delivers nothing and its observation entries are 0.  So one step only has to
// touch the OCCUPIED stations (~15 % of the 54).  Per env the engine keeps a 64-bit occupancy mask
// and a cursor into the day's arrival-sorted session list; one thread advances one env:
//   * walk the set bits of the mask: sparse loads of sess_idx / action / delivered and the session
//     record, battery arithmetic, sparse stores; no shared-memory tiles -> high occupancy;
//   * departures clear bits, arrivals (cursor) set them;
//   * the observation rows are zero-filled with coalesced vector stores by the whole warp and the
//     few non-zero entries are patched by the owning thread; the MOER tails are copied row by row.
// Envs whose clipped action violates the network are queued for the solver kernel untouched.
// =================================================================================================
struct SparseShared {
    double con_re[SG_EV_MAX_CONSTRAINTS * kMC], con_im[SG_EV_MAX_CONSTRAINTS * kMC], lim[SG_EV_MAX_CONSTRAINTS];
    double rate[40], energy[40];
    uint8_t cls[SG_EV_MAX_STATIONS], type[SG_EV_MAX_STATIONS];
};
constexpr int kSparseWarps = 4;

__global__ void __launch_bounds__(kSparseWarps * 32, 4)
ev_step_sparse_kernel(const sg_ev_model m, const sg_ev_state s, const long long B, const void* __restrict__ action,
                      const int flags, const int bins, float* __restrict__ obs, double* __restrict__ reward,
                      uint8_t* __restrict__ done, uint8_t* __restrict__ pilots, float* __restrict__ final_obs,
                      uint8_t* __restrict__ status, double* __restrict__ stats, int32_t* __restrict__ todo) {
    __shared__ SparseShared sh;
    __shared__ double sstat[11 * kSparseWarps];
    const int n = m.n_stations, kf = m.moer_steps, OBS = 2 * n + kf + 2, C = m.n_classes, J = m.n_constraints;
    const int K = m.max_sessions;
    for (int i = threadIdx.x; i < SG_EV_MAX_CONSTRAINTS * kMC; i += blockDim.x) { sh.con_re[i] = m.con_re[i]; sh.con_im[i] = m.con_im[i]; }
    if (threadIdx.x < SG_EV_MAX_CONSTRAINTS) sh.lim[threadIdx.x] = m.magnitudes[threadIdx.x];
    if (threadIdx.x < SG_EV_MAX_STATIONS) {
        sh.cls[threadIdx.x] = (int)threadIdx.x < n ? m.station_class[threadIdx.x] : 0;
        sh.type[threadIdx.x] = (int)threadIdx.x < n ? m.station_type[threadIdx.x] : 0;
    }
    if (threadIdx.x <= 32) {
        const double pp = div_c((double)threadIdx.x * kVoltage, 1000.0, kInv1000);
        const double rate = div_c(pp * 1000.0, kVoltage, kInv208);
        sh.rate[threadIdx.x] = rate;
        sh.energy[threadIdx.x] = div_c(rate * kVoltage, 1000.0, kInv1000) * kPeriodH;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const bool i64 = (flags & 1) == SG_EV_ACTION_I64;
    const uint4* recs = reinterpret_cast<const uint4*>(m.sessions);
    const long long ntiles = (B + 31) / 32;
    const long long wstride = (long long)gridDim.x * kSparseWarps;
    const bool col = 2 * lane < n;
    StatsAcc acc;

    struct Stn { int i, k, arr, dep, est; double energy, del; float a; bool plugged; };
    constexpr int V = 4;

    for (long long tile = (long long)blockIdx.x * kSparseWarps + (threadIdx.x >> 5); tile < ntiles; tile += wstride) {
        const long long b0 = tile * 32;
        const int nv = (int)min(32ll, B - b0);
        const long long b = b0 + lane;
        const bool env = lane < nv;
        int4 clk = make_int4(kEpisode, 0, 0, 0);
        unsigned long long occ = 0ull;
        if (env) { clk = *reinterpret_cast<const int4*>(s.clock + b * 4); occ = s.occ[b]; }
        const int t = clk.x, tn = t + 1;
        int day = clk.y, md = clk.z, nidx = clk.w;
        const bool live = env && t < kEpisode;
        bool commit = live;
        if (!live) occ = 0ull;
        const int sbase = day * K * n;
        float moer0f = 0.f;
        if (live) moer0f = __ldg(m.moer + ((long long)md * SG_EV_MOER_ROWS + tn) * SG_EV_MOER_COLS);

        // batch loader over the set bits of `mask` (V stations, all loads of a level in flight together)
        auto load_batch = [&](unsigned long long& mask, Stn (&q)[V]) {
#pragma unroll
            for (int u = 0; u < V; ++u) {
                q[u].i = -1;
                if (mask) { q[u].i = __ffsll((long long)mask) - 1; mask &= mask - 1; }
            }
#pragma unroll
            for (int u = 0; u < V; ++u) {
                q[u].k = 255; q[u].a = 0.f; q[u].del = 0.0;
                if (q[u].i >= 0) {
                    const long long o = b * n + q[u].i;
                    q[u].k = s.sess_idx[o];
                    q[u].del = s.delivered[o];
                    q[u].a = i64 ? __fdiv_rn((float)reinterpret_cast<const long long*>(action)[o], (float)(bins - 1))
                                 : __ldg(reinterpret_cast<const float*>(action) + o);
                }
            }
            uint4 raw[V];
#pragma unroll
            for (int u = 0; u < V; ++u) {
                raw[u] = make_uint4(0x7fff7fffu, 0x7fffu, 0u, 0u);
                if (q[u].i >= 0 && q[u].k < K) raw[u] = __ldg(recs + (sbase + q[u].k * n + q[u].i));
            }
#pragma unroll
            for (int u = 0; u < V; ++u) {
                q[u].arr = (int)(short)(raw[u].x & 0xffffu);
                q[u].dep = (int)(short)(raw[u].x >> 16);
                q[u].est = (int)(short)(raw[u].y & 0xffffu);
                q[u].energy = __hiloint2double((int)raw[u].w, (int)raw[u].z);
                q[u].plugged = q[u].i >= 0 && t >= 1 && q[u].arr <= t && t < q[u].dep;
            }
        };
        auto clipped = [&](const Stn& q) {
            const double rem = q.energy - q.del;
            const float dem = rem > kFullEps ? (float)rem : 0.f;                 // obs 'demands', env.py:388
            return fmin(fmax((double)q.a, 0.0), upper_bound(dem));
        };

        // ---- pass A: class sums of the clipped action -> is it inside the network constraints? ----
        {
            double S[kMC];
#pragma unroll
            for (int c = 0; c < kMC; ++c) S[c] = 0.0;
            unsigned long long mask = occ;
            while (__any_sync(kFull, mask != 0ull)) {
                Stn q[V];
                load_batch(mask, q);
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    if (q[u].plugged) {
                        const double x = clipped(q[u]);
                        const int c = sh.cls[q[u].i];
#pragma unroll
                        for (int cc = 0; cc < kMC; ++cc) S[cc] += (cc == c) ? x : 0.0;
                    }
                }
            }
            if (live) {
                bool ok = true;
                for (int j = 0; j < J; ++j) {
                    double wr = 0, wi = 0;
#pragma unroll
                    for (int c = 0; c < kMC; ++c) { wr = fma(sh.con_re[j * kMC + c], S[c], wr); wi = fma(sh.con_im[j * kMC + c], S[c], wi); }
                    const double r = sh.lim[j] * (1.0 / kActionScale);
                    ok = ok && (fma(wr, wr, wi * wi) <= r * r);
                }
                if (!ok) { commit = false; todo[2 + atomicAdd(todo, 1)] = (int)b; }
            }
        }
        const unsigned commit_mask = __ballot_sync(kFull, commit);
        const bool dn = tn >= kEpisode;
        const bool autoreset = commit && dn && (flags & SG_EV_FLAG_AUTORESET);
        const unsigned ar_mask = __ballot_sync(kFull, autoreset);

        // ---- zero-fill the committed rows (coalesced); the owners patch the few non-zero entries ----
        for (int r = 0; r < nv; ++r) {
            if (!((commit_mask >> r) & 1u)) continue;
            float* orow = (((ar_mask >> r) & 1u) && final_obs != nullptr) ? final_obs + (b0 + r) * OBS : obs + (b0 + r) * OBS;
            if (col) {
                stg_stream_f2(reinterpret_cast<float2*>(orow) + lane, make_float2(0.f, 0.f));
                stg_stream_f2(reinterpret_cast<float2*>(orow + n) + lane, make_float2(0.f, 0.f));
                if (pilots != nullptr) reinterpret_cast<uchar2*>(pilots + (b0 + r) * n)[lane] = make_uchar2(0, 0);
            }
        }
        __syncwarp();
        float* myrow = (autoreset && final_obs != nullptr) ? final_obs + b * OBS : obs + b * OBS;
        const bool patch = commit && !(autoreset && final_obs == nullptr);      // terminal obs kept only if asked for

        // ---- pass B: pilots, charging, departures for the occupied stations -------------------------
        double rate_sum = 0.0;
        int P[kMC];
#pragma unroll
        for (int c = 0; c < kMC; ++c) P[c] = 0;
        unsigned long long occ_new = commit ? occ : 0ull;
        {
            unsigned long long mask = commit ? occ : 0ull;
            while (__any_sync(kFull, mask != 0ull)) {
                Stn q[V];
                load_batch(mask, q);
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    if (q[u].i < 0) continue;
                    const int i = q[u].i;
                    double del = q[u].del;
                    if (q[u].plugged) {
                        const double pilot = to_pilot(clipped(q[u]), sh.type[i]);
                        const int pi = __double2int_rn(pilot);
                        const int c = sh.cls[i];
#pragma unroll
                        for (int cc = 0; cc < kMC; ++cc) P[cc] += (cc == c) ? pi : 0;
                        if (pilots != nullptr && pi != 0) pilots[b * n + i] = (uint8_t)pi;
                        const double charge = fmax(0.0, kCapacity - q[u].energy) + del;
                        if (kCapacity - charge > 2.0) {              // neither battery limit can bind
                            del += sh.energy[pi];
                            rate_sum += sh.rate[pi];
                        } else {
                            const double rate_to_full = div_c(kCapacity - charge, kPeriodH, kInvPeriodH);
                            const double soc = div_c(charge, kCapacity, kInv100);
                            const double p_pilot = div_c(pilot * kVoltage, 1000.0, kInv1000);
                            const double cap2 = soc < kTransitionSoc ? kMaxPower
                                                                     : div_c(1.0 - soc, kOneMinusTrans, kInvOMT) * kMaxPower;
                            const double power = fmin(fmin(p_pilot, cap2), rate_to_full);
                            const double rate = div_c(power * 1000.0, kVoltage, kInv208);
                            del += div_c(rate * kVoltage, 1000.0, kInv1000) * kPeriodH;
                            rate_sum += rate;
                        }
                    }
                    if (q[u].dep <= tn) {                             // Unplug event (timestamp = departure)
                        occ_new &= ~(1ull << i);
                        s.delivered[b * n + i] = 0.0;
                    } else {
                        if (q[u].plugged) s.delivered[b * n + i] = del;
                        const double rem = q[u].energy - del;
                        if (patch && rem > kFullEps) {                // Interface.active_sessions()
                            myrow[i] = (float)rem;
                            myrow[n + i] = (float)(q[u].est - tn);
                        }
                    }
                }
            }
        }
        // ---- arrivals: Plugin events with timestamp <= t+1 (after the Unplug events) ------------------
        if (commit && !autoreset) {
            const unsigned* arow = m.arrivals + (long long)day * (m.max_day_sessions + 1);
            unsigned ent = __ldg(arow + nidx);
            while ((int)(ent & 0x7fffu) <= tn) {
                const int i = (ent >> 16) & 0xff, k = ent >> 24;
                const uint4 r = __ldg(recs + (sbase + k * n + i));
                const int dep = (int)(short)(r.x >> 16), est = (int)(short)(r.y & 0xffffu);
                s.sess_idx[b * n + i] = (uint8_t)k;
                if (dep > tn) {                                       // (zero-length sessions never charge)
                    occ_new |= 1ull << i;
                    const double energy = __hiloint2double((int)r.w, (int)r.z);
                    if (patch && energy > kFullEps) { myrow[i] = (float)energy; myrow[n + i] = (float)(est - tn); }
                }
                ++nidx;
                ent = __ldg(arow + nidx);
            }
        }
        // ---- per-env epilogue ---------------------------------------------------------------------------
        int md_out = md;
        if (commit) {
            double excess = 0.0;
            for (int j = 0; j < J; ++j) {
                double wr = 0, wi = 0;
#pragma unroll
                for (int c = 0; c < kMC; ++c) { wr = fma(sh.con_re[j * kMC + c], (double)P[c], wr); wi = fma(sh.con_im[j * kMC + c], (double)P[c], wi); }
                const double lim = sh.lim[j], sq = fma(wr, wr, wi * wi);
                if (sq > lim * lim) excess += fmax(0.0, sqrt(sq) - lim);
            }
            const double profit = kProfitFactor * rate_sum;
            const double excess_charge = excess * kViolationFactor;
            const double carbon = kCarbonFactor * rate_sum * (double)moer0f;
            const double rew = profit - carbon - excess_charge;
            double* bd = s.breakdown + b * 3;
            if (t == 0) { bd[0] = profit; bd[1] = carbon; bd[2] = excess_charge; }
            else { bd[0] += profit; bd[1] += carbon; bd[2] += excess_charge; }
            int t_out = tn;
            if (autoreset) {
                day = (day + 1) % m.n_days;                           // sequential days, event_generation.py:284-291
                md_out = (md + 1) % m.n_moer_days;
                t_out = 0; nidx = 0; occ_new = 0ull;
            }
            s.occ[b] = occ_new;
            *reinterpret_cast<int4*>(s.clock + b * 4) = make_int4(t_out, day, md_out, nidx);
            reward[b] = rew;
            done[b] = dn ? 1 : 0;
            if (status != nullptr) status[b] = 0;
            if (stats != nullptr) acc.add(rew, dn, profit, carbon, excess_charge);
        } else if (env && t >= kEpisode) {                            // finished episode, not reset: stays terminal
            reward[b] = 0.0; done[b] = 1;
            if (status != nullptr) status[b] = 0;
        }
        // ---- MOER tails (and the reset rows of auto-reset envs), 8 rows of loads in flight -------------
        const float ts_new = (float)((double)tn / (double)kEpisode);            // env.py:392
        for (int r0 = 0; r0 < nv; r0 += 8) {
            float tail0[8], tail1[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = min(r0 + u, 31);
                const int md_r = __shfl_sync(kFull, md, r), tn_r = __shfl_sync(kFull, tn, r);
                tail0[u] = tail1[u] = 0.f;
                if (r0 + u < nv && ((commit_mask >> r) & 1u)) {
                    const float* row = m.moer + ((long long)md_r * SG_EV_MOER_ROWS + tn_r) * SG_EV_MOER_COLS;
                    if (lane < kf) tail0[u] = __ldg(row + 1 + lane); else if (lane == kf) tail0[u] = __ldg(row);
                    const int e = lane + 32;
                    if (e < kf) tail1[u] = __ldg(row + 1 + e); else if (e == kf) tail1[u] = __ldg(row);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + u;
                if (r >= nv || !((commit_mask >> r) & 1u)) continue;
                const bool ar = (ar_mask >> r) & 1u;
                const float ts_r = __shfl_sync(kFull, ts_new, r);
                if (!(ar && final_obs == nullptr)) {
                    float* orow = (ar ? final_obs : obs) + (b0 + r) * OBS;
                    if (lane <= kf) orow[2 * n + lane] = tail0[u]; else if (lane == kf + 1) orow[2 * n + lane] = ts_r;
                    if (lane + 32 <= kf) orow[2 * n + lane + 32] = tail1[u]; else if (lane + 32 == kf + 1) orow[2 * n + lane + 32] = ts_r;
                }
                if (ar) {                                             // reset observation of the next episode
                    const int mdo_r = __shfl_sync(kFull, md_out, r);
                    const float* row = m.moer + ((long long)mdo_r * SG_EV_MOER_ROWS) * SG_EV_MOER_COLS;
                    float* o2 = obs + (b0 + r) * OBS;
                    if (col && final_obs != nullptr) {                // (already zero when the terminal obs is not kept)
                        stg_stream_f2(reinterpret_cast<float2*>(o2) + lane, make_float2(0.f, 0.f));
                        stg_stream_f2(reinterpret_cast<float2*>(o2 + n) + lane, make_float2(0.f, 0.f));
                    }
#pragma unroll
                    for (int pass = 0; pass < 2; ++pass) {
                        const int e = lane + 32 * pass;
                        if (e < kf + 2) o2[2 * n + e] = e < kf ? __ldg(row + 1 + e) : (e == kf ? __ldg(row) : 0.f);
                    }
                    for (int i = lane; i < n; i += 32) s.sess_idx[(b0 + r) * n + i] = 0;    // new day: slot 0 everywhere
                }
            }
        }
        __syncwarp();
    }
    if (stats != nullptr) {
        __syncthreads();
        stats_flush_block(acc, stats, sstat);
    }
}

// Persistent CTAs: every warp strides over the batch, so the table staging and the statistics
// flush are paid once per warp instead of once per env.
template <int MODE>
__global__ void __launch_bounds__(kThreads, 3)
ev_step_kernel(const sg_ev_model m, const sg_ev_state s, const long long B, const void* __restrict__ action,
               const int flags, const int bins, float* __restrict__ obs, double* __restrict__ reward,
               uint8_t* __restrict__ done, uint8_t* __restrict__ pilots, float* __restrict__ final_obs,
               uint8_t* __restrict__ status, double* __restrict__ stats, int32_t* __restrict__ todo) {
    __shared__ EvShared sh;
    load_shared(m, sh);
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * kWarpsPerBlock;
    StatsAcc acc;
    for (long long b = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); b < B; b += stride)
        ev_env_step<MODE>(m, s, sh, b, lane, action, flags, bins, obs, reward, done, pilots, final_obs, status, todo, acc);
    if (stats != nullptr) {
        __syncthreads();
        stats_flush_block(acc, stats, sh.stat);
    }
}

// Work-list kernel: envs whose clipped action was not certainly feasible (MODE 1 deferred them).
__global__ void __launch_bounds__(128)
ev_step_solve_kernel(const sg_ev_model m, const sg_ev_state s, const void* __restrict__ action, const int flags,
                     const int bins, float* __restrict__ obs, double* __restrict__ reward,
                     uint8_t* __restrict__ done, uint8_t* __restrict__ pilots, float* __restrict__ final_obs,
                     uint8_t* __restrict__ status, double* __restrict__ stats, int32_t* __restrict__ todo) {
    __shared__ EvShared sh;
    load_shared(m, sh);
    const int lane = threadIdx.x & 31;
    const int count = todo[0];
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    StatsAcc acc;
    // warp w of CTA c takes item c + w*gridDim.x: the first items spread over different SMs
    for (int w = blockIdx.x + (threadIdx.x >> 5) * gridDim.x; w < count; w += nwarps) {
        const long long b = todo[2 + w];
        ev_env_step<2>(m, s, sh, b, lane, action, flags, bins, obs, reward, done, pilots, final_obs, status, nullptr, acc);
    }
    if (stats != nullptr) {
        __syncthreads();
        stats_flush_block(acc, stats, sh.stat);
    }
    // last CTA to finish re-arms the work list for the next step (no memset node needed)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(todo + 1, 1) == (int)gridDim.x - 1) { todo[0] = 0; todo[1] = 0; __threadfence(); }
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
    for (int i = lane; i < n; i += 32) { s.delivered[b * n + i] = 0.0; s.sess_idx[b * n + i] = 0; }
    float* orow = obs + b * OBS;
    for (int i = lane; i < 2 * n; i += 32) orow[i] = 0.f;
    const float* mrow = m.moer + ((long long)md * SG_EV_MOER_ROWS) * SG_EV_MOER_COLS;
    for (int e = lane; e < kf + 2; e += 32) orow[2 * n + e] = e < kf ? mrow[1 + e] : (e == kf ? mrow[0] : 0.f);
    if (lane == 0) {
        s.clock[b * 4 + 0] = 0; s.clock[b * 4 + 1] = day; s.clock[b * 4 + 2] = md; s.clock[b * 4 + 3] = 0;
        if (s.occ != nullptr) s.occ[b] = 0ull;
        s.breakdown[b * 3] = s.breakdown[b * 3 + 1] = s.breakdown[b * 3 + 2] = 0.0;
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

}  // namespace
}  // namespace sg

extern "C" {

int sg_ev_reset(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs, const uint8_t* reset_mask,
                const int32_t* day_in, const int32_t* moer_day_in, float* obs, void* stream) {
    if (int rc = sg::check_model(m)) return rc;
    SG_CHECK_ARG(s && obs && m->moer, "ev_reset: NULL argument");
    SG_CHECK_ARG(m->n_days >= 1 && m->n_moer_days >= 1, "ev_reset: empty trace/MOER pool");
    if (num_envs <= 0) return 0;
    const unsigned grid = (unsigned)((num_envs + 3) / 4);
    sg::ev_reset_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*m, *s, num_envs, reset_mask,
                                                                                day_in, moer_day_in, obs);
    return sg::launch_status("ev_reset_kernel");
}

int sg_ev_step(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs, const void* action, int32_t flags,
               int32_t discrete_bins, float* obs, double* reward, uint8_t* done, uint8_t* pilots,
               float* final_obs, uint8_t* status, double* stats, int32_t* todo, void* stream) {
    if (int rc = sg::check_model(m)) return rc;
    SG_CHECK_ARG(s && action && obs && reward && done, "ev_step: NULL argument");
    SG_CHECK_ARG(m->sessions && m->moer && m->n_days >= 1 && m->n_moer_days >= 1 && m->max_sessions >= 1,
                 "ev_step: empty trace/MOER pool");
    SG_CHECK_ARG(!((flags & 1) == SG_EV_ACTION_I64) || discrete_bins >= 2, "ev_step: discrete_bins must be >= 2");
    SG_CHECK_ARG(!m->project || todo != nullptr, "ev_step: projection needs the todo scratch buffer");
    if (num_envs <= 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int sms = sg_device_sm_count();
    if (sms <= 0) sms = 148;
    // persistent grid: 3 resident CTAs per SM (multiple of the SM count), fewer for tiny batches
    const unsigned grid = (unsigned)std::min<long long>((num_envs + sg::kWarpsPerBlock - 1) / sg::kWarpsPerBlock,
                                                        (long long)sms * 3);
    // specialised kernel: even n, <= 5 classes, <= 32 constraints, 32-bit record index; the model
    // flag bit 8 (SG_EV_FLAG_GENERIC_KERNEL) forces the generic one (used by the tests)
    const bool small_index = (long long)m->n_days * m->max_sessions * m->n_stations < (1ll << 31);
    const bool dense_ok = (m->n_stations % 2 == 0) && small_index && m->max_sessions < 255 &&
                          !(flags & SG_EV_FLAG_GENERIC_KERNEL);
    // projection on: event-driven kernel (work proportional to the occupied stations)
    if (m->project && dense_ok && m->arrivals != nullptr && s->occ != nullptr && !(flags & SG_EV_FLAG_TILE_KERNEL)) {
        const long long ntiles = (num_envs + 31) / 32;
        const unsigned sgrid = (unsigned)std::min<long long>((ntiles + sg::kSparseWarps - 1) / sg::kSparseWarps, (long long)sms * 4);
        sg::ev_step_sparse_kernel<<<sgrid, sg::kSparseWarps * 32, 0, st>>>(*m, *s, num_envs, action, flags, discrete_bins, obs,
                                                                        reward, done, pilots, final_obs, status, stats, todo);
        return sg::launch_status("ev_step_sparse_kernel");
    }
    // dense tile kernel: one thread per env on 32-env tiles staged through shared memory
    if (dense_ok) {
        const sg::TileLayout L(m->n_stations);
        const int per_sm = std::max(1, std::min(16, (220 * 1024) / (L.bytes + 1024)));
        const long long ntiles = (num_envs + sg::kTileE - 1) / sg::kTileE;
        const unsigned tgrid = (unsigned)std::min<long long>(ntiles, (long long)sms * per_sm);
        if (!m->project) {
            auto kern = sg::ev_step_tile_kernel<0>;
            if (L.bytes > 48 * 1024) SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.bytes));
            kern<<<tgrid, sg::kTileE, L.bytes, st>>>(*m, *s, num_envs, action, flags, discrete_bins, obs, reward, done,
                                                     pilots, final_obs, status, stats, todo);
        } else {
            auto kern = sg::ev_step_tile_kernel<1>;
            if (L.bytes > 48 * 1024) SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.bytes));
            kern<<<tgrid, sg::kTileE, L.bytes, st>>>(*m, *s, num_envs, action, flags, discrete_bins, obs, reward, done,
                                                     pilots, final_obs, status, stats, todo);
        }
        return sg::launch_status("ev_step_tile_kernel");
    }
    // generic warp-per-env kernel (any network / odd station counts)
    if (!m->project) {
        sg::ev_step_kernel<0><<<grid, sg::kThreads, 0, st>>>(*m, *s, num_envs, action, flags, discrete_bins, obs,
                                                           reward, done, pilots, final_obs, status, stats, todo);
        return sg::launch_status("ev_step_kernel<0>");
    }
    sg::ev_step_kernel<1><<<grid, sg::kThreads, 0, st>>>(*m, *s, num_envs, action, flags, discrete_bins, obs, reward,
                                                       done, pilots, final_obs, status, stats, todo);
    return sg::launch_status("ev_step_kernel<1>");
}

int sg_ev_step_finish(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs, const void* action,
                      int32_t flags, int32_t discrete_bins, float* obs, double* reward, uint8_t* done,
                      uint8_t* pilots, float* final_obs, uint8_t* status, double* stats, int32_t* todo,
                      void* stream) {
    if (int rc = sg::check_model(m)) return rc;
    if (!m->project || num_envs <= 0) return 0;          // nothing is ever deferred without projection
    SG_CHECK_ARG(s && action && obs && reward && done && todo, "ev_step_finish: NULL argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int sms = sg_device_sm_count();
    if (sms <= 0) sms = 148;
    const unsigned sgrid = (unsigned)std::min<long long>((num_envs + 3) / 4, (long long)sms * 2);
    sg::ev_step_solve_kernel<<<sgrid, 128, 0, st>>>(*m, *s, action, flags, discrete_bins, obs, reward, done, pilots,
                                                  final_obs, status, stats, todo);
    return sg::launch_status("ev_step_solve_kernel");
}

int sg_ev_step_host(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs, const void* h_action,
                    int32_t flags, int32_t discrete_bins, float* h_obs, double* h_reward, uint8_t* h_done,
                    void* d_action, float* d_obs, double* d_reward, uint8_t* d_done, uint8_t* d_status,
                    double* d_stats, int32_t* d_todo, void* stream) {
    SG_CHECK_ARG(m && h_action && h_obs && h_reward && h_done && d_action && d_obs && d_reward && d_done,
                 "ev_step_host: NULL argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t B = (size_t)num_envs, n = (size_t)m->n_stations, OBS = 2 * n + m->moer_steps + 2;
    const size_t abytes = B * n * (((flags & 1) == SG_EV_ACTION_I64) ? 8 : 4);
    SG_CUDA(cudaMemcpyAsync(d_action, h_action, abytes, cudaMemcpyHostToDevice, st));
    if (int rc = sg_ev_step(m, s, num_envs, d_action, flags, discrete_bins, d_obs, d_reward, d_done, nullptr,
                            nullptr, d_status, d_stats, d_todo, stream)) return rc;
    if (int rc = sg_ev_step_finish(m, s, num_envs, d_action, flags, discrete_bins, d_obs, d_reward, d_done, nullptr,
                                   nullptr, d_status, d_stats, d_todo, stream)) return rc;
    SG_CUDA(cudaMemcpyAsync(h_obs, d_obs, B * OBS * 4, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(h_reward, d_reward, B * 8, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(h_done, d_done, B, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    return 0;
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
