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
// One kernel launch per step (ev_step_main_kernel, documented at its definition).  Tile CTAs: a warp owns a tile of
// 16 consecutive envs (8 in a thin last round), lists the OCCUPIED stations of the tile (64-bit occupancy masks), and
// works one (env, station) item per lane on packed per-env lists of 32-byte session records; the observation / pilot
// rows of the tile are assembled in shared memory and leave as one bulk (TMA) copy each.  The network constraints
// depend on the schedule only through <= 8 per-class sums, kept per env in shared memory.
//
// The projection has a fast path (the box-clipped action already satisfies the network: it IS the projection),
// certified per env with integer screens and an f32 margin test, an exact single-class (pod) solve run by the tile
// warp itself, and — for the few envs that need it — a primal-dual interior point on the QCQP (Newton systems reduced
// to n_classes x n_classes + n_constraints x n_constraints Cholesky factorisations) run by helper CTAs of the same
// launch, which take those envs from a work list while the tile warps go on.
//
// All divisions by constants use the FMA-corrected reciprocal (Markstein) which reproduces the
// IEEE quotient the reference's Python arithmetic produces (checked on 2e7 samples per constant,
// tests/test_div_const.py), so results stay bit-comparable with the oracle.
//
// Bound: HBM.  Algorithmic bytes per env-step: SURVEY.md 8d (1709 B at n=54, k=36).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

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

// layout of the `work` scratch buffer (SG_EV_WORK_INTS(B) int32, zero-initialised once, re-armed by every step)
// (three 128-byte lines: the helper warps poll the third while the tile warps hammer the second with ticket requests)
constexpr int kWorkCount = 0;        // entries appended to the exact-projection work list
constexpr int kWorkTaken = 1;        // ... and entries claimed by the helper warps
constexpr int kWorkNextTile = 32;    // dynamic tile scheduler of ev_step_main_kernel: tiles handed out so far
constexpr int kWorkRetired = 64;     // tile warps that have left the tile loop
constexpr int kWorkHelpersDone = 65; // helper warps that have finished (the last one re-arms the buffer)
constexpr int kWorkList = SG_EV_WORK_HEADER_INTS;   // first entry of the work list: env index + 1, 0 = not yet written

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

// Per-model constant tables, built once by sg_ev_prepare_tables (ev_prepare_kernel) and copied to
// shared memory by every CTA with 16-byte vector loads.  con_re / con_im are zero-padded to
// [SG_EV_MAX_CONSTRAINTS][SG_EV_MAX_CLASSES] and lim to [SG_EV_MAX_CONSTRAINTS].
constexpr float kPreMargin = 2e-3f;                  // normalised units; covers the f32 accumulation error
struct __align__(16) EvTables {
    double con_re[SG_EV_MAX_CONSTRAINTS * kMC];
    double con_im[SG_EV_MAX_CONSTRAINTS * kMC];
    double lim[SG_EV_MAX_CONSTRAINTS];               // magnitudes in A
    double ppilot[40];                               // pilot (A) -> kW at 208 V, the reference's IEEE operation sequence
    float re[SG_EV_MAX_CONSTRAINTS * kMC], im[SG_EV_MAX_CONSTRAINTS * kMC];   // f32 copies for the pre-checks
    float r_in[SG_EV_MAX_CONSTRAINTS];               // (limit/32 - margin)^2 in normalised action units
    float lim2_lo[SG_EV_MAX_CONSTRAINTS];            // 0.999 * limit^2 (A^2): below it the rounded schedule has no excess
    double simple_R[SG_EV_MAX_CONSTRAINTS];          // single-class constraints: limit on the normalised class sum
    int8_t simple_cls[SG_EV_MAX_CONSTRAINTS];        // class of a constraint with ONE non-zero coefficient (a pod), else -1
    uint8_t cls[SG_EV_MAX_STATIONS];                 // class of each station, 255 = padding
    uint8_t type[SG_EV_MAX_STATIONS];
    // Integer screens of the per-env constraint pre-checks (owner lane, see env_checks).  Class sums are kept as
    // integers (clipped action: 2^-20 fixed point; pilots: A), so the common case is a handful of compares:
    //   single-class constraint q (a pod): certainly satisfied iff the class sum <= simple_S_fix[q] / simple_P_lim[q];
    //   coupled constraints: |sum_c coef_c S_c| <= max_c|coef_c| * sum_c S_c, so all of them are certainly satisfied
    //   iff the sum over ALL classes <= scr_S_fix / scr_P_lim; only otherwise they are evaluated (f32, with margin).
    int simple_S_fix[SG_EV_MAX_CONSTRAINTS];
    int simple_P_lim[SG_EV_MAX_CONSTRAINTS];
    uint8_t simple_j[SG_EV_MAX_CONSTRAINTS];         // the single-class constraints, n_simple of them
    int n_simple, scr_S_fix, scr_P_lim, pad_;
};
static_assert(sizeof(EvTables) % 16 == 0 && sizeof(EvTables) <= SG_EV_TABLE_BYTES, "table block size");
using EvShared = EvTables;

__global__ void ev_prepare_kernel(const sg_ev_model m, EvTables* __restrict__ tb) {
    for (int i = threadIdx.x; i < SG_EV_MAX_CONSTRAINTS * kMC; i += blockDim.x) {
        tb->con_re[i] = m.con_re[i]; tb->con_im[i] = m.con_im[i];
        tb->re[i] = (float)m.con_re[i]; tb->im[i] = (float)m.con_im[i];
    }
    for (int i = threadIdx.x; i < SG_EV_MAX_CONSTRAINTS; i += blockDim.x) {
        const double lim = m.magnitudes[i];
        const double r = lim * (1.0 / kActionScale) - (double)kPreMargin;
        tb->lim[i] = lim;
        tb->r_in[i] = r > 0.0 ? (float)(r * r * 0.99999) : -1.f;
        tb->lim2_lo[i] = (float)(lim * lim * 0.999);
        int nz = 0, cstar = -1;
        for (int c = 0; c < kMC; ++c)
            if (m.con_re[i * kMC + c] != 0.0 || m.con_im[i * kMC + c] != 0.0) { nz++; cstar = c; }
        tb->simple_cls[i] = (int8_t)(nz == 1 && i < m.n_constraints ? cstar : -1);
        double kappa = 1.0;
        if (nz == 1) kappa = sqrt(m.con_re[i * kMC + cstar] * m.con_re[i * kMC + cstar] + m.con_im[i * kMC + cstar] * m.con_im[i * kMC + cstar]);
        tb->simple_R[i] = lim * (1.0 / kActionScale) / kappa;
    }
    if (threadIdx.x == 0) {
        int ns = 0;
        double scrS = 1e30, scrP = 1e30;
        for (int j = 0; j < SG_EV_MAX_CONSTRAINTS; ++j) { tb->simple_j[j] = 0; tb->simple_S_fix[j] = 0; tb->simple_P_lim[j] = 0; }
        for (int j = 0; j < m.n_constraints; ++j) {
            int nz = 0;
            double kmax = 0.0;
            for (int c = 0; c < kMC; ++c) {
                const double mag = sqrt(m.con_re[j * kMC + c] * m.con_re[j * kMC + c] + m.con_im[j * kMC + c] * m.con_im[j * kMC + c]);
                if (mag != 0.0) nz++;
                kmax = fmax(kmax, mag);
            }
            const double lim = m.magnitudes[j];
            // thresholds mirror the f32 tests they stand in for: (limit/32 - margin) for the clipped action,
            // sqrt(0.999) * limit for the rounded schedule; both scaled down a little more, rounded down
            const double rS = (lim * (1.0 / kActionScale) - (double)kPreMargin) * 0.99999;
            const double rP = sqrt(0.999) * lim * 0.99999;
            if (nz == 0) continue;                        // empty row: never violated
            if (nz == 1) {
                tb->simple_j[ns] = (uint8_t)j;
                tb->simple_S_fix[ns] = rS > 0.0 ? (int)fmin(floor(rS / kmax * kFixScale), 2.0e9) : -1;
                tb->simple_P_lim[ns] = (int)fmin(floor(rP / kmax), 2.0e9);
                ns++;
            } else {
                scrS = fmin(scrS, rS > 0.0 ? rS / kmax : -1.0);
                scrP = fmin(scrP, rP / kmax);
            }
        }
        tb->n_simple = ns;
        tb->scr_S_fix = scrS < 0.0 ? -1 : (int)fmin(floor(scrS * kFixScale), 2.0e9);
        tb->scr_P_lim = (int)fmin(floor(scrP), 2.0e9);
        tb->pad_ = 0;
    }
    for (int i = threadIdx.x; i < 40; i += blockDim.x) tb->ppilot[i] = div_c((double)i * kVoltage, 1000.0, kInv1000);
    for (int i = threadIdx.x; i < SG_EV_MAX_STATIONS; i += blockDim.x) {
        const bool in = i < m.n_stations;
        tb->cls[i] = in ? m.station_class[i] : 255;
        tb->type[i] = in ? m.station_type[i] : 0;
    }
}

// Observation-tail table: row (md, t) = forecasted_moer[k] | prev_moer | t / 288 (| padding), env.py:389-392
__global__ void ev_prepare_tail_kernel(const float* __restrict__ moer, const int n_moer_days, const int kf,
                                       float* __restrict__ tail) {
    const int TL = SG_EV_TAIL_FLOATS(kf);
    const long long rows = (long long)n_moer_days * SG_EV_MOER_ROWS;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < rows * TL;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / TL;
        const int q = (int)(idx % TL), t = (int)(r % SG_EV_MOER_ROWS);
        const float* row = moer + r * SG_EV_MOER_COLS;
        tail[idx] = q < kf ? row[1 + q] : (q == kf ? row[0] : (q == kf + 1 ? (float)((double)t * (1.0 / (double)kEpisode)) : 0.f));
    }
}

// CTA-wide copy of the table block into shared memory (caller synchronises)
__device__ __forceinline__ void load_tables(const void* __restrict__ src, EvTables& sh) {
    const uint4* g = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(&sh);
    for (int i = threadIdx.x; i < (int)(sizeof(EvTables) / 16); i += blockDim.x) d[i] = __ldg(g + i);
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
// kept per env as a PACKED LIST in station order — the record of station i is entry popc(occ & ((1 << i) - 1)) of
// the env's [n]-entry block — so the ~6 live records of an env are contiguous (2-3 cache lines instead of 6
// scattered sectors: the kernel is bound by the NUMBER of memory requests, profiles/r2/dram_random.cu), and
// updated IN PLACE: a record is written back right after it was read,
// while its line is still in L2 (measured on B200: writing the records to a second, not yet cached
// copy of the state cost 22 us per 65 536-env step, in place it is free).  Only the occupied
// stations (bit set in the env's 64-bit occupancy mask) hold live records, so one step touches
// ~occupancy x 32 B of state instead of a dense [n] row.  The records of an env are written only
// after its clipped action passed the constraint check, so the rare exact-projection path finds
// the state of the other envs untouched and needs no undo log.
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

// Same record as two 16-byte stores.  Used on the rare path: inside the large non-inlined fallback
// function ptxas 12.9 was observed to narrow the 256-bit asm store to its first 8 bytes (SASS
// STG.E.64), which the parity tests caught; the split form compiles to two ST.E.128.
__device__ __forceinline__ void st_slot_2x16(Slot* p, const Slot& s) {
    uint4 lo, hi;
    lo.x = (unsigned)__double2loint(s.energy); lo.y = (unsigned)__double2hiint(s.energy);
    lo.z = (unsigned)__double2loint(s.delivered); lo.w = (unsigned)__double2hiint(s.delivered);
    hi.x = (unsigned)s.dep; hi.y = (unsigned)s.est; hi.z = (unsigned)s.arr; hi.w = (unsigned)s.k;
    reinterpret_cast<uint4*>(p)[0] = lo;
    reinterpret_cast<uint4*>(p)[1] = hi;
}

// Per-env scalars, one 64-byte record per env, read at the start of a step and written back in place at its
// end.  Two 32-byte halves, each moved with ONE 256-bit access per owner lane (a tile's 16 headers are then 4 + 4
// line requests instead of 3 x 8 with 16-byte loads; the step kernel is bound by the number of memory requests):
//   hot  { i32 t, day, moer_day, arrival cursor | u64 occupancy mask | pad }
//   cold { f64 cumulative profit, carbon_cost, excess_charge | pad }
struct __align__(32) EnvHeader {
    int t, day, moer_day, nidx;                      // env.py:168 t; trace / MOER table index; arrival cursor
    unsigned long long occ;                          // bit i: station i has a plugged-in session
    unsigned long long pad0;
    double bd[3];                                    // cumulative profit, carbon_cost, excess_charge (env.py:460-462)
    unsigned long long pad1;
};
static_assert(sizeof(EnvHeader) == 64, "EnvHeader must be 64 bytes");

__device__ __forceinline__ void ld_v4u64(const void* p, unsigned long long (&v)[4]) {
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st_v4u64(void* p, const unsigned long long (&v)[4]) {
    asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v[0]), "l"(v[1]), "l"(v[2]), "l"(v[3]) : "memory");
}
__device__ __forceinline__ EnvHeader ld_header(const EnvHeader* p) {
    unsigned long long a[4], c[4];
    ld_v4u64(p, a);
    ld_v4u64(reinterpret_cast<const unsigned char*>(p) + 32, c);
    EnvHeader h;
    h.t = (int)(unsigned)(a[0] & 0xffffffffull); h.day = (int)(unsigned)(a[0] >> 32);
    h.moer_day = (int)(unsigned)(a[1] & 0xffffffffull); h.nidx = (int)(unsigned)(a[1] >> 32);
    h.occ = a[2]; h.pad0 = 0ull;
    h.bd[0] = __longlong_as_double((long long)c[0]); h.bd[1] = __longlong_as_double((long long)c[1]);
    h.bd[2] = __longlong_as_double((long long)c[2]); h.pad1 = 0ull;
    return h;
}
__device__ __forceinline__ void st_header(EnvHeader* p, const EnvHeader& h) {
    const unsigned long long a[4] = {(unsigned long long)(unsigned)h.t | ((unsigned long long)(unsigned)h.day << 32),
                                     (unsigned long long)(unsigned)h.moer_day | ((unsigned long long)(unsigned)h.nidx << 32),
                                     h.occ, 0ull};
    const unsigned long long c[4] = {(unsigned long long)__double_as_longlong(h.bd[0]), (unsigned long long)__double_as_longlong(h.bd[1]),
                                     (unsigned long long)__double_as_longlong(h.bd[2]), 0ull};
    st_v4u64(p, a);
    st_v4u64(reinterpret_cast<unsigned char*>(p) + 32, c);
}
__device__ __forceinline__ EnvHeader empty_header() {
    EnvHeader h;
    h.t = kEpisode; h.day = 0; h.moer_day = 0; h.nidx = 0; h.occ = 0ull;
    h.bd[0] = h.bd[1] = h.bd[2] = 0.0; h.pad0 = h.pad1 = 0ull;
    return h;
}
// position of station i's record in the env's packed list: the occupied stations below it
__device__ __forceinline__ int slot_rank(const unsigned long long occ, const int i) {
    return __popcll(occ & ((1ull << i) - 1ull));
}

#ifndef SG_EV_KE
#define SG_EV_KE 16
#endif
#ifndef SG_EV_MINB
#define SG_EV_MINB 3     // resident CTAs per SM the kernel is compiled for (3 x 128 threads -> <= 168 registers, no spills)
#endif
#ifndef SG_EV_KV
#define SG_EV_KV 2
#endif
#ifndef SG_EV_ABL
#define SG_EV_ABL 0      // timing ablations (diagnostic builds only; results are wrong when non-zero)
#endif
constexpr int kE = SG_EV_KE;                         // envs per warp tile
constexpr int kQ = 32 / kE;                          // lanes per env in the spread phases
constexpr int kSegBits = 64 / kQ;                    // occupancy bits enumerated by one lane
#ifndef SG_EV_WARPS
#define SG_EV_WARPS 4
#endif
constexpr int kMainWarps = SG_EV_WARPS;              // warps per CTA of the main kernel
#ifndef SG_EV_HELPER_CTAS
#define SG_EV_HELPER_CTAS 4                          // helper CTAs of a full-size launch (exact projections, see the kernel)
#endif
#ifndef SG_EV_RATESUM_SPLIT
#define SG_EV_RATESUM_SPLIT 1
#endif
#ifndef SG_EV_INLINE_POD
#define SG_EV_INLINE_POD 1                           // 1: single-class (pod) projections on the tile warp itself; 0: every env that
                                                     // fails the margin test goes to the helper warps (measured: the tile warps then
                                                     // finish at 43 us instead of 54, but a helper needs ~25 us per env — cold code
                                                     // on SMs it shares with tile CTAs — and the launch ends at 60-67 us)
#endif
constexpr int kMainThreads = kMainWarps * 32;
constexpr int kV = SG_EV_KV;                         // work items per lane and chunk (loads in flight)
constexpr int kItemCap = 32 * kV;                    // (env, station) work items staged per warp and chunk
constexpr int kTailMax = (SG_EV_MOER_COLS + 1) / 2;  // float2 pairs in the longest MOER + timestep tail
constexpr int kTailPerLane = (kTailMax + kQ - 1) / kQ;
static_assert(kE == 4 || kE == 8 || kE == 16 || kE == 32, "tile size");

struct WarpShared {
    double rate[kItemCap];                           // charging rate (A) of item j
    unsigned short key[kItemCap];                    // env << 8 | station of item j
    int S[kMC][kE];                                  // per-env class sums of the clipped action, 2^-20 fixed point
    int P[kMC][kE];                                  // per-env class sums of the integer pilots
    unsigned occ_clr[kE][2];                         // occupancy bits cleared by Unplug events
    unsigned long long occ_fin[kE];                  // occupancy after this step's Unplug and Plugin events (list ranks)
    float xa[kItemCap];                              // rare path: action of item j (its upper bound goes to rate[j])
    unsigned char icls[kItemCap];                    // rare path: station class of item j
    double nu[kMC];                                  // rare path: class multipliers of one env's projection
    double stat[12];                                 // this warp's reward statistics (SG_STATS layout)
};

// the fields of sg_ev_model / sg_ev_state / launch arguments the rare path needs, passed by value
struct EvCtx {
    int n, kf, C, J, K, n_days, n_moer_days, max_day_sessions, flags, bins;
    long long B;
    const float* moer;
    const uint4* recs;
    const unsigned* arrivals;
    const void* action;
    Slot* slots;
    EnvHeader* header;
    float* obs;
    float* final_obs;
    double* reward;
    uint8_t* done;
    uint8_t* pilots;
    uint8_t* status;
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

__device__ __forceinline__ float act_value(const void* __restrict__ action, const bool i64, const int bins, const long long o) {
    return i64 ? __fdiv_rn((float)reinterpret_cast<const long long*>(action)[o], (float)(bins - 1))   // wrappers.py:45
               : __ldg(reinterpret_cast<const float*>(action) + o);
}

__device__ __forceinline__ float moer_tail(const float* __restrict__ row, const int q, const int kf, const float ts) {
    return q < kf ? __ldg(row + 1 + q) : (q == kf ? __ldg(row) : ts);    // forecasted_moer | prev_moer | timestep
}

// =================================================================================================
// Rare path: the complete step of ONE env by a whole warp (lane l owns stations 2l and 2l+1), with
// the exact action projection: exact feasibility test of the box-clipped action, the single-class
// (pod) projection, else the interior-point solver.  Called for the envs the main path could not
// certify as feasible; re-reads the env's state (the main path never overwrites the state copy of
// the current step) and rewrites everything the main path wrote for this env.  The WITH_IPM = false
// instance (small) returns false without side effects when the interior-point solver is needed;
// the caller then runs the WITH_IPM = true instance.
// =================================================================================================
__device__ __forceinline__ int pod_solve(const EvTables& sh, WarpShared& ws, const int lo, const int hi, const int J,
                                         const unsigned suspects, const int lane);

template <bool WITH_IPM>
__device__ __noinline__ bool fallback_env(const EvTables& sh, WarpShared& ws, const EvCtx cx, const long long b, const int lane,
                                           double* wstat) {
    const int n = cx.n, kf = cx.kf, OBS = 2 * n + kf + 2, C = cx.C, J = cx.J;
    const bool i64 = (cx.flags & 1) == SG_EV_ACTION_I64;
    const EnvHeader hd = cx.header[b];
    const unsigned long long occ = hd.occ;
    const int t = hd.t, tn = t + 1, md = hd.moer_day;
    int day = hd.day, nidx = hd.nidx;
    const bool dn = tn >= kEpisode;
    const bool autoreset = dn && (cx.flags & SG_EV_FLAG_AUTORESET);
    const int md_out = autoreset ? (md + 1) % cx.n_moer_days : md;
    Slot* cur = cx.slots + b * n;
    Slot* nxt = cur;                                   // records are updated in place
    const int i0 = 2 * lane;
    const bool valid[2] = {i0 < n, i0 + 1 < n};
    double a[2], u[2], x[2];
    int cls[2], type[2];
    bool occd[2];
    Slot q[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int i = i0 + p;
        a[p] = 0.0; u[p] = 0.0; cls[p] = 255; type[p] = 0;
        occd[p] = valid[p] && ((occ >> i) & 1ull);
        q[p].energy = 0.0; q[p].delivered = 0.0; q[p].dep = q[p].est = q[p].arr = q[p].k = 0;
        if (valid[p]) {
            a[p] = (double)act_value(cx.action, i64, cx.bins, b * n + i);
            cls[p] = sh.cls[i]; type[p] = sh.type[i];
            if (occd[p]) {
                q[p] = ld_slot(cur + slot_rank(occ, i));
                const double rem = q[p].energy - q[p].delivered;
                const float dem = rem > kFullEps ? (float)rem : 0.f;          // obs 'demands', env.py:388
                u[p] = upper_bound(dem);
            }
        }
        x[p] = valid[p] ? fmin(fmax(a[p], 0.0), u[p]) : 0.0;
    }
    int st = 0;
    {
        // exact feasibility test of the clipped action and, when only single-class constraints (the pods) are violated,
        // the exact projection: the occupied stations go to the warp's item lists, pod_solve runs on the whole warp
        const unsigned m0 = __ballot_sync(kFull, occd[0]), m1 = __ballot_sync(kFull, occd[1]);
        const unsigned lt = (1u << lane) - 1u;
        const int pos0 = __popc(m0 & lt) + __popc(m1 & lt);
#pragma unroll
        for (int p = 0; p < 2; ++p)
            if (occd[p]) {
                const int pos = pos0 + (p == 1 && occd[0] ? 1 : 0);
                ws.xa[pos] = (float)a[p]; ws.rate[pos] = u[p]; ws.icls[pos] = (unsigned char)cls[p];
            }
        __syncwarp();
        const int res = pod_solve(sh, ws, 0, __popc(m0) + __popc(m1), J, J >= 32 ? 0xffffffffu : (1u << J) - 1u, lane);
        __syncwarp();
        if (res == 1) {
#pragma unroll
            for (int p = 0; p < 2; ++p) if (occd[p]) x[p] = fmin(fmax(a[p] - ws.nu[cls[p]], 0.0), u[p]);
            st = SG_EV_STATUS_PROJECTED;
        } else if (res == 2) {
            if (!WITH_IPM) return false;                 // (warp-uniform)
            bool cap = false;
            if (WITH_IPM) ipm_dispatch(sh, C, J, a, u, cls, valid, x, &cap);
            st = SG_EV_STATUS_PROJECTED | (cap ? SG_EV_STATUS_MAXITER : 0);
        }
        __syncwarp();
    }
    // schedule (env.py:368-378), charging, Unplug events
    unsigned pil[2] = {0u, 0u};
    float dem_o[2] = {0.f, 0.f}, est_o[2] = {0.f, 0.f};
    double rate_part = 0.0;
    unsigned long long mine = 0ull;
    bool wr[2] = {false, false};                       // q[p] is live after this step (written once the final mask is known)
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        if (!valid[p]) continue;
        const int i = i0 + p;
        const int pi = min(max(__double2int_rn(to_pilot(x[p], type[p])), 0), 32);
        pil[p] = (unsigned)pi;
        if (occd[p]) {
            rate_part += charge_period(q[p].energy, q[p].delivered, sh.ppilot[pi]);
            if (q[p].dep > tn) {
                wr[p] = true;
                mine |= 1ull << i;
                const double rem = q[p].energy - q[p].delivered;
                if (rem > kFullEps) { dem_o[p] = (float)rem; est_o[p] = (float)(q[p].est - tn); }   // Interface.active_sessions()
            }
        }
    }
    const double total_rate = warp_sum(rate_part);
    // Plugin events with timestamp <= t+1; the lane that owns the station takes the event
    if (!autoreset) {
        const unsigned* arow = cx.arrivals + (long long)day * (cx.max_day_sessions + 1);
        unsigned ent = __ldg(arow + nidx);
        while ((int)(ent & 0x7fffu) <= tn) {
            const int i = (ent >> 16) & 0xff, k = ent >> 24;
            if ((i >> 1) == lane) {
                const uint4 rr = __ldg(cx.recs + ((long long)day * cx.K + k) * n + i);
                Slot ns;
                ns.arr = (int)(short)(rr.x & 0xffffu); ns.dep = (int)(short)(rr.x >> 16);
                ns.est = (int)(short)(rr.y & 0xffffu); ns.k = k;
                ns.energy = __hiloint2double((int)rr.w, (int)rr.z); ns.delivered = 0.0;
                if (ns.dep > tn) {
                    mine |= 1ull << i;
                    const int p = i & 1;
                    if (p == 0) { q[0] = ns; wr[0] = true; } else { q[1] = ns; wr[1] = true; }
                    if (ns.energy > kFullEps) { dem_o[p] = (float)ns.energy; est_o[p] = (float)(ns.est - tn); }
                }
            }
            ++nidx;
            ent = __ldg(arow + nidx);
        }
    }
    const unsigned occ_lo = __reduce_or_sync(kFull, (unsigned)(mine & 0xffffffffull));
    const unsigned occ_hi = __reduce_or_sync(kFull, (unsigned)(mine >> 32));
    {   // the packed list of the new occupancy (every record of the old list was read above, by all lanes)
        const unsigned long long occ_fin = ((unsigned long long)occ_hi << 32) | occ_lo;
#pragma unroll
        for (int p = 0; p < 2; ++p)
            if (wr[p]) st_slot_2x16(nxt + slot_rank(occ_fin, i0 + p), q[p]);
    }
    // excess current from the rounded schedule (env.py:449-452)
    double excess = 0.0;
    {
        double S[kMC];
        class_sums_pilots(pil, cls, C, S);
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
        // (constraints beyond lane 31 cannot exist: J <= 16)
    }
    // observation row(s) and pilot row
    const float* mrow = cx.moer + ((long long)md * SG_EV_MOER_ROWS + min(tn, SG_EV_MOER_ROWS - 1)) * SG_EV_MOER_COLS;
    const float ts = timestep_frac(tn);
    float* orow = autoreset ? (cx.final_obs != nullptr ? cx.final_obs + b * OBS : nullptr) : cx.obs + b * OBS;
    if (orow != nullptr) {
#pragma unroll
        for (int p = 0; p < 2; ++p) if (valid[p]) { orow[i0 + p] = dem_o[p]; orow[n + i0 + p] = est_o[p]; }
        for (int e = lane; e < kf + 2; e += 32) orow[2 * n + e] = moer_tail(mrow, e, kf, ts);
    }
    if (autoreset) {                                   // reset observation of the next episode (env.py:293-338)
        float* o2 = cx.obs + b * OBS;
        const float* r0 = cx.moer + (long long)md_out * SG_EV_MOER_ROWS * SG_EV_MOER_COLS;
        for (int e = lane; e < OBS; e += 32) o2[e] = e < 2 * n ? 0.f : moer_tail(r0, e - 2 * n, kf, 0.f);
    }
    if (cx.pilots != nullptr) {
#pragma unroll
        for (int p = 0; p < 2; ++p) if (valid[p]) cx.pilots[b * n + i0 + p] = (uint8_t)pil[p];
    }
    if (lane == 0) {
        const double moer0 = (double)__ldg(mrow);
        const double profit = kProfitFactor * total_rate;
        const double excess_charge = excess * kViolationFactor;
        const double carbon = kCarbonFactor * total_rate * moer0;
        const double rew = profit - carbon - excess_charge;
        EnvHeader ho;
        ho.bd[0] = (t == 0 ? 0.0 : hd.bd[0]) + profit;
        ho.bd[1] = (t == 0 ? 0.0 : hd.bd[1]) + carbon;
        ho.bd[2] = (t == 0 ? 0.0 : hd.bd[2]) + excess_charge;
        ho.t = tn; ho.occ = ((unsigned long long)occ_hi << 32) | occ_lo;
        if (autoreset) { day = (day + 1) % cx.n_days; ho.t = 0; nidx = 0; ho.occ = 0ull; }
        ho.day = day; ho.moer_day = md_out; ho.nidx = nidx; ho.pad0 = ho.pad1 = 0ull;
        cx.header[b] = ho;
        cx.reward[b] = rew;
        cx.done[b] = dn ? 1 : 0;
        if (cx.status != nullptr) cx.status[b] = (uint8_t)st;
        if (wstat != nullptr) {
            wstat[0] += 1.0; wstat[1] += rew; wstat[2] += rew * rew;
            wstat[3] = fmin(wstat[3], rew); wstat[4] = fmax(wstat[4], rew);
            wstat[5] += dn ? 1.0 : 0.0;
            wstat[6] += (st & SG_EV_STATUS_PROJECTED) ? 1.0 : 0.0;
            wstat[7] += (st & SG_EV_STATUS_MAXITER) ? 1.0 : 0.0;
            wstat[8] += profit; wstat[9] += carbon; wstat[10] += excess_charge;
        }
    }
    __syncwarp();
    return true;
}

// Exact projection of ONE env when only single-class constraints (the pods) are violated — the algorithm of
// simple_project above, run by the WHOLE warp on the env's items j in [lo, hi) of the warp's work list (action
// ws.xa[j], upper bound ws.rate[j], class ws.icls[j]); all arguments are warp-uniform.  Lanes stand for classes
// (exact class sums), for break points of the piecewise-linear sum phi (all evaluated at once) and for constraints
// (checks): a thread-serial version of this took 12-17 us on one lane and, with ~60 such envs per 65 536-env step,
// set the end of the launch.  Returns 0: the clipped action is feasible (exact test), nothing to do; 1: projected,
// x_i = clip(a_i - ws.nu[class_i], 0, u_i); 2: a coupled constraint is (or became) violated -> interior point
// (helper warps).
__device__ __forceinline__ int pod_solve(const EvTables& sh, WarpShared& ws, const int lo, const int hi, const int J,
                                         const unsigned suspects, const int lane) {
    // exact class sums of the clipped action: lane c < kMC sums class c over the items (broadcast reads)
    double mine = 0.0;
    for (int j = lo; j < hi; ++j) {
        const double v = fmin(fmax((double)ws.xa[j], 0.0), ws.rate[j]);
        mine += ws.icls[j] == lane ? v : 0.0;
    }
    double Sx[kMC];
#pragma unroll
    for (int c = 0; c < kMC; ++c) Sx[c] = __shfl_sync(kFull, mine, c);
    if (lane < kMC) ws.nu[lane] = 0.0;
    __syncwarp();
    // exact magnitude of constraint jc at the class sums Sx (every lane evaluates the same expression)
    auto magnitude = [&](const int jc) {
        double wr = 0, wi = 0;
#pragma unroll
        for (int c = 0; c < kMC; ++c) { wr = fma(sh.con_re[jc * kMC + c], Sx[c], wr); wi = fma(sh.con_im[jc * kMC + c], Sx[c], wi); }
        return sqrt(wr * wr + wi * wi);
    };
    bool any = false;
    // only the constraints the margin test flagged can be violated by the clipped action
    for (unsigned todo = suspects; todo; todo &= todo - 1) {
        const int jc = __ffs((int)todo) - 1;
        if (jc >= J) break;
        const double r = sh.lim[jc] * (1.0 / kActionScale);
        if (!(magnitude(jc) > r)) continue;                  // satisfied
        const int cs = sh.simple_cls[jc];
        if (cs < 0) return 2;                                // a coupled constraint is violated
        const double R = sh.simple_R[jc];
        // phi(nu) = sum over the class of clip(a_i - nu, 0, u_i): continuous, non-increasing, linear between the
        // break points {a_i - u_i, a_i}, phi(0) > R >= phi(max a_i)
        auto phi = [&](const double nu, double& s_mid, int& n_mid, double& s_up) {
            s_mid = 0.0; s_up = 0.0; n_mid = 0;
            for (int j = lo; j < hi; ++j) {
                const double a = (double)ws.xa[j], ub = ws.rate[j];
                if (ws.icls[j] != cs || !(ub > 0.0)) continue;
                const double v = a - nu;
                if (v >= ub) s_up += ub;
                else if (v > 0.0) { s_mid += a; ++n_mid; }
            }
            return s_mid - n_mid * nu + s_up;
        };
        double lo_n = 0.0, hi_n = 0.0, s_mid, s_up;
        int n_mid;
        for (int j = lo; j < hi; ++j)
            if (ws.icls[j] == cs && ws.rate[j] > 0.0) hi_n = fmax(hi_n, (double)ws.xa[j]);
        // lane l evaluates phi at break point l & 1 of item lo + (l >> 1) (+ 16 per round); the bracket shrinks to
        // the largest break point with phi > R and the smallest with phi <= R
        for (int j0 = lo; j0 < hi; j0 += 16) {
            const int j = j0 + (lane >> 1);
            double bp = -1.0;
            if (j < hi && ws.icls[j] == cs && ws.rate[j] > 0.0) bp = (lane & 1) ? (double)ws.xa[j] : (double)ws.xa[j] - ws.rate[j];
            const bool in = bp > lo_n && bp < hi_n;
            const double v = phi(in ? bp : 0.0, s_mid, n_mid, s_up);
            lo_n = warp_max(in && v > R ? bp : lo_n);
            hi_n = warp_min(in && !(v > R) ? bp : hi_n);
        }
        phi(0.5 * (lo_n + hi_n), s_mid, n_mid, s_up);        // the linear piece that is left
        const double nu = n_mid > 0 ? fmin(fmax((s_mid + s_up - R) / n_mid, lo_n), hi_n) : hi_n;
        if (lane == 0) ws.nu[cs] = nu;
        double sc = 0.0;
        for (int j = lo; j < hi; ++j)
            if (ws.icls[j] == cs) sc += fmin(fmax((double)ws.xa[j] - nu, 0.0), ws.rate[j]);
#pragma unroll
        for (int cc = 0; cc < kMC; ++cc) Sx[cc] = cc == cs ? sc : Sx[cc];
        any = true;
    }
    if (!any) return 0;
    // all constraints again at the projected point (relative slack 1e-12), constraint l on lane l
    bool viol = false;
    if (lane < J) viol = magnitude(lane) > sh.lim[lane] * (1.0 / kActionScale) * (1.0 + 1e-12);
    return __any_sync(kFull, viol) ? 2 : 1;
}

// =================================================================================================
// Main kernel: one launch = one step() of all envs.  A warp owns a tile of kE consecutive envs (tiles are
// dealt by a global ticket counter; a thin last round is cut into 8-env tiles, see sg_ev_step); lane
// e < kE keeps env e's scalars ("owner"), and the phases that have more parallelism than envs are
// spread over all 32 lanes:
//
//   stage    the tile's observation rows and pilot rows (both contiguous in global memory) are
//            assembled in shared memory: zero-filled with 16-byte stores (an empty station reads 0
//            and, with the projection on, has pilot 0), the MOER row + timestep tail copied
//            asynchronously (cp.async) from the observation-tail table; with the projection off the
//            pilot rows are computed densely with their class sums.  At the end of the tile ONE bulk
//            asynchronous copy (cp.async.bulk, the TMA engine) per array writes the finished rows to
//            global memory: full aligned sectors, each byte written once, no store instructions on
//            the SM; its completion is waited for only when the staging rows are needed again;
//   items    the occupied stations of the tile's envs (bits of the per-env 64-bit masks) are
//            enumerated into one work list and processed one (env, station) item per lane, so lanes
//            stay balanced whatever the per-env occupancy: one 32-byte session record (single
//            256-bit load, position = rank of the station among the env's occupied ones) + one
//            action value in, battery arithmetic (branch-free, exact), the two observation entries
//            and the pilot byte patched into the staged rows, class sums of the clipped action /
//            pilots accumulated with shared-memory integer atomics, the rate to a list in shared
//            memory; the records go back into the env's packed list of the NEW occupancy once the
//            checks have passed;
//   rates    kQ lanes per env add the env's rates in a fixed order (deterministic f64 sum);
//   check    owner lane e tests its env: integer screens on the class sums first, the coupled
//            constraints in f32 with a margin only when the total current could violate one (and the
//            same for the excess current of the rounded schedule); an env that is not certainly
//            feasible gets the exact pod projection here (pod_solve) or goes to the helper warps;
//   arrivals Plugin events with timestamp <= t+1 from the day's arrival-sorted list;
//   epilogue reward, cumulative breakdown, clock, mask, statistics.
// Envs that need the interior-point solver are appended to a work list and finished by the helper CTAs of the same
// launch (fallback_env), so that path's register and stack needs do not sit on the tile warps' critical path.
// =================================================================================================
#ifdef SG_EV_TRACE
__device__ unsigned long long* g_trace = nullptr;     // [ntiles][16] phase timestamps (diagnostic builds only)
#define EV_TRACE(k)                                                                                      \
    do {                                                                                                 \
        if (g_trace != nullptr && lane == 0) {                                                           \
            unsigned long long gt;                                                                       \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));                                       \
            g_trace[tile * 16 + (k)] = gt;                                                               \
        }                                                                                                \
    } while (0)
#else
#define EV_TRACE(k) do { } while (0)
#endif

template <int MODE>
__global__ void __launch_bounds__(kMainThreads, SG_EV_MINB)
ev_step_main_kernel(const sg_ev_model m, const sg_ev_state s, const long long B, const void* __restrict__ action,
                    const int flags, const int bins, float* __restrict__ obs, double* __restrict__ reward,
                    uint8_t* __restrict__ done, uint8_t* __restrict__ pilots, float* __restrict__ final_obs,
                    uint8_t* __restrict__ status, double* __restrict__ stats, int32_t* __restrict__ work,
                    const int n_helpers, const long long t_full, const int tail_envs) {
    __shared__ EvTables sh;
    __shared__ WarpShared wsm[kMainWarps];
    extern __shared__ __align__(128) unsigned char stage_raw[];       // per warp: kE observation rows | kE pilot rows
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // Tiling (chosen by sg_ev_step): the first t_full tiles hold kE envs, the envs left after them are cut into tiles
    // of tail_envs <= kE envs.  A tile costs a warp ~10 us of dependent latency plus ~0.25 us per env, so when the
    // last round of kE-env tiles would occupy only part of the warps, the same envs spread over ALL warps in small
    // tiles end the launch earlier.
    const long long tail_first = t_full * kE;                       // first env of the small tiles
    const long long ntiles = t_full + (B > tail_first ? (B - tail_first + tail_envs - 1) / tail_envs : 0);
    auto tile_b0 = [&](const long long t) { return t < t_full ? t * kE : tail_first + (t - t_full) * tail_envs; };
    auto tile_cap = [&](const long long t) { return t < t_full ? kE : tail_envs; };
    const long long wstride = (long long)((int)gridDim.x - n_helpers) * kMainWarps;   // tile warps of the launch
    if (MODE != 0 && (int)blockIdx.x < n_helpers) {
        // ---------------- helper CTAs: the envs the tile warps could not finish (exact projection needed) --------
        // The first n_helpers CTAs of the grid take no tiles: their warps claim entries of the work list as the tile
        // warps append them and run the complete step of those envs (fallback_env: exact test, single-class solve,
        // interior point), WHILE the tile warps go on.  On a tile warp that path ran on cold instruction-cache
        // lines (~12 us per env, measured with the phase trace) and, hitting ~60 tiles of a 65 536-env step, set
        // the end of the launch; here the code stays resident in the few SMs that run it.
        load_tables(m.tables, sh);
        if (lane < 12) wsm[warp].stat[lane] = lane == 3 ? 1e300 : (lane == 4 ? -1e300 : 0.0);
        __syncthreads();
        EvCtx cx;
        cx.n = m.n_stations; cx.kf = m.moer_steps; cx.C = m.n_classes; cx.J = m.n_constraints; cx.K = m.max_sessions;
        cx.n_days = m.n_days; cx.n_moer_days = m.n_moer_days; cx.max_day_sessions = m.max_day_sessions;
        cx.flags = flags; cx.bins = bins; cx.B = B; cx.moer = m.moer; cx.recs = reinterpret_cast<const uint4*>(m.sessions);
        cx.arrivals = m.arrivals; cx.action = action; cx.slots = reinterpret_cast<Slot*>(s.slots);
        cx.header = reinterpret_cast<EnvHeader*>(s.header);
        cx.obs = obs; cx.final_obs = final_obs; cx.reward = reward; cx.done = done; cx.pilots = pilots; cx.status = status;
        volatile int32_t* vw = work;
        for (;;) {
            int k = 0, v = 0;
            if (lane == 0) {
                k = atomicAdd(work + kWorkTaken, 1);
                const bool in_list = k < B;                   // (the list holds at most one entry per env)
                for (int spin = 0;; ++spin) {                 // entry k: written, or never coming
                    v = in_list ? vw[kWorkList + k] : 0;
                    if (v != 0) break;
                    if ((spin & 3) == 3 && vw[kWorkRetired] >= (int)wstride) { v = in_list ? vw[kWorkList + k] : 0; break; }   // (appends precede retirement)
                    __nanosleep(500);
                }
                if (v != 0) vw[kWorkList + k] = 0;            // re-armed for the next step
            }
            v = __shfl_sync(kFull, v, 0);
            if (v == 0) break;
            __threadfence();                                  // the tile warp's rows of this env are complete (see append)
            const long long bq = v - 1;
            double* wst = stats != nullptr ? wsm[warp].stat : nullptr;
#ifdef SG_EV_TRACE
            k = __shfl_sync(kFull, k, 0);
            if (g_trace != nullptr && lane == 0 && k < 4096) {
                unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                g_trace[(ntiles + k) * 16 + 0] = gt; g_trace[(ntiles + k) * 16 + 2] = blockIdx.x * kMainWarps + warp;
            }
#endif
            if (!fallback_env<false>(sh, wsm[warp], cx, bq, lane, wst)) fallback_env<true>(sh, wsm[warp], cx, bq, lane, wst);
#ifdef SG_EV_TRACE
            if (g_trace != nullptr && lane == 0 && k < 4096) {
                unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                g_trace[(ntiles + k) * 16 + 1] = gt;
            }
#endif
        }
        if (lane == 0) {
            const int nhw = n_helpers * kMainWarps;
            if (atomicAdd(work + kWorkHelpersDone, 1) == nhw - 1) {     // last helper warp: every counter back to zero
                work[kWorkCount] = 0; work[kWorkTaken] = 0; work[kWorkNextTile] = 0; work[kWorkRetired] = 0;
                work[kWorkHelpersDone] = 0;
            }
        }
    } else {
    const int cta = (int)blockIdx.x - n_helpers;
    // Tile scheduling.  Every warp starts on the tile of its own index; further tiles come from a global counter
    // (work[kWorkNextTile]) when the scratch buffer is there, else from a grid stride.  Static strides leave the
    // SMs unevenly loaded — B200 places consecutive CTAs on the same SM, so with 2.3 tiles per warp the SMs that
    // hold the low-numbered CTAs run 9 tile rounds against 6 elsewhere (ncu: SMs busy 62 % of the launch).
    int32_t* sched = (work != nullptr && !(flags & SG_EV_FLAG_STATIC_TILES)) ? work + kWorkNextTile : nullptr;
    long long tile = (long long)cta * kMainWarps + warp;
    EV_TRACE(10);
    // the first tile's env headers are requested before the table copy
    EnvHeader* __restrict__ headers = reinterpret_cast<EnvHeader*>(s.header);
    EnvHeader hd = empty_header();
    if (tile < ntiles && tile_b0(tile) + lane < B && lane < tile_cap(tile)) hd = ld_header(headers + tile_b0(tile) + lane);
    bool first_tile = true;
    load_tables(m.tables, sh);
    WarpShared& ws = wsm[warp];
    if (lane < 12) ws.stat[lane] = lane == 3 ? 1e300 : (lane == 4 ? -1e300 : 0.0);
    __syncthreads();

    const int n = m.n_stations, kf = m.moer_steps, OBS = 2 * n + kf + 2, J = m.n_constraints;
    const int eq = lane / kQ, sub = lane % kQ;         // spread phases: lane works for env eq
    const bool i64 = (flags & 1) == SG_EV_ACTION_I64;
    const float* __restrict__ moer = m.moer;
    const uint4* __restrict__ recs = reinterpret_cast<const uint4*>(m.sessions);
    Slot* slots = reinterpret_cast<Slot*>(s.slots);
    const int tp = (kf + 3) >> 1;                      // float2-sized pairs covering a row's MOER + timestep tail
    constexpr unsigned kOwnerMask = kE == 32 ? 0xffffffffu : ((1u << kE) - 1u);
    const int row_bytes = (kE * OBS * 4 + 15) & ~15, pil_bytes = (kE * n + 15) & ~15;
    float* srows = reinterpret_cast<float*>(stage_raw + (size_t)warp * (row_bytes + pil_bytes));
    uint8_t* spil = reinterpret_cast<uint8_t*>(srows) + row_bytes;
    // the bulk copies need 16-byte aligned, 16-byte multiple global ranges
    const bool bulk_rows = ((kE * OBS) & 3) == 0 && (reinterpret_cast<uintptr_t>(obs) & 15u) == 0;
    const bool bulk_pil = pilots != nullptr && ((kE * n) & 15) == 0 && (reinterpret_cast<uintptr_t>(pilots) & 15u) == 0;
    // 8-byte asynchronous copies of the observation tails need an even row length (8-byte aligned tails)
    const int tail_floats = SG_EV_TAIL_FLOATS(kf);
    #ifdef SG_EV_NO_TAIL_ASYNC
    const bool tail_async = false;
#else
    const bool tail_async = m.obs_tail != nullptr && (OBS & 1) == 0 && !(SG_EV_ABL & 512);
#endif

    while (tile < ntiles) {
        // the next tile is requested now and picked up at the end of this one (the atomic's latency is hidden)
        int next_ticket = 0;
        if (sched != nullptr && lane == 0) next_ticket = atomicAdd(sched, 1);
        EV_TRACE(5);
#ifdef SG_EV_TRACE
        if (g_trace != nullptr && lane == 0) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); g_trace[tile * 16 + 1] = sm; }
#endif
        const long long b0 = tile_b0(tile);
        const int nv = (int)min((long long)tile_cap(tile), B - b0);
        const long long b = b0 + lane;
        const bool env = lane < nv;                    // owner lanes
        if (!first_tile) {
            hd = empty_header();
            if (env) hd = ld_header(headers + b);
        }
        first_tile = false;
        const int t = hd.t, tn = t + 1;
        int day = hd.day, nidx = hd.nidx;
        const int md = hd.moer_day;
        const bool live = env && t < kEpisode;
        const unsigned long long occ = live ? hd.occ : 0ull;
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
        EV_TRACE(0);

        // ---------------- stage: zeroed rows, MOER + timestep tails, dense pilots (projection off) -------
        {
            for (int i = lane; i < kMC * kE; i += 32) { (&ws.S[0][0])[i] = 0; (&ws.P[0][0])[i] = 0; }
            if (lane < kE) { ws.occ_clr[lane][0] = 0u; ws.occ_clr[lane][1] = 0u; }
            // the previous tile's bulk copies must have read the staging rows before they are cleared; the wait sits
            // here, behind this tile's header / MOER / arrival loads, instead of at the end of the previous tile
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            uint4* z = reinterpret_cast<uint4*>(srows);
            for (int i = lane; i < (row_bytes + pil_bytes) >> 4; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
            (void)tp;
            if (tail_async) {
                // MOER + timestep tails straight from the observation-tail table into the staged rows: kQ lanes per
                // env issue 8-byte asynchronous copies (no registers, nothing waits on them until the write-out)
                const float* src = m.obs_tail + ((long long)__shfl_sync(kFull, md, eq) * SG_EV_MOER_ROWS +
                                                 min(__shfl_sync(kFull, tn, eq), SG_EV_MOER_ROWS - 1)) * tail_floats;
                const unsigned dst = (unsigned)__cvta_generic_to_shared(srows + eq * OBS + 2 * n);
                if (eq < nv)
                    for (int q2 = sub; q2 < (kf + 2) >> 1; q2 += kQ)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * q2), "l"(src + 2 * q2) : "memory");
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
                        pv[p] = (unsigned)pilot_f32(act_value(action, i64, bins, (b0 + r) * n + i0 + p), sh.type[i0 + p]);
                        spil[r * n + i0 + p] = (uint8_t)pv[p];
                    }
#pragma unroll
                for (int w = 0; w < kMC / 2; ++w) {
                    unsigned part = 0;
#pragma unroll
                    for (int p = 0; p < 2; ++p)
                        if ((cl[p] >> 1) == w) part += pv[p] << ((cl[p] & 1) * 16);
                    const unsigned tot = __reduce_add_sync(kFull, part);
                    if (lane == 0) { ws.P[2 * w][r] = (int)(tot & 0xffffu); ws.P[2 * w + 1][r] = (int)(tot >> 16); }
                }
            }
        }
        __syncwarp();
        EV_TRACE(2);

        // ---------------- items: one (env, occupied station) per lane ------------------------------------
        // The tile's envs are processed in groups of consecutive envs holding at most kItemCap items
        // (usually one group): load + arithmetic for the whole group, then the constraint checks, then
        // the session records go back IN PLACE (their lines are in L2: no write-allocate traffic) — only
        // for envs that passed the check, so the rare path finds the state of the others untouched.
        const int cnt = __popcll(occ);
        int off = cnt;                                                 // inclusive scan over the owner lanes
#pragma unroll
        for (int d = 1; d < kE; d <<= 1) {
            const int v = __shfl_up_sync(kFull, off, d);
            if (lane >= d) off += v;
        }
        off -= cnt;
        double rate_sum = 0.0;
        unsigned bad = 0u, near = 0u;
        // this lane's segment of env eq's occupancy bits, and the list position of its first item
        const unsigned long long occ_e = __shfl_sync(kFull, occ, eq);
        const unsigned long long seg = kQ == 1 ? occ_e : (occ_e >> (sub * kSegBits)) & ((1ull << (kSegBits & 63)) - 1ull);
        const int j0 = __shfl_sync(kFull, off, eq) +
                       (kQ == 1 ? 0 : __popcll(occ_e & ((1ull << ((sub * kSegBits) & 63)) - 1ull)));
        EV_TRACE(11);
#ifdef SG_EV_TRACE_KEYS
        if (j0 == -12345) return;
        EV_TRACE(14);
#endif
        unsigned redo_all = 0u;                                       // envs whose action was projected here (status)
        unsigned long long arr_mask = 0ull;                           // owner lane: stations that get a new session now
        for (int g0 = 0; g0 < kE;) {
            const int base = __shfl_sync(kFull, off, g0);
            const unsigned fit = __ballot_sync(kFull, lane >= g0 && lane < kE && off + cnt - base <= kItemCap);
            int g1 = g0 + __ffs((int)~(fit >> g0)) - 1;                // first env that does not fit any more
            g1 = min(max(g1, g0 + 1), kE);
            const int nitems = (SG_EV_ABL & 4) ? 0 : __shfl_sync(kFull, off + cnt, g1 - 1) - base;
            const unsigned gmask = ((g1 >= 32 ? 0u : (1u << g1)) - 1u) & ~((1u << g0) - 1u);
#ifdef SG_EV_TRACE_KEYS
            if (nitems == -12345) return;
            EV_TRACE(13);
#endif
            if (nitems > 0 && eq >= g0 && eq < g1) {                   // keys of this group
                int j = j0 - base;
                if (kSegBits <= 32) {
                    for (unsigned m2 = (unsigned)seg; m2; m2 &= m2 - 1, ++j)
                        ws.key[j] = (unsigned short)((eq << 8) | (sub * kSegBits + __ffs((int)m2) - 1));
                } else {
                    for (unsigned long long m2 = seg; m2; m2 &= m2 - 1, ++j)
                        ws.key[j] = (unsigned short)((eq << 8) | (__ffsll((long long)m2) - 1));
                }
            }
            EV_TRACE(12);
            __syncwarp();
            EV_TRACE(3);
            int e[kV], si[kV], tne[kV], rk[kV];
            Slot q[kV];
            float af[kV];
            bool on[kV];
#pragma unroll
            for (int u = 0; u < kV; ++u) {
                const int j = u * 32 + lane;
                on[u] = j < nitems;
                const unsigned key = on[u] ? ws.key[j] : 0u;
                e[u] = (int)(key >> 8); si[u] = (int)(key & 0xffu);
                tne[u] = __shfl_sync(kFull, tn, e[u]);
                rk[u] = base + j - __shfl_sync(kFull, off, e[u]);        // items are listed in station order: list rank
            }
#pragma unroll
            for (int u = 0; u < kV; ++u) {
                af[u] = 0.f;
                q[u].energy = 0.0; q[u].delivered = 0.0; q[u].dep = q[u].est = q[u].arr = q[u].k = 0;
                if (on[u]) {
                    const long long be = b0 + e[u];
                    q[u] = ld_slot(slots + be * n + rk[u]);
                    af[u] = act_value(action, i64, bins, be * n + si[u]);
                }
            }
            if (g0 == 0 && live && (int)(ent & 0x7fffu) <= tn) {
                // Plugin events of this step, first pass (their trace-record loads overlap the session-record loads
                // above): which stations get a new session, i.e. the occupancy the packed lists are written for
                unsigned en = ent, en2 = ent2;
                int ni = nidx;
                do {
                    const int i = (en >> 16) & 0xff, k = en >> 24;
                    const unsigned rx = __ldg(reinterpret_cast<const unsigned*>(recs + ((long long)day * m.max_sessions + k) * n + i));
                    if ((int)(short)(rx >> 16) > tn) arr_mask |= 1ull << i;      // (zero-length sessions never charge)
                    ++ni;
                    en = en2;
                    en2 = (int)(en & 0x7fffu) <= tn ? __ldg(arow + ni + 1) : 0x7fffu;
                } while ((int)(en & 0x7fffu) <= tn);
            }
            if (g0 == 0 && !tail_async) {
                // MOER + timestep tails of the staged rows, kQ lanes per env: placed here so that their loads
                // (L2, sometimes DRAM) are in flight together with the session-record loads above
                const float* p = moer + __shfl_sync(kFull, moff, eq);
                const float ts_e = __shfl_sync(kFull, tsf, eq);
                float tv[2 * kTailPerLane];
#pragma unroll
                for (int k = 0; k < kTailPerLane; ++k) {
                    const int q0 = 2 * (sub + k * kQ);
                    tv[2 * k] = (q0 < kf + 2 && !(SG_EV_ABL & 512)) ? moer_tail(p, q0, kf, ts_e) : 0.f;
                    tv[2 * k + 1] = (q0 + 1 < kf + 2 && !(SG_EV_ABL & 512)) ? moer_tail(p, q0 + 1, kf, ts_e) : 0.f;
                }
                float* dst = srows + eq * OBS + 2 * n;
#pragma unroll
                for (int k = 0; k < kTailPerLane; ++k) {
                    const int q0 = 2 * (sub + k * kQ);
                    if (q0 < kf + 2) dst[q0] = tv[2 * k];
                    if (q0 + 1 < kf + 2) dst[q0 + 1] = tv[2 * k + 1];
                }
            }
#if defined(SG_EV_TRACE) && !defined(SG_EV_TRACE_KEYS)
            EV_TRACE(13);
            if (q[0].dep + q[kV - 1].dep + (int)af[0] == -123456) return;   // wait for the loads here
            EV_TRACE(14);
#endif
            bool keep[kV];
            double del0[kV], rt[kV];                                   // delivered before / rate after the first pass
#pragma unroll
            for (int u = 0; u < kV; ++u) {
                keep[u] = false; del0[u] = q[u].delivered; rt[u] = 0.0;
                if (!on[u]) continue;
                const int i = si[u];
                const int type = sh.type[i], cls = sh.cls[i];
                int pi;
                if (MODE == 0) {
                    pi = pilot_f32(af[u], type);
                } else {
                    const double rem = q[u].energy - q[u].delivered;
                    const float dem = rem > kFullEps ? (float)rem : 0.f;                  // obs 'demands', env.py:388
                    float xf;
                    pi = pilot_clipped(af[u], dem, type, &xf);
                    atomicAdd(&ws.S[cls][e[u]], __float2int_rn(xf * 1048576.f));
                    atomicAdd(&ws.P[cls][e[u]], pi);
                    spil[e[u] * n + i] = (uint8_t)pi;
                }
                rt[u] = charge_period(q[u].energy, q[u].delivered, sh.ppilot[pi]);
                ws.rate[u * 32 + lane] = rt[u];
                if (q[u].dep <= tne[u]) {                                                 // Unplug event (timestamp = departure)
                    atomicOr(&ws.occ_clr[e[u]][i >> 5], 1u << (i & 31));
                } else {
                    keep[u] = true;
                    const double rem = q[u].energy - q[u].delivered;
                    if (rem > kFullEps) {                                                 // Interface.active_sessions()
                        srows[e[u] * OBS + i] = (float)rem;
                        srows[e[u] * OBS + n + i] = (float)(q[u].est - tne[u]);
                    }
                }
            }
            __syncwarp();
            EV_TRACE(4);
            // network constraints: one (env, constraint) pair per lane and iteration.  bad: the clipped action
            // is not certainly feasible (projection needed); near: the rounded schedule is close to or over
            // a limit (exact excess current needed)
            unsigned gbad = 0u, gnear = 0u;
            unsigned jsus = 0u;                                         // owner lane: constraints its env may violate
            // Owner lane e tests its env: integer screens first (a few compares: the pods' class sums against their
            // limits, the total current against the tightest coupled limit), the coupled constraints in f32 with a
            // margin only when the total is large enough for one of them to be violated at all.
            auto pair_checks = [&]() {
                bool mybad = false, mynear = false;
                jsus = 0u;
                if (lane < kE && !(SG_EV_ABL & 16)) {
                    int St = 0, Pt = 0;
#pragma unroll
                    for (int c = 0; c < kMC; ++c) { St += ws.S[c][lane]; Pt += ws.P[c][lane]; }
                    for (int q = 0; q < sh.n_simple; ++q) {
                        const int j = sh.simple_j[q], c = sh.simple_cls[j];
                        if (MODE != 0 && ws.S[c][lane] > sh.simple_S_fix[q]) { mybad = true; jsus |= 1u << j; }
                        mynear = mynear || ws.P[c][lane] > sh.simple_P_lim[q];
                    }
                    if ((MODE != 0 && St > sh.scr_S_fix) || Pt > sh.scr_P_lim) {
                        for (int j = 0; j < J; ++j) {
                            if (sh.simple_cls[j] >= 0) continue;
                            float wr = 0.f, wi = 0.f, vr = 0.f, vi = 0.f;
#pragma unroll
                            for (int c = 0; c < kMC; ++c) {
                                const float cre = sh.re[j * kMC + c], cim = sh.im[j * kMC + c];
                                const float pv = (float)ws.P[c][lane];
                                vr = fmaf(cre, pv, vr); vi = fmaf(cim, pv, vi);
                                if (MODE != 0) { const float sv = (float)ws.S[c][lane] * (1.f / 1048576.f); wr = fmaf(cre, sv, wr); wi = fmaf(cim, sv, wi); }
                            }
                            if (MODE != 0 && !(fmaf(wr, wr, wi * wi) <= sh.r_in[j])) { mybad = true; jsus |= 1u << j; }
                            mynear = mynear || (fmaf(vr, vr, vi * vi) > sh.lim2_lo[j]);
                        }
                    }
                }
                gbad = __ballot_sync(kFull, mybad) & gmask & live_mask;
                gnear = __ballot_sync(kFull, mynear) & gmask;
            };
            pair_checks();
            // ---- envs whose clipped action is not certainly feasible (rare): the exact test and the exact
            // single-class (pod) projection run here, one env at a time on the whole warp; envs that need the
            // interior-point solver go to the helper warps' work list (gfail)
            unsigned gfail = 0u;
            if (!SG_EV_INLINE_POD) {
                gfail = MODE != 0 ? gbad : 0u;        // every env that is not certainly feasible goes to the helper warps
            } else if (MODE != 0 && gbad) {
                unsigned gredo = 0u;
                for (unsigned todo = gbad; todo; todo &= todo - 1) {
                    const int r = __ffs((int)todo) - 1;
#pragma unroll
                    for (int u = 0; u < kV; ++u)
                        if (on[u] && e[u] == r) {                        // publish (action, upper bound) of env r's items
                            const double rem = q[u].energy - del0[u];
                            ws.rate[u * 32 + lane] = upper_bound(rem > kFullEps ? (float)rem : 0.f);
                            ws.xa[u * 32 + lane] = af[u];
                            ws.icls[u * 32 + lane] = (unsigned char)sh.cls[si[u]];
                        }
                    __syncwarp();
                    const int lo_r = __shfl_sync(kFull, off, r) - base;
                    const int res = pod_solve(sh, ws, lo_r, lo_r + __shfl_sync(kFull, cnt, r), J, __shfl_sync(kFull, jsus, r), lane);
                    if (res == 1 && lane < kMC) ws.P[lane][r] = 0;
                    __syncwarp();
                    if (res == 2) gfail |= 1u << r;
                    if (res == 1) gredo |= 1u << r;
#pragma unroll
                    for (int u = 0; u < kV; ++u) {
                        if (!(on[u] && e[u] == r)) continue;
                        if (res != 1) { ws.rate[u * 32 + lane] = rt[u]; continue; }   // first-pass results stand
                        // re-do the item with x = clip(a - nu_class, 0, u)  (KKT point of the projection)
                        const int i = si[u];
                        const int type = sh.type[i], cls = sh.cls[i];
                        const double rem0 = q[u].energy - del0[u];
                        const double ub = upper_bound(rem0 > kFullEps ? (float)rem0 : 0.f);
                        const double x = fmin(fmax((double)af[u] - ws.nu[cls], 0.0), ub);
                        const int pi = min(max(__double2int_rn(to_pilot(x, type)), 0), 32);
                        atomicAdd(&ws.P[cls][r], pi);
                        spil[r * n + i] = (uint8_t)pi;
                        q[u].delivered = del0[u];
                        ws.rate[u * 32 + lane] = charge_period(q[u].energy, q[u].delivered, sh.ppilot[pi]);
                        if (keep[u]) {
                            const double rem = q[u].energy - q[u].delivered;
                            const bool active = rem > kFullEps;
                            srows[r * OBS + i] = active ? (float)rem : 0.f;
                            srows[r * OBS + n + i] = active ? (float)(q[u].est - tne[u]) : 0.f;
                        }
                    }
                    __syncwarp();
                }
                if (gredo) {                                            // pilots changed: excess-current pre-check again
                    const unsigned keep_bad = gbad;
                    pair_checks();
                    gbad = keep_bad;
                    redo_all |= gredo;
                }
            }
            bad |= gfail;
            near |= gnear;
            // owner: lane e sums the rates of env e in station order
#if SG_EV_RATESUM_SPLIT
            // kQ lanes per env: lane (eq, sub) adds the items sub, sub + kQ, .. of env eq's list, the owner lane then
            // adds the kQ partial sums in order (a fixed summation order: results do not depend on timing)
            if (nitems > 0) {
                const int c_e = __shfl_sync(kFull, cnt, eq), lo = __shfl_sync(kFull, off, eq) - base;
                const bool mine = eq >= g0 && eq < g1;
                const int hi = mine ? lo + c_e : lo;
                const int rounds = (__reduce_max_sync(kFull, (unsigned)(mine ? c_e : 0)) + kQ - 1) / kQ;
                double part = 0.0;
                for (int r = 0, j = lo + sub; r < rounds; ++r, j += kQ)
                    if (j < hi) part += ws.rate[j];
#pragma unroll
                for (int q2 = 0; q2 < kQ; ++q2) {
                    const double v = __shfl_sync(kFull, part, (lane * kQ + q2) & 31);
                    if (lane >= g0 && lane < g1) rate_sum += v;
                }
            }
#else
            if (nitems > 0) {
                const bool mine = lane >= g0 && lane < g1;
                const int lo = off - base, hi = mine ? lo + cnt : lo;
                for (int j = lo; __any_sync(kFull, j < hi); ++j)
                    if (j < hi) rate_sum += ws.rate[j];
            }
#endif
            // session records back into the env's packed list of the NEW occupancy (Unplug events of the items
            // phase removed, this step's Plugin events included), for the envs of the group that are finished here;
            // every record of these envs was read at the top of the group
            if (lane >= g0 && lane < g1)
                ws.occ_fin[lane] = (occ & ~(((unsigned long long)ws.occ_clr[lane][1] << 32) | ws.occ_clr[lane][0])) | arr_mask;
            __syncwarp();
#pragma unroll
            for (int u = 0; u < kV; ++u)
                if (keep[u] && !((gfail >> e[u]) & 1u) && !(SG_EV_ABL & 2))
                    st_slot(slots + (b0 + e[u]) * n + slot_rank(ws.occ_fin[e[u]], si[u]), q[u]);
            __syncwarp();
            g0 = g1;
        }
        EV_TRACE(6);

        const bool mine_ok = live && !((bad >> lane) & 1u);              // this owner lane finishes its env here
        EV_TRACE(6);

        // ---------------- the next tile: pick up the ticket and pull its headers, the head of its session lists and its
        // action rows into L2 while this tile finishes (checks, arrivals, epilogue, write-out).  The lists are packed, so
        // the first two lines of an env's block hold its first 6-8 sessions — all of them for a typical garage.
        const long long tile_next = sched != nullptr ? wstride + (long long)__shfl_sync(kFull, next_ticket, 0) : tile + wstride;
        if (tile_next < ntiles && tile_b0(tile_next) + kE <= B && !(SG_EV_ABL & 1024)) {
            const long long bn = tile_b0(tile_next);
            const unsigned char* hp = reinterpret_cast<const unsigned char*>(headers + bn) + lane * 128;
            if (lane < (kE * 64 + 127) / 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(hp));
            const int eper = lane / (32 / kE), part = lane % (32 / kE);            // 32 / kE lines per env
            const unsigned char* sp = reinterpret_cast<const unsigned char*>(slots + (bn + eper) * n) + part * 128;
            // a line holds ~4 records: as many lines as the env in the same position of THIS tile would need (the envs
            // of a batch share the time of day, so its occupancy predicts the next tile's), at least one
            if (4 * part <= __shfl_sync(kFull, cnt, eper))
                asm volatile("prefetch.global.L2 [%0];" ::"l"(sp));
            const long long abytes = (long long)kE * n * (i64 ? 8 : 4);
            const unsigned char* ap = reinterpret_cast<const unsigned char*>(action) + bn * n * (i64 ? 8 : 4);
            for (long long o = (long long)lane * 128; o < abytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(ap + o));
        }

        // ---------------- arrivals: Plugin events with timestamp <= t+1 (after the Unplug events) -------
        unsigned long long occ_new = (occ & ~(((unsigned long long)ws.occ_clr[lane & (kE - 1)][1] << 32) |
                                              ws.occ_clr[lane & (kE - 1)][0])) | arr_mask;
        if (mine_ok && (int)(ent & 0x7fffu) <= tn) {
            float* row = srows + lane * OBS;
            Slot* nxt = slots + b * n;
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
                    st_slot(nxt + slot_rank(occ_new, i), ns);
                    if (ns.energy > kFullEps) { row[i] = (float)ns.energy; row[n + i] = (float)(ns.est - tn); }
                }
                ++nidx;
                ent = ent2;
                ent2 = (int)(ent & 0x7fffu) <= tn ? __ldg(arow + nidx + 1) : 0x7fffu;
            } while ((int)(ent & 0x7fffu) <= tn);
        }
        EV_TRACE(7);

        // ---------------- per-env epilogue: excess current, reward, clock ---------------------------------
        double rew = 0.0, profit = 0.0, carbon = 0.0, excess_charge = 0.0;
        if (mine_ok) {
            double excess = 0.0;
            if ((near >> lane) & 1u) {                                    // exact (env.py:449-452)
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
            profit = kProfitFactor * rate_sum;
            excess_charge = excess * kViolationFactor;
            carbon = kCarbonFactor * rate_sum * (double)moer0f;
            rew = profit - carbon - excess_charge;
            // cumulative reward breakdown (env.py:460-462); restarted on the first step of an episode so that
            // after an auto-reset the finished episode's totals stay readable until the next step
            EnvHeader ho;
            ho.bd[0] = (t == 0 ? 0.0 : hd.bd[0]) + profit;
            ho.bd[1] = (t == 0 ? 0.0 : hd.bd[1]) + carbon;
            ho.bd[2] = (t == 0 ? 0.0 : hd.bd[2]) + excess_charge;
            ho.t = tn;
            if (autoreset) {
                day = (day + 1) % m.n_days;                               // sequential days, event_generation.py:284-291
                ho.t = 0; nidx = 0; occ_new = 0ull;
            }
            ho.day = day; ho.moer_day = md_out; ho.nidx = nidx; ho.occ = occ_new; ho.pad0 = ho.pad1 = 0ull;
            if (!(SG_EV_ABL & 128)) st_header(headers + b, ho);           // in place: the line is in L2
            reward[b] = rew;
            done[b] = dn ? 1 : 0;
            if (status != nullptr) status[b] = (uint8_t)(((redo_all >> lane) & 1u) ? SG_EV_STATUS_PROJECTED : 0);
        } else if (env && !live) {                                        // finished episode, not reset: stays terminal
            reward[b] = 0.0; done[b] = 1;
            if (status != nullptr) status[b] = 0;
        }
        if (stats != nullptr) {                                           // warp-level reduction into this warp's accumulators
            const unsigned okm = __ballot_sync(kFull, mine_ok);
            const unsigned dnm = __ballot_sync(kFull, mine_ok && dn);
            if (okm) {
                const double v[5] = {rew, rew * rew, profit, carbon, excess_charge};
                double r5[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) r5[k] = warp_sum(v[k]);
                const double mn = warp_min(mine_ok ? rew : 1e300), mx = warp_max(mine_ok ? rew : -1e300);
                if (lane == 0) {
                    ws.stat[0] += (double)__popc(okm); ws.stat[1] += r5[0]; ws.stat[2] += r5[1];
                    ws.stat[3] = fmin(ws.stat[3], mn); ws.stat[4] = fmax(ws.stat[4], mx);
                    ws.stat[5] += (double)__popc(dnm);
                    ws.stat[6] += (double)__popc(redo_all & okm);
                    ws.stat[8] += r5[2]; ws.stat[9] += r5[3]; ws.stat[10] += r5[4];
                }
            }
        }
        __syncwarp();
        EV_TRACE(8);

        // ---------------- write-out of the staged rows --------------------------------------------------------
        if (tail_async) asm volatile("cp.async.wait_all;" ::: "memory");  // the tails have landed in the staged rows
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // staged rows -> visible to the bulk-copy engine
        __syncwarp();
        // every env of the tile live and none at the end of its day: the tile's rows go out as ONE bulk copy — also for
        // a short tile, if its byte range is 16-byte aligned and a multiple of 16 bytes (even env counts)
        const unsigned nv_mask = nv >= 32 ? 0xffffffffu : ((1u << nv) - 1u);
        const bool whole = nv > 0 && live_mask == nv_mask && ar_mask == 0u;
        const bool rows_ok = (reinterpret_cast<uintptr_t>(obs) & 15u) == 0 && (((long long)nv * OBS) & 3) == 0 && ((b0 * OBS) & 3) == 0;
        const bool pil_ok = pilots != nullptr && (reinterpret_cast<uintptr_t>(pilots) & 15u) == 0 && ((nv * n) & 15) == 0 && ((b0 * n) & 15) == 0;
        (void)bulk_rows; (void)bulk_pil;
        if (SG_EV_ABL & 256) {                          // (ablation: no observation / pilot rows written)
        } else
        if (whole && rows_ok) {
            if (lane == 0)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(obs + b0 * OBS),
                             "r"((unsigned)__cvta_generic_to_shared(srows)), "r"(nv * OBS * 4) : "memory");
        } else {
            for (int r = 0; r < nv; ++r) {
                if (!((live_mask >> r) & 1u)) continue;
                const bool ar = (ar_mask >> r) & 1u;
                float* orow = ar ? (final_obs != nullptr ? final_obs + (b0 + r) * OBS : nullptr) : obs + (b0 + r) * OBS;
                if (orow != nullptr)
                    for (int e2 = lane; e2 < OBS; e2 += 32) orow[e2] = srows[r * OBS + e2];
                if (ar) {                                  // reset observation of the next episode (env.py:293-338)
                    const int mdo_r = __shfl_sync(kFull, md_out, r);
                    const float* r0 = moer + (long long)mdo_r * SG_EV_MOER_ROWS * SG_EV_MOER_COLS;
                    float* o2 = obs + (b0 + r) * OBS;
                    for (int e2 = lane; e2 < OBS; e2 += 32) o2[e2] = e2 < 2 * n ? 0.f : moer_tail(r0, e2 - 2 * n, kf, 0.f);
                }
            }
        }
        if (pilots != nullptr && !(SG_EV_ABL & 256)) {
            if (whole && pil_ok) {
                if (lane == 0)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(pilots + b0 * n),
                                 "r"((unsigned)__cvta_generic_to_shared(spil)), "r"(nv * n) : "memory");
            } else {
                for (int r = 0; r < nv; ++r)
                    if ((live_mask >> r) & 1u)
                        for (int i = lane; i < n; i += 32) pilots[(b0 + r) * n + i] = spil[r * n + i];
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");   // (waited for before the rows are cleared again)
        __syncwarp();
        // ---------------- rare path: envs that need the exact projection go to the helper warps' work list — once this
        // tile's rows (which hold first-pass values for those envs) are in global memory, so that the helper's rows
        // are the ones that stay
        if (bad) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __syncwarp();
            __threadfence();
            if (lane == 0) {
                const int first = atomicAdd(work + kWorkCount, __popc(bad));
                int k = 0;
                for (unsigned todo = bad; todo; todo &= todo - 1, ++k)
                    atomicExch(work + kWorkList + first + k, (int)(b0 + __ffs((int)todo) - 1) + 1);
            }
        }
        EV_TRACE(15);
        tile = tile_next;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging rows stay valid until read
    if (work != nullptr && lane == 0) {
        __threadfence();                                  // this warp's appends are visible before its retirement
        // without helper CTAs the last warp to leave re-arms the scheduler (every other warp has made its last
        // request); with helpers the last helper warp re-arms the whole buffer
        if (atomicAdd(work + kWorkRetired, 1) == (int)wstride - 1 && n_helpers == 0) {
            work[kWorkNextTile] = 0;
            work[kWorkRetired] = 0;
        }
    }
    }   // tile CTAs
    if (stats != nullptr) {
        __syncthreads();
        // 11 statistics x kMainWarps partials -> one set of atomics per CTA
        if (threadIdx.x < 11) {
            const int i = threadIdx.x;
            double r = wsm[0].stat[i];
            for (int w = 1; w < kMainWarps; ++w) {
                const double o = wsm[w].stat[i];
                r = (i == 3) ? fmin(r, o) : (i == 4) ? fmax(r, o) : r + o;
            }
            if (i == 3) { if (r < 1e300) atomic_min_double(stats + 3, r); }
            else if (i == 4) { if (r > -1e300) atomic_max_double(stats + 4, r); }
            else if (r != 0.0) atomicAdd(stats + i, r);
        }
    }
#ifdef SG_EV_TRACE
    {
        long long tile = (long long)((int)blockIdx.x - n_helpers) * kMainWarps + warp;
        if (tile >= 0 && tile < ntiles) EV_TRACE(9);
    }
#endif
}

__global__ void ev_reset_kernel(const sg_ev_model m, const sg_ev_state s, const long long B,
                                const uint8_t* __restrict__ mask, const int32_t* __restrict__ day_in,
                                const int32_t* __restrict__ moer_day_in, float* __restrict__ obs) {
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    if (mask != nullptr && mask[b] == 0) return;
    const int n = m.n_stations, kf = m.moer_steps, OBS = 2 * n + kf + 2;
    EnvHeader* headers = reinterpret_cast<EnvHeader*>(s.header);
    int day = day_in ? day_in[b] : headers[b].day;
    int md = moer_day_in ? moer_day_in[b] : headers[b].moer_day;
    day = ((day % m.n_days) + m.n_days) % m.n_days;
    md = ((md % m.n_moer_days) + m.n_moer_days) % m.n_moer_days;
    float* orow = obs + b * OBS;
    for (int i = lane; i < 2 * n; i += 32) orow[i] = 0.f;
    const float* mrow = m.moer + ((long long)md * SG_EV_MOER_ROWS) * SG_EV_MOER_COLS;
    for (int e = lane; e < kf + 2; e += 32) orow[2 * n + e] = e < kf ? mrow[1 + e] : (e == kf ? mrow[0] : 0.f);
    __syncwarp();
    if (lane == 0) {
        // empty garage: no occupied station, so no session record is live (the records need no clearing)
        EnvHeader h = empty_header();
        h.t = 0; h.day = day; h.moer_day = md;
        st_header(headers + b, h);
    }
}

// Dense per-station views of the sparse state (debugging, checkpoint inspection, multi-agent info):
// delivered kWh and trace slot of the session plugged in at each station, 0 / 255 where empty.
__global__ void ev_export_kernel(const sg_ev_model m, const sg_ev_state s, const long long B,
                                 double* __restrict__ delivered, uint8_t* __restrict__ sess_idx) {
    const int n = m.n_stations;
    const long long total = B * n;
    const Slot* slots = reinterpret_cast<const Slot*>(s.slots);
    const EnvHeader* headers = reinterpret_cast<const EnvHeader*>(s.header);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / n;
        const int i = (int)(idx % n);
        const bool on = (headers[b].occ >> i) & 1ull;
        double d = 0.0;
        int k = 255;
        if (on) { const Slot q = ld_slot(slots + b * n + slot_rank(headers[b].occ, i)); d = q.delivered; k = q.k; }
        if (delivered != nullptr) delivered[idx] = d;
        if (sess_idx != nullptr) sess_idx[idx] = (uint8_t)k;
    }
}

// Stand-alone projection operator (env.py:200-221): one warp per problem.
__global__ void __launch_bounds__(128)
ev_project_kernel(const sg_ev_model m, const long long B, const float* __restrict__ action,
                  const float* __restrict__ demands, double* __restrict__ x_out, uint8_t* __restrict__ status,
                  int32_t* __restrict__ iters) {
    __shared__ EvTables sh;
    load_tables(m.tables, sh);
    __syncthreads();
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


// Per-agent observations of MultiAgentEVChargingEnv as DOCUMENTED (multiagent_env.py:102-148): agent i sees the other
// stations' est_departure and demand entries of `periods_delay` steps ago; its own station's entries, the MOER
// fields and the timestep are current.  out[b, i, :] for all 54 agents.  (What the reference actually returns is
// the current observation for everybody — its deque aliases the live arrays; see envs/evcharging/multiagent_env.py.)
__global__ void ev_agent_obs_kernel(const float* __restrict__ cur, const float* __restrict__ past, const long long B,
                                    const int n, const int obs_dim, float* __restrict__ out) {
    const long long total = B * n * obs_dim;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(idx % obs_dim);
        const long long bi = idx / obs_dim;
        const int i = (int)(bi % n);
        const long long b = bi / n;
        const bool own = (e == i) || (e == n + i) || e >= 2 * n;     // only demands / est_departures are delayed (:134-139)
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

int check_tables(const sg_ev_model* m) {
    SG_CHECK_ARG(m->tables != nullptr, "ev: model.tables is NULL (call sg_ev_prepare_tables once per model)");
    SG_CHECK_ARG((reinterpret_cast<uintptr_t>(m->tables) & 15u) == 0, "ev: model.tables must be 16-byte aligned");
    SG_CHECK_ARG((reinterpret_cast<uintptr_t>(m->obs_tail) & 7u) == 0, "ev: model.obs_tail must be 8-byte aligned");
    return 0;
}

int check_state(const sg_ev_state* s) {
    SG_CHECK_ARG(s != nullptr && s->slots && s->header, "ev: NULL state array");
    SG_CHECK_ARG((reinterpret_cast<uintptr_t>(s->slots) & 31u) == 0, "ev: state.slots must be 32-byte aligned");
    SG_CHECK_ARG((reinterpret_cast<uintptr_t>(s->header) & 63u) == 0, "ev: state.header must be 64-byte aligned");
    return 0;
}

}  // namespace
}  // namespace sg

extern "C" {

int sg_ev_prepare_tables(const sg_ev_model* m, void* tables, void* stream) {
    if (int rc = sg::check_model(m)) return rc;
    SG_CHECK_ARG(tables != nullptr && (reinterpret_cast<uintptr_t>(tables) & 15u) == 0, "ev_prepare_tables: tables must be a 16-byte aligned device buffer");
    sg::ev_prepare_kernel<<<1, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*m, reinterpret_cast<sg::EvTables*>(tables));
    return sg::launch_status("ev_prepare_kernel");
}

int sg_ev_prepare_obs_tail(const sg_ev_model* m, float* obs_tail, void* stream) {
    if (int rc = sg::check_model(m)) return rc;
    SG_CHECK_ARG(obs_tail != nullptr && m->moer != nullptr && m->n_moer_days >= 1, "ev_prepare_obs_tail: NULL argument / empty MOER pool");
    SG_CHECK_ARG((reinterpret_cast<uintptr_t>(obs_tail) & 7u) == 0, "ev_prepare_obs_tail: the table must be 8-byte aligned");
    const long long total = (long long)m->n_moer_days * SG_EV_MOER_ROWS * SG_EV_TAIL_FLOATS(m->moer_steps);
    const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148ll * 8);
    sg::ev_prepare_tail_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(m->moer, m->n_moer_days, m->moer_steps, obs_tail);
    return sg::launch_status("ev_prepare_tail_kernel");
}

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
               float* final_obs, uint8_t* status, double* stats, int32_t* work, void* stream) {
    if (int rc = sg::check_model(m)) return rc;
    if (int rc = sg::check_state(s)) return rc;
    SG_CHECK_ARG(action && obs && reward && done, "ev_step: NULL argument");
    SG_CHECK_ARG(m->sessions && m->moer && m->arrivals && m->n_days >= 1 && m->n_moer_days >= 1 && m->max_sessions >= 1,
                 "ev_step: empty trace/MOER pool");
    SG_CHECK_ARG(m->max_sessions <= 256, "ev_step: more than 256 session slots per station per day");
    if (int rc = sg::check_tables(m)) return rc;
    SG_CHECK_ARG(2 * num_envs * (int64_t)m->n_stations < (1ll << 40), "ev_step: batch too large");
    SG_CHECK_ARG(!((flags & 1) == SG_EV_ACTION_I64) || discrete_bins >= 2, "ev_step: discrete_bins must be >= 2");
    SG_CHECK_ARG(!m->project || work != nullptr, "ev_step: the projection needs the work-list scratch buffer");
    if (num_envs <= 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int sms = sg_device_sm_count();
    if (sms <= 0) sms = 148;
    // one warp per kE-env tile; SG_EV_MINB resident CTAs per SM, larger batches loop (persistent).  With the
    // projection on, the first n_helpers CTAs of the grid are helpers (exact projection of the envs the tile warps
    // hand over, see the kernel); they come out of the same residency budget so that every CTA is resident at once
    long long ntiles = (num_envs + sg::kE - 1) / sg::kE;
    const long long want = (ntiles + sg::kMainWarps - 1) / sg::kMainWarps;
    const long long resident = (long long)sms * SG_EV_MINB;
    int n_helpers = 0;
    if (m->project) n_helpers = (int)std::max<long long>(1, std::min<long long>(SG_EV_HELPER_CTAS, want / 64));
    unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(want, resident - n_helpers));
    // SG_EV_MAX_CTAS=<n> caps the tile CTAs (test hook): every warp then loops over many tiles, which is the code
    // path the BASELINE batch runs on, at any batch size
    if (const char* cap = getenv("SG_EV_MAX_CTAS")) {
        const long v = atol(cap);
        if (v >= 1) grid = (unsigned)std::min<long long>(grid, v);
    }
    // Tiling.  W tile warps take the kE-env tiles in rounds; when the last round would be less than 3/4 full, its envs
    // are cut into tiles of 8 instead if those fit one round (8 envs: observation AND pilot rows of a tile stay 16-byte
    // aligned multiples of 16 bytes, so they still leave as bulk copies).  Measured at 65 536 envs (3 520 + 576 tiles on
    // 1 760 warps): 50.9 us uniform, 49.8 us with the last 9 216 envs in 8-env tiles, 51.3 us in 6-env tiles (pilot rows
    // by plain stores), 55.0 / 59.1 us with 8- / 12-env tiles throughout (profiles/r2/tiling_probe.sh).
    long long t_full = ntiles;
    int tail_envs = sg::kE;
    {
        const long long W = (long long)grid * sg::kMainWarps, full = num_envs / sg::kE;
        const long long rounds = (ntiles + W - 1) / W, last = ntiles - (rounds - 1) * W;
        if (rounds >= 2 && 4 * last < 3 * W && (rounds - 1) * W <= full) {
            t_full = (rounds - 1) * W;
            const long long rem = num_envs - t_full * sg::kE;
            tail_envs = (int)std::min<long long>(sg::kE, 8 * ((rem + 8 * W - 1) / (8 * W)));
            if (tail_envs >= sg::kE) t_full = ntiles;
        }
        if (const char* ov = getenv("SG_EV_TILING")) {            // "t_full,tail_envs" (diagnostic override)
            long long a = 0; int c = 0;
            if (sscanf(ov, "%lld,%d", &a, &c) == 2 && a >= 0 && a * sg::kE <= num_envs && c >= 1 && c <= sg::kE) { t_full = a; tail_envs = c; }
        }
    }
    grid += (unsigned)n_helpers;
    // staging rows in dynamic shared memory: per warp kE observation rows + kE pilot rows
    const int OBS = 2 * m->n_stations + m->moer_steps + 2;
    const size_t dyn = (size_t)sg::kMainWarps * (((sg::kE * OBS * 4 + 15) & ~15) + ((sg::kE * m->n_stations + 15) & ~15));
    if (m->project) {
        auto kern = sg::ev_step_main_kernel<1>;
        SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        kern<<<grid, sg::kMainThreads, dyn, st>>>(*m, *s, num_envs, action, flags, discrete_bins, obs, reward, done, pilots,
                                                  final_obs, status, stats, work, n_helpers, t_full, tail_envs);
    } else {
        auto kern = sg::ev_step_main_kernel<0>;
        SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        kern<<<grid, sg::kMainThreads, dyn, st>>>(*m, *s, num_envs, action, flags, discrete_bins, obs, reward, done, pilots,
                                                  final_obs, status, stats, work, 0, t_full, tail_envs);
    }
    return sg::launch_status("ev_step_main_kernel");
}

static int ev_step_host_impl(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs, const void* h_action,
                             int32_t flags, int32_t discrete_bins, float* h_obs, double* h_reward, uint8_t* h_done,
                             void* d_action, float* d_obs, double* d_reward, uint8_t* d_done, uint8_t* d_status,
                             double* d_stats, int32_t* d_work, void* stream, bool sync) {
    SG_CHECK_ARG(m && h_action && h_obs && h_reward && h_done && d_action && d_obs && d_reward && d_done,
                 "ev_step_host: NULL argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t B = (size_t)num_envs, n = (size_t)m->n_stations, OBS = 2 * n + m->moer_steps + 2;
    const size_t abytes = B * n * (((flags & 1) == SG_EV_ACTION_I64) ? 8 : 4);
    SG_CUDA(cudaMemcpyAsync(d_action, h_action, abytes, cudaMemcpyHostToDevice, st));
    if (int rc = sg_ev_step(m, s, num_envs, d_action, flags, discrete_bins, d_obs, d_reward, d_done, nullptr,
                            nullptr, d_status, d_stats, d_work, stream)) return rc;
    SG_CUDA(cudaMemcpyAsync(h_obs, d_obs, B * OBS * 4, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(h_reward, d_reward, B * 8, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(h_done, d_done, B, cudaMemcpyDeviceToHost, st));
    if (sync) SG_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int sg_ev_step_host(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs, const void* h_action,
                    int32_t flags, int32_t discrete_bins, float* h_obs, double* h_reward, uint8_t* h_done,
                    void* d_action, float* d_obs, double* d_reward, uint8_t* d_done, uint8_t* d_status,
                    double* d_stats, int32_t* d_work, void* stream) {
    return ev_step_host_impl(m, s, num_envs, h_action, flags, discrete_bins, h_obs, h_reward, h_done, d_action, d_obs,
                             d_reward, d_done, d_status, d_stats, d_work, stream, true);
}

int sg_ev_step_host_async(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs, const void* h_action,
                          int32_t flags, int32_t discrete_bins, float* h_obs, double* h_reward, uint8_t* h_done,
                          void* d_action, float* d_obs, double* d_reward, uint8_t* d_done, uint8_t* d_status,
                          double* d_stats, int32_t* d_work, void* stream) {
    return ev_step_host_impl(m, s, num_envs, h_action, flags, discrete_bins, h_obs, h_reward, h_done, d_action, d_obs,
                             d_reward, d_done, d_status, d_stats, d_work, stream, false);
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

#ifdef SG_EV_TRACE
int sg_ev_set_trace(unsigned long long* buf) {
    return cudaMemcpyToSymbol(sg::g_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : 1;
}
#endif

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
    if (int rc = sg::check_tables(m)) return rc;
    if (num_envs <= 0) return 0;
    const unsigned grid = (unsigned)((num_envs + 3) / 4);
    sg::ev_project_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*m, num_envs, action, demands,
                                                                                  x_out, status, iters);
    return sg::launch_status("ev_project_kernel");
}

}  // extern "C"
