// BuildingEnv batched step()/reset() for sm_100a.
//
// Replaces sustaingym/envs/building/env.py:205-312 (step), :314-386 (reset), :411-434
// (_calc_occupower) of the reference.  One thread advances one env instance; a CTA (one warp) owns
// a tile of 32 consecutive envs.  Fused rollouts (steps > 1) are software-pipelined: the action tiles of the
// next 3 steps are in flight (cp.async ring in shared memory) and the weather rows of the next 2 steps
// are prefetched into registers, so a step's critical path is its arithmetic, not a DRAM round trip.
// The env-major rows the callers hand us (action [B,n], state
// [B,n+4]) are moved global<->shared with fully coalesced accesses and transposed through
// shared memory so that each thread reads its own row; the discretised RC model (A_d, BD_d,
// target, ac_map) is staged once per CTA in shared memory and read as warp-uniform broadcasts.
// The weather table is packed [T,4] so an env's four exogenous values are one 32-byte sector.
//
// Bound: HBM (105 B per env-step at n = 5: SURVEY.md 8d).  All arithmetic is f64 except where
// the reference itself computes in f32 (avg_temp, the NEP-50 f32 products inside
// _calc_occupower, the f32 action norm); those are mirrored with explicit __f*_rn intrinsics.
#include <cooperative_groups.h>

#include "common.cuh"
#include "tc.cuh"

namespace cg = cooperative_groups;

namespace sg {
namespace {

#ifndef SG_BLD_TILE
#define SG_BLD_TILE 32     // one warp per CTA: tiles advance independently (measured +9 % over 128 on the fused rollout)
#endif
constexpr int kTile = SG_BLD_TILE;  // envs per CTA
constexpr int kPre = 3;     // rollout software pipeline: action tiles of the next kPre steps are in flight (cp.async)
constexpr int kRing = kPre + 1;
constexpr int kTileStep = 128;   // envs per CTA tile of the single-step kernels (see launch_building)

__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__constant__ double kOccu[8] = {6.461927, 0.946892, 0.0000255737, 0.0627909,
                                0.0000589172, 0.19855, 0.000940018, 0.00000149532};

// env.py:424-433 with temp an f32 scalar (NumPy >= 2 promotion: Python-float * f32 -> f32).
__device__ __forceinline__ double occupower_f32temp(float t, double m) {
    const float t2 = __fmul_rn(t, t);
    const double m2 = __dmul_rn(m, m);
    double h = __dadd_rn(kOccu[0], __dmul_rn(kOccu[1], m));
    h = __dadd_rn(h, __dmul_rn(kOccu[2], m2));
    h = __dsub_rn(h, __dmul_rn((double)__fmul_rn((float)kOccu[3], t), m));
    h = __dadd_rn(h, __dmul_rn((double)__fmul_rn((float)kOccu[4], t), m2));
    h = __dsub_rn(h, (double)__fmul_rn((float)kOccu[5], t2));
    h = __dadd_rn(h, __dmul_rn((double)__fmul_rn((float)kOccu[6], t2), m));
    h = __dsub_rn(h, __dmul_rn((double)__fmul_rn((float)kOccu[7], t2), m2));
    return h;
}
// same polynomial with an f64 temperature (reset path, env.py:358-364)
__device__ __forceinline__ double occupower_f64temp(double t, double m) {
    const double t2 = __dmul_rn(t, t), m2 = __dmul_rn(m, m);
    double h = __dadd_rn(kOccu[0], __dmul_rn(kOccu[1], m));
    h = __dadd_rn(h, __dmul_rn(kOccu[2], m2));
    h = __dsub_rn(h, __dmul_rn(__dmul_rn(kOccu[3], t), m));
    h = __dadd_rn(h, __dmul_rn(__dmul_rn(kOccu[4], t), m2));
    h = __dsub_rn(h, __dmul_rn(kOccu[5], t2));
    h = __dadd_rn(h, __dmul_rn(__dmul_rn(kOccu[6], t2), m));
    h = __dsub_rn(h, __dmul_rn(__dmul_rn(kOccu[7], t2), m2));
    return h;
}

// np.sum over an f32 vector: NumPy's pairwise_sum (sequential below 8 elements, 8 running
// accumulators up to 128) — reproduced so avg_temp is bit-identical to env.py:249.
template <int MAXN>
__device__ __forceinline__ float numpy_sum_f32(const float (&x)[MAXN], int n) {
    if (MAXN < 8 || n < 8) {
        float r = 0.f;
#pragma unroll
        for (int i = 0; i < (MAXN < 8 ? MAXN : 7); ++i)
            if (i < n) r = __fadd_rn(r, x[i]);
        return r;
    }
    if constexpr (MAXN >= 8) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = x[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], x[i + j]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, x[i]);
    return res;
    } else {
        return 0.f;
    }
}

struct Weather4 { double out, ground, ghi, meta; };
__device__ __forceinline__ Weather4 load_weather(const double* w, int e) {
    const double2* p = reinterpret_cast<const double2*>(w + 4ll * e);
    const double2 a = __ldg(p), b = __ldg(p + 1);
    return {a.x, a.y, b.x, b.y};
}

// One env, one step (env.py:233-312): action row -> exogenous inputs + occupancy heat -> X_new = A_d X + BD_d Y ->
// reward -> cursor / counters (+ the auto-reset state, env.py:346-378).  `w` is the weather row of the current cursor,
// `w_after` the row after it (read only by an auto-reset).  Shared by the rollout kernel and the single-step kernel.
template <int N, bool DISCRETE, int MAXN>
__device__ __forceinline__ void bld_env_step(const sg_building_model& m, const int n, const double* __restrict__ sA,
                                             const double* __restrict__ sBD, const double* __restrict__ sTarget,
                                             const double* __restrict__ sAc, const unsigned char* __restrict__ act_row,
                                             float (&x)[MAXN], float (&tail)[4], int& e, int& r, const Weather4& w,
                                             const Weather4& w_after, const int flags, StatsAcc& acc, double& rew, bool& dn) {
    const int row = n + 4;
    // ---- action: env.py:234-235 ----------------------------------------------------
    double act[MAXN];
    float a_dot = 0.f;       // f32 x.dot(x) of LA.norm (continuous actions are f32)
    double a_dot64 = 0.0;
#pragma unroll
    for (int j = 0; j < n; ++j) {
        if constexpr (DISCRETE) {
            const long long ai = reinterpret_cast<const long long*>(act_row)[j];
            const float qlow100 = __fmul_rn(-(float)sAc[j], 100.f);
            act[j] = ((double)ai + (double)qlow100) / 100.0;
            a_dot64 = fma(act[j], act[j], a_dot64);
        } else {
            const float af = reinterpret_cast<const float*>(act_row)[j];
            act[j] = (double)af;
            a_dot = __fmaf_rn(af, af, a_dot);
            a_dot64 = fma(act[j], act[j], a_dot64);
        }
    }
    // ---- exogenous inputs + occupancy heat: env.py:244-263 --------------------------
    const float avg_temp = __fdiv_rn(numpy_sum_f32<MAXN>(x, n), (float)n);
    const double occ = occupower_f32temp(avg_temp, w.meta);
    // ---- X_new = A_d X + BD_d Y: env.py:266 ; error/reward: env.py:272-276 -----------
    double err2 = 0.0, errp = 0.0;
    float xn[MAXN];
    const bool p2 = (m.pnorm == 2.0);
#pragma unroll
    for (int i = 0; i < n; ++i) {
        const double* Ai = sA + i * n;
        const double* Bi = sBD + i * row;
        double v = Bi[0] * occ;
        v = fma(Bi[1], w.ground, v);
        v = fma(Bi[2], w.out, v);
        v = fma(Bi[n + 3], w.ghi, v);
        double va = 0.0;
#pragma unroll
        for (int j = 0; j < n; ++j) { va = fma(Ai[j], (double)x[j], va); v = fma(Bi[3 + j], act[j], v); }
        v += va;
        const double er = v * sAc[i] - sTarget[i] * sAc[i];
        err2 = fma(er, er, err2);
        if (!p2) errp += pow(fabs(er), m.pnorm);
        xn[i] = (float)v;
    }
    double a_term, e_term;
    if (p2) {
        e_term = sqrt(err2) * m.error_rate;
        if constexpr (DISCRETE) a_term = sqrt(a_dot64) * m.q_rate;
        else a_term = (double)__fmul_rn(__fsqrt_rn(a_dot), (float)m.q_rate);
    } else {
        double ap = 0.0;
        for (int j = 0; j < n; ++j) ap += pow(fabs(act[j]), m.pnorm);
        a_term = pow(ap, 1.0 / m.pnorm) * m.q_rate;
        e_term = pow(errp, 1.0 / m.pnorm) * m.error_rate;
    }
    rew = -(a_term + e_term);
    // ---- new state: env.py:286-296 ; cursor/counters: env.py:302-309 ------------------
#pragma unroll
    for (int j = 0; j < n; ++j) x[j] = xn[j];
    tail[0] = (float)w.out; tail[1] = (float)w.ground; tail[2] = (float)w.ghi;
    tail[3] = (float)(occ / 1000.0);
    e += 1; r += 1;
    if (e >= m.T) e = 0;
    dn = r >= m.episode_len;
    acc.add(rew, dn, -e_term, -a_term, 0.0);
    if (dn && (flags & SG_BUILDING_FLAG_AUTORESET)) {   // env.py:346-378 at the current cursor
        r = 0;
        double sum = 0.0;
        for (int j = 0; j < n; ++j) sum += sTarget[j];
        const Weather4 w2 = w_after;                     // (row e: already prefetched)
        const double occ0 = occupower_f64temp(sum / n, w2.meta);
#pragma unroll
        for (int j = 0; j < n; ++j) x[j] = (float)sTarget[j];
        tail[0] = (float)w2.out; tail[1] = (float)w2.ground; tail[2] = (float)w2.ghi;
        tail[3] = (float)(occ0 / 1000.0);
    }
}

// N > 0: compile-time zone count, everything in registers, rows staged through shared memory.
// N == 0: generic fallback (n <= 128), per-thread arrays in local memory, direct global rows.
#ifndef SG_BLD_STEP_MINB
#define SG_BLD_STEP_MINB 4   // resident 128-thread CTAs per SM the single-step instance is compiled for (see kTileStep)
#endif
template <int N, bool DISCRETE, int TILE>
// (one-warp CTAs of the rollout: 12 per SM = 3 warps per scheduler needs <= 170 registers; left to itself ptxas takes
// 190, i.e. 2 warps per scheduler and 1.5x the waves of CTAs at 262 144 envs — measured 0.38 against 0.49 of HBM)
__global__ void __launch_bounds__(TILE, TILE >= 128 ? SG_BLD_STEP_MINB : 12)
building_step_kernel(const sg_building_model m, const long long B, const int steps_arg,
                     const void* __restrict__ action_v, float* __restrict__ state,
                     int32_t* __restrict__ epoch_g, int32_t* __restrict__ runs_g,
                     float* __restrict__ obs_out, double* __restrict__ reward_g,
                     uint8_t* __restrict__ done_g, const int flags, double* __restrict__ stats) {
    constexpr int MAXN = N > 0 ? N : 128;
    // the 128-env instance is the single-step one (launch_building): with the step count a compile-time 1 the step
    // loop disappears and the model matrices are read from shared memory where they are used instead of being
    // hoisted into ~140 registers for the whole loop — which is what capped that instance at 3-4 CTAs per SM
    const int steps = TILE >= 128 ? 1 : steps_arg;
    const int n = N > 0 ? N : m.n;
    const int row = n + 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);            // [n*n]
    double* sBD = sA + n * n;                                      // [n*(n+4)]
    double* sTarget = sBD + n * row;                               // [n]
    double* sAc = sTarget + n;                                     // [n]
    double* sStat = sAc + n;                                       // [11 * 8]
    float* tile = reinterpret_cast<float*>(sStat + 88);            // [TILE*row]   (N > 0 only)
    // action tile follows: f32 [TILE*n] or i64 [TILE*n]
    unsigned char* act_tile_raw = reinterpret_cast<unsigned char*>(tile + (N > 0 ? TILE * row : 0));
    act_tile_raw = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(act_tile_raw) + 15) & ~uintptr_t(15));

    const int tid = threadIdx.x;
    const long long b0 = (long long)blockIdx.x * TILE;
    const int nvalid = (int)min((long long)TILE, B - b0);
    const long long b = b0 + tid;
    const bool valid = tid < nvalid;

    // Everything the first step needs is requested up front, in dependency order, so that a CTA's start-up costs two
    // memory round trips instead of four (a single step is start-up plus one step; with ~3.5 waves of CTAs that
    // latency IS the launch time): the weather cursor first (the weather rows depend on it), then the model matrices
    // and the state tile as asynchronous copies straight into shared memory, then the action tiles, then — once the
    // cursor has arrived — the weather rows; the copies are waited for last.
    float x[MAXN];
    int e = 0, r = 0;
    if (valid) {
        e = epoch_g[b];
        r = runs_g[b];
    }
    for (int i = tid; i < n * n; i += TILE) cp_async_8(sA + i, m.A_d + i);
    for (int i = tid; i < n * row; i += TILE) cp_async_8(sBD + i, m.BD_d + i);
    for (int i = tid; i < n; i += TILE) { cp_async_8(sTarget + i, m.target + i); cp_async_8(sAc + i, m.ac_map + i); }
    if constexpr (N > 0) {
        const float* g = state + b0 * row;
        const int nbytes = nvalid * row * 4;
        if ((nbytes & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
            for (int i = tid * 16; i < nbytes; i += TILE * 16)
                cp_async_16(reinterpret_cast<unsigned char*>(tile) + i, reinterpret_cast<const unsigned char*>(g) + i);
        } else {
            for (int i = tid; i < nvalid * row; i += TILE) cp_async_4(tile + i, g + i);
        }
    }
    cp_async_commit();
    StatsAcc acc;
    float tail[4] = {0.f, 0.f, 0.f, 0.f};

    // action tile ring (N > 0): slot (s % kRing) holds step s; every call commits one group (maybe empty)
    constexpr int esz = DISCRETE ? 8 : 4;
    const size_t act_tile_bytes = ((size_t)TILE * n * esz + 15) & ~(size_t)15;
    auto stage_actions = [&](const int sp) {
        if (sp < steps) {
            const char* g = reinterpret_cast<const char*>(action_v) + ((long long)sp * B + b0) * n * esz;
            unsigned char* t = act_tile_raw + (size_t)(sp % kRing) * act_tile_bytes;
            const int nbytes = nvalid * n * esz;
            if ((nbytes & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
                for (int i = tid * 16; i < nbytes; i += TILE * 16) cp_async_16(t + i, g + i);
            } else {
                for (int i = tid * 4; i < nbytes; i += TILE * 4) cp_async_4(t + i, g + i);
            }
        }
        cp_async_commit();
    };
    if constexpr (N > 0) {
#pragma unroll
        for (int d = 0; d < kPre; ++d) stage_actions(d);
    }
    // weather rows of this step and the next one (the cursor advances by one row per step)
    Weather4 w_cur = {0, 0, 0, 0}, w_nxt = {0, 0, 0, 0};
    if (valid) {
        w_cur = load_weather(m.weather, e);
        // (a single step needs the next row only for an auto-reset at the end of the episode)
        if (TILE < 128 || r + 1 >= m.episode_len) w_nxt = load_weather(m.weather, (e + 1) % m.T);
    }
    // model matrices + state tile have landed (the action tiles of the ring may still be in flight)
    if constexpr (N > 0) cp_async_wait<kPre>(); else cp_async_wait<0>();
    __syncthreads();
    if (valid) {
        if constexpr (N > 0) {
#pragma unroll
            for (int j = 0; j < N; ++j) x[j] = tile[tid * row + j];
        } else {
            for (int j = 0; j < n; ++j) x[j] = state[b * row + j];
        }
    }

#ifdef SG_BLD_TRACE2
    long long tp[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tq = clock64(), tq2;
#define SG_P2(i) tq2 = clock64(); tp[i] += tq2 - tq; tq = tq2
#else
#define SG_P2(i)
#endif
    for (int s = 0; s < steps; ++s) {
        // ---- this step's action rows: landed in the ring (issued kPre steps ago) --------------------
        const unsigned char* act_tile = act_tile_raw;
        if constexpr (N > 0) {
            cp_async_wait<kPre - 1>();   // this thread's copies for step s
            __syncthreads();             // everyone's copies landed; previous step's readers of the tiles are done
            stage_actions(s + kPre);     // refills the slot step s-1 used
            act_tile = act_tile_raw + (size_t)(s % kRing) * act_tile_bytes;
        }
        SG_P2(0);
        double rew = 0.0;
        bool dn = false;
        if (valid) {
            const unsigned char* act_row;
            if constexpr (N > 0) act_row = act_tile + (size_t)tid * n * esz;
            else act_row = reinterpret_cast<const unsigned char*>(action_v) + ((long long)s * B + b) * n * esz;
            // exogenous inputs of this step; the row the step after next will read is requested now
            const Weather4 w = w_cur;
            w_cur = w_nxt;
            // (unconditional in the rollout instance: a test on the step index here costs the rollout 20 % — the load
            // is no longer scheduled ahead of the step's arithmetic; the row past the last step is read and dropped)
            if (TILE < 128) w_nxt = load_weather(m.weather, (e + 2) % m.T);
            bld_env_step<N, DISCRETE, MAXN>(m, n, sA, sBD, sTarget, sAc, act_row, x, tail, e, r, w, w_cur, flags, acc, rew, dn);
            reward_g[(long long)s * B + b] = rew;
            done_g[(long long)s * B + b] = dn ? 1 : 0;
        }
        SG_P2(4);
        // ---- emit the observation row (per step for rollouts, final state always) ------------
        const bool last = (s == steps - 1);
        if (obs_out != nullptr || last) {
            if constexpr (N > 0) {
                if (valid) {
#pragma unroll
                    for (int j = 0; j < N; ++j) tile[tid * row + j] = x[j];
#pragma unroll
                    for (int j = 0; j < 4; ++j) tile[tid * row + N + j] = tail[j];
                }
                __syncthreads();
                if (obs_out != nullptr) {
                    float* g = obs_out + ((long long)s * B + b0) * row;
                    for (int i = tid; i < nvalid * row; i += TILE) g[i] = tile[i];
                }
                if (last) {
                    float* g = state + b0 * row;
                    for (int i = tid; i < nvalid * row; i += TILE) g[i] = tile[i];
                }
            } else if (valid) {
                if (obs_out != nullptr) {
                    float* g = obs_out + ((long long)s * B + b) * row;
                    for (int j = 0; j < n; ++j) g[j] = x[j];
                    for (int j = 0; j < 4; ++j) g[n + j] = tail[j];
                }
                if (last) {
                    float* g = state + b * row;
                    for (int j = 0; j < n; ++j) g[j] = x[j];
                    for (int j = 0; j < 4; ++j) g[n + j] = tail[j];
                }
            }
        }
        SG_P2(5);
    }
#ifdef SG_BLD_TRACE2
    if (blockIdx.x == 5 && tid == 0) printf("steps %d: top %lld act %lld occ %lld matvec %lld reward %lld emit %lld (cycles/step)\n", steps, tp[0] / steps, tp[1] / steps, tp[2] / steps, tp[3] / steps, tp[4] / steps, tp[5] / steps);
#endif
    if (valid) { epoch_g[b] = e; runs_g[b] = r; }
    if (stats != nullptr) {
        __syncthreads();
        stats_flush_block(acc, stats, sStat);
    }
}

// Single step, N > 0, large batches: a persistent kernel.  The one-tile-per-CTA form above spends a CTA's life on
// start-up latency (model matrices, state tile, cursor -> weather row: ~6 us from launch to exit, measured at 32 768
// envs where every CTA is resident at once) and a 262 144-env step was 2.8 such waves.  Here a CTA keeps the model in
// shared memory and walks over tiles of 128 envs with a two-slot ring: while tile i is computed, the state rows, action
// rows and cursors of tile i+1 are already on their way (cp.async / plain loads into registers), so a tile costs its
// arithmetic plus the weather-row lookup instead of the whole chain.  Same arithmetic (bld_env_step), same results.
template <int N, bool DISCRETE>
__global__ void __launch_bounds__(kTileStep, SG_BLD_STEP_MINB)
building_step1_kernel(const sg_building_model m, const long long B, const void* __restrict__ action_v,
                      float* __restrict__ state, int32_t* __restrict__ epoch_g, int32_t* __restrict__ runs_g,
                      double* __restrict__ reward_g, uint8_t* __restrict__ done_g, const int flags,
                      double* __restrict__ stats) {
    constexpr int TILE = kTileStep, n = N, row = N + 4, esz = DISCRETE ? 8 : 4;
    constexpr int tile_bytes = TILE * row * 4, act_bytes = (TILE * n * esz + 15) & ~15;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);
    double* sBD = sA + n * n;
    double* sTarget = sBD + n * row;
    double* sAc = sTarget + n;
    double* sStat = sAc + n;                                       // [11 * 8]
    unsigned char* ring = reinterpret_cast<unsigned char*>(sStat + 88);   // 2 x (state tile | action tile)
    const int tid = threadIdx.x;
    const long long ntiles = (B + TILE - 1) / TILE;

    // tile t into ring slot `slot`: this thread's cursors (weather row, step count) into registers, the state and
    // action rows as asynchronous copies (one commit group, maybe empty)
    auto issue = [&](const long long t, const int slot, int& e_out, int& r_out) {
        if (t < ntiles) {
            const long long b0 = t * TILE;
            const int nvalid = (int)min((long long)TILE, B - b0);
            if (tid < nvalid) { e_out = epoch_g[b0 + tid]; r_out = runs_g[b0 + tid]; }
            unsigned char* st = ring + (size_t)slot * (tile_bytes + act_bytes);
            const unsigned char* gs = reinterpret_cast<const unsigned char*>(state + b0 * row);
            const unsigned char* ga = reinterpret_cast<const unsigned char*>(action_v) + b0 * n * esz;
            const int sb = nvalid * row * 4, ab = nvalid * n * esz;
            if ((sb & 15) == 0 && (reinterpret_cast<uintptr_t>(gs) & 15u) == 0)
                for (int i = tid * 16; i < sb; i += TILE * 16) cp_async_16(st + i, gs + i);
            else
                for (int i = tid * 4; i < sb; i += TILE * 4) cp_async_4(st + i, gs + i);
            if ((ab & 15) == 0 && (reinterpret_cast<uintptr_t>(ga) & 15u) == 0)
                for (int i = tid * 16; i < ab; i += TILE * 16) cp_async_16(st + tile_bytes + i, ga + i);
            else
                for (int i = tid * 4; i < ab; i += TILE * 4) cp_async_4(st + tile_bytes + i, ga + i);
        }
        cp_async_commit();
    };

    // (measured: requesting the weather row one tile ahead and the cursors two tiles ahead, or 5-6 CTAs per SM with
    // fewer registers, made the step slower — 16.2-19 us against 14.4 us; the kernel is bound by instruction issue
    // (~800 warp instructions per 32 envs), not by these latencies)
    int e_nx = 0, r_nx = 0;
    long long t = blockIdx.x;
    for (int i = tid; i < n * n; i += TILE) cp_async_8(sA + i, m.A_d + i);
    for (int i = tid; i < n * row; i += TILE) cp_async_8(sBD + i, m.BD_d + i);
    for (int i = tid; i < n; i += TILE) { cp_async_8(sTarget + i, m.target + i); cp_async_8(sAc + i, m.ac_map + i); }
    issue(t, 0, e_nx, r_nx);                                       // (the model copies ride in the first group)
    StatsAcc acc;
    for (int it = 0; t < ntiles; ++it, t += gridDim.x) {
        const int slot = it & 1;
        const long long b0 = t * TILE;
        const int nvalid = (int)min((long long)TILE, B - b0);
        const bool valid = tid < nvalid;
        const long long b = b0 + tid;
        int e = e_nx, r = r_nx;
        // weather rows of this tile's cursors (requested before anything waits; the next row only for an auto-reset)
        Weather4 w = {0, 0, 0, 0}, w_after = {0, 0, 0, 0};
        if (valid) {
            w = load_weather(m.weather, e);
            if (r + 1 >= m.episode_len) w_after = load_weather(m.weather, (e + 1) % m.T);
        }
        cp_async_wait<0>();
        __syncthreads();                                           // tile `it` landed; the other slot is free again
        issue(t + gridDim.x, slot ^ 1, e_nx, r_nx);
        float* tile = reinterpret_cast<float*>(ring + (size_t)slot * (tile_bytes + act_bytes));
        const unsigned char* act_tile = reinterpret_cast<const unsigned char*>(tile) + tile_bytes;
        if (valid) {
            float x[N], tail[4];
#pragma unroll
            for (int j = 0; j < N; ++j) x[j] = tile[tid * row + j];
            double rew;
            bool dn;
            bld_env_step<N, DISCRETE, N>(m, n, sA, sBD, sTarget, sAc, act_tile + (size_t)tid * n * esz, x, tail, e, r, w,
                                         w_after, flags, acc, rew, dn);
            reward_g[b] = rew;
            done_g[b] = dn ? 1 : 0;
            epoch_g[b] = e;
            runs_g[b] = r;
#pragma unroll
            for (int j = 0; j < N; ++j) tile[tid * row + j] = x[j];
#pragma unroll
            for (int j = 0; j < 4; ++j) tile[tid * row + N + j] = tail[j];
        }
        __syncthreads();
        {   // the new rows go back with coalesced 16-byte stores
            float* g = state + b0 * row;
            const int nb = nvalid * row * 4;
            if ((nb & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15u) == 0)
                for (int i = tid * 16; i < nb; i += TILE * 16)
                    *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(g) + i) =
                        *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(tile) + i);
            else
                for (int i = tid; i < nvalid * row; i += TILE) g[i] = tile[i];
        }
        // (the slot is refilled two iterations from now, after the __syncthreads at the top of the next one)
    }
    cp_async_wait<0>();
    if (stats != nullptr) {
        __syncthreads();
        stats_flush_block(acc, stats, sStat);
    }
}

__global__ void building_reset_kernel(const sg_building_model m, const long long B,
                                      const int32_t* __restrict__ epoch_in,
                                      const double* __restrict__ T_initial, float* __restrict__ state,
                                      int32_t* __restrict__ epoch_g, int32_t* __restrict__ runs_g) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int n = m.n, row = n + 4;
    int e = epoch_in[b];
    e = min(max(e, 0), m.T - 1);
    double sum = 0.0;   // np.sum(T_initial) in f64, env.py:358
    for (int j = 0; j < n; ++j) {
        const double t0 = T_initial ? T_initial[b * n + j] : m.target[j];
        sum += t0;
        state[b * row + j] = (float)t0;
    }
    const Weather4 w = load_weather(m.weather, e);
    const double occ = occupower_f64temp(sum / n, w.meta);
    state[b * row + n + 0] = (float)w.out;
    state[b * row + n + 1] = (float)w.ground;
    state[b * row + n + 2] = (float)w.ghi;
    state[b * row + n + 3] = (float)(occ / 1000.0);
    epoch_g[b] = e;
    runs_g[b] = 0;
}

// ---------------------------------------------------------------------------------------------------
// Large buildings (n >= 16, continuous actions, Euclidean reward): the zone coupling
//     X+ = A_d X + BD_d[:, 3:3+n] a + (4 rank-1 exogenous terms)
// is a dense f64 contraction with K = 2n, and runs on the tensor cores (DMMA, mma.sync m8n8k4 f64;
// f64 because the reference computes this product in f64 from its f32 state, env.py:266).  The f64
// weights of the biggest shipped building (n = 118: 2n x n x 8 B = 223 KB) do not fit one SM's shared
// memory next to the state tile, so a CLUSTER OF TWO CTAs owns a tile of 32 envs: each CTA keeps the
// weight columns of half of the zones resident in shared memory, computes that half of X+ for all 32
// envs, and writes its f32 results into both CTAs' next-state tiles through distributed shared memory;
// one cluster barrier per step.  Per-env scalar work (occupancy polynomial, action norm, reward,
// cursors, auto-reset) is done redundantly by warp 0 of both CTAs with the thread-per-env code's exact
// arithmetic; rank 0 writes rewards / done flags, the ranks split the observation rows.
// Summation order inside a dot product differs from the thread-per-env kernel (and from BLAS), i.e.
// f64 differences of ~1e-16 before the f32 rounding of the state.
constexpr int kMmaEnvs = 32;       // envs per cluster tile: 4 m-tiles of 8
constexpr int kMmaWarps = 8;        // contraction warps: warp -> (m-tile = warp & 3, n-tiles of parity warp >> 2)
constexpr int kMmaThreads = (kMmaWarps + 2) * 32;   // + the per-env scalar warp + the copy warp
constexpr int kBarOcc = 1, kBarErr = 2, kBarTail = 3;   // named barriers (0 is __syncthreads)
constexpr int kMmaMaxNT = 8;       // n-tiles of 8 zones per CTA (64 >= 59 = ceil(118 / 2))

struct MmaLayout {
    int n4, XP, NH, NT, NHp, NS;
    int AB;      // floats per action buffer: 32 packed rows + 4 zeroed pad floats (fragment loads run to n4)
    size_t off_R4, off_tgt, off_ac, off_tgtAll, off_env, off_errw, off_errp, off_stat, off_mbar, off_tail, off_X, off_A, off_O, total;
};
__host__ __device__ inline MmaLayout mma_layout(const int n) {
    MmaLayout L;
    L.n4 = (n + 3) & ~3;
    L.XP = (L.n4 % 8 == 4) ? L.n4 : L.n4 + 4;            // row stride of the f32 tiles: 16-byte bank groups of 8 rows distinct
    L.NH = (n + 1) / 2;
    L.NT = ((L.NH + 15) / 16) * 2;                       // n-tiles of 8 zones, even: the two warp groups take TI = NT / 2 each
    L.NHp = L.NT * 8;
    L.NS = L.NHp + 4;                                    // row stride of W (NHp % 16 == 0): 32 bytes mod 128, so the 4 k-rows x 4 zones
                                                         // a half-warp reads in one LDS.64 wavefront fall into distinct banks
    size_t o = sizeof(double) * (size_t)(2 * L.n4) * L.NS;                       // W [2 n4][NS]
    L.off_R4 = o; o += sizeof(double) * 4 * L.NHp;
    L.off_tgt = o; o += sizeof(double) * L.NHp;
    L.off_ac = o; o += sizeof(double) * L.NHp;
    L.off_tgtAll = o; o += sizeof(double) * L.n4;
    L.off_env = o; o += sizeof(double) * 5 * kMmaEnvs;                            // occ, ground, out, ghi, episode-ends flag per env
    L.off_errw = o; o += sizeof(double) * 2 * kMmaEnvs;
    L.off_errp = o; o += sizeof(double) * 2 * 2 * kMmaEnvs;                       // [step parity][rank][env]
    L.off_stat = o; o += sizeof(double) * 11 * (kMmaThreads / 32);
    L.off_mbar = o; o += 16;
    L.off_tail = o; o += sizeof(float) * 4 * kMmaEnvs;
    L.off_X = o; o += sizeof(float) * 2 * kMmaEnvs * L.XP;
    L.AB = kMmaEnvs * n + 4;
    L.off_A = o; o += sizeof(float) * 2 * L.AB;                                   // action tiles, packed [32][n] (one bulk copy each)
    L.off_O = o; o += (sizeof(float) * (kMmaEnvs / 2) * (n + 4) + 15) & ~(size_t)15;   // my half of the observation rows, packed
    L.total = o;
    return L;
}

__device__ __forceinline__ void named_bar_sync(const int id, const int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(const int id, const int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// numpy_sum_f32 for a row in shared memory (n >= 8): same operation order
__device__ __forceinline__ float numpy_sum_f32_row(const float* x, const int n) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = x[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], x[i + j]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, x[i]);
    return res;
}

template <int TI>       // n-tiles per warp (NT = 2 TI); tiles beyond the real zones multiply zero weight columns
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMmaThreads, 1)
building_mma_kernel(const sg_building_model m, const long long B, const int steps, const float* __restrict__ action,
                    float* __restrict__ state, int32_t* __restrict__ epoch_g, int32_t* __restrict__ runs_g,
                    float* __restrict__ obs_out, double* __restrict__ reward_g, uint8_t* __restrict__ done_g,
                    const int flags, double* __restrict__ stats) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int n = m.n, row = n + 4;
    const MmaLayout L = mma_layout(n);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* W = reinterpret_cast<double*>(smem_raw);
    double* R4 = reinterpret_cast<double*>(smem_raw + L.off_R4);
    double* sTgt = reinterpret_cast<double*>(smem_raw + L.off_tgt);
    double* sAc = reinterpret_cast<double*>(smem_raw + L.off_ac);
    double* sTgtAll = reinterpret_cast<double*>(smem_raw + L.off_tgtAll);
    double* sEnv = reinterpret_cast<double*>(smem_raw + L.off_env);
    double* sErrW = reinterpret_cast<double*>(smem_raw + L.off_errw);
    double* sErrP = reinterpret_cast<double*>(smem_raw + L.off_errp);
    double* sStat = reinterpret_cast<double*>(smem_raw + L.off_stat);
    float* sTail = reinterpret_cast<float*>(smem_raw + L.off_tail);
    float* Xs = reinterpret_cast<float*>(smem_raw + L.off_X);
    float* As = reinterpret_cast<float*>(smem_raw + L.off_A);
    float* Obuf = reinterpret_cast<float*>(smem_raw + L.off_O);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + L.off_mbar);     // [2]: action tile s & 1 has landed
    float* Xs_peer = cluster.map_shared_rank(Xs, rank ^ 1);
    double* sErrP_peer = cluster.map_shared_rank(sErrP, rank ^ 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long b0 = (long long)(blockIdx.x >> 1) * kMmaEnvs;
    const int nvalid = (int)min((long long)kMmaEnvs, B - b0);
    const int z_lo = rank * L.NH, nz = min(L.NH, n - z_lo);
    const int XP = L.XP, NS = L.NS, n4 = L.n4;

    // ---- resident weights of my zones: W[k][z], k < n4: A_d[z][k]; k >= n4: BD_d[z][3 + k - n4] ----
    for (int idx = tid; idx < 2 * n4 * NS; idx += kMmaThreads) {
        const int k = idx / NS, z = idx % NS;
        double v = 0.0;
        if (z < nz) {
            if (k < n) v = m.A_d[(z_lo + z) * n + k];
            else if (k >= n4 && k - n4 < n) v = m.BD_d[(z_lo + z) * row + 3 + (k - n4)];
        }
        W[idx] = v;
    }
    for (int z = tid; z < L.NHp; z += kMmaThreads) {
        const bool ok = z < nz;
        const double* bd = m.BD_d + (long long)(z_lo + z) * row;
        R4[0 * L.NHp + z] = ok ? bd[0] : 0.0;
        R4[1 * L.NHp + z] = ok ? bd[1] : 0.0;
        R4[2 * L.NHp + z] = ok ? bd[2] : 0.0;
        R4[3 * L.NHp + z] = ok ? bd[n + 3] : 0.0;
        sTgt[z] = ok ? m.target[z_lo + z] : 0.0;
        sAc[z] = ok ? m.ac_map[z_lo + z] : 0.0;
    }
    for (int j = tid; j < n4; j += kMmaThreads) sTgtAll[j] = j < n ? m.target[j] : 0.0;
    for (int i = tid; i < 2 * kMmaEnvs * XP; i += kMmaThreads) Xs[i] = 0.f;
    for (int i = tid; i < 2 * L.AB; i += kMmaThreads) As[i] = 0.f;
    if (tid == 0) { tc::mbar_init(&mbar[0], 1); tc::mbar_init(&mbar[1], 1); tc::mbar_fence_init(); }
    __syncthreads();
    for (int env = warp; env < nvalid; env += kMmaThreads / 32)
        for (int j = lane; j < n; j += 32) Xs[env * XP + j] = state[(b0 + env) * row + j];
    // The action tile of a step is one contiguous range of global memory: a single bulk asynchronous copy
    // (TMA) when its address and size are multiples of 16 bytes, 4/8-byte cp.async otherwise.
    const bool pair_ok = (n % 2 == 0) && ((reinterpret_cast<uintptr_t>(action) & 7u) == 0);
    const bool bulk_ok = nvalid == kMmaEnvs && ((B * n) % 4 == 0) && ((reinterpret_cast<uintptr_t>(action) & 15u) == 0);
    auto stage_actions = [&](const int sp) {       // copy warp: action tile of step sp -> As[sp & 1]
        if (sp >= steps) return;
        const float* g = action + ((long long)sp * B + b0) * n;
        float* t = As + (size_t)(sp & 1) * L.AB;
        if (bulk_ok) {
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)(kMmaEnvs * n * sizeof(float)), mba = tc::smem_u32(&mbar[sp & 1]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mba), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 tc::smem_u32(t)), "l"(g), "r"(bytes), "r"(mba) : "memory");
            }
        } else {
            for (int env = 0; env < nvalid; ++env) {
                if (pair_ok) {
                    for (int j = 2 * lane; j < n; j += 64) cp_async_8(t + env * n + j, g + env * n + j);
                } else {
                    for (int j = lane; j < n; j += 32) cp_async_4(t + env * n + j, g + env * n + j);
                }
            }
            cp_async_commit();
        }
    };
    auto wait_actions = [&](const int sp) {        // copy warp: the tile of step sp is in shared memory
        if (sp >= steps) return;
        if (bulk_ok) tc::mbar_wait(&mbar[sp & 1], (uint32_t)((sp >> 1) & 1));
        else cp_async_wait<0>();
    };
    // Roles: warps 0-7 tensor-core contraction + epilogue; warp 8 (lane = env) per-env scalars: occupancy,
    // weather, action norm, reward, cursors; warp 9 copies: action tiles in, observation rows out.  Hand-offs
    // inside a step are named barriers; the step boundary is the cluster barrier (every thread of both CTAs).
    const bool is_mma = warp < kMmaWarps, is_scalar = warp == kMmaWarps, is_copy = warp == kMmaWarps + 1;
    __syncthreads();                                 // zero fill and mbarrier init precede the first copy
    if (is_copy) { stage_actions(0); wait_actions(0); }
    const bool envthr = is_scalar && lane < nvalid;
    const long long b = b0 + lane;
    int e = 0, r = 0;
    Weather4 w_cur = {0, 0, 0, 0}, w_nxt = {0, 0, 0, 0};
    if (envthr) {
        e = epoch_g[b]; r = runs_g[b];
        w_cur = load_weather(m.weather, e);
        w_nxt = load_weather(m.weather, (e + 1) % m.T);
    }
    StatsAcc acc;
    const int mt = warp & 3, nt0 = (warp >> 2) & 1;
    const bool autoreset = (flags & SG_BUILDING_FLAG_AUTORESET) != 0;
    auto write_obs_rows = [&](const int s, const float* Xrow_src) {      // copy warp: my half of the rows of step s
        const bool last = (s == steps - 1);
        if (obs_out == nullptr && !last) return;
        const int e_lo = rank * (kMmaEnvs / 2), e_hi = min(nvalid, e_lo + kMmaEnvs / 2);
        if (e_hi <= e_lo) return;
        float* og = obs_out != nullptr ? obs_out + ((long long)s * B + b0 + e_lo) * row : nullptr;
        float* sg_ = last ? state + (b0 + e_lo) * row : nullptr;
        const uint32_t bytes = (uint32_t)((e_hi - e_lo) * row * sizeof(float));
        const bool bulk = (bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(og) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(sg_) & 15u) == 0);
        if (bulk) {   // rows assembled in shared memory, then ONE bulk asynchronous store per destination
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous store has read Obuf
            __syncwarp();
            for (int env = e_lo; env < e_hi; ++env)
                for (int j = lane; j < row; j += 32)
                    Obuf[(env - e_lo) * row + j] = j < n ? Xrow_src[env * XP + j] : sTail[env * 4 + (j - n)];
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (og != nullptr)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(og), "r"(tc::smem_u32(Obuf)), "r"(bytes) : "memory");
                if (sg_ != nullptr)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(sg_), "r"(tc::smem_u32(Obuf)), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            for (int env = e_lo; env < e_hi; ++env)
                for (int j = lane; j < row; j += 32) {
                    const float v = j < n ? Xrow_src[env * XP + j] : sTail[env * 4 + (j - n)];
                    if (og != nullptr) og[(env - e_lo) * row + j] = v;
                    if (sg_ != nullptr) sg_[(env - e_lo) * row + j] = v;
                }
        }
    };
    cluster.sync();                                  // both CTAs initialised, action tile 0 landed
#ifdef SG_BLD_TRACE
    long long tph[6] = {0, 0, 0, 0, 0, 0}, tq = clock64(), tq2;
#define SG_PH(i) tq2 = clock64(); tph[i] += tq2 - tq; tq = tq2
#else
#define SG_PH(i)
#endif

    for (int s = 0; s < steps; ++s) {
        const int cur = s & 1, nxt = cur ^ 1;
        float* Xc = Xs + (size_t)cur * kMmaEnvs * XP;
        float* Xn = Xs + (size_t)nxt * kMmaEnvs * XP;
        float* Xn_peer = Xs_peer + (size_t)nxt * kMmaEnvs * XP;
        const float* Ac = As + (size_t)cur * L.AB;      // packed rows of n floats
        if (is_mma) {
            // ---- tensor-core contraction: C[env, zone] over k = (x | action) ----------------------------
            // two accumulator sets (even / odd k-blocks): 8 independent DMMA chains per warp hide the pipe latency
            double c[TI][2], c2[TI][2];
#pragma unroll
            for (int ti = 0; ti < TI; ++ti) c[ti][0] = c[ti][1] = c2[ti][0] = c2[ti][1] = 0.0;
            const int xrow = (mt * 8 + (lane >> 2)) * XP + (lane & 3), acrow = (mt * 8 + (lane >> 2)) * n + (lane & 3);
            const int bcol = (lane & 3) * NS + (lane >> 2);
#pragma unroll 1
            for (int part = 0; part < 2; ++part) {
                const float* src = part == 0 ? Xc + xrow : Ac + acrow;
                const double* wb = W + (size_t)part * n4 * NS + bcol;
                int kb = 0;
#pragma unroll 3
                for (; kb + 1 < n4 / 4; kb += 2) {
                    const double a0 = (double)src[kb * 4], a1 = (double)src[kb * 4 + 4];
                    const double* wk = wb + (size_t)kb * 4 * NS;
#pragma unroll
                    for (int ti = 0; ti < TI; ++ti) {
                        const int t = nt0 + 2 * ti;
                        dmma_m8n8k4(c[ti][0], c[ti][1], a0, wk[t * 8]);
                        dmma_m8n8k4(c2[ti][0], c2[ti][1], a1, wk[4 * NS + t * 8]);
                    }
                }
                if (kb < n4 / 4) {
                    const double a0 = (double)src[kb * 4];
                    const double* wk = wb + (size_t)kb * 4 * NS;
#pragma unroll
                    for (int ti = 0; ti < TI; ++ti) dmma_m8n8k4(c[ti][0], c[ti][1], a0, wk[(nt0 + 2 * ti) * 8]);
                }
            }
#pragma unroll
            for (int ti = 0; ti < TI; ++ti) { c[ti][0] += c2[ti][0]; c[ti][1] += c2[ti][1]; }
            SG_PH(0);
            named_bar_sync(kBarOcc, (kMmaWarps + 1) * 32);
            SG_PH(1);      // the scalar warp has published this step's exogenous inputs
            // ---- epilogue: exogenous rank-1 terms, comfort error, new f32 state to both CTAs -------------
            {
                const int env = mt * 8 + (lane >> 2);
                const double occ_e = sEnv[0 * kMmaEnvs + env], gr_e = sEnv[1 * kMmaEnvs + env];
                const double out_e = sEnv[2 * kMmaEnvs + env], ghi_e = sEnv[3 * kMmaEnvs + env];
                const bool ends = sEnv[4 * kMmaEnvs + env] != 0.0;    // episode ends with this step and auto-reset is on
                double e2 = 0.0;
#pragma unroll
                for (int ti = 0; ti < TI; ++ti) {
                    const int t = nt0 + 2 * ti;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int z = t * 8 + 2 * (lane & 3) + q;
                        if (z < nz) {
                            double v = R4[0 * L.NHp + z] * occ_e;
                            v = fma(R4[1 * L.NHp + z], gr_e, v);
                            v = fma(R4[2 * L.NHp + z], out_e, v);
                            v = fma(R4[3 * L.NHp + z], ghi_e, v);
                            v += c[ti][q];
                            const double er = v * sAc[z] - sTgt[z] * sAc[z];
                            e2 = fma(er, er, e2);
                            const float xf = ends ? (float)sTgt[z] : (float)v;    // env.py:346-378: next episode starts at the targets
                            Xn[env * XP + z_lo + z] = xf;
                            Xn_peer[env * XP + z_lo + z] = xf;
                        }
                    }
                }
                e2 += __shfl_xor_sync(kFull, e2, 1);
                e2 += __shfl_xor_sync(kFull, e2, 2);
                if ((lane & 3) == 0) sErrW[nt0 * kMmaEnvs + env] = e2;
            }
            SG_PH(2);
            named_bar_sync(kBarErr, kMmaWarps * 32);
            if (warp == 0) {                         // my half's comfort error -> both CTAs
                const double part = sErrW[lane] + sErrW[kMmaEnvs + lane];
                sErrP[(cur * 2 + rank) * kMmaEnvs + lane] = part;
                sErrP_peer[(cur * 2 + rank) * kMmaEnvs + lane] = part;
            }
            SG_PH(3);
            cluster.sync();                          // X[nxt] and the error partials are complete in both CTAs
            SG_PH(4);
        } else if (is_scalar) {
            // ---- per-env exogenous inputs (env.py:244-263) -> MMA epilogue; the f32 action norm -----------
            Weather4 w = {0, 0, 0, 0};
            double occ = 0.0;
            float a_dot = 0.f;
            bool dn = false;
            if (envthr) {
                w = w_cur; w_cur = w_nxt;
                w_nxt = load_weather(m.weather, (e + 2) % m.T);
                const float avg_temp = __fdiv_rn(numpy_sum_f32_row(Xc + lane * XP, n), (float)n);
                occ = occupower_f32temp(avg_temp, w.meta);
                dn = r + 1 >= m.episode_len;
                sEnv[0 * kMmaEnvs + lane] = occ; sEnv[1 * kMmaEnvs + lane] = w.ground;
                sEnv[2 * kMmaEnvs + lane] = w.out; sEnv[3 * kMmaEnvs + lane] = w.ghi;
                sEnv[4 * kMmaEnvs + lane] = (dn && autoreset) ? 1.0 : 0.0;
            }
            __threadfence_block();
            named_bar_arrive(kBarOcc, (kMmaWarps + 1) * 32);
            SG_PH(0);
            if (envthr) {
                const float* ar = Ac + lane * n;
                for (int j = 0; j < n; ++j) a_dot = __fmaf_rn(ar[j], ar[j], a_dot);
            }
            SG_PH(1);
            cluster.sync();
            SG_PH(2);
            // ---- reward, cursors, auto-reset bookkeeping (env.py:272-309, 346-378) -------------------------
            if (envthr) {
                const double err2 = sErrP[(cur * 2 + 0) * kMmaEnvs + lane] + sErrP[(cur * 2 + 1) * kMmaEnvs + lane];
                const double e_term = sqrt(err2) * m.error_rate;
                const double a_term = (double)__fmul_rn(__fsqrt_rn(a_dot), (float)m.q_rate);
                const double rew = -(a_term + e_term);
                float t0 = (float)w.out, t1 = (float)w.ground, t2 = (float)w.ghi, t3 = (float)(occ / 1000.0);
                e += 1; r += 1;
                if (e >= m.T) e = 0;
                if (dn && autoreset) {
                    r = 0;
                    double sum = 0.0;
                    for (int j = 0; j < n; ++j) sum += sTgtAll[j];
                    const double occ0 = occupower_f64temp(sum / n, w_cur.meta);
                    t0 = (float)w_cur.out; t1 = (float)w_cur.ground; t2 = (float)w_cur.ghi; t3 = (float)(occ0 / 1000.0);
                }
                sTail[lane * 4 + 0] = t0; sTail[lane * 4 + 1] = t1; sTail[lane * 4 + 2] = t2; sTail[lane * 4 + 3] = t3;
                if (rank == 0) {
                    acc.add(rew, dn, -e_term, -a_term, 0.0);
                    reward_g[(long long)s * B + b] = rew;
                    done_g[(long long)s * B + b] = dn ? 1 : 0;
                }
            }
            __threadfence_block();
            named_bar_arrive(kBarTail, 64);          // -> copy warp: tails of step s are in place
            SG_PH(3);
        } else {
            // ---- copy warp: next action tile in, observation rows of the previous step out -------------------
            stage_actions(s + 1);
            SG_PH(0);
            if (s > 0) {
                named_bar_sync(kBarTail, 64);
                SG_PH(1);
                write_obs_rows(s - 1, Xc);           // X[cur] of step s is the state step s-1 produced
            }
            SG_PH(2);
            wait_actions(s + 1);
            SG_PH(3);
            cluster.sync();
            SG_PH(4);
        }
    }
#ifdef SG_BLD_TRACE
    if (blockIdx.x == 6 && lane == 0 && (warp == 0 || warp == 5 || warp >= kMmaWarps))
        printf("warp %d: %lld %lld %lld %lld %lld %lld (cycles/step)\n", warp, tph[0] / steps, tph[1] / steps, tph[2] / steps, tph[3] / steps, tph[4] / steps, tph[5] / steps);
#endif
    if (is_copy) {
        named_bar_sync(kBarTail, 64);
        write_obs_rows(steps - 1, Xs + (size_t)(steps & 1) * kMmaEnvs * XP);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // bulk stores complete before the CTA exits
    }
    if (envthr && rank == 0) { epoch_g[b] = e; runs_g[b] = r; }
    if (stats != nullptr && rank == 0) {
        __syncthreads();
        stats_flush_block(acc, stats, sStat);
    }
    cluster.sync();                                  // nobody exits while its shared memory may still be addressed
}

int launch_building_mma(const sg_building_model* m, long long B, int steps, const void* action, float* state,
                        int32_t* epoch, int32_t* runs, float* obs_out, double* reward, uint8_t* done,
                        int flags, double* stats, cudaStream_t st) {
    const MmaLayout L = mma_layout(m->n);
    const unsigned grid = 2u * (unsigned)((B + kMmaEnvs - 1) / kMmaEnvs);
    const float* act = static_cast<const float*>(action);
#define SG_MMA_CASE(TI_)                                                                                              \
    case TI_: {                                                                                                       \
        /* per-device attribute: set on every launch (a process may drive several GPUs) */                            \
        cudaError_t err = cudaFuncSetAttribute(building_mma_kernel<TI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total); \
        if (err != cudaSuccess) { set_error("building: %zu B of shared memory unavailable", L.total); return 2; }     \
        building_mma_kernel<TI_><<<grid, kMmaThreads, L.total, st>>>(*m, B, steps, act, state, epoch, runs, obs_out,  \
                                                                     reward, done, flags, stats);                     \
        break;                                                                                                        \
    }
    switch (L.NT / 2) {
        SG_MMA_CASE(1)
        SG_MMA_CASE(2)
        SG_MMA_CASE(3)
        SG_MMA_CASE(4)
        default:
            set_error("building: n=%d too large for the tensor-core kernel", m->n);
            return 1;
    }
#undef SG_MMA_CASE
    return launch_status("building_mma_kernel");
}

// Fused rollouts run one warp per CTA (tiles advance independently through the steps: measured +9 % over 128 threads);
// a single step() has no step loop to decouple, only per-CTA set-up (model -> shared memory, barriers) to amortise: 128
// envs per CTA (262 144 envs: 31.7 us with 8 192 one-warp CTAs).  Same code, same arithmetic.

template <int N, bool DISCRETE, int TILE>
int launch_building_tile(const sg_building_model* m, long long B, int steps, const void* action, float* state,
                         int32_t* epoch, int32_t* runs, float* obs_out, double* reward, uint8_t* done,
                         int flags, double* stats, cudaStream_t st) {
    const int n = m->n, row = n + 4;
    size_t smem = sizeof(double) * (n * n + n * row + 2 * n + 88);
    if (N > 0) smem += sizeof(float) * TILE * row + 16 + (size_t)kRing * ((((DISCRETE ? 8 : 4) * (size_t)TILE * n) + 15) & ~(size_t)15);
    auto kern = building_step_kernel<N, DISCRETE, TILE>;
    if (smem > 48 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) { set_error("building: %zu B of shared memory unavailable", smem); return 2; }
    }
    const unsigned grid = (unsigned)((B + TILE - 1) / TILE);
    kern<<<grid, TILE, smem, st>>>(*m, B, steps, action, state, epoch, runs, obs_out, reward, done, flags, stats);
    return launch_status("building_step_kernel");
}

template <int N, bool DISCRETE>
int launch_building(const sg_building_model* m, long long B, int steps, const void* action, float* state,
                    int32_t* epoch, int32_t* runs, float* obs_out, double* reward, uint8_t* done,
                    int flags, double* stats, cudaStream_t st) {
    if constexpr (N > 0) {
        if (steps == 1 && obs_out == nullptr && B >= 4 * kTileStep && !getenv("SG_BLD_NO_PERSISTENT")) {
            // persistent single-step kernel: every CTA resident, tiles of kTileStep envs dealt round-robin
            constexpr int esz = DISCRETE ? 8 : 4;
            const size_t smem = sizeof(double) * (N * N + N * (N + 4) + 2 * N + 88) +
                                2 * ((size_t)kTileStep * (N + 4) * 4 + (((size_t)kTileStep * N * esz + 15) & ~(size_t)15));
            auto kern = building_step1_kernel<N, DISCRETE>;
            if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                set_error("building: %zu B of shared memory unavailable", smem);
                return 2;
            }
            int sms = sg_device_sm_count();
            if (sms <= 0) sms = 148;
            const long long ntiles = (B + kTileStep - 1) / kTileStep;
            const unsigned grid = (unsigned)std::min<long long>(ntiles, (long long)sms * SG_BLD_STEP_MINB);
            kern<<<grid, kTileStep, smem, st>>>(*m, B, action, state, epoch, runs, reward, done, flags, stats);
            return launch_status("building_step1_kernel");
        }
        if (steps == 1 && B >= 4 * kTileStep)
            return launch_building_tile<N, DISCRETE, kTileStep>(m, B, steps, action, state, epoch, runs, obs_out, reward, done, flags, stats, st);
    }
    return launch_building_tile<N, DISCRETE, kTile>(m, B, steps, action, state, epoch, runs, obs_out, reward, done, flags, stats, st);
}

int dispatch_building(const sg_building_model* m, long long B, int steps, const void* action, float* state,
                      int32_t* epoch, int32_t* runs, float* obs_out, double* reward, uint8_t* done,
                      int flags, double* stats, void* stream) {
    SG_CHECK_ARG(m != nullptr, "building: model is NULL");
    SG_CHECK_ARG(m->n >= 1 && m->n <= 128, "building: n=%d outside [1,128]", m->n);
    SG_CHECK_ARG(m->T >= 1, "building: weather table is empty");
    SG_CHECK_ARG(B >= 0 && steps >= 1, "building: bad num_envs/steps");
    SG_CHECK_ARG(action && state && epoch && runs && reward && done, "building: NULL buffer");
    if (B == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (m->n >= 16 && (m->n + 1) / 2 <= 8 * kMmaMaxNT && !m->discrete && m->pnorm == 2.0 && !(flags & SG_BUILDING_FLAG_NO_MMA))
        return launch_building_mma(m, B, steps, action, state, epoch, runs, obs_out, reward, done, flags, stats, st);
#define SG_B_CASE(NN)                                                                                  \
    case NN:                                                                                           \
        return m->discrete ? launch_building<NN, true>(m, B, steps, action, state, epoch, runs, obs_out, \
                                                       reward, done, flags, stats, st)                 \
                           : launch_building<NN, false>(m, B, steps, action, state, epoch, runs, obs_out, \
                                                        reward, done, flags, stats, st);
    switch (m->n) {
        SG_B_CASE(3)
        SG_B_CASE(5)
        SG_B_CASE(6)
        SG_B_CASE(10)
        default:
            return m->discrete ? launch_building<0, true>(m, B, steps, action, state, epoch, runs, obs_out,
                                                          reward, done, flags, stats, st)
                               : launch_building<0, false>(m, B, steps, action, state, epoch, runs, obs_out,
                                                           reward, done, flags, stats, st);
    }
#undef SG_B_CASE
}

}  // namespace
}  // namespace sg

extern "C" {

int sg_building_reset(const sg_building_model* m, int64_t num_envs, const int32_t* epoch_in,
                      const double* T_initial, float* state, int32_t* epoch, int32_t* runs, void* stream) {
    SG_CHECK_ARG(m && epoch_in && state && epoch && runs, "building_reset: NULL argument");
    SG_CHECK_ARG(m->n >= 1 && m->n <= 128, "building_reset: n=%d outside [1,128]", m->n);
    if (num_envs <= 0) return 0;
    const unsigned grid = (unsigned)((num_envs + 127) / 128);
    sg::building_reset_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        *m, num_envs, epoch_in, T_initial, state, epoch, runs);
    return sg::launch_status("building_reset_kernel");
}

int sg_building_step(const sg_building_model* m, int64_t num_envs, const void* action, float* state,
                     int32_t* epoch, int32_t* runs, double* reward, uint8_t* done, int32_t flags,
                     double* stats, void* stream) {
    return sg::dispatch_building(m, num_envs, 1, action, state, epoch, runs, nullptr, reward, done, flags,
                                 stats, stream);
}

int sg_building_rollout(const sg_building_model* m, int64_t num_envs, int32_t steps, const void* action,
                        float* state, int32_t* epoch, int32_t* runs, float* obs_out, double* reward,
                        uint8_t* done, int32_t flags, double* stats, void* stream) {
    return sg::dispatch_building(m, num_envs, steps, action, state, epoch, runs, obs_out, reward, done,
                                 flags, stats, stream);
}

int sg_building_step_host(const sg_building_model* m, int64_t num_envs, const void* h_action, float* h_obs,
                          double* h_reward, uint8_t* h_done, void* d_action, float* d_state,
                          int32_t* d_epoch, int32_t* d_runs, double* d_reward, uint8_t* d_done,
                          int32_t flags, double* d_stats, void* stream) {
    SG_CHECK_ARG(m && h_action && h_obs && h_reward && h_done && d_action, "building_step_host: NULL argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t B = (size_t)num_envs, n = (size_t)m->n;
    SG_CUDA(cudaMemcpyAsync(d_action, h_action, B * n * (m->discrete ? 8 : 4), cudaMemcpyHostToDevice, st));
    int rc = sg::dispatch_building(m, num_envs, 1, d_action, d_state, d_epoch, d_runs, nullptr, d_reward,
                                   d_done, flags, d_stats, stream);
    if (rc) return rc;
    SG_CUDA(cudaMemcpyAsync(h_obs, d_state, B * (n + 4) * 4, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(h_reward, d_reward, B * 8, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(h_done, d_done, B, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
