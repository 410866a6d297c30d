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
#include "common.cuh"

namespace sg {
namespace {

#ifndef SG_BLD_TILE
#define SG_BLD_TILE 32     // one warp per CTA: tiles advance independently (measured +9 % over 128 on the fused rollout)
#endif
constexpr int kTile = SG_BLD_TILE;  // envs per CTA
constexpr int kPre = 3;     // rollout software pipeline: action tiles of the next kPre steps are in flight (cp.async)
constexpr int kRing = kPre + 1;

__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
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

// N > 0: compile-time zone count, everything in registers, rows staged through shared memory.
// N == 0: generic fallback (n <= 128), per-thread arrays in local memory, direct global rows.
template <int N, bool DISCRETE>
__global__ void __launch_bounds__(kTile)
building_step_kernel(const sg_building_model m, const long long B, const int steps,
                     const void* __restrict__ action_v, float* __restrict__ state,
                     int32_t* __restrict__ epoch_g, int32_t* __restrict__ runs_g,
                     float* __restrict__ obs_out, double* __restrict__ reward_g,
                     uint8_t* __restrict__ done_g, const int flags, double* __restrict__ stats) {
    constexpr int MAXN = N > 0 ? N : 128;
    const int n = N > 0 ? N : m.n;
    const int row = n + 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);            // [n*n]
    double* sBD = sA + n * n;                                      // [n*(n+4)]
    double* sTarget = sBD + n * row;                               // [n]
    double* sAc = sTarget + n;                                     // [n]
    double* sStat = sAc + n;                                       // [11 * 4]
    float* tile = reinterpret_cast<float*>(sStat + 44);            // [kTile*row]   (N > 0 only)
    // action tile follows: f32 [kTile*n] or i64 [kTile*n]
    unsigned char* act_tile_raw = reinterpret_cast<unsigned char*>(tile + (N > 0 ? kTile * row : 0));
    act_tile_raw = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(act_tile_raw) + 15) & ~uintptr_t(15));

    const int tid = threadIdx.x;
    for (int i = tid; i < n * n; i += kTile) sA[i] = m.A_d[i];
    for (int i = tid; i < n * row; i += kTile) sBD[i] = m.BD_d[i];
    for (int i = tid; i < n; i += kTile) { sTarget[i] = m.target[i]; sAc[i] = m.ac_map[i]; }

    const long long b0 = (long long)blockIdx.x * kTile;
    const int nvalid = (int)min((long long)kTile, B - b0);
    const long long b = b0 + tid;
    const bool valid = tid < nvalid;

    float x[MAXN];
    int e = 0, r = 0;
    if constexpr (N > 0) {
        const float* g = state + b0 * row;
        for (int i = tid; i < nvalid * row; i += kTile) tile[i] = g[i];
    }
    __syncthreads();
    if (valid) {
        e = epoch_g[b];
        r = runs_g[b];
        if constexpr (N > 0) {
#pragma unroll
            for (int j = 0; j < N; ++j) x[j] = tile[tid * row + j];
        } else {
            for (int j = 0; j < n; ++j) x[j] = state[b * row + j];
        }
    }
    StatsAcc acc;
    float tail[4] = {0.f, 0.f, 0.f, 0.f};

    // action tile ring (N > 0): slot (s % kRing) holds step s; every call commits one group (maybe empty)
    constexpr int esz = DISCRETE ? 8 : 4;
    const size_t act_tile_bytes = ((size_t)kTile * n * esz + 15) & ~(size_t)15;
    auto stage_actions = [&](const int sp) {
        if (sp < steps) {
            const char* g = reinterpret_cast<const char*>(action_v) + ((long long)sp * B + b0) * n * esz;
            unsigned char* t = act_tile_raw + (size_t)(sp % kRing) * act_tile_bytes;
            const int nbytes = nvalid * n * esz;
            if ((nbytes & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
                for (int i = tid * 16; i < nbytes; i += kTile * 16) cp_async_16(t + i, g + i);
            } else {
                for (int i = tid * 4; i < nbytes; i += kTile * 4) cp_async_4(t + i, g + i);
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
        w_nxt = load_weather(m.weather, (e + 1) % m.T);
    }

    for (int s = 0; s < steps; ++s) {
        // ---- this step's action rows: landed in the ring (issued kPre steps ago) --------------------
        const unsigned char* act_tile = act_tile_raw;
        if constexpr (N > 0) {
            cp_async_wait<kPre - 1>();   // this thread's copies for step s
            __syncthreads();             // everyone's copies landed; previous step's readers of the tiles are done
            stage_actions(s + kPre);     // refills the slot step s-1 used
            act_tile = act_tile_raw + (size_t)(s % kRing) * act_tile_bytes;
        }
        double rew = 0.0;
        bool dn = false;
        if (valid) {
            // ---- action: env.py:234-235 ----------------------------------------------------
            double act[MAXN];
            float a_dot = 0.f;       // f32 x.dot(x) of LA.norm (continuous actions are f32)
            double a_dot64 = 0.0;
#pragma unroll
            for (int j = 0; j < n; ++j) {
                {
                    if constexpr (DISCRETE) {
                        long long ai;
                        if constexpr (N > 0) ai = reinterpret_cast<const long long*>(act_tile)[tid * n + j];
                        else ai = reinterpret_cast<const long long*>(action_v)[((long long)s * B + b) * n + j];
                        const float qlow100 = __fmul_rn(-(float)sAc[j], 100.f);
                        act[j] = ((double)ai + (double)qlow100) / 100.0;
                        a_dot64 = fma(act[j], act[j], a_dot64);
                    } else {
                        float af;
                        if constexpr (N > 0) af = reinterpret_cast<const float*>(act_tile)[tid * n + j];
                        else af = reinterpret_cast<const float*>(action_v)[((long long)s * B + b) * n + j];
                        act[j] = (double)af;
                        a_dot = __fmaf_rn(af, af, a_dot);
                        a_dot64 = fma(act[j], act[j], a_dot64);
                    }
                }
            }
            // ---- exogenous inputs + occupancy heat: env.py:244-263 --------------------------
            const Weather4 w = w_cur;
            w_cur = w_nxt;
            {   // row the step after next will read
                w_nxt = load_weather(m.weather, (e + 2) % m.T);
            }
            const float avg_temp = __fdiv_rn(numpy_sum_f32<MAXN>(x, n), (float)n);
            const double occ = occupower_f32temp(avg_temp, w.meta);
            // ---- X_new = A_d X + BD_d Y: env.py:266 ; error/reward: env.py:272-276 -----------
            double err2 = 0.0, errp = 0.0;
            float xn[MAXN];
            const bool p2 = (m.pnorm == 2.0);
#pragma unroll
            for (int i = 0; i < n; ++i) {
                {
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
                const Weather4 w2 = w_cur;                       // (row e: already prefetched)
                const double occ0 = occupower_f64temp(sum / n, w2.meta);
#pragma unroll
                for (int j = 0; j < n; ++j) x[j] = (float)sTarget[j];
                tail[0] = (float)w2.out; tail[1] = (float)w2.ground; tail[2] = (float)w2.ghi;
                tail[3] = (float)(occ0 / 1000.0);
            }
            reward_g[(long long)s * B + b] = rew;
            done_g[(long long)s * B + b] = dn ? 1 : 0;
        }
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
                    for (int i = tid; i < nvalid * row; i += kTile) g[i] = tile[i];
                }
                if (last) {
                    float* g = state + b0 * row;
                    for (int i = tid; i < nvalid * row; i += kTile) g[i] = tile[i];
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
    }
    if (valid) { epoch_g[b] = e; runs_g[b] = r; }
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

template <int N, bool DISCRETE>
int launch_building(const sg_building_model* m, long long B, int steps, const void* action, float* state,
                    int32_t* epoch, int32_t* runs, float* obs_out, double* reward, uint8_t* done,
                    int flags, double* stats, cudaStream_t st) {
    const int n = m->n, row = n + 4;
    size_t smem = sizeof(double) * (n * n + n * row + 2 * n + 44);
    if (N > 0) smem += sizeof(float) * kTile * row + 16 + (size_t)kRing * ((((DISCRETE ? 8 : 4) * (size_t)kTile * n) + 15) & ~(size_t)15);
    auto kern = building_step_kernel<N, DISCRETE>;
    if (smem > 48 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) { set_error("building: %zu B of shared memory unavailable", smem); return 2; }
    }
    const unsigned grid = (unsigned)((B + kTile - 1) / kTile);
    kern<<<grid, kTile, smem, st>>>(*m, B, steps, action, state, epoch, runs, obs_out, reward, done, flags, stats);
    return launch_status("building_step_kernel");
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
                          void* stream) {
    SG_CHECK_ARG(m && h_action && h_obs && h_reward && h_done && d_action, "building_step_host: NULL argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t B = (size_t)num_envs, n = (size_t)m->n;
    SG_CUDA(cudaMemcpyAsync(d_action, h_action, B * n * (m->discrete ? 8 : 4), cudaMemcpyHostToDevice, st));
    int rc = sg::dispatch_building(m, num_envs, 1, d_action, d_state, d_epoch, d_runs, nullptr, d_reward,
                                   d_done, 0, nullptr, stream);
    if (rc) return rc;
    SG_CUDA(cudaMemcpyAsync(h_obs, d_state, B * (n + 4) * 4, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(h_reward, d_reward, B * 8, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(h_done, d_done, B, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
