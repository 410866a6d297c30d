// CogenEnv batched step()/reset() for sm_100a.
//
// Replaces sustaingym/envs/cogen/env.py:276-353 (_compute_reward incl. the ONNX Runtime call at
// :305), :232-274 (_dyn_constraint_volation), :145-184 (_forecast_from_time/_get_obs), :355-387
// (step) and sustaingym/envs/cogen/multiagent_env.py:97-101 (per-agent reward split).
//
// The plant model is 23 independent 3-layer MLPs (1.84 M MAC per env-step)
// — a dense contraction, the one compute-bound piece of the whole engine.  Two kernels:
//
// cogen_mlp_tc_kernel (default): layer 2 (>95% of the FLOPs) on the 5th-generation tensor cores.  One
//   CTA = 128 envs x one sub-model, warp-specialised: 8 compute warps (two threads per env) + 1 producer /
//   MMA-issuer warp; per chunk of 16 hidden-1 units the compute threads evaluate layer 1 in fp32, split the
//   activations into tf32 (hi, lo) pairs and store them as the K-major A operand; the pre-split W2 chunk
//   (hi, lo) and the chunk's W1/b1 arrive by bulk asynchronous copies (TMA) into mbarrier-guarded rings
//   (W2: 2 stages tied to MMA completion, W1/b1: 4 slots); the issuer thread issues 3 tcgen05.mma kind::tf32 per K=8
//   (lo*hi, hi*lo, hi*hi: 3xTF32, because a single TF32/BF16 pass cannot meet the 1e-5 tolerance,
//   SURVEY.md H4) accumulating a 128 x H2 fp32 tile in tensor memory; tcgen05.commit -> mbarrier frees
//   the stage; the epilogue reads the accumulator rows with tcgen05.ld and fuses bias, activation and
//   layer 3.
// cogen_mlp_kernel (fallback when the model carries no pre-split weights): fp32 FFMA.  One CTA = 64
//   envs x one sub-model:
//   layer 1 (K = d <= 14) is recomputed per 32-wide chunk of hidden-1 units straight into shared
//   memory (transposed, so the layer-2 A operand is one float4 broadcast per thread);
//   layer 2 (256x128 or 512x256, >95% of the FLOPs) streams its weight rows global->shared with
//   16-byte cp.async, double buffered, and accumulates a 4 env x (H2/16) col register tile/thread;
//   layer 3 (H2 -> 1) is fused into the epilogue with a 16-lane shuffle reduction.
// Weights (7.4 MB) are shared by every env and stay L2-resident.  A second, memory-bound kernel
// (one thread per env) applies the gates, derived outputs, constraint hinges, rewards and writes
// the flattened observation.
#include "common.cuh"
#include "tc.cuh"

namespace sg {
namespace {

constexpr int kTM = 64;        // envs per CTA
constexpr int kKC = 32;        // hidden-1 units per chunk
constexpr int kThreadsC = 256;
constexpr int kNSub = 23;
constexpr int kDPad = 16;
constexpr int kSteps = 96;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float act_fn(float v, int act) {
    // act 0: sigmoid evaluated as the reference graph does, 1/(1+exp(-v)); act 1: relu
    return act == 0 ? __fdividef(1.f, 1.f + expf(-v)) : fmaxf(v, 0.f);
}
// same with the special-function unit: ex2.approx of a pre-scaled argument + rcp.approx (4 instructions);
// absolute error of the sigmoid <= ~4e-7, below the 3xTF32 operand representation error of the tc kernel
__device__ __forceinline__ float act_fn_fast(float v, int act) {
#ifdef SG_COGEN_EXACT_ACT
    return act_fn(v, act);
#else
    if (act != 0) return fmaxf(v, 0.f);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));      // e = +inf -> 0
    return r;
#endif
}
__device__ __forceinline__ float gate_fn(float v) { return v != 0.f ? 1.f : 0.f; }     // |sign(v)|

// extended feature vector E[25] from the 18 model inputs (gen_cogen_pack.py docstring)
__device__ __forceinline__ float ext_feature(const float* x, int f) {
    if (f < 18) return x[f];
    if (f < 21) return x[3 + 3 * (f - 18)] * gate_fn(x[5 + 3 * (f - 18)]);
    if (f < 24) return x[4 + 3 * (f - 21)] * gate_fn(x[5 + 3 * (f - 21)]);
    return x[16] * gate_fn(x[15]);
}

// X row of env b: [TAMB, PAMB, RHAMB](f32 of the ambient row at (day,t)) ++ action[15]
__device__ __forceinline__ void build_x(const sg_cogen_model& m, const float* X_in, const float* action,
                                        const int32_t* day_t, long long b, float* x) {
    if (X_in != nullptr) {
#pragma unroll
        for (int i = 0; i < 18; ++i) x[i] = X_in[b * 18 + i];
        return;
    }
    const int day = day_t[b * 2], t = day_t[b * 2 + 1];
    const double* row = m.ambient + ((long long)day * kSteps + min(t, kSteps - 1)) * 7;
    x[0] = (float)row[0]; x[1] = (float)row[1]; x[2] = (float)row[2];
#pragma unroll
    for (int i = 0; i < 15; ++i) x[3 + i] = action[b * 15 + i];
}

template <int H2>
__device__ __forceinline__ void mlp_tile(const sg_cogen_model& m, const int j, const long long b0, const int nvalid,
                                         const float* X_in, const float* action, const int32_t* day_t,
                                         float* __restrict__ raw, unsigned char* smem) {
    constexpr int NJ = H2 / 64;                 // float4 column groups per thread
    float* xsT = reinterpret_cast<float*>(smem);                // [kDPad][kTM]  normalised inputs, transposed
    float* h1T = xsT + kTM * kDPad;                             // [kKC][kTM]    hidden-1 chunk, transposed
    float* w1s = h1T + kKC * kTM;                               // [kDPad][kKC]  layer-1 weights of the chunk
    float* w2s = w1s + kDPad * kKC;                             // [2][kKC][H2]  layer-2 weight rows (2 stages)
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int d = m.meta[j * 6 + 0], H1 = m.meta[j * 6 + 1], act = m.meta[j * 6 + 3];
    const float* W1 = m.blob + m.offsets[j * 5 + 0];
    const float* b1 = m.blob + m.offsets[j * 5 + 1];
    const float* W2 = m.blob + m.offsets[j * 5 + 2];
    const float* b2 = m.blob + m.offsets[j * 5 + 3];
    const float* W3 = m.blob + m.offsets[j * 5 + 4];

    // normalised inputs of this sub-model for the 64 envs of the tile
    for (int idx = tid; idx < kTM * kDPad; idx += kThreadsC) {
        const int dd = idx / kTM, mm = idx % kTM;
        float v = 0.f;
        if (mm < nvalid && dd < d) {
            float x[18];
            build_x(m, X_in, action, day_t, b0 + mm, x);
            v = (ext_feature(x, m.feats[j * kDPad + dd]) - m.mu[j * kDPad + dd]) * m.sinv[j * kDPad + dd];
        }
        xsT[idx] = v;
    }
    auto issue_w2 = [&](int chunk, int stage) {
        const float* src = W2 + (long long)chunk * kKC * H2;
        float* dst = w2s + stage * kKC * H2;
        for (int i = tid * 4; i < kKC * H2; i += kThreadsC * 4) cp_async16(dst + i, src + i);
        cp_async_commit();
    };
    const int nchunks = H1 / kKC;
    issue_w2(0, 0);
    float acc[4][NJ * 4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < NJ * 4; ++c) acc[i][c] = 0.f;
    __syncthreads();

    for (int ch = 0; ch < nchunks; ++ch) {
        const int stage = ch & 1;
        if (ch + 1 < nchunks) issue_w2(ch + 1, stage ^ 1);
        // layer 1 for hidden units [ch*32, ch*32+32): 64 envs x 32 units, 8 units per thread
        for (int idx = tid; idx < kDPad * kKC; idx += kThreadsC) {
            const int dd = idx / kKC, kk = idx % kKC;
            w1s[idx] = dd < d ? __ldg(W1 + dd * H1 + ch * kKC + kk) : 0.f;
        }
        __syncthreads();
        {
            const int mm = tid & (kTM - 1), kg = (tid >> 6) * 8;
            float xr[kDPad];
#pragma unroll
            for (int dd = 0; dd < kDPad; ++dd) xr[dd] = xsT[dd * kTM + mm];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int kk = kg + i;
                float v = __ldg(b1 + ch * kKC + kk);
#pragma unroll
                for (int dd = 0; dd < kDPad; ++dd) v = fmaf(xr[dd], w1s[dd * kKC + kk], v);
                h1T[kk * kTM + mm] = act_fn(v, act);
            }
        }
        if (ch + 1 < nchunks) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        const float* wb = w2s + stage * kKC * H2;
#pragma unroll 4
        for (int k = 0; k < kKC; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(h1T + k * kTM + ty * 4);
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) {
                const float4 bv = *reinterpret_cast<const float4*>(wb + k * H2 + jj * 64 + tx * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][jj * 4 + 0] = fmaf(av[i], bv.x, acc[i][jj * 4 + 0]);
                    acc[i][jj * 4 + 1] = fmaf(av[i], bv.y, acc[i][jj * 4 + 1]);
                    acc[i][jj * 4 + 2] = fmaf(av[i], bv.z, acc[i][jj * 4 + 2]);
                    acc[i][jj * 4 + 3] = fmaf(av[i], bv.w, acc[i][jj * 4 + 3]);
                }
            }
        }
        __syncthreads();
    }
    // epilogue: h2 = act(acc + b2); y = h2 . W3 + b3, reduced over the 16 column-threads of a row
    float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b2 + jj * 64 + tx * 4));
        const float4 ww = __ldg(reinterpret_cast<const float4*>(W3 + jj * 64 + tx * 4));
        const float bbv[4] = {bb.x, bb.y, bb.z, bb.w}, wwv[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) part[i] = fmaf(act_fn(acc[i][jj * 4 + c] + bbv[c], act), wwv[c], part[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) part[i] += __shfl_xor_sync(kFull, part[i], o);
        const int mm = ty * 4 + i;
        if (tx == 0 && mm < nvalid) raw[(b0 + mm) * kNSub + j] = part[i] + m.b3[j];
    }
}

__global__ void __launch_bounds__(kThreadsC, 2)
cogen_mlp_kernel(const sg_cogen_model m, const long long B, const float* __restrict__ X_in,
                 const float* __restrict__ action, const int32_t* __restrict__ day_t, float* __restrict__ raw) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int j = kNSub - 1 - blockIdx.y;        // big sub-models (packed last) are scheduled first
    const long long b0 = (long long)blockIdx.x * kTM;
    const int nvalid = (int)min((long long)kTM, B - b0);
    if (m.meta[j * 6 + 2] == 256) mlp_tile<256>(m, j, b0, nvalid, X_in, action, day_t, raw, smem);
    else mlp_tile<128>(m, j, b0, nvalid, X_in, action, day_t, raw, smem);
}

// =================================================================================================
// Tensor-core path (tcgen05 + TMEM).  Pre-split weight blob, per sub-model j at tc_offsets[j] (floats):
// H1/16 chunks of  [ W1 chunk [16 inputs][16 units] | b1 chunk [16] | W2 chunk hi | W2 chunk lo ],
// the W2 parts in the K-major no-swizzle UMMA layout [4 K-groups][H2][4]: element (n, k) = W2[k][n].
// =================================================================================================
constexpr int kTcM = 128;      // envs per CTA = UMMA M
constexpr int kTcKC = 16;      // hidden-1 units per chunk (two K=8 instructions)
__host__ __device__ constexpr int tc_chunk_floats(const int H2) { return kDPad * kTcKC + kTcKC + 2 * kTcKC * H2; }

__global__ void cogen_tc_pack_kernel(const sg_cogen_model m, float* __restrict__ out) {
    const int j = blockIdx.x;
    const int d = m.meta[j * 6 + 0], H1 = m.meta[j * 6 + 1], H2 = m.meta[j * 6 + 2];
    const float* W1 = m.blob + m.offsets[j * 5 + 0];
    const float* b1 = m.blob + m.offsets[j * 5 + 1];
    const float* W2 = m.blob + m.offsets[j * 5 + 2];
    float* dst = out + m.tc_offsets[j];
    const int chf = tc_chunk_floats(H2), nch = H1 / kTcKC;
    for (long long idx = threadIdx.x; idx < (long long)nch * chf; idx += blockDim.x) {
        const int c = (int)(idx / chf), r = (int)(idx % chf);
        float v;
        if (r < kDPad * kTcKC) {                                  // W1 chunk [dd][kk]
            const int dd = r / kTcKC, kk = r % kTcKC;
            v = dd < d ? W1[dd * H1 + c * kTcKC + kk] : 0.f;
        } else if (r < kDPad * kTcKC + kTcKC) {                   // b1 chunk
            v = b1[c * kTcKC + (r - kDPad * kTcKC)];
        } else {                                                  // W2 chunk, hi then lo
            int q = r - (kDPad * kTcKC + kTcKC);
            const bool lo_part = q >= kTcKC * H2;
            if (lo_part) q -= kTcKC * H2;
            const int kg = q / (H2 * 4), nn = (q / 4) % H2, ki = q & 3;
            float hi, lo;
            tc::split_tf32(W2[(long long)(c * kTcKC + kg * 4 + ki) * H2 + nn], hi, lo);
            v = lo_part ? lo : hi;
        }
        dst[idx] = v;
    }
}

// Normalised, padded inputs of every sub-model, once per env: xn[j][b][16].  (Building them inside the
// tensor-core CTAs cost ~15k cycles of dependent global loads per CTA -- 40% of a small sub-model's CTA.)
__global__ void __launch_bounds__(128)
cogen_tc_inputs_kernel(const sg_cogen_model m, const long long B, const float* __restrict__ X_in,
                       const float* __restrict__ action, const int32_t* __restrict__ day_t, float* __restrict__ xn) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // thread = (env, sub-model)
    const int j = blockIdx.y;
    if (b >= B) return;
    float x[18];
    build_x(m, X_in, action, day_t, b, x);
    const int d = m.meta[j * 6 + 0];
    float4* dst = reinterpret_cast<float4*>(xn + ((long long)j * B + b) * kDPad);
#pragma unroll
    for (int q = 0; q < kDPad / 4; ++q) {
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int dd = q * 4 + i;
            v[i] = dd < d ? (ext_feature(x, m.feats[j * kDPad + dd]) - m.mu[j * kDPad + dd]) * m.sinv[j * kDPad + dd] : 0.f;
        }
        dst[q] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

constexpr int kTcW1Ring = 4;   // W1 | b1 chunks in flight (their slots are not tied to MMA completion)
#ifndef SG_COGEN_RT
#define SG_COGEN_RT 2
#endif
constexpr int kTcRT = SG_COGEN_RT;            // compute threads per env row (each owns kTcKC / kTcRT hidden units of a chunk)
constexpr int kTcCompute = kTcRT * kTcM;      // compute threads (kTcRT * 4 warps; 2 and 4 per row time the same: the loop is bound by the shared-memory pipe)
constexpr int kTcThreads = kTcCompute + 32;   // + 1 producer / MMA-issuer warp

struct TcSync {
    uint64_t w1full[kTcW1Ring];   // slot q: the chunk's W1 | b1 bulk copy has landed
    uint64_t full[2];             // stage s: the chunk's W2 (hi | lo) bulk copy has landed
    uint64_t a_ready[2];          // stage s: all compute threads have written their part of the A operand
    uint64_t freed[2];            // stage s: the MMAs that read it have completed
    uint64_t done;                // all MMAs of the tile have completed
    uint32_t tmem;                // TMEM base address of the accumulator
};

// Warp-specialised tile: the compute warps evaluate layer 1 -> A operand (tf32 hi | lo) and the epilogue; lane 0 of
// the last warp streams the weights (TMA bulk copies: W1 | b1 chunks 3 ahead, W2 chunks into the 2-stage ring as
// soon as the MMAs that read a stage have completed) and issues the tcgen05.mma instructions.  All hand-offs
// are mbarriers; there is no CTA-wide barrier inside the chunk loop.
template <int H2, int ACT>
__device__ __forceinline__ void mlp_tile_tc(const sg_cogen_model& m, const int j, const long long b0, const int nvalid,
                                            const long long B, const float* __restrict__ xn,
                                            float* __restrict__ raw, unsigned char* smem, TcSync& sy) {
    constexpr int W1F = kDPad * kTcKC + kTcKC;                  // floats of W1 | b1 of one chunk
    constexpr int W2F = 2 * kTcKC * H2;                         // floats of W2 hi | lo of one chunk
    constexpr int CHF = tc_chunk_floats(H2);                    // = W1F + W2F, one chunk in the blob
    constexpr int AF = 2 * kTcM * kTcKC;                        // floats of the A operand hi | lo
    constexpr uint32_t LBO_A = kTcM * 16, LBO_B = H2 * 16, SBO = 128;
    float* const w2base = reinterpret_cast<float*>(smem);       // [2][W2F]
    float* const abase = w2base + 2 * W2F;                      // [2][AF]
    float* const w1r = abase + 2 * AF;                          // [kTcW1Ring][W1F]
    const int tid = threadIdx.x, warp = tid >> 5, row = tid & (kTcM - 1), part = (tid >> 7) & (kTcRT - 1);
    const bool producer = tid >= kTcCompute;
    const int d = m.meta[j * 6 + 0], H1 = m.meta[j * 6 + 1];
    constexpr int act = ACT;
    const int nch = H1 / kTcKC;
    const float* blob = m.tc_blob + m.tc_offsets[j];
#ifdef SG_COGEN_TRACE
    const long long tp0 = clock64();
#endif
    // Layer-1 mapping (independent of the epilogue's TMEM-lane mapping): a thread owns the two env rows
    // ra, ra + 64 and 4 of the chunk's 16 hidden units, so one broadcast LDS.128 of W1 feeds 8 fmas.
    static_assert(kTcCompute == 256 && kTcKC == 16, "layer-1 mapping: 64 row pairs x 4 unit groups");
    const int ra = tid & 63, ug = (tid >> 6) & 3;
    float xa[kDPad], xb[kDPad];
    {   // normalised inputs of my two rows (cogen_tc_inputs_kernel), issued before anything else
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* pa = reinterpret_cast<const float4*>(xn + ((long long)j * B + b0 + ra) * kDPad);
        const bool va = !producer && ra < nvalid, vb = !producer && ra + 64 < nvalid;
#pragma unroll
        for (int q = 0; q < kDPad / 4; ++q) {
            const float4 a4 = va ? __ldg(pa + q) : z4, b4 = vb ? __ldg(pa + 64 * (kDPad / 4) + q) : z4;
            xa[q * 4 + 0] = a4.x; xa[q * 4 + 1] = a4.y; xa[q * 4 + 2] = a4.z; xa[q * 4 + 3] = a4.w;
            xb[q * 4 + 0] = b4.x; xb[q * 4 + 1] = b4.y; xb[q * 4 + 2] = b4.z; xb[q * 4 + 3] = b4.w;
        }
    }
    if (warp == 0) tc::tmem_alloc(&sy.tmem, H2);
#ifdef SG_COGEN_TRACE
    const long long tp1 = clock64();
#endif
    auto bulk = [&](float* dst, const float* src, const uint32_t bytes, uint64_t* mb) {
        const uint32_t mba = tc::smem_u32(mb);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mba), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         tc::smem_u32(dst)), "l"(src), "r"(bytes), "r"(mba) : "memory");
    };
    auto load_w1 = [&](const int c) { bulk(w1r + (c % kTcW1Ring) * W1F, blob + (long long)c * CHF, W1F * 4, &sy.w1full[c % kTcW1Ring]); };
    auto load_w2 = [&](const int c) { bulk(w2base + (c & 1) * W2F, blob + (long long)c * CHF + W1F, W2F * 4, &sy.full[c & 1]); };
    if (tid == kTcCompute) {                                    // the producer thread owns the barriers: init, first loads
        for (int q = 0; q < kTcW1Ring; ++q) tc::mbar_init(&sy.w1full[q], 1);
        for (int s2 = 0; s2 < 2; ++s2) {
            tc::mbar_init(&sy.full[s2], 1); tc::mbar_init(&sy.a_ready[s2], kTcCompute); tc::mbar_init(&sy.freed[s2], 1);
        }
        tc::mbar_init(&sy.done, 1);
        tc::mbar_fence_init();
        for (int c = 0; c < kTcW1Ring - 1 && c < nch; ++c) load_w1(c);
        load_w2(0);
        if (nch > 1) load_w2(1);
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = sy.tmem;
#ifdef SG_COGEN_TRACE
    const long long tp2 = clock64();
#endif

    if (producer) {
        if (tid == kTcCompute) {
            const uint32_t idesc = tc::make_idesc_tf32(kTcM, H2);
#ifdef SG_COGEN_TRACE
            long long tr[32][4]; const long long tr0 = clock64();
#define SG_TR(i) tr[c][i] = clock64() - tr0
#else
#define SG_TR(i)
#endif
            for (int c = 0; c < nch; ++c) {
                const int s = c & 1;
                tc::mbar_wait(&sy.a_ready[s], (c >> 1) & 1);
                SG_TR(0);        // A(c) is written: compute threads are done with W1 slot c % ring
                if (c + kTcW1Ring - 1 < nch) load_w1(c + kTcW1Ring - 1);   // -> slot (c - 1) % ring, released one iteration ago
                tc::mbar_wait(&sy.full[s], (c >> 1) & 1);           // W2(c) landed
                SG_TR(1);
                tc::fence_after_sync();
                const uint32_t ah = tc::smem_u32(abase + s * AF), al = ah + kTcM * kTcKC * 4;
                const uint32_t bh = tc::smem_u32(w2base + s * W2F), bl = bh + kTcKC * H2 * 4;
#pragma unroll
                for (int kk = 0; kk < kTcKC / 8; ++kk) {
                    const uint64_t dah = tc::make_smem_desc(ah + kk * 2 * LBO_A, LBO_A, SBO);
                    const uint64_t dal = tc::make_smem_desc(al + kk * 2 * LBO_A, LBO_A, SBO);
                    const uint64_t dbh = tc::make_smem_desc(bh + kk * 2 * LBO_B, LBO_B, SBO);
                    const uint64_t dbl = tc::make_smem_desc(bl + kk * 2 * LBO_B, LBO_B, SBO);
                    tc::mma_tf32(tmem_d, dal, dbh, idesc, (c | kk) != 0);   // small terms first
                    tc::mma_tf32(tmem_d, dah, dbl, idesc, 1u);
                    tc::mma_tf32(tmem_d, dah, dbh, idesc, 1u);
                }
                tc::mma_commit(&sy.freed[s]);
                if (c == nch - 1) tc::mma_commit(&sy.done);
                SG_TR(2);
                if (c + 2 < nch) {                                  // refill W2 stage s once its MMAs have completed
                    tc::mbar_wait(&sy.freed[s], (c >> 1) & 1);
                    load_w2(c + 2);
                }
                SG_TR(3);
            }
#ifdef SG_COGEN_TRACE
            if (SG_COGEN_TRACE > 1 && blockIdx.x == 7 && (blockIdx.y == 0 || blockIdx.y == 22))
                for (int c = 0; c < nch; ++c)
                    printf("j=%d c=%2d a_ready %6lld full %6lld issued %6lld freed %6lld\n", j, c, tr[c][0], tr[c][1], tr[c][2], tr[c][3]);
#endif
        }
    } else {
        for (int c = 0; c < nch; ++c) {
            const int s = c & 1, q = c % kTcW1Ring;
            tc::mbar_wait(&sy.w1full[q], (c / kTcW1Ring) & 1);      // W1 | b1 of chunk c
            if (c >= 2) tc::mbar_wait(&sy.freed[s], ((c - 2) >> 1) & 1);   // the A operand of stage s is free again
            const float* w1s = w1r + q * W1F;                       // [16][16]
            const float* b1s = w1s + kDPad * kTcKC;
            const float4 bb = *reinterpret_cast<const float4*>(b1s + ug * 4);
            float ha[4] = {bb.x, bb.y, bb.z, bb.w}, hb[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int dd = 0; dd < kDPad; ++dd) {
                if (dd >= 6 && dd >= d) break;                      // CTA-uniform: every sub-model has 6 <= d <= 14 inputs
                const float4 w = *reinterpret_cast<const float4*>(w1s + dd * kTcKC + ug * 4);
                ha[0] = fmaf(xa[dd], w.x, ha[0]); hb[0] = fmaf(xb[dd], w.x, hb[0]);
                ha[1] = fmaf(xa[dd], w.y, ha[1]); hb[1] = fmaf(xb[dd], w.y, hb[1]);
                ha[2] = fmaf(xa[dd], w.z, ha[2]); hb[2] = fmaf(xb[dd], w.z, hb[2]);
                ha[3] = fmaf(xa[dd], w.w, ha[3]); hb[3] = fmaf(xb[dd], w.w, hb[3]);
            }
            float4* a_hi = reinterpret_cast<float4*>(abase + s * AF) + ug * kTcM + ra;   // k-group ug, row ra
            float4* a_lo = a_hi + kTcM * (kTcKC / 4);
            float hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) tc::split_tf32(act_fn_fast(ha[i], act), hi[i], lo[i]);
            a_hi[0] = make_float4(hi[0], hi[1], hi[2], hi[3]);
            a_lo[0] = make_float4(lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) tc::split_tf32(act_fn_fast(hb[i], act), hi[i], lo[i]);
            a_hi[64] = make_float4(hi[0], hi[1], hi[2], hi[3]);
            a_lo[64] = make_float4(lo[0], lo[1], lo[2], lo[3]);
            tc::fence_async_smem();                                 // my part of the A operand -> visible to the tensor core
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&sy.a_ready[s])) : "memory");
        }
        tc::mbar_wait(&sy.done, 0);
        tc::fence_after_sync();
    }
#ifdef SG_COGEN_TRACE
    const long long tp3 = clock64();
#endif
    // epilogue (compute warps): my env's accumulator row, my share of the columns; h2 = act(acc + b2); y = h2 . W3 + b3
    float y = 0.f;
    if (!producer) {
        const float* b2 = m.blob + m.offsets[j * 5 + 3];
        const float* W3 = m.blob + m.offsets[j * 5 + 4];
        for (int c0 = part * (H2 / kTcRT); c0 < (part + 1) * (H2 / kTcRT); c0 += 16) {
            float v[16];
            tc::tmem_ld16(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + c0, v);
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(b2 + c0) + qq);
                const float4 ww = __ldg(reinterpret_cast<const float4*>(W3 + c0) + qq);
                y = fmaf(act_fn_fast(v[qq * 4 + 0] + bb.x, act), ww.x, y);
                y = fmaf(act_fn_fast(v[qq * 4 + 1] + bb.y, act), ww.y, y);
                y = fmaf(act_fn_fast(v[qq * 4 + 2] + bb.z, act), ww.z, y);
                y = fmaf(act_fn_fast(v[qq * 4 + 3] + bb.w, act), ww.w, y);
            }
        }
    }
    float* ypart = abase;                                           // [kTcRT - 1][kTcM] (all MMAs are done: the stages are free)
    if (!producer && part != 0) ypart[(part - 1) * kTcM + row] = y;
    tc::fence_before_sync();
    __syncthreads();
    if (!producer && part == 0 && row < nvalid) {
#pragma unroll
        for (int pp = 0; pp < kTcRT - 1; ++pp) y += ypart[pp * kTcM + row];
        raw[(b0 + row) * kNSub + j] = y + m.b3[j];
    }
    if (warp == 0) tc::tmem_dealloc(tmem_d, H2);
#ifdef SG_COGEN_TRACE
    if (tid == 0 && blockIdx.x % 64 == 7 && (blockIdx.y == 0 || blockIdx.y == 22 || blockIdx.y == 9))
        printf("CTA j=%d tile=%d alloc %lld prologue %lld loop %lld epilogue %lld\n", j, blockIdx.x, tp1 - tp0, tp2 - tp1, tp3 - tp2, clock64() - tp3);
#endif
}

__global__ void __launch_bounds__(kTcThreads, 2)
cogen_mlp_tc_kernel(const sg_cogen_model m, const long long B, const float* __restrict__ xn, float* __restrict__ raw) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ TcSync sy;
    const int j = kNSub - 1 - blockIdx.y;        // big sub-models (packed last) are scheduled first
    const long long b0 = (long long)blockIdx.x * kTcM;
    const int nvalid = (int)min((long long)kTcM, B - b0);
    const bool relu = m.meta[j * 6 + 3] != 0;
    if (m.meta[j * 6 + 2] == 256) {
        if (relu) mlp_tile_tc<256, 1>(m, j, b0, nvalid, B, xn, raw, smem, sy);
        else mlp_tile_tc<256, 0>(m, j, b0, nvalid, B, xn, raw, smem, sy);
    } else {
        if (relu) mlp_tile_tc<128, 1>(m, j, b0, nvalid, B, xn, raw, smem, sy);
        else mlp_tile_tc<128, 0>(m, j, b0, nvalid, B, xn, raw, smem, sy);
    }
}

// Y[29] from the 23 raw sub-model outputs (gates + derived sums; Appendix B)
__device__ __forceinline__ void assemble_y(const sg_cogen_model& m, const float* x, const float* raw, float* Y) {
    const float gGT[3] = {gate_fn(x[5]), gate_fn(x[8]), gate_fn(x[11])};
    const float gHR[3] = {gate_fn(x[12]), gate_fn(x[13]), gate_fn(x[13])};   // sic: HR3's gate reads HR2
    float gates[6];
#pragma unroll
    for (int i = 0; i < 3; ++i) { gates[i] = gGT[i]; gates[3 + i] = gate_fn(gGT[i] + gHR[i]); }
#pragma unroll
    for (int i = 0; i < 29; ++i) Y[i] = 0.f;
    for (int j = 0; j < kNSub; ++j) {
        const int yi = m.meta[j * 6 + 4], g = m.meta[j * 6 + 5];
        Y[yi] = g >= 0 ? raw[j] * gates[g] : raw[j];
    }
    Y[6] = Y[0] + Y[3]; Y[7] = Y[1] + Y[4]; Y[8] = Y[2] + Y[5];
    Y[21] = (Y[6] + Y[7]) + Y[8];
    Y[27] = (((x[5] + x[8]) + x[11]) + x[15]) - Y[26];
    Y[28] = ((x[12] + x[13]) + x[14]) + x[16];
}

__global__ void cogen_plant_finish_kernel(const sg_cogen_model m, const long long B, const float* __restrict__ X,
                                          const float* __restrict__ raw, float* __restrict__ Y) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float x[18], r[kNSub], y[29];
    for (int i = 0; i < 18; ++i) x[i] = X[b * 18 + i];
    for (int i = 0; i < kNSub; ++i) r[i] = raw[b * kNSub + i];
    assemble_y(m, x, r, y);
    for (int i = 0; i < 29; ++i) Y[b * 29 + i] = y[i];
}

// flat observation (gymnasium flatten of the Dict, sorted keys; cogen/env.py:164-184):
// Energy_Price[k+1] | Gas_Price[k+1] | PAMB[k+1] | Prev_Action[32] | RHAMB[k+1] | TAMB[k+1] |
// Target_Power[k+1] | Target_Steam[k+1] | Time[1]; Prev_Action (sorted keys): CT_NrBays one-hot 12,
// then per GT i: EVC one-hot 2, PAC one-hot 2, PWR; HR1..3; IPPROC_M; ST_PWR.
__device__ void write_obs(const sg_cogen_model& m, const int day, const int t, const float* a /*15*/,
                          const double* noise /* [k,7] or NULL */, float* o) {
    const int k = m.forecast_horizon, k1 = k + 1;
    const int col_of_group[7] = {5, 6, 1, 2, 0, 3, 4};       // ambient column of each forecast group, in obs order
    int pos = 0;
    for (int g = 0; g < 7; ++g) {
        if (g == 3) {                                        // Prev_Action sits after PAMB
            const int bays = (int)a[14];
            for (int i = 0; i < 12; ++i) o[pos++] = (i == bays - 1) ? 1.f : 0.f;
            for (int u = 0; u < 3; ++u) {
                const int evc = (int)a[3 * u + 1], pac = (int)a[3 * u + 0];
                o[pos++] = evc == 0 ? 1.f : 0.f; o[pos++] = evc == 1 ? 1.f : 0.f;
                o[pos++] = pac == 0 ? 1.f : 0.f; o[pos++] = pac == 1 ? 1.f : 0.f;
                o[pos++] = a[3 * u + 2];
            }
            o[pos++] = a[9]; o[pos++] = a[10]; o[pos++] = a[11]; o[pos++] = a[13]; o[pos++] = a[12];
        }
        const int c = col_of_group[g];
        for (int r = 0; r < k1; ++r) {
            int tt = t + r, dd = day;
            if (tt >= kSteps) { tt -= kSteps; dd += 1; }     // spill into the next day (env.py:152-155)
            dd = min(dd, m.n_days - 1);                      // the reference raises IndexError here; clamp
            double v = m.ambient[((long long)dd * kSteps + tt) * 7 + c];
            if (noise != nullptr && r >= 1) v += noise[(r - 1) * 7 + c];
            o[pos++] = (float)v;
        }
    }
    o[pos++] = (float)((double)t / (double)kSteps);
}

__global__ void cogen_step_finish_kernel(const sg_cogen_model m, const long long B, const float* __restrict__ action,
                                         float* __restrict__ prev_action, int32_t* __restrict__ day_t,
                                         const double* __restrict__ noise, const float* __restrict__ raw,
                                         float* __restrict__ obs, double* __restrict__ reward,
                                         double* __restrict__ agent_reward, uint8_t* __restrict__ done,
                                         float* __restrict__ Y_out, const int flags, double* __restrict__ stats) {
    __shared__ double sstat[11 * 4];
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    StatsAcc acc;
    if (b < B) {
        const int day = day_t[b * 2], t = day_t[b * 2 + 1];
        float x[18], r[kNSub], Y[29], a[15], pa[15];
        build_x(m, nullptr, action, day_t, b, x);
        for (int i = 0; i < 15; ++i) { a[i] = x[3 + i]; pa[i] = prev_action[b * 15 + i]; }
        for (int i = 0; i < kNSub; ++i) r[i] = raw[b * kNSub + i];
        assemble_y(m, x, r, Y);
        // targets of the current row (obs[...][0], env.py:341-342)
        const double* row = m.ambient + ((long long)day * kSteps + min(t, kSteps - 1)) * 7;
        const float target_power = (float)row[3], target_steam = (float)row[4];
        // ---- reward: env.py:309-346 (dtypes as NumPy >= 2 produces them) ---------------------------
        const float rp = (float)m.ramp_penalty;
        const int pw[4] = {2, 5, 8, 12};
        float ramp[4];
        for (int i = 0; i < 4; ++i) ramp[i] = __fmul_rn(rp, fabsf(__fsub_rn(a[pw[i]], pa[pw[i]])));
        const float total_ramp = __fadd_rn(__fadd_rn(__fadd_rn(ramp[0], ramp[1]), ramp[2]), ramp[3]);
        const int lo[7] = {9, 15, 11, 17, 13, 19, 24}, xi[7] = {5, 12, 8, 13, 11, 14, 15}, hi[7] = {10, 16, 12, 18, 14, 20, 25};
        double cv[16];
        for (int i = 0; i < 7; ++i) {
            cv[2 * i] = (double)fmaxf(0.f, __fsub_rn(Y[lo[i]], x[xi[i]]));
            cv[2 * i + 1] = (double)fmaxf(0.f, __fsub_rn(x[xi[i]], Y[hi[i]]));
        }
        cv[14] = (double)fmaxf(0.f, __fsub_rn(x[16], Y[22]));
        cv[15] = (double)fmaxf(0.f, __fsub_rn(x[16], Y[23]));
        double dyn[4];
        for (int u = 0; u < 4; ++u) dyn[u] = m.constraint_violation_penalty * (((cv[4 * u] + cv[4 * u + 1]) + cv[4 * u + 2]) + cv[4 * u + 3]);
        const double total_dyn = ((dyn[0] + dyn[1]) + dyn[2]) + dyn[3];
        const float steam_pen = fmaxf(0.f, __fsub_rn(target_steam, Y[28]));
        const float energy_pen = fmaxf(0.f, __fsub_rn(target_power, Y[27]));
        const float non_delivery = __fmul_rn((float)m.supply_imbalance_penalty, __fadd_rn(steam_pen, energy_pen));
        const double rew = -((double)__fadd_rn(__fadd_rn(Y[21], total_ramp), non_delivery) + total_dyn);
        const float fuel[4] = {Y[6], Y[7], Y[8], 0.f};
        const double nd4 = (double)__fdiv_rn(non_delivery, 4.f);
        for (int u = 0; u < 4; ++u)
            agent_reward[b * 4 + u] = -((((double)fuel[u] + (double)ramp[u]) + dyn[u]) + nd4);   // multiagent_env.py:97-101
        reward[b] = rew;
        if (Y_out != nullptr) for (int i = 0; i < 29; ++i) Y_out[b * 29 + i] = Y[i];
        // ---- advance: env.py:373-382 ------------------------------------------------------------
        const int tn = t + 1;
        const bool term = tn >= kSteps;
        for (int i = 0; i < 15; ++i) prev_action[b * 15 + i] = a[i];
        const int obs_dim = 7 * (m.forecast_horizon + 1) + 33;
        write_obs(m, day, tn, a, noise ? noise + b * m.forecast_horizon * 7 : nullptr, obs + b * obs_dim);
        day_t[b * 2 + 1] = tn;
        done[b] = term ? 1 : 0;
        acc.add(rew, term, -(double)Y[21], -(double)total_ramp, -((double)non_delivery + total_dyn));
    }
    if (stats != nullptr) {
        __syncthreads();
        stats_flush_block(acc, stats, sstat);
    }
}

__global__ void cogen_reset_kernel(const sg_cogen_model m, const long long B, const int32_t* __restrict__ day_in,
                                   const float* __restrict__ init_action, float* __restrict__ prev_action,
                                   int32_t* __restrict__ day_t, float* __restrict__ obs) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int day = min(max(day_in[b], 0), m.n_days - 1);
    float a[15];
    for (int i = 0; i < 15; ++i) { a[i] = init_action[b * 15 + i]; prev_action[b * 15 + i] = a[i]; }
    day_t[b * 2] = day; day_t[b * 2 + 1] = 0;
    const int obs_dim = 7 * (m.forecast_horizon + 1) + 33;
    write_obs(m, day, 0, a, nullptr, obs + b * obs_dim);
}

int check_cogen(const sg_cogen_model* m) {
    SG_CHECK_ARG(m != nullptr, "cogen: model is NULL");
    SG_CHECK_ARG(m->n_sub == kNSub, "cogen: expected 23 sub-models, got %d", m->n_sub);
    SG_CHECK_ARG(m->forecast_horizon >= 0 && m->forecast_horizon < kSteps - 1, "cogen: forecast_horizon out of range");
    SG_CHECK_ARG(m->blob && m->offsets && m->meta && m->feats && m->mu && m->sinv && m->b3, "cogen: NULL weight table");
    return 0;
}

int launch_mlp(const sg_cogen_model* m, long long B, const float* X, const float* action, const int32_t* day_t,
               float* raw, cudaStream_t st) {
    if (m->tc_blob != nullptr && m->tc_offsets != nullptr) {     // tensor-core path
        const size_t smem_tc = sizeof(float) * (2 * (2 * kTcKC * 256 + 2 * kTcM * kTcKC) + kTcW1Ring * (kDPad * kTcKC + kTcKC));
        // per-device attribute: set on every launch (a process may drive several GPUs)
        SG_CUDA(cudaFuncSetAttribute(cogen_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc));
        float* xn = raw + ((B * kNSub + 3) & ~3LL);               // scratch = raw [B,23] | xn [23,B,16] (16-byte aligned)
        cogen_tc_inputs_kernel<<<dim3((unsigned)((B + 127) / 128), kNSub), 128, 0, st>>>(*m, B, X, action, day_t, xn);
        dim3 grid((unsigned)((B + kTcM - 1) / kTcM), kNSub);
        cogen_mlp_tc_kernel<<<grid, kTcThreads, smem_tc, st>>>(*m, B, xn, raw);
        return launch_status("cogen_mlp_tc_kernel");
    }
    const size_t smem = sizeof(float) * (kTM * kDPad + kKC * kTM + kDPad * kKC + 2 * kKC * 256);
    SG_CUDA(cudaFuncSetAttribute(cogen_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((B + kTM - 1) / kTM), kNSub);
    cogen_mlp_kernel<<<grid, kThreadsC, smem, st>>>(*m, B, X, action, day_t, raw);
    return launch_status("cogen_mlp_kernel");
}

}  // namespace
}  // namespace sg

extern "C" {

int sg_cogen_plant(const sg_cogen_model* m, int64_t num_envs, const float* X, float* Y, float* scratch, void* stream) {
    if (int rc = sg::check_cogen(m)) return rc;
    SG_CHECK_ARG(X && Y && scratch, "cogen_plant: NULL argument");
    if (num_envs <= 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (int rc = sg::launch_mlp(m, num_envs, X, nullptr, nullptr, scratch, st)) return rc;
    sg::cogen_plant_finish_kernel<<<(unsigned)((num_envs + 127) / 128), 128, 0, st>>>(*m, num_envs, X, scratch, Y);
    return sg::launch_status("cogen_plant_finish_kernel");
}

int sg_cogen_prepare_tc(const sg_cogen_model* m, float* tc_blob, void* stream) {
    if (int rc = sg::check_cogen(m)) return rc;
    SG_CHECK_ARG(tc_blob != nullptr && m->tc_offsets != nullptr, "cogen_prepare_tc: NULL argument");
    SG_CHECK_ARG((reinterpret_cast<uintptr_t>(tc_blob) & 15u) == 0, "cogen_prepare_tc: tc_blob must be 16-byte aligned");
    sg::cogen_tc_pack_kernel<<<sg::kNSub, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*m, tc_blob);
    return sg::launch_status("cogen_tc_pack_kernel");
}

int sg_cogen_reset(const sg_cogen_model* m, int64_t num_envs, const int32_t* day_in, const float* init_action,
                   float* prev_action, int32_t* day_t, float* obs, void* stream) {
    if (int rc = sg::check_cogen(m)) return rc;
    SG_CHECK_ARG(day_in && init_action && prev_action && day_t && obs && m->ambient, "cogen_reset: NULL argument");
    if (num_envs <= 0) return 0;
    sg::cogen_reset_kernel<<<(unsigned)((num_envs + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        *m, num_envs, day_in, init_action, prev_action, day_t, obs);
    return sg::launch_status("cogen_reset_kernel");
}

int sg_cogen_step(const sg_cogen_model* m, int64_t num_envs, const float* action, float* prev_action,
                  int32_t* day_t, const double* noise, float* obs, double* reward, double* agent_reward,
                  uint8_t* done, float* Y_out, float* scratch, int32_t flags, double* stats, void* stream) {
    if (int rc = sg::check_cogen(m)) return rc;
    SG_CHECK_ARG(action && prev_action && day_t && obs && reward && agent_reward && done && scratch && m->ambient,
                 "cogen_step: NULL argument");
    if (num_envs <= 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (int rc = sg::launch_mlp(m, num_envs, nullptr, action, day_t, scratch, st)) return rc;
    sg::cogen_step_finish_kernel<<<(unsigned)((num_envs + 127) / 128), 128, 0, st>>>(
        *m, num_envs, action, prev_action, day_t, noise, scratch, obs, reward, agent_reward, done, Y_out, flags, stats);
    return sg::launch_status("cogen_step_finish_kernel");
}

}  // extern "C"
