// CogenEnv plant model on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// sg_tc_gemm_selftest: one 128 x N x K GEMM tile through exactly the machinery the plant-model kernel
// uses (3xTF32 split operands in the K-major no-swizzle shared-memory layout, tcgen05.mma kind::tf32
// accumulating in tensor memory, tcgen05.ld read-back) — the unit test of the descriptor encodings.
#include "common.cuh"
#include "tc.cuh"

namespace sg {
namespace {

// operand tile in shared memory: [K/4][rows][4] tf32 — a column of 8-row x 16-byte core matrices per K-group
__device__ __forceinline__ int op_index(const int row, const int k, const int rows) { return ((k >> 2) * rows + row) * 4 + (k & 3); }

template <int N>
__global__ void __launch_bounds__(128)
tc_gemm_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ D, const int K) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* a_hi = reinterpret_cast<float*>(smem);            // [K/4][128][4]
    float* a_lo = a_hi + 128 * K;
    float* b_hi = a_lo + 128 * K;                            // [K/4][N][4]
    float* b_lo = b_hi + N * K;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tmem_slot, N < 32 ? 32 : N);
    if (tid == 0) { tc::mbar_init(&mbar, 1); tc::mbar_fence_init(); }
    for (int idx = tid; idx < 128 * K; idx += 128) {
        const int r = idx / K, k = idx % K;
        float hi, lo;
        tc::split_tf32(A[idx], hi, lo);
        a_hi[op_index(r, k, 128)] = hi; a_lo[op_index(r, k, 128)] = lo;
    }
    for (int idx = tid; idx < N * K; idx += 128) {
        const int r = idx / K, k = idx % K;
        float hi, lo;
        tc::split_tf32(Bm[idx], hi, lo);
        b_hi[op_index(r, k, N)] = hi; b_lo[op_index(r, k, N)] = lo;
    }
    tc::fence_async_smem();                                  // generic-proxy writes -> visible to the tensor core
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(128, N);
        const uint32_t lbo_a = 128 * 16, lbo_b = N * 16, sbo = 128;
        uint32_t acc = 0;
        for (int kk = 0; kk < K / 8; ++kk) {                 // one instruction = 8 K values = 2 K-groups
            const uint64_t dah = tc::make_smem_desc(tc::smem_u32(a_hi) + kk * 2 * lbo_a, lbo_a, sbo);
            const uint64_t dal = tc::make_smem_desc(tc::smem_u32(a_lo) + kk * 2 * lbo_a, lbo_a, sbo);
            const uint64_t dbh = tc::make_smem_desc(tc::smem_u32(b_hi) + kk * 2 * lbo_b, lbo_b, sbo);
            const uint64_t dbl = tc::make_smem_desc(tc::smem_u32(b_lo) + kk * 2 * lbo_b, lbo_b, sbo);
            tc::mma_tf32(tmem_d, dal, dbh, idesc, acc); acc = 1;   // small terms first
            tc::mma_tf32(tmem_d, dah, dbl, idesc, acc);
            tc::mma_tf32(tmem_d, dah, dbh, idesc, acc);
        }
        tc::mma_commit(&mbar);
    }
    tc::mbar_wait(&mbar, 0);
    tc::fence_after_sync();
    // thread t reads row t: lanes of its warp's TMEM quarter, 16 columns at a time
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) D[tid * N + c0 + i] = v[i];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, N < 32 ? 32 : N);
}

}  // namespace
}  // namespace sg

extern "C" int sg_tc_gemm_selftest(const float* A, const float* Bm, float* D, int32_t N, int32_t K, void* stream) {
    SG_CHECK_ARG(A && Bm && D, "tc_gemm_selftest: NULL argument");
    SG_CHECK_ARG((N == 128 || N == 256) && K % 8 == 0 && K >= 8 && K <= 64, "tc_gemm_selftest: N in {128,256}, K multiple of 8 <= 64");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)(2 * 128 * K + 2 * N * K) * sizeof(float);
    if (N == 128) {
        SG_CUDA(cudaFuncSetAttribute(sg::tc_gemm_selftest_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sg::tc_gemm_selftest_kernel<128><<<1, 128, smem, st>>>(A, Bm, D, K);
    } else {
        SG_CUDA(cudaFuncSetAttribute(sg::tc_gemm_selftest_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sg::tc_gemm_selftest_kernel<256><<<1, 128, smem, st>>>(A, Bm, D, K);
    }
    return sg::launch_status("tc_gemm_selftest_kernel");
}
