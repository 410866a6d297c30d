// Minimal tcgen05 / TMEM helpers for sm_100a (5th-generation tensor cores), used by the CogenEnv plant
// model: kind::tf32 UMMA with both operands in shared memory (K-major, no swizzle) and the fp32
// accumulator in tensor memory.  Bit layouts follow the PTX ISA tcgen05 descriptors (cross-checked
// against cute/arch/mma_sm100_desc.hpp of the CUTLASS tree shipped in this image).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sg {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor, K-major operand, SWIZZLE_NONE ("interleaved" canonical layout):
// the operand is a grid of 8-row x 16-byte core matrices (8 x 4 tf32), each stored as 128 contiguous
// bytes; `sbo` = byte distance between core matrices adjacent along M/N, `lbo` = byte distance between
// the two core matrices adjacent along K that one K=8 instruction consumes.
__device__ __forceinline__ uint64_t make_smem_desc(const uint32_t saddr, const uint32_t lbo, const uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);            // start address        bits [0,14)
    d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;        // leading byte offset  bits [16,30)
    d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;        // stride byte offset   bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    return d;                                           // base offset 0, lbo mode 0, layout type 0 = no swizzle
}

// Instruction descriptor: D = F32, A = B = TF32, both K-major, dense, M x N tile.
__device__ __forceinline__ uint32_t make_idesc_tf32(const int M, const int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T   (one thread issues for the CTA)
__device__ __forceinline__ void mma_tf32(const uint32_t tmem_d, const uint64_t desc_a, const uint64_t desc_b,
                                         const uint32_t idesc, const uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// all previously issued MMAs of this thread -> one arrival on the mbarrier when they have completed
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* mbar, const uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, const uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(mbar)), "r"(parity)
            : "memory");
    } while (!done);
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMEM allocation: one full warp; `ncols` a power of two in [32, 512]; the address lands in *slot (shared)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, const uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(const uint32_t taddr, const uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// accumulator read-back: the 32 lanes of this warp's TMEM quarter x 16 consecutive columns (one row per thread)
__device__ __forceinline__ void tmem_ld16(const uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// fp32 -> (hi, lo) tf32 pair with hi + lo == v to ~2^-22 relative: the 3xTF32 split
__device__ __forceinline__ void split_tf32(const float v, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    hi = __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hi));
    lo = __uint_as_float(l);
}

}  // namespace tc
}  // namespace sg
