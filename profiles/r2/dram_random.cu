// Microbenchmark: throughput of scattered DRAM accesses on B200 — random 32-byte sectors vs runs of adjacent sectors,
// reads and read-modify-writes, at the concurrency of the EV step kernel and at full occupancy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_bin/dram_random profiles/r2/dram_random.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long mix(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}

// each thread does `iters` accesses; access = `run` adjacent 32-byte sectors at a random 32*run-aligned offset
template <int RUN, bool WRITE>
__global__ void k(uint4* buf, unsigned long long nchunks, int iters, unsigned long long seed, unsigned* sink) {
    unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        unsigned long long c = mix(tid * 0x9E3779B97F4A7C15ULL + it + seed) % nchunks;
        uint4* p = buf + c * (2 * RUN);
#pragma unroll
        for (int r = 0; r < RUN; ++r) {
            uint4 a = p[2 * r], b = p[2 * r + 1];
            acc += a.x + b.w;
            if (WRITE) { a.x += 1; p[2 * r] = a; }
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int RUN, bool WRITE>
void run(uint4* buf, size_t bytes, int blocks, int threads, int iters, unsigned* sink, const char* label) {
    unsigned long long nchunks = bytes / (32ull * RUN);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<RUN, WRITE><<<blocks, threads>>>(buf, nchunks, iters, 1, sink);
    cudaEventRecord(e0);
    for (int rep = 0; rep < 5; ++rep) k<RUN, WRITE><<<blocks, threads>>>(buf, nchunks, iters, 7 + rep, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double acc = (double)blocks * threads * iters;
    printf("%-28s run=%d sectors %s blocks=%5d thr=%4d: %7.3f ms  %7.2f G accesses/s  %7.2f G sectors/s  %7.1f GB/s\n", label, RUN,
           WRITE ? "rmw " : "read", blocks, threads, ms, acc / ms / 1e6, acc * RUN / ms / 1e6, acc * RUN * 32 / ms / 1e6);
}

// warp-coalesced: the warp reads WSEC adjacent sectors (lane l -> sector l % WSEC of chunk l / WSEC) at random chunk offsets
template <int WSEC>
__global__ void kw(uint4* buf, unsigned long long nchunks, int iters, unsigned long long seed, unsigned* sink) {
    unsigned long long wid = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) / WSEC;
    const int l = threadIdx.x % WSEC;
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        unsigned long long c = mix(wid * 0x9E3779B97F4A7C15ULL + it + seed) % nchunks;
        uint4* p = buf + (c * WSEC + l) * 2;
        uint4 a = p[0], b = p[1];
        acc += a.x + b.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}
template <int WSEC>
void runw(uint4* buf, size_t bytes, int blocks, int threads, int iters, unsigned* sink) {
    unsigned long long nchunks = bytes / (32ull * WSEC);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kw<WSEC><<<blocks, threads>>>(buf, nchunks, iters, 1, sink);
    cudaEventRecord(e0);
    for (int rep = 0; rep < 5; ++rep) kw<WSEC><<<blocks, threads>>>(buf, nchunks, iters, 7 + rep, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double acc = (double)blocks * threads * iters;
    printf("lanes share %4d-byte chunks     read blocks=%5d thr=%4d: %7.3f ms  %7.2f G sectors/s  %7.1f GB/s\n", WSEC * 32, blocks, threads, ms,
           acc / ms / 1e6, acc * 32 / ms / 1e6);
}

int main() {
    uint4* buf; unsigned* sink;
    cudaMalloc(&buf, 4ull << 30); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, 4ull << 30);
    const size_t sizes[3] = {192ull << 20, 1ull << 30, 4ull << 30};
    for (size_t bytes : sizes) {
        printf("---- buffer %zu MB (L2 126 MB, TLB reach ~256 MB)\n", bytes >> 20);
        const int cfg[2][2] = {{148 * 3, 128}, {148 * 16, 128}};
        for (auto& c : cfg) {
            int iters = 64;
            run<1, false>(buf, bytes, c[0], c[1], iters, sink, "per-lane random sector");
            run<8, false>(buf, bytes, c[0], c[1], iters / 2, sink, "per-lane random 256 B");
            run<1, true>(buf, bytes, c[0], c[1], iters, sink, "per-lane random sector");
            runw<2>(buf, bytes, c[0], c[1], iters, sink);
            runw<8>(buf, bytes, c[0], c[1], iters, sink);
            runw<32>(buf, bytes, c[0], c[1], iters, sink);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
