// gather_bench.cu -- random-access HBM microbenchmark (SURVEY 8d asks for a measured random-sector peak).
// Each thread loads `BYTES` contiguous, BYTES-aligned bytes at a pseudo-random position of a large buffer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu && ./gather_bench [GiB] [granularity]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33; return h; }

template <int BYTES, int MODE> // MODE 0: ld.global  1: ld.global.nc  2: ld.global.nc.L2::64B  3: ld.global.nc.L1::no_allocate
__global__ void gather(const uint4 *buf, uint64_t nblocks, uint64_t per_thread, uint64_t *out)
{
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (uint64_t i = 0; i < per_thread; i++) {
        uint64_t b = __umul64hi(mix(tid * per_thread + i + 1), nblocks);
        const uint4 *p = buf + b * (BYTES / 16);
#pragma unroll
        for (int k = 0; k < BYTES / 16; k++) {
            uint4 v;
            if (MODE == 0) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + k));
            if (MODE == 1) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + k));
            if (MODE == 2) asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + k));
            if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + k));
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    if (acc == 0x1234567) out[0] = acc;
}

template <int BYTES, int MODE>
static void run(const uint4 *buf, size_t bytes, uint64_t *out, const char *name)
{
    const uint64_t nblocks = bytes / BYTES;
    const int threads = 256, grid = 148 * 8 * 4;
    const uint64_t per_thread = 256;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    gather<BYTES, MODE><<<grid, threads>>>(buf, nblocks, 16, out);
    cudaEventRecord(a);
    gather<BYTES, MODE><<<grid, threads>>>(buf, nblocks, per_thread, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double n = (double)grid * threads * per_thread;
    printf("%-28s access %3d B: %8.1f M accesses/ms  %7.1f GB/s useful  (%.3f ms)\n", name, BYTES, n / ms / 1e6, n * BYTES / ms / 1e6, ms);
}

int main(int argc, char **argv)
{
    double gib = argc > 1 ? atof(argv[1]) : 4.0;
    int gran = argc > 2 ? atoi(argv[2]) : 0;
    if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); printf("set L2 fetch granularity %d -> %s\n", gran, cudaGetErrorString(e)); }
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity = %zu\n", g);
    size_t bytes = (size_t)(gib * (1ull << 30));
    uint4 *buf; uint64_t *out;
    cudaMalloc(&buf, bytes); cudaMalloc(&out, 8);
    cudaMemset(buf, 1, bytes);
    run<32, 0>(buf, bytes, out, "ld.global");
    run<32, 1>(buf, bytes, out, "ld.global.nc");
    run<32, 2>(buf, bytes, out, "ld.global.nc.L2::64B");
    run<32, 3>(buf, bytes, out, "ld.global.nc.L1::no_allocate");
    run<64, 1>(buf, bytes, out, "ld.global.nc");
    run<64, 2>(buf, bytes, out, "ld.global.nc.L2::64B");
    run<128, 1>(buf, bytes, out, "ld.global.nc");
    run<16, 1>(buf, bytes, out, "ld.global.nc");
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
