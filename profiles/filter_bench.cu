// filter_bench.cu -- what does the probe stage's memory system cost at its floor?
// Every thread issues, per iteration, one 4-byte look-up into a presence filter of F bytes (meant to live in L2) and, with
// probability PASS, one 32-byte random bucket read from a table of T bytes (streams from HBM).  Sweeps F and the L2
// policy (plain / evict_last hint / persisting access-policy window) so that the filter size the probe kernel uses is a
// measured choice and not a guess about how much of the 126 MB L2 a hot structure can keep while 1.2 GB streams by.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o filter_bench filter_bench.cu && ./filter_bench
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33; return h; }

// POLICY 0: ld.global.nc  1: ld.global.nc + evict_last hint  (the persisting window is set by the host on the stream)
template <int POLICY>
__global__ void probe(const uint32_t *filter, uint64_t fwords, const uint4 *table, uint64_t nbuckets, uint32_t pass_thresh,
                      uint64_t per_thread, uint64_t *out)
{
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    uint64_t pol_last, pol_first;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    for (uint64_t i = 0; i < per_thread; i++) {
        const uint64_t h = mix(tid * per_thread + i + 1);
        const uint32_t *fp = filter + __umul64hi(h, fwords);
        uint32_t w;
        if (POLICY == 1) asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(w) : "l"(fp), "l"(pol_last));
        else asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(w) : "l"(fp));
        acc += w;
        if ((uint32_t)(h >> 7) + (w & 1) < pass_thresh) { // w & 1: make the bucket read depend on the filter word
            const uint4 *p = table + 2 * __umul64hi(h * 0x9E3779B97F4A7C15ULL, nbuckets);
            uint64_t a, b, c, d;
            asm volatile("ld.global.nc.L2::cache_hint.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p), "l"(pol_first));
            acc += a ^ b ^ c ^ d;
        }
    }
    if (acc == 0x1234567) out[0] = acc;
}

int main(int argc, char **argv)
{
    const double table_gb = argc > 1 ? atof(argv[1]) : 1.2;
    const double pass = argc > 2 ? atof(argv[2]) : 0.36;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("L2 %d MB, persisting max %d MB, window max %d MB\n", prop.l2CacheSize >> 20, prop.persistingL2CacheMaxSize >> 20, prop.accessPolicyMaxWindowSize >> 20);
    const size_t tbytes = (size_t)(table_gb * 1e9) & ~(size_t)31;
    uint4 *table; uint32_t *filter; uint64_t *out;
    const size_t fmax = 256u << 20;
    cudaMalloc(&table, tbytes); cudaMalloc(&filter, fmax); cudaMalloc(&out, 8);
    cudaMemset(table, 1, tbytes); cudaMemset(filter, 0, fmax);
    cudaStream_t s; cudaStreamCreate(&s);
    const int threads = 256, grid = 148 * 8;
    const uint64_t per_thread = 2048; // 620 M look-ups per launch
    const uint32_t thresh = (uint32_t)(pass * 4294967295.0);
    for (int mode = 0; mode < 3; mode++) {
        for (int fmb = 4; fmb <= 128; fmb *= 2) {
            const size_t fbytes = (size_t)fmb << 20;
            if (mode == 2) {
                size_t want = fbytes < (size_t)prop.persistingL2CacheMaxSize ? fbytes : (size_t)prop.persistingL2CacheMaxSize;
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
                cudaStreamAttrValue a{};
                a.accessPolicyWindow.base_ptr = filter;
                a.accessPolicyWindow.num_bytes = fbytes < (size_t)prop.accessPolicyMaxWindowSize ? fbytes : (size_t)prop.accessPolicyMaxWindowSize;
                a.accessPolicyWindow.hitRatio = 1.0f;
                a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                cudaError_t e = cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &a);
                if (e != cudaSuccess) printf("window: %s\n", cudaGetErrorString(e));
            }
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            auto launch = [&](uint64_t pt) {
                if (mode == 1) probe<1><<<grid, threads, 0, s>>>(filter, fbytes / 4, table, tbytes / 32, thresh, pt, out);
                else probe<0><<<grid, threads, 0, s>>>(filter, fbytes / 4, table, tbytes / 32, thresh, pt, out);
            };
            launch(256);
            cudaEventRecord(a, s);
            launch(per_thread);
            cudaEventRecord(b, s);
            cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            const double n = (double)grid * threads * per_thread;
            printf("mode %d (%s) filter %3d MB: %7.2f G filter look-ups/s, %6.2f G bucket reads/s  (%.2f ms for %.0f M look-ups)\n", mode,
                   mode == 0 ? "plain" : mode == 1 ? "evict_last" : "persisting window", fmb, n / ms / 1e6, n * pass / ms / 1e6, ms, n / 1e6);
            if (mode == 2) {
                cudaStreamAttrValue z{};
                z.accessPolicyWindow.num_bytes = 0;
                cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &z);
                cudaCtxResetPersistingL2Cache();
            }
        }
    }
    // the filter alone (no bucket reads): the L2 look-up rate
    for (int fmb = 4; fmb <= 128; fmb *= 2) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        probe<0><<<grid, threads, 0, s>>>(filter, ((size_t)fmb << 20) / 4, table, tbytes / 32, 0, 256, out);
        cudaEventRecord(a, s);
        probe<0><<<grid, threads, 0, s>>>(filter, ((size_t)fmb << 20) / 4, table, tbytes / 32, 0, per_thread, out);
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double n = (double)grid * threads * per_thread;
        printf("filter only %3d MB: %7.2f G look-ups/s\n", fmb, n / ms / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
