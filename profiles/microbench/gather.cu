// gather.cu — microbenchmark: random 8-byte gathers from a 800 MB array (the read phase's access pattern).
// Question: how many DRAM bytes does one random 8 B gather cost on B200, and do load qualifiers or
// cudaLimitMaxL2FetchGranularity change it?   usage: gather [limit_bytes]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__device__ __forceinline__ uint64_t mix(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k; }
template <int V> __device__ __forceinline__ double ld(const double* p) {
    if (V == 0) return *p;
    if (V == 1) return __ldg(p);
    if (V == 2) return __ldcg(p);
    if (V == 3) return __ldcs(p);
    if (V == 4) return __ldlu(p);
    if (V == 5) { double r; asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p)); return r; }
    if (V == 6) { double r; asm volatile("ld.global.nc.L1::no_allocate.L2::64B.f64 %0, [%1];" : "=d"(r) : "l"(p)); return r; }
    if (V == 7) { double r; asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(r) : "l"(p)); return r; }
    return 0;
}
template <int V> __global__ void gather(const double* __restrict__ a, uint64_t n, uint64_t per, double* out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0;
    for (uint64_t i = 0; i < per; i += 4) {
        const uint64_t i0 = mix(t * per + i) % n, i1 = mix(t * per + i + 1) % n, i2 = mix(t * per + i + 2) % n, i3 = mix(t * per + i + 3) % n;
        acc += ld<V>(a + i0) + ld<V>(a + i1) + ld<V>(a + i2) + ld<V>(a + i3);
    }
    if (acc == 12345.678) out[0] = acc;
}
template <int V> void run(const double* a, uint64_t n, double* out, const char* name) {
    const uint64_t threads = 148ull * 2048 * 8, per = 64;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<V><<<threads / 256, 256>>>(a, n, per, out);
    cudaEventRecord(e0);
    gather<V><<<threads / 256, 256>>>(a, n, per, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double g = (double)threads * per;
    printf("%-28s %8.3f ms  %7.2f Ggather/s  => %6.1f B/gather at 6555 GB/s if DRAM bound\n", name, ms, g / ms / 1e6, 6555e9 * ms * 1e-3 / g);
}
int main(int argc, char** argv) {
    if (argc > 1) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[1])); printf("set limit %s -> %s\n", argv[1], cudaGetErrorString(e)); }
    size_t lim = 0; cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit = %zu\n", lim);
    const uint64_t n = 100000000;
    double *a, *out; cudaMalloc(&a, n * 8); cudaMalloc(&out, 8); cudaMemset(a, 0, n * 8);
    run<0>(a, n, out, "ld.global"); run<1>(a, n, out, "ld.global.nc (__ldg)"); run<2>(a, n, out, "ld.global.cg"); run<3>(a, n, out, "ld.global.cs");
    run<4>(a, n, out, "ld.global.lu"); run<5>(a, n, out, "L1::no_allocate"); run<6>(a, n, out, "nc.L1::no_allocate.L2::64B"); run<7>(a, n, out, "nc.L1::evict_last");
    return 0;
}
