// wavefront.cu — microbenchmark: what bounds divergent one-byte gathers on an SM?
// The prefiltered read phase gathers one key byte per edge from an L2-resident column and runs at ~170-195 G
// gathers/s; the model in DESIGN.md says the ceiling is the L1 wavefront rate (one tag lookup per clock and SM = 148 x 1.965 GHz =
// 290 G/s).  This program measures the pieces of that model so that the next kernel shape is chosen on numbers:
//   (1) LDG.U8 gathers, fully divergent, against window size (32 KB = L1-resident ... 50 MB = L2-resident ... 800 MB = DRAM)
//   (2) the same with 2 / 4 / 8 / 32 lanes sharing a 32 B sector, and with lanes sharing a 128 B line but not a sector
//       (is a wavefront a sector or a line?)
//   (3) the same gathers through the texture path (tex1Dfetch<unsigned char> on a linear texture): separate address path,
//       historically a higher divergent-fetch rate than the LSU
//   (4) shared-memory gathers of one byte (the upper bound if a key block could live in shared memory)
//   usage: wavefront            (prints a table; a few seconds)
// Written in round 1 after the GPU budget was spent: NOT RUN YET.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
__host__ __device__ __forceinline__ uint64_t mix(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k; }

__device__ __forceinline__ uint32_t ld_u8_keep(const uint8_t* p, uint64_t pol) {
    uint32_t r; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol)); return r;
}
__device__ __forceinline__ uint32_t ld_u8_l1(const uint8_t* p) { return __ldg(p); }

// address of gather g of thread t: `share` consecutive lanes fall into one unit of `unit` bytes (unit = 32: a sector, 128: a line),
// each on its own byte; with spread = 1 the lanes sharing a 128 B line are put on different sectors of it.
__device__ __forceinline__ uint64_t gather_index(uint64_t t, uint32_t g, uint64_t window, uint32_t share, uint32_t unit, uint32_t spread) {
    const uint32_t lane = (uint32_t)(t & 31);
    const uint64_t group = (t >> 5) * 32 + (lane / share) * share;          // the same for the lanes that share
    uint64_t base = __umul64hi(mix(group * 977 + g), window / unit) * unit;  // random unit inside the window
    const uint32_t within = lane % share;
    return base + (spread ? (within * 32u) % unit + within / 4u : within % unit);
}

template <int MODE>   // 0: LDG no_allocate + evict_last, 1: LDG via L1, 2: texture
__global__ void __launch_bounds__(256) gather_kernel(const uint8_t* __restrict__ a, cudaTextureObject_t tex, uint64_t window, uint32_t per, uint32_t share, uint32_t unit,
                                                     uint32_t spread, uint32_t* out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    uint32_t acc = 0;
    for (uint32_t g = 0; g < per; g += 4) {
        const uint64_t i0 = gather_index(t, g, window, share, unit, spread), i1 = gather_index(t, g + 1, window, share, unit, spread),
                       i2 = gather_index(t, g + 2, window, share, unit, spread), i3 = gather_index(t, g + 3, window, share, unit, spread);
        if (MODE == 0) acc += ld_u8_keep(a + i0, pol) + ld_u8_keep(a + i1, pol) + ld_u8_keep(a + i2, pol) + ld_u8_keep(a + i3, pol);
        else if (MODE == 1) acc += ld_u8_l1(a + i0) + ld_u8_l1(a + i1) + ld_u8_l1(a + i2) + ld_u8_l1(a + i3);
        else acc += tex1Dfetch<unsigned char>(tex, (int)i0) + tex1Dfetch<unsigned char>(tex, (int)i1) + tex1Dfetch<unsigned char>(tex, (int)i2) + tex1Dfetch<unsigned char>(tex, (int)i3);
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// the index arithmetic alone (what the gathers cost on top of it)
__global__ void __launch_bounds__(256) index_only_kernel(uint64_t window, uint32_t per, uint32_t share, uint32_t unit, uint32_t spread, uint32_t* out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (uint32_t g = 0; g < per; ++g) acc += gather_index(t, g, window, share, unit, spread);
    if (acc == 0x12345678u) out[0] = (uint32_t)acc;
}

// shared-memory byte gathers from a 64 KB table
__global__ void __launch_bounds__(256) smem_gather_kernel(const uint8_t* __restrict__ a, uint32_t per, uint32_t* out) {
    extern __shared__ uint8_t tab[];
    for (uint32_t i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tab)[i] = reinterpret_cast<const uint32_t*>(a)[i];
    __syncthreads();
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t g = 0; g < per; g += 4) {
        const uint64_t h = mix(t * 977 + g);
        acc += tab[h & 65535] + tab[(h >> 16) & 65535] + tab[(h >> 32) & 65535] + tab[(h >> 48) & 65535];
    }
    if (acc == 0x12345678u) out[0] = acc;
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    const uint64_t maxwin = 800ull * 1000 * 1000;
    uint8_t* a; uint32_t* out;
    CK(cudaMalloc(&a, maxwin)); CK(cudaMalloc(&out, 64));
    CK(cudaMemset(a, 1, maxwin));
    {
        int maxp = 0; cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, 0);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp);
    }
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const uint64_t threads = 148ull * 2048 * 8;
    const uint32_t per = 64;
    auto timeit = [&](auto&& launch) {
        float best = 1e9f, ms;
        for (int r = 0; r < 3; ++r) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
        CK(cudaGetLastError());
        return best;
    };
    const float t_idx = timeit([&] { index_only_kernel<<<(unsigned)(threads / 256), 256>>>(50000000ull, per, 1, 32, 0, out); });
    printf("index arithmetic alone: %.3f ms for %.1f G indices\n", t_idx, threads * per / 1e9);
    auto report = [&](const char* what, float ms) { printf("%-72s %8.3f ms  %7.1f G gathers/s\n", what, ms, (double)threads * per / ms / 1e6); };

    printf("== (1) fully divergent byte gathers against window size ==\n");
    for (uint64_t win : {32768ull, 1000000ull, 8000000ull, 50000000ull, 100000000ull, 800000000ull}) {
        cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = a; rd.res.linear.desc = cudaCreateChannelDesc<unsigned char>();
        rd.res.linear.sizeInBytes = win < (1ull << 27) ? win : (1ull << 27);
        cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tex = 0;
        const bool has_tex = cudaCreateTextureObject(&tex, &rd, &td, nullptr) == cudaSuccess && win <= (1ull << 27);
        cudaGetLastError();
        char nm[128];
        snprintf(nm, sizeof nm, "window %9llu B  LDG.U8 no_allocate, L2 evict_last", (unsigned long long)win);
        report(nm, timeit([&] { gather_kernel<0><<<(unsigned)(threads / 256), 256>>>(a, 0, win, per, 1, 32, 0, out); }));
        snprintf(nm, sizeof nm, "window %9llu B  LDG.U8 through L1 (__ldg)", (unsigned long long)win);
        report(nm, timeit([&] { gather_kernel<1><<<(unsigned)(threads / 256), 256>>>(a, 0, win, per, 1, 32, 0, out); }));
        if (has_tex) {
            snprintf(nm, sizeof nm, "window %9llu B  tex1Dfetch<unsigned char>", (unsigned long long)win);
            report(nm, timeit([&] { gather_kernel<2><<<(unsigned)(threads / 256), 256>>>(a, tex, win, per, 1, 32, 0, out); }));
        }
        if (tex) cudaDestroyTextureObject(tex);
    }
    printf("== (2) lanes sharing a sector / a line (50 MB window, LDG.U8 no_allocate) ==\n");
    for (uint32_t share : {1u, 2u, 4u, 8u, 32u}) {
        char nm[128];
        snprintf(nm, sizeof nm, "%2u lanes per 32 B sector", share);
        report(nm, timeit([&] { gather_kernel<0><<<(unsigned)(threads / 256), 256>>>(a, 0, 50000000ull, per, share, 32, 0, out); }));
        if (share > 1 && share <= 4) {
            snprintf(nm, sizeof nm, "%2u lanes per 128 B line, each on its own sector", share);
            report(nm, timeit([&] { gather_kernel<0><<<(unsigned)(threads / 256), 256>>>(a, 0, 50000000ull, per, share, 128, 1, out); }));
        }
    }
    printf("== (4) shared-memory byte gathers from a 64 KB table ==\n");
    CK(cudaFuncSetAttribute(smem_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    report("LDS.U8, random bank", timeit([&] { smem_gather_kernel<<<(unsigned)(threads / 256), 256, 65536>>>(a, per, out); }));
    return 0;
}
