// peer_gather.cu — microbenchmark (2 GPUs, one process): what does an 8-byte gather from a PEER's HBM over NVLink cost?
// Question behind it: with the prefilter a rank needs the one-byte key of every ghost but the exact 8-byte state of
// only the ~5 % that pass.  Can the flush of the sweep read those states straight from the owner's state column (P2P loads) instead
// of receiving all 8-byte states in a halo phase of its own?
//   (1) random 8 B loads from the peer against loads in flight (grid size x unroll), window 100 MB and 800 MB
//   (2) the same loads from local HBM (DRAM-random rate) for comparison
//   (3) peer stores: contiguous push of 8 B words (what halo_push_kernel does) and random 8 B stores
//   (4) mixed: local L2-resident byte gathers with a 1/16 share of peer 8 B loads in the same warps (the sweep's instruction mix)
//   usage: peer_gather        (needs two GPUs with peer access)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
__host__ __device__ __forceinline__ uint64_t mix(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k; }

__device__ __forceinline__ uint64_t ld_nc64(const uint64_t* p) { uint64_t r; asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(r) : "l"(p)); return r; }
__device__ __forceinline__ uint32_t ld_u8_keep(const uint8_t* p, uint64_t pol) {
    uint32_t r; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol)); return r;
}

template <int U>
__global__ void __launch_bounds__(256) gather64_kernel(const uint64_t* __restrict__ a, uint64_t words, uint32_t per, uint64_t* out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (uint32_t g = 0; g < per; g += U) {
        uint64_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld_nc64(a + __umul64hi(mix(t * 977 + g + u), words));
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u];
    }
    if (acc == 0x12345678u) out[0] = acc;
}
__global__ void __launch_bounds__(256) push_kernel(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void __launch_bounds__(256) push_u8_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void __launch_bounds__(256) scatter64_kernel(uint64_t* __restrict__ dst, uint64_t words, uint32_t per) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t g = 0; g < per; ++g) dst[__umul64hi(mix(t * 977 + g), words)] = t;
}
// the sweep's mix: every lane gathers `per` key bytes from a local L2-resident column; one in `share` of them is followed by an
// 8-byte load of the exact state from `st` (local or peer)
__global__ void __launch_bounds__(64, 32) mixed_kernel(const uint8_t* __restrict__ keys, uint64_t nkeys, const uint64_t* __restrict__ st, uint64_t words,
                                                        uint32_t per, uint32_t share, uint64_t* out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    uint64_t acc = 0;
    for (uint32_t g = 0; g < per; g += 2) {
        const uint64_t h0 = mix(t * 977 + g), h1 = mix(t * 977 + g + 1);
        const uint32_t k0 = ld_u8_keep(keys + __umul64hi(h0, nkeys), pol), k1 = ld_u8_keep(keys + __umul64hi(h1, nkeys), pol);
        acc += k0 + k1;
        if (share && (h0 >> 7) % share == 0) acc += ld_nc64(st + __umul64hi(h0 * 31, words));
        if (share && (h1 >> 7) % share == 0) acc += ld_nc64(st + __umul64hi(h1 * 31, words));
    }
    if (acc == 0x12345678u) out[0] = acc;
}

static float timeit(cudaStream_t s, void (*launch)(void*), void* arg, int reps = 3) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(arg); CK(cudaStreamSynchronize(s));
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0, s)); launch(arg); CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}
struct G { const uint64_t* a; uint64_t words; uint32_t per; unsigned grid; int unroll; uint64_t* out; };
static void launch_g(void* p) {
    G& g = *(G*)p;
    if (g.unroll == 1) gather64_kernel<1><<<g.grid, 256>>>(g.a, g.words, g.per, g.out);
    else if (g.unroll == 4) gather64_kernel<4><<<g.grid, 256>>>(g.a, g.words, g.per, g.out);
    else gather64_kernel<8><<<g.grid, 256>>>(g.a, g.words, g.per, g.out);
}
struct P { const uint64_t* s; uint64_t* d; uint64_t n; unsigned grid; };
static void launch_p(void* p) { P& q = *(P*)p; push_kernel<<<q.grid, 256>>>(q.s, q.d, q.n); }
struct S { uint64_t* d; uint64_t words; uint32_t per; unsigned grid; };
static void launch_s(void* p) { S& q = *(S*)p; scatter64_kernel<<<q.grid, 256>>>(q.d, q.words, q.per); }
struct M { const uint8_t* k; uint64_t nk; const uint64_t* st; uint64_t words; uint32_t per, share; unsigned grid; uint64_t* out; };
static void launch_m(void* p) { M& q = *(M*)p; mixed_kernel<<<q.grid, 64>>>(q.k, q.nk, q.st, q.words, q.per, q.share, q.out); }

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    int nd = 0; CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
    int can = 0; CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    printf("peer access 0 -> 1: %d\n", can);
    if (!can) return 0;
    const uint64_t bytes = 800ull * 1000 * 1000, words = bytes / 8;
    uint64_t *local, *peer, *out; uint8_t* keys;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&peer, bytes)); CK(cudaMemset(peer, 1, bytes)); CK(cudaDeviceSynchronize());
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
    CK(cudaMalloc(&local, bytes)); CK(cudaMemset(local, 1, bytes)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&keys, 50000000)); CK(cudaMemset(keys, 3, 50000000));
    {
        int maxp = 0; cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, 0);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp);
    }
    printf("== (1)/(2) random 8 B loads: peer HBM over NVLink vs local HBM ==\n");
    for (int where = 0; where < 2; ++where)
        for (uint64_t win : {(uint64_t)(100000000ull / 8), (uint64_t)words})
            for (unsigned grid : {148u * 2, 148u * 8, 148u * 32})
                for (int unroll : {1, 4, 8}) {
                    const uint32_t per = (uint32_t)(where == 0 ? (64u << 20) : (256u << 20)) / (grid * 256u) / 8 * 8;
                    G g{where == 0 ? peer : local, win, per < 8 ? 8 : per, grid, unroll, out};
                    const float ms = timeit(0, launch_g, &g);
                    const double n = (double)g.per * grid * 256;
                    printf("%-5s window %4llu MB  grid %5u x 256, unroll %d   %8.3f ms  %7.2f G loads/s  %7.1f GB/s of 8 B  (%6.1f GB/s of 32 B sectors)\n", where == 0 ? "peer" : "local",
                           (unsigned long long)(win * 8 / 1000000), grid, unroll, ms, n / ms * 1e-6, n * 8 / ms * 1e-6, n * 32 / ms * 1e-6);
                }
    printf("== (3) stores into the peer ==\n");
    for (unsigned grid : {148u * 4, 148u * 16}) {
        P p{local, peer, words, grid};
        const float ms = timeit(0, launch_p, &p);
        printf("contiguous 8 B stores, 800 MB, grid %5u   %8.3f ms  %7.1f GB/s\n", grid, ms, bytes / ms * 1e-6);
    }
    {
        P p{local, local + words / 2, words / 2, 148u * 16};
        const float ms = timeit(0, launch_p, &p);
        printf("(local copy of 400 MB for comparison      %8.3f ms  %7.1f GB/s read+write)\n", ms, bytes / ms * 1e-6);
    }
    for (unsigned grid : {148u * 8}) {
        S s{peer, words, 64, grid};
        const float ms = timeit(0, launch_s, &s);
        const double n = 64.0 * grid * 256;
        printf("random 8 B stores into the peer, grid %5u  %8.3f ms  %7.2f G stores/s\n", grid, ms, n / ms * 1e-6);
    }
    printf("== (4) the sweep's mix: L2-resident key gathers (50 MB column) + a share of exact 8 B loads ==\n");
    for (uint32_t share : {0u, 16u, 8u, 2u})
        for (int where = 0; where < 2; ++where) {
            if (share == 0 && where == 1) continue;
            M m{keys, 50000000ull, where == 0 ? local : peer, words, 256, share, 148u * 32 * 8, out};
            const float ms = timeit(0, launch_m, &m);
            const double n = 256.0 * m.grid * 64;
            printf("key gathers + 1/%-2u exact loads from %-5s   %8.3f ms  %7.2f G key gathers/s  %6.2f G exact loads/s\n", share, share == 0 ? "-" : where == 0 ? "local" : "peer", ms,
                   n / ms * 1e-6, share ? n / share / ms * 1e-6 : 0.0);
        }
    return 0;
}
