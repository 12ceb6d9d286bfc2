// prefilter.cu — microbenchmark: quantised-key prefilter for selective reduce transitions (Hegselmann-Krause).
// Question: the HK fold only accepts a neighbour when |o_s - o_t| < eps (4 % of the neighbours at eps = 0.02). If the
// engine keeps a 1-byte monotone quantisation of the gathered field next to the state column (1e8 agents = 100 MB,
// L2-sized), a single direct pass can gather the KEY of every source from L2 and fetch the exact 8 B state from DRAM only
// for the few edges whose keys are within the band.  The result stays bit-identical to the unfiltered fold: the decision
// is always taken on the exact value, the key only rules out edges that cannot pass.
//   usage: prefilter [n_agents] [degree] [persist_mb]     env POWERLAW=1 -> Pareto degrees as in blocked.cu
// Prints: direct (unfiltered) step, key build, prefilter with L lanes per row, prefilter with the warp-staged
// edge-parallel shape (ordered queue, sequential per-row fold), and the mismatches of each against the direct step.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
__host__ __device__ __forceinline__ uint64_t mix(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k; }
__device__ __forceinline__ double unit(uint64_t k) { return (double)(mix(k) >> 11) * (1.0 / 9007199254740992.0); }
__device__ __forceinline__ uint32_t source_of(uint64_t e, uint64_t n) {   // hub-skewed like hk_source: n * u^2
    const double u = unit(e);
    uint64_t s = (uint64_t)((double)n * u * u);
    return (uint32_t)(s >= n ? n - 1 : s);
}
__device__ __forceinline__ double ld_gather(const double* p) {
    double r; asm volatile("ld.global.nc.L1::no_allocate.L2::64B.f64 %0, [%1];" : "=d"(r) : "l"(p)); return r;
}
template <int MODE> __device__ __forceinline__ uint32_t ld_key(const uint8_t* p, uint64_t pol) {
    uint32_t r;
    if (MODE == 0) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    else if (MODE == 1) asm volatile("ld.global.nc.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));   // L1 allocate: hub keys hit L1
    else if (MODE == 4) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::64B.u8 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    else r = __ldg(p);
    return r;
}
__device__ __forceinline__ uint64_t keep_policy() { uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol)); return pol; }
__host__ __device__ __forceinline__ uint32_t quant(double o) { const double q = o * 256.0; return q >= 255.0 ? 255u : (q <= 0.0 ? 0u : (uint32_t)q); }

__global__ void fill_degrees(uint32_t* d, uint64_t n, uint32_t deg, int powerlaw) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n) return;
    if (t == n) { d[t] = 0; return; }
    if (!powerlaw) { d[t] = deg + 1; return; }
    const double u = unit(t ^ 0x9e3779b97f4a7c15ull);
    const double x = 6.8333 * pow(1.0 - u, -2.0 / 3.0);
    d[t] = (x >= 16383.0 ? 16383u : (uint32_t)x) + 1;
}
__global__ void fill_direct(uint32_t* src, const uint32_t* __restrict__ roff, uint64_t n) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t b = roff[t], e = roff[t + 1];
    for (uint32_t k = b; k + 1 < e; ++k) src[k] = source_of(k, n);
    src[e - 1] = (uint32_t)t;
}
__global__ void fill_state(double* s, uint64_t n) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) s[t] = unit(t + 0x51ed270b1ull);
}
__global__ void build_keys(const double* __restrict__ s, uint8_t* __restrict__ key, uint64_t n) {     // 4 agents per thread
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (t + 3 < n) {
        const double2 a = *reinterpret_cast<const double2*>(s + t), b = *reinterpret_cast<const double2*>(s + t + 2);
        *reinterpret_cast<uint32_t*>(key + t) = quant(a.x) | (quant(a.y) << 8) | (quant(b.x) << 16) | (quant(b.y) << 24);
    } else for (uint64_t i = t; i < n; ++i) key[i] = (uint8_t)quant(s[i]);
}

// (0) direct: 8 lanes per row, exact gathers for every edge
__global__ void direct_step(const uint32_t* __restrict__ src, const uint32_t* __restrict__ roff, const double* __restrict__ state,
                            double* __restrict__ out, uint32_t* __restrict__ outc, uint64_t n, double eps) {
    const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const uint32_t lane = threadIdx.x & 7;
    if (g >= n) return;
    const double own = state[g];
    double s = 0; uint32_t c = 0;
    const uint32_t rb = roff[g], re = roff[g + 1];
    for (uint32_t k = rb + lane; k < re; k += 8) {
        const double v = ld_gather(state + __ldcs(src + k));
        if (fabs(v - own) < eps) { s += v; c += 1; }
    }
    for (int o = 4; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if (lane == 0) { out[g] = s / (double)c; outc[g] = c; }
}

// (1) prefilter, L lanes per row, two edges per lane and round, exact gather inline
template <int L, int MODE>
__global__ void __launch_bounds__(256) prefilter_lanes(const uint32_t* __restrict__ src, const uint32_t* __restrict__ roff, const double* __restrict__ state,
                                                       const uint8_t* __restrict__ key, double* __restrict__ out, uint32_t* __restrict__ outc, uint64_t n, double eps, int band) {
    const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
    const uint32_t lane = threadIdx.x % L;
    if (g >= n) return;
    const uint64_t pol = keep_policy();
    const double own = state[g];
    const int ok = (int)quant(own);
    double s = 0; uint32_t c = 0;
    const uint32_t rb = roff[g], re = roff[g + 1];
    for (uint32_t k = rb + lane; k < re; k += 2 * L) {
        const bool two = k + L < re;
        const uint32_t i0 = __ldcs(src + k), i1 = two ? __ldcs(src + k + L) : 0u;
        const int k0 = (int)ld_key<MODE>(key + i0, pol), k1 = two ? (int)ld_key<MODE>(key + i1, pol) : 100000;
        if (abs(k0 - ok) <= band) { const double v = ld_gather(state + i0); if (fabs(v - own) < eps) { s += v; c += 1; } }
        if (abs(k1 - ok) <= band) { const double v = ld_gather(state + i1); if (fabs(v - own) < eps) { s += v; c += 1; } }
    }
    for (int o = L / 2; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if (lane == 0) { out[g] = s / (double)c; outc[g] = c; }
}

// (2) prefilter, warp-staged: a warp owns 32 consecutive rows and walks their (contiguous) entries edge-parallel in chunks
// of 128; edges whose key is within the band are queued in entry order (ballot compaction) with their row tag; a flush
// gathers the exact states of the queue edge-parallel and every lane folds its own row's segment in entry order, i.e. in
// exactly the order of the sequential reference fold.
constexpr int WPB = 8;
template <int QCAP> struct WarpStage { uint32_t off[33]; uint32_t qcnt[32]; uint32_t key[32]; uint32_t nz[32]; uint32_t qidx[QCAP]; double qval[QCAP]; };
// RF (row find) = 0: the row that owns an entry by a 5-step search of the 33 offsets in shared memory (5 dependent LDS per entry);
// RF = 1: per 32-entry sub-chunk one REDUX builds the mask of positions where a non-empty row starts and one ballot counts the rows
// started before it; an entry's row is then nz[started_before + popc(mask & lanes_le) - 1] with nz = the table of non-empty rows
// (one LDS per entry).  Index arithmetic checked against brute force on the CPU (random rows incl. empty ones).
template <int MODE, bool FIRST = true, bool LAST = true, int CHUNK = 128, int WPB = 8, int RF = 0>
__global__ void __launch_bounds__(32 * WPB) prefilter_warp(const uint32_t* __restrict__ src, const uint32_t* __restrict__ roff, const double* __restrict__ state,
                                                           const uint8_t* __restrict__ key, double* __restrict__ out, uint32_t* __restrict__ outc, uint64_t n, double eps, int band,
                                                           double* __restrict__ sum = nullptr, uint32_t* __restrict__ cnt = nullptr, uint32_t head = 0xffffffffu,
                                                           cudaTextureObject_t ktex = 0) {
    constexpr int QCAP = CHUNK + 32, U = CHUNK / 32;
    __shared__ WarpStage<QCAP> stage[WPB];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t r0 = ((uint64_t)blockIdx.x * WPB + wib) * 32;
    if (r0 >= n) return;
    WarpStage<QCAP>& sm = stage[wib];
    const uint32_t nr = (uint32_t)(n - r0 < 32 ? n - r0 : 32);
    const uint64_t pol_keep = keep_policy();
    uint64_t pol_first; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    sm.off[lane] = __ldcs(roff + r0 + (lane < nr ? lane : nr));
    if (lane == 0) sm.off[32] = __ldcs(roff + r0 + nr);
    const double own = lane < nr ? __ldcs(state + r0 + lane) : 0.0;
    sm.key[lane] = quant(own);
    sm.qcnt[lane] = 0;
    __syncwarp();
    const uint32_t e0 = sm.off[0], e1 = sm.off[32];
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t my_lo = sm.off[lane];
    const bool nonempty = sm.off[lane + 1] > my_lo;
    if (RF == 1) {
        const uint32_t nzmask = __ballot_sync(0xffffffffu, nonempty);
        if (nonempty) sm.nz[__popc(nzmask & lt)] = lane;
        __syncwarp();
    }
    double s = 0; uint32_t c = 0, qn = 0;
    if (!FIRST && lane < nr) { s = __ldcs(sum + r0 + lane); c = __ldcs(cnt + r0 + lane); }
    for (uint32_t base = e0; base < e1; base += CHUNK) {
        uint32_t idx[U]; int ks[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const uint32_t e = base + u * 32 + lane; idx[u] = e < e1 ? __ldcs(src + e) : 0xffffffffu; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (MODE == 5) ks[u] = idx[u] != 0xffffffffu ? (int)tex1Dfetch<unsigned char>(ktex, (int)idx[u]) : 100000;     // texture path (NOT RUN YET)
            else ks[u] = idx[u] != 0xffffffffu ? (int)ld_key<(MODE == 3 ? 0 : MODE)>(key + idx[u], MODE == 3 && idx[u] >= head ? pol_first : pol_keep) : 100000;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t e = base + u * 32 + lane;
            uint32_t r = 0;
            if (RF == 0) {
#pragma unroll
                for (int st = 16; st; st >>= 1) if (sm.off[r + st] <= e) r += st;
            } else {
                const uint32_t x0 = base + u * 32, d = my_lo - x0;
                const uint32_t before = __popc(__ballot_sync(0xffffffffu, nonempty && my_lo < x0));
                const uint32_t starts = __reduce_or_sync(0xffffffffu, (nonempty && my_lo >= x0 && d < 32u) ? (1u << d) : 0u);
                r = sm.nz[(before + __popc(starts & (lt | (1u << lane))) - 1u) & 31u];
            }
            const bool pass = abs(ks[u] - (int)sm.key[r]) <= band;
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            if (pass) { sm.qidx[qn + __popc(m & lt)] = idx[u] | (r << 27); atomicAdd(&sm.qcnt[r], 1u); }
            qn += __popc(m);
        }
        __syncwarp();
        if (qn > QCAP - CHUNK || base + CHUNK >= e1) {
            for (uint32_t i = lane; i < qn; i += 32) sm.qval[i] = ld_gather(state + (sm.qidx[i] & 0x7ffffffu));
            __syncwarp();
            const uint32_t mine = sm.qcnt[lane];
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += t; }
            const uint32_t start = incl - mine;
            for (uint32_t j = 0; j < mine; ++j) { const double v = sm.qval[start + j]; if (fabs(v - own) < eps) { s += v; c += 1; } }
            sm.qcnt[lane] = 0; qn = 0;
            __syncwarp();
        }
    }
    if (lane < nr) {
        if (LAST) { __stcs(out + r0 + lane, s / (double)c); __stcs(outc + r0 + lane, c); }
        else { __stcs(sum + r0 + lane, s); __stcs(cnt + r0 + lane, c); }
    }
}

// source-blocked layout for the key-blocked variant (as in blocked.cu): per block b the offsets off[b][n+4] and the entries
__global__ void count_blocks(const uint32_t* __restrict__ src, const uint32_t* __restrict__ roff, uint64_t n, uint32_t bsize, uint32_t* cnt) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    for (uint32_t k = roff[t]; k < roff[t + 1]; ++k) { const uint32_t b = src[k] / bsize; cnt[(uint64_t)b * (n + 4) + t] += 1; }
}
__global__ void fill_blocks(const uint32_t* __restrict__ src, const uint32_t* __restrict__ roff, uint64_t n, uint32_t bsize, uint32_t nb, const uint32_t* off, const uint64_t* base, uint32_t* bsrc) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint32_t fill[8];
    for (uint32_t b = 0; b < nb; ++b) fill[b] = 0;
    for (uint32_t k = roff[t]; k < roff[t + 1]; ++k) {
        const uint32_t s = src[k], b = s / bsize;
        bsrc[base[b] + off[(uint64_t)b * (n + 4) + t] + fill[b]++] = s;
    }
}

__global__ void compare(const double* __restrict__ a, const uint32_t* __restrict__ ac, const double* __restrict__ b, const uint32_t* __restrict__ bc, uint64_t n, unsigned long long* bad) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (ac[t] != bc[t]) atomicAdd(bad, 1ull);
    else if (fabs(a[t] - b[t]) > 1e-12 * fabs(a[t])) atomicAdd(bad + 1, 1ull);
    if (a[t] != b[t]) atomicAdd(bad + 2, 1ull);
}

struct Bench {
    cudaEvent_t e0, e1; uint64_t n, E; const double* ref; const uint32_t* refc; double* out; uint32_t* outc; unsigned long long* bad;
    template <class F> void run(const char* name, F&& launch) {
        float best = 1e9f, ms;
        CK(cudaMemset(out, 0, n * 8)); CK(cudaMemset(outc, 0, n * 4));
        for (int r = 0; r < 4; ++r) {
            CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        CK(cudaGetLastError());
        unsigned long long h[3] = {0, 0, 0};
        if (ref) {
            CK(cudaMemset(bad, 0, 24));
            compare<<<(unsigned)((n + 255) / 256), 256>>>(ref, refc, out, outc, n, bad);
            CK(cudaMemcpy(h, bad, 24, cudaMemcpyDeviceToHost));
        }
        printf("%-44s %8.3f ms  %7.2f Gedges/s   count mismatches %llu, value > 1e-12 rel %llu, not bit-equal %llu\n", name, best, (double)E / best / 1e6, h[0], h[1], h[2]);
    }
};

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IONBF, 0);
    const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 100000000ull;
    const uint32_t deg = argc > 2 ? atoi(argv[2]) : 20;
    const int powerlaw = getenv("POWERLAW") != nullptr;
    if (n >= (1ull << 27)) { printf("n must stay below 2^27 (row tag packing of the queue)\n"); return 1; }
    const double eps = 0.02;
    const int band = (int)std::floor(eps * 256.0) + 1;      // |o_s - o_t| < eps  =>  |q(o_s) - q(o_t)| <= floor(256 eps) + 1
    {
        int maxp = 0, l2 = 0;
        cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, 0); cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, 0);
        size_t want = argc > 3 ? (size_t)atoll(argv[3]) * 1000000 : (size_t)maxp;
        cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want); size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
        printf("L2 %d B, max persisting %d B, set-aside want %zu -> %s, now %zu\n", l2, maxp, want, cudaGetErrorString(e), got);
    }
    uint32_t* roff; CK(cudaMalloc(&roff, (n + 1) * 4));
    fill_degrees<<<(unsigned)((n + 256) / 256), 256>>>(roff, n, deg, powerlaw);
    { void* t = nullptr; size_t tz = 0; cub::DeviceScan::ExclusiveSum(t, tz, roff, roff, (int)(n + 1)); CK(cudaMalloc(&t, tz)); cub::DeviceScan::ExclusiveSum(t, tz, roff, roff, (int)(n + 1)); CK(cudaFree(t)); }
    uint32_t E32; CK(cudaMemcpy(&E32, roff + n, 4, cudaMemcpyDeviceToHost));
    const uint64_t E = E32;
    double *state, *ref, *out; uint32_t *refc, *outc, *src; uint8_t* key; unsigned long long* bad;
    CK(cudaMalloc(&state, n * 8)); CK(cudaMalloc(&ref, n * 8)); CK(cudaMalloc(&out, n * 8)); CK(cudaMalloc(&refc, n * 4)); CK(cudaMalloc(&outc, n * 4));
    CK(cudaMalloc(&src, E * 4)); CK(cudaMalloc(&key, n + 64)); CK(cudaMalloc(&bad, 24));
    fill_state<<<(unsigned)((n + 255) / 256), 256>>>(state, n);
    fill_direct<<<(unsigned)((n + 255) / 256), 256>>>(src, roff, n);
    CK(cudaDeviceSynchronize());
    printf("n=%llu E=%llu powerlaw=%d eps=%.3f band=%d (key pass rate ~%.1f %%)\n", (unsigned long long)n, (unsigned long long)E, powerlaw, eps, band, 100.0 * (2 * band + 1) / 256.0);
    Bench b; CK(cudaEventCreate(&b.e0)); CK(cudaEventCreate(&b.e1)); b.n = n; b.E = E; b.ref = nullptr; b.refc = nullptr; b.out = ref; b.outc = refc; b.bad = bad;
    b.run("direct, 8 lanes per row (no prefilter)", [&] { direct_step<<<(unsigned)((n * 8 + 255) / 256), 256>>>(src, roff, state, ref, refc, n, eps); });
    b.ref = ref; b.refc = refc; b.out = out; b.outc = outc;
    {
        float ms; CK(cudaEventRecord(b.e0)); build_keys<<<(unsigned)((n / 4 + 256) / 256), 256>>>(state, key, n); CK(cudaEventRecord(b.e1)); CK(cudaEventSynchronize(b.e1));
        CK(cudaEventElapsedTime(&ms, b.e0, b.e1)); printf("build_keys: %.3f ms (%.1f MB of keys)\n", ms, n / 1e6);
    }
    if (!getenv("SHAPES2")) {
    b.run("prefilter  4 lanes/row, keys L2 evict_last", [&] { prefilter_lanes<4, 0><<<(unsigned)((n * 4 + 255) / 256), 256>>>(src, roff, state, key, out, outc, n, eps, band); });
    b.run("prefilter  8 lanes/row, keys L2 evict_last", [&] { prefilter_lanes<8, 0><<<(unsigned)((n * 8 + 255) / 256), 256>>>(src, roff, state, key, out, outc, n, eps, band); });
    b.run("prefilter  8 lanes/row, keys L1+L2 evict_last", [&] { prefilter_lanes<8, 1><<<(unsigned)((n * 8 + 255) / 256), 256>>>(src, roff, state, key, out, outc, n, eps, band); });
    b.run("prefilter  8 lanes/row, keys plain ldg", [&] { prefilter_lanes<8, 2><<<(unsigned)((n * 8 + 255) / 256), 256>>>(src, roff, state, key, out, outc, n, eps, band); });
    b.run("prefilter 16 lanes/row, keys L2 evict_last", [&] { prefilter_lanes<16, 0><<<(unsigned)((n * 16 + 255) / 256), 256>>>(src, roff, state, key, out, outc, n, eps, band); });
    }
    const unsigned wgrid = (unsigned)((n + 32 * WPB - 1) / (32 * WPB));
    if (!getenv("SHAPES2")) {
    b.run("prefilter warp-staged, keys L2 evict_last", [&] { prefilter_warp<0><<<wgrid, 32 * WPB>>>(src, roff, state, key, out, outc, n, eps, band); });
    b.run("prefilter warp-staged, keys L1+L2 evict_last", [&] { prefilter_warp<1><<<wgrid, 32 * WPB>>>(src, roff, state, key, out, outc, n, eps, band); });
    b.run("prefilter warp-staged, keys plain ldg", [&] { prefilter_warp<2><<<wgrid, 32 * WPB>>>(src, roff, state, key, out, outc, n, eps, band); });
    for (uint32_t head_mb : {56u, 64u, 72u, 80u})
        if (n > head_mb * 1000000ull) {
            char nm[96]; snprintf(nm, sizeof nm, "warp-staged, evict_last for the first %u MB only", head_mb);
            b.run(nm, [&] { prefilter_warp<3><<<wgrid, 32 * WPB>>>(src, roff, state, key, out, outc, n, eps, band, nullptr, nullptr, head_mb * 1000000u); });
        }
    b.run("warp-staged, keys evict_last + L2::64B", [&] { prefilter_warp<4><<<wgrid, 32 * WPB>>>(src, roff, state, key, out, outc, n, eps, band); });
    for (int co : {0, 15, 25, 35, 50, 70}) {
        CK(cudaFuncSetAttribute(prefilter_warp<0>, cudaFuncAttributePreferredSharedMemoryCarveout, co));
        char nm[96]; snprintf(nm, sizeof nm, "warp-staged, evict_last, smem carveout %d %%", co);
        int nblk = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, prefilter_warp<0>, 32 * WPB, 0);
        b.run(nm, [&] { prefilter_warp<0><<<wgrid, 32 * WPB>>>(src, roff, state, key, out, outc, n, eps, band); });
        printf("   (occupancy calculator: %d CTAs/SM)\n", nblk);
    }
    CK(cudaFuncSetAttribute(prefilter_warp<0>, cudaFuncAttributePreferredSharedMemoryCarveout, -1));
    }
    // key-blocked: nb sweeps, each over the rows' entries whose source lies in block b (keys of one block = n / nb bytes)
    double* sum; uint32_t* cnt; CK(cudaMalloc(&sum, n * 8)); CK(cudaMalloc(&cnt, n * 4));
    uint32_t* bsrc; CK(cudaMalloc(&bsrc, E * 4 + 64));
    for (uint32_t nb : {2u}) {
        const uint32_t bsize = (uint32_t)((n + nb - 1) / nb);
        uint32_t* off; uint64_t* base;
        CK(cudaMalloc(&off, (uint64_t)nb * (n + 4) * 4)); CK(cudaMalloc(&base, nb * 8));
        CK(cudaMemset(off, 0, (uint64_t)nb * (n + 4) * 4));
        count_blocks<<<(unsigned)((n + 255) / 256), 256>>>(src, roff, n, bsize, off);
        void* tmp = nullptr; size_t tmpsz = 0;
        cub::DeviceScan::ExclusiveSum(tmp, tmpsz, off, off, (int)(n + 4));
        CK(cudaMalloc(&tmp, tmpsz));
        std::vector<uint64_t> hbase(nb + 1, 0);
        for (uint32_t k = 0; k < nb; ++k) {
            cub::DeviceScan::ExclusiveSum(tmp, tmpsz, off + (uint64_t)k * (n + 4), off + (uint64_t)k * (n + 4), (int)(n + 4));
            uint32_t tot; CK(cudaMemcpy(&tot, off + (uint64_t)k * (n + 4) + n, 4, cudaMemcpyDeviceToHost));
            hbase[k + 1] = hbase[k] + tot;
        }
        CK(cudaMemcpy(base, hbase.data(), nb * 8, cudaMemcpyHostToDevice));
        fill_blocks<<<(unsigned)((n + 255) / 256), 256>>>(src, roff, n, bsize, nb, off, base, bsrc);
        CK(cudaDeviceSynchronize());
        char nm[96]; snprintf(nm, sizeof nm, "warp-staged, key-blocked nb=%u (%.0f MB of keys each)", nb, bsize / 1e6);
        b.run(nm, [&] {
            for (uint32_t k = 0; k < nb; ++k) {
                const uint32_t* o = off + (uint64_t)k * (n + 4); const uint32_t* sp = bsrc + hbase[k];
                if (k == 0) prefilter_warp<0, true, false><<<wgrid, 32 * WPB>>>(sp, o, state, key, out, outc, n, eps, band, sum, cnt);
                else if (k + 1 == nb) prefilter_warp<0, false, true><<<wgrid, 32 * WPB>>>(sp, o, state, key, out, outc, n, eps, band, sum, cnt);
                else prefilter_warp<0, false, false><<<wgrid, 32 * WPB>>>(sp, o, state, key, out, outc, n, eps, band, sum, cnt);
            }
        });
        auto shape = [&](auto first, auto last, const char* what, int wpb, int co) {
            cudaFuncSetAttribute(first, cudaFuncAttributePreferredSharedMemoryCarveout, co); cudaFuncSetAttribute(last, cudaFuncAttributePreferredSharedMemoryCarveout, co);
            char nm2[96]; snprintf(nm2, sizeof nm2, "   nb=2 %s, carveout %d", what, co);
            const unsigned g = (unsigned)((n + 32 * wpb - 1) / (32 * wpb));
            b.run(nm2, [&] {
                first<<<g, 32 * wpb>>>(bsrc + hbase[0], off, state, key, out, outc, n, eps, band, sum, cnt, 0xffffffffu, (cudaTextureObject_t)0);
                last<<<g, 32 * wpb>>>(bsrc + hbase[1], off + (n + 4), state, key, out, outc, n, eps, band, sum, cnt, 0xffffffffu, (cudaTextureObject_t)0);
            });
        };
        if (nb == 2 && !getenv("SHAPES2")) for (int co : {-1, 15, 25, 35, 50}) {
            shape(prefilter_warp<0, true, false, 128, 8>, prefilter_warp<0, false, true, 128, 8>, "chunk 128, 8 warps", 8, co);
            shape(prefilter_warp<0, true, false, 64, 8>, prefilter_warp<0, false, true, 64, 8>, "chunk  64, 8 warps", 8, co);
            shape(prefilter_warp<0, true, false, 256, 8>, prefilter_warp<0, false, true, 256, 8>, "chunk 256, 8 warps", 8, co);
            shape(prefilter_warp<0, true, false, 128, 4>, prefilter_warp<0, false, true, 128, 4>, "chunk 128, 4 warps", 4, co);
        }
        if (nb == 2 && getenv("SHAPES2")) {     // CTA size: a CTA lives as long as its slowest warp (power-law rows), so small CTAs idle less
            shape(prefilter_warp<0, true, false, 64, 8>, prefilter_warp<0, false, true, 64, 8>, "chunk  64, 8 warps", 8, -1);
            shape(prefilter_warp<0, true, false, 64, 4>, prefilter_warp<0, false, true, 64, 4>, "chunk  64, 4 warps", 4, -1);
            shape(prefilter_warp<0, true, false, 64, 2>, prefilter_warp<0, false, true, 64, 2>, "chunk  64, 2 warps", 2, -1);
            shape(prefilter_warp<0, true, false, 128, 4>, prefilter_warp<0, false, true, 128, 4>, "chunk 128, 4 warps", 4, -1);
            shape(prefilter_warp<0, true, false, 128, 2>, prefilter_warp<0, false, true, 128, 2>, "chunk 128, 2 warps", 2, -1);
            shape(prefilter_warp<0, true, false, 128, 1>, prefilter_warp<0, false, true, 128, 1>, "chunk 128, 1 warp ", 1, -1);
            shape(prefilter_warp<0, true, false, 96, 4>, prefilter_warp<0, false, true, 96, 4>, "chunk  96, 4 warps", 4, -1);
            shape(prefilter_warp<0, true, false, 96, 2>, prefilter_warp<0, false, true, 96, 2>, "chunk  96, 2 warps", 2, -1);
            {   // NOT RUN YET: keys through the texture path (linear texture over the key column; 2^27 texels at most)
                cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = key; rd.res.linear.desc = cudaCreateChannelDesc<unsigned char>();
                rd.res.linear.sizeInBytes = n;
                cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
                cudaTextureObject_t ktex = 0;
                if (cudaCreateTextureObject(&ktex, &rd, &td, nullptr) == cudaSuccess) {
                    const unsigned g = (unsigned)((n + 63) / 64);
                    b.run("   nb=2 chunk  64, 2 warps, keys via tex1Dfetch", [&] {
                        prefilter_warp<5, true, false, 64, 2><<<g, 64>>>(bsrc + hbase[0], off, state, key, out, outc, n, eps, band, sum, cnt, 0xffffffffu, ktex);
                        prefilter_warp<5, false, true, 64, 2><<<g, 64>>>(bsrc + hbase[1], off + (n + 4), state, key, out, outc, n, eps, band, sum, cnt, 0xffffffffu, ktex);
                    });
                    cudaDestroyTextureObject(ktex);
                } else { printf("   texture object over %llu keys refused: %s\n", (unsigned long long)n, cudaGetErrorString(cudaGetLastError())); }
            }
            // NOT RUN YET (added after the round's GPU budget was spent): mask / popc row find instead of the offset search
            shape(prefilter_warp<0, true, false, 64, 2, 1>, prefilter_warp<0, false, true, 64, 2, 1>, "chunk  64, 2 warps, mask row find", 2, -1);
            shape(prefilter_warp<0, true, false, 64, 4, 1>, prefilter_warp<0, false, true, 64, 4, 1>, "chunk  64, 4 warps, mask row find", 4, -1);
            shape(prefilter_warp<0, true, false, 96, 2, 1>, prefilter_warp<0, false, true, 96, 2, 1>, "chunk  96, 2 warps, mask row find", 2, -1);
            shape(prefilter_warp<0, true, false, 128, 4, 1>, prefilter_warp<0, false, true, 128, 4, 1>, "chunk 128, 4 warps, mask row find", 4, -1);
        }
        {   // per-sweep times
            std::vector<cudaEvent_t> ev(nb + 1); for (auto& e : ev) CK(cudaEventCreate(&e));
            for (uint32_t k = 0; k < nb; ++k) {
                const uint32_t* o = off + (uint64_t)k * (n + 4); const uint32_t* sp = bsrc + hbase[k];
                CK(cudaEventRecord(ev[k]));
                if (k == 0) prefilter_warp<0, true, false><<<wgrid, 32 * WPB>>>(sp, o, state, key, out, outc, n, eps, band, sum, cnt);
                else if (k + 1 == nb) prefilter_warp<0, false, true><<<wgrid, 32 * WPB>>>(sp, o, state, key, out, outc, n, eps, band, sum, cnt);
                else prefilter_warp<0, false, false><<<wgrid, 32 * WPB>>>(sp, o, state, key, out, outc, n, eps, band, sum, cnt);
            }
            CK(cudaEventRecord(ev[nb])); CK(cudaEventSynchronize(ev[nb]));
            printf("   per sweep [ms | M edges]:");
            for (uint32_t k = 0; k < nb; ++k) { float ms; CK(cudaEventElapsedTime(&ms, ev[k], ev[k + 1])); printf(" %.2f|%.0f", ms, (hbase[k + 1] - hbase[k]) / 1e6); }
            printf("\n");
        }
        CK(cudaFree(off)); CK(cudaFree(base)); CK(cudaFree(tmp));
    }
    return 0;
}
