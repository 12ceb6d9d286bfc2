// blocked.cu — microbenchmark for the source-blocked read phase.
// Question: the HK-100M read phase gathers 2.1e9 random 8 B states from an 800 MB array and is bound by the DRAM
// random-access rate (gather.cu). If the CSR is split by SOURCE BLOCK (blocks of the state array that fit in L2) and
// the targets are swept once per block with (sum, count) accumulators streamed through HBM, how fast is a step?
//   usage: blocked [n_agents] [degree]
// Prints (1) random-gather rate against window size (effective L2 capacity), (2) the direct kernel, (3) the blocked
// sweep for several block counts.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
__host__ __device__ __forceinline__ uint64_t mix(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k; }
__device__ __forceinline__ uint32_t source_of(uint64_t e, uint64_t n) {   // hub-skewed like hk_source: n * u^2
    const double u = (double)(mix(e) >> 11) * (1.0 / 9007199254740992.0);
    uint64_t s = (uint64_t)((double)n * u * u);
    return (uint32_t)(s >= n ? n - 1 : s);
}
__device__ __forceinline__ double ld_keep(const double* p) {
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    double r; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol)); return r;
}
__device__ __forceinline__ double ld_gather(const double* p) {
    double r; asm volatile("ld.global.nc.L1::no_allocate.L2::64B.f64 %0, [%1];" : "=d"(r) : "l"(p)); return r;
}

// (1) gather rate vs window
__global__ void window_gather(const double* __restrict__ a, uint64_t window, uint64_t per, double* out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0;
    for (uint64_t i = 0; i < per; i += 4) {
        const uint64_t i0 = __umul64hi(mix(t * per + i), window), i1 = __umul64hi(mix(t * per + i + 1), window), i2 = __umul64hi(mix(t * per + i + 2), window), i3 = __umul64hi(mix(t * per + i + 3), window);
        acc += ld_keep(a + i0) + ld_keep(a + i1) + ld_keep(a + i2) + ld_keep(a + i3);
    }
    if (acc == 12345.678) out[0] = acc;
}

// graph: row t has deg(t) random sources + itself; deg = `deg` (uniform) or a Pareto(1.5) draw with mean ~20 capped at 16383 (powerlaw)
__global__ void fill_degrees(uint32_t* d, uint64_t n, uint32_t deg, int powerlaw) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n) return;
    if (t == n) { d[t] = 0; return; }
    if (!powerlaw) { d[t] = deg + 1; return; }
    const double u = (double)(mix(t ^ 0x9e3779b97f4a7c15ull) >> 11) * (1.0 / 9007199254740992.0);
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
__global__ void count_blocks(const uint32_t* __restrict__ src, const uint32_t* __restrict__ roff, uint64_t n, uint32_t bsize, uint32_t nb, uint32_t* cnt /*[nb][n+4]*/) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    for (uint32_t k = roff[t]; k < roff[t + 1]; ++k) { const uint32_t b = src[k] / bsize; cnt[(uint64_t)b * (n + 4) + t] += 1; }
}
__global__ void fill_blocks(const uint32_t* __restrict__ src, const uint32_t* __restrict__ roff, uint64_t n, uint32_t bsize, uint32_t nb, const uint32_t* off, const uint64_t* base, uint32_t* bsrc) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint32_t fill[64];
    for (uint32_t b = 0; b < nb; ++b) fill[b] = 0;
    for (uint32_t k = roff[t]; k < roff[t + 1]; ++k) {
        const uint32_t s = src[k], b = s / bsize;
        bsrc[base[b] + off[(uint64_t)b * (n + 4) + t] + fill[b]++] = s;
    }
}

// (2) direct: 8 lanes per target
__global__ void direct_step(const uint32_t* __restrict__ src, const uint32_t* __restrict__ roff, const double* __restrict__ state, double* __restrict__ out, uint64_t n, double eps) {
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
    if (lane == 0) out[g] = s / (double)c;
}

// (3) blocked sweep: thread per target
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(256) blocked_pass(const uint32_t* __restrict__ off, const uint32_t* __restrict__ bsrc, const double* __restrict__ state,
                                                    double* __restrict__ sum, uint32_t* __restrict__ cnt, double* __restrict__ out, uint64_t n, double eps) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t lo = __ldcs(off + t), hi = __ldcs(off + t + 1);
    const double own = __ldcs(state + t);
    double s = FIRST ? 0.0 : __ldcs(sum + t);
    uint32_t c = FIRST ? 0u : __ldcs(cnt + t);
    for (uint32_t e = lo; e < hi; ++e) {
        const double v = ld_keep(state + __ldcs(bsrc + e));
        if (fabs(v - own) < eps) { s += v; c += 1; }
    }
    if (LAST) __stcs(out + t, s / (double)c);
    else { __stcs(sum + t, s); __stcs(cnt + t, c); }
}

// (4) blocked sweep, staged: a CTA owns TPB consecutive targets; all edges of the chunk are gathered edge-parallel into
// shared memory (coalesced index loads, independent gathers), then each thread folds the segments of its targets.
template <bool FIRST, bool LAST, int TPB, int CAP>
__global__ void __launch_bounds__(256) blocked_pass2(const uint32_t* __restrict__ off, const uint32_t* __restrict__ bsrc, const double* __restrict__ state,
                                                     double* __restrict__ sum, uint32_t* __restrict__ cnt, double* __restrict__ out, uint64_t n, double eps) {
    __shared__ uint32_t soff[TPB + 1];
    __shared__ double sval[CAP];
    const uint64_t t0 = (uint64_t)blockIdx.x * TPB;
    const uint32_t nt = (uint32_t)(n - t0 < TPB ? n - t0 : TPB);
    for (uint32_t i = threadIdx.x; i <= nt; i += 256) soff[i] = __ldcs(off + t0 + i);
    __syncthreads();
    const uint32_t e0 = soff[0], e1 = soff[nt];
    double own[TPB / 256], s[TPB / 256]; uint32_t c[TPB / 256];
#pragma unroll
    for (int j = 0; j < TPB / 256; ++j) {
        const uint32_t i = threadIdx.x + j * 256;
        if (i < nt) { own[j] = __ldcs(state + t0 + i); s[j] = FIRST ? 0.0 : __ldcs(sum + t0 + i); c[j] = FIRST ? 0u : __ldcs(cnt + t0 + i); }
    }
    for (uint32_t base = e0; base < e1; base += CAP) {
        const uint32_t m = e1 - base < CAP ? e1 - base : CAP;
        if (base != e0) __syncthreads();
        for (uint32_t i = threadIdx.x; i < m; i += 256) sval[i] = ld_keep(state + __ldcs(bsrc + base + i));
        __syncthreads();
#pragma unroll
        for (int j = 0; j < TPB / 256; ++j) {
            const uint32_t i = threadIdx.x + j * 256;
            if (i < nt) {
                uint32_t lo = soff[i], hi = soff[i + 1];
                lo = lo < base ? base : lo; hi = hi > base + m ? base + m : hi;
                for (uint32_t e = lo; e < hi; ++e) { const double v = sval[e - base]; if (fabs(v - own[j]) < eps) { s[j] += v; c[j] += 1; } }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < TPB / 256; ++j) {
        const uint32_t i = threadIdx.x + j * 256;
        if (i < nt) {
            if (LAST) __stcs(out + t0 + i, s[j] / (double)c[j]);
            else { __stcs(sum + t0 + i, s[j]); __stcs(cnt + t0 + i, c[j]); }
        }
    }
}
template <int TPB, int CAP>
float run_pass2(uint32_t nb, const uint32_t* off, const uint32_t* bsrc, const std::vector<uint64_t>& hbase, const double* state, double* sum, uint32_t* cnt, double* out, uint64_t n,
                cudaEvent_t e0, cudaEvent_t e1) {
    float best = 1e9f, ms;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        for (uint32_t b = 0; b < nb; ++b) {
            const uint32_t* o = off + (uint64_t)b * (n + 4); const uint32_t* s = bsrc + hbase[b];
            const unsigned grid = (unsigned)((n + TPB - 1) / TPB);
            if (b == 0) blocked_pass2<true, false, TPB, CAP><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
            else if (b + 1 == nb) blocked_pass2<false, true, TPB, CAP><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
            else blocked_pass2<false, false, TPB, CAP><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

// (5) blocked sweep, 4 consecutive targets per thread, 128-bit streaming loads
__device__ __forceinline__ uint4 ldcs4(const uint32_t* p) { return __ldcs(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ double2 ldcs2(const double* p) { return __ldcs(reinterpret_cast<const double2*>(p)); }
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(256) blocked_pass4(const uint32_t* __restrict__ off, const uint32_t* __restrict__ bsrc, const double* __restrict__ state,
                                                     double* __restrict__ sum, uint32_t* __restrict__ cnt, double* __restrict__ out, uint64_t n, double eps) {
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (t >= n) return;   // n % 4 == 0 in this benchmark
    const uint4 o4 = ldcs4(off + t); const uint32_t o5 = __ldcs(off + t + 4);
    const double2 a = ldcs2(state + t), b = ldcs2(state + t + 2);
    double own[4] = {a.x, a.y, b.x, b.y}, s[4] = {0, 0, 0, 0}; uint32_t c[4] = {0, 0, 0, 0};
    if (!FIRST) {
        const double2 sa = ldcs2(sum + t), sb = ldcs2(sum + t + 2); const uint4 cc = ldcs4(cnt + t);
        s[0] = sa.x; s[1] = sa.y; s[2] = sb.x; s[3] = sb.y; c[0] = cc.x; c[1] = cc.y; c[2] = cc.z; c[3] = cc.w;
    }
    const uint32_t o[5] = {o4.x, o4.y, o4.z, o4.w, o5};
#pragma unroll
    for (int j = 0; j < 4; ++j)
        for (uint32_t e = o[j]; e < o[j + 1]; ++e) {
            const double v = ld_keep(state + __ldcs(bsrc + e));
            if (fabs(v - own[j]) < eps) { s[j] += v; c[j] += 1; }
        }
    if (LAST) {
        __stcs(reinterpret_cast<double2*>(out + t), make_double2(s[0] / c[0], s[1] / c[1]));
        __stcs(reinterpret_cast<double2*>(out + t + 2), make_double2(s[2] / c[2], s[3] / c[3]));
    } else {
        __stcs(reinterpret_cast<double2*>(sum + t), make_double2(s[0], s[1]));
        __stcs(reinterpret_cast<double2*>(sum + t + 2), make_double2(s[2], s[3]));
        __stcs(reinterpret_cast<uint4*>(cnt + t), make_uint4(c[0], c[1], c[2], c[3]));
    }
}
// (6) same, flat edge loop: one loop over the thread's whole edge range, target index advanced on the fly
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(256) blocked_pass4f(const uint32_t* __restrict__ off, const uint32_t* __restrict__ bsrc, const double* __restrict__ state,
                                                      double* __restrict__ sum, uint32_t* __restrict__ cnt, double* __restrict__ out, uint64_t n, double eps) {
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (t >= n) return;
    const uint4 o4 = ldcs4(off + t); const uint32_t o5 = __ldcs(off + t + 4);
    const double2 a = ldcs2(state + t), b = ldcs2(state + t + 2);
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0; uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if (!FIRST) {
        const double2 sa = ldcs2(sum + t), sb = ldcs2(sum + t + 2); const uint4 cc = ldcs4(cnt + t);
        s0 = sa.x; s1 = sa.y; s2 = sb.x; s3 = sb.y; c0 = cc.x; c1 = cc.y; c2 = cc.z; c3 = cc.w;
    }
    for (uint32_t e = o4.x; e < o5; ++e) {
        const double v = ld_keep(state + __ldcs(bsrc + e));
        const int j = (e >= o4.y) + (e >= o4.z) + (e >= o4.w);
        const double own = j == 0 ? a.x : j == 1 ? a.y : j == 2 ? b.x : b.y;
        const bool ok = fabs(v - own) < eps;
        const double add = ok ? v : 0.0;
        s0 += j == 0 ? add : 0.0; s1 += j == 1 ? add : 0.0; s2 += j == 2 ? add : 0.0; s3 += j == 3 ? add : 0.0;
        c0 += (j == 0) & ok; c1 += (j == 1) & ok; c2 += (j == 2) & ok; c3 += (j == 3) & ok;
    }
    if (LAST) {
        __stcs(reinterpret_cast<double2*>(out + t), make_double2(s0 / c0, s1 / c1));
        __stcs(reinterpret_cast<double2*>(out + t + 2), make_double2(s2 / c2, s3 / c3));
    } else {
        __stcs(reinterpret_cast<double2*>(sum + t), make_double2(s0, s1));
        __stcs(reinterpret_cast<double2*>(sum + t + 2), make_double2(s2, s3));
        __stcs(reinterpret_cast<uint4*>(cnt + t), make_uint4(c0, c1, c2, c3));
    }
}
template <int V>
float run_pass4(uint32_t nb, const uint32_t* off, const uint32_t* bsrc, const std::vector<uint64_t>& hbase, const double* state, double* sum, uint32_t* cnt, double* out, uint64_t n,
                cudaEvent_t e0, cudaEvent_t e1) {
    float best = 1e9f, ms;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        for (uint32_t b = 0; b < nb; ++b) {
            const uint32_t* o = off + (uint64_t)b * (n + 4); const uint32_t* s = bsrc + hbase[b];
            const unsigned grid = (unsigned)((n / 4 + 255) / 256);
            if (V == 0) {
                if (b == 0) blocked_pass4<true, false><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
                else if (b + 1 == nb) blocked_pass4<false, true><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
                else blocked_pass4<false, false><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
            } else {
                if (b == 0) blocked_pass4f<true, false><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
                else if (b + 1 == nb) blocked_pass4f<false, true><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
                else blocked_pass4f<false, false><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
            }
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

// (7) blocked sweep, U interleaved targets per thread (stride = grid), joint edge loop => U independent chains in flight
template <bool FIRST, bool LAST, int U>
__global__ void __launch_bounds__(256) blocked_pass_u(const uint32_t* __restrict__ off, const uint32_t* __restrict__ bsrc, const double* __restrict__ state,
                                                      double* __restrict__ sum, uint32_t* __restrict__ cnt, double* __restrict__ out, uint64_t n, double eps) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lo[U], hi[U]; double own[U], s[U]; uint32_t c[U];
#pragma unroll
    for (int k = 0; k < U; ++k) { const uint64_t t = t0 + k * stride; lo[k] = hi[k] = 0; if (t < n) { lo[k] = __ldg(off + t); hi[k] = __ldg(off + t + 1); } }
#pragma unroll
    for (int k = 0; k < U; ++k) {
        const uint64_t t = t0 + k * stride; own[k] = 0; s[k] = 0; c[k] = 0;
        if (t < n) { own[k] = __ldcs(state + t); if (!FIRST) { s[k] = __ldcs(sum + t); c[k] = __ldcs(cnt + t); } }
    }
    bool more = true;
    while (more) {
        more = false;
        uint32_t sidx[U];
#pragma unroll
        for (int k = 0; k < U; ++k) if (lo[k] < hi[k]) sidx[k] = __ldg(bsrc + lo[k]);
        double v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) if (lo[k] < hi[k]) v[k] = ld_keep(state + sidx[k]);
#pragma unroll
        for (int k = 0; k < U; ++k) if (lo[k] < hi[k]) { if (fabs(v[k] - own[k]) < eps) { s[k] += v[k]; c[k] += 1; } lo[k] += 1; more |= lo[k] < hi[k]; }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
        const uint64_t t = t0 + k * stride;
        if (t < n) { if (LAST) __stcs(out + t, s[k] / (double)c[k]); else { __stcs(sum + t, s[k]); __stcs(cnt + t, c[k]); } }
    }
}
template <int U>
float run_pass_u(uint32_t nb, const uint32_t* off, const uint32_t* bsrc, const std::vector<uint64_t>& hbase, const double* state, double* sum, uint32_t* cnt, double* out, uint64_t n,
                 cudaEvent_t e0, cudaEvent_t e1) {
    float best = 1e9f, ms;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        for (uint32_t b = 0; b < nb; ++b) {
            const uint32_t* o = off + (uint64_t)b * (n + 4); const uint32_t* s = bsrc + hbase[b];
            const unsigned grid = (unsigned)((n + 256 * U - 1) / (256 * U));
            if (b == 0) blocked_pass_u<true, false, U><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
            else if (b + 1 == nb) blocked_pass_u<false, true, U><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
            else blocked_pass_u<false, false, U><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

// (8) warp-pipelined persistent sweep: a warp owns items of 32 consecutive rows; per loop iteration it issues the offset loads of
// item i+2, the own/acc/source-index loads of item i+1 and the gathers of item i, so three dependent hops of three items overlap.
__device__ __forceinline__ double ld_keep_pol(const double* p, uint64_t pol) {
    double r; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol)); return r;
}
template <bool FIRST, bool LAST, int ABL = 0>   // ABL 1: no gathers (streams only); 2: no own/acc streams (offsets, indices, gathers only)
__global__ void __launch_bounds__(256) blocked_pass_w(const uint32_t* __restrict__ off, const uint32_t* __restrict__ bsrc, const double* __restrict__ state,
                                                      double* __restrict__ sum, uint32_t* __restrict__ cnt, double* __restrict__ out, uint64_t n, double eps) {
    constexpr int CH = 128;                                   // edges staged per item (4 per lane)
    __shared__ double sval[8][CH];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t W = gridDim.x * 8, w = blockIdx.x * 8 + wib;
    const uint32_t nitems = (uint32_t)(n / 32);
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    double* sv = sval[wib];
    // stage A registers (offsets), stage B registers (own, acc, src)
    uint32_t a_lo = 0, a_hi = 0;
    uint32_t b_lo = 0, b_hi = 0, b_src[4] = {0, 0, 0, 0}, b_c = 0; double b_own = 0, b_s = 0;
    auto loadA = [&](uint32_t item) { if (item < nitems) { a_lo = __ldcs(off + (uint64_t)item * 32 + lane); a_hi = __ldcs(off + (uint64_t)item * 32 + lane + 1); } };
    auto loadB = [&](uint32_t item) {     // consumes a_lo/a_hi of this item
        b_lo = a_lo; b_hi = a_hi;
        if (item < nitems) {
            const uint64_t t = (uint64_t)item * 32 + lane;
            if (ABL == 2) { b_own = 0.5; b_s = 0; b_c = 0; } else {
            b_own = __ldcs(state + t);
            if (!FIRST) { b_s = __ldcs(sum + t); b_c = __ldcs(cnt + t); } else { b_s = 0; b_c = 0; } }
            const uint32_t e0 = __shfl_sync(0xffffffffu, b_lo, 0), e1 = __shfl_sync(0xffffffffu, b_hi, 31);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const uint32_t e = e0 + k * 32 + lane; if (e < e1) b_src[k] = __ldcs(bsrc + e); }
        }
    };
    loadA(w); loadB(w); loadA(w + W);
    for (uint32_t item = w; item < nitems; item += W) {
        // C registers = B registers of this item
        const uint32_t lo = b_lo, hi = b_hi; const double own = b_own; double s = b_s; uint32_t c = b_c;
        uint32_t src[4] = {b_src[0], b_src[1], b_src[2], b_src[3]};
        const uint32_t e0 = __shfl_sync(0xffffffffu, lo, 0), e1 = __shfl_sync(0xffffffffu, hi, 31);
        // gathers of this item first (their addresses are ready), then the loads of the following items
        double v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const uint32_t e = e0 + k * 32 + lane; if (e < e1) v[k] = ABL == 1 ? (double)src[k] : ld_keep_pol(state + src[k], pol); }
        loadB(item + W);          // uses the offsets loaded one iteration ago
        loadA(item + 2 * W);
#pragma unroll
        for (int k = 0; k < 4; ++k) { const uint32_t e = e0 + k * 32 + lane; if (e < e1) sv[k * 32 + lane] = v[k]; }
        __syncwarp();
        const uint32_t stop = hi < e0 + CH ? hi : e0 + CH;
        for (uint32_t e = lo; e < stop; ++e) { const double x = sv[e - e0]; if (fabs(x - own) < eps) { s += x; c += 1; } }
        for (uint32_t e = (lo > e0 + CH ? lo : e0 + CH); e < hi; ++e) {      // rare: more than CH edges in 32 rows
            const double x = ld_keep_pol(state + __ldcs(bsrc + e), pol); if (fabs(x - own) < eps) { s += x; c += 1; }
        }
        __syncwarp();
        const uint64_t t = (uint64_t)item * 32 + lane;
        if (ABL == 2) { if (s == 12345.678) out[t] = s; }
        else if (LAST) __stcs(out + t, s / (double)c); else { __stcs(sum + t, s); __stcs(cnt + t, c); }
    }
}
template <int CTAS_PER_SM, int ABL = 0>
float run_pass_w(uint32_t nb, const uint32_t* off, const uint32_t* bsrc, const std::vector<uint64_t>& hbase, const double* state, double* sum, uint32_t* cnt, double* out, uint64_t n,
                 cudaEvent_t e0, cudaEvent_t e1) {
    float best = 1e9f, ms;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        for (uint32_t b = 0; b < nb; ++b) {
            const uint32_t* o = off + (uint64_t)b * (n + 4); const uint32_t* s = bsrc + hbase[b];
            const unsigned grid = 148 * CTAS_PER_SM;
            if (b == 0) blocked_pass_w<true, false, ABL><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
            else if (b + 1 == nb) blocked_pass_w<false, true, ABL><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
            else blocked_pass_w<false, false, ABL><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

// (9) TMA-staged persistent sweep: a CTA owns tiles of T rows; one elected thread streams the tile's offset / own / accumulator
// columns and its source-index range into shared memory with cp.async.bulk (mbarrier complete_tx), NST tiles deep, so tens of KB
// per SM are always in flight; the CTA gathers the staged indices edge-parallel (independent L2 hits), then folds row-parallel.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int T, int CAP> struct TileStage {
    alignas(16) uint32_t off[T + 4];
    alignas(16) double own[T];
    alignas(16) double sum[T];
    alignas(16) uint32_t cnt[T];
    alignas(16) uint32_t src[CAP + 8];
};
template <bool FIRST, bool LAST, int T, int CAP, int NST>
__global__ void __launch_bounds__(256) blocked_pass_t(const uint32_t* __restrict__ off, const uint32_t* __restrict__ bsrc, const double* __restrict__ state,
                                                      double* __restrict__ sum, uint32_t* __restrict__ cnt, double* __restrict__ out, uint64_t n, double eps, uint64_t hb) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    TileStage<T, CAP>* st = reinterpret_cast<TileStage<T, CAP>*>(smem_raw);
    double* sval = reinterpret_cast<double*>(smem_raw + sizeof(TileStage<T, CAP>) * NST);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sval + CAP);
    const uint32_t ntiles = (uint32_t)((n + T - 1) / T);     // n % T == 0 in this benchmark
    const uint32_t tid = threadIdx.x;
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    if (tid == 0) { for (int i = 0; i < NST; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    // producer state (thread 0): edge range of the next tile to issue
    uint32_t pe0 = 0, pe1 = 0;
    auto range_of = [&](uint32_t tile) { if (tile < ntiles) { pe0 = __ldg(off + (uint64_t)tile * T); pe1 = __ldg(off + (uint64_t)tile * T + T); } };
    auto issue = [&](uint32_t tile, int sidx) {     // thread 0 only; pe0/pe1 hold the tile's range
        if (tile >= ntiles) return;
        TileStage<T, CAP>& S = st[sidx];
        const uint64_t t0 = (uint64_t)tile * T;
        const uint64_t a0 = (hb + pe0) & ~3ull;
        uint64_t a1 = (hb + pe1 + 3u) & ~3ull; if (a1 - a0 > CAP + 4) a1 = a0 + CAP + 4;
        const uint32_t nsrc = (uint32_t)(a1 - a0) * 4;
        const uint32_t bytes = (T + 4) * 4 + T * 8 + (FIRST ? 0 : T * 8 + T * 4) + nsrc;
        mbar_expect_tx(&bars[sidx], bytes);
        bulk_g2s(S.off, off + t0, (T + 4) * 4, &bars[sidx]);
        bulk_g2s(S.own, state + t0, T * 8, &bars[sidx]);
        if (!FIRST) { bulk_g2s(S.sum, sum + t0, T * 8, &bars[sidx]); bulk_g2s(S.cnt, cnt + t0, T * 4, &bars[sidx]); }
        if (nsrc) bulk_g2s(S.src, bsrc + a0, nsrc, &bars[sidx]);
    };
    if (tid == 0) {
        for (int j = 0; j < NST - 1; ++j) { range_of(blockIdx.x + j * gridDim.x); issue(blockIdx.x + j * gridDim.x, j); }
        range_of(blockIdx.x + (NST - 1) * gridDim.x);
    }
    uint32_t k = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++k) {
        const int sidx = k % NST;
        if (tid == 0) {     // refill the stage consumed in the previous iteration
            issue(tile + (NST - 1) * gridDim.x, (k + NST - 1) % NST);
            range_of(tile + NST * gridDim.x);
        }
        mbar_wait(&bars[sidx], (k / NST) & 1);
        TileStage<T, CAP>& S = st[sidx];
        const uint32_t e0 = S.off[0], e1 = S.off[T];
        const uint32_t shift = (uint32_t)((hb + e0) & 3ull);
        uint32_t m = e1 - e0; { const uint32_t avail = CAP + 4 - shift; if (m > avail) m = avail; if (m > CAP) m = CAP; }
        for (uint32_t i = tid; i < m; i += 256 * 4) {
            double v[4]; uint32_t ix[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) if (i + u * 256 < m) ix[u] = S.src[shift + i + u * 256];
#pragma unroll
            for (int u = 0; u < 4; ++u) if (i + u * 256 < m) v[u] = ld_keep_pol(state + ix[u], pol);
#pragma unroll
            for (int u = 0; u < 4; ++u) if (i + u * 256 < m) sval[i + u * 256] = v[u];
        }
        __syncthreads();
        const uint64_t t0 = (uint64_t)tile * T;
#pragma unroll
        for (int j = 0; j < T / 256; ++j) {
            const uint32_t r = tid + j * 256;
            const uint32_t lo = S.off[r] - e0, hi = S.off[r + 1] - e0;
            const double own = S.own[r];
            double s = FIRST ? 0.0 : S.sum[r]; uint32_t c = FIRST ? 0u : S.cnt[r];
            const uint32_t stop = hi < m ? hi : m;
            for (uint32_t e = lo; e < stop; ++e) { const double x = sval[e]; if (fabs(x - own) < eps) { s += x; c += 1; } }
            for (uint32_t e = (lo > m ? lo : m); e < hi; ++e) { const double x = ld_keep_pol(state + __ldg(bsrc + hb + e0 + e), pol); if (fabs(x - own) < eps) { s += x; c += 1; } }
            if (LAST) __stcs(out + t0 + r, s / (double)c); else { __stcs(sum + t0 + r, s); __stcs(cnt + t0 + r, c); }
        }
        __syncthreads();
    }
}
template <int T, int CAP, int NST>
float run_pass_t(uint32_t nb, const uint32_t* off, const uint32_t* bsrc, const std::vector<uint64_t>& hbase, const double* state, double* sum, uint32_t* cnt, double* out, uint64_t n,
                 cudaEvent_t e0, cudaEvent_t e1, int ctas_per_sm) {
    float best = 1e9f, ms;
    const size_t smem = sizeof(TileStage<T, CAP>) * NST + CAP * 8 + NST * 8;
    CK(cudaFuncSetAttribute(blocked_pass_t<true, false, T, CAP, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(blocked_pass_t<false, true, T, CAP, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(blocked_pass_t<false, false, T, CAP, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        for (uint32_t b = 0; b < nb; ++b) {
            const uint32_t* o = off + (uint64_t)b * (n + 4); const uint32_t* s = bsrc; const uint64_t hb = hbase[b];
            const unsigned grid = 148 * ctas_per_sm;
            if (b == 0) blocked_pass_t<true, false, T, CAP, NST><<<grid, 256, smem>>>(o, s, state, sum, cnt, out, n, 0.02, hb);
            else if (b + 1 == nb) blocked_pass_t<false, true, T, CAP, NST><<<grid, 256, smem>>>(o, s, state, sum, cnt, out, n, 0.02, hb);
            else blocked_pass_t<false, false, T, CAP, NST><<<grid, 256, smem>>>(o, s, state, sum, cnt, out, n, 0.02, hb);
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    printf("   TMA-staged T=%d CAP=%d NST=%d smem=%zu ctas/SM=%d: %8.3f ms\n", T, CAP, NST, smem, ctas_per_sm, best);
    return best;
}

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IONBF, 0);
    const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 100000000ull;
    const uint32_t deg = argc > 2 ? atoi(argv[2]) : 20;
    const int powerlaw = getenv("POWERLAW") != nullptr;
    uint32_t* roff; CK(cudaMalloc(&roff, (n + 1) * 4));
    fill_degrees<<<(unsigned)((n + 256) / 256), 256>>>(roff, n, deg, powerlaw);
    { void* t = nullptr; size_t tz = 0; cub::DeviceScan::ExclusiveSum(t, tz, roff, roff, (int)(n + 1)); CK(cudaMalloc(&t, tz)); cub::DeviceScan::ExclusiveSum(t, tz, roff, roff, (int)(n + 1)); CK(cudaFree(t)); }
    uint32_t E32; CK(cudaMemcpy(&E32, roff + n, 4, cudaMemcpyDeviceToHost));
    const uint64_t E = E32;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    {
        int maxp = 0, l2 = 0, maxwin = 0;
        cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, 0); cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, 0);
        cudaDeviceGetAttribute(&maxwin, cudaDevAttrMaxAccessPolicyWindowSize, 0);
        printf("L2 %d B, max persisting %d B, max window %d B\n", l2, maxp, maxwin);
        if (argc > 4) { const size_t want = (size_t)atoll(argv[4]) * 1000000; cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want); size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize); printf("persisting L2 set-aside: want %zu -> %s, now %zu\n", want, cudaGetErrorString(e), got); }
    }
    const bool reset_between = argc > 5 && atoi(argv[5]);
    double *state, *out, *sum; uint32_t *cnt, *src, *bsrc;
    CK(cudaMalloc(&state, n * 8)); CK(cudaMalloc(&out, n * 8)); CK(cudaMalloc(&sum, n * 8)); CK(cudaMalloc(&cnt, n * 4));
    CK(cudaMalloc(&src, E * 4)); CK(cudaMalloc(&bsrc, E * 4 + 64));
    CK(cudaMemset(state, 0, n * 8));

    printf("== random-gather rate vs window (L2::evict_last) ==\n");
    for (uint64_t mb : {16, 32, 48, 56, 64, 72, 80, 96, 112, 128, 160, 800}) {
        if (argc > 3) break;
        const uint64_t window = mb * 1000000ull / 8 > n ? n : mb * 1000000ull / 8;
        const uint64_t threads = 148ull * 2048 * 4, per = 64;
        for (int r = 0; r < 2; ++r) {
            CK(cudaEventRecord(e0));
            window_gather<<<threads / 256, 256>>>(state, window, per, out);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        }
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("window %4llu MB: %8.3f ms  %7.1f Ggather/s\n", (unsigned long long)mb, ms, (double)threads * per / ms / 1e6);
    }

    fill_direct<<<(unsigned)((n + 255) / 256), 256>>>(src, roff, n);
    CK(cudaDeviceSynchronize());
    printf("== direct (8 lanes per target), n=%llu E=%llu ==\n", (unsigned long long)n, (unsigned long long)E);
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        direct_step<<<(unsigned)((n * 8 + 255) / 256), 256>>>(src, roff, state, out, n, 0.02);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("direct: %8.3f ms  %7.2f Gedges/s\n", ms, (double)E / ms / 1e6);

    std::vector<uint32_t> nbs = {8u, 10u, 13u, 16u};
    if (argc > 3) nbs = {(uint32_t)atoi(argv[3])};
    for (uint32_t nb : nbs) {
        const uint32_t bsize = (uint32_t)((n + nb - 1) / nb);
        uint32_t* off; uint64_t* base;
        CK(cudaMalloc(&off, (uint64_t)nb * (n + 4) * 4)); CK(cudaMalloc(&base, nb * 8));
        CK(cudaMemset(off, 0, (uint64_t)nb * (n + 4) * 4));
        count_blocks<<<(unsigned)((n + 255) / 256), 256>>>(src, roff, n, bsize, nb, off);
        void* tmp = nullptr; size_t tmpsz = 0;
        cub::DeviceScan::ExclusiveSum(tmp, tmpsz, off, off, (int)(n + 4));
        CK(cudaMalloc(&tmp, tmpsz));
        std::vector<uint64_t> hbase(nb + 1, 0);
        for (uint32_t b = 0; b < nb; ++b) {
            cub::DeviceScan::ExclusiveSum(tmp, tmpsz, off + (uint64_t)b * (n + 4), off + (uint64_t)b * (n + 4), (int)(n + 4));
            uint32_t tot; CK(cudaMemcpy(&tot, off + (uint64_t)b * (n + 4) + n, 4, cudaMemcpyDeviceToHost));
            hbase[b + 1] = hbase[b] + tot;
        }
        CK(cudaMemcpy(base, hbase.data(), nb * 8, cudaMemcpyHostToDevice));
        fill_blocks<<<(unsigned)((n + 255) / 256), 256>>>(src, roff, n, bsize, nb, off, base, bsrc);
        CK(cudaDeviceSynchronize());
        float best = 1e9f;
        for (int r = 0; r < 3; ++r) {
            CK(cudaEventRecord(e0));
            for (uint32_t b = 0; b < nb; ++b) {
                const uint32_t* o = off + (uint64_t)b * (n + 4); const uint32_t* s = bsrc + hbase[b];
                const unsigned grid = (unsigned)((n + 255) / 256);
                if (b == 0) blocked_pass<true, false><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
                else if (b + 1 == nb) blocked_pass<false, true><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
                else blocked_pass<false, false><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
            }
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        {   // per-pass times
            std::vector<cudaEvent_t> ev(nb + 1);
            for (auto& e : ev) CK(cudaEventCreate(&e));
            for (uint32_t b = 0; b < nb; ++b) {
                const uint32_t* o = off + (uint64_t)b * (n + 4); const uint32_t* s = bsrc + hbase[b];
                const unsigned grid = (unsigned)((n + 255) / 256);
                if (reset_between) CK(cudaCtxResetPersistingL2Cache());
                CK(cudaEventRecord(ev[b]));
                if (b == 0) blocked_pass<true, false><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
                else if (b + 1 == nb) blocked_pass<false, true><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
                else blocked_pass<false, false><<<grid, 256>>>(o, s, state, sum, cnt, out, n, 0.02);
            }
            CK(cudaEventRecord(ev[nb])); CK(cudaEventSynchronize(ev[nb]));
            printf("   per pass [ms | M edges]:");
            for (uint32_t b = 0; b < nb; ++b) { CK(cudaEventElapsedTime(&ms, ev[b], ev[b + 1])); printf(" %.2f|%.0f", ms, (hbase[b + 1] - hbase[b]) / 1e6); }
            printf("\n");
        }
        printf("blocked nb=%2u (block %6.1f MB, edges in block 0: %5.1f%%): %8.3f ms  %7.2f Gedges/s\n", nb, bsize * 8.0 / 1e6,
               100.0 * hbase[1] / E, best, (double)E / best / 1e6);
        if (getenv("QUICK")) { CK(cudaFree(off)); CK(cudaFree(base)); CK(cudaFree(tmp)); continue; }
        float b2;
        run_pass_t<512, 2048, 2>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1, 3);
        run_pass_t<512, 2048, 3>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1, 2);
        run_pass_t<256, 2048, 3>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1, 3);
        run_pass_t<256, 1024, 3>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1, 5);
        run_pass_t<256, 1024, 4>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1, 4);
        b2 = run_pass_w<8>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1);
        printf("   warp-pipelined 8 CTA/SM:  %8.3f ms  %7.2f Gedges/s\n", b2, (double)E / b2 / 1e6);
        b2 = run_pass_w<4, 1>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1);
        printf("   warp-pipelined, no gathers:        %8.3f ms\n", b2);
        b2 = run_pass_w<4, 2>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1);
        printf("   warp-pipelined, no own/acc streams: %8.3f ms\n", b2);
        b2 = run_pass_w<4>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1);
        printf("   warp-pipelined 4 CTA/SM:  %8.3f ms  %7.2f Gedges/s\n", b2, (double)E / b2 / 1e6);
        { cudaMemcpy(sum, out, 8 * 1000, cudaMemcpyDeviceToDevice); }
        if (argc > 4) { CK(cudaFree(off)); CK(cudaFree(base)); CK(cudaFree(tmp)); continue; }
        b2 = run_pass_u<1>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1);
        printf("   U=1 interleaved:  %8.3f ms  %7.2f Gedges/s\n", b2, (double)E / b2 / 1e6);
        b2 = run_pass_u<2>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1);
        printf("   U=2 interleaved:  %8.3f ms  %7.2f Gedges/s\n", b2, (double)E / b2 / 1e6);
        b2 = run_pass_u<4>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1);
        printf("   U=4 interleaved:  %8.3f ms  %7.2f Gedges/s\n", b2, (double)E / b2 / 1e6);
        b2 = run_pass_u<8>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1);
        printf("   U=8 interleaved:  %8.3f ms  %7.2f Gedges/s\n", b2, (double)E / b2 / 1e6);
        b2 = run_pass4<0>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1);
        printf("   4 targets/thread:        %8.3f ms  %7.2f Gedges/s\n", b2, (double)E / b2 / 1e6);
        b2 = run_pass4<1>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1);
        printf("   4 targets/thread, flat:  %8.3f ms  %7.2f Gedges/s\n", b2, (double)E / b2 / 1e6);
        b2 = run_pass2<512, 2048>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1);
        printf("   staged TPB= 512: %8.3f ms  %7.2f Gedges/s\n", b2, (double)E / b2 / 1e6);
        b2 = run_pass2<1024, 3072>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1);
        printf("   staged TPB=1024: %8.3f ms  %7.2f Gedges/s\n", b2, (double)E / b2 / 1e6);
        b2 = run_pass2<2048, 4096>(nb, off, bsrc, hbase, state, sum, cnt, out, n, e0, e1);
        printf("   staged TPB=2048: %8.3f ms  %7.2f Gedges/s\n", b2, (double)E / b2 / 1e6);
        CK(cudaFree(off)); CK(cudaFree(base)); CK(cudaFree(tmp));
    }
    return 0;
}
