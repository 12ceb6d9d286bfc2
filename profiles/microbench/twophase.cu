// twophase.cu — microbenchmark: the prefiltered read phase split into two kernels (Hegselmann-Krause on the 1e8 / 2e9 power-law graph).
//   phase A ("mark"):  per key block, every entry gathers the one-byte key of its source from the L2-resident key column and tests it
//                      against the probe of its row; passing entries are written, compacted and in entry order, to a candidate
//                      buffer (per 32-segment group: at the start of the group's own region).  No accumulators, no shared-memory
//                      queue, no exact states: the kernel is a pure stream + gather.  Variant X also fetches the exact state of a
//                      candidate right away (fire and forget) and writes the VALUE next to the candidate.
//   phase B ("fold"):  per group, the candidates of all blocks are read back (coalesced), exact states gathered (or read from the
//                      value buffer), every lane folds the candidates of its own segment in entry order, finish() runs.
// Why: (1) multi-GPU — phase A needs only the KEYS of the ghosts (1 B each), so the 8-byte state halo can travel while it runs;
//      (2) the fused kernel's queue / flush / per-row counters cost issue slots and occupancy; (3) hub rows: an entry carries the
//      index of its segment inside the group (5 bits above the 27-bit source slot), rows longer than SEGMAX are cut into segments,
//      so no warp ever walks more than 32 x SEGMAX entries and no side pass is needed.
//   usage: twophase [n_agents] [segmax]      env DMAX=<cap of the Pareto degrees, default 1000000>
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
__device__ __forceinline__ uint32_t source_of(uint64_t e, uint64_t n) { const double u = unit(e); uint64_t s = (uint64_t)((double)n * u * u); return (uint32_t)(s >= n ? n - 1 : s); }
__device__ __forceinline__ double ld_gather(const double* p) { double r; asm volatile("ld.global.nc.L1::no_allocate.L2::64B.f64 %0, [%1];" : "=d"(r) : "l"(p)); return r; }
__device__ __forceinline__ uint32_t ld_key(const uint8_t* p, uint64_t pol) { uint32_t r; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol)); return r; }
__device__ __forceinline__ uint64_t keep_policy() { uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol)); return pol; }
__host__ __device__ __forceinline__ uint32_t quant(double o) { const double q = o * 256.0; return q >= 255.0 ? 255u : (q <= 0.0 ? 0u : (uint32_t)q); }
constexpr uint32_t SRC_MASK = 0x7ffffffu;

__global__ void fill_degrees(uint32_t* d, uint64_t n, uint32_t dmax) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n) return;
    if (t == n) { d[t] = 0; return; }
    const double u = unit(t ^ 0x9e3779b97f4a7c15ull);
    const double x = 6.8333 * pow(1.0 - u, -2.0 / 3.0);
    d[t] = (x >= (double)dmax ? dmax : (uint32_t)x) + 1;
}
__global__ void fill_direct(uint32_t* src, const uint32_t* __restrict__ roff, uint64_t n) {      // a warp per row (hub rows are long)
    const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (w >= n) return;
    const uint32_t b = roff[w], e = roff[w + 1];
    for (uint32_t k = b + lane; k + 1 < e; k += 32) src[k] = source_of(k, n);
    if (lane == 0) src[e - 1] = (uint32_t)w;
}
__global__ void fill_state(double* s, uint64_t n) { const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (t < n) s[t] = unit(t + 0x51ed270b1ull); }
__global__ void build_keys(const double* __restrict__ s, uint8_t* __restrict__ key, uint64_t n) {
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (t + 3 < n) {
        const double2 a = *reinterpret_cast<const double2*>(s + t), b = *reinterpret_cast<const double2*>(s + t + 2);
        *reinterpret_cast<uint32_t*>(key + t) = quant(a.x) | (quant(a.y) << 8) | (quant(b.x) << 16) | (quant(b.y) << 24);
    } else for (uint64_t i = t; i < n; ++i) key[i] = (uint8_t)quant(s[i]);
}
// reference: a warp per row, exact gathers for every entry
__global__ void direct_step(const uint32_t* __restrict__ src, const uint32_t* __restrict__ roff, const double* __restrict__ state, double* __restrict__ out, uint32_t* __restrict__ outc,
                            uint64_t n, double eps) {
    const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (g >= n) return;
    const double own = state[g];
    double s = 0; uint32_t c = 0;
    for (uint32_t k = roff[g] + lane; k < roff[g + 1]; k += 32) { const double v = ld_gather(state + __ldcs(src + k)); if (fabs(v - own) < eps) { s += v; c += 1; } }
    for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if (lane == 0) { out[g] = s / (double)c; outc[g] = c; }
}

// ---- segments: row r is cut into max(1, ceil(len / segmax)) segments -------------------------------------------------------------
__global__ void seg_counts(const uint32_t* __restrict__ roff, uint64_t n, uint32_t segmax, uint32_t* __restrict__ cnt) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n) return;
    if (t == n) { cnt[t] = 0; return; }
    const uint32_t len = roff[t + 1] - roff[t];
    cnt[t] = len <= segmax ? 1u : (len + segmax - 1) / segmax;
}
// seg_row[s] = row | MULTI for segments of rows with several segments; seg_lo[s] = first entry (CSR position); seg_lo[nseg] = E
constexpr uint32_t MULTI = 0x80000000u;
__global__ void seg_fill(const uint32_t* __restrict__ roff, const uint32_t* __restrict__ sfirst, uint64_t n, uint32_t segmax, uint32_t* __restrict__ seg_row, uint32_t* __restrict__ seg_lo) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t s0 = sfirst[t], ns = sfirst[t + 1] - s0, lo = roff[t];
    for (uint32_t k = 0; k < ns; ++k) { seg_row[s0 + k] = (uint32_t)t | (ns > 1 ? MULTI : 0u); seg_lo[s0 + k] = lo + k * segmax; }
    if (t == n - 1) seg_lo[s0 + ns] = roff[n];
}
__global__ void count_blocks(const uint32_t* __restrict__ src, const uint32_t* __restrict__ seg_lo, uint64_t nseg, uint64_t pad, uint32_t bsize, uint32_t* cnt) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nseg) return;
    for (uint32_t k = seg_lo[t]; k < seg_lo[t + 1]; ++k) { const uint32_t b = src[k] / bsize; cnt[(uint64_t)b * pad + t] += 1; }
}
__global__ void fill_blocks(const uint32_t* __restrict__ src, const uint32_t* __restrict__ seg_lo, uint64_t nseg, uint64_t pad, uint32_t bsize, uint32_t nb, const uint32_t* off,
                            const uint64_t* base, uint32_t* bsrc) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nseg) return;
    uint32_t fill[8];
    for (uint32_t b = 0; b < nb; ++b) fill[b] = 0;
    const uint32_t tag = (uint32_t)(t & 31) << 27;
    for (uint32_t k = seg_lo[t]; k < seg_lo[t + 1]; ++k) {
        const uint32_t s = src[k], b = s / bsize;
        bsrc[base[b] + off[(uint64_t)b * pad + t] + fill[b]++] = (s - b * bsize) | tag;
    }
}

// ---- phase A -----------------------------------------------------------------------------------------------------------------------
// bsrc / off / cand / gcnt are the block's own (already offset) arrays.  X = 1: also fetch the exact state of every candidate.
template <int U, int WPB, int MINB, int X>
__global__ void __launch_bounds__(32 * WPB, MINB) phase_a(const uint32_t* __restrict__ bsrc, const uint32_t* __restrict__ off, const uint32_t* __restrict__ seg_row,
                                                           const uint8_t* __restrict__ key_all, const uint8_t* __restrict__ key_blk, const double* __restrict__ state_blk,
                                                           uint32_t* __restrict__ cand, double* __restrict__ cval, uint32_t* __restrict__ gcnt, uint64_t nseg, int band,
                                                           uint64_t g_first = 0, uint64_t g_end = ~0ull) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t g = g_first + (uint64_t)blockIdx.x * WPB + (threadIdx.x >> 5);
    if (g >= g_end) return;
    const uint64_t s0 = g * 32;
    if (s0 >= nseg) return;
    const uint32_t ns = (uint32_t)(nseg - s0 < 32 ? nseg - s0 : 32);
    const uint64_t pol = keep_policy();
    const uint32_t e0 = __ldcs(off + s0), e1 = __ldcs(off + s0 + ns);
    const int own = lane < ns ? (int)key_all[__ldcs(seg_row + s0 + lane) & ~MULTI] : -100000;       // the probe of my segment's row
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t cn = 0;
    for (uint32_t base = e0; base < e1; base += 32 * U) {
        uint32_t idx[U]; int ks[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const uint32_t e = base + u * 32 + lane; idx[u] = e < e1 ? __ldcs(bsrc + e) : 0xffffffffu; }
#pragma unroll
        for (int u = 0; u < U; ++u) ks[u] = idx[u] != 0xffffffffu ? (int)ld_key(key_blk + (idx[u] & SRC_MASK), pol) : 1000000;
        double v[U]; uint32_t pos[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int probe = __shfl_sync(0xffffffffu, own, idx[u] >> 27);
            const bool pass = abs(ks[u] - probe) <= band;
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            pos[u] = pass ? e0 + cn + __popc(m & lt) : 0xffffffffu;
            if (pass) { cand[pos[u]] = idx[u]; if (X) v[u] = ld_gather(state_blk + (idx[u] & SRC_MASK)); }
            cn += __popc(m);
        }
        if (X) {
#pragma unroll
            for (int u = 0; u < U; ++u) if (pos[u] != 0xffffffffu) cval[pos[u]] = v[u];
        }
    }
    if (lane == 0) gcnt[g] = cn;
}

// ---- phase B -----------------------------------------------------------------------------------------------------------------------
struct BArgs {
    const uint32_t* off[8]; const uint32_t* cand[8]; const double* cval[8]; const uint32_t* gcnt[8]; uint32_t nb, bsize;
};
template <int X>
__global__ void __launch_bounds__(256) phase_b(const BArgs a, const uint32_t* __restrict__ seg_row, const double* __restrict__ state, double* __restrict__ out, uint32_t* __restrict__ outc,
                                               double* __restrict__ seg_sum, uint32_t* __restrict__ seg_cnt, uint64_t nseg, double eps) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t s0 = g * 32;
    if (s0 >= nseg) return;
    const uint32_t ns = (uint32_t)(nseg - s0 < 32 ? nseg - s0 : 32);
    const uint32_t rowm = lane < ns ? __ldcs(seg_row + s0 + lane) : 0u;
    const double own = lane < ns ? state[rowm & ~MULTI] : 0.0;
    double s = 0; uint32_t c = 0;
    for (uint32_t b = 0; b < a.nb; ++b) {
        const uint32_t e0 = __ldcs(a.off[b] + s0), cn = __ldcs(a.gcnt[b] + g);
        for (uint32_t i0 = 0; i0 < cn; i0 += 32) {
            const uint32_t i = i0 + lane;
            uint32_t e = 0xffffffffu; double v = 0.0;
            if (i < cn) { e = __ldcs(a.cand[b] + e0 + i); v = X ? __ldcs(a.cval[b] + e0 + i) : ld_gather(state + (size_t)b * a.bsize + (e & SRC_MASK)); }
            const uint32_t m = cn - i0 < 32 ? cn - i0 : 32;
            for (uint32_t j = 0; j < m; ++j) {                          // entry order; the owner lane of candidate j folds it
                const uint32_t tag = __shfl_sync(0xffffffffu, e, j) >> 27;
                const double vj = __shfl_sync(0xffffffffu, v, j);
                if (lane == tag && fabs(vj - own) < eps) { s += vj; c += 1; }
            }
        }
    }
    if (lane < ns) {
        if (rowm & MULTI) { seg_sum[s0 + lane] = s; seg_cnt[s0 + lane] = c; }
        else { __stcs(out + rowm, s / (double)c); __stcs(outc + rowm, c); }
    }
}

// phase B, second shape: the first candidate chunk of every block is loaded up front (the blocks' chains cand -> exact state run
// side by side), and a lane finds the run of its own segment's candidates in a chunk by a binary search over the (ascending) tags
// held by the lanes instead of walking all candidates.
template <int NB>
__global__ void __launch_bounds__(128) phase_b2(const BArgs a, const uint32_t* __restrict__ seg_row, const double* __restrict__ state, double* __restrict__ out, uint32_t* __restrict__ outc,
                                                double* __restrict__ seg_sum, uint32_t* __restrict__ seg_cnt, uint64_t nseg, double eps, uint64_t g_first, uint64_t g_end) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t g = g_first + (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (g >= g_end) return;
    const uint64_t s0 = g * 32;
    const uint32_t ns = (uint32_t)(nseg - s0 < 32 ? nseg - s0 : 32);
    uint32_t e0[NB], cn[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) { e0[b] = __ldcs(a.off[b] + s0); cn[b] = __ldcs(a.gcnt[b] + g); }
    const uint32_t rowm = lane < ns ? __ldcs(seg_row + s0 + lane) : 0u;
    const double own = lane < ns ? state[rowm & ~MULTI] : 0.0;
    double s = 0; uint32_t c = 0;
    uint32_t e[NB]; double v[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) e[b] = lane < cn[b] ? __ldcs(a.cand[b] + e0[b] + lane) : 0xffffffffu;
#pragma unroll
    for (int b = 0; b < NB; ++b) v[b] = lane < cn[b] ? ld_gather(state + (size_t)b * a.bsize + (e[b] & SRC_MASK)) : 0.0;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        for (uint32_t i0 = 0; i0 < cn[b]; i0 += 32) {
            if (i0) {
                const uint32_t i = i0 + lane;
                e[b] = i < cn[b] ? __ldcs(a.cand[b] + e0[b] + i) : 0xffffffffu;
                v[b] = i < cn[b] ? ld_gather(state + (size_t)b * a.bsize + (e[b] & SRC_MASK)) : 0.0;
            }
            const uint32_t m = cn[b] - i0 < 32 ? cn[b] - i0 : 32;
            const uint32_t tagv = lane < m ? e[b] >> 27 : 32u;
            uint32_t pos = 0;                                            // first candidate of the chunk whose tag is >= my lane
#pragma unroll
            for (int st = 16; st; st >>= 1) { const uint32_t t = __shfl_sync(0xffffffffu, tagv, pos + st - 1); if (t < lane) pos += st; }
            { const uint32_t t = __shfl_sync(0xffffffffu, tagv, pos); if (t < lane) pos += 1; }
            uint32_t end = __shfl_down_sync(0xffffffffu, pos, 1);
            if (lane == 31) end = m;
            const uint32_t mine = end - pos;
            const uint32_t maxc = __reduce_max_sync(0xffffffffu, mine);
            for (uint32_t j = 0; j < maxc; ++j) {
                const double vj = __shfl_sync(0xffffffffu, v[b], (pos + j) & 31);
                if (j < mine && fabs(vj - own) < eps) { s += vj; c += 1; }
            }
        }
    }
    if (lane < ns) {
        if (rowm & MULTI) { seg_sum[s0 + lane] = s; seg_cnt[s0 + lane] = c; }
        else { __stcs(out + rowm, s / (double)c); __stcs(outc + rowm, c); }
    }
}
// ---- fused, deferred: one kernel per key block as in the engine's prefiltered sweep, but (1) the segment of an entry is its 5-bit
// tag, the probe comes from the owner lane by shuffle (no offset search, no shared memory), (2) the exact states of a chunk's
// candidates are requested at the end of iteration i and folded in iteration i + 1 after the next chunk's index loads and key
// gathers have been issued: the DRAM round trip of the exact states overlaps the L2 round trip of the next keys.
template <int U, int WPB, int MINB, bool FIRST, bool LAST>
__global__ void __launch_bounds__(32 * WPB, MINB) fused_deferred(const uint32_t* __restrict__ bsrc, const uint32_t* __restrict__ off, const uint32_t* __restrict__ seg_row,
                                                                  const uint8_t* __restrict__ key_blk, const double* __restrict__ state, const double* __restrict__ state_blk,
                                                                  double* __restrict__ out, uint32_t* __restrict__ outc, double* __restrict__ seg_sum, uint32_t* __restrict__ seg_cnt,
                                                                  uint64_t nseg, double eps, int band) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t g = (uint64_t)blockIdx.x * WPB + (threadIdx.x >> 5);
    const uint64_t s0 = g * 32;
    if (s0 >= nseg) return;
    const uint32_t ns = (uint32_t)(nseg - s0 < 32 ? nseg - s0 : 32);
    const uint64_t pol = keep_policy();
    const uint32_t e0 = __ldcs(off + s0), e1 = __ldcs(off + s0 + ns);
    const uint32_t rowm = lane < ns ? __ldcs(seg_row + s0 + lane) : 0u;
    const double own = lane < ns ? state[rowm & ~MULTI] : -1e300;
    const int probe = lane < ns ? (int)quant(own) : -100000;
    double s = 0; uint32_t c = 0;
    if (!FIRST && lane < ns) { s = __ldcs(seg_sum + s0 + lane); c = __ldcs(seg_cnt + s0 + lane); }
    uint32_t pidx[U], pmask[U]; double pv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { pmask[u] = 0; pidx[u] = 0; pv[u] = 0.0; }
    for (uint32_t base = e0; base < e1; base += 32 * U) {
        uint32_t idx[U]; int ks[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const uint32_t e = base + u * 32 + lane; idx[u] = e < e1 ? __ldcs(bsrc + e) : 0xffffffffu; }
#pragma unroll
        for (int u = 0; u < U; ++u) ks[u] = idx[u] != 0xffffffffu ? (int)ld_key(key_blk + (idx[u] & SRC_MASK), pol) : 1000000;
        // fold the previous chunk's candidates (entry order: u ascending, lanes ascending)
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t m = pmask[u];
            while (m) {
                const int j = __ffs(m) - 1; m &= m - 1;
                const uint32_t tag = __shfl_sync(0xffffffffu, pidx[u], j) >> 27;
                const double vj = __shfl_sync(0xffffffffu, pv[u], j);
                if (lane == tag && fabs(vj - own) < eps) { s += vj; c += 1; }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pr = __shfl_sync(0xffffffffu, probe, idx[u] >> 27);
            const bool pass = abs(ks[u] - pr) <= band;
            pmask[u] = __ballot_sync(0xffffffffu, pass);
            pidx[u] = idx[u];
            if (pass) pv[u] = ld_gather(state_blk + (idx[u] & SRC_MASK));
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        uint32_t m = pmask[u];
        while (m) {
            const int j = __ffs(m) - 1; m &= m - 1;
            const uint32_t tag = __shfl_sync(0xffffffffu, pidx[u], j) >> 27;
            const double vj = __shfl_sync(0xffffffffu, pv[u], j);
            if (lane == tag && fabs(vj - own) < eps) { s += vj; c += 1; }
        }
    }
    if (lane < ns) {
        if (!LAST || (rowm & MULTI)) { __stcs(seg_sum + s0 + lane, s); __stcs(seg_cnt + s0 + lane, c); }
        else { __stcs(out + rowm, s / (double)c); __stcs(outc + rowm, c); }
    }
}
// ---- fused, queued: the engine's sweep shape (ordered queue in shared memory, exact states fetched 32 at a time when the queue fills)
// with the tag / shuffle row lookup and without per-row counters: at a flush a lane finds the run of its segment by a binary search
// over the (ascending) tags of the queue.
template <int U, int QEXTRA> struct QStage { uint32_t qe[32 * U + QEXTRA]; double qv[32 * U + QEXTRA]; };
template <int U, int QEXTRA, int WPB, int MINB, bool FIRST, bool LAST>
__global__ void __launch_bounds__(32 * WPB, MINB) fused_queued(const uint32_t* __restrict__ bsrc, const uint32_t* __restrict__ off, const uint32_t* __restrict__ seg_row,
                                                                const uint8_t* __restrict__ key_blk, const double* __restrict__ state, const double* __restrict__ state_blk,
                                                                double* __restrict__ out, uint32_t* __restrict__ outc, double* __restrict__ seg_sum, uint32_t* __restrict__ seg_cnt,
                                                                uint64_t nseg, double eps, int band) {
    constexpr int QCAP = 32 * U + QEXTRA;
    __shared__ QStage<U, QEXTRA> stages[WPB];
    QStage<U, QEXTRA>& sm = stages[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t g = (uint64_t)blockIdx.x * WPB + (threadIdx.x >> 5);
    const uint64_t s0 = g * 32;
    if (s0 >= nseg) return;
    const uint32_t ns = (uint32_t)(nseg - s0 < 32 ? nseg - s0 : 32);
    const uint64_t pol = keep_policy();
    const uint32_t e0 = __ldcs(off + s0), e1 = __ldcs(off + s0 + ns);
    const uint32_t rowm = lane < ns ? __ldcs(seg_row + s0 + lane) : 0u;
    const double own = lane < ns ? state[rowm & ~MULTI] : -1e300;
    const int probe = lane < ns ? (int)quant(own) : -100000;
    const uint32_t lt = (1u << lane) - 1u;
    double s = 0; uint32_t c = 0, qn = 0;
    if (!FIRST && lane < ns) { s = __ldcs(seg_sum + s0 + lane); c = __ldcs(seg_cnt + s0 + lane); }
    for (uint32_t base = e0; base < e1; base += 32 * U) {
        uint32_t idx[U]; int ks[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const uint32_t e = base + u * 32 + lane; idx[u] = e < e1 ? __ldcs(bsrc + e) : 0xffffffffu; }
#pragma unroll
        for (int u = 0; u < U; ++u) ks[u] = idx[u] != 0xffffffffu ? (int)ld_key(key_blk + (idx[u] & SRC_MASK), pol) : 1000000;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pr = __shfl_sync(0xffffffffu, probe, idx[u] >> 27);
            const bool pass = abs(ks[u] - pr) <= band;
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            if (pass) sm.qe[qn + __popc(m & lt)] = idx[u];
            qn += __popc(m);
        }
        __syncwarp();
        if (qn > (uint32_t)QEXTRA || base + 32 * U >= e1) {
            for (uint32_t i = lane; i < qn; i += 32) sm.qv[i] = ld_gather(state_blk + (sm.qe[i] & SRC_MASK));
            uint32_t lo = 0, hi = qn;                                     // first queued candidate whose tag is >= lane
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((sm.qe[mid] >> 27) < lane) lo = mid + 1; else hi = mid; }
            uint32_t end = __shfl_down_sync(0xffffffffu, lo, 1);
            if (lane == 31) end = qn;
            __syncwarp();
            for (uint32_t j = lo; j < end; ++j) { const double v = sm.qv[j]; if (fabs(v - own) < eps) { s += v; c += 1; } }
            qn = 0;
            __syncwarp();
        }
    }
    if (lane < ns) {
        if (!LAST || (rowm & MULTI)) { __stcs(seg_sum + s0 + lane, s); __stcs(seg_cnt + s0 + lane, c); }
        else { __stcs(out + rowm, s / (double)c); __stcs(outc + rowm, c); }
    }
}
// rows with several segments: merge the partial accumulators in segment order, finish
__global__ void merge_hubs(const uint32_t* __restrict__ hub_rows, uint32_t nh, const uint32_t* __restrict__ sfirst, const double* __restrict__ seg_sum, const uint32_t* __restrict__ seg_cnt,
                           double* __restrict__ out, uint32_t* __restrict__ outc) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nh) return;
    const uint32_t r = hub_rows[t];
    double s = 0; uint32_t c = 0;
    for (uint32_t k = sfirst[r]; k < sfirst[r + 1]; ++k) { s += seg_sum[k]; c += seg_cnt[k]; }
    out[r] = s / (double)c; outc[r] = c;
}
__global__ void hub_flags(const uint32_t* __restrict__ sfirst, uint64_t n, uint32_t* __restrict__ flag) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) flag[t] = sfirst[t + 1] - sfirst[t] > 1 ? 1u : 0u;
}
__global__ void compact_rows(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos, uint64_t n, uint32_t* __restrict__ outrows) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && flag[t]) outrows[pos[t]] = (uint32_t)t;
}
__global__ void compare(const double* __restrict__ a, const uint32_t* __restrict__ ac, const double* __restrict__ b, const uint32_t* __restrict__ bc, uint64_t n, unsigned long long* bad) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (ac[t] != bc[t]) atomicAdd(bad, 1ull);
    else if (fabs(a[t] - b[t]) > 1e-12 * fabs(a[t])) atomicAdd(bad + 1, 1ull);
}
static void scan_u32(uint32_t* p, uint64_t n) {
    void* t = nullptr; size_t tz = 0;
    cub::DeviceScan::ExclusiveSum(t, tz, p, p, (int)n); CK(cudaMalloc(&t, tz)); cub::DeviceScan::ExclusiveSum(t, tz, p, p, (int)n); CK(cudaFree(t));
}

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IONBF, 0);
    const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 100000000ull;
    const uint32_t segmax = argc > 2 ? atoi(argv[2]) : 2048;
    const uint32_t dmax = getenv("DMAX") ? atoi(getenv("DMAX")) : 1000000;
    const uint32_t nb = getenv("NB") ? atoi(getenv("NB")) : 2;
    const double eps = getenv("EPS") ? atof(getenv("EPS")) : 0.02;
    const int band = (int)std::floor(eps * 256.0) + 1;
    if (n >= (1ull << 27) * nb) { printf("a key block must stay below 2^27 slots\n"); return 1; }
    { int maxp = 0; cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, 0); cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp); }
    uint32_t* roff; CK(cudaMalloc(&roff, (n + 1) * 4));
    fill_degrees<<<(unsigned)((n + 256) / 256), 256>>>(roff, n, dmax);
    scan_u32(roff, n + 1);
    uint32_t E32; CK(cudaMemcpy(&E32, roff + n, 4, cudaMemcpyDeviceToHost));
    const uint64_t E = E32;
    double *state, *ref, *out; uint32_t *refc, *outc, *src; uint8_t* key; unsigned long long* bad;
    CK(cudaMalloc(&state, n * 8)); CK(cudaMalloc(&ref, n * 8)); CK(cudaMalloc(&out, n * 8)); CK(cudaMalloc(&refc, n * 4)); CK(cudaMalloc(&outc, n * 4));
    CK(cudaMalloc(&src, E * 4)); CK(cudaMalloc(&key, n + 64)); CK(cudaMalloc(&bad, 16));
    fill_state<<<(unsigned)((n + 255) / 256), 256>>>(state, n);
    fill_direct<<<(unsigned)((n * 32 + 255) / 256), 256>>>(src, roff, n);
    CK(cudaDeviceSynchronize());
    printf("n=%llu E=%llu dmax=%u segmax=%u nb=%u eps=%.3f band=%d\n", (unsigned long long)n, (unsigned long long)E, dmax, segmax, nb, eps, band);
    cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    float ms;
    CK(cudaEventRecord(e0)); direct_step<<<(unsigned)((n * 32 + 255) / 256), 256>>>(src, roff, state, ref, refc, n, eps); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1)); printf("direct, warp per row (reference values)        %8.3f ms\n", ms);
    CK(cudaEventRecord(e0)); build_keys<<<(unsigned)((n / 4 + 256) / 256), 256>>>(state, key, n); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1)); printf("build_keys                                     %8.3f ms\n", ms);
    // segments
    uint32_t* sfirst; CK(cudaMalloc(&sfirst, (n + 1) * 4));
    seg_counts<<<(unsigned)((n + 256) / 256), 256>>>(roff, n, segmax, sfirst);
    scan_u32(sfirst, n + 1);
    uint32_t nseg32; CK(cudaMemcpy(&nseg32, sfirst + n, 4, cudaMemcpyDeviceToHost));
    const uint64_t nseg = nseg32, pad = (nseg + 4 + 3) & ~3ull;
    uint32_t *seg_row, *seg_lo; CK(cudaMalloc(&seg_row, (nseg + 1) * 4)); CK(cudaMalloc(&seg_lo, (nseg + 1) * 4));
    seg_fill<<<(unsigned)((n + 255) / 256), 256>>>(roff, sfirst, n, segmax, seg_row, seg_lo);
    // hub rows (several segments)
    uint32_t *hflag, *hpos, *hub_rows; CK(cudaMalloc(&hflag, (n + 1) * 4)); CK(cudaMalloc(&hpos, (n + 1) * 4));
    CK(cudaMemset(hflag, 0, (n + 1) * 4));
    hub_flags<<<(unsigned)((n + 255) / 256), 256>>>(sfirst, n, hflag);
    CK(cudaMemcpy(hpos, hflag, (n + 1) * 4, cudaMemcpyDeviceToDevice)); scan_u32(hpos, n + 1);
    uint32_t nh; CK(cudaMemcpy(&nh, hpos + n, 4, cudaMemcpyDeviceToHost));
    CK(cudaMalloc(&hub_rows, (size_t)(nh + 1) * 4));
    compact_rows<<<(unsigned)((n + 255) / 256), 256>>>(hflag, hpos, n, hub_rows);
    CK(cudaFree(hflag)); CK(cudaFree(hpos));
    printf("segments: %llu (%llu more than rows), rows with several segments: %u\n", (unsigned long long)nseg, (unsigned long long)(nseg - n), nh);
    // blocked view over segments
    const uint32_t bsize = (uint32_t)((n + nb - 1) / nb);
    uint32_t* off; uint64_t* base; CK(cudaMalloc(&off, (uint64_t)nb * pad * 4)); CK(cudaMalloc(&base, nb * 8));
    CK(cudaMemset(off, 0, (uint64_t)nb * pad * 4));
    count_blocks<<<(unsigned)((nseg + 255) / 256), 256>>>(src, seg_lo, nseg, pad, bsize, off);
    std::vector<uint64_t> hbase(nb + 1, 0);
    for (uint32_t k = 0; k < nb; ++k) {
        scan_u32(off + (uint64_t)k * pad, nseg + 1);
        uint32_t tot; CK(cudaMemcpy(&tot, off + (uint64_t)k * pad + nseg, 4, cudaMemcpyDeviceToHost));
        hbase[k + 1] = hbase[k] + tot;
    }
    CK(cudaMemcpy(base, hbase.data(), nb * 8, cudaMemcpyHostToDevice));
    uint32_t *bsrc, *cand, *gcnt, *seg_cnt; double *cval, *seg_sum;
    const uint64_t ngroups = (nseg + 31) / 32;
    CK(cudaMalloc(&bsrc, E * 4 + 64)); CK(cudaMalloc(&cand, E * 4 + 64)); CK(cudaMalloc(&cval, E * 8 + 64)); CK(cudaMalloc(&gcnt, (uint64_t)nb * ngroups * 4));
    CK(cudaMalloc(&seg_sum, nseg * 8)); CK(cudaMalloc(&seg_cnt, nseg * 4));
    fill_blocks<<<(unsigned)((nseg + 255) / 256), 256>>>(src, seg_lo, nseg, pad, bsize, nb, off, base, bsrc);
    CK(cudaDeviceSynchronize());
    printf("block entries [M]:"); for (uint32_t k = 0; k < nb; ++k) printf(" %.0f", (hbase[k + 1] - hbase[k]) / 1e6); printf("\n");

    BArgs ba{}; ba.nb = nb; ba.bsize = bsize;
    for (uint32_t k = 0; k < nb; ++k) { ba.off[k] = off + (uint64_t)k * pad; ba.cand[k] = cand + hbase[k]; ba.cval[k] = cval + hbase[k]; ba.gcnt[k] = gcnt + (uint64_t)k * ngroups; }
    auto check = [&](const char* name, float msa, float msb) {
        unsigned long long h[2] = {0, 0};
        CK(cudaMemset(bad, 0, 16));
        compare<<<(unsigned)((n + 255) / 256), 256>>>(ref, refc, out, outc, n, bad);
        CK(cudaMemcpy(h, bad, 16, cudaMemcpyDeviceToHost));
        printf("%-46s A %7.3f ms (%6.1f G entries/s)  B %6.3f ms  total %7.3f ms   count mismatches %llu, value > 1e-12 rel %llu\n", name, msa, (double)E / msa / 1e6, msb, msa + msb, h[0], h[1]);
    };
    auto run = [&](const char* name, auto kernA, int wpb, int x) {
        float best_a = 1e9f, best_b = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaMemset(out, 0, n * 8)); CK(cudaMemset(outc, 0, n * 4));
            CK(cudaEventRecord(e0));
            for (uint32_t k = 0; k < nb; ++k)
                kernA<<<(unsigned)((ngroups + wpb - 1) / wpb), 32 * wpb>>>(bsrc + hbase[k], off + (uint64_t)k * pad, seg_row, key, key + (uint64_t)k * bsize, state + (uint64_t)k * bsize,
                                                                           cand + hbase[k], cval + hbase[k], gcnt + (uint64_t)k * ngroups, nseg, band, 0, ~0ull);
            CK(cudaEventRecord(e1));
            if (x) phase_b<1><<<(unsigned)((ngroups * 32 + 255) / 256), 256>>>(ba, seg_row, state, out, outc, seg_sum, seg_cnt, nseg, eps);
            else phase_b<0><<<(unsigned)((ngroups * 32 + 255) / 256), 256>>>(ba, seg_row, state, out, outc, seg_sum, seg_cnt, nseg, eps);
            if (nh) merge_hubs<<<(nh + 255) / 256, 256>>>(hub_rows, nh, sfirst, seg_sum, seg_cnt, out, outc);
            CK(cudaEventRecord(e2)); CK(cudaEventSynchronize(e2));
            CK(cudaGetLastError());
            float a, b; CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, e1, e2));
            if (a + b < best_a + best_b) { best_a = a; best_b = b; }
        }
        check(name, best_a, best_b);
    };
    if (getenv("SWEEP")) {
    run("A: U=2, 2 warps/CTA, 32 CTAs/SM; B gathers", phase_a<2, 2, 32, 0>, 2, 0);
    run("A: U=4, 2 warps/CTA, 32 CTAs/SM; B gathers", phase_a<4, 2, 32, 0>, 2, 0);
    run("A: U=4, 2 warps/CTA, 24 CTAs/SM; B gathers", phase_a<4, 2, 24, 0>, 2, 0);
    run("A: U=8, 2 warps/CTA, 24 CTAs/SM; B gathers", phase_a<8, 2, 24, 0>, 2, 0);
    run("A: U=8, 2 warps/CTA, 16 CTAs/SM; B gathers", phase_a<8, 2, 16, 0>, 2, 0);
    run("A: U=4, 4 warps/CTA, 16 CTAs/SM; B gathers", phase_a<4, 4, 16, 0>, 4, 0);
    run("A: U=4, 8 warps/CTA,  8 CTAs/SM; B gathers", phase_a<4, 8, 8, 0>, 8, 0);
    run("A: U=4, 1 warp /CTA, 32 CTAs/SM; B gathers", phase_a<4, 1, 32, 0>, 1, 0);
    run("AX: U=2, 2 warps/CTA, 32 CTAs/SM; B streams", phase_a<2, 2, 32, 1>, 2, 1);
    run("AX: U=4, 2 warps/CTA, 24 CTAs/SM; B streams", phase_a<4, 2, 24, 1>, 2, 1);
    run("AX: U=4, 2 warps/CTA, 16 CTAs/SM; B streams", phase_a<4, 2, 16, 1>, 2, 1);
    run("AX: U=8, 2 warps/CTA, 16 CTAs/SM; B streams", phase_a<8, 2, 16, 1>, 2, 1);
    run("AX: U=4, 4 warps/CTA, 12 CTAs/SM; B streams", phase_a<4, 4, 12, 1>, 4, 1);
    } else run("A: U=8, 2 warps/CTA, 24 CTAs/SM; B gathers", phase_a<8, 2, 24, 0>, 2, 0);
    // ---- fused, deferred folds ----------------------------------------------------------------------------------------------------
    {
        auto fused = [&](const char* name, auto kf, auto km, auto kl, int wpb) {
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaMemset(out, 0, n * 8)); CK(cudaMemset(outc, 0, n * 4));
                CK(cudaEventRecord(e0));
                for (uint32_t k = 0; k < nb; ++k) {
                    const unsigned grid = (unsigned)((ngroups + wpb - 1) / wpb);
                    const uint32_t* sp = bsrc + hbase[k]; const uint32_t* op = off + (uint64_t)k * pad;
                    if (k == 0) kf<<<grid, 32 * wpb>>>(sp, op, seg_row, key + (uint64_t)k * bsize, state, state + (uint64_t)k * bsize, out, outc, seg_sum, seg_cnt, nseg, eps, band);
                    else if (k + 1 == nb) kl<<<grid, 32 * wpb>>>(sp, op, seg_row, key + (uint64_t)k * bsize, state, state + (uint64_t)k * bsize, out, outc, seg_sum, seg_cnt, nseg, eps, band);
                    else km<<<grid, 32 * wpb>>>(sp, op, seg_row, key + (uint64_t)k * bsize, state, state + (uint64_t)k * bsize, out, outc, seg_sum, seg_cnt, nseg, eps, band);
                }
                if (nh) merge_hubs<<<(nh + 255) / 256, 256>>>(hub_rows, nh, sfirst, seg_sum, seg_cnt, out, outc);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
                float t; CK(cudaEventElapsedTime(&t, e0, e1)); if (t < best) best = t;
            }
            check(name, best, 0.f);
        };
        if (getenv("SWEEP")) fused("fused deferred, U=2, 2 warps, 24 CTAs/SM", fused_deferred<2, 2, 24, true, false>, fused_deferred<2, 2, 24, false, false>, fused_deferred<2, 2, 24, false, true>, 2);
        if (getenv("SWEEP")) fused("fused deferred, U=4, 2 warps, 24 CTAs/SM", fused_deferred<4, 2, 24, true, false>, fused_deferred<4, 2, 24, false, false>, fused_deferred<4, 2, 24, false, true>, 2);
        fused("fused deferred, U=4, 2 warps, 20 CTAs/SM", fused_deferred<4, 2, 20, true, false>, fused_deferred<4, 2, 20, false, false>, fused_deferred<4, 2, 20, false, true>, 2);
        if (getenv("SWEEP")) fused("fused deferred, U=4, 2 warps, 16 CTAs/SM", fused_deferred<4, 2, 16, true, false>, fused_deferred<4, 2, 16, false, false>, fused_deferred<4, 2, 16, false, true>, 2);
        if (getenv("SWEEP")) fused("fused deferred, U=8, 2 warps, 16 CTAs/SM", fused_deferred<8, 2, 16, true, false>, fused_deferred<8, 2, 16, false, false>, fused_deferred<8, 2, 16, false, true>, 2);
        if (getenv("SWEEP")) fused("fused deferred, U=4, 4 warps, 12 CTAs/SM", fused_deferred<4, 4, 12, true, false>, fused_deferred<4, 4, 12, false, false>, fused_deferred<4, 4, 12, false, true>, 4);
        fused("fused deferred, U=4, 1 warp , 32 CTAs/SM", fused_deferred<4, 1, 32, true, false>, fused_deferred<4, 1, 32, false, false>, fused_deferred<4, 1, 32, false, true>, 1);

        fused("fused queued, U=2 q+32, 2 warps, 32 CTAs/SM", fused_queued<2, 32, 2, 32, true, false>, fused_queued<2, 32, 2, 32, false, false>, fused_queued<2, 32, 2, 32, false, true>, 2);
        fused("fused queued, U=2 q+64, 2 warps, 32 CTAs/SM", fused_queued<2, 64, 2, 32, true, false>, fused_queued<2, 64, 2, 32, false, false>, fused_queued<2, 64, 2, 32, false, true>, 2);
        fused("fused queued, U=4 q+32, 2 warps, 24 CTAs/SM", fused_queued<4, 32, 2, 24, true, false>, fused_queued<4, 32, 2, 24, false, false>, fused_queued<4, 32, 2, 24, false, true>, 2);
        fused("fused queued, U=4 q+64, 2 warps, 24 CTAs/SM", fused_queued<4, 64, 2, 24, true, false>, fused_queued<4, 64, 2, 24, false, false>, fused_queued<4, 64, 2, 24, false, true>, 2);
        fused("fused queued, U=4 q+64, 2 warps, 32 CTAs/SM", fused_queued<4, 64, 2, 32, true, false>, fused_queued<4, 64, 2, 32, false, false>, fused_queued<4, 64, 2, 32, false, true>, 2);
        fused("fused queued, U=4 q+128, 2 warps, 24 CTAs/SM", fused_queued<4, 128, 2, 24, true, false>, fused_queued<4, 128, 2, 24, false, false>, fused_queued<4, 128, 2, 24, false, true>, 2);
        fused("fused queued, U=8 q+64, 2 warps, 24 CTAs/SM", fused_queued<8, 64, 2, 24, true, false>, fused_queued<8, 64, 2, 24, false, false>, fused_queued<8, 64, 2, 24, false, true>, 2);
        fused("fused queued, U=4 q+64, 4 warps, 12 CTAs/SM", fused_queued<4, 64, 4, 12, true, false>, fused_queued<4, 64, 4, 12, false, false>, fused_queued<4, 64, 4, 12, false, true>, 4);
        fused("fused queued, U=4 q+64, 1 warp , 32 CTAs/SM", fused_queued<4, 64, 1, 32, true, false>, fused_queued<4, 64, 1, 32, false, false>, fused_queued<4, 64, 1, 32, false, true>, 1);
    }
    // ---- B second shape, alone and pipelined behind the last block's sweeps on a high-priority stream ------------------------------
    if (nb == 2 || nb == 3) {
        auto launch_b2 = [&](cudaStream_t st, uint64_t g0, uint64_t g1) {
            const unsigned grid = (unsigned)(((g1 - g0) * 32 + 127) / 128);
            if (nb == 2) phase_b2<2><<<grid, 128, 0, st>>>(ba, seg_row, state, out, outc, seg_sum, seg_cnt, nseg, eps, g0, g1);
            else phase_b2<3><<<grid, 128, 0, st>>>(ba, seg_row, state, out, outc, seg_sum, seg_cnt, nseg, eps, g0, g1);
        };
        auto launch_a = [&](cudaStream_t st, uint32_t k, uint64_t g0, uint64_t g1) {
            phase_a<8, 2, 24, 0><<<(unsigned)((g1 - g0 + 1) / 2), 64, 0, st>>>(bsrc + hbase[k], off + (uint64_t)k * pad, seg_row, key, key + (uint64_t)k * bsize, state + (uint64_t)k * bsize,
                                                                            cand + hbase[k], cval + hbase[k], gcnt + (uint64_t)k * ngroups, nseg, band, g0, g1);
        };
        {
            float best_a = 1e9f, best_b = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaMemset(out, 0, n * 8)); CK(cudaMemset(outc, 0, n * 4));
                CK(cudaEventRecord(e0));
                for (uint32_t k = 0; k < nb; ++k) launch_a(0, k, 0, ngroups);
                CK(cudaEventRecord(e1));
                launch_b2(0, 0, ngroups);
                if (nh) merge_hubs<<<(nh + 255) / 256, 256>>>(hub_rows, nh, sfirst, seg_sum, seg_cnt, out, outc);
                CK(cudaEventRecord(e2)); CK(cudaEventSynchronize(e2)); CK(cudaGetLastError());
                float a, b; CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, e1, e2));
                if (a + b < best_a + best_b) { best_a = a; best_b = b; }
            }
            check("A: U=8 2w 24; B2 (runs by search), serial", best_a, best_b);
        }
        int lo_p = 0, hi_p = 0; cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
        cudaStream_t sA, sB; CK(cudaStreamCreateWithPriority(&sA, cudaStreamNonBlocking, lo_p)); CK(cudaStreamCreateWithPriority(&sB, cudaStreamNonBlocking, hi_p));
        for (int order = 0; order < (getenv("SWEEP") ? 2 : 0); ++order)
            for (uint32_t S : {4u, 8u, 16u}) {
                std::vector<cudaEvent_t> ev(S); for (auto& x : ev) CK(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
                cudaEvent_t evB; CK(cudaEventCreateWithFlags(&evB, cudaEventDisableTiming));
                float best = 1e9f;
                for (int rep = 0; rep < 3; ++rep) {
                    CK(cudaMemset(out, 0, n * 8)); CK(cudaMemset(outc, 0, n * 4)); CK(cudaDeviceSynchronize());
                    CK(cudaEventRecord(e0, sA));
                    if (order == 0) for (uint32_t k = 0; k + 1 < nb; ++k) launch_a(sA, k, 0, ngroups);      // block-major: slices only in the last block's sweep
                    for (uint32_t sl = 0; sl < S; ++sl) {
                        const uint64_t g0 = ngroups * sl / S, g1 = ngroups * (sl + 1) / S;
                        if (order == 1) for (uint32_t k = 0; k + 1 < nb; ++k) launch_a(sA, k, g0, g1);         // slice-major
                        launch_a(sA, nb - 1, g0, g1);
                        CK(cudaEventRecord(ev[sl], sA));
                        CK(cudaStreamWaitEvent(sB, ev[sl], 0));
                        launch_b2(sB, g0, g1);
                    }
                    if (nh) merge_hubs<<<(nh + 255) / 256, 256, 0, sB>>>(hub_rows, nh, sfirst, seg_sum, seg_cnt, out, outc);
                    CK(cudaEventRecord(evB, sB)); CK(cudaStreamWaitEvent(sA, evB, 0));
                    CK(cudaEventRecord(e1, sA)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
                    float t; CK(cudaEventElapsedTime(&t, e0, e1)); if (t < best) best = t;
                }
                char nm[96]; snprintf(nm, sizeof nm, "pipelined, %s, %u slices (A+B2 total in A)", order == 0 ? "block-major" : "slice-major", S);
                check(nm, best, 0.f);
            }
    }
    return 0;
}
