// primitives.cuh — device building blocks of finish_write! and the count->scan->emit write phase:
// exclusive scan, stable LSD radix sort (8-bit digits) with payload, CSR offset construction,
// AoS<->SoA transposition, stream compaction of flags.  All hand-written for sm_100a; every kernel is
// HBM-bound (no tensor-core work on this path).
//
// Replaces, on device, what the reference does with Dict/Vector containers on the host:
//   per-target push! order        src/EdgeMethods.jl:495-496,518-519   -> stable sort on target row
//   Dict{AgentID,Vector} lookup   src/EdgeMethods.jl:303-371            -> CSR offsets
//   reuseable slot list           src/AgentMethods.jl:171,430           -> ordered compaction of died flags
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace vbp {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t x = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += x;
    }
    return v;
}
// exclusive scan of one value per thread across a 256-thread block; returns the block total in `total`
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t& total) {
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t inc = warp_incl_scan(v, lane);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < SCAN_THREADS / 32 ? wsum[lane] : 0;
        s = warp_incl_scan(s, lane);
        if (lane < SCAN_THREADS / 32) wsum[lane] = s;
    }
    __syncthreads();
    const uint32_t woff = w ? wsum[w - 1] : 0;
    total = wsum[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return woff + inc - v;
}

static __global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ bsum) {
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const uint64_t k = base + (uint64_t)i * SCAN_THREADS + threadIdx.x;
        if (k < n) s += in[k];
    }
    uint32_t total;
    block_excl_scan(s, total);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}
// in-place exclusive scan of up to SCAN_TILE values by one block; writes the grand total to *total (if non-null)
static __global__ void __launch_bounds__(SCAN_THREADS) scan_single_kernel(uint32_t* __restrict__ data, uint32_t n, uint32_t* __restrict__ total_out) {
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const uint32_t k = threadIdx.x * SCAN_ITEMS + i;
        v[i] = k < n ? data[k] : 0;
        s += v[i];
    }
    uint32_t total;
    uint32_t ex = block_excl_scan(s, total);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const uint32_t k = threadIdx.x * SCAN_ITEMS + i;
        if (k < n) data[k] = ex;
        ex += v[i];
    }
    if (total_out && threadIdx.x == 0) *total_out = total;
}
static __global__ void __launch_bounds__(SCAN_THREADS) scan_downsweep_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t n,
                                                                    const uint32_t* __restrict__ boff) {
    // thread t owns items [t*ITEMS, t*ITEMS+ITEMS) of the tile (blocked), so the per-thread scan is sequential
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = (base + i < n) ? in[base + i] : 0; s += v[i]; }
    uint32_t total;
    uint32_t ex = block_excl_scan(s, total) + boff[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
}
static __global__ void scan_total_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ out, uint64_t n, uint32_t* total) {
    *total = n ? out[n - 1] + in[n - 1] : 0;
}

// scratch needed by exclusive_scan for n elements (in uint32 words)
inline uint64_t scan_scratch_words(uint64_t n) {
    uint64_t words = 0;
    while (n > SCAN_TILE) { n = (n + SCAN_TILE - 1) / SCAN_TILE; words += n; }
    return words + 1;
}
// out[i] = sum(in[0..i)); in and out may alias only if identical.  `total` (device, optional) receives the sum.
inline void exclusive_scan(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* total, uint32_t* scratch, cudaStream_t st) {
    if (n == 0) { if (total) cudaMemsetAsync(total, 0, 4, st); return; }
    if (n <= SCAN_TILE) {
        if (in != out) cudaMemcpyAsync(out, in, n * 4, cudaMemcpyDeviceToDevice, st);
        scan_single_kernel<<<1, SCAN_THREADS, 0, st>>>(out, (uint32_t)n, total);
        return;
    }
    const uint64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_reduce_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, scratch);
    exclusive_scan(scratch, scratch, nb, total, scratch + nb, st);
    scan_downsweep_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, n, scratch);
}

// ---- stable LSD radix sort, 8-bit digits, u32 keys, up to two payload arrays (4 B and 4/8 B) -----------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;                          // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;       // 2048 keys per block
constexpr int RS_SEG = RS_TILE / RS_WARPS;           // contiguous keys per warp

static __global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint32_t* __restrict__ keys, uint64_t n, int shift, uint32_t* __restrict__ hist,
                                                                  uint32_t nblocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const uint64_t k = base + (uint64_t)i * RS_THREADS + threadIdx.x;
        if (k < n) atomicAdd(&h[(__ldcs(keys + k) >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// lanes of the warp holding the same 9-bit digit value (bit 8 = "no key"): nine ballots instead of match.any, whose cost on
// sm_100 grows with the number of distinct values in the warp (profiles: the scatter pass was issue-bound at 84 % SM throughput)
__device__ __forceinline__ uint32_t match_digit(uint32_t d) {
#ifdef VB_SORT_MATCH_ANY
    return __match_any_sync(0xffffffffu, d);
#else
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 9; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
        m &= bit ? bal : ~bal;
    }
    return m;
#endif
}
struct NoPayload { uint8_t x; };
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) { return __ldcs(p); }
__device__ __forceinline__ uint64_t ld_stream(const uint64_t* p) { return (uint64_t)__ldcs(reinterpret_cast<const unsigned long long*>(p)); }
__device__ __forceinline__ NoPayload ld_stream(const NoPayload* p) { return *p; }

template <class P1, class P2>   // payload word types; NoPayload for absent
struct RsArgs {
    const uint32_t* kin; uint32_t* kout;
    const P1* p1in; P1* p1out;
    const P2* p2in; P2* p2out;
    uint64_t n; int shift; const uint32_t* hist; uint32_t nblocks;
};
// One pass of the stable LSD sort for one tile: all loads are issued up front, ranks come from warp-private digit
// counters (match.any), key and payloads are staged together in shared memory in tile-sorted order and written out
// as coalesced digit runs.  4 blocks per SM (<= 64 registers, 45 KB smem).
template <class P1, class P2, bool HAS1, bool HAS2>
static __global__ void __launch_bounds__(RS_THREADS, 4) rs_scatter_kernel(const RsArgs<P1, P2> a) {
    __shared__ uint32_t wh[RS_WARPS][256];
    __shared__ uint32_t dstart[256];
    __shared__ uint32_t gbase[256];
    __shared__ uint32_t skey[RS_TILE];
    __shared__ P1 sp1[HAS1 ? RS_TILE : 1];
    __shared__ P2 sp2[HAS2 ? RS_TILE : 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wh[0][0])[i] = 0;
    const uint64_t tile0 = (uint64_t)blockIdx.x * RS_TILE;
    const uint32_t count = (uint32_t)((a.n - tile0) < (uint64_t)RS_TILE ? (a.n - tile0) : (uint64_t)RS_TILE);
    uint32_t key[RS_ITEMS];
    P1 v1[HAS1 ? RS_ITEMS : 1];
    P2 v2[HAS2 ? RS_ITEMS : 1];
    // warp w owns the contiguous segment [w*RS_SEG, (w+1)*RS_SEG) of the tile; round r covers 32 consecutive keys
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t li = w * RS_SEG + r * 32 + lane;
        const bool valid = li < count;
        key[r] = valid ? __ldcs(a.kin + tile0 + li) : 0xffffffffu;
        if (HAS1) v1[r] = valid ? ld_stream(a.p1in + tile0 + li) : P1();
        if (HAS2) v2[r] = valid ? ld_stream(a.p2in + tile0 + li) : P2();
    }
    __syncthreads();
    // phase 1: stable rank of every key among the keys of its digit inside the warp segment
    uint32_t pos[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t li = w * RS_SEG + r * 32 + lane;
        const bool valid = li < count;
        const uint32_t d = valid ? ((key[r] >> a.shift) & 255u) : 256u;
        const uint32_t m = match_digit(d);
        const int leader = __ffs(m) - 1;
        uint32_t prev = 0;
        if (valid && lane == leader) prev = wh[w][d];
        prev = __shfl_sync(0xffffffffu, prev, leader);
        if (valid && lane == leader) wh[w][d] = prev + __popc(m);
        __syncwarp();
        pos[r] = prev + __popc(m & ((1u << lane) - 1u));
    }
    __syncthreads();
    // phase 2: exclusive prefix over warps per digit, digit starts inside the tile, global bases
    {
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) { const uint32_t t = wh[ww][d]; wh[ww][d] = run; run += t; }
        uint32_t total;
        const uint32_t ex = block_excl_scan(run, total);
        dstart[d] = ex;
        gbase[d] = a.hist[(uint64_t)d * a.nblocks + blockIdx.x];
    }
    __syncthreads();
    // phase 3: stage key + payloads in tile-sorted order
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t li = w * RS_SEG + r * 32 + lane;
        if (li < count) {
            const uint32_t d = (key[r] >> a.shift) & 255u;
            const uint32_t p = pos[r] + dstart[d] + wh[w][d];
            skey[p] = key[r];
            if (HAS1) sp1[p] = v1[r];
            if (HAS2) sp2[p] = v2[r];
        }
    }
    __syncthreads();
    // phase 4: element i of the sorted tile goes to gbase[digit] + (i - dstart[digit]): consecutive threads write consecutive addresses
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = r * RS_THREADS + threadIdx.x;
        if (i < count) {
            const uint32_t k = skey[i];
            const uint32_t d = (k >> a.shift) & 255u;
            const uint32_t dst = gbase[d] + (i - dstart[d]);
            a.kout[dst] = k;
            if (HAS1) a.p1out[dst] = sp1[i];
            if (HAS2) a.p2out[dst] = sp2[i];
        }
    }
}

// ---- onesweep: one kernel per digit pass (Adinets & Merrill).  A first kernel reads the keys once and builds the global
//      digit histograms of all passes; each pass kernel then takes tiles in ticket order, publishes its per-digit counts in a
//      status array and finds its exclusive per-digit prefix by decoupled look-back over the preceding tiles. -------------------
constexpr int OS_MAX_PASSES = 4;
constexpr unsigned long long OS_FLAG_AGG = 1ull << 62, OS_FLAG_PREFIX = 2ull << 62, OS_VALUE_MASK = (1ull << 62) - 1;

static __global__ void __launch_bounds__(RS_THREADS) os_hist_kernel(const uint32_t* __restrict__ keys, uint64_t n, int npass, uint32_t* __restrict__ ghist) {
    __shared__ uint32_t h[OS_MAX_PASSES][256];
    for (int i = threadIdx.x; i < OS_MAX_PASSES * 256; i += RS_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    for (uint64_t k = (uint64_t)blockIdx.x * RS_THREADS + threadIdx.x; k < n; k += (uint64_t)gridDim.x * RS_THREADS) {
        const uint32_t key = __ldcs(keys + k);
#pragma unroll
        for (int p = 0; p < OS_MAX_PASSES; ++p) if (p < npass) atomicAdd(&h[p][(key >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * 256; i += RS_THREADS) { const uint32_t v = (&h[0][0])[i]; if (v) atomicAdd(&ghist[i], v); }
}
// exclusive scan of each pass's 256 digit counts, in place (one block of 256 threads per pass)
static __global__ void __launch_bounds__(256) os_scan_kernel(uint32_t* __restrict__ ghist) {
    uint32_t total;
    const uint32_t v = ghist[blockIdx.x * 256 + threadIdx.x];
    const uint32_t ex = block_excl_scan(v, total);
    ghist[blockIdx.x * 256 + threadIdx.x] = ex;
}

template <class P1, class P2>
struct OsArgs {
    const uint32_t* kin; uint32_t* kout;
    const P1* p1in; P1* p1out;
    const P2* p2in; P2* p2out;
    uint64_t n; int shift; const uint32_t* gbase;   // gbase: exclusive digit bases of this pass [256]
    unsigned long long* status;                     // [ntiles][256], zeroed before the pass
    uint32_t* ticket;                               // zeroed before the pass
    uint32_t* error;                                // set if a look-back spin gives up (never expected)
};

template <class P1, class P2, bool HAS1, bool HAS2>
static __global__ void __launch_bounds__(RS_THREADS, 4) os_pass_kernel(const OsArgs<P1, P2> a) {
    __shared__ uint32_t wh[RS_WARPS][256];
    __shared__ uint32_t dstart[256];
    __shared__ uint32_t gbase[256];
    __shared__ uint32_t skey[RS_TILE];
    __shared__ P1 sp1[HAS1 ? RS_TILE : 1];
    __shared__ P2 sp2[HAS2 ? RS_TILE : 1];
    __shared__ uint32_t s_tile;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1u);   // tiles are taken in launch order: a predecessor is always running or done
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wh[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile0 = (uint64_t)tile * RS_TILE;
    const uint32_t count = (uint32_t)((a.n - tile0) < (uint64_t)RS_TILE ? (a.n - tile0) : (uint64_t)RS_TILE);
    uint32_t key[RS_ITEMS];
    P1 v1[HAS1 ? RS_ITEMS : 1];
    P2 v2[HAS2 ? RS_ITEMS : 1];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t li = w * RS_SEG + r * 32 + lane;
        const bool valid = li < count;
        key[r] = valid ? __ldcs(a.kin + tile0 + li) : 0xffffffffu;
        if (HAS1) v1[r] = valid ? ld_stream(a.p1in + tile0 + li) : P1();
        if (HAS2) v2[r] = valid ? ld_stream(a.p2in + tile0 + li) : P2();
    }
    // stable rank of every key among the keys of its digit inside the warp segment
    uint32_t pos[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t li = w * RS_SEG + r * 32 + lane;
        const bool valid = li < count;
        const uint32_t d = valid ? ((key[r] >> a.shift) & 255u) : 256u;
        const uint32_t m = match_digit(d);
        const int leader = __ffs(m) - 1;
        uint32_t prev = 0;
        if (valid && lane == leader) prev = wh[w][d];
        prev = __shfl_sync(0xffffffffu, prev, leader);
        if (valid && lane == leader) wh[w][d] = prev + __popc(m);
        __syncwarp();
        pos[r] = prev + __popc(m & ((1u << lane) - 1u));
    }
    __syncthreads();
    // per digit: prefix over the warps, publish the tile's count, look back for the exclusive prefix over earlier tiles
    {
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) { const uint32_t t = wh[ww][d]; wh[ww][d] = run; run += t; }
        unsigned long long* mine = a.status + (size_t)tile * 256 + d;
        if (tile == 0) {
            __stcg(mine, OS_FLAG_PREFIX | run);
            gbase[d] = a.gbase[d];
        } else {
            __stcg(mine, OS_FLAG_AGG | run);
            unsigned long long excl = 0;
            long long t = (long long)tile - 1;
            bool done = false;
            while (t >= 0 && !done) {
                // four predecessors per round trip: their status words are loaded together, then consumed in order
                unsigned long long v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = (t - j >= 0) ? __ldcg(a.status + (size_t)(t - j) * 256 + d) : 0ull;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (done || t - j < 0) break;
                    unsigned long long x = v[j];
                    unsigned spins = 0;
                    while ((x >> 62) == 0) {   // not published yet: wait for this one
                        if (++spins > (1u << 26)) { atomicExch(a.error, 1u); break; }
                        __nanosleep(20);
                        x = __ldcg(a.status + (size_t)(t - j) * 256 + d);
                    }
                    excl += x & OS_VALUE_MASK;
                    if ((x >> 62) != 1) done = true;   // an inclusive prefix (or the give-up path) ends the walk
                }
                t -= 4;
            }
            __stcg(mine, OS_FLAG_PREFIX | (excl + run));
            gbase[d] = a.gbase[d] + (uint32_t)excl;
        }
        uint32_t total;
        dstart[d] = block_excl_scan(run, total);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t li = w * RS_SEG + r * 32 + lane;
        if (li < count) {
            const uint32_t d = (key[r] >> a.shift) & 255u;
            const uint32_t p = pos[r] + dstart[d] + wh[w][d];
            skey[p] = key[r];
            if (HAS1) sp1[p] = v1[r];
            if (HAS2) sp2[p] = v2[r];
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = r * RS_THREADS + threadIdx.x;
        if (i < count) {
            const uint32_t k = skey[i];
            const uint32_t d = (k >> a.shift) & 255u;
            const uint32_t dst = gbase[d] + (i - dstart[d]);
            a.kout[dst] = k;
            if (HAS1) a.p1out[dst] = sp1[i];
            if (HAS2) a.p2out[dst] = sp2[i];
        }
    }
}

inline uint64_t rs_scratch_words(uint64_t n) {
    const uint64_t nb = (n + RS_TILE - 1) / RS_TILE;
    const uint64_t classic = nb * 256 + scan_scratch_words(nb * 256) + 16;
    const uint64_t onesweep = nb * 256 * 2 + OS_MAX_PASSES * 256 + 64;
    return classic > onesweep ? classic : onesweep;
}
inline int bits_for(uint64_t maxkey_plus1) {
    int b = 1;
    while (b < 32 && (1ull << b) < maxkey_plus1) ++b;
    return b;
}

template <class P1, class P2, bool HAS1, bool HAS2>
inline void os_launch(const OsArgs<P1, P2>& a, uint32_t nb, cudaStream_t st) { os_pass_kernel<P1, P2, HAS1, HAS2><<<nb, RS_THREADS, 0, st>>>(a); }

// One sort = ceil(bits/8) passes ping-ponging between (k0,p10,p20) and (k1,p11,p21); returns which buffer set holds the
// result (0 or 1).  p2 word size: 0 (none), 4 or 8 bytes.  p1 (4 B) may be null.  `error` (device word, optional) is set if a
// look-back ever gives up.  Default: three kernels per pass (tile histogram, scan, scatter) — measured faster on B200 at the
// 2048-key tiles this kernel uses (SIR config 5, 1e8 edges: 4.56 ms vs 5.06 ms per finish_write!); VB_ONESWEEP=1 selects the
// single-kernel-per-pass onesweep variant (decoupled look-back).
inline int radix_sort(uint32_t* k0, uint32_t* k1, uint32_t* p10, uint32_t* p11, void* p20, void* p21, int p2_bytes, uint64_t n, int bits,
                      uint32_t* scratch, cudaStream_t st, uint32_t* error = nullptr) {
    if (n == 0) return 0;
    const uint32_t nb = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    static const bool classic = getenv("VB_ONESWEEP") == nullptr;
    const int npass = (bits + 7) / 8;
    int cur = 0;
    if (!classic && npass <= OS_MAX_PASSES) {
        // scratch layout: [status: nb*256 u64][ghist: 4*256 u32][ticket][error]
        unsigned long long* status = reinterpret_cast<unsigned long long*>(scratch);
        uint32_t* ghist = scratch + (uint64_t)nb * 256 * 2;
        uint32_t* ticket = ghist + OS_MAX_PASSES * 256;
        uint32_t* err = error ? error : ticket + 1;
        cudaMemsetAsync(ghist, 0, (OS_MAX_PASSES * 256 + 8) * 4, st);
        const unsigned hb = (unsigned)((n + RS_THREADS * 16 - 1) / (RS_THREADS * 16));
        os_hist_kernel<<<hb > 4096 ? 4096 : (hb ? hb : 1), RS_THREADS, 0, st>>>(k0, n, npass, ghist);
        os_scan_kernel<<<npass, 256, 0, st>>>(ghist);
        for (int p = 0; p < npass; ++p) {
            uint32_t* kin = cur ? k1 : k0; uint32_t* kout = cur ? k0 : k1;
            uint32_t* p1in = cur ? p11 : p10; uint32_t* p1out = cur ? p10 : p11;
            void* p2in = cur ? p21 : p20; void* p2out = cur ? p20 : p21;
            cudaMemsetAsync(status, 0, (size_t)nb * 256 * 8, st);
            cudaMemsetAsync(ticket, 0, 4, st);
            const bool h1 = p10 != nullptr;
            if (p2_bytes == 8) {
                OsArgs<uint32_t, uint64_t> a{kin, kout, p1in, p1out, (const uint64_t*)p2in, (uint64_t*)p2out, n, 8 * p, ghist + p * 256, status, ticket, err};
                if (h1) os_launch<uint32_t, uint64_t, true, true>(a, nb, st); else os_launch<uint32_t, uint64_t, false, true>(a, nb, st);
            } else if (p2_bytes == 4) {
                OsArgs<uint32_t, uint32_t> a{kin, kout, p1in, p1out, (const uint32_t*)p2in, (uint32_t*)p2out, n, 8 * p, ghist + p * 256, status, ticket, err};
                if (h1) os_launch<uint32_t, uint32_t, true, true>(a, nb, st); else os_launch<uint32_t, uint32_t, false, true>(a, nb, st);
            } else {
                OsArgs<uint32_t, NoPayload> a{kin, kout, p1in, p1out, nullptr, nullptr, n, 8 * p, ghist + p * 256, status, ticket, err};
                if (h1) os_launch<uint32_t, NoPayload, true, false>(a, nb, st); else os_launch<uint32_t, NoPayload, false, false>(a, nb, st);
            }
            cur ^= 1;
        }
        return cur;
    }
    uint32_t* hist = scratch;
    uint32_t* sscr = scratch + (uint64_t)nb * 256;
    for (int shift = 0; shift < bits; shift += 8) {
        uint32_t* kin = cur ? k1 : k0; uint32_t* kout = cur ? k0 : k1;
        uint32_t* p1in = cur ? p11 : p10; uint32_t* p1out = cur ? p10 : p11;
        void* p2in = cur ? p21 : p20; void* p2out = cur ? p20 : p21;
        rs_hist_kernel<<<nb, RS_THREADS, 0, st>>>(kin, n, shift, hist, nb);
        exclusive_scan(hist, hist, (uint64_t)nb * 256, nullptr, sscr, st);
        const bool h1 = p10 != nullptr;
        if (p2_bytes == 8) {
            RsArgs<uint32_t, uint64_t> a{kin, kout, p1in, p1out, (const uint64_t*)p2in, (uint64_t*)p2out, n, shift, hist, nb};
            if (h1) rs_scatter_kernel<uint32_t, uint64_t, true, true><<<nb, RS_THREADS, 0, st>>>(a);
            else rs_scatter_kernel<uint32_t, uint64_t, false, true><<<nb, RS_THREADS, 0, st>>>(a);
        } else if (p2_bytes == 4) {
            RsArgs<uint32_t, uint32_t> a{kin, kout, p1in, p1out, (const uint32_t*)p2in, (uint32_t*)p2out, n, shift, hist, nb};
            if (h1) rs_scatter_kernel<uint32_t, uint32_t, true, true><<<nb, RS_THREADS, 0, st>>>(a);
            else rs_scatter_kernel<uint32_t, uint32_t, false, true><<<nb, RS_THREADS, 0, st>>>(a);
        } else {
            RsArgs<uint32_t, NoPayload> a{kin, kout, p1in, p1out, nullptr, nullptr, n, shift, hist, nb};
            if (h1) rs_scatter_kernel<uint32_t, NoPayload, true, false><<<nb, RS_THREADS, 0, st>>>(a);
            else rs_scatter_kernel<uint32_t, NoPayload, false, false><<<nb, RS_THREADS, 0, st>>>(a);
        }
        cur ^= 1;
    }
    return cur;
}

// ---- CSR row counts from sorted row keys: cnt[r] += (#entries with key r).  Two atomics per non-empty row
//      (run start subtracts its index, run end adds index+1), no per-entry atomics; an exclusive scan of cnt
//      gives the offsets.  cnt must be zero (or hold the counts of an existing CSR being merged). ------------
static __global__ void csr_run_counts_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ cnt) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = keys[i];
    if (i == 0 || keys[i - 1] != k) atomicSub(&cnt[k], (uint32_t)i);
    if (i == n - 1 || keys[i + 1] != k) atomicAdd(&cnt[k], (uint32_t)i + 1u);
}
// ---- CSR offsets straight from the sorted keys: off[r] = first position whose key is >= r (off[rows] = n).  Thread i looks at the
//      boundary between keys[i - 1] and keys[i] (i = 0: before the first key, i = n: behind the last one) and writes the offsets of the
//      rows that start there.  No atomics, no counts, no scan over the rows.  A boundary that skips more than 32 empty rows is parked
//      in `gaps` ((first row, last row, value) triples, at most rows / 32 + 2 of them) and filled by csr_offset_gaps_kernel. ------------
static __global__ void csr_offsets_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t rows, uint32_t* __restrict__ off,
                                          uint32_t* __restrict__ gaps, uint32_t* __restrict__ gap_count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const uint32_t cur = i < n ? __ldcs(keys + i) : rows;
    uint32_t lo = 0;
    if (i) { const uint32_t prev = __ldcs(keys + i - 1); if (prev == cur) return; lo = prev + 1; }
    if (cur - lo < 32u) { for (uint32_t r = lo; r <= cur; ++r) off[r] = (uint32_t)i; return; }
    const uint32_t g = atomicAdd(gap_count, 1u);
    gaps[3 * g] = lo; gaps[3 * g + 1] = cur; gaps[3 * g + 2] = (uint32_t)i;
}
static __global__ void csr_offset_gaps_kernel(uint32_t* __restrict__ off, const uint32_t* __restrict__ gaps, const uint32_t* __restrict__ gap_count) {
    const uint32_t ng = *gap_count;
    for (uint32_t g = blockIdx.x; g < ng; g += gridDim.x) {
        const uint32_t lo = gaps[3 * g], hi = gaps[3 * g + 1], v = gaps[3 * g + 2];
        for (uint64_t r = (uint64_t)lo + threadIdx.x; r <= hi; r += blockDim.x) off[r] = v;
    }
}
static __global__ void fill_u32_kernel(uint32_t* p, uint64_t n, uint32_t v) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
static __global__ void iota_u32_kernel(uint32_t* p, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}
// row lengths -> counts (for merging an existing CSR with newly sorted edges)
static __global__ void row_counts_kernel(const uint32_t* __restrict__ off, uint32_t rows, uint32_t* __restrict__ cnt) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) cnt[r] = off[r + 1] - off[r];
}

// ---- AoS (host records) <-> SoA word columns ------------------------------------------------------------------
// dst column c of record slot0+i at cols + c*stride*word + (slot0+i)*word
static __global__ void aos_to_soa_kernel(const uint8_t* __restrict__ aos, uint8_t* __restrict__ cols, uint64_t stride, uint64_t slot0, uint64_t n,
                                  uint32_t size, uint32_t word) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ncols = size / word;
    if (t >= n * ncols) return;
    const uint64_t i = t / ncols;
    const uint32_t c = (uint32_t)(t % ncols);
    const uint8_t* s = aos + i * size + (uint64_t)c * word;
    uint8_t* d = cols + (uint64_t)c * stride * word + (slot0 + i) * word;
    for (uint32_t b = 0; b < word; ++b) d[b] = s[b];
}
static __global__ void soa_to_aos_kernel(const uint8_t* __restrict__ cols, uint8_t* __restrict__ aos, uint64_t stride, uint64_t slot0, uint64_t n,
                                  uint32_t size, uint32_t word) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ncols = size / word;
    if (t >= n * ncols) return;
    const uint64_t i = t / ncols;
    const uint32_t c = (uint32_t)(t % ncols);
    uint8_t* d = aos + i * size + (uint64_t)c * word;
    const uint8_t* s = cols + (uint64_t)c * stride * word + (slot0 + i) * word;
    for (uint32_t b = 0; b < word; ++b) d[b] = s[b];
}
// gather records through an index list (compacted read-out): aos[i] = record idx[i]
static __global__ void soa_gather_aos_kernel(const uint8_t* __restrict__ cols, uint8_t* __restrict__ aos, uint64_t stride, const uint32_t* __restrict__ idx,
                                      uint64_t n, uint32_t size, uint32_t word) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ncols = size / word;
    if (t >= n * ncols) return;
    const uint64_t i = t / ncols;
    const uint32_t c = (uint32_t)(t % ncols);
    uint8_t* d = aos + i * size + (uint64_t)c * word;
    const uint8_t* s = cols + (uint64_t)c * stride * word + (uint64_t)idx[i] * word;
    for (uint32_t b = 0; b < word; ++b) d[b] = s[b];
}
// column-wise copy between two SoA buffers with different strides (capacity growth, log -> CSR moves)
static __global__ void soa_copy_kernel(const uint8_t* __restrict__ src, uint64_t sstride, uint8_t* __restrict__ dst, uint64_t dstride, uint64_t n,
                                uint32_t ncols, uint32_t word, uint64_t soff, uint64_t doff) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per = n * word;
    if (t >= per * ncols) return;
    const uint32_t c = (uint32_t)(t / per);
    const uint64_t b = t % per;
    dst[(uint64_t)c * dstride * word + doff * word + b] = src[(uint64_t)c * sstride * word + soff * word + b];
}

// ---- flags -> ordered index list (newly died slots appended to the reuse stack in ascending order) ------------
// flag[i] = died_w[i] && !died_r[i] for i < n_r (slots beyond the read length cannot die this step)
static __global__ void newly_died_flags_kernel(const uint8_t* __restrict__ died_r, const uint8_t* __restrict__ died_w, uint32_t n, uint32_t* __restrict__ flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (died_w[i] && !died_r[i]) ? 1u : 0u;
}
// how many slots died in this apply (same predicate), without materialising the flags: one 16-byte load of each array per 16 slots.
// finish_write! of a mortal type whose transition killed nobody (the common case) then skips the flags + scan + compact passes
// (1.8 GB of traffic at 1e8 slots against 0.2 GB for this count).  Both arrays are 16-byte aligned (pool allocations).
static __device__ __forceinline__ uint32_t nonzero_bytes(uint32_t x) { return (x | ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u; }
static __global__ void __launch_bounds__(256) count_newly_died_kernel(const uint8_t* __restrict__ died_r, const uint8_t* __restrict__ died_w, uint32_t n,
                                                                      uint32_t* __restrict__ count) {
    const uint32_t nvec = n / 16;
    const uint4* r4 = reinterpret_cast<const uint4*>(died_r);
    const uint4* w4 = reinterpret_cast<const uint4*>(died_w);
    uint32_t c = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 r = r4[i], w = w4[i];
        c += __popc(nonzero_bytes(w.x) & ~nonzero_bytes(r.x)) + __popc(nonzero_bytes(w.y) & ~nonzero_bytes(r.y)) +
             __popc(nonzero_bytes(w.z) & ~nonzero_bytes(r.z)) + __popc(nonzero_bytes(w.w) & ~nonzero_bytes(r.w));
    }
    if (blockIdx.x == 0 && threadIdx.x < n - nvec * 16) {      // tail of fewer than 16 slots
        const uint32_t i = nvec * 16 + threadIdx.x;
        c += (died_w[i] && !died_r[i]) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}
static __global__ void alive_flags_kernel(const uint8_t* __restrict__ died, uint32_t n, uint32_t* __restrict__ flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = died ? (died[i] ? 0u : 1u) : 1u;
}
static __global__ void compact_indices_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos, uint32_t n, uint32_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) out[pos[i]] = (uint32_t)i;
}

inline unsigned nblk(uint64_t n, unsigned t = 256) { return (unsigned)((n + t - 1) / t); }

}  // namespace vbp
