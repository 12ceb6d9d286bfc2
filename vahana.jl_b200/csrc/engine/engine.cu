// engine.cu — the B200 execution engine behind include/vahana_b200.h.
//
// Host-side phase logic of apply! (src/Simulation.jl:720-821) and finish_write!
// (src/AgentMethods.jl:300-439, src/EdgeMethods.jl:639-684) driving hand-written sm_100a kernels:
//   read phase / write phase   transition_kernel<F,...> (include/vahana_device.cuh), count -> scan -> emit
//   finish_write! (edges)      stable radix sort on target row + CSR build / merge (primitives.cuh)
//   finish_write! (agents)     buffer swap, died flags, ordered reuse-stack append, dead-agent edge purge
//   mapreduce                  block reduction kernels
// There is no CPU fallback: every state-touching entry point needs the CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and enums only: the library is resolved at run time (dlopen), see Nccl below

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/vahana_b200.h"
#include "../../../include/vahana_device.cuh"
#include "primitives.cuh"

using vb::AgentID;
using vbp::nblk;

// ------------------------------------------------------------------------------------------------------
namespace {

struct AssertionError : std::runtime_error { explicit AssertionError(const std::string& m) : std::runtime_error(m) {} };
struct ArgError : std::runtime_error { explicit ArgError(const std::string& m) : std::runtime_error(m) {} };
struct CudaError : std::runtime_error { explicit CudaError(const std::string& m) : std::runtime_error(m) {} };

thread_local std::string g_err;
int g_device = -1;
cudaStream_t g_stream = nullptr;
unsigned long long g_launches = 0;   // kernels of ours launched (bench reports the per-step count)

#define CK(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) throw CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e));     \
    } while (0)
#define LAUNCH_CHECK()                 \
    do {                               \
        ++g_launches;                  \
        CK(cudaGetLastError());        \
    } while (0)

// ---- NCCL over NVLink: replaces the reference's MPI exchanges (src/MPI.jl, src/MPIinit.jl) -----------------
// Resolved with dlopen so the engine has no link-time dependency: inside a torch process the bundled
// libnccl.so.2 is already mapped; other hosts set VB_NCCL_LIB or have libnccl.so.2 on the loader path.
struct Nccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, int /*ncclRedOp_t: 0 = sum*/, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (h) return true;
        const char* env = getenv("VB_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { if (n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break; }
        if (!h) return false;
#define VB_SYM(field, name) field = (decltype(field))dlsym(h, name); if (!field) return false;
        VB_SYM(GetUniqueId, "ncclGetUniqueId") VB_SYM(CommInitRank, "ncclCommInitRank") VB_SYM(CommDestroy, "ncclCommDestroy")
        VB_SYM(Send, "ncclSend") VB_SYM(Recv, "ncclRecv") VB_SYM(GroupStart, "ncclGroupStart") VB_SYM(GroupEnd, "ncclGroupEnd")
        VB_SYM(AllGather, "ncclAllGather") VB_SYM(AllReduce, "ncclAllReduce") VB_SYM(GetErrorString, "ncclGetErrorString")
#undef VB_SYM
        return true;
    }
};
Nccl g_nccl;
ncclComm_t g_comm = nullptr;
int g_rank = 0, g_nranks = 1;
#define NK(expr)                                                                                                  \
    do {                                                                                                          \
        ncclResult_t _r = (expr);                                                                                 \
        if (_r != ncclSuccess) throw CudaError(std::string(#expr) + ": " + g_nccl.GetErrorString(_r));           \
    } while (0)

void allgather8_host(const void* mine, std::vector<uint64_t>& all);

// VB_TRACE=1: per-phase host wall times of apply! on stderr (the analogue of the reference's <Begin>/<End> duration log, src/Logging.jl:30-73)
struct Trace {
    bool on = getenv("VB_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0;
    void begin() { if (on) { cudaStreamSynchronize(g_stream); t0 = std::chrono::steady_clock::now(); } }
    void end(const char* what, const std::string& detail = "") {
        if (!on) return;
        cudaStreamSynchronize(g_stream);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[vb trace] %-28s %-24s %9.3f ms\n", what, detail.c_str(), ms);
        t0 = std::chrono::steady_clock::now();
    }
    // VB_TRACE=2: device times between marks on the main stream (CUDA events), printed by flush_marks() after a synchronize
    bool marks_on = getenv("VB_TRACE") != nullptr && atoi(getenv("VB_TRACE")) >= 2;
    std::vector<std::pair<std::string, cudaEvent_t>> marks, side_marks;
    void mark(const std::string& what, cudaStream_t side = nullptr) {
        if (!marks_on) return;
        cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, side ? side : g_stream);
        (side ? side_marks : marks).emplace_back(what, e);
    }
    void flush_marks() {
        if (!marks_on || marks.empty()) return;
        cudaDeviceSynchronize();
        for (size_t i = 1; i < marks.size(); ++i) {
            float ms = 0; cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
            fprintf(stderr, "[vb mark ] %-54s %9.3f ms\n", marks[i].first.c_str(), ms);
        }
        for (size_t i = 0; i < side_marks.size(); ++i) {    // halo stream: since the previous mark there, and since the first mark of the main stream
            float ms = 0, at = 0;
            if (i) cudaEventElapsedTime(&ms, side_marks[i - 1].second, side_marks[i].second);
            cudaEventElapsedTime(&at, marks[0].second, side_marks[i].second);
            fprintf(stderr, "[vb halo ] %-54s %9.3f ms   (at %8.3f ms)\n", side_marks[i].first.c_str(), ms, at);
        }
        { float at = 0; cudaEventElapsedTime(&at, marks[0].second, marks.back().second); fprintf(stderr, "[vb mark ] %-54s             (at %8.3f ms)\n", "last mark of the main stream", at); }
        for (auto& m : marks) cudaEventDestroy(m.second);
        for (auto& m : side_marks) cudaEventDestroy(m.second);
        marks.clear(); side_marks.clear();
    }
};
Trace g_trace;

// persisting-L2 set-aside: only touched when the wanted size changes (cudaDeviceSetLimit synchronises the device)
void set_persisting_l2_mb(int mb) {
    static int current = -1;
    if (current == mb) return;
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)mb << 20);
    cudaGetLastError();
    current = mb;
}
void require_device() {
    if (g_device < 0) throw CudaError("vahana_b200: vb_init() has not been called or no CUDA device is available (no CPU fallback)");
}

// ---- caching device allocator: finish_write! ping-pongs buffers of a few recurring sizes every apply ----
struct Pool {
    std::multimap<size_t, void*> free_;
    std::unordered_map<void*, size_t> size_;
    // size classes: four per octave above 1 MB (<= 25 % slack), so the slowly growing buffers of a growing population keep
    // hitting cached blocks instead of calling cudaMalloc every step
    static size_t round(size_t b) {
        if (b <= (1u << 20)) return ((b + 255) / 256) * 256;
        int lg = 63 - __builtin_clzll((unsigned long long)b);
        const size_t g = (size_t)1 << (lg - 2);
        return ((b + g - 1) / g) * g;
    }
    void* alloc(size_t bytes) {
        if (bytes == 0) bytes = 256;
        const size_t r = round(bytes);
        auto it = free_.lower_bound(r);
        if (it != free_.end() && it->first <= r + r / 2) {
            void* p = it->second;
            free_.erase(it);
            return p;
        }
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, r);
        if (e != cudaSuccess) {
            release_all();
            e = cudaMalloc(&p, r);
            if (e != cudaSuccess) throw CudaError(std::string("cudaMalloc(") + std::to_string(r) + "): " + cudaGetErrorString(e));
        }
        size_[p] = r;
        return p;
    }
    void free(void* p) {
        if (!p) return;
        auto it = size_.find(p);
        if (it == size_.end()) return;
        free_.insert({it->second, p});
    }
    void release_all() {
        for (auto& kv : free_) { cudaFree(kv.second); size_.erase(kv.second); }
        free_.clear();
    }
    size_t cached_bytes() const { size_t b = 0; for (auto& kv : free_) b += kv.first; return b; }
};
Pool g_pool;

template <class T> T* dalloc(size_t n) { return (T*)g_pool.alloc(n * sizeof(T)); }
inline void dfree(void* p) { g_pool.free(p); }

std::map<std::pair<std::string, std::string>, const vb::TransitionInfo*>& registry() {
    static std::map<std::pair<std::string, std::string>, const vb::TransitionInfo*> r;
    return r;
}

std::map<std::pair<std::string, std::string>, const vb::MapInfo*>& map_registry() {
    static std::map<std::pair<std::string, std::string>, const vb::MapInfo*> r;
    return r;
}

// ---- engine-side kernels -------------------------------------------------------------------------------
// raw (AgentID) adds -> composite append log.  Validation as in the reference's add_edge! (ids must
// name an existing slot of a registered type; :SingleType target must match).
struct TranslateArgs {
    const uint64_t* to; const uint64_t* from; uint64_t n;
    uint32_t* log_to; uint32_t* log_from; uint64_t pos0;
    uint32_t base[vb::MAX_AGENT_TYPES + 2]; uint32_t nslots[vb::MAX_AGENT_TYPES + 1];
    uint32_t lcap[vb::MAX_AGENT_TYPES + 1]; uint32_t nghost[vb::MAX_AGENT_TYPES + 1]; const uint64_t* ghost_ids[vb::MAX_AGENT_TYPES + 1];
    uint32_t ntypes; int32_t target; int ignore_from; uint32_t rank; uint32_t* error;
};
__global__ void translate_edges_kernel(const TranslateArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const uint64_t to = a.to[i];
    const uint32_t tt = vb::type_nr(to);
    const uint64_t tnr = vb::agent_nr(to);
    uint32_t row = 0;
    if (tt < 1 || tt > a.ntypes || tnr < 1 || tnr > a.nslots[tt]) { atomicOr(a.error, (uint32_t)vb::DERR_BAD_ID); }
    else if (vb::process_nr(to) != a.rank) { atomicOr(a.error, (uint32_t)vb::DERR_REMOTE); }   // edges are stored on the rank of their target
    else if (a.target) { if ((int)tt != a.target) atomicOr(a.error, (uint32_t)vb::DERR_SINGLETYPE_MISMATCH); row = (uint32_t)(tnr - 1); }
    else row = a.base[tt] + (uint32_t)(tnr - 1);
    a.log_to[a.pos0 + i] = row;
    if (!a.ignore_from) {
        const uint64_t fr = a.from[i];
        const uint32_t ft = vb::type_nr(fr);
        const uint64_t fnr = vb::agent_nr(fr);
        uint32_t c = 0;
        if (ft < 1 || ft > a.ntypes || fnr < 1) atomicOr(a.error, (uint32_t)vb::DERR_BAD_ID);
        else if (vb::process_nr(fr) != a.rank) {   // remote source: its slot in the ghost segment
            uint32_t lo = 0, hi = a.nghost[ft];
            const uint64_t* g = a.ghost_ids[ft];
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (g[mid] < fr) lo = mid + 1; else hi = mid; }
            if (lo >= a.nghost[ft] || g[lo] != fr) atomicOr(a.error, (uint32_t)vb::DERR_BAD_ID);
            else c = a.base[ft] + a.lcap[ft] + lo;
        }
        else if (fnr > a.nslots[ft]) atomicOr(a.error, (uint32_t)vb::DERR_BAD_ID);
        else c = a.base[ft] + (uint32_t)(fnr - 1);
        a.log_from[a.pos0 + i] = c;
    }
}
__global__ void count_adds_kernel(const uint32_t* __restrict__ rows, uint64_t n, uint32_t* __restrict__ cnt, int flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flag) cnt[rows[i]] = 1u; else atomicAdd(&cnt[rows[i]], 1u);
}
// remove_edges! records that arrived from other ranks (removeedges_alltoall!, src/MPI.jl:462-470: every received (source | 0, target)
// is applied with the ordinary local remove_edges!) -> records of the local remove log.  An id that names nobody on this rank
// (unknown type / slot, a remote source this rank never mirrored, a :SingleType mismatch) matches no entry: row = 0xffffffff.
struct TranslateRmArgs {
    const uint64_t* to; const uint64_t* from; uint32_t n;
    uint32_t* rm_row; uint32_t* rm_from; uint32_t* rm_mark; uint64_t* rm_to64; uint64_t* rm_from64; uint32_t pos0; uint32_t mark;
    uint32_t base[vb::MAX_AGENT_TYPES + 2]; uint32_t lcap[vb::MAX_AGENT_TYPES + 1]; uint32_t nghost[vb::MAX_AGENT_TYPES + 1];
    const uint64_t* ghost_ids[vb::MAX_AGENT_TYPES + 1];
    uint32_t ntypes; int32_t target; uint32_t rank;
};
__global__ void translate_removes_kernel(const TranslateRmArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const uint64_t to = a.to[i], fr = a.from[i];
    const uint32_t tt = vb::type_nr(to);
    const uint64_t tnr = vb::agent_nr(to);
    uint32_t row = 0xffffffffu, fcomp = 0xffffffffu;
    if (vb::process_nr(to) == a.rank && tt >= 1 && tt <= a.ntypes && tnr >= 1 && tnr <= a.lcap[tt] && (!a.target || (int)tt == a.target))
        row = (a.target ? 0u : a.base[tt]) + (uint32_t)(tnr - 1);
    if (fr) {
        const uint32_t ft = vb::type_nr(fr);
        const uint64_t fnr = vb::agent_nr(fr);
        bool ok = ft >= 1 && ft <= a.ntypes && fnr >= 1;
        if (ok && vb::process_nr(fr) != a.rank) {
            uint32_t lo = 0, hi = a.nghost[ft];
            const uint64_t* g = a.ghost_ids[ft];
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (g[mid] < fr) lo = mid + 1; else hi = mid; }
            ok = lo < a.nghost[ft] && g[lo] == fr;
            if (ok) fcomp = a.base[ft] + a.lcap[ft] + lo;
        } else if (ok) {
            ok = fnr <= a.lcap[ft];
            if (ok) fcomp = a.base[ft] + (uint32_t)(fnr - 1);
        }
        if (!ok) row = 0xffffffffu;
    }
    const uint32_t p = a.pos0 + i;
    a.rm_row[p] = row; a.rm_from[p] = fcomp; a.rm_mark[p] = a.mark;
    if (a.rm_to64) { a.rm_to64[p] = 0ull; a.rm_from64[p] = 0ull; }
}
// records that travel: a target id was parked and names an existing rank (a target on a rank that does not exist matches nothing)
__global__ void remote_remove_flags_kernel(const uint64_t* __restrict__ to, uint32_t n, uint32_t nranks, uint32_t* __restrict__ flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (to[i] != 0ull && vb::process_nr(to[i]) < nranks) ? 1u : 0u;
}
// compacts the flagged (to, from) pairs and notes the destination rank of each
__global__ void compact_removes_kernel(const uint64_t* __restrict__ to, const uint64_t* __restrict__ from, const uint32_t* __restrict__ flag,
                                       const uint32_t* __restrict__ pos, uint32_t n, uint64_t* __restrict__ oto, uint64_t* __restrict__ ofrom,
                                       uint32_t* __restrict__ odst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    oto[pos[i]] = to[i]; ofrom[pos[i]] = from[i]; odst[pos[i]] = vb::process_nr(to[i]);
}
// keep only the last entry of every run of equal keys (:SingleEdge "overwrite the slot", EdgeMethods.jl:441-445);
// flags a run whose entries differ (second add with another value asserts for Dict containers, :267-293)
__global__ void last_of_run_flags_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == n - 1 || keys[i + 1] != keys[i]) ? 1u : 0u;
}
__global__ void single_edge_conflict_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ from, const uint8_t* __restrict__ st,
                                            uint32_t st_stride, uint32_t word, uint32_t ncols, uint32_t n, uint32_t* conflict) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n || keys[i] != keys[i + 1]) return;
    bool diff = from && from[i] != from[i + 1];
    for (uint32_t c = 0; st && c < ncols && !diff; ++c)
        for (uint32_t b = 0; b < word; ++b)
            if (st[(size_t)c * st_stride * word + (size_t)i * word + b] != st[(size_t)c * st_stride * word + (size_t)(i + 1) * word + b]) { diff = true; break; }
    if (diff) *conflict = 1u;
}
__global__ void compact_u32_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos, uint32_t n,
                                   uint32_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) out[pos[i]] = in[i];
}
__global__ void compact_soa_kernel(const uint8_t* __restrict__ in, uint32_t istride, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos,
                                   uint32_t n, uint8_t* __restrict__ out, uint32_t ostride, uint32_t word, uint32_t ncols) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n * ncols) return;
    const uint32_t i = (uint32_t)(t % n), c = (uint32_t)(t / n);
    if (!flag[i]) return;
    for (uint32_t b = 0; b < word; ++b) out[(size_t)c * ostride * word + (size_t)pos[i] * word + b] = in[(size_t)c * istride * word + (size_t)i * word + b];
}
// 1- and 2-byte edge states ride through the sort as a 32-bit payload (cheaper than a permutation + random gather)
__global__ void widen_state_kernel(const uint8_t* __restrict__ in, uint32_t n, uint32_t word, uint32_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = word == 1 ? (uint32_t)in[i] : (uint32_t)reinterpret_cast<const uint16_t*>(in)[i];
}
// a one- or two-byte edge state rides in the key's unused high bits through the sort (the passes only look at the low `bits` bits): the
// log is sorted as (key, from) instead of (key, from, state widened to 4 bytes)
__global__ void pack_state_into_key_kernel(uint32_t* __restrict__ key, const uint8_t* __restrict__ st, uint32_t n, uint32_t word, int bits) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = word == 1 ? (uint32_t)st[i] : (uint32_t)reinterpret_cast<const uint16_t*>(st)[i];
    key[i] |= v << bits;
}
__global__ void unpack_state_from_key_kernel(uint32_t* __restrict__ key, uint8_t* __restrict__ st, uint32_t n, uint32_t word, int bits) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = key[i];
    if (word == 1) st[i] = (uint8_t)(k >> bits); else reinterpret_cast<uint16_t*>(st)[i] = (uint16_t)(k >> bits);
    key[i] = k & ((1u << bits) - 1u);
}
__global__ void narrow_state_kernel(const uint32_t* __restrict__ in, uint32_t n, uint32_t word, uint8_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (word == 1) out[i] = (uint8_t)in[i]; else reinterpret_cast<uint16_t*>(out)[i] = (uint16_t)in[i];
}
__global__ void gather_soa_kernel(const uint8_t* __restrict__ in, uint32_t istride, const uint32_t* __restrict__ perm, uint32_t n, uint8_t* __restrict__ out,
                                  uint32_t ostride, uint32_t word, uint32_t ncols) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n * ncols) return;
    const uint32_t i = (uint32_t)(t % n), c = (uint32_t)(t / n);
    const uint32_t s = perm[i];
    for (uint32_t b = 0; b < word; ++b) out[(size_t)c * ostride * word + (size_t)i * word + b] = in[(size_t)c * istride * word + (size_t)s * word + b];
}
// merge an existing CSR (old) with newly sorted edges (new, run offsets noff): row r of the result is
// old row r followed by the new entries of r — push! order (EdgeMethods.jl:495-496).  single_edge: a new
// entry replaces the old one.
struct MergeArgs {
    const uint32_t* ooff; const uint32_t* osrc; const uint8_t* ost; uint32_t ostride; uint32_t orows;
    const uint32_t* noff; const uint32_t* nsrc; const uint8_t* nst; uint32_t nstride;
    const uint32_t* off; uint32_t* src; uint8_t* st; uint32_t stride;
    uint32_t rows; uint32_t word, ncols; int single_edge;
    uint32_t* conflict;     // :SingleEdge with add_existing: set when a row's kept entry and its new entry differ (_can_add, EdgeMethods.jl:267-293)
};
__global__ void merge_counts_kernel(const uint32_t* __restrict__ ooff, uint32_t orows, const uint32_t* __restrict__ ncnt, uint32_t rows, int single_edge,
                                    uint32_t* __restrict__ cnt) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > rows) return;
    if (r == rows) { cnt[r] = 0; return; }
    const uint32_t o = r < orows ? ooff[r + 1] - ooff[r] : 0;
    const uint32_t nn = ncnt[r];
    cnt[r] = single_edge ? (nn ? 1u : o) : o + nn;
}
__global__ void merge_copy_kernel(const MergeArgs a) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.rows) return;
    uint32_t d = a.off[r];
    const uint32_t nb = a.noff[r], ne = a.noff[r + 1];
    if (a.conflict && a.single_edge && ne > nb && r < a.orows && a.ooff[r + 1] > a.ooff[r]) {
        // the write container already held an edge for this target (add_existing) and the transition added another one: allowed only
        // when the value is identical
        const uint32_t ko = a.ooff[r], kn = ne - 1;
        bool same = !a.src || a.osrc[ko] == a.nsrc[kn];
        for (uint32_t c = 0; a.st && c < a.ncols; ++c)
            for (uint32_t b = 0; b < a.word; ++b)
                same &= a.ost[(size_t)c * a.ostride * a.word + (size_t)ko * a.word + b] == a.nst[(size_t)c * a.nstride * a.word + (size_t)kn * a.word + b];
        if (!same) atomicOr(a.conflict, 1u);
    }
    if (!(a.single_edge && ne > nb) && r < a.orows) {
        for (uint32_t k = a.ooff[r]; k < a.ooff[r + 1]; ++k, ++d) {
            if (a.src) a.src[d] = a.osrc[k];
            for (uint32_t c = 0; a.st && c < a.ncols; ++c)
                for (uint32_t b = 0; b < a.word; ++b)
                    a.st[(size_t)c * a.stride * a.word + (size_t)d * a.word + b] = a.ost[(size_t)c * a.ostride * a.word + (size_t)k * a.word + b];
        }
    }
    for (uint32_t k = a.single_edge ? (ne > nb ? ne - 1 : ne) : nb; k < ne; ++k, ++d) {
        if (a.src) a.src[d] = a.nsrc[k];
        for (uint32_t c = 0; a.st && c < a.ncols; ++c)
            for (uint32_t b = 0; b < a.word; ++b)
                a.st[(size_t)c * a.stride * a.word + (size_t)d * a.word + b] = a.nst[(size_t)c * a.nstride * a.word + (size_t)k * a.word + b];
    }
}
// remove_edges! (src/EdgeMethods.jl:527-599) applied at finish_write!: a record (row, from|ALL, P) deletes the existing entries of
// the row (matching `from`) and the appended ones at log positions < P, i.e. what was in the write container when it was called.
__global__ void rm_cut_kernel(const uint32_t* __restrict__ row, const uint32_t* __restrict__ from, const uint32_t* __restrict__ mark, uint32_t n,
                              uint32_t* __restrict__ cutA, uint32_t* __restrict__ pairflag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool valid = row[i] != 0xffffffffu;
    if (valid && from[i] == 0xffffffffu) atomicMax(&cutA[row[i]], mark[i] + 1u);
    pairflag[i] = (valid && from[i] != 0xffffffffu) ? 1u : 0u;
}
struct RmArgs {
    const uint32_t* cutA;                                           // per row: 0 = no remove_edges!(to), else max P + 1
    const uint32_t* poff; const uint32_t* pfrom; const uint32_t* pmark;   // (from,to) records grouped by row (nullptr if none)
};
__device__ __forceinline__ bool rm_hits(const RmArgs& a, uint32_t row, uint32_t from, bool has_from, long long p) {
    const uint32_t c = a.cutA[row];
    if (c && p + 1 < (long long)c) return true;
    if (a.poff && has_from)
        for (uint32_t k = a.poff[row]; k < a.poff[row + 1]; ++k)
            if (a.pfrom[k] == from && p < (long long)a.pmark[k]) return true;
    return false;
}
__global__ void rm_filter_log_kernel(const uint32_t* __restrict__ to, const uint32_t* __restrict__ from, uint32_t n, const RmArgs a, uint32_t* __restrict__ keep) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keep[i] = rm_hits(a, to[i], from ? from[i] : 0u, from != nullptr, (long long)i) ? 0u : 1u;
}
__global__ void rm_filter_old_count_kernel(const uint32_t* __restrict__ off, const uint32_t* __restrict__ src, uint32_t rows, const RmArgs a,
                                           uint32_t* __restrict__ cnt) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > rows) return;
    if (r == rows) { cnt[r] = 0; return; }
    uint32_t keep = 0;
    for (uint32_t k = off[r]; k < off[r + 1]; ++k) keep += rm_hits(a, (uint32_t)r, src ? src[k] : 0u, src != nullptr, -1) ? 0u : 1u;
    cnt[r] = keep;
}
// dead-agent edge purge (src/AgentMethods.jl:314-358, src/EdgeMethods.jl:606-631,897-920): drop rows whose
// target died this apply and entries whose source died; `dead` is indexed by composite agent index.
struct PurgeArgs {
    const uint32_t* off; const uint32_t* src; const uint8_t* st; uint32_t stride; uint32_t rows;
    const uint8_t* dead; uint32_t row_base;   // composite index of row 0 (SingleType: base[target], else 0)
    int check_src;
    uint32_t* cnt;                            // out: kept entries per row [rows + 1]
    const uint32_t* noff; uint32_t* nsrc; uint8_t* nst; uint32_t nstride; uint32_t word, ncols;
};
__global__ void purge_count_kernel(const PurgeArgs a) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > a.rows) return;
    if (r == a.rows) { a.cnt[r] = 0; return; }
    uint32_t keep = 0;
    if (!a.dead[a.row_base + r]) {
        if (!a.check_src) keep = a.off[r + 1] - a.off[r];
        else for (uint32_t k = a.off[r]; k < a.off[r + 1]; ++k) keep += a.dead[a.src[k]] ? 0u : 1u;
    }
    a.cnt[r] = keep;
}
__global__ void purge_copy_kernel(const PurgeArgs a) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.rows || a.dead[a.row_base + r]) return;
    uint32_t d = a.noff[r];
    for (uint32_t k = a.off[r]; k < a.off[r + 1]; ++k) {
        if (a.check_src && a.dead[a.src[k]]) continue;
        if (a.nsrc) a.nsrc[d] = a.src[k];
        for (uint32_t c = 0; a.nst && c < a.ncols; ++c)
            for (uint32_t b = 0; b < a.word; ++b)
                a.nst[(size_t)c * a.nstride * a.word + (size_t)d * a.word + b] = a.st[(size_t)c * a.stride * a.word + (size_t)k * a.word + b];
        ++d;
    }
}
__global__ void rm_filter_old_copy_kernel(const PurgeArgs a, const RmArgs rm) {   // PurgeArgs reused: off/src/st -> noff/nsrc/nst
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.rows) return;
    uint32_t d = a.noff[r];
    for (uint32_t k = a.off[r]; k < a.off[r + 1]; ++k) {
        if (rm_hits(rm, (uint32_t)r, a.src ? a.src[k] : 0u, a.src != nullptr, -1)) continue;
        if (a.nsrc) a.nsrc[d] = a.src[k];
        for (uint32_t c = 0; a.nst && c < a.ncols; ++c)
            for (uint32_t b = 0; b < a.word; ++b)
                a.nst[(size_t)c * a.nstride * a.word + (size_t)d * a.word + b] = a.st[(size_t)c * a.stride * a.word + (size_t)k * a.word + b];
        ++d;
    }
}
__global__ void rm_zero_rows_kernel(uint32_t* __restrict__ cnt, uint32_t rows, const uint32_t* __restrict__ cutA) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows && cutA[r]) cnt[r] = 0;
}
__global__ void purge_rows_cnt_kernel(uint32_t* __restrict__ cnt, uint32_t rows, const uint8_t* __restrict__ dead, uint32_t row_base) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows && dead[row_base + r]) cnt[r] = 0;
}
// multi-GPU ghost discovery: remote source ids of raw adds
__global__ void remote_flags_kernel(const uint64_t* __restrict__ ids, uint64_t n, uint32_t rank, uint32_t* __restrict__ flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = vb::process_nr(ids[i]) != rank ? 1u : 0u;
}
__global__ void compact_u64_split_kernel(const uint64_t* __restrict__ in, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos, uint64_t n,
                                         uint32_t* __restrict__ lo, uint32_t* __restrict__ hi, uint64_t out0) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) { lo[out0 + pos[i]] = (uint32_t)in[i]; hi[out0 + pos[i]] = (uint32_t)(in[i] >> 32); }
}
__global__ void unique_flags64_kernel(const uint32_t* __restrict__ hi, const uint32_t* __restrict__ lo, uint64_t n, uint32_t* __restrict__ flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || hi[i] != hi[i - 1] || lo[i] != lo[i - 1]) ? 1u : 0u;
}
__global__ void compact_join64_kernel(const uint32_t* __restrict__ hi, const uint32_t* __restrict__ lo, const uint32_t* __restrict__ flag,
                                      const uint32_t* __restrict__ pos, uint64_t n, uint64_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) out[pos[i]] = ((uint64_t)hi[i] << 32) | lo[i];
}
// lower_bound of (type, rank, nr = 0) boundaries in the sorted ghost id list
__global__ void ghost_bounds_kernel(const uint64_t* __restrict__ ids, uint32_t n, uint32_t ntypes, uint32_t nranks, uint32_t* __restrict__ bounds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > ntypes * nranks) return;
    const uint32_t t = i / nranks + 1, r = i % nranks;
    const uint64_t key = i == ntypes * nranks ? ~0ull : vb::agent_id(t, r, 0);
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (ids[mid] < key) lo = mid + 1; else hi = mid; }
    bounds[i] = lo;
}
__global__ void find_ghost_kernel(const uint64_t* __restrict__ ids, uint32_t n, uint64_t id, int64_t* __restrict__ out) {   // binary search of the sorted ghost table
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (ids[mid] < id) lo = mid + 1; else hi = mid; }
    *out = (lo < n && ids[lo] == id) ? (int64_t)lo : -1;
}
__global__ void ids_to_slots_kernel(const uint64_t* __restrict__ ids, uint32_t n, uint32_t* __restrict__ slots) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slots[i] = (uint32_t)(vb::agent_nr(ids[i]) - 1);
}
// halo pack: send[c][i] = state[c][slot[i]] for every SoA column c
__global__ void halo_pack_kernel(const uint8_t* __restrict__ cols, uint32_t stride, const uint32_t* __restrict__ slots, uint32_t n, uint8_t* __restrict__ out,
                                 uint32_t word, uint32_t ncols) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n * ncols) return;
    const uint32_t i = (uint32_t)(t % n), c = (uint32_t)(t / n);
    const uint8_t* sp = cols + (size_t)c * stride * word + (size_t)slots[i] * word;
    uint8_t* dp = out + (size_t)c * n * word + (size_t)i * word;
    if (word == 8) *reinterpret_cast<uint64_t*>(dp) = *reinterpret_cast<const uint64_t*>(sp);
    else if (word == 4) *reinterpret_cast<uint32_t*>(dp) = *reinterpret_cast<const uint32_t*>(sp);
    else for (uint32_t b = 0; b < word; ++b) dp[b] = sp[b];
}
// halo push: the owner writes the requested states straight into the peers' ghost segments (peer memory over NVLink / NVSwitch):
// pack and transfer are one kernel, the stores of a warp are contiguous in the peer's buffer (transmit_agents!, src/MPI.jl:155-267)
struct HaloPushArgs {
    const uint8_t* cols; uint32_t stride; const uint32_t* slots; uint32_t n, word, ncols, npeers;
    // peers in rotated order (rank + 1, rank + 2, ...): at any moment every receiver is fed by a different sender, instead of all
    // ranks pushing into rank 0 first (which shares one ingress between all senders: measured 300 GB/s instead of 800 GB/s per GPU)
    uint32_t voff[17];                        // [npeers + 1] prefix of the per-peer counts in that order
    uint32_t first[16];                       // position in `slots` of the first state requested by that peer
    uint8_t* remote[16]; uint32_t rstride[16]; uint32_t rghost0[16];
};
__global__ void halo_push_kernel(const HaloPushArgs a) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)a.n * a.ncols) return;
    const uint32_t v = (uint32_t)(t % a.n), c = (uint32_t)(t / a.n);
    uint32_t k = 0;
    while (k + 1 < a.npeers && v >= a.voff[k + 1]) ++k;
    const uint32_t j = v - a.voff[k];
    const uint8_t* sp = a.cols + (size_t)c * a.stride * a.word + (size_t)a.slots[a.first[k] + j] * a.word;
    uint8_t* dp = a.remote[k] + (size_t)c * a.rstride[k] * a.word + (size_t)(a.rghost0[k] + j) * a.word;
    if (a.word == 8) *reinterpret_cast<uint64_t*>(dp) = *reinterpret_cast<const uint64_t*>(sp);
    else if (a.word == 4) *reinterpret_cast<uint32_t*>(dp) = *reinterpret_cast<const uint32_t*>(sp);
    else if (a.word == 16) *reinterpret_cast<uint4*>(dp) = *reinterpret_cast<const uint4*>(sp);
    else for (uint32_t b = 0; b < a.word; ++b) dp[b] = sp[b];
}
// Cross-rank barrier over peer memory: every rank owns a small flag array mapped into the peers (CUDA IPC); a rank announces epoch E
// by storing it into slot [rank] of every peer's array and waits until its own array holds >= E in every peer's slot.  One warp, a few
// microseconds over NVLink, stream-ordered like any kernel (the ncclAllGather it replaces costs a collective launch per barrier).
struct PeerFlagPtrs { unsigned long long* p[16]; };
__global__ void peer_barrier_kernel(unsigned long long* mine, const PeerFlagPtrs peers, uint32_t rank, uint32_t P, unsigned long long epoch) {
    const uint32_t r = threadIdx.x;
    __threadfence_system();                                   // everything this stream wrote before (also into peer memory) is visible first
    if (r < P && r != rank) {
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.p[r] + rank), "l"(epoch) : "memory");
        unsigned long long seen = 0;
        do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine + r) : "memory"); } while (seen < epoch);
    }
    __syncwarp();
    __threadfence_system();
}
// the same selection (one phase's sub-range of every peer's list), packed into the send buffer at the list positions: the copy engines
// then move every peer's part over NVLink without occupying an SM (halo_exchange)
__global__ void halo_pack_ranges_kernel(const HaloPushArgs a, uint8_t* __restrict__ out, uint32_t ns) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)a.n * a.ncols) return;
    const uint32_t v = (uint32_t)(t % a.n), c = (uint32_t)(t / a.n);
    uint32_t k = 0;
    while (k + 1 < a.npeers && v >= a.voff[k + 1]) ++k;
    const uint32_t i = a.first[k] + (v - a.voff[k]);
    const uint8_t* sp = a.cols + (size_t)c * a.stride * a.word + (size_t)a.slots[i] * a.word;
    uint8_t* dp = out + ((size_t)c * ns + i) * a.word;
    if (a.word == 8) *reinterpret_cast<uint64_t*>(dp) = *reinterpret_cast<const uint64_t*>(sp);
    else if (a.word == 4) *reinterpret_cast<uint32_t*>(dp) = *reinterpret_cast<const uint32_t*>(sp);
    else if (a.word == 16) *reinterpret_cast<uint4*>(dp) = *reinterpret_cast<const uint4*>(sp);
    else for (uint32_t b = 0; b < a.word; ++b) dp[b] = sp[b];
}
// died-agent ids of the other ranks (C7: join(aids), src/AgentMethods.jl:338): mark the ghosts that mirror them
struct MarkDeadArgs {
    const uint64_t* ids; uint32_t n; uint8_t* dead; uint32_t rank; uint32_t ntypes;
    uint32_t base[vb::MAX_AGENT_TYPES + 2]; uint32_t lcap[vb::MAX_AGENT_TYPES + 1]; uint32_t nghost[vb::MAX_AGENT_TYPES + 1]; const uint64_t* ghost_ids[vb::MAX_AGENT_TYPES + 1];
};
__global__ void mark_remote_dead_kernel(const MarkDeadArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const uint64_t id = a.ids[i];
    const uint32_t t = vb::type_nr(id);
    if (id == 0 || t < 1 || t > a.ntypes || vb::process_nr(id) == a.rank) return;
    const uint64_t* g = a.ghost_ids[t];
    uint32_t lo = 0, hi = a.nghost[t];
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (g[mid] < id) lo = mid + 1; else hi = mid; }
    if (lo < a.nghost[t] && g[lo] == id) a.dead[a.base[t] + a.lcap[t] + lo] = 1;
}
__global__ void slots_to_ids_kernel(const uint32_t* __restrict__ slots, uint32_t n, uint32_t type, uint32_t rank, uint64_t* __restrict__ ids) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ids[i] = vb::agent_id(type, rank, (uint64_t)slots[i] + 1);
}
__global__ void gather_u64_kernel(const uint64_t* __restrict__ in, const uint32_t* __restrict__ perm, uint32_t n, uint64_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}
__global__ void gather_soa_to_aos_kernel(const uint8_t* __restrict__ cols, uint32_t stride, const uint32_t* __restrict__ perm, uint32_t n, uint8_t* __restrict__ aos,
                                         uint32_t size, uint32_t word) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ncols = size / word;
    if (t >= (uint64_t)n * ncols) return;
    const uint32_t i = (uint32_t)(t / ncols), c = (uint32_t)(t % ncols);
    const uint8_t* sp = cols + (size_t)c * stride * word + (size_t)perm[i] * word;
    uint8_t* dp = aos + (size_t)i * size + (size_t)c * word;
    for (uint32_t b = 0; b < word; ++b) dp[b] = sp[b];
}
__global__ void heavy_flags_kernel(const uint32_t* __restrict__ off, uint32_t row0, uint32_t n, uint32_t rows, uint32_t heavy_min, uint32_t* __restrict__ flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = row0 + (uint32_t)i;
    flag[i] = (r < rows && off[r + 1] - off[r] >= heavy_min) ? 1u : 0u;
}

// ---- source-blocked view of a CSR (reduce transitions): count -> scan -> fill, one thread per called row ----------------------
// Rows with >= heavy_min entries stay with the block-per-agent pass and contribute nothing here.
__global__ void blk_heavy_bits_kernel(const uint32_t* __restrict__ off, uint32_t row0, uint32_t n, uint32_t rows, uint32_t heavy_min, uint32_t* __restrict__ bits) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = row0 + (uint32_t)i;
    const bool heavy = i < n && r < rows && off[r + 1] - off[r] >= heavy_min;
    const unsigned b = __ballot_sync(0xffffffffu, heavy);
    if ((threadIdx.x & 31) == 0 && i < n) bits[i >> 5] = b;
}
struct BlkBuildArgs {
    const uint32_t* off; const uint32_t* src; uint32_t row0, n, rows, heavy_min;
    uint32_t tb, nsl;            // composite base and slot count (local + ghosts) of the source type
    uint32_t bsize, nb, rpad;
    uint32_t* boff; uint32_t* bsrc; uint32_t* error;
};
__global__ void blk_active_flags_kernel(const uint32_t* __restrict__ boff, uint32_t n, uint32_t* __restrict__ flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = boff[i + 1] > boff[i] ? 1u : 0u;
}
__global__ void blk_listed_offsets_kernel(const uint32_t* __restrict__ boff, const uint32_t* __restrict__ rows, uint32_t cnt, uint32_t end, uint32_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) out[i] = boff[rows[i]];
    else if (i == cnt) out[i] = end;
}
__global__ void blk_count_kernel(const BlkBuildArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const uint32_t r = a.row0 + (uint32_t)i;
    if (r >= a.rows) return;
    const uint32_t b0 = a.off[r], b1 = a.off[r + 1];
    if (b1 - b0 >= a.heavy_min) return;
    for (uint32_t k = b0; k < b1; ++k) {
        const uint32_t s = a.src[k] - a.tb;
        if (s >= a.nsl) { atomicOr(a.error, 1u); continue; }     // a source of another agent type
        a.boff[(size_t)(s / a.bsize) * a.rpad + i] += 1;          // one thread per row: no atomics
    }
}
__global__ void blk_fill_kernel(const BlkBuildArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const uint32_t r = a.row0 + (uint32_t)i;
    if (r >= a.rows) return;
    const uint32_t b0 = a.off[r], b1 = a.off[r + 1];
    if (b1 - b0 >= a.heavy_min) return;
    uint16_t cur[64];                                             // entries of this row already placed, per block (row < heavy_min <= 65535)
    for (uint32_t b = 0; b < a.nb; ++b) cur[b] = 0;
    for (uint32_t k = b0; k < b1; ++k) {
        const uint32_t s = a.src[k] - a.tb;
        if (s >= a.nsl) continue;
        const uint32_t b = s / a.bsize;
        a.bsrc[a.boff[(size_t)b * a.rpad + i] + cur[b]++] = s;    // row order is kept inside every block
    }
}
// ---- segmented source-blocked view (prefiltered sweeps): rows of more than seg_len entries are cut into segments ---------------------
struct SegBuildArgs {
    const uint32_t* off; const uint32_t* src; uint32_t row0, n, rows, seg_len;
    uint32_t tb, nsl, nb, spad, nseg;
    // blocks of the source type's slots: nbl blocks of bsize_l slots over the local slots [0, lcap); behind them ng ghost blocks, block
    // nbl + j = the j-th part (gpart[p] slots) of every owner p's range [goff[p], goff[p + 1]) of the ghost segment.  Entries of a
    // ghost block are stored relative to lcap.
    // The ghost blocks start at block gfirst: nbl (behind the local blocks) or nbl - 1 (the first ghost part shares the last local
    // block: one sweep less).  absolute != 0: entries hold the slot itself (all slots < 2^27), which is what lets a block mix ranges.
    uint32_t lcap, bsize_l, nbl, ng, nowners, gfirst, absolute;
    uint32_t goff[17], gfirst_len[16], gpart[16];      // owner p's range, the length of its first part and of each later part
    const uint32_t* sfirst;          // [n + 1] first segment of every called row
    __device__ __forceinline__ uint32_t block_of(uint32_t s, uint32_t& first) const {
        uint32_t b;
        if (s < lcap) { b = s / bsize_l; if (b >= nbl) b = nbl - 1; first = b * bsize_l; }
        else {
            const uint32_t g = s - lcap;
            uint32_t p = 0;
            while (p + 1 < nowners && g >= goff[p + 1]) ++p;
            const uint32_t x = g - goff[p];
            b = x < gfirst_len[p] ? 0u : 1u + (x - gfirst_len[p]) / gpart[p];
            if (b >= ng) b = ng - 1;
            b += gfirst; first = lcap;
        }
        if (absolute) first = 0;
        return b;
    }
    uint32_t* seg_row; uint32_t* seg_lo; uint32_t* seg_hi;
    uint32_t* boff; uint32_t* bsrc; uint32_t* error;
};
__global__ void seg_count_kernel(const uint32_t* __restrict__ off, uint32_t row0, uint32_t n, uint32_t rows, uint32_t seg_len, uint32_t* __restrict__ cnt) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { cnt[i] = 0; return; }
    const uint32_t r = row0 + (uint32_t)i;
    const uint32_t len = r < rows ? off[r + 1] - off[r] : 0u;
    cnt[i] = len <= seg_len ? 1u : (len + seg_len - 1) / seg_len;
}
__global__ void seg_fill_kernel(const SegBuildArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const uint32_t r = a.row0 + (uint32_t)i;
    const uint32_t lo = r < a.rows ? a.off[r] : 0u, hi = r < a.rows ? a.off[r + 1] : 0u;
    const uint32_t s0 = a.sfirst[i], ns = a.sfirst[i + 1] - s0;
    for (uint32_t k = 0; k < ns; ++k) {
        a.seg_row[s0 + k] = (uint32_t)i | (ns > 1 ? 0x80000000u : 0u);
        const uint32_t l = lo + k * a.seg_len;
        a.seg_lo[s0 + k] = l;
        a.seg_hi[s0 + k] = (hi - l > a.seg_len) ? l + a.seg_len : hi;
    }
}
__global__ void seg_blk_count_kernel(const SegBuildArgs a) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nseg) return;
    for (uint32_t k = a.seg_lo[t]; k < a.seg_hi[t]; ++k) {
        const uint32_t s = a.src[k] - a.tb;
        if (s >= a.nsl) { atomicOr(a.error, 1u); continue; }     // a source of another agent type
        uint32_t first;
        a.boff[(size_t)a.block_of(s, first) * a.spad + t] += 1;   // one thread per segment: no atomics
    }
}
__global__ void seg_blk_fill_kernel(const SegBuildArgs a) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nseg) return;
    uint16_t cur[64];                                             // entries of this segment already placed, per block (seg_len <= 65535)
    for (uint32_t b = 0; b < a.nb; ++b) cur[b] = 0;
    const uint32_t tag = ((uint32_t)t & 31u) << 27;
    for (uint32_t k = a.seg_lo[t]; k < a.seg_hi[t]; ++k) {
        const uint32_t s = a.src[k] - a.tb;
        if (s >= a.nsl) continue;
        uint32_t first;
        const uint32_t b = a.block_of(s, first);
        a.bsrc[a.boff[(size_t)b * a.spad + t] + cur[b]++] = (s - first) | tag;          // entry order is kept inside every block
    }
}
__global__ void seg_hub_flags_kernel(const uint32_t* __restrict__ sfirst, uint32_t n, uint32_t* __restrict__ flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = sfirst[i + 1] - sfirst[i] > 1 ? 1u : 0u;
}
__global__ void seg_hub_first_kernel(const uint32_t* __restrict__ hub_rows, uint32_t nhub, const uint32_t* __restrict__ sfirst, uint32_t* __restrict__ hub_seg) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nhub) { hub_seg[2 * i] = sfirst[hub_rows[i]]; hub_seg[2 * i + 1] = sfirst[hub_rows[i] + 1]; }    // [first, end) of every hub row
}
__global__ void mark_dead_kernel(const uint32_t* __restrict__ flag, uint32_t n, uint8_t* __restrict__ dead, uint32_t base) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) dead[base + i] = 1;
}
// composite index rebase after an agent type's capacity grew (bases of later types shift)
struct RebaseArgs {
    uint32_t old_base[vb::MAX_AGENT_TYPES + 2]; uint32_t new_base[vb::MAX_AGENT_TYPES + 2]; uint32_t ntypes;
    // ghost slots sit behind the local capacity: they move when the local capacity grows and are renumbered (remap) when
    // the sorted ghost table of a type gains entries
    uint32_t old_lcap[vb::MAX_AGENT_TYPES + 1]; uint32_t new_lcap[vb::MAX_AGENT_TYPES + 1]; const uint32_t* remap[vb::MAX_AGENT_TYPES + 1];
};
__global__ void rebase_values_kernel(uint32_t* __restrict__ v, uint64_t n, const RebaseArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = v[i];
    if (c == 0xffffffffu) return;   // "nobody" / "all" markers of the remove records
    uint32_t t = 1;
    while (t < a.ntypes && c >= a.old_base[t + 1]) ++t;
    uint32_t slot = c - a.old_base[t];
    if (slot >= a.old_lcap[t]) {
        uint32_t g = slot - a.old_lcap[t];
        if (a.remap[t]) g = a.remap[t][g];
        slot = a.new_lcap[t] + g;
    }
    v[i] = a.new_base[t] + slot;
}
__global__ void ghost_remap_kernel(const uint64_t* __restrict__ old_ids, uint32_t n_old, const uint64_t* __restrict__ new_ids, uint32_t n_new,
                                   uint32_t* __restrict__ remap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_old) return;
    const uint64_t id = old_ids[i];
    uint32_t lo = 0, hi = n_new;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (new_ids[mid] < id) lo = mid + 1; else hi = mid; }
    remap[i] = lo;
}
__global__ void split64_kernel(const uint64_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ lo, uint32_t* __restrict__ hi, uint64_t out0) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { lo[out0 + i] = (uint32_t)in[i]; hi[out0 + i] = (uint32_t)(in[i] >> 32); }
}
__global__ void rebase_rows_kernel(const uint32_t* __restrict__ ooff, uint32_t orows, uint32_t* __restrict__ noff, uint32_t nrows, const RebaseArgs a) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > nrows) return;
    if (r == nrows) { noff[r] = ooff[orows]; return; }
    uint32_t t = 1;
    while (t < a.ntypes && r >= a.new_base[t + 1]) ++t;
    const uint32_t slot = (uint32_t)r - a.new_base[t];
    const uint32_t ocap = a.old_base[t + 1] - a.old_base[t];
    // rows that did not exist before point at the end of their type's old segment (empty row)
    noff[r] = slot < ocap ? ooff[a.old_base[t] + slot] : ooff[a.old_base[t + 1]];
}
__global__ void rebase_cnt_kernel(const uint32_t* __restrict__ ocnt, uint32_t orows, uint32_t* __restrict__ ncnt, uint32_t nrows, const RebaseArgs a) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    uint32_t t = 1;
    while (t < a.ntypes && r >= a.new_base[t + 1]) ++t;
    const uint32_t slot = (uint32_t)r - a.new_base[t];
    const uint32_t ocap = a.old_base[t + 1] - a.old_base[t];
    ncnt[r] = (slot < ocap && a.old_base[t] + slot < orows) ? ocnt[a.old_base[t] + slot] : 0;
}
// composite index -> AgentID: a local slot is (type, this rank, slot + 1); a ghost slot (>= the local capacity old_lcap[t]) is the id the
// sorted ghost table of the type holds for it (ghost_ids passed through `remap`, reinterpreted: two words per id)
__global__ void comp_to_id_kernel(const uint32_t* __restrict__ comp, uint64_t n, uint64_t* __restrict__ ids, const RebaseArgs a, uint32_t rank) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = comp[i];
    uint32_t t = 1;
    while (t < a.ntypes && c >= a.old_base[t + 1]) ++t;
    const uint32_t slot = c - a.old_base[t];
    if (slot >= a.old_lcap[t] && a.remap[t]) ids[i] = reinterpret_cast<const uint64_t*>(a.remap[t])[slot - a.old_lcap[t]];
    else ids[i] = vb::agent_id(t, rank, (uint64_t)slot + 1);
}
__global__ void raster_cells_kernel(const uint64_t* __restrict__ ids, uint64_t n, uint32_t* __restrict__ cells, const RebaseArgs a, uint32_t rank) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cells[i] = vb::process_nr(ids[i]) == rank ? a.new_base[vb::type_nr(ids[i])] + (uint32_t)(vb::agent_nr(ids[i]) - 1) : 0xffffffffu;
}

// mapreduce (src/AgentMethods.jl:533-565, src/EdgeMethods.jl:972-994): map = field at byte offset (optionally
// compared with a constant, or the constant 1), fold with op.  Integer results are exact and order free;
// floating point uses a fixed reduction tree (deterministic run to run, within 1e-12 relative of the
// reference's left fold — tests state the tolerance).
struct MapArgs {
    const uint8_t* cols; uint64_t stride; uint32_t word; uint64_t n; const uint8_t* died;
    int offset; int dt; int has_cmp; long long cmp; int op;
};
template <bool FLOAT> struct Acc { typedef long long T; };
template <> struct Acc<true> { typedef double T; };
template <bool FLOAT>
__device__ __forceinline__ typename Acc<FLOAT>::T mr_identity(int op) {
    typedef typename Acc<FLOAT>::T T;
    switch (op) {
        case vb::OP_PROD: return (T)1;
        case vb::OP_MIN: return FLOAT ? (T)INFINITY : (T)INT64_MAX;
        case vb::OP_MAX: return FLOAT ? (T)-INFINITY : (T)(-INT64_MAX - 1);
        case vb::OP_AND: return (T)-1;
        default: return (T)0;
    }
}
template <bool FLOAT>
__device__ __forceinline__ typename Acc<FLOAT>::T mr_fold(typename Acc<FLOAT>::T a, typename Acc<FLOAT>::T b, int op) {
    typedef typename Acc<FLOAT>::T T;
    switch (op) {
        case vb::OP_SUM: return FLOAT ? a + b : (T)((unsigned long long)a + (unsigned long long)b);
        case vb::OP_PROD: return FLOAT ? a * b : (T)((unsigned long long)a * (unsigned long long)b);
        case vb::OP_MIN: return a < b ? a : b;
        case vb::OP_MAX: return a > b ? a : b;
        case vb::OP_AND: return (T)((long long)a & (long long)b);
        default: return (T)((long long)a | (long long)b);
    }
}
template <bool FLOAT>
__device__ __forceinline__ typename Acc<FLOAT>::T mr_load(const MapArgs& a, uint64_t i) {
    typedef typename Acc<FLOAT>::T T;
    if (a.dt < 0) return (T)1;
    long long iv = 0;
    double fv = 0;
    bool isf = false;
    const uint32_t w = a.word;
    switch (a.dt) {
        case vb::DT_I64: iv = vb::soa_load_field<long long>(a.cols, (uint32_t)a.stride, w, (uint32_t)i, a.offset); break;
        case vb::DT_F64: fv = vb::soa_load_field<double>(a.cols, (uint32_t)a.stride, w, (uint32_t)i, a.offset); isf = true; break;
        case vb::DT_I32: iv = vb::soa_load_field<int>(a.cols, (uint32_t)a.stride, w, (uint32_t)i, a.offset); break;
        case vb::DT_F32: fv = vb::soa_load_field<float>(a.cols, (uint32_t)a.stride, w, (uint32_t)i, a.offset); isf = true; break;
        default: iv = vb::soa_load_field<uint8_t>(a.cols, (uint32_t)a.stride, w, (uint32_t)i, a.offset); break;
    }
    if (a.has_cmp) { iv = isf ? (fv == (double)a.cmp) : (iv == a.cmp); isf = false; }
    return isf ? (T)fv : (T)iv;
}
template <bool FLOAT>
__global__ void __launch_bounds__(256) mapreduce_kernel(const MapArgs a, typename Acc<FLOAT>::T* __restrict__ partial) {
    typedef typename Acc<FLOAT>::T T;
    __shared__ T sm[8];
    T acc = mr_identity<FLOAT>(a.op);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (a.died && a.died[i]) continue;
        acc = mr_fold<FLOAT>(mr_load<FLOAT>(a, i), acc, a.op);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = mr_fold<FLOAT>(acc, __shfl_xor_sync(0xffffffffu, acc, o), a.op);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        T v = threadIdx.x < 8 ? sm[threadIdx.x] : mr_identity<FLOAT>(a.op);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v = mr_fold<FLOAT>(v, __shfl_xor_sync(0xffffffffu, v, o), a.op);
        if (threadIdx.x == 0) partial[blockIdx.x] = v;
    }
}
template <bool FLOAT>
__global__ void __launch_bounds__(256) mapreduce_final_kernel(const typename Acc<FLOAT>::T* __restrict__ partial, uint32_t n, int op,
                                                            typename Acc<FLOAT>::T* __restrict__ out) {
    typedef typename Acc<FLOAT>::T T;
    __shared__ T sm[8];
    T acc = mr_identity<FLOAT>(op);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) acc = mr_fold<FLOAT>(partial[i], acc, op);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = mr_fold<FLOAT>(acc, __shfl_xor_sync(0xffffffffu, acc, o), op);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        T v = sm[0];
        for (int i = 1; i < 8; ++i) v = mr_fold<FLOAT>(v, sm[i], op);
        *out = v;
    }
}
__global__ void field_out_kernel(const uint8_t* __restrict__ cols, uint32_t stride, uint32_t word, const uint32_t* __restrict__ cells, uint32_t cbase,
                                 uint64_t n, int offset, uint32_t fsize, uint8_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (cells[i] == 0xffffffffu) { for (uint32_t b = 0; b < fsize; ++b) out[i * fsize + b] = 0; return; }     // a cell of another rank (joined afterwards)
    const uint32_t s = cells[i] - cbase;
    for (uint32_t b = 0; b < fsize; ++b) {
        const uint32_t p = offset + b, c = p / word;
        out[i * fsize + b] = cols[(size_t)c * stride * word + (size_t)s * word + (p - c * word)];
    }
}
__global__ void raster_num_edges_kernel(const uint32_t* __restrict__ cells, uint64_t n, const uint32_t* __restrict__ off, const uint32_t* __restrict__ cnt,
                                        uint32_t rows, uint32_t row_shift, long long* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = cells[i] - row_shift;
    out[i] = (cells[i] != 0xffffffffu && r < rows) ? (off ? (long long)(off[r + 1] - off[r]) : (long long)cnt[r]) : 0;
}

// ------------------------------------------------------------------------------------------------------
struct AgentStore {
    std::string name;
    uint32_t size = 0, hints = 0, word = 0, ncols = 0;
    bool immortal = false, independent = false, stateless = false;
    uint32_t cap = 0;
    uint8_t* state[2] = {nullptr, nullptr};
    uint8_t* died[2] = {nullptr, nullptr};
    int cur = 0;                 // state[cur] / died[cur] is the read buffer
    uint32_t* reuse = nullptr;
    uint32_t reuse_cap = 0, n_reuse = 0;
    uint32_t nslots = 0;         // length(read.state) == length(write.state) between applies
    uint64_t nextid = 1;
    bool write_stale = false;    // the write buffer does not hold the current states (after a swap)
    bool writeable = false, prepared = false;
    int64_t last_change = 0;
    uint32_t births = 0;         // births of the running apply
    uint64_t uoffset = 0;        // offset of this rank's block in the global per-agent uniform table
    // multi-GPU: ghost segment [cap, cap + nghost) mirrors remote agents referenced by local edges
    uint32_t gcap = 0, nghost = 0;
    uint64_t* ghost_ids = nullptr;            // device, ascending (rank-major)
    std::vector<uint32_t> ghost_off;          // [nranks + 1] ghost range per owner rank
    uint32_t* send_slots = nullptr;           // device: local slots requested by the peers, grouped by peer
    std::vector<uint32_t> send_off;           // [nranks + 1]
    // peer-memory halo (one kernel packs and pushes over NVLink): the peers' state buffers mapped into this process
    struct PeerMap {
        std::vector<uint8_t*> base[2];        // [nranks] peer r's state[0] / state[1] (nullptr for this rank)
        std::vector<uint32_t> stride, ghost0; // peer r's column stride; first ghost slot of MY agents in peer r's buffers
        std::vector<void*> opened;            // cudaIpcOpenMemHandle results to close on refresh
        uint64_t sig = 0;                     // layout signature the map was exchanged for (0 = never)
        bool ok = false;
        bool check = true;                    // some rank's layout may have changed since the exchange: the next halo re-checks (collective)
        // the halo travels in `ng` phases (the same count on every rank): phase j carries the j-th part of every owner's range of a
        // receiver's ghost segment (ghost_part() slots of it), so every rank sends in every phase; the receiver's ghost key block j is
        // the union of those parts and can be swept while the later phases are still on the wire
        uint32_t ng = 1;
    } peers;
    // slots per phase of a range of `len` ghosts travelling in `ng` phases (the same arithmetic on the sender and on the receiver)
    // The first part is smaller (VB_HALO_FIRST, 0.3 of the range when there are several phases): the first sweep that needs ghosts
    // waits for it with nothing to overlap, the later parts travel beside sweeps.
    static uint32_t ghost_first(uint32_t len, uint32_t ng) {
        static const double frac = getenv("VB_HALO_FIRST") ? std::min(1.0, std::max(0.01, atof(getenv("VB_HALO_FIRST")))) : 0.3;
        if (ng <= 1) return len;
        return std::min<uint32_t>(len, ((uint32_t)std::ceil((double)len * std::min(frac, 1.0 / ng)) + 63u) & ~63u);     // never more than an equal share
    }
    static uint32_t ghost_rest(uint32_t len, uint32_t ng) {     // slots of each later part
        if (ng <= 1) return 1;
        const uint32_t rest = len - ghost_first(len, ng);
        return (std::max<uint32_t>((rest + ng - 2) / (ng - 1), 1u) + 63u) & ~63u;
    }
    static void ghost_part_range(uint32_t len, uint32_t ng, uint32_t j, uint32_t& lo, uint32_t& hi) {     // part j of a range of len slots
        const uint64_t f = ghost_first(len, ng), r = ghost_rest(len, ng);
        if (j == 0) { lo = 0; hi = (uint32_t)f; return; }
        lo = (uint32_t)std::min<uint64_t>(f + (uint64_t)(j - 1) * r, len); hi = (uint32_t)std::min<uint64_t>(f + (uint64_t)j * r, len);
        if (j + 1 == ng) hi = len;
    }
    std::vector<cudaEvent_t> ev_phase;        // recorded on the halo stream after phase j has landed everywhere
    uint32_t halo_pending = 0, halo_waited = 0;   // phases started by this apply / already awaited by the main stream
    uint8_t* send_buf = nullptr;              // packed states for the halo exchange
    bool halo_dirty = true;
    uint32_t stride() const { return cap + gcap; }
    uint8_t* rstate() const { return state[cur]; }
    uint8_t* wstate() const { return independent ? state[cur] : state[cur ^ 1]; }
    uint8_t* rdied() const { return died[cur]; }
    uint8_t* wdied() const { return died[cur ^ 1]; }
};

struct RawChunk { uint64_t* to = nullptr; uint64_t* from = nullptr; uint8_t* st = nullptr; uint64_t n = 0; };

struct EdgeStore {
    std::string name;
    uint32_t size = 0, hints = 0, word = 0, ncols = 0;
    int32_t target = 0;
    uint64_t size_hint = 0;
    bool stateless = false, ignorefrom = false, singleedge = false, singletype = false;
    uint8_t kind = vb::KIND_CSR;
    // read container
    uint32_t* off = nullptr; uint32_t rows = 0;
    uint32_t* src = nullptr; uint8_t* st = nullptr; uint32_t nnz = 0, st_cap = 0;
    uint32_t* cnt = nullptr;
    // write container of the running apply
    uint32_t* log_to = nullptr; uint32_t* log_from = nullptr; uint8_t* log_st = nullptr;
    uint32_t log_n = 0, log_cap = 0;
    uint32_t* wcnt = nullptr; uint32_t rows_w = 0;
    // remove_edges! records of the running apply
    uint32_t* rm_row = nullptr; uint32_t* rm_from = nullptr; uint32_t* rm_mark = nullptr; uint32_t rm_n = 0, rm_cap = 0;
    uint64_t* rm_to64 = nullptr; uint64_t* rm_from64 = nullptr;   // multi-GPU: AgentIDs of the records that leave the rank (same positions)
    // connect_raster_neighbors! kept implicit (KIND_STENCIL on device) until something needs explicit rows
    bool implicit_stencil = false; int st_raster = -1; int st_metric = 0; double st_distance = 0; bool st_periodic = true;
    std::vector<int8_t> st_off_host; int8_t* st_off = nullptr; int st_n = 0; int st_reach = 0; uint32_t st_slot0 = 0;
    // multi-GPU: appended edges that leave the rank (AgentIDs), exchanged by transmit_edges
    uint64_t* rlog_to = nullptr; uint64_t* rlog_from = nullptr; uint8_t* rlog_st = nullptr; uint32_t* rlog_dst = nullptr; uint32_t rlog_n = 0, rlog_cap = 0;
    bool ordered_log = false;   // count/flag container whose appends must be ordered this apply (a transition removes edges of it)
    // raw adds from the host API (AgentIDs, AoS states) awaiting translation, in call order
    std::vector<uint64_t> h_to, h_from; std::vector<uint8_t> h_st;
    std::vector<RawChunk> chunks;
    uint64_t raw_n = 0;
    // :SingleEdge double-add detection at add time for host adds (EdgeMethods.jl:267-293)
    std::unordered_map<uint64_t, std::pair<uint64_t, std::vector<uint8_t>>> single_seen;
    bool readable = false, writeable = false, add_existing = false;
    int64_t last_change = 0;
    // degree-binning cache: slots of agent type `heavy_type` whose row has >= HEAVY_MIN entries
    uint32_t* heavy_rows = nullptr; uint32_t heavy_n = 0; int heavy_type = 0; uint32_t heavy_min = 0; uint64_t heavy_version = ~0ull; uint64_t version = 0;
    // source-blocked view of the CSR columns for reduce transitions (include/vahana_device.cuh): per block b of the
    // source type's slots an offset array boff + b * rpad ([n + 1] positions into bsrc) and the rows' sources inside that block
    struct Blocked {
        uint32_t* boff = nullptr; uint32_t* bsrc = nullptr; uint32_t* heavy_bits = nullptr; uint8_t* acc = nullptr;
        uint8_t* key = nullptr; uint32_t key_n = 0;   // prefiltered sweeps: one key byte per slot of the source type (rebuilt by every apply)
        // segmented form (prefiltered sweeps): boff / acc are indexed by segment, rpad = padded segment count
        bool segmented = false;
        uint32_t nseg = 0, seg_len = 0, nhub = 0;
        // blocks: nbl of bsize slots over the local slots [0, lcap); then ng ghost blocks, block nbl + j = the j-th part of every owner's
        // range of the ghost segment (goff = the ghost table's ranges the view was built for); entries of ghost blocks are relative to lcap
        uint32_t lcap = 0, nbl = 0, ng = 0, gfirst = 0;   // ghost part j lives in block gfirst + j (gfirst = nbl - 1: it shares the last local block)
        bool absolute = false;                             // entries hold the slot itself, not the slot relative to the block
        std::vector<uint32_t> goff;
        uint32_t block_first(uint32_t b) const { return absolute ? 0u : (b < nbl ? b * bsize : lcap); }
        // local block b: the slots that can hold a source (the last local block takes the tail up to `used`, the slots in use)
        uint32_t local_block_slots(uint32_t b, uint32_t used) const {
            const uint32_t f = b * bsize, lend = std::min(lcap, used);
            const uint32_t lim = b + 1 == nbl ? lend : std::min(lend, f + bsize);
            return lim > f ? lim - f : 0u;
        }
        uint32_t* seg_row = nullptr;                  // [nseg] called row | 0x80000000 when the row has several segments
        uint32_t* hub_rows = nullptr; uint32_t* hub_seg = nullptr;   // rows with several segments, [2 nhub] their segment ranges [first, end)
        uint32_t nb = 0, bsize = 0, rpad = 0, n = 0, acc_bytes = 0, heavy_min = 0;
        std::vector<uint32_t> bstart;                 // position in bsrc where each block's entries start (nb + 1 values)
        std::vector<uint32_t*> arows;                 // per block: ascending list of the rows that own an entry in it
        std::vector<uint32_t*> aoff;                  //   and the positions of their entries ([acount + 1], contiguous in bsrc)
        std::vector<uint32_t> acount;
        int called = 0, source = 0;
        uint64_t version = ~0ull, epoch = ~0ull;      // container version / sim layout epoch the view was built for
        uint64_t seen_version = ~0ull; uint32_t seen = 0;   // how many applies found the same container (static network => worth building)
        bool refused = false;                          // the build found a source of another type: stay on the direct path
    } blk;
    // prefilter policy: the sampled pass rate decides between prefiltered and unfiltered sweeps (checked every 16th apply)
    bool pf_off = false; uint32_t pf_check_in = 0; double pf_rate = -1.0;
    bool has_src() const { return !ignorefrom; }
    bool has_state() const { return !stateless && size > 0; }
};

struct RasterStore {
    std::string name;
    std::vector<int64_t> dims;
    int type = 0;
    std::vector<uint64_t> ids;    // host copy (column-major)
    uint32_t* cells = nullptr;    // device composite indices (rebuilt when bases change); 0xffffffff = a cell of another rank
    uint64_t* cell_ids = nullptr; // device copy of `ids`, kept when the cells are spread over the ranks (vb_set_raster)
    bool distributed = false;     // read-outs join the ranks (collective)
};

}  // namespace

struct vb_sim {
    std::string name;
    std::vector<AgentStore> agents;   // index = type id - 1
    std::vector<EdgeStore> edges;
    std::vector<RasterStore> rasters;
    std::vector<uint8_t> params;
    bool initialized = false, intransition = false;
    int64_t num_transitions = 0;
    bool asserts_enabled = true, check_readable = true;
    bool all_immortal = true;
    uint32_t rank = 0;
    uint32_t base[vb::MAX_AGENT_TYPES + 2] = {0};
    vb::DeviceSim h_ds;                       // host copy of the view; passed by value in every transition launch
    uint32_t* d_error = nullptr;
    uint32_t* d_scalars = nullptr;          // small device scratch for totals
    unsigned long long* d_stats = nullptr;
    // stats of the last apply
    double ms_rw = 0, ms_fin = 0;
    bool stats_pending = false, times_pending = false;
    void fetch_times() {
        if (!times_pending) return;
        times_pending = false;
        float m0 = 0, m1 = 0;
        if (cudaEventSynchronize(ev[2]) == cudaSuccess) { cudaEventElapsedTime(&m0, ev[0], ev[1]); cudaEventElapsedTime(&m1, ev[1], ev[2]); }
        cudaGetLastError();
        ms_rw = m0; ms_fin = m1;
    }
    uint64_t st_edges_read = 0, st_edges_appended = 0, st_agents_called = 0, st_launches = 0;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t evk[2] = {nullptr, nullptr};   // around the transition kernels themselves
    double ms_kernel = 0;

    AgentStore& A(int t) { if (t < 1 || t > (int)agents.size()) throw ArgError("unknown agent type id"); return agents[t - 1]; }
    EdgeStore& E(int e) { if (e < 0 || e >= (int)edges.size()) throw ArgError("unknown edge type index"); return edges[e]; }
    void mayassert(bool c, const char* m) const { if (asserts_enabled && !c) throw AssertionError(m); }
    uint32_t total_slots() const { return base[agents.size() + 1]; }
    uint32_t rows_of(const EdgeStore& e) const { return e.singletype ? agents[e.target - 1].cap : base[agents.size() + 1]; }
    uint32_t row_base_of(const EdgeStore& e) const { return e.singletype ? base[e.target] : 0; }

    ~vb_sim();
    void compute_bases(uint32_t* out) const;
    void ensure_agent_cap(int t, uint64_t need, uint32_t ghost_need = 0, bool do_rebase = true);
    void build_ghosts(const uint64_t* const* extra = nullptr, const uint64_t* extra_n = nullptr, int n_extra = 0);
    void halo_exchange(int t);                // starts the exchange (peer-memory path: on the halo stream, in phases); halo_wait() joins
    void halo_wait(int t, uint32_t phases);   // the main stream waits until the first `phases` phases of type t's halo have landed
    void halo_wait_all();
    bool refresh_peer_map(int t);
    // some rank's buffers / ghost tables may have changed (called at the same events on every rank): the next halo exchange of every
    // agent type re-checks the peer maps
    void mark_peer_check() { for (auto& a : agents) a.peers.check = true; }
    uint64_t halo_bytes = 0;   // bytes received by the last apply's halo exchanges
    cudaEvent_t ev_halo[2] = {nullptr, nullptr};   // on the halo stream: behind the entry barrier / behind the last phase's barrier
    bool halo_timed = false;
    double blk_block_mb = -1, blk_min_mb = -1; int blk_eager = -1;   // vb_set_read_blocking (negative = environment / default)
    int blk_prefilter = -1;                                          // vb_set_read_prefilter (negative = environment / default)
    bool prefilter_on(const vb::TransitionInfo* ti) const;
    uint32_t last_blocked_nb = 0;   // source blocks swept by the last apply's read phase (0 = direct path)
    bool last_prefiltered = false;  // those sweeps gathered keys (prefilter) instead of states
    double last_pass_rate = -1.0;   // last estimate of the prefilter's pass rate (vb_last_pass_rate)
    uint64_t layout_epoch = 0; // bumped whenever stored composite indices are renumbered (rebase): invalidates blocked views
    bool ensure_blocked(int ei, const vb::TransitionInfo* ti, int C, uint32_t n, uint32_t heavy_min, const vb::LaunchArgs* la = nullptr, uint64_t seed = 0);
    void rebase(const uint32_t* old_base, const uint32_t* old_lcap = nullptr, const uint32_t* const* remap = nullptr);
    void exchange_ghost_requests();
    void transmit_edges(int e);
    void transmit_removes(int e);
    void ensure_rm(EdgeStore& e, uint64_t need);
    void upload_view(uint64_t seed);
    void check_device_error(const char* where);
    void flush_raw(int e);
    void merge_pending(int e);
    void merge_all_pending() { for (size_t e = 0; e < edges.size(); ++e) merge_pending((int)e); }
    void ensure_log(EdgeStore& es, uint64_t need);
    void build_container(int e, bool add_existing);
    void apply_removes(int e);
    void materialize_stencil(int e);
    void emit_raster_edges(int e, int raster, double distance, int metric, bool periodic, const void* st);
    std::vector<uint32_t> stencil_row_host(const EdgeStore& e, uint64_t lin) const;
    uint64_t stencil_total(const EdgeStore& e) const;
    void purge_dead(const uint8_t* dead);
    uint64_t edge_total(int e, bool write);
};

namespace {

void free_agent(AgentStore& a) {
    for (void* q : a.peers.opened) cudaIpcCloseMemHandle(q);
    cudaGetLastError();
    a.peers = AgentStore::PeerMap{};
    for (cudaEvent_t e : a.ev_phase) cudaEventDestroy(e);
    a.ev_phase.clear();
    dfree(a.state[0]); if (a.state[1] != a.state[0]) dfree(a.state[1]);
    dfree(a.died[0]); dfree(a.died[1]); dfree(a.reuse); dfree(a.ghost_ids); dfree(a.send_slots); dfree(a.send_buf);
    a.state[0] = a.state[1] = a.died[0] = a.died[1] = nullptr; a.reuse = nullptr; a.ghost_ids = nullptr; a.send_slots = nullptr; a.send_buf = nullptr;
}
void free_blocked(EdgeStore& e) {
    for (auto p : e.blk.arows) dfree(p);
    for (auto p : e.blk.aoff) dfree(p);
    dfree(e.blk.boff); dfree(e.blk.bsrc); dfree(e.blk.heavy_bits); dfree(e.blk.acc); dfree(e.blk.key);
    dfree(e.blk.seg_row); dfree(e.blk.hub_rows); dfree(e.blk.hub_seg);
    e.blk = EdgeStore::Blocked{};
}
void free_edge_read(EdgeStore& e) {
    ++e.version;
    if (e.blk.boff) free_blocked(e);
    dfree(e.off); dfree(e.src); dfree(e.st); dfree(e.cnt);
    e.off = e.src = e.cnt = nullptr; e.st = nullptr; e.nnz = e.st_cap = 0; e.rows = 0;
}
void free_edge_log(EdgeStore& e) {
    dfree(e.log_to); dfree(e.log_from); dfree(e.log_st); dfree(e.wcnt); dfree(e.rm_row); dfree(e.rm_from); dfree(e.rm_mark); dfree(e.rm_to64); dfree(e.rm_from64);
    e.log_to = e.log_from = e.wcnt = nullptr; e.log_st = nullptr; e.log_n = e.log_cap = 0; e.rows_w = 0;
    e.rm_row = e.rm_from = e.rm_mark = nullptr; e.rm_to64 = e.rm_from64 = nullptr; e.rm_n = e.rm_cap = 0;
    dfree(e.rlog_to); dfree(e.rlog_from); dfree(e.rlog_st); dfree(e.rlog_dst);
    e.rlog_to = e.rlog_from = nullptr; e.rlog_st = nullptr; e.rlog_dst = nullptr; e.rlog_n = e.rlog_cap = 0;
}
void free_chunks(EdgeStore& e) {
    for (auto& c : e.chunks) { dfree(c.to); dfree(c.from); dfree(c.st); }
    e.chunks.clear();
}

}  // namespace

vb_sim::~vb_sim() {
    for (auto& a : agents) free_agent(a);
    for (auto& e : edges) { free_edge_read(e); free_edge_log(e); free_chunks(e); dfree(e.st_off); dfree(e.heavy_rows); free_blocked(e); }
    for (auto& r : rasters) { dfree(r.cells); dfree(r.cell_ids); }
    dfree(d_error); dfree(d_scalars); dfree(d_stats);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    for (auto& e : evk) if (e) cudaEventDestroy(e);
    for (auto& e : ev_halo) if (e) cudaEventDestroy(e);
}

void vb_sim::compute_bases(uint32_t* out) const {
    uint64_t run = 0;
    out[0] = 0;
    for (size_t t = 1; t <= agents.size(); ++t) { out[t] = (uint32_t)run; run += agents[t - 1].stride(); }
    if (run >= 0xffffffffull) throw ArgError("more than 2^32-1 agent slots on one rank are not supported by the 32-bit composite index");
    out[agents.size() + 1] = (uint32_t)run;
}

// grow the buffers of agent type t to hold `need` slots; composite bases of later types shift (rebase)
void vb_sim::ensure_agent_cap(int t, uint64_t need, uint32_t ghost_need, bool do_rebase) {
    AgentStore& a = A(t);
    if (need <= a.cap && ghost_need <= a.gcap) return;
    uint64_t ncap = a.cap;
    if (need > a.cap) { ncap = std::max<uint64_t>(need, (uint64_t)a.cap + a.cap / 2); ncap = (ncap + 255) / 256 * 256; }
    const uint32_t ngcap = std::max(a.gcap, ghost_need);
    const uint64_t nstride = ncap + ngcap;
    if (ncap >= 0xffffffffull) throw ArgError("agent type too large for 32-bit slots");
    uint32_t old_base[vb::MAX_AGENT_TYPES + 2];
    std::memcpy(old_base, base, sizeof(old_base));
    const bool had = a.cap > 0;
    const int nb = a.independent ? 1 : 2;
    if (a.size) {
        for (int b = 0; b < nb; ++b) {
            uint8_t* n = (uint8_t*)g_pool.alloc(nstride * a.size);
            if (had) {   // the local part moves; ghosts are refilled by the next halo exchange
                vbp::soa_copy_kernel<<<nblk((uint64_t)a.cap * a.size), 256, 0, g_stream>>>(a.state[b], a.stride(), n, nstride, a.cap, a.ncols, a.word, 0, 0);
                LAUNCH_CHECK();
            }
            dfree(a.state[b]);
            a.state[b] = n;
        }
        if (a.independent) a.state[1] = a.state[0];
    }
    if (!a.immortal) {
        for (int b = 0; b < 2; ++b) {
            uint8_t* n = (uint8_t*)g_pool.alloc(nstride);
            CK(cudaMemsetAsync(n, 0, nstride, g_stream));
            if (had) CK(cudaMemcpyAsync(n, a.died[b], a.cap, cudaMemcpyDeviceToDevice, g_stream));   // local flags; ghosts count as alive
            dfree(a.died[b]);
            a.died[b] = n;
        }
        uint32_t* r = dalloc<uint32_t>(ncap);
        if (a.n_reuse) CK(cudaMemcpyAsync(r, a.reuse, (size_t)a.n_reuse * 4, cudaMemcpyDeviceToDevice, g_stream));
        dfree(a.reuse);
        a.reuse = r;
        a.reuse_cap = (uint32_t)ncap;
    }
    uint32_t old_lcap[vb::MAX_AGENT_TYPES + 1] = {0};
    for (size_t k = 1; k <= agents.size(); ++k) old_lcap[k] = agents[k - 1].cap;
    a.cap = (uint32_t)ncap;
    a.gcap = ngcap;
    a.halo_dirty = true;
    compute_bases(base);
    if (do_rebase) rebase(old_base, old_lcap);
}

// after capacities changed: remap every stored composite index (CSR columns, rows of containers that are
// keyed by all agent types, append logs, raster cell tables)
void vb_sim::rebase(const uint32_t* old_base, const uint32_t* old_lcap, const uint32_t* const* remap) {
    RebaseArgs ra{};
    std::memcpy(ra.old_base, old_base, sizeof(ra.old_base));
    std::memcpy(ra.new_base, base, sizeof(ra.new_base));
    ra.ntypes = (uint32_t)agents.size();
    bool same = true;
    for (size_t t = 0; t <= agents.size() + 1; ++t) same &= ra.old_base[t] == ra.new_base[t];
    for (size_t t = 1; t <= agents.size(); ++t) {
        ra.new_lcap[t] = agents[t - 1].cap;
        ra.old_lcap[t] = old_lcap ? old_lcap[t] : agents[t - 1].cap;
        ra.remap[t] = remap ? remap[t] : nullptr;
        same &= ra.old_lcap[t] == ra.new_lcap[t] && !ra.remap[t];
    }
    if (!same) ++layout_epoch;
    for (auto& e : edges) {
        if (!same && e.src && e.nnz) { rebase_values_kernel<<<nblk(e.nnz), 256, 0, g_stream>>>(e.src, e.nnz, ra); LAUNCH_CHECK(); }
        if (!same && e.log_from && e.log_n) { rebase_values_kernel<<<nblk(e.log_n), 256, 0, g_stream>>>(e.log_from, e.log_n, ra); LAUNCH_CHECK(); }
        if (!same && e.rm_from && e.rm_n) { rebase_values_kernel<<<nblk(e.rm_n), 256, 0, g_stream>>>(e.rm_from, e.rm_n, ra); LAUNCH_CHECK(); }
        if (!same && !e.singletype && e.rm_row && e.rm_n) { rebase_values_kernel<<<nblk(e.rm_n), 256, 0, g_stream>>>(e.rm_row, e.rm_n, ra); LAUNCH_CHECK(); }
        const uint32_t nrows = rows_of(e);
        if (!e.singletype) {
            if (!same && e.log_to && e.log_n) { rebase_values_kernel<<<nblk(e.log_n), 256, 0, g_stream>>>(e.log_to, e.log_n, ra); LAUNCH_CHECK(); }
            if (e.off && nrows != e.rows) {
                uint32_t* noff = dalloc<uint32_t>((size_t)nrows + 1);
                rebase_rows_kernel<<<nblk((uint64_t)nrows + 1), 256, 0, g_stream>>>(e.off, e.rows, noff, nrows, ra); LAUNCH_CHECK();
                dfree(e.off); e.off = noff; e.rows = nrows;
            }
            if (e.cnt && nrows != e.rows) {
                uint32_t* nc = dalloc<uint32_t>((size_t)nrows + 1);
                rebase_cnt_kernel<<<nblk(nrows), 256, 0, g_stream>>>(e.cnt, e.rows, nc, nrows, ra); LAUNCH_CHECK();
                dfree(e.cnt); e.cnt = nc; e.rows = nrows;
            }
            if (e.wcnt && nrows != e.rows_w) {
                uint32_t* nc = dalloc<uint32_t>((size_t)nrows + 1);
                rebase_cnt_kernel<<<nblk(nrows), 256, 0, g_stream>>>(e.wcnt, e.rows_w, nc, nrows, ra); LAUNCH_CHECK();
                dfree(e.wcnt); e.wcnt = nc; e.rows_w = nrows;
            }
        } else {
            // rows are the slots of the target type only: extend with empty rows
            if (e.off && nrows > e.rows) {
                uint32_t* noff = dalloc<uint32_t>((size_t)nrows + 1);
                CK(cudaMemcpyAsync(noff, e.off, ((size_t)e.rows + 1) * 4, cudaMemcpyDeviceToDevice, g_stream));
                vbp::fill_u32_kernel<<<nblk(nrows - e.rows), 256, 0, g_stream>>>(noff + e.rows + 1, nrows - e.rows, e.nnz); LAUNCH_CHECK();
                dfree(e.off); e.off = noff; e.rows = nrows;
            }
            if (e.cnt && nrows > e.rows) {
                uint32_t* nc = dalloc<uint32_t>((size_t)nrows + 1);
                CK(cudaMemsetAsync(nc, 0, ((size_t)nrows + 1) * 4, g_stream));
                CK(cudaMemcpyAsync(nc, e.cnt, (size_t)e.rows * 4, cudaMemcpyDeviceToDevice, g_stream));
                dfree(e.cnt); e.cnt = nc; e.rows = nrows;
            }
            if (e.wcnt && nrows > e.rows_w) {
                uint32_t* nc = dalloc<uint32_t>((size_t)nrows + 1);
                CK(cudaMemsetAsync(nc, 0, ((size_t)nrows + 1) * 4, g_stream));
                CK(cudaMemcpyAsync(nc, e.wcnt, (size_t)e.rows_w * 4, cudaMemcpyDeviceToDevice, g_stream));
                dfree(e.wcnt); e.wcnt = nc; e.rows_w = nrows;
            }
        }
    }
    for (auto& r : rasters) {
        if (!r.cells) r.cells = dalloc<uint32_t>(r.ids.size());
        uint64_t* tmp = dalloc<uint64_t>(r.ids.size());
        CK(cudaMemcpyAsync(tmp, r.ids.data(), r.ids.size() * 8, cudaMemcpyHostToDevice, g_stream));
        raster_cells_kernel<<<nblk(r.ids.size()), 256, 0, g_stream>>>(tmp, r.ids.size(), r.cells, ra, rank); LAUNCH_CHECK();
        CK(cudaStreamSynchronize(g_stream));
        if (r.distributed && !r.cell_ids) r.cell_ids = tmp; else dfree(tmp);
    }
}

void vb_sim::upload_view(uint64_t seed) {
    vb::DeviceSim& h = h_ds;
    std::memset(&h, 0, sizeof(h));
    for (size_t t = 1; t <= agents.size(); ++t) {
        const AgentStore& a = agents[t - 1];
        vb::AgentView& v = h.agents[t];
        v.state_r = a.rstate(); v.state_w = a.wstate();
        v.died_r = a.immortal ? nullptr : a.rdied(); v.died_w = a.immortal ? nullptr : a.wdied();
        v.reuse = a.reuse; v.cap = a.stride(); v.lcap = a.cap; v.nghost = a.nghost; v.ghost_ids = a.ghost_ids; v.nslots_r = a.nslots;
        v.n_reuse = a.n_reuse; v.next0 = (uint32_t)(a.nextid - 1);
        v.size = a.size; v.word = a.word ? a.word : 1; v.ncols = a.ncols; v.uoffset = a.uoffset;
        v.immortal = a.immortal; v.independent = a.independent; v.readable = a.prepared; v.writeable = a.writeable;
    }
    int n_tab = 0;
    for (size_t i = 0; i < edges.size(); ++i) {
        const EdgeStore& e = edges[i];
        vb::EdgeView& v = h.edges[i];
        v.off = e.off; v.src = e.src; v.st = e.st; v.cnt = e.cnt; v.rows = e.rows; v.st_cap = e.st_cap;
        v.log_to = e.log_to; v.log_from = e.log_from; v.log_st = e.log_st; v.wcnt = e.wcnt; v.log_cap = e.log_cap; v.rows_w = e.rows_w;
        v.rm_row = e.rm_row; v.rm_from = e.rm_from; v.rm_mark = e.rm_mark; v.rm_to64 = e.rm_to64; v.rm_from64 = e.rm_from64;
        v.rlog_to = e.rlog_to; v.rlog_from = e.rlog_from; v.rlog_st = e.rlog_st; v.rlog_dst = e.rlog_dst; v.rlog_cap = e.rlog_cap;
        v.size = e.size; v.word = e.word ? e.word : 1; v.ncols = e.ncols; v.target = e.singletype ? e.target : 0;
        v.hints = (uint8_t)e.hints; v.kind = e.implicit_stencil ? (uint8_t)vb::KIND_STENCIL : e.kind; v.readable = e.readable; v.writeable = e.writeable;
        v.st_tab = 0; v.st_n = e.st_n; v.st_raster = e.st_raster; v.st_slot0 = e.st_slot0; v.st_periodic = e.st_periodic; v.st_reach = (uint8_t)e.st_reach;
        if (e.implicit_stencil) {
            v.rows = 0xffffffffu;
            v.st_tab = n_tab;
            vb::StencilTab& tb = h.stencils[n_tab++];
            const RasterStore& r = rasters[e.st_raster];
            for (int si = 0; si < e.st_n; ++si) {
                long long lin = 0, stride = 1;
                for (size_t k = 0; k < r.dims.size(); ++k) {
                    tb.off[si][k] = e.st_off_host[(size_t)si * vb::MAX_RASTER_DIMS + k];
                    lin += (long long)tb.off[si][k] * stride; stride *= r.dims[k];
                }
                tb.lin[si] = (int32_t)lin;
            }
        }
    }
    for (size_t i = 0; i < rasters.size(); ++i) {
        vb::RasterView& v = h.rasters[i];
        v.cells = rasters[i].cells; v.cell_ids = rasters[i].cell_ids; v.ndims = (int)rasters[i].dims.size(); v.type = rasters[i].type;
        uint64_t st = 1;
        for (int k = 0; k < vb::MAX_RASTER_DIMS; ++k) { v.dim32[k] = 1; v.stride32[k] = 0; }
        for (size_t k = 0; k < rasters[i].dims.size(); ++k) { v.dims[k] = rasters[i].dims[k]; v.dim32[k] = (uint32_t)rasters[i].dims[k]; v.stride32[k] = (uint32_t)st; st *= (uint64_t)rasters[i].dims[k]; }
        v.ncells = (uint32_t)std::min<uint64_t>(st, 0xffffffffull);
    }
    std::memcpy(h.base, base, sizeof(h.base));
    h.n_agent_types = (uint32_t)agents.size(); h.n_edge_types = (uint32_t)edges.size(); h.n_rasters = (uint32_t)rasters.size();
    h.rank = rank; h.nranks = (uint32_t)g_nranks; h.check = asserts_enabled && check_readable; h.error = d_error; h.seed = seed;
    if (!params.empty()) std::memcpy(h.params, params.data(), params.size());
}

void vb_sim::check_device_error(const char* where) {
    uint32_t err = 0;
    CK(cudaMemcpyAsync(&err, d_error, 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    if (!err) return;
    CK(cudaMemsetAsync(d_error, 0, 4, g_stream));
    std::string m = std::string(where) + ":";
    if (err & vb::DERR_EDGE_NOT_READABLE) m += " an edge type was accessed that is not in the `read` argument;";
    if (err & vb::DERR_AGENT_NOT_READABLE) m += " an agent type was accessed that is not in the `read` argument;";
    if (err & vb::DERR_AGENT_TYPE_MISMATCH) m += " the id of an agent does not match the given type;";
    if (err & vb::DERR_AGENT_DIED) m += " agentstate was requested for an agent that has been removed;";
    if (err & vb::DERR_IMMORTAL_DIED) m += " `nothing` was returned for an :Immortal agent;";
    if (err & vb::DERR_ACCESSOR_UNAVAILABLE) m += " an accessor was used that is not defined for the edge type's hint combination;";
    if (err & vb::DERR_BAD_ID) m += " an agent id does not name an existing agent;";
    if (err & vb::DERR_EDGE_NOT_DECLARED) m += " add_edge/add_agent on a type the transition did not declare in EdgeWrites/AgentWrites;";
    if (err & vb::DERR_SINGLETYPE_MISMATCH) m += " :SingleType edge used with an agent of another type;";
    if (err & vb::DERR_RASTER_POS) m += " raster position out of range;";
    if (err & vb::DERR_INDEX) m += " neighbour index out of range;";
    if (err & vb::DERR_MODEL_ASSERT) m += " an assertion inside the transition function failed (ctx.require);";
    if (err & vb::DERR_REMOTE) m += " the edges of an agent of another rank were read, or an edge of another rank was named outside of a transition (edges live on the rank of their target);";
    throw AssertionError(m);
}

void vb_sim::ensure_log(EdgeStore& es, uint64_t need) {
    if (need <= es.log_cap) return;
    if (need >= 0xffffffffull) throw ArgError("more than 2^32-1 edges of one type on one rank are not supported");
    uint64_t ncap = std::max<uint64_t>(need, (uint64_t)es.log_cap + es.log_cap / 2);
    ncap = (ncap + 1023) / 1024 * 1024;
    uint32_t* nt = dalloc<uint32_t>(ncap);
    if (es.log_n) CK(cudaMemcpyAsync(nt, es.log_to, (size_t)es.log_n * 4, cudaMemcpyDeviceToDevice, g_stream));
    dfree(es.log_to); es.log_to = nt;
    if (es.has_src()) {
        uint32_t* nf = dalloc<uint32_t>(ncap);
        if (es.log_n) CK(cudaMemcpyAsync(nf, es.log_from, (size_t)es.log_n * 4, cudaMemcpyDeviceToDevice, g_stream));
        dfree(es.log_from); es.log_from = nf;
    }
    if (es.has_state()) {
        uint8_t* ns = (uint8_t*)g_pool.alloc(ncap * es.size);
        if (es.log_n) { vbp::soa_copy_kernel<<<nblk((uint64_t)es.log_n * es.size), 256, 0, g_stream>>>(es.log_st, es.log_cap, ns, ncap, es.log_n, es.ncols, es.word, 0, 0); LAUNCH_CHECK(); }
        dfree(es.log_st); es.log_st = ns;
    }
    es.log_cap = (uint32_t)ncap;
}

// host-staged raw adds -> one device chunk
void vb_sim::flush_raw(int ei) {
    EdgeStore& e = E(ei);
    const uint64_t n = e.h_to.size();
    if (!n) return;
    RawChunk c;
    c.n = n;
    c.to = dalloc<uint64_t>(n);
    CK(cudaMemcpyAsync(c.to, e.h_to.data(), n * 8, cudaMemcpyHostToDevice, g_stream));
    if (e.has_src()) { c.from = dalloc<uint64_t>(n); CK(cudaMemcpyAsync(c.from, e.h_from.data(), n * 8, cudaMemcpyHostToDevice, g_stream)); }
    if (e.has_state()) { c.st = (uint8_t*)g_pool.alloc(n * e.size); CK(cudaMemcpyAsync(c.st, e.h_st.data(), n * e.size, cudaMemcpyHostToDevice, g_stream)); }
    CK(cudaStreamSynchronize(g_stream));
    e.h_to.clear(); e.h_from.clear(); e.h_st.clear();
    e.chunks.push_back(c);
}

// translate all raw chunks into the append log and fold them into the container as an add_existing write:
// this is what makes add_edge! outside of transitions (init phase; test hacks after finish_init!) visible.
void vb_sim::merge_pending(int ei) {
    EdgeStore& e = E(ei);
    flush_raw(ei);
    if (e.chunks.empty()) return;
    uint64_t total = 0;
    for (auto& c : e.chunks) total += c.n;
    const uint32_t rows = rows_of(e);
    if (e.kind == vb::KIND_CSR) ensure_log(e, e.log_n + total);
    else if (!e.wcnt) {
        e.wcnt = dalloc<uint32_t>((size_t)rows + 1); e.rows_w = rows;
        CK(cudaMemsetAsync(e.wcnt, 0, ((size_t)rows + 1) * 4, g_stream));
        if (e.cnt) CK(cudaMemcpyAsync(e.wcnt, e.cnt, (size_t)std::min(rows, e.rows) * 4, cudaMemcpyDeviceToDevice, g_stream));
    }
    uint32_t* tmp_rows = e.kind == vb::KIND_CSR ? nullptr : dalloc<uint32_t>(total);
    uint64_t pos = e.kind == vb::KIND_CSR ? e.log_n : 0;
    for (auto& c : e.chunks) {
        TranslateArgs ta{};
        ta.to = c.to; ta.from = c.from; ta.n = c.n;
        ta.log_to = e.kind == vb::KIND_CSR ? e.log_to : tmp_rows; ta.log_from = e.log_from; ta.pos0 = pos;
        std::memcpy(ta.base, base, sizeof(ta.base));
        for (size_t t = 1; t <= agents.size(); ++t) ta.nslots[t] = (uint32_t)std::max<uint64_t>(agents[t - 1].nslots, agents[t - 1].nextid - 1);
        for (size_t t = 1; t <= agents.size(); ++t) { ta.lcap[t] = agents[t - 1].cap; ta.nghost[t] = agents[t - 1].nghost; ta.ghost_ids[t] = agents[t - 1].ghost_ids; }
        ta.ntypes = (uint32_t)agents.size(); ta.target = e.singletype ? e.target : 0; ta.ignore_from = !e.has_src() || e.kind != vb::KIND_CSR; ta.rank = rank; ta.error = d_error;
        translate_edges_kernel<<<nblk(c.n), 256, 0, g_stream>>>(ta); LAUNCH_CHECK();
        if (e.kind == vb::KIND_CSR && e.has_state()) {
            vbp::aos_to_soa_kernel<<<nblk(c.n * e.ncols), 256, 0, g_stream>>>(c.st, e.log_st, e.log_cap, pos, c.n, e.size, e.word); LAUNCH_CHECK();
        }
        pos += c.n;
    }
    check_device_error("add_edge!");
    if (e.kind == vb::KIND_CSR) {
        e.log_n = (uint32_t)pos;
        build_container(ei, true);
    } else {
        count_adds_kernel<<<nblk(total), 256, 0, g_stream>>>(tmp_rows, total, e.wcnt, e.kind == vb::KIND_FLAG); LAUNCH_CHECK();
        dfree(tmp_rows);
        dfree(e.cnt); e.cnt = e.wcnt; e.rows = e.rows_w; e.wcnt = nullptr; e.rows_w = 0;
    }
    CK(cudaStreamSynchronize(g_stream));
    free_chunks(e);
    e.raw_n = 0;
}

// remove_edges! records -> filter the append log and the existing container (which is being merged: removes require add_existing)
void vb_sim::apply_removes(int ei) {
    EdgeStore& e = E(ei);
    const uint32_t nrec = e.rm_n;
    if (!nrec) return;
    e.rm_n = 0;
    const uint32_t rows = rows_of(e);
    uint32_t* cutA = dalloc<uint32_t>((size_t)rows + 2);
    CK(cudaMemsetAsync(cutA, 0, ((size_t)rows + 2) * 4, g_stream));
    uint32_t* pflag = dalloc<uint32_t>(nrec); uint32_t* ppos = dalloc<uint32_t>(nrec);
    uint32_t* scr = dalloc<uint32_t>(vbp::scan_scratch_words(std::max<uint64_t>(nrec, (uint64_t)rows + 1)));
    rm_cut_kernel<<<nblk(nrec), 256, 0, g_stream>>>(e.rm_row, e.rm_from, e.rm_mark, nrec, cutA, pflag); LAUNCH_CHECK();
    vbp::exclusive_scan(pflag, ppos, nrec, d_scalars, scr, g_stream); g_launches += 3;
    uint32_t npair = 0;
    CK(cudaMemcpyAsync(&npair, d_scalars, 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    RmArgs rm{cutA, nullptr, nullptr, nullptr};
    uint32_t *prow = nullptr, *pfrom = nullptr, *pmark = nullptr, *poff = nullptr, *t0 = nullptr, *t1 = nullptr, *t2 = nullptr;
    if (npair) {   // group the (from,to) records by row: stable sort on the row, payloads from + mark
        prow = dalloc<uint32_t>(npair); pfrom = dalloc<uint32_t>(npair); pmark = dalloc<uint32_t>(npair);
        t0 = dalloc<uint32_t>(npair); t1 = dalloc<uint32_t>(npair); t2 = dalloc<uint32_t>(npair);
        compact_u32_kernel<<<nblk(nrec), 256, 0, g_stream>>>(e.rm_row, pflag, ppos, nrec, prow); LAUNCH_CHECK();
        compact_u32_kernel<<<nblk(nrec), 256, 0, g_stream>>>(e.rm_from, pflag, ppos, nrec, pfrom); LAUNCH_CHECK();
        compact_u32_kernel<<<nblk(nrec), 256, 0, g_stream>>>(e.rm_mark, pflag, ppos, nrec, pmark); LAUNCH_CHECK();
        uint32_t* sscr = dalloc<uint32_t>(vbp::rs_scratch_words(npair));
        const int res = vbp::radix_sort(prow, t0, pfrom, t1, pmark, t2, 4, npair, vbp::bits_for(rows), sscr, g_stream);
        CK(cudaGetLastError());
        dfree(sscr);
        if (res) { std::swap(prow, t0); std::swap(pfrom, t1); std::swap(pmark, t2); }
        uint32_t* pc = dalloc<uint32_t>((size_t)rows + 2);
        CK(cudaMemsetAsync(pc, 0, ((size_t)rows + 2) * 4, g_stream));
        vbp::csr_run_counts_kernel<<<nblk(npair), 256, 0, g_stream>>>(prow, npair, pc); LAUNCH_CHECK();
        poff = dalloc<uint32_t>((size_t)rows + 2);
        vbp::exclusive_scan(pc, poff, (uint64_t)rows + 1, nullptr, scr, g_stream); g_launches += 3;
        dfree(pc);
        rm.poff = poff; rm.pfrom = pfrom; rm.pmark = pmark;
    }
    // --- the append log ---
    if (e.log_n) {
        const uint32_t n = e.log_n;
        uint32_t* keep = dalloc<uint32_t>(n); uint32_t* pos = dalloc<uint32_t>(n);
        uint32_t* scr2 = dalloc<uint32_t>(vbp::scan_scratch_words(n));
        rm_filter_log_kernel<<<nblk(n), 256, 0, g_stream>>>(e.log_to, e.has_src() ? e.log_from : nullptr, n, rm, keep); LAUNCH_CHECK();
        vbp::exclusive_scan(keep, pos, n, d_scalars, scr2, g_stream); g_launches += 3;
        uint32_t kept = 0;
        CK(cudaMemcpyAsync(&kept, d_scalars, 4, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        if (kept != n) {
            uint32_t* nt = dalloc<uint32_t>(e.log_cap);
            compact_u32_kernel<<<nblk(n), 256, 0, g_stream>>>(e.log_to, keep, pos, n, nt); LAUNCH_CHECK();
            dfree(e.log_to); e.log_to = nt;
            if (e.has_src() && e.log_from) {
                uint32_t* nf = dalloc<uint32_t>(e.log_cap);
                compact_u32_kernel<<<nblk(n), 256, 0, g_stream>>>(e.log_from, keep, pos, n, nf); LAUNCH_CHECK();
                dfree(e.log_from); e.log_from = nf;
            }
            if (e.has_state() && e.log_st) {
                uint8_t* ns = (uint8_t*)g_pool.alloc((size_t)e.log_cap * e.size);
                compact_soa_kernel<<<nblk((uint64_t)n * e.ncols), 256, 0, g_stream>>>(e.log_st, e.log_cap, keep, pos, n, ns, e.log_cap, e.word, e.ncols); LAUNCH_CHECK();
                dfree(e.log_st); e.log_st = ns;
            }
            e.log_n = kept;
        }
        CK(cudaStreamSynchronize(g_stream));
        dfree(keep); dfree(pos); dfree(scr2);
    }
    // --- the existing container ---
    if (e.kind != vb::KIND_CSR) {
        if (e.wcnt) { rm_zero_rows_kernel<<<nblk(std::min(rows, e.rows_w)), 256, 0, g_stream>>>(e.wcnt, std::min(rows, e.rows_w), cutA); LAUNCH_CHECK(); }
    } else if (e.off && e.nnz) {
        PurgeArgs pa{};
        pa.off = e.off; pa.src = e.src; pa.st = e.st; pa.stride = e.st_cap; pa.rows = e.rows;
        uint32_t* cnt = dalloc<uint32_t>((size_t)e.rows + 2);
        rm_filter_old_count_kernel<<<nblk((uint64_t)e.rows + 1), 256, 0, g_stream>>>(e.off, e.src, e.rows, rm, cnt); LAUNCH_CHECK();
        uint32_t* noff = dalloc<uint32_t>((size_t)e.rows + 2);
        vbp::exclusive_scan(cnt, noff, (uint64_t)e.rows + 1, d_scalars, scr, g_stream); g_launches += 3;
        uint32_t total = 0;
        CK(cudaMemcpyAsync(&total, d_scalars, 4, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        if (total != e.nnz) {
            const uint32_t cap = std::max<uint32_t>(total, 1);
            pa.noff = noff; pa.nsrc = e.has_src() ? dalloc<uint32_t>(cap) : nullptr;
            pa.nst = e.has_state() ? (uint8_t*)g_pool.alloc((size_t)cap * e.size) : nullptr;
            pa.nstride = cap; pa.word = e.word; pa.ncols = e.ncols;
            rm_filter_old_copy_kernel<<<nblk(e.rows), 256, 0, g_stream>>>(pa, rm); LAUNCH_CHECK();
            CK(cudaStreamSynchronize(g_stream));
            dfree(e.off); dfree(e.src); dfree(e.st); ++e.version;
            e.off = noff; e.src = pa.nsrc; e.st = pa.nst; e.st_cap = cap; e.nnz = total;
            noff = nullptr;
        }
        dfree(cnt); dfree(noff);
    }
    CK(cudaStreamSynchronize(g_stream));
    dfree(cutA); dfree(pflag); dfree(ppos); dfree(scr); dfree(prow); dfree(pfrom); dfree(pmark); dfree(poff); dfree(t0); dfree(t1); dfree(t2);
}

// finish_write! for one edge type: sorted append log (+ the existing container when add_existing) -> new
// read container.  Per-target order = append order (stable sort), old entries first.
void vb_sim::build_container(int ei, bool add_existing) {
    EdgeStore& e = E(ei);
    const uint32_t rows = rows_of(e);
    apply_removes(ei);
    if (e.kind != vb::KIND_CSR) {   // count / flag containers were written in place by the transition
        if (e.wcnt && e.log_to && e.log_n) {   // ordered path (an apply with remove_edges!): the surviving appends are counted now
            count_adds_kernel<<<nblk(e.log_n), 256, 0, g_stream>>>(e.log_to, e.log_n, e.wcnt, e.kind == vb::KIND_FLAG); LAUNCH_CHECK();
        }
        dfree(e.log_to); e.log_to = nullptr; e.log_n = 0; e.log_cap = 0;
        if (!e.wcnt) { e.wcnt = dalloc<uint32_t>((size_t)rows + 1); e.rows_w = rows; CK(cudaMemsetAsync(e.wcnt, 0, ((size_t)rows + 1) * 4, g_stream)); }
        dfree(e.cnt); e.cnt = e.wcnt; e.rows = e.rows_w; e.wcnt = nullptr; e.rows_w = 0;
        return;
    }
    uint32_t n = e.log_n;
    const bool have_old = add_existing && e.off && e.nnz > 0;
    // --- sort the log by target row (stable) ---
    uint32_t* skey = e.log_to; uint32_t* sfrom = e.log_from; uint8_t* sst = e.log_st; const uint32_t sstride = e.log_cap;
    if (n > 1) {
        const int bits = vbp::bits_for(rows);
        const bool direct = e.has_state() && e.ncols == 1 && (e.word == 4 || e.word == 8);
        const int pbits = (bits + 7) & ~7;                 // whole digits: the passes look at all of the key's low pbits bits
        const bool packed = e.has_state() && e.ncols == 1 && (e.word == 1 || e.word == 2) && pbits + 8 * (int)e.word <= 32;     // the state fits above the sorted digits
        const bool widened = e.has_state() && e.ncols == 1 && (e.word == 1 || e.word == 2) && !packed;
        const bool via_perm = e.has_state() && !direct && !widened && !packed;
        // buffer set A = the log itself, set B = scratch of the same capacity; the sort ping-pongs between them
        uint32_t* kA = e.log_to; uint32_t* kB = dalloc<uint32_t>(e.log_cap);
        uint32_t* fA = e.log_from; uint32_t* fB = e.has_src() ? dalloc<uint32_t>(e.log_cap) : nullptr;
        void* pA = nullptr; void* pB = nullptr; int p2b = 0;
        if (direct) { p2b = (int)e.word; pA = e.log_st; pB = g_pool.alloc((size_t)e.log_cap * e.word); }
        else if (packed) { pack_state_into_key_kernel<<<nblk(n), 256, 0, g_stream>>>(kA, e.log_st, n, e.word, pbits); LAUNCH_CHECK(); }
        else if (widened) {
            p2b = 4; pA = dalloc<uint32_t>(e.log_cap); pB = dalloc<uint32_t>(e.log_cap);
            widen_state_kernel<<<nblk(n), 256, 0, g_stream>>>(e.log_st, n, e.word, (uint32_t*)pA); LAUNCH_CHECK();
        }
        else if (via_perm) {
            p2b = 4; pA = dalloc<uint32_t>(e.log_cap); pB = dalloc<uint32_t>(e.log_cap);
            vbp::iota_u32_kernel<<<nblk(n), 256, 0, g_stream>>>((uint32_t*)pA, n); LAUNCH_CHECK();
        }
        uint32_t* scratch = dalloc<uint32_t>(vbp::rs_scratch_words(n));
        g_launches += (unsigned long long)((bits + 7) / 8) * 5;
        const int res = vbp::radix_sort(kA, kB, fA, fB, pA, pB, p2b, n, bits, scratch, g_stream);
        CK(cudaGetLastError());
        dfree(scratch);
        if (res == 1) { std::swap(kA, kB); std::swap(fA, fB); std::swap(pA, pB); }   // A now holds the sorted data
        e.log_to = kA; e.log_from = fA;
        dfree(kB); dfree(fB);
        if (direct) { e.log_st = (uint8_t*)pA; dfree(pB); }
        else if (packed) { unpack_state_from_key_kernel<<<nblk(n), 256, 0, g_stream>>>(kA, e.log_st, n, e.word, pbits); LAUNCH_CHECK(); }
        else if (widened) {
            narrow_state_kernel<<<nblk(n), 256, 0, g_stream>>>((const uint32_t*)pA, n, e.word, e.log_st); LAUNCH_CHECK();
            dfree(pA); dfree(pB);
        }
        else if (via_perm) {   // gather the state columns through the sort permutation
            uint8_t* g = (uint8_t*)g_pool.alloc((size_t)e.log_cap * e.size);
            gather_soa_kernel<<<nblk((uint64_t)n * e.ncols), 256, 0, g_stream>>>(e.log_st, e.log_cap, (const uint32_t*)pA, n, g, e.log_cap, e.word, e.ncols); LAUNCH_CHECK();
            dfree(e.log_st); e.log_st = g;
            dfree(pA); dfree(pB);
        }
        skey = e.log_to; sfrom = e.log_from; sst = e.log_st;
    }
    // --- :SingleEdge: the last add to a target wins; two different values for one target assert (Dict containers) ---
    if (e.singleedge && n > 1) {
        if (!e.singletype && asserts_enabled) {
            CK(cudaMemsetAsync(d_scalars, 0, 4, g_stream));
            single_edge_conflict_kernel<<<nblk(n), 256, 0, g_stream>>>(skey, sfrom, e.has_state() ? sst : nullptr, sstride, e.word, e.ncols, n, d_scalars); LAUNCH_CHECK();
            uint32_t conflict = 0;
            CK(cudaMemcpyAsync(&conflict, d_scalars, 4, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            if (conflict) { e.log_n = 0; throw AssertionError("An edge has already been added to this agent (the edge type has the :SingleEdge hint)"); }
        }
        uint32_t* flag = dalloc<uint32_t>(n); uint32_t* pos = dalloc<uint32_t>(n);
        uint32_t* scr = dalloc<uint32_t>(vbp::scan_scratch_words(n));
        last_of_run_flags_kernel<<<nblk(n), 256, 0, g_stream>>>(skey, n, flag); LAUNCH_CHECK();
        vbp::exclusive_scan(flag, pos, n, d_scalars, scr, g_stream); g_launches += 3;
        uint32_t kept = 0;
        CK(cudaMemcpyAsync(&kept, d_scalars, 4, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        if (kept != n) {
            uint32_t* nk = dalloc<uint32_t>(e.log_cap);
            compact_u32_kernel<<<nblk(n), 256, 0, g_stream>>>(skey, flag, pos, n, nk); LAUNCH_CHECK();
            dfree(e.log_to); e.log_to = nk; skey = nk;
            if (e.has_src()) {
                uint32_t* nf = dalloc<uint32_t>(e.log_cap);
                compact_u32_kernel<<<nblk(n), 256, 0, g_stream>>>(sfrom, flag, pos, n, nf); LAUNCH_CHECK();
                dfree(e.log_from); e.log_from = nf; sfrom = nf;
            }
            if (e.has_state()) {
                uint8_t* ns = (uint8_t*)g_pool.alloc((size_t)e.log_cap * e.size);
                compact_soa_kernel<<<nblk((uint64_t)n * e.ncols), 256, 0, g_stream>>>(sst, sstride, flag, pos, n, ns, e.log_cap, e.word, e.ncols); LAUNCH_CHECK();
                dfree(e.log_st); e.log_st = ns; sst = ns;
            }
            n = kept; e.log_n = kept;
        }
        dfree(flag); dfree(pos); dfree(scr);
    }
    if (!have_old) {
        // --- the sorted log becomes the container: its offsets come straight from the sorted keys (no counts, no scan over the rows) ---
        uint32_t* off = dalloc<uint32_t>((size_t)rows + 2);
        if (n == 0) CK(cudaMemsetAsync(off, 0, ((size_t)rows + 2) * 4, g_stream));
        else {
            uint32_t* gaps = dalloc<uint32_t>(3 * ((size_t)rows / 32 + 2) + 1);
            uint32_t* gap_count = d_scalars + 2;
            CK(cudaMemsetAsync(gap_count, 0, 4, g_stream));
            vbp::csr_offsets_kernel<<<nblk((uint64_t)n + 1), 256, 0, g_stream>>>(skey, n, rows, off, gaps, gap_count); LAUNCH_CHECK();
            vbp::csr_offset_gaps_kernel<<<296, 256, 0, g_stream>>>(off, gaps, gap_count); LAUNCH_CHECK();
            CK(cudaMemsetAsync(off + rows + 1, 0, 4, g_stream));
            dfree(gaps);
        }
        CK(cudaGetLastError());
        free_edge_read(e);
        e.off = off; e.rows = rows; e.nnz = n;
        e.src = e.log_from; e.st = e.log_st; e.st_cap = e.log_cap;       // the sorted log becomes the container
        e.log_from = nullptr; e.log_st = nullptr;
        dfree(e.log_to); e.log_to = nullptr; e.log_n = 0; e.log_cap = 0;
        return;
    }
    // --- add_existing: row counts of the new entries -> run offsets, merged with the old rows ---
    uint32_t* ncnt = dalloc<uint32_t>((size_t)rows + 2);
    CK(cudaMemsetAsync(ncnt, 0, ((size_t)rows + 2) * 4, g_stream));
    if (n) { vbp::csr_run_counts_kernel<<<nblk(n), 256, 0, g_stream>>>(skey, n, ncnt); LAUNCH_CHECK(); }
    uint32_t* scr = dalloc<uint32_t>(vbp::scan_scratch_words((uint64_t)rows + 1));
    {
        uint32_t* noff = dalloc<uint32_t>((size_t)rows + 2);
        vbp::exclusive_scan(ncnt, noff, (uint64_t)rows + 1, nullptr, scr, g_stream); g_launches += 3;   // run offsets of the new entries
        uint32_t* cnt = dalloc<uint32_t>((size_t)rows + 2);
        merge_counts_kernel<<<nblk((uint64_t)rows + 1), 256, 0, g_stream>>>(e.off, e.rows, ncnt, rows, e.singleedge, cnt); LAUNCH_CHECK();
        uint32_t* off = dalloc<uint32_t>((size_t)rows + 2);
        vbp::exclusive_scan(cnt, off, (uint64_t)rows + 1, d_scalars, scr, g_stream); g_launches += 3;
        uint32_t total = 0;
        CK(cudaMemcpyAsync(&total, d_scalars, 4, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        const uint32_t cap = std::max<uint32_t>(total, 1);
        MergeArgs ma{};
        ma.ooff = e.off; ma.osrc = e.src; ma.ost = e.st; ma.ostride = e.st_cap; ma.orows = e.rows;
        ma.noff = noff; ma.nsrc = sfrom; ma.nst = sst; ma.nstride = sstride;
        ma.off = off; ma.src = e.has_src() ? dalloc<uint32_t>(cap) : nullptr; ma.st = e.has_state() ? (uint8_t*)g_pool.alloc((size_t)cap * e.size) : nullptr;
        ma.stride = cap; ma.rows = rows; ma.word = e.word; ma.ncols = e.ncols; ma.single_edge = e.singleedge;
        const bool check_single = e.singleedge && !e.singletype && asserts_enabled && n > 0;
        if (check_single) { ma.conflict = d_scalars + 1; CK(cudaMemsetAsync(ma.conflict, 0, 4, g_stream)); }
        merge_copy_kernel<<<nblk(rows), 256, 0, g_stream>>>(ma); LAUNCH_CHECK();
        if (check_single) {
            uint32_t conflict = 0;
            CK(cudaMemcpyAsync(&conflict, ma.conflict, 4, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            if (conflict) {
                dfree(ma.src); dfree(ma.st); dfree(noff); dfree(cnt); dfree(off); dfree(ncnt); dfree(scr);
                free_edge_log(e);
                throw AssertionError("An edge has already been added to this agent (the edge type has the :SingleEdge hint)");
            }
        }
        free_edge_read(e);
        e.off = off; e.rows = rows; e.nnz = total; e.src = ma.src; e.st = ma.st; e.st_cap = cap;
        dfree(noff); dfree(cnt);
        free_edge_log(e);
    }
    dfree(ncnt); dfree(scr);
}

// purge edges of agents that died in this apply from every edge type's container
void vb_sim::purge_dead(const uint8_t* dead) {
    for (auto& e : edges) {
        const uint32_t rb = row_base_of(e);
        if (e.kind != vb::KIND_CSR) {
            if (e.cnt && e.rows) { purge_rows_cnt_kernel<<<nblk(e.rows), 256, 0, g_stream>>>(e.cnt, e.rows, dead, rb); LAUNCH_CHECK(); }
            continue;
        }
        if (!e.off || !e.nnz) continue;
        PurgeArgs pa{};
        pa.off = e.off; pa.src = e.src; pa.st = e.st; pa.stride = e.st_cap; pa.rows = e.rows; pa.dead = dead; pa.row_base = rb;
        pa.check_src = e.has_src() && !all_immortal;
        pa.cnt = dalloc<uint32_t>((size_t)e.rows + 2);
        purge_count_kernel<<<nblk((uint64_t)e.rows + 1), 256, 0, g_stream>>>(pa); LAUNCH_CHECK();
        uint32_t* noff = dalloc<uint32_t>((size_t)e.rows + 2);
        uint32_t* scr = dalloc<uint32_t>(vbp::scan_scratch_words((uint64_t)e.rows + 1));
        vbp::exclusive_scan(pa.cnt, noff, (uint64_t)e.rows + 1, d_scalars, scr, g_stream); g_launches += 3;
        uint32_t total = 0;
        CK(cudaMemcpyAsync(&total, d_scalars, 4, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        if (total != e.nnz) {
            const uint32_t cap = std::max<uint32_t>(total, 1);
            pa.noff = noff; pa.nsrc = e.has_src() ? dalloc<uint32_t>(cap) : nullptr;
            pa.nst = e.has_state() ? (uint8_t*)g_pool.alloc((size_t)cap * e.size) : nullptr;
            pa.nstride = cap; pa.word = e.word; pa.ncols = e.ncols;
            purge_copy_kernel<<<nblk(e.rows), 256, 0, g_stream>>>(pa); LAUNCH_CHECK();
            dfree(e.off); dfree(e.src); dfree(e.st); ++e.version;
            e.off = noff; e.src = pa.nsrc; e.st = pa.nst; e.st_cap = cap; e.nnz = total;
            e.last_change = num_transitions;
            noff = nullptr;
        }
        dfree(pa.cnt); dfree(noff); dfree(scr);
    }
}

// Multi-GPU (one process per GPU).  Agents are owned by one rank (rank bits of the id, src/Agent.jl:39-40), every edge is
// stored on its target's rank (src/EdgeMethods.jl:396-398).  The state of remote *sources* is mirrored in a ghost segment
// behind the local slots of each agent type; this replaces the reference's request/reply halo (transmit_agents!,
// src/MPI.jl:155-267): the request lists are exchanged once here, per step only packed states travel (halo_exchange).
void vb_sim::build_ghosts(const uint64_t* const* extra, const uint64_t* extra_n, int n_extra) {
    if (g_nranks <= 1) return;
    const uint32_t NT = (uint32_t)agents.size(), P = (uint32_t)g_nranks;
    // 1. candidate ids: remote sources of the raw adds + the caller's id arrays (edges received from other ranks) + the
    //    ghosts already mirrored; split into 32-bit halves for the radix sort
    uint64_t cap = 0;
    for (auto& e : edges) { if (!e.has_src()) continue; flush_raw((int)(&e - &edges[0])); for (auto& c : e.chunks) cap += c.n; }
    for (int i = 0; i < n_extra; ++i) cap += extra_n[i];
    for (auto& a : agents) cap += a.nghost;
    uint32_t* lo = dalloc<uint32_t>(std::max<uint64_t>(cap, 1)); uint32_t* hi = dalloc<uint32_t>(std::max<uint64_t>(cap, 1));
    uint64_t nrem = 0;
    auto add_filtered = [&](const uint64_t* ids, uint64_t n) {
        if (!n) return;
        if (n >= 0xffffffffull) throw ArgError("id chunk too large");
        uint32_t* flag = dalloc<uint32_t>(n); uint32_t* pos = dalloc<uint32_t>(n);
        uint32_t* scr = dalloc<uint32_t>(vbp::scan_scratch_words(n));
        remote_flags_kernel<<<nblk(n), 256, 0, g_stream>>>(ids, n, rank, flag); LAUNCH_CHECK();
        vbp::exclusive_scan(flag, pos, n, d_scalars, scr, g_stream);
        uint32_t k = 0;
        CK(cudaMemcpyAsync(&k, d_scalars, 4, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        if (k) { compact_u64_split_kernel<<<nblk(n), 256, 0, g_stream>>>(ids, flag, pos, n, lo, hi, nrem); LAUNCH_CHECK(); }
        nrem += k;
        CK(cudaStreamSynchronize(g_stream));
        dfree(flag); dfree(pos); dfree(scr);
    };
    for (auto& e : edges) { if (!e.has_src()) continue; for (auto& c : e.chunks) add_filtered(c.from, c.n); }
    for (int i = 0; i < n_extra; ++i) add_filtered(extra[i], extra_n[i]);
    for (auto& a : agents)
        if (a.nghost) { split64_kernel<<<nblk(a.nghost), 256, 0, g_stream>>>(a.ghost_ids, a.nghost, lo, hi, nrem); LAUNCH_CHECK(); nrem += a.nghost; }
    if (nrem >= 0xffffffffull) throw ArgError("too many remote edge sources on one rank");
    // 2. sort by (hi, lo) = ascending AgentID (two stable 32-bit sorts), then unique
    uint64_t* G = nullptr; uint32_t ng = 0;
    if (nrem) {
        uint32_t* lo2 = dalloc<uint32_t>(nrem); uint32_t* hi2 = dalloc<uint32_t>(nrem);
        uint32_t* scratch = dalloc<uint32_t>(vbp::rs_scratch_words(nrem));
        int r = vbp::radix_sort(lo, lo2, hi, hi2, nullptr, nullptr, 0, nrem, 32, scratch, g_stream);     // key = low half, payload = high half
        uint32_t* slo = r ? lo2 : lo; uint32_t* shi = r ? hi2 : hi; uint32_t* tlo = r ? lo : lo2; uint32_t* thi = r ? hi : hi2;
        r = vbp::radix_sort(shi, thi, slo, tlo, nullptr, nullptr, 0, nrem, 32, scratch, g_stream);       // key = high half (stable)
        uint32_t* fhi = r ? thi : shi; uint32_t* flo = r ? tlo : slo;
        CK(cudaGetLastError());
        uint32_t* flag = dalloc<uint32_t>(nrem); uint32_t* pos = dalloc<uint32_t>(nrem);
        uint32_t* scr = dalloc<uint32_t>(vbp::scan_scratch_words(nrem));
        unique_flags64_kernel<<<nblk(nrem), 256, 0, g_stream>>>(fhi, flo, nrem, flag); LAUNCH_CHECK();
        vbp::exclusive_scan(flag, pos, nrem, d_scalars, scr, g_stream);
        CK(cudaMemcpyAsync(&ng, d_scalars, 4, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        G = dalloc<uint64_t>(std::max<uint32_t>(ng, 1));
        compact_join64_kernel<<<nblk(nrem), 256, 0, g_stream>>>(fhi, flo, flag, pos, nrem, G); LAUNCH_CHECK();
        CK(cudaStreamSynchronize(g_stream));
        dfree(lo2); dfree(hi2); dfree(scratch); dfree(flag); dfree(pos); dfree(scr);
    }
    dfree(lo); dfree(hi);
    // 3. boundaries per (type, owner rank); remap of the previous ghost numbering
    std::vector<uint32_t> bounds((size_t)NT * P + 1, 0);
    if (ng) {
        uint32_t* db = dalloc<uint32_t>(bounds.size());
        ghost_bounds_kernel<<<nblk(bounds.size()), 256, 0, g_stream>>>(G, ng, NT, P, db); LAUNCH_CHECK();
        CK(cudaMemcpyAsync(bounds.data(), db, bounds.size() * 4, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        dfree(db);
    }
    uint32_t old_base[vb::MAX_AGENT_TYPES + 2]; uint32_t old_lcap[vb::MAX_AGENT_TYPES + 1] = {0};
    const uint32_t* remap[vb::MAX_AGENT_TYPES + 1] = {nullptr};
    std::memcpy(old_base, base, sizeof(old_base));
    bool changed = false;
    for (uint32_t t = 1; t <= NT; ++t) {
        AgentStore& a = agents[t - 1];
        old_lcap[t] = a.cap;
        const uint32_t b0 = bounds[(size_t)(t - 1) * P], b1 = bounds[(size_t)t * P];
        const uint32_t nn = b1 - b0;
        if (nn != a.nghost) {
            changed = true;
            uint64_t* ng_ids = dalloc<uint64_t>(std::max<uint32_t>(nn, 1));
            if (nn) CK(cudaMemcpyAsync(ng_ids, G + b0, (size_t)nn * 8, cudaMemcpyDeviceToDevice, g_stream));
            if (a.nghost) {
                uint32_t* rm = dalloc<uint32_t>(a.nghost);
                ghost_remap_kernel<<<nblk(a.nghost), 256, 0, g_stream>>>(a.ghost_ids, a.nghost, ng_ids, nn, rm); LAUNCH_CHECK();
                remap[t] = rm;
            }
            CK(cudaStreamSynchronize(g_stream));
            dfree(a.ghost_ids);
            a.ghost_ids = ng_ids;
            a.nghost = nn;
        }
        a.ghost_off.assign(P + 1, 0);
        for (uint32_t r = 0; r <= P; ++r) a.ghost_off[r] = bounds[(size_t)(t - 1) * P + r] - b0;
        ensure_agent_cap((int)t, a.cap, a.nghost + a.nghost / 4, false);   // some headroom: the table only grows
    }
    compute_bases(base);
    rebase(old_base, old_lcap, remap);
    CK(cudaStreamSynchronize(g_stream));
    for (uint32_t t = 1; t <= NT; ++t) dfree((void*)remap[t]);
    dfree(G);
    // 4. all ranks agree whether anybody's tables changed; if so the request lists are exchanged again
    uint64_t mine = changed ? 1 : 0;
    std::vector<uint64_t> all;
    allgather8_host(&mine, all);
    bool any = false;
    for (uint64_t v : all) any |= v != 0;
    if (any || agents[0].send_off.empty()) exchange_ghost_requests();
}

// tell every owner which of its agents this rank mirrors: counts (all-gather), then the id lists (grouped send/recv)
void vb_sim::exchange_ghost_requests() {
    const uint32_t NT = (uint32_t)agents.size(), P = (uint32_t)g_nranks;
    uint32_t* dcnt = dalloc<uint32_t>((size_t)P * NT); uint32_t* dall = dalloc<uint32_t>((size_t)P * P * NT);
    std::vector<uint32_t> mine((size_t)P * NT), all((size_t)P * P * NT);
    for (uint32_t t = 1; t <= NT; ++t) for (uint32_t r = 0; r < P; ++r) mine[(size_t)(t - 1) * P + r] = agents[t - 1].ghost_off[r + 1] - agents[t - 1].ghost_off[r];
    CK(cudaMemcpyAsync(dcnt, mine.data(), mine.size() * 4, cudaMemcpyHostToDevice, g_stream));
    NK(g_nccl.AllGather(dcnt, dall, mine.size() * 4, ncclUint8, g_comm, g_stream));
    CK(cudaMemcpyAsync(all.data(), dall, all.size() * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    dfree(dcnt); dfree(dall);
    for (uint32_t t = 1; t <= NT; ++t) {
        AgentStore& a = agents[t - 1];
        a.send_off.assign(P + 1, 0);
        for (uint32_t r = 0; r < P; ++r) a.send_off[r + 1] = a.send_off[r] + all[(size_t)r * P * NT + (size_t)(t - 1) * P + rank];   // what rank r wants from me
        const uint32_t ns = a.send_off[P];
        dfree(a.send_slots); a.send_slots = nullptr; dfree(a.send_buf); a.send_buf = nullptr;
        uint64_t* req = dalloc<uint64_t>(std::max<uint32_t>(ns, 1));
        NK(g_nccl.GroupStart());
        for (uint32_t r = 0; r < P; ++r) {
            if (r == rank) continue;
            const uint32_t want = a.ghost_off[r + 1] - a.ghost_off[r], give = a.send_off[r + 1] - a.send_off[r];
            if (want) NK(g_nccl.Send(a.ghost_ids + a.ghost_off[r], (size_t)want * 8, ncclUint8, (int)r, g_comm, g_stream));
            if (give) NK(g_nccl.Recv(req + a.send_off[r], (size_t)give * 8, ncclUint8, (int)r, g_comm, g_stream));
        }
        NK(g_nccl.GroupEnd());
        if (ns) {
            a.send_slots = dalloc<uint32_t>(ns);
            ids_to_slots_kernel<<<nblk(ns), 256, 0, g_stream>>>(req, ns, a.send_slots); LAUNCH_CHECK();
            a.send_buf = (uint8_t*)g_pool.alloc((size_t)ns * std::max<uint32_t>(a.size, 1));
        }
        CK(cudaStreamSynchronize(g_stream));
        dfree(req);
        a.halo_dirty = true;
    }
}

// room for `need` records in the remove log of `e` (on several ranks: plus the AgentID columns of the records that travel)
void vb_sim::ensure_rm(EdgeStore& e, uint64_t need) {
    const bool want64 = g_nranks > 1;
    if (need <= e.rm_cap && (!want64 || e.rm_to64 || e.rm_cap == 0)) return;
    if (need >= 0xffffffffull) throw ArgError("too many remove_edges! records in one apply");
    const uint32_t ncap = need > e.rm_cap ? (uint32_t)std::max<uint64_t>(need, (uint64_t)e.rm_cap * 2 + 1024) : e.rm_cap;
    uint32_t* nr = dalloc<uint32_t>(ncap); uint32_t* nf = dalloc<uint32_t>(ncap); uint32_t* nm = dalloc<uint32_t>(ncap);
    uint64_t* nt64 = want64 ? dalloc<uint64_t>(ncap) : nullptr; uint64_t* nf64 = want64 ? dalloc<uint64_t>(ncap) : nullptr;
    if (e.rm_n) {
        CK(cudaMemcpyAsync(nr, e.rm_row, (size_t)e.rm_n * 4, cudaMemcpyDeviceToDevice, g_stream));
        CK(cudaMemcpyAsync(nf, e.rm_from, (size_t)e.rm_n * 4, cudaMemcpyDeviceToDevice, g_stream));
        CK(cudaMemcpyAsync(nm, e.rm_mark, (size_t)e.rm_n * 4, cudaMemcpyDeviceToDevice, g_stream));
        if (want64) {
            if (e.rm_to64) {
                CK(cudaMemcpyAsync(nt64, e.rm_to64, (size_t)e.rm_n * 8, cudaMemcpyDeviceToDevice, g_stream));
                CK(cudaMemcpyAsync(nf64, e.rm_from64, (size_t)e.rm_n * 8, cudaMemcpyDeviceToDevice, g_stream));
            } else {   // records written before the columns existed are local ones
                CK(cudaMemsetAsync(nt64, 0, (size_t)e.rm_n * 8, g_stream));
                CK(cudaMemsetAsync(nf64, 0, (size_t)e.rm_n * 8, g_stream));
            }
        }
    }
    dfree(e.rm_row); dfree(e.rm_from); dfree(e.rm_mark); dfree(e.rm_to64); dfree(e.rm_from64);
    e.rm_row = nr; e.rm_from = nf; e.rm_mark = nm; e.rm_to64 = nt64; e.rm_from64 = nf64; e.rm_cap = ncap;
}

// transmit_remove_edges! / removeedges_alltoall! (src/MPI.jl:432-479, called at src/Simulation.jl:792-795): the remove_edges! calls of
// this apply whose target lives on another rank were parked as AgentID pairs (from | 0, to) beside the local records.  They are
// bucketed by destination rank (stable), exchanged with grouped ncclSend/ncclRecv (all-to-all-v) and become ordinary records of
// the receiver's remove log.  The reference applies them with the local remove_edges! after the transition loop and before
// transmit_edges!: a received record sees the existing entries and *every* local add of this apply, none of the edges that
// arrive afterwards -> mark = length of the local append log now.  Collective.
void vb_sim::transmit_removes(int ei) {
    EdgeStore& e = E(ei);
    const uint32_t P = (uint32_t)g_nranks, n = e.rm_to64 ? e.rm_n : 0u;
    std::vector<uint32_t> scnt(P + 1, 0);
    uint64_t* sto = nullptr; uint64_t* sfrom = nullptr;
    uint32_t nsend = 0;
    if (n) {
        uint32_t* flag = dalloc<uint32_t>(n); uint32_t* pos = dalloc<uint32_t>(n);
        uint32_t* scr = dalloc<uint32_t>(vbp::scan_scratch_words(n));
        remote_remove_flags_kernel<<<nblk(n), 256, 0, g_stream>>>(e.rm_to64, n, P, flag); LAUNCH_CHECK();
        vbp::exclusive_scan(flag, pos, n, d_scalars, scr, g_stream); g_launches += 3;
        CK(cudaMemcpyAsync(&nsend, d_scalars, 4, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        if (nsend) {
            uint64_t* cto = dalloc<uint64_t>(nsend); uint64_t* cfrom = dalloc<uint64_t>(nsend); uint32_t* cdst = dalloc<uint32_t>(nsend);
            compact_removes_kernel<<<nblk(n), 256, 0, g_stream>>>(e.rm_to64, e.rm_from64, flag, pos, n, cto, cfrom, cdst); LAUNCH_CHECK();
            uint32_t* perm = dalloc<uint32_t>(nsend); uint32_t* k1 = dalloc<uint32_t>(nsend); uint32_t* p1 = dalloc<uint32_t>(nsend);
            vbp::iota_u32_kernel<<<nblk(nsend), 256, 0, g_stream>>>(perm, nsend); LAUNCH_CHECK();
            uint32_t* scratch = dalloc<uint32_t>(vbp::rs_scratch_words(nsend));
            const int res = vbp::radix_sort(cdst, k1, perm, p1, nullptr, nullptr, 0, nsend, vbp::bits_for(P), scratch, g_stream);
            CK(cudaGetLastError());
            dfree(scratch);
            const uint32_t* sdst = res ? k1 : cdst; const uint32_t* sperm = res ? p1 : perm;
            uint32_t* dc = dalloc<uint32_t>(P + 2);
            CK(cudaMemsetAsync(dc, 0, (P + 2) * 4, g_stream));
            vbp::csr_run_counts_kernel<<<nblk(nsend), 256, 0, g_stream>>>(sdst, nsend, dc); LAUNCH_CHECK();
            CK(cudaMemcpyAsync(scnt.data(), dc, P * 4, cudaMemcpyDeviceToHost, g_stream));
            sto = dalloc<uint64_t>(nsend); sfrom = dalloc<uint64_t>(nsend);
            gather_u64_kernel<<<nblk(nsend), 256, 0, g_stream>>>(cto, sperm, nsend, sto); LAUNCH_CHECK();
            gather_u64_kernel<<<nblk(nsend), 256, 0, g_stream>>>(cfrom, sperm, nsend, sfrom); LAUNCH_CHECK();
            CK(cudaStreamSynchronize(g_stream));
            dfree(cto); dfree(cfrom); dfree(cdst); dfree(perm); dfree(k1); dfree(p1); dfree(dc);
        }
        dfree(flag); dfree(pos); dfree(scr);
    }
    // counts: everybody learns the whole P x P matrix
    uint32_t* dcnt = dalloc<uint32_t>(P); uint32_t* dall = dalloc<uint32_t>((size_t)P * P);
    std::vector<uint32_t> all((size_t)P * P);
    CK(cudaMemcpyAsync(dcnt, scnt.data(), P * 4, cudaMemcpyHostToDevice, g_stream));
    NK(g_nccl.AllGather(dcnt, dall, (size_t)P * 4, ncclUint8, g_comm, g_stream));
    CK(cudaMemcpyAsync(all.data(), dall, all.size() * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    dfree(dcnt); dfree(dall);
    std::vector<uint32_t> soff(P + 1, 0), roff(P + 1, 0);
    for (uint32_t r = 0; r < P; ++r) { soff[r + 1] = soff[r] + scnt[r]; roff[r + 1] = roff[r] + (r == rank ? 0u : all[(size_t)r * P + rank]); }
    const uint32_t nrecv = roff[P];
    uint64_t total = 0;
    for (uint32_t v : all) total += v;
    if (total) {
        uint64_t* rto = dalloc<uint64_t>(std::max<uint32_t>(nrecv, 1)); uint64_t* rfrom = dalloc<uint64_t>(std::max<uint32_t>(nrecv, 1));
        NK(g_nccl.GroupStart());
        for (uint32_t r = 0; r < P; ++r) {
            if (r == rank) continue;
            const uint32_t give = scnt[r], want = roff[r + 1] - roff[r];
            if (give) {
                NK(g_nccl.Send(sto + soff[r], (size_t)give * 8, ncclUint8, (int)r, g_comm, g_stream));
                NK(g_nccl.Send(sfrom + soff[r], (size_t)give * 8, ncclUint8, (int)r, g_comm, g_stream));
            }
            if (want) {
                NK(g_nccl.Recv(rto + roff[r], (size_t)want * 8, ncclUint8, (int)r, g_comm, g_stream));
                NK(g_nccl.Recv(rfrom + roff[r], (size_t)want * 8, ncclUint8, (int)r, g_comm, g_stream));
            }
        }
        NK(g_nccl.GroupEnd());
        CK(cudaStreamSynchronize(g_stream));
        halo_bytes += (uint64_t)nrecv * 16;
        if (nrecv) {
            ensure_rm(e, (uint64_t)e.rm_n + nrecv);
            TranslateRmArgs ta{};
            ta.to = rto; ta.from = rfrom; ta.n = nrecv;
            ta.rm_row = e.rm_row; ta.rm_from = e.rm_from; ta.rm_mark = e.rm_mark; ta.rm_to64 = e.rm_to64; ta.rm_from64 = e.rm_from64;
            ta.pos0 = e.rm_n; ta.mark = (e.kind == vb::KIND_CSR || e.ordered_log) ? e.log_n : 0u;
            std::memcpy(ta.base, base, sizeof(ta.base));
            for (size_t t = 1; t <= agents.size(); ++t) {
                ta.lcap[t] = agents[t - 1].cap; ta.nghost[t] = agents[t - 1].nghost; ta.ghost_ids[t] = agents[t - 1].ghost_ids;
            }
            ta.ntypes = (uint32_t)agents.size(); ta.target = e.singletype ? e.target : 0; ta.rank = rank;
            translate_removes_kernel<<<nblk(nrecv), 256, 0, g_stream>>>(ta); LAUNCH_CHECK();
            e.rm_n += nrecv;
            CK(cudaStreamSynchronize(g_stream));
        }
        dfree(rto); dfree(rfrom);
    }
    dfree(sto); dfree(sfrom);
}

// transmit_edges! (src/EdgeMethods.jl:686-689, edges_alltoall! src/MPI.jl:353-430): the appended edges that belong to another rank are
// bucketed by destination (stable: each rank's batch keeps its call order), exchanged with grouped ncclSend/ncclRecv (all-to-all-v)
// and appended behind the local adds in source-rank order (A-15).  Collective: every rank calls it for every written edge type.
void vb_sim::transmit_edges(int ei) {
    EdgeStore& e = E(ei);
    const uint32_t P = (uint32_t)g_nranks, n = e.rlog_n;
    e.rlog_n = 0;
    std::vector<uint32_t> scnt(P + 1, 0);
    uint64_t* sto = nullptr; uint64_t* sfrom = nullptr; uint8_t* sst = nullptr;
    if (n) {
        uint32_t* perm = dalloc<uint32_t>(n); uint32_t* k1 = dalloc<uint32_t>(n); uint32_t* p1 = dalloc<uint32_t>(n);
        vbp::iota_u32_kernel<<<nblk(n), 256, 0, g_stream>>>(perm, n); LAUNCH_CHECK();
        uint32_t* scratch = dalloc<uint32_t>(vbp::rs_scratch_words(n));
        const int res = vbp::radix_sort(e.rlog_dst, k1, perm, p1, nullptr, nullptr, 0, n, vbp::bits_for(P), scratch, g_stream);
        CK(cudaGetLastError());
        dfree(scratch);
        const uint32_t* sdst = res ? k1 : e.rlog_dst; const uint32_t* sperm = res ? p1 : perm;
        uint32_t* dc = dalloc<uint32_t>(P + 2);
        CK(cudaMemsetAsync(dc, 0, (P + 2) * 4, g_stream));
        vbp::csr_run_counts_kernel<<<nblk(n), 256, 0, g_stream>>>(sdst, n, dc); LAUNCH_CHECK();
        CK(cudaMemcpyAsync(scnt.data(), dc, P * 4, cudaMemcpyDeviceToHost, g_stream));
        sto = dalloc<uint64_t>(n);
        gather_u64_kernel<<<nblk(n), 256, 0, g_stream>>>(e.rlog_to, sperm, n, sto); LAUNCH_CHECK();
        if (e.has_src()) { sfrom = dalloc<uint64_t>(n); gather_u64_kernel<<<nblk(n), 256, 0, g_stream>>>(e.rlog_from, sperm, n, sfrom); LAUNCH_CHECK(); }
        if (e.has_state()) {
            sst = (uint8_t*)g_pool.alloc((size_t)n * e.size);
            gather_soa_to_aos_kernel<<<nblk((uint64_t)n * e.ncols), 256, 0, g_stream>>>(e.rlog_st, e.rlog_cap, sperm, n, sst, e.size, e.word); LAUNCH_CHECK();
        }
        CK(cudaStreamSynchronize(g_stream));
        dfree(perm); dfree(k1); dfree(p1); dfree(dc);
    }
    // counts: everybody learns the whole P x P matrix
    uint32_t* dcnt = dalloc<uint32_t>(P); uint32_t* dall = dalloc<uint32_t>((size_t)P * P);
    std::vector<uint32_t> all((size_t)P * P);
    CK(cudaMemcpyAsync(dcnt, scnt.data(), P * 4, cudaMemcpyHostToDevice, g_stream));
    NK(g_nccl.AllGather(dcnt, dall, (size_t)P * 4, ncclUint8, g_comm, g_stream));
    CK(cudaMemcpyAsync(all.data(), dall, all.size() * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    dfree(dcnt); dfree(dall);
    std::vector<uint32_t> soff(P + 1, 0), roff(P + 1, 0);
    for (uint32_t r = 0; r < P; ++r) { soff[r + 1] = soff[r] + scnt[r]; roff[r + 1] = roff[r] + all[(size_t)r * P + rank]; }
    const uint32_t nrecv = roff[P];
    uint64_t* rto = dalloc<uint64_t>(std::max<uint32_t>(nrecv, 1));
    uint64_t* rfrom = e.has_src() ? dalloc<uint64_t>(std::max<uint32_t>(nrecv, 1)) : nullptr;
    uint8_t* rst = e.has_state() ? (uint8_t*)g_pool.alloc((size_t)std::max<uint32_t>(nrecv, 1) * e.size) : nullptr;
    NK(g_nccl.GroupStart());
    for (uint32_t r = 0; r < P; ++r) {
        const uint32_t give = scnt[r], want = roff[r + 1] - roff[r];
        if (r == rank) continue;
        if (give) {
            NK(g_nccl.Send(sto + soff[r], (size_t)give * 8, ncclUint8, (int)r, g_comm, g_stream));
            if (sfrom) NK(g_nccl.Send(sfrom + soff[r], (size_t)give * 8, ncclUint8, (int)r, g_comm, g_stream));
            if (sst) NK(g_nccl.Send(sst + (size_t)soff[r] * e.size, (size_t)give * e.size, ncclUint8, (int)r, g_comm, g_stream));
        }
        if (want) {
            NK(g_nccl.Recv(rto + roff[r], (size_t)want * 8, ncclUint8, (int)r, g_comm, g_stream));
            if (rfrom) NK(g_nccl.Recv(rfrom + roff[r], (size_t)want * 8, ncclUint8, (int)r, g_comm, g_stream));
            if (rst) NK(g_nccl.Recv(rst + (size_t)roff[r] * e.size, (size_t)want * e.size, ncclUint8, (int)r, g_comm, g_stream));
        }
    }
    NK(g_nccl.GroupEnd());
    if (scnt[rank]) {   // edges kept on this rank because their source was not mirrored yet
        CK(cudaMemcpyAsync(rto + roff[rank], sto + soff[rank], (size_t)scnt[rank] * 8, cudaMemcpyDeviceToDevice, g_stream));
        if (rfrom) CK(cudaMemcpyAsync(rfrom + roff[rank], sfrom + soff[rank], (size_t)scnt[rank] * 8, cudaMemcpyDeviceToDevice, g_stream));
        if (rst) CK(cudaMemcpyAsync(rst + (size_t)roff[rank] * e.size, sst + (size_t)soff[rank] * e.size, (size_t)scnt[rank] * e.size, cudaMemcpyDeviceToDevice, g_stream));
    }
    CK(cudaStreamSynchronize(g_stream));
    dfree(sto); dfree(sfrom); dfree(sst);
    halo_bytes += (uint64_t)nrecv * (8 + (e.has_src() ? 8 : 0) + (e.has_state() ? e.size : 0));
    // the sources of the received edges that live on other ranks become ghosts (collective)
    const uint64_t* extra[1] = {rfrom};
    const uint64_t extra_n[1] = {rfrom ? nrecv : 0u};
    build_ghosts(extra, extra_n, 1);
    // translate and append behind the local adds
    if (nrecv) {
        uint32_t* tmp_rows = nullptr;
        if (e.kind == vb::KIND_CSR || e.ordered_log) ensure_log(e, (uint64_t)e.log_n + nrecv);
        else tmp_rows = dalloc<uint32_t>(nrecv);
        TranslateArgs ta{};
        ta.to = rto; ta.from = rfrom; ta.n = nrecv;
        ta.log_to = tmp_rows ? tmp_rows : e.log_to; ta.log_from = e.log_from; ta.pos0 = tmp_rows ? 0 : e.log_n;
        std::memcpy(ta.base, base, sizeof(ta.base));
        for (size_t t = 1; t <= agents.size(); ++t) {
            ta.nslots[t] = (uint32_t)std::max<uint64_t>(agents[t - 1].nslots, agents[t - 1].nextid - 1 + agents[t - 1].births);
            ta.lcap[t] = agents[t - 1].cap; ta.nghost[t] = agents[t - 1].nghost; ta.ghost_ids[t] = agents[t - 1].ghost_ids;
        }
        ta.ntypes = (uint32_t)agents.size(); ta.target = e.singletype ? e.target : 0; ta.ignore_from = !e.has_src() || e.kind != vb::KIND_CSR; ta.rank = rank; ta.error = d_error;
        translate_edges_kernel<<<nblk(nrecv), 256, 0, g_stream>>>(ta); LAUNCH_CHECK();
        if (tmp_rows) {
            count_adds_kernel<<<nblk(nrecv), 256, 0, g_stream>>>(tmp_rows, nrecv, e.wcnt, e.kind == vb::KIND_FLAG); LAUNCH_CHECK();
        } else {
            if (e.kind == vb::KIND_CSR && e.has_state()) { vbp::aos_to_soa_kernel<<<nblk((uint64_t)nrecv * e.ncols), 256, 0, g_stream>>>(rst, e.log_st, e.log_cap, e.log_n, nrecv, e.size, e.word); LAUNCH_CHECK(); }
            e.log_n += nrecv;
        }
        check_device_error("transmit_edges!");
        CK(cudaStreamSynchronize(g_stream));
        dfree(tmp_rows);
    }
    dfree(rto); dfree(rfrom); dfree(rst);
}

// per-step halo: pack the states the peers mirror, grouped ncclSend/ncclRecv (all-to-all-v) straight into the ghost segments

// Source-blocked view of edge type `ei` for the reduce transition `ti` called on agent type C (n slots).  Returns true when the view
// is ready.  Policy: only for gather-bound shapes (the source type's state array is several times the L2 set-aside) and only once
// the same container has been seen by two applies (a network rebuilt every step never amortises the build).
// VB_BLOCK=0 disables, VB_BLOCK_MB sets the block size (default 75 MB of source states), VB_BLOCK_MIN_MB the activation threshold
// (default 192 MB), VB_BLOCK_EAGER=1 builds at first sight (tests).
// Prefiltered sweeps (functors with kPrefilter, include/vahana_model.h) gather one key byte per entry instead of the state, so their
// blocks are sized in key bytes: VB_KEY_BLOCK_MB (default 52 MB of keys = 52 M slots; profiles/r1_prefilter: 50 MB of keys hold the
// L2 gather rate, 100 MB do not).  VB_PREFILTER=0 / vb_set_read_prefilter(sim, 0) keeps the unfiltered sweeps.
bool vb_sim::prefilter_on(const vb::TransitionInfo* ti) const {
    static const bool env_on = !(getenv("VB_PREFILTER") && atoi(getenv("VB_PREFILTER")) == 0);
    return ti->prefilter && ti->launch_keys && (blk_prefilter >= 0 ? blk_prefilter != 0 : env_on);
}
bool vb_sim::ensure_blocked(int ei, const vb::TransitionInfo* ti, int C, uint32_t n, uint32_t heavy_min, const vb::LaunchArgs* la, uint64_t seed) {
    static const double env_key_block_mb = getenv("VB_KEY_BLOCK_MB") ? atof(getenv("VB_KEY_BLOCK_MB")) : 52.0;
    static const bool enabled = !(getenv("VB_BLOCK") && atoi(getenv("VB_BLOCK")) == 0);
    static const double env_block_mb = getenv("VB_BLOCK_MB") ? atof(getenv("VB_BLOCK_MB")) : 75.0;
    static const double env_min_mb = getenv("VB_BLOCK_MIN_MB") ? atof(getenv("VB_BLOCK_MIN_MB")) : 192.0;
    static const bool env_eager = getenv("VB_BLOCK_EAGER") && atoi(getenv("VB_BLOCK_EAGER")) != 0;
    const double block_mb = blk_block_mb > 0 ? blk_block_mb : env_block_mb;       // vb_set_read_blocking overrides the environment
    const double min_mb = blk_min_mb >= 0 ? blk_min_mb : env_min_mb;
    const bool eager = blk_eager >= 0 ? blk_eager != 0 : env_eager;
    if (!enabled || blk_block_mb == 0 || !ti->reduce || !ti->launch_blocked) return false;
    EdgeStore& pe = E(ei);
    AgentStore& src = A(ti->source_type);
    AgentStore& a = A(C);
    if (pe.kind != vb::KIND_CSR || !pe.off || !pe.src || pe.implicit_stencil || !pe.nnz) return false;
    if (pe.singletype && pe.target != C) return false;
    if (!src.size || src.size != ti->source_size) return false;
    if (a.independent && ti->source_type == C) return false;          // in-place states: a later sweep would read updated sources
    const uint32_t nsl = src.cap + src.nghost;
    if ((double)nsl * src.size < min_mb * 1e6) return false;
    if (la && ti->launch_passrate && prefilter_on(ti) && blk_prefilter < 0) {
        // Is the prefilter selective at the moment?  Every 16th apply samples 65 536 entries (one launch, one 16-byte read-back) and
        // switches between the prefiltered and the unfiltered sweeps with some hysteresis.  vb_set_read_prefilter(sim, 1 / 0) pins a form.
        static const double max_pass = getenv("VB_PF_MAX_PASS") ? atof(getenv("VB_PF_MAX_PASS")) : 0.30;
        static const double min_pass = getenv("VB_PF_MIN_PASS") ? atof(getenv("VB_PF_MIN_PASS")) : 0.22;
        if (pe.pf_check_in == 0) {
            unsigned long long* cnt = (unsigned long long*)(d_scalars + 48);
            CK(cudaMemsetAsync(cnt, 0, 16, g_stream));
            vb::LaunchArgs lp = *la;
            lp.stats = cnt;
            upload_view(seed);
            CK(ti->launch_passrate(lp)); ++g_launches;
            unsigned long long h2[2] = {0, 0};
            CK(cudaMemcpyAsync(h2, cnt, 16, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            if (h2[0]) {
                pe.pf_rate = (double)h2[1] / (double)h2[0];
                if (pe.pf_rate > max_pass) pe.pf_off = true; else if (pe.pf_rate < min_pass) pe.pf_off = false;
            }
            pe.pf_check_in = 16;
        }
        --pe.pf_check_in;
        last_pass_rate = pe.pf_rate;
    }
    const bool pf = prefilter_on(ti) && !pe.pf_off;
    uint32_t bsize = (uint32_t)std::max<double>(1.0, block_mb * 1e6 / src.size);
    uint32_t nb = (nsl + bsize - 1) / bsize;
    if (nb > 64) { bsize = (nsl + 63) / 64; nb = (nsl + bsize - 1) / bsize; }
    if (nb < 2 && !pf) return false;
    if (pf) {     // blocks of keys: as few and as even as the key budget allows (an explicit vb_set_read_blocking size scales 1 : sizeof(state))
        const double key_slots = std::max(1.0, (blk_block_mb > 0 ? block_mb / 75.0 : 1.0) * env_key_block_mb * 1e6);
        const uint32_t used = src.nghost ? nsl : std::max<uint32_t>(1u, std::min<uint32_t>(nsl, src.nslots));   // capacity beyond the slots in use holds no source
        nb = (uint32_t)std::max<double>(1.0, std::ceil((double)used / key_slots));
        if (nb > 64) nb = 64;
        bsize = (used + nb - 1) / nb;
        nb = (nsl + bsize - 1) / bsize;                // blocks still cover the capacity; the ones past `used` stay empty and are skipped
        if (nb > 64) { bsize = (nsl + 63) / 64; nb = (nsl + bsize - 1) / bsize; }
    }
    EdgeStore::Blocked& k = pe.blk;
    // prefiltered sweeps walk the SEGMENTED view (rows cut into segments of seg_len entries, no hub pass); VB_PF_SEG=0 keeps the
    // row-based view with the block-per-agent pass for hub rows
    static const bool env_seg = !(getenv("VB_PF_SEG") && atoi(getenv("VB_PF_SEG")) == 0);
    static const uint32_t env_seg_len = getenv("VB_SEG_LEN") ? (uint32_t)std::min(65535, std::max(32, atoi(getenv("VB_SEG_LEN")))) : 2048u;
    bool seg = pf && env_seg;
    uint32_t g_lcap = 0, g_nbl = 0, g_bsl = 0, g_ng = 0, g_gfirst = 0;
    bool g_abs = false;
    if (seg) {
        // blocks of keys: the local slots in as few even blocks as the key budget allows; the ghosts in the blocks the halo travels in
        // (AgentStore::PeerMap: the same count on every rank), so that a ghost block can be swept as soon as its phase has landed
        const double key_slots = std::max(1.0, (blk_block_mb > 0 ? block_mb / 75.0 : 1.0) * env_key_block_mb * 1e6);
        g_lcap = src.nghost ? src.cap : nsl;
        const uint32_t used = std::max<uint32_t>(1u, std::min<uint32_t>(g_lcap, std::max<uint32_t>(src.nslots, 1u)));   // capacity beyond the slots in use holds no source
        g_nbl = (uint32_t)std::max<double>(1.0, std::ceil((double)used / key_slots));
        g_bsl = (((used + g_nbl - 1) / g_nbl) + 63u) & ~63u;
        if (src.nghost) g_ng = src.peers.ok ? src.peers.ng : (uint32_t)std::min<double>(16.0, std::max<double>(1.0, std::ceil((double)src.nghost / key_slots)));
        g_abs = nsl <= (1u << 27);
        g_gfirst = g_nbl;
        if (g_ng && g_abs) {
            // the first ghost part joins the last local block when both fit one key block (VB_KEY_BLOCK_MAX_MB, 60 MB of keys: still
            // inside the L2 set-aside): every block less is one pass less over the rows' offsets, states and parked accumulators
            static const double env_key_max_mb = getenv("VB_KEY_BLOCK_MAX_MB") ? atof(getenv("VB_KEY_BLOCK_MAX_MB")) : 60.0;
            const double last_local = (double)used - (double)(g_nbl - 1) * g_bsl;
            double part0 = 0;
            for (size_t p2 = 0; p2 + 1 < src.ghost_off.size(); ++p2) part0 += AgentStore::ghost_first(src.ghost_off[p2 + 1] - src.ghost_off[p2], g_ng);
            if (last_local + part0 <= std::max(key_slots, env_key_max_mb * 1e6 * (key_slots / (env_key_block_mb * 1e6)))) g_gfirst = g_nbl - 1;
        }
        if (g_gfirst + g_ng > 64 || g_bsl > (1u << 27) || src.nghost > (1u << 27) || g_ng > 16 || (src.nghost && src.ghost_off.size() > 17))
            seg = false;                                                  // (the row-based view below takes over)
        else { nb = g_gfirst + g_ng; bsize = g_bsl; }
    }
    if (k.boff && k.version == pe.version && k.epoch == layout_epoch && k.called == C && k.source == ti->source_type && k.n == n &&
        k.acc_bytes == ti->acc_bytes && k.bsize == bsize && k.heavy_min == heavy_min && (k.key != nullptr) == pf && k.segmented == seg &&
        (!seg || (k.nb == nb && k.lcap == g_lcap && k.nbl == g_nbl && k.ng == g_ng && k.gfirst == g_gfirst && k.absolute == g_abs && (!g_ng || k.goff == src.ghost_off))))
        return true;
    if (k.seen_version != pe.version) { k.seen_version = pe.version; k.seen = 1; k.refused = false; }
    else if (k.seen < 0xffffffffu) ++k.seen;
    if (k.refused || (!eager && k.seen < 2)) return false;
    const uint64_t rpad = ((uint64_t)n + 1 + 3) & ~3ull;
    if (rpad * nb >= 0xfffffff0ull) return false;
    {   // drop a stale view, keep the bookkeeping
        const uint64_t sv = k.seen_version; const uint32_t sn = k.seen;
        free_blocked(pe);
        k.seen_version = sv; k.seen = sn;
    }
    g_trace.begin();
    if (seg) {
        uint32_t *sfirst = nullptr, *seg_lo = nullptr, *seg_hi = nullptr, *scr = nullptr, *flag = nullptr, *pos = nullptr;
        try {
            // segments of the called rows: count -> scan -> fill
            sfirst = dalloc<uint32_t>((uint64_t)n + 2);
            const uint32_t row0 = pe.singletype ? 0u : base[C];
            seg_count_kernel<<<nblk((uint64_t)n + 1), 256, 0, g_stream>>>(pe.off, row0, n, pe.rows, env_seg_len, sfirst); LAUNCH_CHECK();
            scr = dalloc<uint32_t>(vbp::scan_scratch_words((uint64_t)n + 1));
            vbp::exclusive_scan(sfirst, sfirst, (uint64_t)n + 1, d_scalars, scr, g_stream); g_launches += 3;
            uint32_t nseg = 0;
            CK(cudaMemcpyAsync(&nseg, d_scalars, 4, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            dfree(scr); scr = nullptr;
            const uint64_t spad = ((uint64_t)nseg + 1 + 3) & ~3ull;
            if (spad * nb >= 0xfffffff0ull) throw CudaError("segmented view too large");
            k.seg_row = dalloc<uint32_t>((uint64_t)nseg + 1);
            seg_lo = dalloc<uint32_t>((uint64_t)nseg + 1); seg_hi = dalloc<uint32_t>((uint64_t)nseg + 1);
            const uint64_t total = spad * nb;
            k.boff = dalloc<uint32_t>(total + 4);
            CK(cudaMemsetAsync(k.boff, 0, (total + 4) * 4, g_stream));
            CK(cudaMemsetAsync(d_scalars, 0, 8, g_stream));
            SegBuildArgs sa{};
            sa.off = pe.off; sa.src = pe.src; sa.row0 = row0; sa.n = n; sa.rows = pe.rows; sa.seg_len = env_seg_len;
            sa.tb = base[ti->source_type]; sa.nsl = nsl; sa.nb = nb; sa.spad = (uint32_t)spad; sa.nseg = nseg;
            sa.lcap = g_lcap; sa.bsize_l = g_bsl; sa.nbl = g_nbl; sa.ng = g_ng; sa.nowners = g_ng ? (uint32_t)src.ghost_off.size() - 1 : 0;
            sa.gfirst = g_gfirst; sa.absolute = g_abs ? 1u : 0u;
            for (uint32_t p2 = 0; p2 < sa.nowners; ++p2) {
                sa.goff[p2] = src.ghost_off[p2]; sa.goff[p2 + 1] = src.ghost_off[p2 + 1];
                sa.gfirst_len[p2] = AgentStore::ghost_first(src.ghost_off[p2 + 1] - src.ghost_off[p2], g_ng);
                sa.gpart[p2] = AgentStore::ghost_rest(src.ghost_off[p2 + 1] - src.ghost_off[p2], g_ng);
            }
            sa.sfirst = sfirst; sa.seg_row = k.seg_row; sa.seg_lo = seg_lo; sa.seg_hi = seg_hi; sa.boff = k.boff; sa.bsrc = nullptr; sa.error = d_scalars + 1;
            seg_fill_kernel<<<nblk(n), 256, 0, g_stream>>>(sa); LAUNCH_CHECK();
            seg_blk_count_kernel<<<nblk(nseg), 256, 0, g_stream>>>(sa); LAUNCH_CHECK();
            scr = dalloc<uint32_t>(vbp::scan_scratch_words(total + 1));
            vbp::exclusive_scan(k.boff, k.boff, total + 1, d_scalars, scr, g_stream); g_launches += 3;
            uint32_t res[2] = {0, 0};
            CK(cudaMemcpyAsync(res, d_scalars, 8, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            dfree(scr); scr = nullptr;
            if (res[1]) throw CudaError("a source of another agent type");
            k.bstart.assign(nb + 1, res[0]);
            for (uint32_t b = 0; b < nb; ++b) CK(cudaMemcpyAsync(&k.bstart[b], k.boff + (size_t)b * spad, 4, cudaMemcpyDeviceToHost, g_stream));
            k.bsrc = dalloc<uint32_t>((uint64_t)res[0] + 64);
            sa.bsrc = k.bsrc;
            seg_blk_fill_kernel<<<nblk(nseg), 256, 0, g_stream>>>(sa); LAUNCH_CHECK();
            // rows with several segments: their accumulators are merged by a pass of their own
            flag = dalloc<uint32_t>(n); pos = dalloc<uint32_t>(n);
            scr = dalloc<uint32_t>(vbp::scan_scratch_words(n));
            seg_hub_flags_kernel<<<nblk(n), 256, 0, g_stream>>>(sfirst, n, flag); LAUNCH_CHECK();
            vbp::exclusive_scan(flag, pos, n, d_scalars, scr, g_stream); g_launches += 3;
            uint32_t nhub = 0;
            CK(cudaMemcpyAsync(&nhub, d_scalars, 4, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            if (nhub) {
                k.hub_rows = dalloc<uint32_t>(nhub); k.hub_seg = dalloc<uint32_t>((uint64_t)nhub * 2);
                vbp::compact_indices_kernel<<<nblk(n), 256, 0, g_stream>>>(flag, pos, n, k.hub_rows); LAUNCH_CHECK();
                seg_hub_first_kernel<<<nblk(nhub), 256, 0, g_stream>>>(k.hub_rows, nhub, sfirst, k.hub_seg); LAUNCH_CHECK();
            }
            k.acc = (uint8_t*)g_pool.alloc((size_t)spad * ti->acc_bytes);
            k.key_n = nsl; k.key = dalloc<uint8_t>((uint64_t)nsl + 64);
            CK(cudaStreamSynchronize(g_stream));
            dfree(sfirst); dfree(seg_lo); dfree(seg_hi); dfree(scr); dfree(flag); dfree(pos);
            k.arows.assign(nb, nullptr); k.aoff.assign(nb, nullptr); k.acount.assign(nb, 0);
            k.segmented = true; k.nseg = nseg; k.seg_len = env_seg_len; k.nhub = nhub; k.rpad = (uint32_t)spad;
            k.lcap = g_lcap; k.nbl = g_nbl; k.ng = g_ng; k.goff = g_ng ? src.ghost_off : std::vector<uint32_t>();
            k.gfirst = g_gfirst; k.absolute = g_abs;
        } catch (...) {
            dfree(sfirst); dfree(seg_lo); dfree(seg_hi); dfree(scr); dfree(flag); dfree(pos);
            free_blocked(pe);
            k.seen_version = pe.version; k.seen = 2; k.refused = true;     // e.g. out of memory: stay on the direct path
            cudaGetLastError();
            return false;
        }
        k.heavy_min = heavy_min;
        k.nb = nb; k.bsize = bsize; k.n = n; k.acc_bytes = ti->acc_bytes; k.called = C; k.source = ti->source_type;
        k.version = pe.version; k.epoch = layout_epoch;
        g_trace.end("build segmented source-blocked view", pe.name);
        return true;
    }
    try {
        const uint64_t total = rpad * nb;
        k.boff = dalloc<uint32_t>(total + 4);
        CK(cudaMemsetAsync(k.boff, 0, (total + 4) * 4, g_stream));
        CK(cudaMemsetAsync(d_scalars, 0, 8, g_stream));
        BlkBuildArgs ba{};
        ba.off = pe.off; ba.src = pe.src; ba.row0 = pe.singletype ? 0u : base[C]; ba.n = n; ba.rows = pe.rows; ba.heavy_min = heavy_min;
        ba.tb = base[ti->source_type]; ba.nsl = nsl; ba.bsize = bsize; ba.nb = nb; ba.rpad = (uint32_t)rpad;
        ba.boff = k.boff; ba.bsrc = nullptr; ba.error = d_scalars + 1;
        blk_count_kernel<<<nblk(n), 256, 0, g_stream>>>(ba); LAUNCH_CHECK();
        uint32_t* scr = dalloc<uint32_t>(vbp::scan_scratch_words(total + 1));
        vbp::exclusive_scan(k.boff, k.boff, total + 1, d_scalars, scr, g_stream); g_launches += 3;
        uint32_t res[2] = {0, 0};
        CK(cudaMemcpyAsync(res, d_scalars, 8, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        dfree(scr);
        if (res[1]) { free_blocked(pe); k.seen_version = pe.version; k.seen = 2; k.refused = true; return false; }
        k.bstart.assign(nb + 1, res[0]);
        for (uint32_t b = 0; b < nb; ++b) CK(cudaMemcpyAsync(&k.bstart[b], k.boff + (size_t)b * rpad, 4, cudaMemcpyDeviceToHost, g_stream));
        k.bsrc = dalloc<uint32_t>((uint64_t)res[0] + 64);
        ba.bsrc = k.bsrc;
        blk_fill_kernel<<<nblk(n), 256, 0, g_stream>>>(ba); LAUNCH_CHECK();
        k.heavy_bits = dalloc<uint32_t>(((uint64_t)n + 31) / 32 + 1);
        blk_heavy_bits_kernel<<<nblk(n), 256, 0, g_stream>>>(pe.off, ba.row0, n, pe.rows, heavy_min, k.heavy_bits); LAUNCH_CHECK();
        k.acc = (uint8_t*)g_pool.alloc((size_t)rpad * ti->acc_bytes);
        if (pf) { k.key_n = nsl; k.key = dalloc<uint8_t>((uint64_t)nsl + 64); }
        CK(cudaStreamSynchronize(g_stream));
        // per block the list of rows that own an entry in it: the middle sweeps visit only those
        k.arows.assign(nb, nullptr); k.aoff.assign(nb, nullptr); k.acount.assign(nb, 0);
        {
            uint32_t* flag = dalloc<uint32_t>(n); uint32_t* pos = dalloc<uint32_t>(n);
            uint32_t* scr2 = dalloc<uint32_t>(vbp::scan_scratch_words(n));
            for (uint32_t b = 0; b < nb; ++b) {
                if (k.bstart[b + 1] == k.bstart[b]) continue;
                blk_active_flags_kernel<<<nblk(n), 256, 0, g_stream>>>(k.boff + (size_t)b * rpad, n, flag); LAUNCH_CHECK();
                vbp::exclusive_scan(flag, pos, n, d_scalars, scr2, g_stream); g_launches += 3;
                uint32_t cnt = 0;
                CK(cudaMemcpyAsync(&cnt, d_scalars, 4, cudaMemcpyDeviceToHost, g_stream));
                CK(cudaStreamSynchronize(g_stream));
                k.acount[b] = cnt;
                if (cnt && (double)cnt < 0.75 * n) {     // a list only pays off when a good part of the rows can be skipped
                    k.arows[b] = dalloc<uint32_t>(cnt);
                    vbp::compact_indices_kernel<<<nblk(n), 256, 0, g_stream>>>(flag, pos, n, k.arows[b]); LAUNCH_CHECK();
                    k.aoff[b] = dalloc<uint32_t>((uint64_t)cnt + 1);
                    blk_listed_offsets_kernel<<<nblk((uint64_t)cnt + 1), 256, 0, g_stream>>>(k.boff + (size_t)b * rpad, k.arows[b], cnt, k.bstart[b + 1], k.aoff[b]); LAUNCH_CHECK();
                }
            }
            CK(cudaStreamSynchronize(g_stream));
            dfree(flag); dfree(pos); dfree(scr2);
        }
    } catch (...) {
        free_blocked(pe);
        k.seen_version = pe.version; k.seen = 2; k.refused = true;     // e.g. out of memory: stay on the direct path
        cudaGetLastError();
        return false;
    }
    k.heavy_min = heavy_min;
    k.nb = nb; k.bsize = bsize; k.rpad = (uint32_t)rpad; k.n = n; k.acc_bytes = ti->acc_bytes; k.called = C; k.source = ti->source_type;
    k.version = pe.version; k.epoch = layout_epoch;
    g_trace.end("build source-blocked view", pe.name);
    return true;
}

// ---- barriers between the ranks --------------------------------------------------------------------------------------------------
// stream-ordered barrier through NCCL: a collective completes on a rank only after every rank's stream has reached it (fallback)
static void nccl_stream_barrier(cudaStream_t st) {
    static uint64_t* buf = nullptr;
    if (!buf) buf = dalloc<uint64_t>((size_t)g_nranks + 1);
    NK(g_nccl.AllGather(buf + g_nranks, buf, 8, ncclUint8, g_comm, st));
}
// flag barrier over peer memory (peer_barrier_kernel): set up once, collectively
static struct PeerBarrier {
    unsigned long long* mine = nullptr; PeerFlagPtrs peers{}; unsigned long long epoch = 0; bool tried = false, ok = false;
} g_pb;
static void peer_barrier_setup() {     // collective
    if (g_pb.tried) return;
    g_pb.tried = true;
    static const bool enabled = !(getenv("VB_PEER_BARRIER") && atoi(getenv("VB_PEER_BARRIER")) == 0);
    const uint32_t P = (uint32_t)g_nranks;
    struct H { cudaIpcMemHandle_t h; uint64_t good; };
    static_assert(sizeof(H) % 8 == 0, "exchanged in 8-byte units");
    H mine{};
    bool ok = enabled && P <= 16;
    if (ok && cudaMalloc(&g_pb.mine, 64 * sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); ok = false; g_pb.mine = nullptr; }
    if (ok) {
        CK(cudaMemsetAsync(g_pb.mine, 0, 64 * sizeof(unsigned long long), g_stream));
        if (cudaIpcGetMemHandle(&mine.h, g_pb.mine) != cudaSuccess) { cudaGetLastError(); ok = false; }
    }
    mine.good = ok ? 1 : 0;
    uint8_t* dsend = (uint8_t*)g_pool.alloc(sizeof(H)); uint8_t* drecv = (uint8_t*)g_pool.alloc(sizeof(H) * P);
    CK(cudaMemcpyAsync(dsend, &mine, sizeof(H), cudaMemcpyHostToDevice, g_stream));
    NK(g_nccl.AllGather(dsend, drecv, sizeof(H), ncclUint8, g_comm, g_stream));      // also: every rank's flags are zero before anybody writes one
    std::vector<H> all(P);
    CK(cudaMemcpyAsync(all.data(), drecv, sizeof(H) * P, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    dfree(dsend); dfree(drecv);
    for (uint32_t r = 0; r < P; ++r) ok &= all[r].good == 1;
    for (uint32_t r = 0; r < P && ok; ++r) {
        if ((int)r == g_rank) { g_pb.peers.p[r] = g_pb.mine; continue; }
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        g_pb.peers.p[r] = (unsigned long long*)q;
    }
    uint64_t good = ok ? 1 : 0;
    std::vector<uint64_t> ag;
    allgather8_host(&good, ag);
    for (uint64_t v : ag) ok &= v != 0;
    g_pb.ok = ok;
}
static void stream_barrier(cudaStream_t st) {
    if (g_pb.ok) {
        ++g_pb.epoch;
        peer_barrier_kernel<<<1, 32, 0, st>>>(g_pb.mine, g_pb.peers, (uint32_t)g_rank, (uint32_t)g_nranks, g_pb.epoch); LAUNCH_CHECK();
    } else nccl_stream_barrier(st);
}
// the stream the halo pushes and their barriers run on (high priority: a push kernel gets SM slots next to the sweeps)
static cudaStream_t halo_stream() {
    static cudaStream_t st = nullptr;
    if (!st) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        CK(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, hi));
    }
    return st;
}

// (Re)maps the peers' state buffers of agent type t when some rank's layout may have changed (collective).  Returns false when the
// peer-memory path is not available (more than 16 ranks, IPC refused): the caller falls back to ncclSend/ncclRecv.
// The check itself costs a host-synchronising all-gather, so it only runs when `peer_check` is set: after finish_init!, after an
// apply! that can add agents or edges (births grow buffers, received edges grow ghost tables) and after host-side additions — the
// same events on every rank.  A static network (the reference skips the exchange of unchanged agents the same way, src/MPI.jl:178-179)
// never pays for it again.
bool vb_sim::refresh_peer_map(int t) {
    AgentStore& a = A(t);
    static const bool enabled = !(getenv("VB_HALO_P2P") && atoi(getenv("VB_HALO_P2P")) == 0);
    const uint32_t P = (uint32_t)g_nranks;
    if (!enabled || P > 16) return false;
    if (!a.peers.check && a.peers.sig != 0) {
        // no collective event since the last exchange: the layout must be what the peers know (a buffer reallocated by a rank-local
        // call in between would make them push into freed memory)
        uint64_t sig0 = 1469598103934665603ull;
        auto mix0 = [&](uint64_t v) { sig0 = (sig0 ^ v) * 1099511628211ull; };
        mix0((uint64_t)(uintptr_t)a.state[0]); mix0((uint64_t)(uintptr_t)a.state[1]); mix0(a.stride()); mix0(a.cap); mix0(a.nghost);
        for (uint32_t r = 0; r <= P; ++r) mix0(a.ghost_off[r]);
        if (sig0 == 0) sig0 = 1;
        if (sig0 != a.peers.sig) throw CudaError("multi-rank: the buffers of agent type " + a.name + " changed outside of a collective operation");
        return a.peers.ok;
    }
    a.peers.check = false;
    peer_barrier_setup();
    // layout signature of this rank: buffers, stride, ghost table
    uint64_t sig = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { sig = (sig ^ v) * 1099511628211ull; };
    mix((uint64_t)(uintptr_t)a.state[0]); mix((uint64_t)(uintptr_t)a.state[1]); mix(a.stride()); mix(a.cap); mix(a.nghost);
    for (uint32_t r = 0; r <= P; ++r) mix(a.ghost_off[r]);
    if (sig == 0) sig = 1;
    uint64_t changed = (a.peers.sig != sig) ? 1 : 0;
    std::vector<uint64_t> all;
    allgather8_host(&changed, all);
    bool any = false;
    for (uint64_t v : all) any |= v != 0;
    if (!any) return a.peers.ok;
    // exchange {ipc handles of both buffers, stride, cap, ghost_off[], wanted halo phases}
    struct Desc { cudaIpcMemHandle_t h[2]; uint32_t same, stride, cap, pad; uint32_t ghost_off[17]; uint32_t ng_want; };
    static_assert(sizeof(Desc) % 8 == 0, "descriptor is exchanged in 8-byte units");
    Desc mine{};
    bool ok = true;
    for (int b = 0; b < 2; ++b) {
        if (b == 1 && a.state[1] == a.state[0]) { mine.same = 1; mine.h[1] = mine.h[0]; continue; }
        if (cudaIpcGetMemHandle(&mine.h[b], a.state[b]) != cudaSuccess) { cudaGetLastError(); ok = false; }
    }
    mine.stride = a.stride(); mine.cap = a.cap; mine.pad = ok ? 1u : 0u;
    for (uint32_t r = 0; r <= P; ++r) mine.ghost_off[r] = a.ghost_off[r];
    {   // halo phases = ghost blocks of the prefiltered sweeps: one per VB_KEY_BLOCK_MB of ghost keys (every further block costs a pass
        // over the rows' offsets, states and parked accumulators).  VB_HALO_PHASES overrides.
        static const double key_mb = getenv("VB_KEY_BLOCK_MB") ? atof(getenv("VB_KEY_BLOCK_MB")) : 52.0;
        static const int env_phases = getenv("VB_HALO_PHASES") ? atoi(getenv("VB_HALO_PHASES")) : 0;
        uint32_t want = (uint32_t)std::max<double>(1.0, std::ceil((double)a.nghost / std::max(1.0, key_mb * 1e6)));
        if (env_phases > 0) want = (uint32_t)env_phases;
        mine.ng_want = std::min(want, 16u);
    }
    uint8_t* dsend = (uint8_t*)g_pool.alloc(sizeof(Desc)); uint8_t* drecv = (uint8_t*)g_pool.alloc(sizeof(Desc) * P);
    CK(cudaMemcpyAsync(dsend, &mine, sizeof(Desc), cudaMemcpyHostToDevice, g_stream));
    NK(g_nccl.AllGather(dsend, drecv, sizeof(Desc), ncclUint8, g_comm, g_stream));
    std::vector<Desc> descs(P);
    CK(cudaMemcpyAsync(descs.data(), drecv, sizeof(Desc) * P, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    dfree(dsend); dfree(drecv);
    for (void* q : a.peers.opened) cudaIpcCloseMemHandle(q);
    cudaGetLastError();
    a.peers.opened.clear();
    a.peers.base[0].assign(P, nullptr); a.peers.base[1].assign(P, nullptr); a.peers.stride.assign(P, 0); a.peers.ghost0.assign(P, 0);
    uint32_t ng = 1;
    for (uint32_t r = 0; r < P; ++r) { ok &= descs[r].pad == 1u; ng = std::max(ng, descs[r].ng_want); }
    a.peers.ng = ng;
    for (uint32_t r = 0; r < P && ok; ++r) {
        if (r == rank) continue;
        for (int b = 0; b < 2; ++b) {
            if (b == 1 && descs[r].same) { a.peers.base[1][r] = a.peers.base[0][r]; continue; }
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, descs[r].h[b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
            a.peers.opened.push_back(q);
            a.peers.base[b][r] = (uint8_t*)q;
        }
        a.peers.stride[r] = descs[r].stride;
        a.peers.ghost0[r] = descs[r].cap + descs[r].ghost_off[rank];     // my agents' slots in peer r's ghost segment
    }
    // every rank must take the same path
    uint64_t good = ok ? 1 : 0;
    allgather8_host(&good, all);
    for (uint64_t v : all) ok &= v != 0;
    a.peers.ok = ok;
    a.peers.sig = sig;
    ++layout_epoch;        // ghost block sizes may have changed: blocked views over [local | ghost] slots are rebuilt
    return ok;
}

// transmit_agents! (src/MPI.jl:155-267) for agent type t.  Peer-memory path: the exchange runs on the halo stream in `ng` phases —
//   barrier (no peer still reads last step's ghosts) -> for every ghost block j: pack + push in one kernel -> barrier (phase j has
//   landed everywhere) -> event j —
// and returns at once; halo_wait(t, j + 1) makes the main stream wait for phases 0..j.  The pushes read this rank's READ buffer, which
// no kernel of the apply writes (states are double buffered; :Independent types are joined before the first kernel), so sweeps over
// the local agents and over ghost blocks that have already arrived overlap with the blocks still on the wire.
void vb_sim::halo_exchange(int t) {
    AgentStore& a = A(t);
    a.halo_pending = 0; a.halo_waited = 0;
    if (g_nranks <= 1 || !a.halo_dirty) return;
    a.halo_dirty = false;
    if (!a.size || a.send_off.empty()) return;
    const uint32_t P = (uint32_t)g_nranks, ns = a.send_off[P];
    if (refresh_peer_map(t)) {
        cudaStream_t hs = halo_stream();
        static cudaEvent_t ev_begin = nullptr;
        if (!ev_begin) CK(cudaEventCreateWithFlags(&ev_begin, cudaEventDisableTiming));
        const uint32_t ng = a.peers.ng;
        while (a.ev_phase.size() < ng) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); a.ev_phase.push_back(e); }
        g_trace.mark("halo exchange of " + a.name + " starts");
        CK(cudaEventRecord(ev_begin, g_stream));
        CK(cudaStreamWaitEvent(hs, ev_begin, 0));
        g_trace.mark("begin", hs);
        stream_barrier(hs);
        g_trace.mark("barrier: every rank is here", hs);
        if (!ev_halo[0]) { CK(cudaEventCreate(&ev_halo[0])); CK(cudaEventCreate(&ev_halo[1])); }
        CK(cudaEventRecord(ev_halo[0], hs));
        for (uint32_t j = 0; j < ng; ++j) {
            HaloPushArgs h{};
            h.cols = a.rstate(); h.stride = a.stride(); h.slots = a.send_slots; h.word = a.word; h.ncols = a.ncols; h.npeers = P - 1;
            for (uint32_t k = 0; k + 1 < P; ++k) {
                const uint32_t r = (rank + 1 + k) % P;
                // the j-th part of the range peer r mirrors of my agents (the part belongs to its ghost key block j)
                uint32_t lo = 0, hi = 0;
                AgentStore::ghost_part_range(a.send_off[r + 1] - a.send_off[r], ng, j, lo, hi);
                h.voff[k + 1] = h.voff[k] + (uint32_t)(hi - lo);
                h.first[k] = a.send_off[r] + (uint32_t)lo;
                h.remote[k] = a.peers.base[a.cur][r]; h.rstride[k] = a.peers.stride[r]; h.rghost0[k] = a.peers.ghost0[r] + (uint32_t)lo;
            }
            h.n = h.voff[P - 1];
            // VB_HALO_CE=1 (default): pack this phase's parts into the send buffer, then one peer copy per peer and column on the copy
            // engines — the transfer occupies no SM, so the sweeps that run beside it keep the whole device (with the push kernel a
            // local sweep ran at a third of its rate while a phase was on the wire, profiles/r2_scaling.md).  0: one kernel packs and pushes.
            static const int env_ce = getenv("VB_HALO_CE") ? atoi(getenv("VB_HALO_CE")) : -1;
            // default: the first phase by the push kernel (nothing runs beside it: the first sweep waits for it; 560 GB/s against 430 GB/s
            // of the copy engines on 4 B200), the later phases, which overlap with sweeps, by the copy engines
            const bool use_ce = env_ce >= 0 ? env_ce != 0 : j > 0;
            if (h.n && use_ce && a.send_buf) {
                halo_pack_ranges_kernel<<<nblk((uint64_t)h.n * a.ncols), 256, 0, hs>>>(h, a.send_buf, ns); LAUNCH_CHECK();
                for (uint32_t k = 0; k + 1 < P; ++k) {
                    const uint32_t cnt = h.voff[k + 1] - h.voff[k];
                    if (!cnt) continue;
                    for (uint32_t c = 0; c < a.ncols; ++c)
                        CK(cudaMemcpyAsync(h.remote[k] + ((size_t)c * h.rstride[k] + h.rghost0[k]) * a.word, a.send_buf + ((size_t)c * ns + h.first[k]) * a.word,
                                           (size_t)cnt * a.word, cudaMemcpyDeviceToDevice, hs));
                }
            } else if (h.n) { halo_push_kernel<<<nblk((uint64_t)h.n * a.ncols), 256, 0, hs>>>(h); LAUNCH_CHECK(); }
            g_trace.mark("push of phase " + std::to_string(j) + ": " + std::to_string((uint64_t)h.n * a.size) + " B", hs);
            stream_barrier(hs);
            g_trace.mark("barrier: phase landed everywhere", hs);
            CK(cudaEventRecord(a.ev_phase[j], hs));
        }
        CK(cudaEventRecord(ev_halo[1], hs));
        halo_timed = true;
        a.halo_pending = ng;
        halo_bytes += (uint64_t)a.nghost * a.size;
        (void)ns;
        return;
    }
    if (ns) { halo_pack_kernel<<<nblk((uint64_t)ns * a.ncols), 256, 0, g_stream>>>(a.rstate(), a.stride(), a.send_slots, ns, a.send_buf, a.word, a.ncols); LAUNCH_CHECK(); }
    NK(g_nccl.GroupStart());
    for (uint32_t c = 0; c < a.ncols; ++c) {
        for (uint32_t r = 0; r < P; ++r) {
            if (r == rank) continue;
            const uint32_t give = a.send_off[r + 1] - a.send_off[r], want = a.ghost_off[r + 1] - a.ghost_off[r];
            if (give) NK(g_nccl.Send(a.send_buf + ((size_t)c * ns + a.send_off[r]) * a.word, (size_t)give * a.word, ncclUint8, (int)r, g_comm, g_stream));
            if (want) NK(g_nccl.Recv(a.rstate() + ((size_t)c * a.stride() + a.cap + a.ghost_off[r]) * a.word, (size_t)want * a.word, ncclUint8, (int)r, g_comm, g_stream));
        }
    }
    NK(g_nccl.GroupEnd());
    halo_bytes += (uint64_t)a.nghost * a.size;
}
void vb_sim::halo_wait(int t, uint32_t phases) {
    AgentStore& a = A(t);
    phases = std::min(phases, a.halo_pending);
    for (; a.halo_waited < phases; ++a.halo_waited) CK(cudaStreamWaitEvent(g_stream, a.ev_phase[a.halo_waited], 0));
}
void vb_sim::halo_wait_all() {
    for (size_t t = 1; t <= agents.size(); ++t) if (agents[t - 1].halo_pending) halo_wait((int)t, agents[t - 1].halo_pending);
}

// rows of an implicit raster stencil, enumerated exactly like Ctx::stencil_row on the device
std::vector<uint32_t> vb_sim::stencil_row_host(const EdgeStore& e, uint64_t lin) const {
    const RasterStore& r = rasters[e.st_raster];
    const int nd = (int)r.dims.size();
    std::vector<int64_t> pos(nd), stride(nd);
    uint64_t rest = lin; int64_t st = 1;
    for (int k = 0; k < nd; ++k) { pos[k] = (int64_t)(rest % (uint64_t)r.dims[k]); rest /= (uint64_t)r.dims[k]; stride[k] = st; st *= r.dims[k]; }
    std::vector<std::pair<uint64_t, int>> keys;
    for (int si = 0; si < e.st_n; ++si) {
        int64_t l = 0; bool ok = true;
        for (int k = 0; k < nd; ++k) {
            int64_t v = pos[k] - e.st_off_host[(size_t)si * vb::MAX_RASTER_DIMS + k];
            if (v < 0 || v >= r.dims[k]) { if (!e.st_periodic) { ok = false; break; } v %= r.dims[k]; if (v < 0) v += r.dims[k]; }
            l += v * stride[k];
        }
        if (ok) keys.push_back({(uint64_t)l, si});
    }
    std::sort(keys.begin(), keys.end());
    std::vector<uint32_t> out;
    for (auto& k : keys) out.push_back((uint32_t)k.first);
    return out;
}
uint64_t vb_sim::stencil_total(const EdgeStore& e) const {
    const RasterStore& r = rasters[e.st_raster];
    uint64_t total = 0;
    for (int si = 0; si < e.st_n; ++si) {
        uint64_t c = 1;
        for (size_t k = 0; k < r.dims.size(); ++k) {
            const int64_t o = std::llabs((long long)e.st_off_host[(size_t)si * vb::MAX_RASTER_DIMS + k]);
            c *= e.st_periodic ? (uint64_t)r.dims[k] : (uint64_t)std::max<int64_t>(0, r.dims[k] - o);
        }
        total += c;
    }
    return total;
}
void vb_sim::materialize_stencil(int ei) {
    EdgeStore& e = E(ei);
    if (!e.implicit_stencil) return;
    e.implicit_stencil = false;
    dfree(e.st_off); e.st_off = nullptr;
    emit_raster_edges(ei, e.st_raster, e.st_distance, e.st_metric, e.st_periodic, nullptr);
    if (initialized) merge_pending(ei);
}

uint64_t vb_sim::edge_total(int ei, bool write) {
    EdgeStore& e = E(ei);
    if (e.implicit_stencil) return (initialized || write) ? stencil_total(e) : 0;
    // before finish_init! everything lives in the write container = raw adds (Edge.jl:376-380)
    if (!initialized) {
        if (!write) return 0;
        if (e.singleedge && e.chunks.empty()) {   // one slot per target: count distinct targets
            std::vector<uint64_t> t = e.h_to;
            std::sort(t.begin(), t.end());
            return (uint64_t)(std::unique(t.begin(), t.end()) - t.begin());
        }
        return e.raw_n;
    }
    merge_pending(ei);
    if (e.kind == vb::KIND_CSR) return e.nnz;
    if (!e.cnt || !e.rows) return 0;
    // sum of the per-row counts
    MapArgs ma{};
    ma.cols = (const uint8_t*)e.cnt; ma.stride = e.rows; ma.word = 4; ma.n = e.rows; ma.died = nullptr; ma.offset = 0; ma.dt = vb::DT_I32; ma.op = vb::OP_SUM;
    long long* part = dalloc<long long>(1024 + 1);
    const unsigned nb = std::min<unsigned>(1024, nblk(e.rows));
    mapreduce_kernel<false><<<nb, 256, 0, g_stream>>>(ma, part); LAUNCH_CHECK();
    mapreduce_final_kernel<false><<<1, 256, 0, g_stream>>>(part, nb, vb::OP_SUM, part + 1024); LAUNCH_CHECK();
    long long r = 0;
    CK(cudaMemcpyAsync(&r, part + 1024, 8, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    dfree(part);
    return (uint64_t)r;
}

// ------------------------------------------------------------------------------------------------------
namespace {

template <class F>
int guard(F&& f) {
    try { f(); return VB_OK; }
    catch (const AssertionError& e) { g_err = e.what(); return VB_ERR_ASSERT; }
    catch (const ArgError& e) { g_err = e.what(); return VB_ERR_ARG; }
    catch (const CudaError& e) { g_err = e.what(); return VB_ERR_CUDA; }
    catch (const std::exception& e) { g_err = e.what(); return VB_ERR_STATE; }
}
bool contains(const std::vector<int>& v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); }

void finish_write_agent(vb_sim& s, int t, std::vector<uint32_t*>& died_flags, std::vector<uint32_t>& died_n, std::vector<uint32_t>& died_cnt) {
    AgentStore& a = s.A(t);
    // births of this apply: pops from the reuse stack first, then fresh slots (AgentMethods.jl:37-63)
    const uint32_t pops = std::min(a.births, a.n_reuse);
    const uint32_t fresh = a.births - pops;
    const uint32_t n_before = a.nslots;
    a.n_reuse -= pops;
    a.nextid += fresh;
    a.nslots = (uint32_t)std::max<uint64_t>(a.nslots, a.nextid - 1);
    a.births = 0;
    if (!a.immortal && n_before > 0 && s.initialized) {
        // nobody died (the common case for models without deaths whose types are not registered :Immortal, e.g. the docs' HK model):
        // one counting pass over the two died arrays instead of flags + scan + compact
        uint32_t* dcount = s.d_scalars + 60;
        CK(cudaMemsetAsync(dcount, 0, 4, g_stream));
        vbp::count_newly_died_kernel<<<std::min<unsigned>(nblk((uint64_t)n_before / 16 + 1), 148 * 8), 256, 0, g_stream>>>(a.rdied(), a.wdied(), n_before, dcount); LAUNCH_CHECK();
        uint32_t any_died = 0;
        CK(cudaMemcpyAsync(&any_died, dcount, 4, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        if (any_died) {
            // agents that died in this apply, in ascending slot order, are appended to read.reuseable (:171,:430)
            uint32_t* flag = dalloc<uint32_t>(n_before); uint32_t* pos = dalloc<uint32_t>(n_before);
            uint32_t* scr = dalloc<uint32_t>(vbp::scan_scratch_words(n_before));
            vbp::newly_died_flags_kernel<<<nblk(n_before), 256, 0, g_stream>>>(a.rdied(), a.wdied(), n_before, flag); LAUNCH_CHECK();
            vbp::exclusive_scan(flag, pos, n_before, s.d_scalars, scr, g_stream); g_launches += 3;
            uint32_t nd = 0;
            CK(cudaMemcpyAsync(&nd, s.d_scalars, 4, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            if (nd) {
                vbp::compact_indices_kernel<<<nblk(n_before), 256, 0, g_stream>>>(flag, pos, n_before, a.reuse + a.n_reuse); LAUNCH_CHECK();
                a.n_reuse += nd;
                died_flags[t] = flag; died_n[t] = n_before; died_cnt[t] = nd;
                flag = nullptr;
            }
            dfree(flag); dfree(pos); dfree(scr);
        }
    }
    a.cur ^= 1;   // read := write by swapping the double buffers (:Independent types share one state buffer)
    a.write_stale = true;
    a.halo_dirty = true;
    a.last_change = s.num_transitions;
    a.writeable = false;
}

void do_apply(vb_sim& s, const std::string& tname, const std::vector<int>& call, const std::vector<int>& read, const std::vector<int>& write,
              const std::vector<int>& add_existing, int with_edge, uint64_t seed) {
    require_device();
    if (!s.initialized) throw AssertionError("You must call finish_init! before apply!");
    if (with_edge >= 0 && s.E(with_edge).singletype) throw AssertionError("The `with_edge` keyword can only be used for edgetypes without the :SingleType hint");
    for (int c : call) if (contains(add_existing, c)) throw AssertionError("a `call` type can not be element of `add_existing`");
    for (int ae : add_existing) if (!contains(write, ae)) throw AssertionError("type is in `add_existing` but not in `write`");
    std::vector<const vb::TransitionInfo*> tis;
    for (int c : call) {
        if (c >= vb::EDGE_REF) throw ArgError("`call` must list agent types");
        AgentStore& a = s.A(c);
        auto it = registry().find({tname, a.name});
        if (it == registry().end()) throw ArgError("transition '" + tname + "' is not registered for agent type " + a.name);
        const vb::TransitionInfo* ti = it->second;
        if (a.size && ti->state_size != a.size) throw ArgError("transition '" + tname + "': sizeof(State) does not match the registered size of " + a.name);
        for (int i = 0; i < ti->n_edge_writes; ++i)   // _can_add: EdgeMethods.jl:258-265
            if (s.asserts_enabled && s.check_readable && !contains(write, vb::EDGE_REF + ti->edge_writes[i]))
                throw AssertionError("edge type " + s.E(ti->edge_writes[i]).name + " must be in the `write` argument of the transition function");
        for (int i = 0; i < ti->n_edge_removes; ++i) {   // _can_remove_edges: EdgeMethods.jl:101-121
            const int er = vb::EDGE_REF + ti->edge_removes[i];
            if (s.asserts_enabled && s.check_readable && !contains(write, er))
                throw AssertionError("Edge of " + s.E(ti->edge_removes[i]).name + " can not removed, as it is not in the `write` argument of the transition function");
            if (s.asserts_enabled && s.check_readable && !contains(add_existing, er))
                throw AssertionError(s.E(ti->edge_removes[i]).name + " must be in the `add_existing` keyword of the transition function");
        }
        for (int i = 0; i < ti->n_agent_writes; ++i)  // add_agent!: AgentMethods.jl:71-77
            if (s.asserts_enabled && !contains(write, ti->agent_writes[i]))
                throw AssertionError("agent type " + s.A(ti->agent_writes[i]).name + " must be in the `write` argument of the transition function");
        tis.push_back(ti);
    }
    for (int w : write) if (w >= vb::EDGE_REF) s.materialize_stencil(w - vb::EDGE_REF);
    if (with_edge >= 0) s.materialize_stencil(with_edge);
    s.merge_all_pending();
    g_trace.begin();
    const unsigned long long launches0 = g_launches;
    s.intransition = true;
    struct Reset {
        vb_sim& s; const std::vector<int>& read;
        ~Reset() {
            s.intransition = false;
            for (int r : read) { if (r >= vb::EDGE_REF) s.edges[r - vb::EDGE_REF].readable = false; else s.agents[r - 1].prepared = false; }
            for (auto& a : s.agents) { a.writeable = false; a.births = 0; }
            for (auto& e : s.edges) e.writeable = false;
        }
    } reset{s, read};
    for (int r : read) { if (r >= vb::EDGE_REF) s.E(r - vb::EDGE_REF).readable = true; else s.A(r).prepared = true; }

    // ---- prepare_write! ----
    for (int w : write) {
        const bool ae = contains(call, w) || contains(add_existing, w);
        if (w < vb::EDGE_REF) {                                                // AgentMethods.jl:271-298
            AgentStore& a = s.A(w);
            if (a.immortal && !ae) throw AssertionError("an :Immortal type in `write` must also be in `add_existing` (or `call`)");
            if (!ae) { a.nslots = 0; a.nextid = 1; a.n_reuse = 0; }
            a.writeable = true;
            a.births = 0;
            if (a.nslots) {
                if (!a.immortal) CK(cudaMemcpyAsync(a.wdied(), a.rdied(), a.nslots, cudaMemcpyDeviceToDevice, g_stream));
                const bool full_call = contains(call, w) && with_edge < 0;
                if (!a.independent && a.size && a.write_stale && !full_call) {
                    CK(cudaMemcpyAsync(a.wstate(), a.rstate(), (size_t)a.stride() * a.size, cudaMemcpyDeviceToDevice, g_stream));
                }
            }
        } else {                                                               // EdgeMethods.jl:639-663
            EdgeStore& e = s.E(w - vb::EDGE_REF);
            e.writeable = true;
            e.add_existing = contains(add_existing, w);
            e.log_n = 0; e.rm_n = 0; e.rlog_n = 0;
            e.ordered_log = false;
            for (auto* ti : tis) for (int i = 0; i < ti->n_edge_removes; ++i) if (ti->edge_removes[i] == w - vb::EDGE_REF) e.ordered_log = e.kind != vb::KIND_CSR;
            if (e.kind != vb::KIND_CSR) {
                const uint32_t rows = s.rows_of(e);
                dfree(e.wcnt);
                e.wcnt = dalloc<uint32_t>((size_t)rows + 1); e.rows_w = rows;
                CK(cudaMemsetAsync(e.wcnt, 0, ((size_t)rows + 1) * 4, g_stream));
                if (e.add_existing && e.cnt) CK(cudaMemcpyAsync(e.wcnt, e.cnt, (size_t)std::min(rows, e.rows) * 4, cudaMemcpyDeviceToDevice, g_stream));
            }
        }
    }
    g_trace.end("prepare_write!", tname);
    CK(cudaMemsetAsync(s.d_stats, 0, 4096 * 8, g_stream));
    CK(cudaEventRecord(s.ev[0], g_stream));
    s.halo_bytes = 0;
    for (int r : read) if (r < vb::EDGE_REF) s.halo_exchange(r);   // prepare_read!: transmit_agents! (AgentMethods.jl:484-496)
    if (g_nranks > 1) g_trace.end("halo exchange", tname);
    s.st_agents_called = 0;
    s.ms_kernel = 0;
    bool kernel_time_pending = false;
    uint64_t appended = 0;

    // ---- the transition loop over `call` (Simulation.jl:774-788) ----
    for (size_t ci = 0; ci < call.size(); ++ci) {
        const int C = call[ci];
        const vb::TransitionInfo* ti = tis[ci];
        AgentStore& a = s.A(C);
        const uint32_t n = a.nslots;
        if (n == 0) continue;
        s.st_agents_called += n;
        // the halo of this apply may still be on the wire (halo stream): only the segmented prefiltered sweeps below start before it has
        // landed (local blocks first, a ghost block when its phase is there); every other form waits here
        static const bool env_overlap = !(getenv("VB_HALO_OVERLAP") && atoi(getenv("VB_HALO_OVERLAP")) == 0);
        const bool pipelined = env_overlap && ti->reduce && ti->prefilter && ti->n_edge_writes + ti->n_agent_writes + ti->n_edge_removes == 0 && call.size() == 1 &&
                               !a.independent && with_edge < 0;
        if (!pipelined) s.halo_wait_all();
        vb::LaunchArgs la{};
        la.ds = &s.h_ds; la.type = C; la.n = n; la.in_read = contains(read, C); la.in_write = contains(write, C);
        la.with_edge = with_edge; la.stats = s.d_stats; la.stream = g_stream;
        const int nw = ti->n_edge_writes + ti->n_agent_writes + ti->n_edge_removes;
        std::vector<uint32_t*> tmp;
        if (nw > 0) {
            set_persisting_l2_mb(0);
            // count pass -> exclusive scans -> totals
            uint32_t* scr = dalloc<uint32_t>(vbp::scan_scratch_words(n));
            tmp.push_back(scr);
            for (int i = 0; i < ti->n_edge_writes; ++i) { la.ecount[i] = dalloc<uint32_t>(n); tmp.push_back(la.ecount[i]); }
            for (int i = 0; i < ti->n_agent_writes; ++i) { la.acount[i] = dalloc<uint32_t>(n); tmp.push_back(la.acount[i]); }
            for (int i = 0; i < ti->n_edge_removes; ++i) { la.rcount[i] = dalloc<uint32_t>(n); tmp.push_back(la.rcount[i]); }
            if (g_nranks > 1) for (int i = 0; i < ti->n_edge_writes; ++i) { la.ercount[i] = dalloc<uint32_t>(n); tmp.push_back(la.ercount[i]); }
            s.upload_view(seed);
            la.mode = vb::MODE_COUNT;
            CK(ti->launch(la)); ++g_launches;
            for (int i = 0; i < ti->n_edge_writes; ++i) { vbp::exclusive_scan(la.ecount[i], la.ecount[i], n, s.d_scalars + i, scr, g_stream); g_launches += 3; }
            for (int i = 0; i < ti->n_agent_writes; ++i) { vbp::exclusive_scan(la.acount[i], la.acount[i], n, s.d_scalars + vb::MAX_EDGE_WRITES + i, scr, g_stream); g_launches += 3; }
            constexpr int RM0 = vb::MAX_EDGE_WRITES + vb::MAX_AGENT_WRITES;
            for (int i = 0; i < ti->n_edge_removes; ++i) { vbp::exclusive_scan(la.rcount[i], la.rcount[i], n, s.d_scalars + RM0 + i, scr, g_stream); g_launches += 3; }
            constexpr int ER0 = RM0 + vb::MAX_EDGE_REMOVES;
            if (g_nranks > 1) for (int i = 0; i < ti->n_edge_writes; ++i) { vbp::exclusive_scan(la.ercount[i], la.ercount[i], n, s.d_scalars + ER0 + i, scr, g_stream); g_launches += 3; }
            uint32_t totals[vb::MAX_EDGE_WRITES + vb::MAX_AGENT_WRITES + vb::MAX_EDGE_REMOVES + vb::MAX_EDGE_WRITES] = {0};
            CK(cudaMemcpyAsync(totals, s.d_scalars, sizeof(totals), cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            s.check_device_error("apply! (count pass)");
            // births: make room (may shift composite bases -> rebase), then logs
            for (int i = 0; i < ti->n_agent_writes; ++i) {
                AgentStore& b = s.A(ti->agent_writes[i]);
                la.abase[i] = b.births;
                const uint32_t total_births = b.births + totals[vb::MAX_EDGE_WRITES + i];
                const uint32_t pops = std::min(total_births, b.n_reuse);
                s.ensure_agent_cap(ti->agent_writes[i], (b.nextid - 1) + (total_births - pops));
            }
            for (int i = 0; i < ti->n_edge_writes; ++i) {
                EdgeStore& e = s.E(ti->edge_writes[i]);
                la.ebase[i] = e.log_n;
                if (e.kind == vb::KIND_CSR || e.ordered_log) s.ensure_log(e, (uint64_t)e.log_n + totals[i] + 1);
                if (e.kind != vb::KIND_CSR && s.rows_of(e) != e.rows_w) throw CudaError("internal: count container not sized");
            }
            if (g_nranks > 1) for (int i = 0; i < ti->n_edge_writes; ++i) {   // edges that leave the rank
                EdgeStore& e = s.E(ti->edge_writes[i]);
                la.erbase[i] = e.rlog_n;
                const uint64_t need = (uint64_t)e.rlog_n + totals[ER0 + i];
                if (need > e.rlog_cap) {
                    const uint32_t ncap = (uint32_t)std::max<uint64_t>(need, (uint64_t)e.rlog_cap * 2 + 1024);
                    uint64_t* nt = dalloc<uint64_t>(ncap); uint64_t* nf = e.has_src() ? dalloc<uint64_t>(ncap) : nullptr;
                    uint32_t* nd = dalloc<uint32_t>(ncap); uint8_t* ns = e.has_state() ? (uint8_t*)g_pool.alloc((size_t)ncap * e.size) : nullptr;
                    if (e.rlog_n) {
                        CK(cudaMemcpyAsync(nt, e.rlog_to, (size_t)e.rlog_n * 8, cudaMemcpyDeviceToDevice, g_stream));
                        if (nf) CK(cudaMemcpyAsync(nf, e.rlog_from, (size_t)e.rlog_n * 8, cudaMemcpyDeviceToDevice, g_stream));
                        CK(cudaMemcpyAsync(nd, e.rlog_dst, (size_t)e.rlog_n * 4, cudaMemcpyDeviceToDevice, g_stream));
                        if (ns) { vbp::soa_copy_kernel<<<nblk((uint64_t)e.rlog_n * e.size), 256, 0, g_stream>>>(e.rlog_st, e.rlog_cap, ns, ncap, e.rlog_n, e.ncols, e.word, 0, 0); LAUNCH_CHECK(); }
                    }
                    dfree(e.rlog_to); dfree(e.rlog_from); dfree(e.rlog_dst); dfree(e.rlog_st);
                    e.rlog_to = nt; e.rlog_from = nf; e.rlog_dst = nd; e.rlog_st = ns; e.rlog_cap = ncap;
                }
            }
            for (int i = 0; i < ti->n_edge_removes; ++i) {
                EdgeStore& e = s.E(ti->edge_removes[i]);
                la.rbase[i] = e.rm_n; la.rmark[i] = e.log_n;
                s.ensure_rm(e, (uint64_t)e.rm_n + totals[RM0 + i]);
            }
            s.upload_view(seed);
            la.mode = vb::MODE_EMIT;
            CK(cudaEventRecord(s.evk[0], g_stream));
            CK(ti->launch(la)); ++g_launches;
            CK(cudaEventRecord(s.evk[1], g_stream));
            for (int i = 0; i < ti->n_edge_writes; ++i) { EdgeStore& e = s.E(ti->edge_writes[i]); if (e.kind == vb::KIND_CSR || e.ordered_log) e.log_n += totals[i]; appended += totals[i]; }
            for (int i = 0; i < ti->n_edge_removes; ++i) s.E(ti->edge_removes[i]).rm_n += totals[RM0 + i];
            if (g_nranks > 1) for (int i = 0; i < ti->n_edge_writes; ++i) { s.E(ti->edge_writes[i]).rlog_n += totals[ER0 + i]; appended += totals[ER0 + i]; }
            for (int i = 0; i < ti->n_agent_writes; ++i) s.A(ti->agent_writes[i]).births += totals[vb::MAX_EDGE_WRITES + i];
        } else {
            la.mode = vb::MODE_DIRECT;
            la.primary_edge = -1; la.heavy_min = 0; la.group = 0; la.rows = nullptr;
            uint32_t heavy_n = 0; const uint32_t* heavy_rows = nullptr;
            bool blocked = false, stencil_reduce = false;
            if (ti->cooperative && ti->primary_edge >= 0 && with_edge < 0) {
                // degree binning (north_star: sub-warp / warp per agent, block per agent for rows >= 1024 entries)
                EdgeStore& pe = s.E(ti->primary_edge);
                if (pe.implicit_stencil) { la.group = 1; stencil_reduce = ti->reduce && ti->launch_stencil; }   // grid stencil: a thread per cell, neighbour loads of a warp are adjacent
                else if (pe.kind == vb::KIND_CSR && pe.off && (!pe.singletype || pe.target == C)) {
                    // Reduce transitions over a static network whose source states dwarf L2 sweep the rows once per L2-sized
                    // source block; only hub rows of >= 16384 entries are left to the block-per-agent pass there (the sweeps fold
                    // long rows warp-cooperatively), the direct path hands over rows of >= 1024 entries.
                    constexpr uint32_t HEAVY_DIRECT = 1024, HEAVY_BLOCKED = 16384;
                    blocked = ti->reduce && s.ensure_blocked(ti->primary_edge, ti, C, n, HEAVY_BLOCKED, &la, seed);
                    const uint32_t HEAVY_MIN = blocked ? HEAVY_BLOCKED : HEAVY_DIRECT;
                    const bool segmented = blocked && pe.blk.segmented;          // hub rows are cut into segments: no pass of their own
                    if (!segmented && (pe.heavy_version != pe.version || pe.heavy_type != C || pe.heavy_min != HEAVY_MIN)) {
                        dfree(pe.heavy_rows); pe.heavy_rows = nullptr; pe.heavy_n = 0;
                        uint32_t* flag = dalloc<uint32_t>(n); uint32_t* pos = dalloc<uint32_t>(n);
                        uint32_t* scr = dalloc<uint32_t>(vbp::scan_scratch_words(n));
                        heavy_flags_kernel<<<nblk(n), 256, 0, g_stream>>>(pe.off, pe.singletype ? 0u : s.base[C], n, pe.rows, HEAVY_MIN, flag); LAUNCH_CHECK();
                        vbp::exclusive_scan(flag, pos, n, s.d_scalars, scr, g_stream); g_launches += 3;
                        uint32_t hn = 0;
                        CK(cudaMemcpyAsync(&hn, s.d_scalars, 4, cudaMemcpyDeviceToHost, g_stream));
                        CK(cudaStreamSynchronize(g_stream));
                        if (hn) {
                            pe.heavy_rows = dalloc<uint32_t>(hn);
                            vbp::compact_indices_kernel<<<nblk(n), 256, 0, g_stream>>>(flag, pos, n, pe.heavy_rows); LAUNCH_CHECK();
                        }
                        pe.heavy_n = hn; pe.heavy_type = C; pe.heavy_version = pe.version; pe.heavy_min = HEAVY_MIN;
                        CK(cudaStreamSynchronize(g_stream));
                        dfree(flag); dfree(pos); dfree(scr);
                    }
                    if (!segmented) { heavy_n = pe.heavy_n; heavy_rows = pe.heavy_rows; }
                    la.primary_edge = ti->primary_edge; la.heavy_min = HEAVY_MIN;
                    const double avg = (double)pe.nnz / std::max<uint32_t>(1, n);
                    la.group = avg < 48.0 ? 8 : 32;
                }
            }
            if (blocked) {
                EdgeStore& pe = s.E(ti->primary_edge);
                const EdgeStore::Blocked& k = pe.blk;
                uint32_t swept = 0;
                static const bool use_lists = !(getenv("VB_BLOCK_LISTS") && atoi(getenv("VB_BLOCK_LISTS")) == 0);
                static const int l2_mb = getenv("VB_BLOCK_L2_MB") ? atoi(getenv("VB_BLOCK_L2_MB")) : 64;
                g_trace.mark("blocked read phase: host reaches the launch sequence");
                {   // room for the evict_last source block (set once: the call synchronises the device, which would serialise the halo stream)
                    set_persisting_l2_mb(l2_mb);
                }
                s.upload_view(seed);
                CK(cudaEventRecord(s.evk[0], g_stream));
                vb::LaunchArgs lb = la;
                lb.blk_src = k.bsrc; lb.blk_acc = k.acc; lb.blk_stride = k.rpad; lb.blk_heavy = heavy_n ? k.heavy_bits : nullptr;
                const bool pf = k.key != nullptr && s.prefilter_on(ti);
                lb.blk_key = pf ? k.key : nullptr; lb.blk_nkeys = pf ? k.key_n : 0; lb.blk_prefilter = pf ? 1 : 0;
                if (k.segmented) { lb.blk_seg_row = k.seg_row; lb.blk_nseg = k.nseg; lb.blk_heavy = nullptr; }
                if (!k.segmented) s.halo_wait_all();
                else for (size_t t2 = 1; t2 <= s.agents.size(); ++t2) if ((int)t2 != ti->source_type) s.halo_wait((int)t2, s.agents[t2 - 1].halo_pending);
                // keys of this step's read states (inside the timed region): the whole column at once, or block by block when the ghost
                // blocks arrive in phases
                g_trace.mark("sweeps begin");
                if (pf && !k.segmented) { lb.blk_key_first = 0; CK(ti->launch_keys(lb)); ++g_launches; g_trace.mark("keys"); }
                // sweeps over blocks that hold no entry (capacity beyond the agents in use) are skipped; the first sweep that runs
                // initialises the accumulators, the last one runs finish()
                std::vector<uint32_t> todo;
                for (uint32_t b = 0; b < k.nb; ++b) if (k.bstart[b + 1] > k.bstart[b]) todo.push_back(b);
                while (todo.size() < (pf ? 1u : 2u)) { uint32_t b = 0; while (std::find(todo.begin(), todo.end(), b) != todo.end()) ++b; todo.push_back(b); std::sort(todo.begin(), todo.end()); }
                // the hub rows left to the block-per-agent pass (a handful of CTAs with long dependent chains) run beside the sweeps
                // on a second stream: they read the same read buffer and write rows the last sweep skips
                static cudaStream_t side = nullptr;
                static cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
                if (heavy_n) {
                    if (!side) { CK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming)); }
                    CK(cudaEventRecord(ev_fork, g_stream));
                    CK(cudaStreamWaitEvent(side, ev_fork, 0));
                    vb::LaunchArgs lh = la;
                    lh.group = 256; lh.rows = heavy_rows; lh.n = heavy_n; lh.heavy_min = 0; lh.stream = side;
                    CK(ti->launch(lh)); ++g_launches;
                    CK(cudaEventRecord(ev_join, side));
                }
                for (size_t i = 0; i < todo.size(); ++i) {
                    lb.blk_off = k.boff + (size_t)todo[i] * k.rpad; lb.blk_first = i == 0; lb.blk_last = i + 1 == todo.size();
                    lb.blk_rows = use_lists ? k.arows[todo[i]] : nullptr; lb.blk_roff = k.aoff[todo[i]]; lb.blk_nrows = k.acount[todo[i]];
                    lb.blk_base = todo[i] * k.bsize;
                    if (k.segmented) {
                        const uint32_t b = todo[i];
                        if (k.ng && b >= k.gfirst) s.halo_wait(ti->source_type, b - k.gfirst + 1);      // a block with a ghost part: its phase of the halo must have landed
                        lb.blk_base = k.block_first(b);
                        g_trace.mark("(wait for the halo phase of block " + std::to_string(b) + ")");
                        if (b < k.nbl) {
                            lb.blk_key_first = b * k.bsize; lb.blk_nkeys = k.local_block_slots(b, s.A(ti->source_type).nslots);
                            CK(ti->launch_keys(lb)); ++g_launches;
                        }
                        if (k.ng && b >= k.gfirst) {                         // the part of every owner's range that belongs to this block
                            const uint32_t j = b - k.gfirst;
                            for (size_t p2 = 0; p2 + 1 < k.goff.size(); ++p2) {
                                uint32_t lo = 0, hi = 0;
                                AgentStore::ghost_part_range(k.goff[p2 + 1] - k.goff[p2], k.ng, j, lo, hi);
                                if (hi == lo) continue;
                                lb.blk_key_first = k.lcap + k.goff[p2] + lo; lb.blk_nkeys = hi - lo;
                                CK(ti->launch_keys(lb)); ++g_launches;
                            }
                        }
                        g_trace.mark("keys of block " + std::to_string(b));
                        lb.blk_nkeys = k.key_n;
                    }
                    CK(ti->launch_blocked(lb)); ++g_launches;
                    g_trace.mark("sweep of block " + std::to_string(todo[i]) + ": " + std::to_string(k.bstart[todo[i] + 1] - k.bstart[todo[i]]) + " entries");
                }
                s.halo_wait_all();
                if (k.segmented && k.nhub) {       // rows cut into several segments: merge their parked accumulators, finish
                    lb.blk_op = 1; lb.blk_hub_rows = k.hub_rows; lb.blk_hub_seg = k.hub_seg; lb.blk_nhub = k.nhub;
                    CK(ti->launch_blocked(lb)); ++g_launches;
                    g_trace.mark("hub rows: merge of " + std::to_string(k.nhub) + " rows");
                }
                swept = (uint32_t)todo.size();
                if (heavy_n) CK(cudaStreamWaitEvent(g_stream, ev_join, 0));
                CK(cudaEventRecord(s.evk[1], g_stream));
                CK(cudaStreamSynchronize(g_stream));
                cudaCtxResetPersistingL2Cache();
                cudaGetLastError();
                s.last_blocked_nb = swept; s.last_prefiltered = pf;
            } else {
            s.halo_wait_all();
            s.last_blocked_nb = 0; s.last_prefiltered = false;
            s.upload_view(seed);
            // Gather-bound read phases (state array far larger than L2): keep the head of the gathered type's state resident in
            // L2 for the duration of the launch.  Power-law / preferential-attachment graphs number their hubs first, so the head
            // takes a disproportionate share of the random gathers (profiles/l2_experiment.sh: 35.7 -> 33.0 ms on HK-100M).
            // VB_L2_PERSIST_MB overrides the window (0 disables).
            static const int persist_env = getenv("VB_L2_PERSIST_MB") ? atoi(getenv("VB_L2_PERSIST_MB")) : -1;
            const size_t state_bytes = (size_t)a.stride() * a.size;
            const int persist_mb = persist_env >= 0 ? persist_env : (ti->cooperative && ti->primary_edge >= 0 && state_bytes > ((size_t)256 << 20) ? 48 : 0);
            bool persisting = false;
            if (persist_mb == 0) set_persisting_l2_mb(0);      // hand a set-aside left by swept read phases back: sorts and scatters want the whole L2
            if (persist_mb > 0 && a.size) {
                set_persisting_l2_mb(persist_mb);
                cudaStreamAttrValue av{};
                av.accessPolicyWindow.base_ptr = a.rstate();
                av.accessPolicyWindow.num_bytes = std::min<size_t>(state_bytes, (size_t)persist_mb << 20);
                av.accessPolicyWindow.hitRatio = 1.0f;
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                persisting = cudaStreamSetAttribute(g_stream, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
                cudaGetLastError();
            }
            CK(cudaEventRecord(s.evk[0], g_stream));
            if (stencil_reduce) CK(ti->launch_stencil(la)); else CK(ti->launch(la));
            ++g_launches;
            if (heavy_n) {
                vb::LaunchArgs lh = la;
                lh.group = 256; lh.rows = heavy_rows; lh.n = heavy_n; lh.heavy_min = 0;
                CK(ti->launch(lh)); ++g_launches;
            }
            CK(cudaEventRecord(s.evk[1], g_stream));
            if (persisting) {   // hand the set-aside back: later kernels see the whole L2
                cudaStreamAttrValue av{};
                av.accessPolicyWindow.num_bytes = 0;
                cudaStreamSetAttribute(g_stream, cudaStreamAttributeAccessPolicyWindow, &av);
                cudaCtxResetPersistingL2Cache();
                cudaGetLastError();
            }
            }
        }
        if (call.size() > 1) {       // the event pair is reused by the next called type
            CK(cudaStreamSynchronize(g_stream));
            float mk = 0; cudaEventElapsedTime(&mk, s.evk[0], s.evk[1]); s.ms_kernel += mk;
        } else kernel_time_pending = true;      // read after the synchronisation of the error check below: one host round trip less per apply!
        for (auto p : tmp) dfree(p);            // (the pool hands buffers out again in stream order)
    }
    s.halo_wait_all();          // (a rank without agents of the called type launched nothing: join the halo stream before the buffers swap)
    CK(cudaEventRecord(s.ev[1], g_stream));
    g_trace.end("transition loop", tname);
    s.check_device_error("apply!");
    if (kernel_time_pending) { float mk = 0; if (cudaEventElapsedTime(&mk, s.evk[0], s.evk[1]) == cudaSuccess) s.ms_kernel += mk; cudaGetLastError(); }
    if (g_nranks > 1) {
        // transmit_remove_edges! (Simulation.jl:792-795) before transmit_edges! (:800).  Collective; every rank runs the same
        // transition, so all of them agree on the edge types whose removes have to travel.
        for (int w : write) if (w >= vb::EDGE_REF) {
            bool removes = false;
            for (auto* ti : tis) for (int i = 0; i < ti->n_edge_removes; ++i) removes |= ti->edge_removes[i] == w - vb::EDGE_REF;
            if (removes) s.transmit_removes(w - vb::EDGE_REF);
        }
        for (int w : write) if (w >= vb::EDGE_REF) s.transmit_edges(w - vb::EDGE_REF);
    }

    // ---- finish_write! agents (Simulation.jl:807), then edges (:809), then the dead-agent purge ----
    std::vector<uint32_t*> died_flags(s.agents.size() + 1, nullptr);
    std::vector<uint32_t> died_n(s.agents.size() + 1, 0);
    std::vector<uint32_t> died_cnt(s.agents.size() + 1, 0);
    g_trace.end("transmit_edges!", tname);
    for (int w : write) if (w < vb::EDGE_REF) { finish_write_agent(s, w, died_flags, died_n, died_cnt); g_trace.end("finish_write! agents", s.A(w).name); }
    for (int w : write) if (w >= vb::EDGE_REF) {
        EdgeStore& e = s.E(w - vb::EDGE_REF);
        s.build_container(w - vb::EDGE_REF, e.add_existing);
        e.last_change = s.num_transitions;
        e.writeable = false;
        g_trace.end("finish_write! edges", e.name);
    }
    bool any_dead = false;
    for (auto p : died_flags) any_dead |= p != nullptr;
    // multi-GPU (C7): the ids of the agents that died on the other ranks, so that edges *from* them are purged here too.
    // Collective whenever a mortal type is written (the same decision on every rank).
    uint64_t* remote_died = nullptr; uint32_t remote_died_n = 0;
    bool mortal_written = false;
    for (int w : write) if (w < vb::EDGE_REF && !s.A(w).immortal) mortal_written = true;
    if (g_nranks > 1 && mortal_written) {
        uint64_t mine = 0;
        for (size_t t = 1; t <= s.agents.size(); ++t) mine += died_cnt[t];
        std::vector<uint64_t> all;
        allgather8_host(&mine, all);
        uint64_t mx = 0;
        for (uint64_t v : all) mx = std::max(mx, v);
        if (mx) {
            uint64_t* sendb = dalloc<uint64_t>(mx);
            CK(cudaMemsetAsync(sendb, 0, mx * 8, g_stream));
            uint64_t o = 0;
            for (size_t t = 1; t <= s.agents.size(); ++t) {
                if (!died_cnt[t]) continue;
                AgentStore& a = s.agents[t - 1];   // this apply's deaths are the top died_cnt entries of the reuse stack
                slots_to_ids_kernel<<<nblk(died_cnt[t]), 256, 0, g_stream>>>(a.reuse + (a.n_reuse - died_cnt[t]), died_cnt[t], (uint32_t)t, s.rank, sendb + o); LAUNCH_CHECK();
                o += died_cnt[t];
            }
            remote_died_n = (uint32_t)(mx * g_nranks);
            remote_died = dalloc<uint64_t>(remote_died_n);
            NK(g_nccl.AllGather(sendb, remote_died, mx * 8, ncclUint8, g_comm, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            dfree(sendb);
            any_dead = true;
        }
    }
    if (any_dead) {
        for (size_t e = 0; e < s.edges.size(); ++e) s.materialize_stencil((int)e);
        const uint32_t tot = s.total_slots();
        uint8_t* dead = (uint8_t*)g_pool.alloc((size_t)tot + 1);
        CK(cudaMemsetAsync(dead, 0, (size_t)tot + 1, g_stream));
        for (size_t t = 1; t <= s.agents.size(); ++t)
            if (died_flags[t]) { mark_dead_kernel<<<nblk(died_n[t]), 256, 0, g_stream>>>(died_flags[t], died_n[t], dead, s.base[t]); LAUNCH_CHECK(); }
        if (remote_died) {
            MarkDeadArgs md{};
            md.ids = remote_died; md.n = remote_died_n; md.dead = dead; md.rank = s.rank; md.ntypes = (uint32_t)s.agents.size();
            std::memcpy(md.base, s.base, sizeof(md.base));
            for (size_t t = 1; t <= s.agents.size(); ++t) { md.lcap[t] = s.agents[t - 1].cap; md.nghost[t] = s.agents[t - 1].nghost; md.ghost_ids[t] = s.agents[t - 1].ghost_ids; }
            mark_remote_dead_kernel<<<nblk(remote_died_n), 256, 0, g_stream>>>(md); LAUNCH_CHECK();
        }
        s.purge_dead(dead);
        CK(cudaStreamSynchronize(g_stream));
        dfree(dead); dfree(remote_died);
        g_trace.end("purge dead agents' edges", tname);
        for (auto p : died_flags) dfree(p);
    }
    CK(cudaEventRecord(s.ev[2], g_stream));
    s.times_pending = true;   // the phase times are read when somebody asks (vb_last_apply_stats): apply! does not wait for its last kernels
    s.stats_pending = true;   // the 1024 spread counters are fetched lazily by vb_last_apply_stats
    const unsigned long long er = 0;
    s.st_edges_read = er; s.st_edges_appended = appended; s.st_launches = g_launches - launches0;
    s.num_transitions += 1;
    g_trace.flush_marks();
    if (g_nranks > 1) {     // births grow buffers, received edges grow ghost tables: the peer maps are re-checked by the next halo (every rank
                            // runs the same transition, so every rank decides the same)
        bool writes = false;
        for (auto* ti : tis) writes |= ti->n_agent_writes > 0 || ti->n_edge_writes > 0 || ti->n_edge_removes > 0;
        if (writes) s.mark_peer_check();
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------
extern "C" {

int vb_register_transition(const vb::TransitionInfo* info) {
    registry()[{info->name, info->agent_type}] = info;
    return 0;
}
int vb_register_map(const vb::MapInfo* info) {
    map_registry()[{info->name, info->type_name}] = info;
    return 0;
}

const char* vb_last_error(void) { return g_err.c_str(); }
const char* vb_backend(void) { return "cuda-sm100a"; }

int vb_init(int device) {
    return guard([&] {
        if (g_device >= 0) return;
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) throw CudaError("vahana_b200: no CUDA device available (this engine has no CPU fallback)");
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
        g_device = device;
    });
}
int vb_shutdown(void) {
    return guard([&] {
        if (g_device < 0) return;
        g_pool.release_all();
        cudaStreamDestroy(g_stream);
        g_stream = nullptr;
        g_device = -1;
    });
}
int vb_set_stream(void* stream) {   // run all engine work on the caller's stream (e.g. torch's current stream)
    return guard([&] {
        require_device();
        CK(cudaStreamSynchronize(g_stream));
        g_stream = (cudaStream_t)stream;
    });
}
int vb_comm_unique_id(uint8_t id_out[128]) {
    return guard([&] {
        if (!g_nccl.load()) throw CudaError("libnccl.so.2 could not be loaded (set VB_NCCL_LIB)");
        ncclUniqueId id;
        NK(g_nccl.GetUniqueId(&id));
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
        std::memcpy(id_out, &id, 128);
    });
}
int vb_comm_init(int rank, int nranks, const uint8_t* idbytes) {
    return guard([&] {
        require_device();
        if (nranks <= 1) { g_rank = 0; g_nranks = 1; return; }
        if (!g_nccl.load()) throw CudaError("libnccl.so.2 could not be loaded (set VB_NCCL_LIB)");
        if (nranks > (1 << vb::BITS_PROCESS)) throw ArgError("too many ranks for the 20-bit process id");
        ncclUniqueId id;
        std::memcpy(&id, idbytes, 128);
        NK(g_nccl.CommInitRank(&g_comm, nranks, id, rank));
        g_rank = rank; g_nranks = nranks;
        // NCCL connects its channels on the first collective (seconds on 8 GPUs): do that here, once, instead of inside the first
        // finish_init! (the scaling bench's build time grew with the rank count for that reason)
        std::vector<uint64_t> warm;
        const uint64_t me = (uint64_t)rank;
        allgather8_host(&me, warm);
        {   // ... and its point-to-point connections on the first ncclSend / ncclRecv between a pair of ranks: one grouped 8-byte exchange
            // with every peer (what the ghost-request exchange of finish_init! and the edge exchanges use)
            uint64_t* buf = dalloc<uint64_t>((size_t)2 * nranks);
            CK(cudaMemsetAsync(buf, 0, (size_t)2 * nranks * 8, g_stream));
            NK(g_nccl.GroupStart());
            for (int r = 0; r < nranks; ++r) {
                if (r == rank) continue;
                NK(g_nccl.Send(buf + r, 8, ncclUint8, r, g_comm, g_stream));
                NK(g_nccl.Recv(buf + nranks + r, 8, ncclUint8, r, g_comm, g_stream));
            }
            NK(g_nccl.GroupEnd());
            CK(cudaStreamSynchronize(g_stream));
            dfree(buf);
        }
    });
}
int vb_comm_rank(int* r, int* n) { *r = g_rank; *n = g_nranks; return VB_OK; }
// fold one 8-byte value per rank with `op` (mapreduce / num_agents / num_edges: MPI.Allreduce in the reference)
static void allgather8(const void* mine, std::vector<uint64_t>& all) { allgather8_host(mine, all); }
}  // extern "C"
namespace {
void allgather8_host(const void* mine, std::vector<uint64_t>& all) {
    all.assign((size_t)g_nranks, 0);
    if (g_nranks <= 1) { std::memcpy(all.data(), mine, 8); return; }
    uint64_t* d = dalloc<uint64_t>((size_t)g_nranks + 1);
    CK(cudaMemcpyAsync(d + g_nranks, mine, 8, cudaMemcpyHostToDevice, g_stream));
    NK(g_nccl.AllGather(d + g_nranks, d, 8, ncclUint8, g_comm, g_stream));
    CK(cudaMemcpyAsync(all.data(), d, (size_t)g_nranks * 8, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    dfree(d);
}
}  // namespace
extern "C" {
int vb_halo_bytes(vb_sim* s, uint64_t* out) { *out = s->halo_bytes; return VB_OK; }
int vb_last_halo_ms(vb_sim* s, double* ms) {     // device time of the last phased halo exchange (entry barrier excluded), negative = none
    *ms = -1.0;
    if (s->halo_timed && s->ev_halo[1] && cudaEventSynchronize(s->ev_halo[1]) == cudaSuccess) {
        float t = 0;
        if (cudaEventElapsedTime(&t, s->ev_halo[0], s->ev_halo[1]) == cudaSuccess) *ms = t;
    }
    cudaGetLastError();
    return VB_OK;
}
int vb_set_uniform_offset(vb_sim* s, int type, uint64_t offset) { return guard([&] { s->A(type).uoffset = offset; }); }

int vb_sim_create(const vb_model_desc* m, const void* params, vb_sim** out) {
    return guard([&] {
        require_device();
        if (m->n_agent_types > vb::MAX_AGENT_TYPES) throw ArgError("more agent types than this build supports (MAX_AGENT_TYPES)");
        if (m->n_edge_types > vb::MAX_EDGE_TYPES) throw ArgError("more edge types than this build supports (MAX_EDGE_TYPES)");
        if (m->param_size > vb::MAX_PARAM_BYTES) throw ArgError("parameter struct too large (MAX_PARAM_BYTES)");
        auto s = std::make_unique<vb_sim>();
        s->name = m->name;
        s->rank = (uint32_t)g_rank;
        for (uint32_t i = 0; i < m->n_agent_types; ++i) {
            AgentStore a;
            a.name = m->agent_types[i].name; a.size = m->agent_types[i].size; a.hints = m->agent_types[i].hints;
            a.word = vb::soa_word(a.size); a.ncols = a.size ? a.size / a.word : 0;
            a.immortal = a.hints & vb::AGENT_IMMORTAL; a.independent = a.hints & vb::AGENT_INDEPENDENT; a.stateless = a.size == 0;
            if (!a.immortal) s->all_immortal = false;
            s->agents.push_back(a);
        }
        for (uint32_t i = 0; i < m->n_edge_types; ++i) {
            EdgeStore e;
            const vb_edgetype_desc& d = m->edge_types[i];
            e.name = d.name; e.size = d.size; e.hints = d.hints; e.target = d.target_type; e.size_hint = d.size_hint;
            e.stateless = (d.hints & vb::EDGE_STATELESS) || d.size == 0; e.ignorefrom = d.hints & vb::EDGE_IGNORE_FROM;
            e.singleedge = d.hints & vb::EDGE_SINGLE_EDGE; e.singletype = d.hints & vb::EDGE_SINGLE_TYPE;
            if (e.singletype && (e.target < 1 || e.target > (int)m->n_agent_types)) throw AssertionError(":SingleType needs the target keyword");
            if (e.singletype && e.singleedge && !((d.hints & vb::EDGE_STATELESS) && e.ignorefrom))
                throw AssertionError(":SingleEdge and :SingleType can only be combined with :Stateless and :IgnoreFrom");
            const bool S = d.hints & vb::EDGE_STATELESS;
            e.stateless = S;
            e.kind = (S && e.ignorefrom) ? (e.singleedge ? vb::KIND_FLAG : vb::KIND_COUNT) : vb::KIND_CSR;
            e.word = vb::soa_word(S ? 0 : e.size); e.ncols = (!S && e.size) ? e.size / e.word : 0;
            s->edges.push_back(std::move(e));
        }
        if (m->param_size) s->params.assign((const uint8_t*)params, (const uint8_t*)params + m->param_size);
        s->d_error = dalloc<uint32_t>(1);
        s->d_scalars = dalloc<uint32_t>(64);
        s->d_stats = dalloc<unsigned long long>(4096);
        CK(cudaMemsetAsync(s->d_error, 0, 4, g_stream));
        CK(cudaMemsetAsync(s->d_stats, 0, 4096 * 8, g_stream));
        for (auto& e : s->ev) CK(cudaEventCreate(&e));
        for (auto& e : s->evk) CK(cudaEventCreate(&e));
        s->compute_bases(s->base);
        *out = s.release();
    });
}

int vb_sim_copy(const vb_sim* src, vb_sim** out) {   // copy_simulation: Simulation.jl:500-510
    return guard([&] {
        require_device();
        vb_sim& o = const_cast<vb_sim&>(*src);
        if (o.initialized) o.merge_all_pending();
        auto s = std::make_unique<vb_sim>();
        s->name = o.name; s->params = o.params; s->initialized = o.initialized; s->num_transitions = o.num_transitions;
        s->asserts_enabled = o.asserts_enabled; s->check_readable = o.check_readable; s->all_immortal = o.all_immortal; s->rank = o.rank;
        std::memcpy(s->base, o.base, sizeof(o.base));
        auto dup = [&](const void* p, size_t bytes) -> void* {
            if (!p) return nullptr;
            void* q = g_pool.alloc(bytes);
            CK(cudaMemcpyAsync(q, p, bytes, cudaMemcpyDeviceToDevice, g_stream));
            return q;
        };
        for (auto& a : o.agents) {
            AgentStore b = a;
            b.peers = AgentStore::PeerMap{};      // the copy maps its peers on its first halo exchange
            b.ev_phase.clear(); b.halo_pending = 0; b.halo_waited = 0;
            if (a.size && a.cap) { b.state[0] = (uint8_t*)dup(a.state[0], (size_t)a.stride() * a.size); b.state[1] = a.independent ? b.state[0] : (uint8_t*)dup(a.state[1], (size_t)a.stride() * a.size); }
            if (!a.immortal && a.cap) { b.died[0] = (uint8_t*)dup(a.died[0], a.stride()); b.died[1] = (uint8_t*)dup(a.died[1], a.stride()); b.reuse = (uint32_t*)dup(a.reuse, (size_t)a.reuse_cap * 4); }
            // the ghost table and the per-peer send lists of a multi-rank simulation are owned per simulation, whatever the type's hints
            b.ghost_ids = a.nghost ? (uint64_t*)dup(a.ghost_ids, (size_t)a.nghost * 8) : nullptr;
            b.send_slots = (a.send_slots && !a.send_off.empty() && a.send_off.back()) ? (uint32_t*)dup(a.send_slots, (size_t)a.send_off.back() * 4) : nullptr;
            b.send_buf = b.send_slots ? (uint8_t*)g_pool.alloc((size_t)a.send_off.back() * std::max<uint32_t>(a.size, 1)) : nullptr;   // pack buffer of the NCCL halo
            s->agents.push_back(b);
        }
        for (auto& e : o.edges) {
            EdgeStore f = e;
            f.chunks.clear();
            f.off = (uint32_t*)dup(e.off, ((size_t)e.rows + 2) * 4);
            f.src = (uint32_t*)dup(e.src, (size_t)std::max<uint32_t>(e.st_cap, e.nnz) * 4);
            f.st = (uint8_t*)dup(e.st, (size_t)e.st_cap * e.size);
            f.cnt = (uint32_t*)dup(e.cnt, ((size_t)e.rows + 1) * 4);
            f.log_to = f.log_from = f.wcnt = nullptr; f.log_st = nullptr; f.log_n = f.log_cap = 0; f.rows_w = 0;
            f.st_off = (int8_t*)dup(e.st_off, e.st_off_host.size());
            f.rlog_to = f.rlog_from = nullptr; f.rlog_st = nullptr; f.rlog_dst = nullptr; f.rlog_n = f.rlog_cap = 0;
            f.rm_row = f.rm_from = f.rm_mark = nullptr; f.rm_to64 = f.rm_from64 = nullptr; f.rm_n = f.rm_cap = 0; f.heavy_rows = nullptr; f.heavy_n = 0; f.heavy_version = ~0ull;
            f.blk = EdgeStore::Blocked{};
            for (auto& c : e.chunks) {
                RawChunk d; d.n = c.n;
                d.to = (uint64_t*)dup(c.to, c.n * 8); d.from = (uint64_t*)dup(c.from, c.n * 8); d.st = (uint8_t*)dup(c.st, c.n * e.size);
                f.chunks.push_back(d);
            }
            s->edges.push_back(std::move(f));
        }
        for (auto& r : o.rasters) { RasterStore q = r; q.cells = (uint32_t*)dup(r.cells, r.ids.size() * 4); s->rasters.push_back(q); }
        s->d_error = dalloc<uint32_t>(1);
        s->d_scalars = dalloc<uint32_t>(64);
        s->d_stats = dalloc<unsigned long long>(4096);
        CK(cudaMemsetAsync(s->d_error, 0, 4, g_stream));
        for (auto& e : s->ev) CK(cudaEventCreate(&e));
        for (auto& e : s->evk) CK(cudaEventCreate(&e));
        CK(cudaStreamSynchronize(g_stream));
        *out = s.release();
    });
}
int vb_sim_destroy(vb_sim* s) {
    delete s;
    // the buffers of a simulation that is gone rarely fit the next one: hand large caches back to the driver now instead of on the first
    // failed cudaMalloc in the middle of somebody's step (predator/prey after HK-100M in one process: steps of 50-450 ms among steps of 7 ms)
    if (g_device >= 0 && g_pool.cached_bytes() > ((size_t)256 << 20)) { cudaStreamSynchronize(g_stream); g_pool.release_all(); }
    return VB_OK;
}

int vb_set_param(vb_sim* s, const void* p, uint32_t size) {
    return guard([&] {
        if (s->initialized) throw AssertionError("set_param! can only be called before finish_init!");
        if (size > vb::MAX_PARAM_BYTES) throw ArgError("parameter struct too large");
        s->params.assign((const uint8_t*)p, (const uint8_t*)p + size);
    });
}
int vb_set_config(vb_sim* s, int asserts, int check_readable) { s->asserts_enabled = asserts; s->check_readable = check_readable; return VB_OK; }
int vb_disable_transition_checks(vb_sim* s, int disable) { s->intransition = disable; s->check_readable = !disable; return VB_OK; }

int vb_add_agents(vb_sim* s, int type, const void* states, uint64_t n, vb_agent_id* ids_out) {
    return guard([&] {   // add_agent!: AgentMethods.jl:65-89 (outside of transitions: init phase)
        require_device();
        AgentStore& a = s->A(type);
        s->mayassert(!s->initialized || s->intransition, "add_agent! only in the initialization phase or within a transition");
        s->mayassert(!s->initialized || a.writeable, "agent type must be in the `write` argument");
        if (n == 0) return;
        const uint64_t first = a.nextid;   // init phase: no reuse (nothing has died yet)
        s->ensure_agent_cap(type, first - 1 + n);
        if (a.size) {
            cudaPointerAttributes pa{};
            const bool dev = cudaPointerGetAttributes(&pa, states) == cudaSuccess && pa.type == cudaMemoryTypeDevice;
            cudaGetLastError();
            uint8_t* tmp = dev ? (uint8_t*)states : (uint8_t*)g_pool.alloc(n * a.size);
            if (!dev) CK(cudaMemcpyAsync(tmp, states, n * a.size, cudaMemcpyHostToDevice, g_stream));
            vbp::aos_to_soa_kernel<<<nblk(n * a.ncols), 256, 0, g_stream>>>(tmp, a.wstate(), a.stride(), first - 1, n, a.size, a.word); LAUNCH_CHECK();
            CK(cudaStreamSynchronize(g_stream));
            if (!dev) dfree(tmp);
        }
        a.nextid += n;
        for (uint64_t i = 0; i < n && ids_out; ++i) ids_out[i] = vb::agent_id((uint32_t)type, s->rank, first + i);   // ids_out may be NULL for bulk adds
    });
}

int vb_add_agent_per_process(vb_sim* s, int type, const void* state, vb_agent_id* id_out) {
    return guard([&] {   // add_agent_per_process!: Agent.jl:363-388 = prepare_write!(add_existing) + add_agent! + finish_write!
        require_device();
        if (!s->initialized) throw AssertionError("add_agent_per_process! can only be called after finish_init!");
        if (s->intransition) throw AssertionError("add_agent_per_process! cannot be called within a transition function");
        AgentStore& a = s->A(type);
        uint32_t slot;
        if (!a.immortal && a.n_reuse) {   // _get_next_id: pop the most recently freed slot
            CK(cudaMemcpyAsync(&slot, a.reuse + (a.n_reuse - 1), 4, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            a.n_reuse -= 1;
        } else {
            slot = (uint32_t)(a.nextid - 1);
            s->ensure_agent_cap(type, a.nextid);
            a.nextid += 1;
        }
        a.nslots = (uint32_t)std::max<uint64_t>(a.nslots, a.nextid - 1);
        if (a.size) {   // read := write afterwards, so the state goes to both buffers
            uint8_t* tmp = (uint8_t*)g_pool.alloc(a.size);
            CK(cudaMemcpyAsync(tmp, state, a.size, cudaMemcpyHostToDevice, g_stream));
            for (int b = 0; b < (a.independent ? 1 : 2); ++b) { vbp::aos_to_soa_kernel<<<1, 64, 0, g_stream>>>(tmp, a.state[b], a.stride(), slot, 1, a.size, a.word); LAUNCH_CHECK(); }
            CK(cudaStreamSynchronize(g_stream));
            dfree(tmp);
        }
        if (!a.immortal) { CK(cudaMemsetAsync(a.died[0] + slot, 0, 1, g_stream)); CK(cudaMemsetAsync(a.died[1] + slot, 0, 1, g_stream)); }
        a.halo_dirty = true;
        s->mark_peer_check();            // collective by contract (one agent on every rank): the buffers may have grown
        a.last_change = s->num_transitions;
        if (id_out) *id_out = vb::agent_id((uint32_t)type, s->rank, (uint64_t)slot + 1);
    });
}

int vb_add_edges(vb_sim* s, int ei, const vb_agent_id* from, const vb_agent_id* to, const void* states, uint64_t n) {
    return guard([&] {   // add_edge!: EdgeMethods.jl:388-523 (outside of transitions)
        require_device();
        EdgeStore& e = s->E(ei);
        s->mayassert(!s->initialized || s->intransition, "add_edge! only in the initialization phase or within a transition");
        s->mayassert(!s->check_readable || !s->initialized || e.writeable, "edge type must be in the `write` argument");
        if (e.implicit_stencil) s->materialize_stencil(ei);
        cudaPointerAttributes pa{};
        const bool dev = cudaPointerGetAttributes(&pa, to) == cudaSuccess && pa.type == cudaMemoryTypeDevice;
        cudaGetLastError();
        if (dev) {   // bulk ingest straight from device memory
            s->flush_raw(ei);
            RawChunk c; c.n = n;
            c.to = dalloc<uint64_t>(n); CK(cudaMemcpyAsync(c.to, to, n * 8, cudaMemcpyDeviceToDevice, g_stream));
            if (e.has_src()) { c.from = dalloc<uint64_t>(n); CK(cudaMemcpyAsync(c.from, from, n * 8, cudaMemcpyDeviceToDevice, g_stream)); }
            if (e.has_state()) { c.st = (uint8_t*)g_pool.alloc(n * e.size); CK(cudaMemcpyAsync(c.st, states, n * e.size, cudaMemcpyDeviceToDevice, g_stream)); }
            CK(cudaStreamSynchronize(g_stream));
            e.chunks.push_back(c); e.raw_n += n;
            return;
        }
        for (uint64_t i = 0; i < n; ++i) {
            const uint64_t t = to[i], f = from ? from[i] : 0;
            if (s->asserts_enabled) {   // ids must name existing agents (the reference dereferences them)
                const uint32_t tt = vb::type_nr(t);
                // the agent number can only be range-checked for agents of this rank: another rank's block may be longer than ours
                // (n % nranks != 0, explicit partitions); unknown remote ids are rejected by the ghost build / translation later
                const bool t_local = vb::process_nr(t) == s->rank, f_local = vb::process_nr(f) == s->rank;
                if (tt < 1 || tt > s->agents.size() || vb::agent_nr(t) < 1 || (t_local && vb::agent_nr(t) >= s->agents[tt - 1].nextid)) throw AssertionError("add_edge!: invalid target id");
                if (e.singletype && (int)tt != e.target) throw AssertionError("add_edge!: the :SingleType hint is set and the target has another type");
                if (e.has_src()) {
                    const uint32_t ft = vb::type_nr(f);
                    if (ft < 1 || ft > s->agents.size() || vb::agent_nr(f) < 1 || (f_local && vb::agent_nr(f) >= s->agents[ft - 1].nextid)) throw AssertionError("add_edge!: invalid source id");
                }
            }
            const uint8_t* st = (e.has_state() && states) ? (const uint8_t*)states + i * e.size : nullptr;
            if (e.singleedge && !e.singletype && !(e.stateless && e.ignorefrom) && s->asserts_enabled) {   // _can_add: :267-293
                auto it = e.single_seen.find(t);
                if (it != e.single_seen.end()) {
                    bool same = (!e.has_src() || it->second.first == f) && (!st || std::memcmp(it->second.second.data(), st, e.size) == 0);
                    if (!same) throw AssertionError("An edge has already been added to this agent (the edge type has the :SingleEdge hint)");
                } else {
                    e.single_seen[t] = {f, st ? std::vector<uint8_t>(st, st + e.size) : std::vector<uint8_t>()};
                }
            }
            e.h_to.push_back(t);
            if (e.has_src()) e.h_from.push_back(f);
            if (e.has_state()) { if (!st) throw ArgError("edge type has a state"); e.h_st.insert(e.h_st.end(), st, st + e.size); }
        }
        e.raw_n += n;
        if (e.h_to.size() >= (1u << 22)) s->flush_raw(ei);
    });
}
int vb_remove_edges(vb_sim* s, int ei, vb_agent_id from, vb_agent_id to) {
    return guard([&] {   // remove_edges! outside of transitions (init phase, or with the transition checks disabled): EdgeMethods.jl:527-599
        require_device();
        EdgeStore& e = s->E(ei);
        s->mayassert(!s->initialized || s->intransition, "remove_edges! only in the initialization phase or within a transition");
        if (from && e.ignorefrom) throw AssertionError("remove_edges! with a source agent is not defined for edgetypes with the :IgnoreFrom hint");
        s->materialize_stencil(ei);
        if (!s->initialized) {   // the adds are still staged on the host as AgentIDs: filter them in place
            if (!e.chunks.empty()) throw ArgError("remove_edges! before finish_init! after a device-side bulk add is not supported");
            size_t w = 0;
            for (size_t i = 0; i < e.h_to.size(); ++i) {
                const bool hit = e.h_to[i] == to && (!from || (e.has_src() && e.h_from[i] == from));
                if (hit) continue;
                e.h_to[w] = e.h_to[i];
                if (e.has_src()) e.h_from[w] = e.h_from[i];
                if (e.has_state()) std::memmove(&e.h_st[w * e.size], &e.h_st[i * e.size], e.size);
                ++w;
            }
            e.raw_n -= e.h_to.size() - w;
            e.h_to.resize(w); if (e.has_src()) e.h_from.resize(w); if (e.has_state()) e.h_st.resize(w * e.size);
            e.single_seen.erase(to);
            return;
        }
        s->merge_pending(ei);
        const uint32_t tt = vb::type_nr(to);
        const uint64_t tnr = vb::agent_nr(to);
        if (tt < 1 || tt > s->agents.size() || tnr < 1 || tnr > s->agents[tt - 1].cap) return;
        if (e.singletype && (int)tt != e.target) return;
        uint32_t rec[3] = {(e.singletype ? 0u : s->base[tt]) + (uint32_t)(tnr - 1), 0xffffffffu, 0u};
        if (from) {
            const uint32_t ft = vb::type_nr(from);
            const uint64_t fnr = vb::agent_nr(from);
            if (ft < 1 || ft > s->agents.size() || fnr < 1 || fnr > s->agents[ft - 1].cap) return;
            rec[1] = s->base[ft] + (uint32_t)(fnr - 1);
        }
        dfree(e.rm_row); dfree(e.rm_from); dfree(e.rm_mark); dfree(e.rm_to64); dfree(e.rm_from64); e.rm_to64 = e.rm_from64 = nullptr;
        e.rm_row = dalloc<uint32_t>(1); e.rm_from = dalloc<uint32_t>(1); e.rm_mark = dalloc<uint32_t>(1); e.rm_cap = 1; e.rm_n = 1;
        CK(cudaMemcpyAsync(e.rm_row, &rec[0], 4, cudaMemcpyHostToDevice, g_stream));
        CK(cudaMemcpyAsync(e.rm_from, &rec[1], 4, cudaMemcpyHostToDevice, g_stream));
        CK(cudaMemcpyAsync(e.rm_mark, &rec[2], 4, cudaMemcpyHostToDevice, g_stream));
        e.log_n = 0;
        if (e.kind != vb::KIND_CSR && !e.wcnt) {
            const uint32_t rows = s->rows_of(e);
            e.wcnt = dalloc<uint32_t>((size_t)rows + 1); e.rows_w = rows;
            CK(cudaMemsetAsync(e.wcnt, 0, ((size_t)rows + 1) * 4, g_stream));
            if (e.cnt) CK(cudaMemcpyAsync(e.wcnt, e.cnt, (size_t)std::min(rows, e.rows) * 4, cudaMemcpyDeviceToDevice, g_stream));
        }
        s->build_container(ei, true);
        CK(cudaStreamSynchronize(g_stream));
    });
}

int vb_add_raster(vb_sim* s, const char* name, int ndims, const int64_t* dims, int type, const void* states, vb_agent_id* ids_out) {
    return guard([&] {   // add_raster!: Raster.jl:32-54
        if (s->initialized) throw AssertionError("add_raster! can be only called before finish_init!");
        if (ndims < 1 || ndims > vb::MAX_RASTER_DIMS) throw ArgError("rasters with 1..4 dimensions are supported");
        if (s->rasters.size() >= vb::MAX_RASTERS) throw ArgError("too many rasters (MAX_RASTERS)");
        if (!name || !dims) throw ArgError("vb_add_raster: name and dims must be given");
        RasterStore r;
        r.name = name; r.dims.assign(dims, dims + ndims); r.type = type;
        uint64_t n = 1;
        for (int i = 0; i < ndims; ++i) { if (dims[i] < 1) throw ArgError("raster dimensions must be positive"); n *= (uint64_t)dims[i]; }
        r.ids.resize(n);
        int rc = vb_add_agents(s, type, states, n, r.ids.data());
        if (rc != VB_OK) throw AssertionError(g_err);
        if (ids_out) std::memcpy(ids_out, r.ids.data(), n * 8);
        s->rasters.push_back(std::move(r));
        uint32_t ob[vb::MAX_AGENT_TYPES + 2];
        std::memcpy(ob, s->base, sizeof(ob));
        s->rebase(ob);   // builds the device cell table
    });
}

int vb_set_raster(vb_sim* s, const char* name, int ndims, const int64_t* dims, int type, const vb_agent_id* ids) {
    return guard([&] {   // broadcastids (src/MPI.jl:59-73): the id grid of a raster whose cells were handed out to the ranks
        if (s->initialized) throw AssertionError("rasters can only be defined before finish_init!");
        if (ndims < 1 || ndims > vb::MAX_RASTER_DIMS) throw ArgError("rasters with 1..4 dimensions are supported");
        if (s->rasters.size() >= vb::MAX_RASTERS) throw ArgError("too many rasters (MAX_RASTERS)");
        if (!name || !dims || !ids) throw ArgError("vb_set_raster: name, dims and ids must be given");
        s->A(type);      // a registered agent type
        RasterStore r;
        r.name = name; r.dims.assign(dims, dims + ndims); r.type = type;
        uint64_t n = 1;
        for (int i = 0; i < ndims; ++i) { if (dims[i] < 1) throw ArgError("raster dimensions must be positive"); n *= (uint64_t)dims[i]; }
        r.ids.assign(ids, ids + n);
        for (uint64_t i = 0; i < n; ++i) {
            if ((int)vb::type_nr(ids[i]) != type) throw AssertionError("vb_set_raster: a cell id of another agent type");
            if (vb::process_nr(ids[i]) == s->rank && vb::agent_nr(ids[i]) >= s->A(type).nextid) throw AssertionError("vb_set_raster: a cell id that names no agent of this rank");
        }
        r.distributed = g_nranks > 1;
        s->rasters.push_back(std::move(r));
        uint32_t ob[vb::MAX_AGENT_TYPES + 2];
        std::memcpy(ob, s->base, sizeof(ob));
        s->rebase(ob);   // builds the device cell table (cells of other ranks are marked) and keeps the id table on the device
    });
}

namespace {
// Raster read-outs of a distributed raster: every rank has filled in the cells it owns and zeros elsewhere; the byte-wise sum over the
// ranks is the joined array (exactly one rank contributes to a byte, so nothing carries).  join(), src/MPI.jl:492-517.  Collective.
void raster_join(const RasterStore& r, void* dev, size_t bytes) {
    if (!r.distributed || g_nranks <= 1 || !bytes) return;
    NK(g_nccl.AllReduce(dev, dev, bytes, ncclUint8, 0 /* ncclSum */, g_comm, g_stream));
}
// host-side stencil enumeration for the init-phase helpers (Raster.jl:82-110)
std::vector<std::vector<int64_t>> stencil(int metric, int n, double distance) {
    int64_t d = (int64_t)std::floor(distance);
    std::vector<std::vector<int64_t>> out;
    if (d < 0) return out;
    std::vector<int64_t> cur(n, -d);
    while (true) {
        bool zero = true; double n2 = 0; int64_t n1 = 0;
        for (int i = 0; i < n; ++i) { zero &= cur[i] == 0; n2 += (double)(cur[i] * cur[i]); n1 += std::llabs(cur[i]); }
        bool keep = !zero;
        if (keep && metric == vb::EUCLIDEAN) keep = std::sqrt(n2) <= distance;
        if (keep && metric == vb::MANHATTEN) keep = (double)n1 <= distance;
        if (keep) out.push_back(cur);
        int i = 0;
        while (i < n && ++cur[i] > d) { cur[i] = -d; ++i; }
        if (i == n) break;
    }
    return out;
}
bool checkpos(std::vector<int64_t>& pos, const std::vector<int64_t>& dims, bool periodic) {   // Raster.jl:479-499
    bool oob = false;
    for (size_t i = 0; i < dims.size(); ++i)
        if (pos[i] < 1 || pos[i] > dims[i]) { oob = true; int64_t m = (pos[i] - 1) % dims[i]; if (m < 0) m += dims[i]; pos[i] = m + 1; }
    return !oob || periodic;
}
size_t linear_index(const std::vector<int64_t>& pos, const std::vector<int64_t>& dims) {
    size_t idx = 0, stride = 1;
    for (size_t i = 0; i < dims.size(); ++i) { idx += (size_t)(pos[i] - 1) * stride; stride *= (size_t)dims[i]; }
    return idx;
}
RasterStore& find_raster(vb_sim* s, const char* name) {
    for (auto& r : s->rasters) if (r.name == name) return r;
    throw ArgError(std::string("unknown raster ") + name);
}
}  // namespace

}  // extern "C" (re-opened below)
// explicit generation of the raster edges (Raster.jl:139-167): for org in cells (column-major), for s in stencil: org -> shifted
void vb_sim::emit_raster_edges(int ei, int raster, double distance, int metric, bool periodic, const void* st) {
    RasterStore& r = rasters[raster];
    EdgeStore& e = E(ei);
    auto sten = stencil(metric, (int)r.dims.size(), distance);
    const size_t n = r.ids.size();
    std::vector<uint64_t> from, to;
    from.reserve(std::min<size_t>(n * sten.size(), (size_t)1 << 24)); to.reserve(from.capacity());
    std::vector<uint8_t> sts;
    std::vector<int64_t> org(r.dims.size(), 1), sh(r.dims.size());
    auto flush = [&] {
        if (to.empty()) return;
        if (e.has_state()) { sts.resize(to.size() * e.size); for (size_t i = 0; i < to.size(); ++i) std::memcpy(&sts[i * e.size], st, e.size); }
        int rc = vb_add_edges(this, ei, from.data(), to.data(), e.has_state() ? sts.data() : nullptr, to.size());
        if (rc != VB_OK) throw AssertionError(g_err);
        from.clear(); to.clear();
    };
    for (size_t i = 0; i < n; ++i) {
        for (auto& o : sten) {
            for (size_t k = 0; k < org.size(); ++k) sh[k] = org[k] + o[k];
            if (checkpos(sh, r.dims, periodic)) { from.push_back(r.ids[i]); to.push_back(r.ids[linear_index(sh, r.dims)]); }
        }
        if (to.size() >= ((size_t)1 << 24)) flush();
        size_t k = 0;
        while (k < org.size() && ++org[k] > r.dims[k]) { org[k] = 1; ++k; }
    }
    flush();
}
extern "C" {
int vb_connect_raster_neighbors(vb_sim* s, const char* name, int ei, double distance, int metric, int periodic, const void* st) {
    return guard([&] {
        RasterStore& r = find_raster(s, name);
        EdgeStore& e = s->E(ei);
        const int ri = (int)(&r - &s->rasters[0]);
        auto sten = stencil(metric, (int)r.dims.size(), distance);
        // Grid-stencil fast path (K7): if these are the only edges of a stateless CSR edge type over consecutively stored cells,
        // keep them implicit — no 8N-edge list, no sort; the read phase enumerates neighbours arithmetically.
        bool contiguous = !r.ids.empty();
        for (size_t i = 1; i < r.ids.size() && contiguous; ++i) contiguous = r.ids[i] == r.ids[0] + i;
        const bool eligible = !getenv("VB_NO_IMPLICIT_STENCIL") && !s->initialized && e.kind == vb::KIND_CSR && !e.has_state() && !e.singleedge &&
                              !e.ignorefrom && e.raw_n == 0 && !e.implicit_stencil && contiguous && !sten.empty() &&
                              sten.size() <= vb::MAX_IMPLICIT_STENCIL && (!e.singletype || e.target == r.type) && r.ids.size() < 0x7fffffffull &&
                              std::count_if(s->edges.begin(), s->edges.end(), [](const EdgeStore& x) { return x.implicit_stencil; }) < vb::MAX_RASTERS;
        if (!eligible) { s->materialize_stencil(ei); s->emit_raster_edges(ei, ri, distance, metric, periodic != 0, st); return; }
        e.implicit_stencil = true; e.st_raster = ri; e.st_metric = metric; e.st_distance = distance; e.st_periodic = periodic != 0;
        e.st_n = (int)sten.size(); e.st_reach = 0; e.st_slot0 = (uint32_t)(vb::agent_nr(r.ids[0]) - 1);
        e.st_off_host.assign(sten.size() * vb::MAX_RASTER_DIMS, 0);
        for (size_t i = 0; i < sten.size(); ++i)
            for (size_t k = 0; k < sten[i].size(); ++k) {
                e.st_off_host[i * vb::MAX_RASTER_DIMS + k] = (int8_t)sten[i][k];
                e.st_reach = std::max<int>(e.st_reach, (int)std::llabs(sten[i][k]));
            }
        dfree(e.st_off);
        e.st_off = (int8_t*)g_pool.alloc(e.st_off_host.size());
        CK(cudaMemcpyAsync(e.st_off, e.st_off_host.data(), e.st_off_host.size(), cudaMemcpyHostToDevice, g_stream));
        CK(cudaStreamSynchronize(g_stream));
    });
}
int vb_move_to(vb_sim* s, const char* name, vb_agent_id id, const int64_t* posv, int e_from, const void* s_from, int e_to, const void* s_to,
               double distance, int metric, int periodic, int only_surrounding) {
    return guard([&] {   // Raster.jl:437-477
        RasterStore& r = find_raster(s, name);
        std::vector<int64_t> pos(posv, posv + r.dims.size());
        auto add = [&](int ei, uint64_t f, uint64_t t, const void* st) {
            int rc = vb_add_edges(s, ei, &f, &t, st, 1);
            if (rc != VB_OK) throw AssertionError(g_err);
        };
        if (!only_surrounding) {
            for (size_t k = 0; k < pos.size(); ++k) if (pos[k] < 1 || pos[k] > r.dims[k]) throw AssertionError("move_to!: position outside the raster");
            const uint64_t cell = r.ids[linear_index(pos, r.dims)];
            if (e_from >= 0) add(e_from, cell, id, s_from);
            if (e_to >= 0) add(e_to, id, cell, s_to);
        }
        if (distance >= 1) {
            std::vector<int64_t> sh(pos.size());
            for (auto& o : stencil(metric, (int)r.dims.size(), distance)) {
                for (size_t k = 0; k < pos.size(); ++k) sh[k] = pos[k] + o[k];
                if (!checkpos(sh, r.dims, periodic)) continue;
                const uint64_t cell = r.ids[linear_index(sh, r.dims)];
                if (e_from >= 0) add(e_from, cell, id, s_from);
                if (e_to >= 0) add(e_to, id, cell, s_to);
            }
        }
    });
}
int vb_cellid(vb_sim* s, const char* name, const int64_t* posv, vb_agent_id* out) {
    return guard([&] {
        RasterStore& r = find_raster(s, name);
        std::vector<int64_t> pos(posv, posv + r.dims.size());
        for (size_t k = 0; k < pos.size(); ++k) if (pos[k] < 1 || pos[k] > r.dims[k]) throw AssertionError("cellid: position outside the raster");
        *out = r.ids[linear_index(pos, r.dims)];
    });
}

int vb_finish_init(vb_sim* s) {
    return guard([&] {   // Simulation.jl:403-476 (single rank): finish_write! all agent types, then all edge types
        require_device();
        if (s->initialized) throw AssertionError("You can not call finish_init! twice for the same simulation");
        for (auto& a : s->agents) {   // read := write
            a.nslots = (uint32_t)(a.nextid - 1);
            a.cur ^= 1;
            a.write_stale = true;
            if (!a.immortal && a.cap) CK(cudaMemsetAsync(a.rdied(), 0, a.stride(), g_stream));
        }
        s->initialized = true;
        s->mark_peer_check();
        s->build_ghosts();
        s->merge_all_pending();
        for (auto& e : s->edges) {   // every container exists after init, even if empty
            if (e.kind == vb::KIND_CSR && !e.off && !e.implicit_stencil) { e.log_n = 0; s->build_container((int)(&e - &s->edges[0]), false); }
            if (e.kind != vb::KIND_CSR && !e.cnt) s->build_container((int)(&e - &s->edges[0]), false);
            e.single_seen.clear();
        }
        CK(cudaStreamSynchronize(g_stream));
        s->num_transitions = 1;
    });
}

int vb_apply(vb_sim* s, const char* transition, const int* call, int ncall, const int* read, int nread, const int* write, int nwrite,
             const int* add_existing, int nadd, int with_edge, uint64_t seed) {
    return guard([&] {
        do_apply(*s, transition, std::vector<int>(call, call + ncall), std::vector<int>(read, read + nread), std::vector<int>(write, write + nwrite),
                 std::vector<int>(add_existing, add_existing + nadd), with_edge, seed);
    });
}
int vb_has_transition(const char* t, const char* a) { return registry().count({t, a}) ? 1 : 0; }
int vb_load_model_library(const char* path) {
    void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) { g_err = std::string("dlopen: ") + dlerror(); return VB_ERR_NOTFOUND; }
    return VB_OK;
}

int vb_num_agents(vb_sim* s, int type, uint64_t* n_out) {
    return guard([&] {   // Agent.jl:324-343
        require_device();
        AgentStore& a = s->A(type);
        auto sum_ranks = [&](uint64_t local) {   // MPI.Allreduce(local_num, +): Agent.jl:338-342
            std::vector<uint64_t> all;
            allgather8(&local, all);
            uint64_t t = 0;
            for (uint64_t v : all) t += v;
            return t;
        };
        if (a.immortal) { *n_out = sum_ranks(a.nextid - 1); return; }
        const uint32_t n = s->initialized ? a.nslots : (uint32_t)(a.nextid - 1);
        if (!n) { *n_out = sum_ranks(0); return; }
        if (!s->initialized) { *n_out = sum_ranks(n); return; }
        MapArgs ma{};
        ma.cols = nullptr; ma.n = n; ma.died = a.rdied(); ma.dt = -1; ma.op = vb::OP_SUM; ma.word = 1;
        long long* part = dalloc<long long>(1024 + 1);
        const unsigned nb = std::min<unsigned>(1024, nblk(n));
        mapreduce_kernel<false><<<nb, 256, 0, g_stream>>>(ma, part); LAUNCH_CHECK();
        mapreduce_final_kernel<false><<<1, 256, 0, g_stream>>>(part, nb, vb::OP_SUM, part + 1024); LAUNCH_CHECK();
        long long r = 0;
        CK(cudaMemcpyAsync(&r, part + 1024, 8, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        dfree(part);
        *n_out = sum_ranks((uint64_t)r);
    });
}

int vb_all_agents(vb_sim* s, int type, void* states_out, vb_agent_id* ids_out, uint64_t cap, uint64_t* n_out) {
    return guard([&] {   // Agent.jl:234-313: live slots in ascending nr
        require_device();
        AgentStore& a = s->A(type);
        const uint32_t n = s->initialized ? a.nslots : (uint32_t)(a.nextid - 1);
        const uint8_t* state = s->initialized ? a.rstate() : a.wstate();
        const uint8_t* died = (a.immortal || !s->initialized) ? nullptr : a.rdied();
        if (!n) { *n_out = 0; return; }
        uint32_t* flag = dalloc<uint32_t>(n); uint32_t* pos = dalloc<uint32_t>(n); uint32_t* idx = dalloc<uint32_t>(n);
        uint32_t* scr = dalloc<uint32_t>(vbp::scan_scratch_words(n));
        vbp::alive_flags_kernel<<<nblk(n), 256, 0, g_stream>>>(died, n, flag); LAUNCH_CHECK();
        vbp::exclusive_scan(flag, pos, n, s->d_scalars, scr, g_stream);
        uint32_t live = 0;
        CK(cudaMemcpyAsync(&live, s->d_scalars, 4, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        *n_out = live;
        if (live && cap >= live) {
            vbp::compact_indices_kernel<<<nblk(n), 256, 0, g_stream>>>(flag, pos, n, idx); LAUNCH_CHECK();
            if (states_out && a.size) {
                uint8_t* tmp = (uint8_t*)g_pool.alloc((size_t)live * a.size);
                vbp::soa_gather_aos_kernel<<<nblk((uint64_t)live * a.ncols), 256, 0, g_stream>>>(state, tmp, a.stride(), idx, live, a.size, a.word); LAUNCH_CHECK();
                CK(cudaMemcpyAsync(states_out, tmp, (size_t)live * a.size, cudaMemcpyDeviceToHost, g_stream));
                CK(cudaStreamSynchronize(g_stream));
                dfree(tmp);
            }
            if (ids_out) {
                std::vector<uint32_t> h(live);
                CK(cudaMemcpyAsync(h.data(), idx, (size_t)live * 4, cudaMemcpyDeviceToHost, g_stream));
                CK(cudaStreamSynchronize(g_stream));
                for (uint32_t i = 0; i < live; ++i) ids_out[i] = vb::agent_id((uint32_t)type, s->rank, (uint64_t)h[i] + 1);
            }
        }
        dfree(flag); dfree(pos); dfree(idx); dfree(scr);
    });
}

int vb_agentstate(vb_sim* s, vb_agent_id id, int type, void* out) {
    return guard([&] {   // AgentMethods.jl:91-124
        require_device();
        AgentStore& a = s->A(type);
        s->mayassert((int)vb::type_nr(id) == type, "The id of the agent does not match the given type");
        s->mayassert(a.prepared || !s->check_readable, "agent type must be in the `read` argument of the transition function");
        uint64_t nr = vb::agent_nr(id);
        bool ghost = false;
        if (g_nranks > 1 && vb::process_nr(id) != s->rank) {
            // an agent of another rank (the reference answers from `foreignstate`, AgentMethods.jl:103-111): the ghost slot that mirrors it,
            // as of the last halo exchange of an apply! that read the type
            int64_t* found = (int64_t*)(s->d_scalars + 56);
            find_ghost_kernel<<<1, 1, 0, g_stream>>>(a.ghost_ids, a.nghost, id, found); LAUNCH_CHECK();
            int64_t k = -1;
            CK(cudaMemcpyAsync(&k, found, 8, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            if (k < 0) throw AssertionError("agentstate: the agent lives on another rank and is not mirrored here (no local edge refers to it)");
            nr = (uint64_t)a.cap + (uint64_t)k + 1;
            ghost = true;
        }
        if (nr < 1 || (!ghost && nr > a.nslots)) throw AssertionError("agentstate: agent does not exist");
        if (!a.immortal && !ghost) {
            uint8_t d = 0;
            CK(cudaMemcpyAsync(&d, a.rdied() + (nr - 1), 1, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            s->mayassert(!d, "agentstate was requested for an agent that has been removed");
        }
        if (!a.size) return;
        uint8_t* tmp = (uint8_t*)g_pool.alloc(a.size);
        vbp::soa_to_aos_kernel<<<1, 64, 0, g_stream>>>(a.rstate(), tmp, a.stride(), nr - 1, 1, a.size, a.word); LAUNCH_CHECK();
        CK(cudaMemcpyAsync(out, tmp, a.size, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        dfree(tmp);
    });
}

int vb_num_edges_total(vb_sim* s, int e, int write, uint64_t* n_out) {
    return guard([&] {   // Edge.jl:373-389 incl. the Allreduce over ranks
        require_device();
        const uint64_t local = s->edge_total(e, write != 0);
        std::vector<uint64_t> all;
        allgather8(&local, all);
        uint64_t t = 0;
        for (uint64_t v : all) t += v;
        *n_out = t;
    });
}

namespace {
// fetch row `to` of edge type e to the host: ids (AgentIDs) and AoS states; returns count or -1 (`nothing`)
int64_t fetch_row(vb_sim* s, EdgeStore& e, vb_agent_id to, std::vector<uint64_t>* from, std::vector<uint8_t>* st, bool count_only) {
    const uint32_t tt = vb::type_nr(to);
    const uint64_t nr = vb::agent_nr(to);
    if (tt < 1 || tt > s->agents.size() || nr < 1) throw AssertionError("invalid agent id");
    if (e.singletype) s->mayassert((int)tt == e.target, "The :SingleType hint is set and the agent has another type");
    if (e.implicit_stencil) {
        const RasterStore& r = s->rasters[e.st_raster];
        if ((int)tt != r.type || nr - 1 < e.st_slot0 || nr - 1 - e.st_slot0 >= r.ids.size()) return -1;
        auto row = s->stencil_row_host(e, nr - 1 - e.st_slot0);
        if (row.empty()) return -1;
        if (from && !count_only) { from->resize(row.size()); for (size_t i = 0; i < row.size(); ++i) (*from)[i] = r.ids[row[i]]; }
        return (int64_t)row.size();
    }
    if (e.singletype && (int)tt != e.target) return -1;
    if (nr > s->agents[tt - 1].cap) return -1;
    const uint32_t row = e.singletype ? (uint32_t)(nr - 1) : s->base[tt] + (uint32_t)(nr - 1);
    if (row >= e.rows) return -1;
    if (e.kind != vb::KIND_CSR) {
        uint32_t c = 0;
        CK(cudaMemcpyAsync(&c, e.cnt + row, 4, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        return c ? (int64_t)c : -1;
    }
    uint32_t o[2];
    CK(cudaMemcpyAsync(o, e.off + row, 8, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    const uint32_t n = o[1] - o[0];
    if (!n) return -1;
    if (count_only) return n;
    if (from && e.has_src()) {
        std::vector<uint32_t> c(n);
        CK(cudaMemcpyAsync(c.data(), e.src + o[0], (size_t)n * 4, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        from->resize(n);
        for (uint32_t i = 0; i < n; ++i) {
            uint32_t t = 1;
            while (t < s->agents.size() && c[i] >= s->base[t + 1]) ++t;
            const AgentStore& ga = s->agents[t - 1];
            const uint32_t slot = c[i] - s->base[t];
            if (slot >= ga.cap && slot - ga.cap < ga.nghost) {        // a source on another rank: the id its ghost slot mirrors
                CK(cudaMemcpyAsync(&(*from)[i], ga.ghost_ids + (slot - ga.cap), 8, cudaMemcpyDeviceToHost, g_stream));
                CK(cudaStreamSynchronize(g_stream));
            } else (*from)[i] = vb::agent_id(t, s->rank, (uint64_t)slot + 1);
        }
    }
    if (st && e.has_state()) {
        uint8_t* tmp = (uint8_t*)g_pool.alloc((size_t)n * e.size);
        vbp::soa_to_aos_kernel<<<nblk((uint64_t)n * e.ncols), 256, 0, g_stream>>>(e.st, tmp, e.st_cap, o[0], n, e.size, e.word); LAUNCH_CHECK();
        st->resize((size_t)n * e.size);
        CK(cudaMemcpyAsync(st->data(), tmp, (size_t)n * e.size, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        dfree(tmp);
    }
    return n;
}
void avail(bool ok, const char* what) { if (!ok) throw AssertionError(std::string(what) + " is not defined for this hint combination"); }
}  // namespace

int vb_edges_of(vb_sim* s, int ei, vb_agent_id to, int what, vb_agent_id* from_out, void* states_out, uint64_t cap, int64_t* n_out) {
    return guard([&] {   // EdgeMethods.jl:699-892
        require_device();
        EdgeStore& e = s->E(ei);
        const bool S = e.stateless, I = e.ignorefrom, E1 = e.singleedge, T = e.singletype;
        switch (what) {
            case VB_ACC_EDGES: avail(!S && !I, "edges"); break;
            case VB_ACC_NEIGHBORIDS: avail(!I, "neighborids"); break;
            case VB_ACC_NEIGHBORIDS_ITER: avail(!I && !E1, "neighborids_iter"); break;
            case VB_ACC_EDGESTATES: avail(!S, "edgestates"); break;
            case VB_ACC_EDGESTATES_ITER: avail(!S && !E1, "edgestates_iter"); break;
            case VB_ACC_NUM_EDGES: avail(!E1, "num_edges"); break;
            case VB_ACC_HAS_EDGE: avail(!(E1 && T && !(S && I)), "has_edge"); break;
            default: throw ArgError("bad accessor");
        }
        s->mayassert(e.readable || !s->check_readable, "edge type is not in the `read` argument of apply!");
        s->merge_pending(ei);
        const bool count_only = what == VB_ACC_NUM_EDGES || what == VB_ACC_HAS_EDGE || (S && I);
        std::vector<uint64_t> from; std::vector<uint8_t> st;
        const int64_t n = fetch_row(s, e, to, &from, &st, count_only || cap == 0);
        if (what == VB_ACC_NUM_EDGES || what == VB_ACC_HAS_EDGE) { *n_out = n < 0 ? 0 : n; return; }
        *n_out = n;
        if (n <= 0 || count_only || cap == 0) return;
        for (int64_t i = 0; i < n && (uint64_t)i < cap; ++i) {
            if (from_out && !I) from_out[i] = from[i];
            if (states_out && e.has_state()) std::memcpy((uint8_t*)states_out + i * e.size, &st[i * e.size], e.size);
        }
    });
}

int vb_export_csr(vb_sim* s, int ei, int target_type, uint64_t* offsets, uint64_t nrows, vb_agent_id* from_out, void* states_out, uint64_t cap) {
    return guard([&] {
        require_device();
        EdgeStore& e = s->E(ei);
        s->merge_pending(ei);
        const uint32_t rb = e.singletype ? 0 : s->base[target_type];
        if (e.singletype && target_type != e.target) { for (uint64_t r = 0; r <= nrows; ++r) offsets[r] = 0; return; }
        if (e.implicit_stencil) {
            const RasterStore& r = s->rasters[e.st_raster];
            uint64_t n = 0;
            for (uint64_t row = 0; row < nrows; ++row) {
                offsets[row] = n;
                if (target_type != r.type || row < e.st_slot0 || row - e.st_slot0 >= r.ids.size()) continue;
                auto src = s->stencil_row_host(e, row - e.st_slot0);
                for (uint32_t c : src) { if (from_out && n < cap) from_out[n] = r.ids[c]; ++n; }
            }
            offsets[nrows] = n;
            return;
        }
        std::vector<uint32_t> off(nrows + 1, 0);
        if (e.kind != vb::KIND_CSR) {
            std::vector<uint32_t> c(nrows, 0);
            const uint64_t avail_rows = rb < e.rows ? std::min<uint64_t>(std::min<uint64_t>(nrows, s->A(target_type).cap), e.rows - rb) : 0;
            if (avail_rows && e.cnt) { CK(cudaMemcpyAsync(c.data(), e.cnt + rb, avail_rows * 4, cudaMemcpyDeviceToHost, g_stream)); CK(cudaStreamSynchronize(g_stream)); }
            uint64_t run = 0;
            for (uint64_t r = 0; r < nrows; ++r) { offsets[r] = run; run += c[r]; }
            offsets[nrows] = run;
            return;
        }
        const uint64_t type_rows = s->A(target_type).cap;   // rows of this target type only (the composite row space continues with the next type)
        const uint64_t avail_rows = (e.off && rb < e.rows) ? std::min<uint64_t>(std::min<uint64_t>(nrows, type_rows), e.rows - rb) : 0;
        if (avail_rows) { CK(cudaMemcpyAsync(off.data(), e.off + rb, (avail_rows + 1) * 4, cudaMemcpyDeviceToHost, g_stream)); CK(cudaStreamSynchronize(g_stream)); }
        const uint32_t o0 = avail_rows ? off[0] : 0;
        for (uint64_t r = 0; r <= nrows; ++r) offsets[r] = r <= avail_rows ? off[r] - o0 : off[avail_rows] - o0;
        const uint64_t n = avail_rows ? off[avail_rows] - o0 : 0;
        if (!n || cap < n) return;
        if (from_out && e.has_src()) {
            uint64_t* ids = dalloc<uint64_t>(n);
            RebaseArgs ra{}; std::memcpy(ra.old_base, s->base, sizeof(ra.old_base)); std::memcpy(ra.new_base, s->base, sizeof(ra.new_base)); ra.ntypes = (uint32_t)s->agents.size();
            for (size_t t = 1; t <= s->agents.size(); ++t) {     // sources on other ranks are ghost slots: their ids come from the ghost table
                const AgentStore& ga = s->agents[t - 1];
                ra.old_lcap[t] = ga.cap; ra.remap[t] = ga.nghost ? reinterpret_cast<const uint32_t*>(ga.ghost_ids) : nullptr;
            }
            comp_to_id_kernel<<<nblk(n), 256, 0, g_stream>>>(e.src + o0, n, ids, ra, s->rank); LAUNCH_CHECK();
            CK(cudaMemcpyAsync(from_out, ids, n * 8, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            dfree(ids);
        }
        if (states_out && e.has_state()) {
            uint8_t* tmp = (uint8_t*)g_pool.alloc(n * e.size);
            vbp::soa_to_aos_kernel<<<nblk(n * e.ncols), 256, 0, g_stream>>>(e.st, tmp, e.st_cap, o0, n, e.size, e.word); LAUNCH_CHECK();
            CK(cudaMemcpyAsync(states_out, tmp, n * e.size, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            dfree(tmp);
        }
    });
}

int vb_all_edges(vb_sim* s, int ei, vb_agent_id* to_out, vb_agent_id* from_out, void* states_out, uint64_t cap, uint64_t* n_out) {
    return guard([&] {   // EdgeMethods.jl:1005-1027; emitted in ascending target id order
        require_device();
        EdgeStore& e = s->E(ei);
        if (!s->initialized) { *n_out = 0; return; }   // the reference returns [] while the read container is empty (num_edges(sim, T) == 0, EdgeMethods.jl:1006-1008)
        s->merge_pending(ei);
        uint64_t n = 0;
        for (size_t t = 1; t <= s->agents.size(); ++t) {
            if (e.singletype && (int)t != e.target) continue;
            const uint64_t nrows = s->agents[t - 1].cap;
            if (!nrows) continue;
            std::vector<uint64_t> off(nrows + 1);
            int rc = vb_export_csr(s, ei, (int)t, off.data(), nrows, nullptr, nullptr, 0);
            if (rc != VB_OK) throw AssertionError(g_err);
            const uint64_t m = off[nrows];
            if (m && n + m <= cap) {
                rc = vb_export_csr(s, ei, (int)t, off.data(), nrows, from_out ? from_out + n : nullptr, states_out ? (uint8_t*)states_out + n * e.size : nullptr, m);
                if (rc != VB_OK) throw AssertionError(g_err);
                if (to_out) for (uint64_t r = 0; r < nrows; ++r) for (uint64_t k = off[r]; k < off[r + 1]; ++k) to_out[n + k] = vb::agent_id((uint32_t)t, s->rank, r + 1);
            }
            n += m;
        }
        *n_out = n;
    });
}

// mapreduce with the map given as a field selector (mi == nullptr) or as a registered functor (mi)
static int mapreduce_common(vb_sim* s, int type_ref, int offset, int dt, int has_cmp, int64_t cmp, int op, int result_dt, const void* init, void* out,
                            const vb::MapInfo* mi) {
    return guard([&] {
        require_device();
        if (s->intransition) throw AssertionError("You can not call mapreduce inside of a transition function.");
        MapArgs ma{};
        ma.offset = offset; ma.dt = dt; ma.has_cmp = has_cmp; ma.cmp = cmp; ma.op = op;
        if (type_ref < vb::EDGE_REF) {
            AgentStore& a = s->A(type_ref);
            ma.cols = a.rstate(); ma.stride = a.stride(); ma.word = a.word ? a.word : 1; ma.n = a.nextid - 1; ma.died = a.immortal ? nullptr : a.rdied();
            if (ma.n > a.nslots) ma.n = a.nslots;
        } else {
            EdgeStore& e = s->E(type_ref - vb::EDGE_REF);
            if (e.stateless) throw AssertionError("mapreduce is not defined for :Stateless edge types");
            s->merge_pending(type_ref - vb::EDGE_REF);
            ma.cols = e.st; ma.stride = e.st_cap; ma.word = e.word; ma.n = e.nnz; ma.died = nullptr;
        }
        const bool isf = result_dt == vb::DT_F64 || result_dt == vb::DT_F32;
        vb::MapLaunchArgs mla{};
        if (mi) {
            const uint32_t esize = type_ref < vb::EDGE_REF ? s->A(type_ref).size : s->E(type_ref - vb::EDGE_REF).size;
            if (mi->elem_size != esize) throw ArgError(std::string("map '") + mi->name + "': sizeof(Elem) does not match the registered size of " + mi->type_name);
            if ((mi->is_float != 0) != isf) throw ArgError(std::string("map '") + mi->name + "': the result datatype must be " + (mi->is_float ? "floating point" : "integral"));
            mla.cols = ma.cols; mla.stride = (uint32_t)ma.stride; mla.n = ma.n; mla.died = ma.died; mla.op = op; mla.stream = g_stream;
        }
        const bool isb = result_dt == vb::DT_BOOL;
        if (isf && (op == vb::OP_AND || op == vb::OP_OR)) throw AssertionError("& and | are only supported for integer and boolean types");
        const unsigned nb = std::max<unsigned>(1, std::min<unsigned>(1024, nblk(ma.n)));
        double fres = 0; long long ires = 0;
        if (isf) {
            double* part = dalloc<double>(1024 + 1);
            if (mi) { mla.partial = part; mla.nblocks = nb; CK(mi->launch(mla)); ++g_launches; }
            else { mapreduce_kernel<true><<<nb, 256, 0, g_stream>>>(ma, part); LAUNCH_CHECK(); }
            mapreduce_final_kernel<true><<<1, 256, 0, g_stream>>>(part, nb, op, part + 1024); LAUNCH_CHECK();
            CK(cudaMemcpyAsync(&fres, part + 1024, 8, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            dfree(part);
        } else {
            long long* part = dalloc<long long>(1024 + 1);
            if (mi) { mla.partial = part; mla.nblocks = nb; CK(mi->launch(mla)); ++g_launches; }
            else { mapreduce_kernel<false><<<nb, 256, 0, g_stream>>>(ma, part); LAUNCH_CHECK(); }
            mapreduce_final_kernel<false><<<1, 256, 0, g_stream>>>(part, nb, op, part + 1024); LAUNCH_CHECK();
            CK(cudaMemcpyAsync(&ires, part + 1024, 8, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            dfree(part);
        }
        // fold in the reference's start value: `init` or the identity of val4empty (Helpers.jl:44-81)
        auto fold_i = [&](long long a, long long b) -> long long {
            switch (op) {
                case vb::OP_SUM: return (long long)((unsigned long long)a + (unsigned long long)b);
                case vb::OP_PROD: return (long long)((unsigned long long)a * (unsigned long long)b);
                case vb::OP_MIN: return std::min(a, b);
                case vb::OP_MAX: return std::max(a, b);
                case vb::OP_AND: return a & b;
                default: return a | b;
            }
        };
        auto fold_f = [&](double a, double b) -> double {
            switch (op) {
                case vb::OP_SUM: return a + b;
                case vb::OP_PROD: return a * b;
                case vb::OP_MIN: return std::fmin(a, b);
                default: return std::fmax(a, b);
            }
        };
        if (isf) {
            double start;
            if (init) { if (result_dt == vb::DT_F64) std::memcpy(&start, init, 8); else { float f; std::memcpy(&f, init, 4); start = f; } }
            else start = op == vb::OP_PROD ? 1.0 : op == vb::OP_MIN ? INFINITY : op == vb::OP_MAX ? -INFINITY : 0.0;
            double r = ma.n ? fold_f(fres, start) : start;
            if (g_nranks > 1) {   // MPI.Allreduce(reduced, op) in rank order (AgentMethods.jl:555-560)
                std::vector<uint64_t> all;
                allgather8(&r, all);
                std::memcpy(&r, &all[0], 8);
                for (int k = 1; k < g_nranks; ++k) { double v; std::memcpy(&v, &all[k], 8); r = fold_f(v, r); }
            }
            if (result_dt == vb::DT_F64) std::memcpy(out, &r, 8); else { float f = (float)r; std::memcpy(out, &f, 4); }
        } else {
            long long start;
            if (init) {
                switch (result_dt) {
                    case vb::DT_I64: std::memcpy(&start, init, 8); break;
                    case vb::DT_I32: { int v; std::memcpy(&v, init, 4); start = v; break; }
                    default: start = *(const uint8_t*)init; break;
                }
            } else {
                switch (op) {
                    case vb::OP_PROD: start = 1; break;
                    case vb::OP_MAX: start = isb ? -1 : -INT64_MAX; break;
                    case vb::OP_MIN: start = isb ? 1 : INT64_MAX; break;
                    case vb::OP_AND: start = isb ? 1 : INT64_MAX; break;
                    default: start = 0; break;
                }
            }
            long long r = fold_i(ires, start);
            if (g_nranks > 1) {
                std::vector<uint64_t> all;
                allgather8(&r, all);
                r = (long long)all[0];
                for (int k = 1; k < g_nranks; ++k) r = fold_i((long long)all[k], r);
            }
            switch (result_dt) {
                case vb::DT_I64: std::memcpy(out, &r, 8); break;
                case vb::DT_I32: { int v = (int)r; std::memcpy(out, &v, 4); break; }
                case vb::DT_BOOL: { uint8_t v = r != 0; std::memcpy(out, &v, 1); break; }
                default: { uint8_t v = (uint8_t)r; std::memcpy(out, &v, 1); break; }
            }
        }
    });
}

int vb_mapreduce(vb_sim* s, int type_ref, int offset, int dt, int has_cmp, int64_t cmp, int op, int result_dt, const void* init, void* out) {
    return mapreduce_common(s, type_ref, offset, dt, has_cmp, cmp, op, result_dt, init, out, nullptr);
}
int vb_mapreduce_fn(vb_sim* s, const char* map_name, int type_ref, int op, int result_dt, const void* init, void* out) {
    const vb::MapInfo* mi = nullptr;
    int rc = guard([&] {
        const std::string tname = type_ref < vb::EDGE_REF ? s->A(type_ref).name : s->E(type_ref - vb::EDGE_REF).name;
        auto it = map_registry().find({map_name ? map_name : "", tname});
        if (it == map_registry().end()) throw ArgError(std::string("map '") + (map_name ? map_name : "") + "' is not registered for type " + tname);
        mi = it->second;
    });
    if (rc != VB_OK) return rc;
    return mapreduce_common(s, type_ref, 0, vb::DT_I64, 0, 0, op, result_dt, init, out, mi);
}

namespace { size_t dt_size(int dt) { return (dt == VB_DT_I64 || dt == VB_DT_F64) ? 8 : (dt == VB_DT_I32 || dt == VB_DT_F32) ? 4 : 1; } }

int vb_rastervalues(vb_sim* s, const char* name, int offset, int dt, void* out) {
    return guard([&] {   // Raster.jl:282-387
        require_device();
        if (!s->initialized) throw AssertionError("rastervalues can be only called after finish_init!");
        RasterStore& r = find_raster(s, name);
        AgentStore& a = s->A(r.type);
        const size_t w = dt_size(dt), n = r.ids.size();
        uint8_t* tmp = (uint8_t*)g_pool.alloc(n * w);
        field_out_kernel<<<nblk(n), 256, 0, g_stream>>>(a.rstate(), a.stride(), a.word, r.cells, s->base[r.type], n, offset, (uint32_t)w, tmp); LAUNCH_CHECK();
        raster_join(r, tmp, n * w);
        CK(cudaMemcpyAsync(out, tmp, n * w, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        dfree(tmp);
    });
}
int vb_calc_rasterstate_fn(vb_sim* s, const char* name, const char* map_name, int is_float_out, void* out) {
    return guard([&] {   // Raster.jl:238-280 with f = a registered map functor
        require_device();
        if (!s->initialized) throw AssertionError("calc_rasterstate can be only called after finish_init!");
        RasterStore& r = find_raster(s, name);
        AgentStore& a = s->A(r.type);
        auto it = map_registry().find({map_name ? map_name : "", a.name});
        if (it == map_registry().end()) throw ArgError(std::string("map '") + (map_name ? map_name : "") + "' is not registered for type " + a.name);
        const vb::MapInfo* mi = it->second;
        if (mi->elem_size != a.size) throw ArgError(std::string("map '") + mi->name + "': sizeof(Elem) does not match the registered size of " + a.name);
        if ((mi->is_float != 0) != (is_float_out != 0)) throw ArgError(std::string("map '") + mi->name + "': the result datatype must be " + (mi->is_float ? "floating point" : "integral"));
        const size_t n = r.ids.size();
        uint8_t* tmp = (uint8_t*)g_pool.alloc(std::max<size_t>(n, 1) * 8);
        vb::MapCellsArgs ca{};
        ca.cols = a.rstate(); ca.stride = a.stride(); ca.cells = r.cells; ca.cbase = s->base[r.type]; ca.n = n; ca.out = tmp; ca.stream = g_stream;
        CK(mi->launch_cells(ca)); ++g_launches;
        raster_join(r, tmp, n * 8);
        CK(cudaMemcpyAsync(out, tmp, n * 8, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        dfree(tmp);
    });
}
int vb_calc_raster_num_edges(vb_sim* s, const char* name, int ei, int64_t* out) {
    return guard([&] {   // Raster.jl:206-236 with f = id -> num_edges(sim, id, E)
        require_device();
        if (!s->initialized) throw AssertionError("calc_raster can be only called after finish_init!");
        RasterStore& r = find_raster(s, name);
        EdgeStore& e = s->E(ei);
        avail(!e.singleedge, "num_edges");
        s->merge_pending(ei);
        const size_t n = r.ids.size();
        if (e.implicit_stencil) {
            for (size_t i = 0; i < n; ++i) {
                const uint64_t slot = vb::agent_nr(r.ids[i]) - 1;
                const RasterStore& er = s->rasters[e.st_raster];
                out[i] = (r.type == er.type && slot >= e.st_slot0 && slot - e.st_slot0 < er.ids.size()) ? (int64_t)s->stencil_row_host(e, slot - e.st_slot0).size() : 0;
            }
            return;
        }
        long long* tmp = dalloc<long long>(n);
        const uint32_t shift = e.singletype ? s->base[e.target] : 0;
        if (e.singletype && e.target != r.type) {
            CK(cudaMemsetAsync(tmp, 0, n * 8, g_stream));
            raster_join(r, tmp, n * 8);               // (collective on a distributed raster: every rank takes part)
            CK(cudaStreamSynchronize(g_stream));
            std::memset(out, 0, n * 8); dfree(tmp); return;
        }
        raster_num_edges_kernel<<<nblk(n), 256, 0, g_stream>>>(r.cells, n, e.kind == vb::KIND_CSR ? e.off : nullptr, e.cnt, e.rows, shift, tmp); LAUNCH_CHECK();
        raster_join(r, tmp, n * 8);
        CK(cudaMemcpyAsync(out, tmp, n * 8, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        dfree(tmp);
    });
}
int vb_raster_info(vb_sim* s, const char* name, int* ndims, int64_t* dims, vb_agent_id* ids) {
    return guard([&] {
        RasterStore& r = find_raster(s, name);
        *ndims = (int)r.dims.size();
        if (dims) for (size_t i = 0; i < r.dims.size(); ++i) dims[i] = r.dims[i];
        if (ids) std::memcpy(ids, r.ids.data(), r.ids.size() * 8);
    });
}
int vb_num_transitions(vb_sim* s, int64_t* n) { *n = s->num_transitions; return VB_OK; }
int vb_last_apply_stats(vb_sim* s, double* ms_rw, double* ms_fin, uint64_t* er, uint64_t* ea, uint64_t* ac, uint64_t* kl) {
    s->fetch_times();
    if (s->stats_pending) {
        s->stats_pending = false;
        std::vector<unsigned long long> hs(4096);
        if (cudaMemcpyAsync(hs.data(), s->d_stats, 4096 * 8, cudaMemcpyDeviceToHost, g_stream) == cudaSuccess && cudaStreamSynchronize(g_stream) == cudaSuccess) {
            unsigned long long t = 0;
            for (int i = 0; i < 1024; ++i) t += hs[(size_t)i * 4];
            s->st_edges_read = t;
        }
    }
    if (ms_rw) *ms_rw = s->ms_rw; if (ms_fin) *ms_fin = s->ms_fin;
    if (er) *er = s->st_edges_read; if (ea) *ea = s->st_edges_appended; if (ac) *ac = s->st_agents_called; if (kl) *kl = s->st_launches;
    return VB_OK;
}
int vb_last_kernel_ms(vb_sim* s, double* ms) { *ms = s->ms_kernel; return VB_OK; }
int vb_set_read_prefilter(vb_sim* s, int on) { if (!s) return VB_ERR_ARG; s->blk_prefilter = on; return VB_OK; }
int vb_last_apply_prefiltered(vb_sim* s, int* on) { if (on) *on = s->last_prefiltered ? 1 : 0; return VB_OK; }
int vb_last_pass_rate(vb_sim* s, double* rate) { if (rate) *rate = s->last_pass_rate; return VB_OK; }
int vb_set_read_blocking(vb_sim* s, double block_mb, double min_mb, int eager) {
    s->blk_block_mb = block_mb; s->blk_min_mb = min_mb; s->blk_eager = eager;
    return VB_OK;
}
int vb_last_apply_blocks(vb_sim* s, uint32_t* nb) { if (nb) *nb = s->last_blocked_nb; return VB_OK; }
uint64_t vb_device_view_bytes(void) { return sizeof(vb::DeviceSim); }   // host->device bytes uploaded per transition launch

}  // extern "C"
