// generators.cu — device-side synthetic workload generation (include/vahana_workloads.h).  Uses only the
// public C-ABI of the engine (vb_add_agents / vb_add_edges with device pointers).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../../include/vahana_workloads.h"
#include "../engine/primitives.cuh"
#include "hk_powerlaw.h"

namespace {
__global__ void hk_degree_kernel(uint64_t seed, uint64_t i0, uint64_t n, double c, uint32_t dmax, uint32_t* __restrict__ deg) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) deg[i] = vbw::hk_degree(seed, i0 + i, c, dmax);
}
__global__ void hk_opinion_kernel(uint64_t seed, uint64_t i0, uint64_t n, double* __restrict__ op) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) op[i] = vbw::hk_opinion(seed, i0 + i);
}
// one warp per target row: row i of the chunk holds deg[i] sources then the self loop; positions include one
// extra slot per row for the self loop: pos = off[i] + i
__global__ void hk_fill_kernel(uint64_t seed, uint64_t i0, uint64_t n, uint64_t nglobal, uint64_t edge0, int type, uint32_t nranks,
                               const uint32_t* __restrict__ off, const uint32_t* __restrict__ deg, uint64_t* __restrict__ from, uint64_t* __restrict__ to) {
    const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (w >= n) return;
    const uint64_t o = off[w], d = deg[w];
    const uint64_t tid = vbw::hk_global_id(type, i0 + w, nglobal, nranks);
    const uint64_t p0 = o + w;
    for (uint64_t k = lane; k < d; k += 32) {
        from[p0 + k] = vbw::hk_global_id(type, vbw::hk_source(seed, edge0 + o + k, nglobal), nglobal, nranks);
        to[p0 + k] = tid;
    }
    if (lane == 0) { from[p0 + d] = tid; to[p0 + d] = tid; }
}
}  // namespace

extern "C" int vbw_hk_powerlaw_build(vb_sim* sim, int agent_type, int edge_type, uint64_t n, uint64_t seed_graph, uint64_t seed_opinion, double c,
                                     uint32_t dmax, uint64_t chunk_targets, uint64_t* n_edges_out) {
    return vbw_hk_powerlaw_build_sharded(sim, agent_type, edge_type, n, seed_graph, seed_opinion, c, dmax, chunk_targets, 0, 1, n_edges_out);
}

// rank `rank` of `nranks` builds the block [lo, hi) of the equal partition: its agents and all edges whose target it owns
extern "C" int vbw_hk_powerlaw_build_sharded(vb_sim* sim, int agent_type, int edge_type, uint64_t nglobal, uint64_t seed_graph, uint64_t seed_opinion,
                                             double c, uint32_t dmax, uint64_t chunk_targets, uint32_t rank, uint32_t nranks, uint64_t* n_edges_out) {
    if (chunk_targets == 0) chunk_targets = 1u << 22;
    cudaStream_t st = nullptr;
    const uint64_t q = nglobal / nranks, rem = nglobal % nranks;
    const uint64_t lo = rank * q + (rank < rem ? rank : rem), hi = lo + q + (rank < rem ? 1 : 0);
    const uint64_t n = hi;   // loops below run over global target indices [lo, hi)
    // agents: opinions generated on device, added in chunks
    {
        double* op = nullptr;
        if (cudaMalloc(&op, chunk_targets * 8) != cudaSuccess) return VB_ERR_CUDA;
        for (uint64_t i0 = lo; i0 < n; i0 += chunk_targets) {
            const uint64_t m = n - i0 < chunk_targets ? n - i0 : chunk_targets;
            hk_opinion_kernel<<<vbp::nblk(m), 256, 0, st>>>(seed_opinion, i0, m, op);
            if (cudaStreamSynchronize(st) != cudaSuccess) { cudaFree(op); return VB_ERR_CUDA; }
            int rc = vb_add_agents(sim, agent_type, op, m, nullptr);
            if (rc != VB_OK) { cudaFree(op); return rc; }
        }
        cudaFree(op);
    }
    uint32_t *deg = nullptr, *off = nullptr, *scr = nullptr, *tot = nullptr;
    cudaMalloc(&deg, chunk_targets * 4); cudaMalloc(&off, chunk_targets * 4); cudaMalloc(&tot, 4);
    cudaMalloc(&scr, vbp::scan_scratch_words(chunk_targets) * 4);
    uint64_t *from = nullptr, *to = nullptr, cap = 0;
    uint64_t edge0 = 0, total = 0;
    int rc = VB_OK;
    // the k-th source of a target is keyed by its global edge index: sum the degrees of all targets before `lo`
    for (uint64_t i0 = 0; i0 < lo; i0 += chunk_targets) {
        const uint64_t m = lo - i0 < chunk_targets ? lo - i0 : chunk_targets;
        hk_degree_kernel<<<vbp::nblk(m), 256, 0, st>>>(seed_graph, i0, m, c, dmax, deg);
        vbp::exclusive_scan(deg, off, m, tot, scr, st);
        uint32_t sum = 0;
        cudaMemcpyAsync(&sum, tot, 4, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) return VB_ERR_CUDA;
        edge0 += sum;
    }
    for (uint64_t i0 = lo; i0 < n && rc == VB_OK; i0 += chunk_targets) {
        const uint64_t m = n - i0 < chunk_targets ? n - i0 : chunk_targets;
        hk_degree_kernel<<<vbp::nblk(m), 256, 0, st>>>(seed_graph, i0, m, c, dmax, deg);
        vbp::exclusive_scan(deg, off, m, tot, scr, st);
        uint32_t sum = 0;
        cudaMemcpyAsync(&sum, tot, 4, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = VB_ERR_CUDA; break; }
        const uint64_t ne = (uint64_t)sum + m;
        if (ne > cap) {
            cudaFree(from); cudaFree(to);
            cap = ne + ne / 8;
            if (cudaMalloc(&from, cap * 8) != cudaSuccess || cudaMalloc(&to, cap * 8) != cudaSuccess) { rc = VB_ERR_CUDA; break; }
        }
        hk_fill_kernel<<<vbp::nblk(m * 32), 256, 0, st>>>(seed_graph, i0, m, nglobal, edge0, agent_type, nranks, off, deg, from, to);
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = VB_ERR_CUDA; break; }
        rc = vb_add_edges(sim, edge_type, from, to, nullptr, ne);
        edge0 += sum;
        total += ne;
    }
    cudaFree(deg); cudaFree(off); cudaFree(scr); cudaFree(tot); cudaFree(from); cudaFree(to);
    if (n_edges_out) *n_edges_out = total;
    return rc;
}

extern "C" int vbw_hk_powerlaw_host(uint64_t n, int agent_type, uint64_t seed_graph, uint64_t seed_opinion, double c, uint32_t dmax,
                                    vb_agent_id* from_out, vb_agent_id* to_out, double* opinions_out, uint64_t* n_edges_out) {
    return vbw::hk_powerlaw_host(n, agent_type, seed_graph, seed_opinion, c, dmax, from_out, to_out, opinions_out, n_edges_out);
}
