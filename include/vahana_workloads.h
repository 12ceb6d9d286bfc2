/* vahana_workloads.h — synthetic workload generators for the BASELINE configs (bench.py and tests).
 * Not part of the reference's API: the reference builds its graphs with Graphs.jl / SNAPDatasets on the
 * host (docs/examples/hegselmann.jl:74-95); at 1e8 agents / 2e9 edges the inputs are generated on
 * device and handed to the engine through the ordinary bulk entry points vb_add_agents / vb_add_edges.
 *
 * HK power-law graph: for target i (0-based) the in-degree is
 *     d_i = min(dmax, floor(c * (1 - u_i)^(-2/3))),  u_i = Philox(seed_graph, i, 0)   (Pareto, gamma = 2.5)
 * its k-th source is  floor(N * v^2), v = Philox(seed_graph, first_edge_index(i) + k, 1)  (hub skewed),
 * followed by one self loop (docs/examples/hegselmann.jl:94); opinion_i = Philox(seed_opinion, i, 0).
 * Row order = sources in k order, then the self loop.  Agent ids: type `agent_type`, rank 0, nr = i + 1. */
#ifndef VAHANA_WORKLOADS_H
#define VAHANA_WORKLOADS_H
#include <stdint.h>
#include "vahana_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* device generation + ingestion into `sim` (CUDA engine only).  Targets [0, n) in chunks of `chunk_targets`. */
int vbw_hk_powerlaw_build(vb_sim* sim, int agent_type, int edge_type, uint64_t n, uint64_t seed_graph, uint64_t seed_opinion,
                          double c, uint32_t dmax, uint64_t chunk_targets, uint64_t* n_edges_out);
/* the same for rank `rank` of `nranks`: builds block `rank` of the contiguous equal partition of the n agents
 * (src/Simulation.jl:353-367) — its agents and every edge whose target it owns; sources carry their owner's rank. */
int vbw_hk_powerlaw_build_sharded(vb_sim* sim, int agent_type, int edge_type, uint64_t n, uint64_t seed_graph, uint64_t seed_opinion,
                                  double c, uint32_t dmax, uint64_t chunk_targets, uint32_t rank, uint32_t nranks, uint64_t* n_edges_out);
/* host generation of the same graph (parity tests, CPU baseline): first call with from/to == NULL to get the
 * edge count, then with caller-allocated arrays.  opinions may be NULL. */
int vbw_hk_powerlaw_host(uint64_t n, int agent_type, uint64_t seed_graph, uint64_t seed_opinion, double c, uint32_t dmax,
                         vb_agent_id* from_out, vb_agent_id* to_out, double* opinions_out, uint64_t* n_edges_out);

#ifdef __cplusplus
}
#endif
#endif
