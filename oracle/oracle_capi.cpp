// oracle_capi.cpp — CPU ORACLE (test infrastructure): exports the same C-ABI as the CUDA engine
// (include/vahana_b200.h) on top of oracle/vahana_oracle.hpp, so the very same host-side test code
// can drive both and compare.  vb_backend() returns "oracle-cpu"; the product refuses to run on it.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../include/vahana_b200.h"
#include "vahana_oracle.hpp"

static thread_local std::string g_err;

struct vb_sim { std::unique_ptr<vo::Sim> s; };

template <class F>
static int guard(F&& f) {
    try { f(); return VB_OK; }
    catch (const vo::AssertionError& e) { g_err = e.what(); return VB_ERR_ASSERT; }
    catch (const vo::ArgError& e) { g_err = e.what(); return VB_ERR_ARG; }
    catch (const std::exception& e) { g_err = e.what(); return VB_ERR_STATE; }
}

extern "C" {

const char* vb_last_error(void) { return g_err.c_str(); }
const char* vb_backend(void) { return "oracle-cpu"; }
int vb_init(int) { return VB_OK; }
int vb_shutdown(void) { return VB_OK; }
int vb_set_stream(void*) { return VB_OK; }
uint64_t vb_device_view_bytes(void) { return 0; }
int vb_last_kernel_ms(vb_sim*, double* ms) { *ms = 0; return VB_OK; }
int vb_comm_unique_id(uint8_t id_out[128]) { std::memset(id_out, 0, 128); return VB_OK; }
int vb_comm_init(int, int nranks, const uint8_t*) { if (nranks != 1) { g_err = "oracle C-ABI is single rank"; return VB_ERR_ARG; } return VB_OK; }
int vb_comm_rank(int* r, int* n) { *r = 0; *n = 1; return VB_OK; }
int vb_halo_bytes(vb_sim*, uint64_t* out) { *out = 0; return VB_OK; }
int vb_last_halo_ms(vb_sim*, double* ms) { if (ms) *ms = -1.0; return VB_OK; }
int vb_set_uniform_offset(vb_sim*, int, uint64_t) { return VB_OK; }

int vb_sim_create(const vb_model_desc* m, const void* params, vb_sim** out) {
    return guard([&] {
        std::vector<vo::AgentTypeDesc> ats;
        std::vector<vo::EdgeTypeDesc> ets;
        for (uint32_t i = 0; i < m->n_agent_types; ++i) ats.push_back({m->agent_types[i].name, m->agent_types[i].size, m->agent_types[i].hints});
        for (uint32_t i = 0; i < m->n_edge_types; ++i)
            ets.push_back({m->edge_types[i].name, m->edge_types[i].size, m->edge_types[i].hints, m->edge_types[i].target_type, m->edge_types[i].size_hint});
        auto* h = new vb_sim;
        h->s = vo::create_sim(m->name, ats, ets, params, m->param_size);
        *out = h;
    });
}
int vb_sim_copy(const vb_sim* sim, vb_sim** out) { return guard([&] { auto* h = new vb_sim; h->s = vo::copy_sim(*sim->s); *out = h; }); }
int vb_sim_destroy(vb_sim* sim) { delete sim; return VB_OK; }
int vb_set_param(vb_sim* sim, const void* p, uint32_t size) {
    return guard([&] {
        if (sim->s->initialized) throw vo::AssertionError("set_param! can only be called before finish_init!");
        sim->s->params.assign((const uint8_t*)p, (const uint8_t*)p + size);
    });
}
int vb_set_config(vb_sim* sim, int asserts, int check_readable) { sim->s->asserts_enabled = asserts; sim->s->check_readable = check_readable; return VB_OK; }
int vb_disable_transition_checks(vb_sim* sim, int disable) {   // Helpers.jl:259: intransition := disable, check_readable := !disable
    sim->s->intransition = disable;
    sim->s->check_readable = !disable;
    return VB_OK;
}

int vb_add_agents(vb_sim* sim, int type, const void* states, uint64_t n, vb_agent_id* ids_out) {
    return guard([&] {
        const uint32_t sz = sim->s->A(type).desc.size;
        for (uint64_t i = 0; i < n; ++i) {
            vb_agent_id id = vo::add_agent(*sim->s, type, sz ? (const uint8_t*)states + i * sz : nullptr);
            if (ids_out) ids_out[i] = id;
        }
    });
}
int vb_add_agent_per_process(vb_sim* sim, int type, const void* state, vb_agent_id* id_out) {
    return guard([&] {   // Agent.jl:363-388
        vo::Sim& s = *sim->s;
        if (!s.initialized) throw vo::AssertionError("add_agent_per_process! can only be called after finish_init!");
        if (s.intransition) throw vo::AssertionError("add_agent_per_process! cannot be called within a transition function");
        vo::prepare_write_agent(s, type, true);
        s.intransition = true;
        vb_agent_id id = vo::add_agent(s, type, state);
        s.intransition = false;
        vo::finish_write_agent(s, type);
        if (id_out) *id_out = id;
    });
}
int vb_add_edges(vb_sim* sim, int e, const vb_agent_id* from, const vb_agent_id* to, const void* states, uint64_t n) {
    return guard([&] {
        const uint32_t sz = sim->s->E(e).desc.size;
        for (uint64_t i = 0; i < n; ++i) vo::add_edge(*sim->s, e, from ? from[i] : 0, to[i], (sz && states) ? (const uint8_t*)states + i * sz : nullptr);
    });
}
int vb_remove_edges(vb_sim* sim, int e, vb_agent_id from, vb_agent_id to) {
    return guard([&] { if (from) vo::remove_edges_from_to(*sim->s, e, from, to); else vo::remove_edges_to(*sim->s, e, to); });
}
int vb_add_raster(vb_sim* sim, const char* name, int ndims, const int64_t* dims, int type, const void* states, vb_agent_id* ids_out) {
    return guard([&] { vo::add_raster(*sim->s, name, std::vector<int64_t>(dims, dims + ndims), type, states, ids_out); });
}
int vb_set_raster(vb_sim* sim, const char* name, int ndims, const int64_t* dims, int type, const vb_agent_id* ids) {
    return guard([&] {   // a raster over existing agents (broadcastids, src/MPI.jl:59-73); one rank: all of them are local
        (void)type;
        if (sim->s->initialized) throw vo::AssertionError("rasters can only be defined before finish_init!");
        vo::Raster r;
        r.name = name; r.dims.assign(dims, dims + ndims);
        size_t n = 1;
        for (int i = 0; i < ndims; ++i) n *= (size_t)dims[i];
        r.ids.assign(ids, ids + n);
        sim->s->rasters.push_back(std::move(r));
    });
}
int vb_connect_raster_neighbors(vb_sim* sim, const char* name, int e, double distance, int metric, int periodic, const void* st) {
    return guard([&] { vo::connect_raster_neighbors(*sim->s, name, e, distance, metric, periodic, st); });
}
int vb_move_to(vb_sim* sim, const char* name, vb_agent_id id, const int64_t* pos, int e_from, const void* s_from, int e_to,
               const void* s_to, double distance, int metric, int periodic, int only_surrounding) {
    return guard([&] { vo::move_to(*sim->s, vo::find_raster(*sim->s, name), id, pos, e_from, s_from, e_to, s_to, distance, metric, periodic, only_surrounding); });
}
int vb_cellid(vb_sim* sim, const char* name, const int64_t* pos, vb_agent_id* out) {
    return guard([&] {
        vo::Raster& r = vo::find_raster(*sim->s, name);
        std::vector<int64_t> p(pos, pos + r.dims.size());
        for (size_t k = 0; k < p.size(); ++k) if (p[k] < 1 || p[k] > r.dims[k]) throw vo::AssertionError("cellid: position outside the raster");
        *out = r.ids[vo::linear_index(p, r.dims)];
    });
}
int vb_finish_init(vb_sim* sim) { return guard([&] { vo::finish_init(*sim->s); }); }

int vb_apply(vb_sim* sim, const char* transition, const int* call, int ncall, const int* read, int nread, const int* write,
             int nwrite, const int* add_existing, int nadd, int with_edge, uint64_t seed) {
    return guard([&] {
        vo::apply(*sim->s, transition, std::vector<int>(call, call + ncall), std::vector<int>(read, read + nread),
                  std::vector<int>(write, write + nwrite), std::vector<int>(add_existing, add_existing + nadd), with_edge, seed);
    });
}
int vb_has_transition(const char* t, const char* a) { return vo::Registry::get().fns.count({t, a}) ? 1 : 0; }
int vb_load_model_library(const char*) { g_err = "oracle: model libraries are compiled in"; return VB_ERR_ARG; }

int vb_num_agents(vb_sim* sim, int type, uint64_t* n) { return guard([&] { *n = vo::num_agents(*sim->s, type); }); }
int vb_all_agents(vb_sim* sim, int type, void* states_out, vb_agent_id* ids_out, uint64_t cap, uint64_t* n_out) {
    return guard([&] {   // Agent.jl:234-313
        vo::Sim& s = *sim->s;
        vo::AgentFields& a = s.A(type);
        const vo::AgentRW& rw = s.initialized ? a.read : a.write;
        const uint32_t sz = a.desc.size;
        uint64_t n = 0;
        for (uint64_t i = 0; i < rw.nslots; ++i) {
            if (!a.immortal && rw.died[i]) continue;
            if (n < cap) {
                if (states_out && sz) std::memcpy((uint8_t*)states_out + n * sz, &rw.state[i * sz], sz);
                if (ids_out) ids_out[n] = vb::agent_id((uint32_t)type, s.rank, i + 1);
            }
            ++n;
        }
        *n_out = n;
    });
}
int vb_agentstate(vb_sim* sim, vb_agent_id id, int type, void* out) {
    return guard([&] { const void* p = vo::agentstate(*sim->s, id, type); std::memcpy(out, p, sim->s->A(type).desc.size); });
}
int vb_num_edges_total(vb_sim* sim, int e, int write, uint64_t* n) {   // Edge.jl:373-389: before init the write container is read
    return guard([&] { *n = vo::num_edges_total(*sim->s, e, write != 0); });
}
int vb_edges_of(vb_sim* sim, int e, vb_agent_id to, int what, vb_agent_id* from_out, void* states_out, uint64_t cap, int64_t* n_out) {
    return guard([&] {
        vo::Sim& s = *sim->s;
        vo::EdgeFields& f = s.E(e);
        const bool S = f.stateless, I = f.ignorefrom, E1 = f.singleedge, T = f.singletype;
        // availability matrix: EdgeMethods.jl:699-892, docs/src/performance.md:129-136
        switch (what) {
            case VB_ACC_EDGES: vo::Ctx::avail(!S && !I, "edges"); break;
            case VB_ACC_NEIGHBORIDS: vo::Ctx::avail(!I, "neighborids"); break;
            case VB_ACC_NEIGHBORIDS_ITER: vo::Ctx::avail(!I && !E1, "neighborids_iter"); break;
            case VB_ACC_EDGESTATES: vo::Ctx::avail(!S, "edgestates"); break;
            case VB_ACC_EDGESTATES_ITER: vo::Ctx::avail(!S && !E1, "edgestates_iter"); break;
            case VB_ACC_NUM_EDGES: vo::Ctx::avail(!E1, "num_edges"); break;
            case VB_ACC_HAS_EDGE: vo::Ctx::avail(!(E1 && T && !(S && I)), "has_edge"); break;
            default: throw vo::ArgError("bad accessor");
        }
        vo::Row* r;
        if (what == VB_ACC_HAS_EDGE && E1 && !T && !(S && I)) {   // haskey(read, to): EdgeMethods.jl:882-885
            s.mayassert(f.readable || !s.check_readable, "edge type is not in the `read` argument of apply!");
            r = f.read->dict.find(to);
            *n_out = r ? 1 : 0;
            return;
        }
        r = vo::get_container(s, f, *f.read, to);
        if (what == VB_ACC_NUM_EDGES || what == VB_ACC_HAS_EDGE) { *n_out = r ? r->count : 0; return; }
        if (!r) { *n_out = -1; return; }
        if (S && I) { *n_out = r->count; return; }
        const uint32_t sz = f.desc.size;
        uint64_t n = I ? (sz ? r->state.size() / sz : 0) : r->from.size();
        for (uint64_t i = 0; i < n && i < cap; ++i) {
            if (from_out && !I) from_out[i] = r->from[i];
            if (states_out && !S && sz) std::memcpy((uint8_t*)states_out + i * sz, &r->state[i * sz], sz);
        }
        *n_out = (int64_t)n;
    });
}
int vb_all_edges(vb_sim* sim, int e, vb_agent_id* to_out, vb_agent_id* from_out, void* states_out, uint64_t cap, uint64_t* n_out) {
    return guard([&] {   // EdgeMethods.jl:1005-1027 / EdgeIterator.jl:45-106, emitted in ascending target order
        vo::Sim& s = *sim->s;
        vo::EdgeFields& f = s.E(e);
        const vo::EdgeContainer& c = s.initialized ? *f.read : *f.write;
        const uint32_t sz = f.desc.size;
        std::vector<std::pair<uint64_t, const vo::Row*>> rows;
        if (f.singletype) { for (size_t i = 0; i < c.vec.size(); ++i) if (c.assigned[i]) rows.push_back({vb::agent_id((uint32_t)f.desc.target, s.rank, i + 1), &c.vec[i]}); }
        else c.dict.for_each([&](uint64_t k, const vo::Row& r) { rows.push_back({k, &r}); });
        std::sort(rows.begin(), rows.end(), [](auto& a, auto& b) { return a.first < b.first; });
        uint64_t n = 0;
        for (auto& kv : rows) {
            const vo::Row& r = *kv.second;
            for (int64_t i = 0; i < r.count; ++i) {
                if (n < cap) {
                    if (to_out) to_out[n] = kv.first;
                    if (from_out) from_out[n] = (!f.ignorefrom && (size_t)i < r.from.size()) ? r.from[i] : 0;
                    if (states_out && sz && !f.stateless && (size_t)(i + 1) * sz <= r.state.size()) std::memcpy((uint8_t*)states_out + n * sz, &r.state[i * sz], sz);
                }
                ++n;
            }
        }
        *n_out = n;
    });
}

static vo::Value to_value(const void* p, int dt) {
    vo::Value v; v.dt = dt;
    switch (dt) {
        case VB_DT_I64: { int64_t x; std::memcpy(&x, p, 8); v.i = x; break; }
        case VB_DT_F64: { double x; std::memcpy(&x, p, 8); v.f = x; break; }
        case VB_DT_BOOL: case VB_DT_U8: v.i = *(const uint8_t*)p; break;
        case VB_DT_I32: { int32_t x; std::memcpy(&x, p, 4); v.i = x; break; }
        case VB_DT_F32: { float x; std::memcpy(&x, p, 4); v.f = x; break; }
    }
    return v;
}
static void from_value(const vo::Value& v, int dt, void* out) {
    switch (dt) {
        case VB_DT_I64: { int64_t x = v.i; std::memcpy(out, &x, 8); break; }
        case VB_DT_F64: { double x = v.f; std::memcpy(out, &x, 8); break; }
        case VB_DT_BOOL: case VB_DT_U8: { uint8_t x = (uint8_t)(v.i != 0 ? (dt == VB_DT_BOOL ? 1 : v.i) : 0); std::memcpy(out, &x, 1); break; }
        case VB_DT_I32: { int32_t x = (int32_t)v.i; std::memcpy(out, &x, 4); break; }
        case VB_DT_F32: { float x = (float)v.f; std::memcpy(out, &x, 4); break; }
    }
}
int vb_mapreduce(vb_sim* sim, int type_ref, int offset, int dt, int has_cmp, int64_t cmp, int op, int result_dt, const void* init, void* out) {
    return guard([&] {
        vo::Value iv;
        if (init) iv = to_value(init, result_dt);
        vo::Value r = vo::mapreduce(*sim->s, type_ref, offset, dt, has_cmp != 0, cmp, op, result_dt, init ? &iv : nullptr);
        from_value(r, result_dt, out);
    });
}

int vb_mapreduce_fn(vb_sim* sim, const char* map_name, int type_ref, int op, int result_dt, const void* init, void* out) {
    return guard([&] {
        vo::Sim& s = *sim->s;
        const std::string tname = type_ref < vb::EDGE_REF ? s.A(type_ref).desc.name : s.E(type_ref - vb::EDGE_REF).desc.name;
        auto& maps = vo::Registry::get().maps;
        auto it = maps.find({map_name ? map_name : "", tname});
        if (it == maps.end()) throw vo::ArgError(std::string("map '") + (map_name ? map_name : "") + "' is not registered for type " + tname);
        vo::Value iv;
        if (init) iv = to_value(init, result_dt);
        vo::Value r = vo::mapreduce(s, type_ref, 0, vb::DT_I64, false, 0, op, result_dt, init ? &iv : nullptr, &it->second);
        from_value(r, result_dt, out);
    });
}

static size_t dt_size(int dt) { return (dt == VB_DT_I64 || dt == VB_DT_F64) ? 8 : (dt == VB_DT_I32 || dt == VB_DT_F32) ? 4 : 1; }
int vb_rastervalues(vb_sim* sim, const char* name, int offset, int dt, void* out) {   // Raster.jl:282-387
    return guard([&] {
        vo::Sim& s = *sim->s;
        if (!s.initialized) throw vo::AssertionError("rastervalues can be only called after finish_init!");
        vo::Raster& r = vo::find_raster(s, name);
        const size_t w = dt_size(dt);
        bool cr = s.check_readable; s.check_readable = false;
        for (size_t i = 0; i < r.ids.size(); ++i) {
            const uint8_t* p = (const uint8_t*)vo::agentstate(s, r.ids[i], (int)vb::type_nr(r.ids[i]));
            std::memcpy((uint8_t*)out + i * w, p + offset, w);
        }
        s.check_readable = cr;
    });
}
int vb_calc_rasterstate_fn(vb_sim* sim, const char* name, const char* map_name, int is_float_out, void* out) {   // Raster.jl:238-280
    return guard([&] {
        vo::Sim& s = *sim->s;
        if (!s.initialized) throw vo::AssertionError("calc_rasterstate can be only called after finish_init!");
        vo::Raster& r = vo::find_raster(s, name);
        if (r.ids.empty()) return;
        const int type = (int)vb::type_nr(r.ids[0]);
        auto& maps = vo::Registry::get().maps;
        auto it = maps.find({map_name ? map_name : "", s.A(type).desc.name});
        if (it == maps.end()) throw vo::ArgError(std::string("map '") + (map_name ? map_name : "") + "' is not registered for type " + s.A(type).desc.name);
        const vo::MapFn& mf = it->second;
        if (mf.elem_size != s.A(type).desc.size) throw vo::ArgError("sizeof(Elem) of the map functor does not match the agent type");
        if (mf.is_float != (is_float_out != 0)) throw vo::ArgError("the result datatype does not match the map functor");
        bool cr = s.check_readable; s.check_readable = false;
        for (size_t i = 0; i < r.ids.size(); ++i) {
            const uint8_t* p = (const uint8_t*)vo::agentstate(s, r.ids[i], type);
            double f = 0; int64_t v = 0;
            mf.fn(p, f, v);
            if (is_float_out) ((double*)out)[i] = f; else ((int64_t*)out)[i] = v;
        }
        s.check_readable = cr;
    });
}
int vb_calc_raster_num_edges(vb_sim* sim, const char* name, int e, int64_t* out) {   // Raster.jl:206-236
    return guard([&] {
        vo::Sim& s = *sim->s;
        if (!s.initialized) throw vo::AssertionError("calc_raster can be only called after finish_init!");
        vo::Raster& r = vo::find_raster(s, name);
        vo::EdgeFields& f = s.E(e);
        bool was = f.readable; f.readable = true;
        vo::Ctx ctx(s);
        for (size_t i = 0; i < r.ids.size(); ++i) out[i] = ctx.num_edges(e, r.ids[i]);
        f.readable = was;
    });
}
int vb_raster_info(vb_sim* sim, const char* name, int* ndims, int64_t* dims, vb_agent_id* ids) {
    return guard([&] {
        vo::Raster& r = vo::find_raster(*sim->s, name);
        *ndims = (int)r.dims.size();
        if (dims) for (size_t i = 0; i < r.dims.size(); ++i) dims[i] = r.dims[i];
        if (ids) std::memcpy(ids, r.ids.data(), r.ids.size() * 8);
    });
}
int vb_num_transitions(vb_sim* sim, int64_t* n) { *n = sim->s->num_transitions; return VB_OK; }

int vb_export_csr(vb_sim* sim, int e, int target_type, uint64_t* offsets, uint64_t nrows, vb_agent_id* from_out, void* states_out, uint64_t cap) {
    return guard([&] {
        vo::Sim& s = *sim->s;
        vo::EdgeFields& f = s.E(e);
        const vo::EdgeContainer& c = *f.read;
        const uint32_t sz = f.desc.size;
        uint64_t n = 0;
        for (uint64_t row = 0; row < nrows; ++row) {
            offsets[row] = n;
            const vo::Row* r = nullptr;
            if (f.singletype) { if (target_type == f.desc.target && row < c.vec.size() && c.assigned[row]) r = &c.vec[row]; }
            else r = c.dict.find(vb::agent_id((uint32_t)target_type, s.rank, row + 1));
            if (!r) continue;
            for (int64_t i = 0; i < r->count; ++i) {
                if (n < cap) {
                    if (from_out) from_out[n] = (!f.ignorefrom && (size_t)i < r->from.size()) ? r->from[i] : 0;
                    if (states_out && sz && !f.stateless && (size_t)(i + 1) * sz <= r->state.size()) std::memcpy((uint8_t*)states_out + n * sz, &r->state[i * sz], sz);
                }
                ++n;
            }
        }
        offsets[nrows] = n;
    });
}
int vb_last_apply_stats(vb_sim* sim, double* ms_rw, double* ms_fin, uint64_t* er, uint64_t* ea, uint64_t* ac, uint64_t* kl) {
    if (ms_rw) *ms_rw = 0; if (ms_fin) *ms_fin = 0;
    if (er) *er = sim->s->st_edges_read; if (ea) *ea = sim->s->st_edges_appended; if (ac) *ac = sim->s->st_agents_called;
    if (kl) *kl = 0;
    return VB_OK;
}
int vb_set_read_prefilter(vb_sim*, int) { return VB_OK; }                        // no keys: every neighbour state is read
int vb_last_apply_prefiltered(vb_sim*, int* on) { if (on) *on = 0; return VB_OK; }
int vb_last_pass_rate(vb_sim*, double* rate) { if (rate) *rate = -1.0; return VB_OK; }
int vb_set_read_blocking(vb_sim*, double, double, int) { return VB_OK; }   // the oracle walks every row left to right
int vb_last_apply_blocks(vb_sim*, uint32_t* nb) { if (nb) *nb = 0; return VB_OK; }

}  // extern "C"
