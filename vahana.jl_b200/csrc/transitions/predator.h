// predator.h — the predator/prey raster model of the reference's docs (BASELINE config 3), restated transition by
// transition from /root/reference/docs/examples/predator.jl (model :55-129, move! :205-209, transitions :252-400, step! :437-469).
//
// Randomness.  The reference draws from Julia's default RNG inside the closures (rand(v), shuffle, rand() * 100); trajectories are
// therefore not reproducible outside Julia ("parity unpinned").  Both implementations here read the same per-agent uniform table
// ctx.uniform(k) and use these documented definitions (SURVEY.md §7 "stochastic models"):
//     rand(v)      := v[floor(u * length(v))]                      (k-th uniform of the agent, k stated at each use)
//     shuffle(v)   := v ordered by ascending per-element keys uniform(16 + index), ties by index
//     rand(Set s)  := the floor(u * |s|)-th remaining element in row order
// try_eat tracks the prey of one cell in a 64-bit mask: at most 64 prey per cell take part (documented bound).
#pragma once
#include "../../../include/vahana_model.h"

namespace pp {

struct Animal { int64_t energy; int64_t pos[2]; };   // Predator and Prey (predator.jl:55-63)
struct Cell { int64_t pos[2]; int64_t countdown; };  // predator.jl:75-78
struct Params {                                      // AllParams / SpeciesParams (predator.jl:131-153), flattened
    int64_t restart;
    int64_t pred_gain, pred_loss, pred_thres, pred_prob;
    int64_t prey_gain, prey_loss, prey_thres, prey_prob;
};
enum : int { T_PREDATOR = 1, T_PREY = 2, T_CELL = 3 };
enum : int { E_POS_PRED = 0, E_POS_PREY = 1, E_VIEW_PRED = 2, E_VIEW_PREY = 3, E_VISIBLE_PREY = 4, E_DIE = 5, E_EAT = 6 };
enum : int { RASTER = 0 };

// move!(sim, id, newpos, Species)  (predator.jl:205-209): 1 Position edge + 2 x 5 View edges
template <class Ctx> VB_HD void move_animal(Ctx& ctx, vb::AgentID id, const int64_t (&pos)[2], int e_position, int e_view) {
    vb::Pos p{{pos[0], pos[1], 0, 0}};
    ctx.move_to(RASTER, id, p, -1, e_position);
    ctx.move_to(RASTER, id, p, e_view, e_view, 1.0, vb::MANHATTEN);
}

// move(state::Prey, id, sim)  (predator.jl:294-310)
struct MovePrey : vb::TransitionBase {
    using State = Animal;
    using EdgeWrites = vb::IntList<E_VIEW_PREY, E_POS_PREY>;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Animal& s, vb::AgentID id) const {
        const Params& pr = ctx.template param<Params>();
        const int64_t e = s.energy - pr.prey_loss;
        if (e <= 0) return false;
        long long nview = 0, ngrass = 0;
        ctx.for_each_neighbor(E_VIEW_PREY, id, [&](vb::AgentID cell) {
            nview += 1;
            if (ctx.template agentstate<Cell>(T_CELL, cell).countdown == 0) ngrass += 1;
        });
        const double u = ctx.uniform(0);
        vb::AgentID next = 0;
        if (ngrass == 0) {
            long long k = (long long)(u * (double)nview);
            if (k >= nview) k = nview - 1;
            next = ctx.neighbor_at(E_VIEW_PREY, id, k);
        } else {
            long long k = (long long)(u * (double)ngrass), seen = 0;
            if (k >= ngrass) k = ngrass - 1;
            ctx.for_each_neighbor(E_VIEW_PREY, id, [&](vb::AgentID cell) {
                if (ctx.template agentstate<Cell>(T_CELL, cell).countdown == 0) { if (seen == k) next = cell; seen += 1; }
            });
        }
        const Cell c = ctx.template agentstate<Cell>(T_CELL, next);
        s.energy = e; s.pos[0] = c.pos[0]; s.pos[1] = c.pos[1];
        move_animal(ctx, id, s.pos, E_POS_PREY, E_VIEW_PREY);
        return true;
    }
};
// find_prey(::Val{Cell}, id, sim)  (predator.jl:252-260)
struct FindPrey : vb::TransitionBase {
    using State = Cell;
    using EdgeWrites = vb::IntList<E_VISIBLE_PREY>;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Cell&, vb::AgentID id) const {
        if (ctx.has_edge(E_POS_PREY, id) && ctx.has_edge(E_VIEW_PRED, id)) {
            ctx.for_each_neighbor(E_POS_PREY, id, [&](vb::AgentID prey) {
                ctx.for_each_neighbor(E_VIEW_PRED, id, [&](vb::AgentID pred) { ctx.add_edge(E_VISIBLE_PREY, prey, pred); });
            });
        }
        return true;
    }
};
// move(state::Predator, id, sim)  (predator.jl:268-285)
struct MovePredator : vb::TransitionBase {
    using State = Animal;
    using EdgeWrites = vb::IntList<E_VIEW_PRED, E_POS_PRED>;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Animal& s, vb::AgentID id) const {
        const Params& pr = ctx.template param<Params>();
        const int64_t e = s.energy - pr.pred_loss;
        if (e <= 0) return false;
        const double u = ctx.uniform(0);
        const long long nprey = ctx.num_edges(E_VISIBLE_PREY, id);
        if (nprey == 0) {   // isnothing(prey): a random visible cell
            const long long nview = ctx.num_edges(E_VIEW_PRED, id);
            long long k = (long long)(u * (double)nview);
            if (k >= nview) k = nview - 1;
            const Cell c = ctx.template agentstate<Cell>(T_CELL, ctx.neighbor_at(E_VIEW_PRED, id, k));
            s.pos[0] = c.pos[0]; s.pos[1] = c.pos[1];
        } else {            // rand(prey).pos
            long long k = (long long)(u * (double)nprey);
            if (k >= nprey) k = nprey - 1;
            const Animal a = ctx.template agentstate<Animal>(T_PREY, ctx.neighbor_at(E_VISIBLE_PREY, id, k));
            s.pos[0] = a.pos[0]; s.pos[1] = a.pos[1];
        }
        s.energy = e;
        move_animal(ctx, id, s.pos, E_POS_PRED, E_VIEW_PRED);
        return true;
    }
};
// grow_food(state::Cell, _, _)  (predator.jl:318-320)
struct GrowFood : vb::TransitionBase {
    using State = Cell;
    template <class Ctx> VB_HD bool operator()(Ctx&, Cell& c, vb::AgentID) const { c.countdown = c.countdown > 1 ? c.countdown - 1 : 0; return true; }
};
// try_eat(state::Cell, id, sim)  (predator.jl:330-354)
struct TryEat : vb::TransitionBase {
    using State = Cell;
    using EdgeWrites = vb::IntList<E_DIE, E_EAT>;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Cell& c, vb::AgentID id) const {
        const Params& pr = ctx.template param<Params>();
        const long long npred = ctx.num_edges(E_POS_PRED, id);
        long long nprey = ctx.num_edges(E_POS_PREY, id);
        if (nprey > 64) nprey = 64;
        uint64_t eaten = 0;
        long long remaining = nprey;
        if (npred > 0 && nprey > 0) {
            // shuffle(predators): visit them in ascending order of their keys
            double last_key = -1.0; long long last_idx = -1;
            for (long long t = 0; t < npred && remaining > 0; ++t) {
                double best = 2.0; long long bi = -1;
                for (long long i = 0; i < npred; ++i) {
                    const double k = ctx.uniform(16 + (int)i);
                    const bool after = k > last_key || (k == last_key && i > last_idx);
                    if (after && (k < best || (k == best && i < bi) || bi < 0)) { best = k; bi = i; }
                }
                last_key = best; last_idx = bi;
                const vb::AgentID pred = ctx.neighbor_at(E_POS_PRED, id, bi);
                long long j = (long long)(ctx.uniform(1 + (int)t < 16 ? 1 + (int)t : 15) * (double)remaining);   // p = rand(prey)
                if (j >= remaining) j = remaining - 1;
                long long idx = 0;
                for (long long i = 0; i < nprey; ++i) { if (eaten >> i & 1) continue; if (j == 0) { idx = i; break; } --j; }
                eaten |= 1ull << idx;
                remaining -= 1;
                ctx.add_edge(E_DIE, id, ctx.neighbor_at(E_POS_PREY, id, idx));
                ctx.add_edge(E_EAT, id, pred);
            }
        }
        if (remaining > 0 && c.countdown == 0) {   // prey left that can eat the grass
            long long j = (long long)(ctx.uniform(0) * (double)remaining);
            if (j >= remaining) j = remaining - 1;
            long long idx = 0;
            for (long long i = 0; i < nprey; ++i) { if (eaten >> i & 1) continue; if (j == 0) { idx = i; break; } --j; }
            ctx.add_edge(E_EAT, id, ctx.neighbor_at(E_POS_PREY, id, idx));
            c.countdown = pr.restart;
        }
        return true;
    }
};
// try_reproduce(state, id, sim)  (predator.jl:364-392); Int64(round(energy / 2)) rounds half to even (SURVEY A-37)
template <int T, int E_POS, int E_VIEW, bool kIsPrey> struct TryReproduce : vb::TransitionBase {
    using State = Animal;
    using EdgeWrites = vb::IntList<E_POS, E_VIEW>;
    using AgentWrites = vb::IntList<T>;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Animal& s, vb::AgentID id) const {
        const Params& pr = ctx.template param<Params>();
        if (kIsPrey && ctx.has_edge(E_DIE, id)) return false;
        const int64_t gain = kIsPrey ? pr.prey_gain : pr.pred_gain, thres = kIsPrey ? pr.prey_thres : pr.pred_thres,
                      prob = kIsPrey ? pr.prey_prob : pr.pred_prob;
        if (ctx.has_edge(E_EAT, id)) s.energy += gain;
        if (s.energy > thres && ctx.uniform(0) * 100.0 < (double)prob) {
            const int64_t q = s.energy / 2;
            const int64_t off = (s.energy & 1) ? q + (q & 1) : q;
            Animal child{off, {s.pos[0], s.pos[1]}};
            const vb::AgentID nid = ctx.add_agent(T, child);
            move_animal(ctx, nid, s.pos, E_POS, E_VIEW);
            s.energy -= off;
        }
        return true;
    }
};

// map closures of the docs example that are more than a field selector (registered maps: mapreduce / calc_rasterstate)
struct HasFood : vb::MapBase {                 // c -> c.countdown == 0   (predator.jl:487-489, cells with food)
    using Elem = Cell;
    using Result = int64_t;
    VB_HD int64_t operator()(const Cell& c) const { return c.countdown == 0 ? 1 : 0; }
};
struct GrowthProgress : vb::MapBase {          // a Float64-valued read-out for plots: 1 = food, falling to 0 right after grazing
    using Elem = Cell;
    using Result = double;
    VB_HD double operator()(const Cell& c) const { return 1.0 / (1.0 + (double)c.countdown); }
};

}  // namespace pp
