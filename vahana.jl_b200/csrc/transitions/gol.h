// gol.h — Conway's Game of Life on a raster (BASELINE config 2).  The reference only sketches it
// (/root/reference/src/Raster.jl:193-200, docs/src/parallel.md:140-144); SURVEY.md Appendix C gives the model
// in the reference's API:
//     struct Cell active::Bool end ; struct Neighbor end
//     register_agenttype!(Cell, :Immortal) ; register_edgetype!(Neighbor, :Stateless, :SingleType; target = Cell)
//     life(c, id, sim) = (n = count(s -> s.active, neighborstates_iter(sim, id, Neighbor, Cell));
//                         Cell(n == 3 || (c.active && n == 2)))
// All-integer: parity with the oracle is bit-exact.
#pragma once
#include "../../../include/vahana_model.h"

namespace gol {

struct Cell { bool active; };
enum : int { T_CELL = 1 };
enum : int { E_NEIGHBOR = 0 };

struct Life : vb::TransitionBase {
    using State = Cell;
    static constexpr bool kCooperative = true;
    static constexpr int kPrimaryEdge = E_NEIGHBOR;
    template <class Ctx>
    VB_HD bool operator()(Ctx& ctx, Cell& self, vb::AgentID id) const {
        int n = 0;
        ctx.template for_each_neighborstate<Cell>(E_NEIGHBOR, T_CELL, id, [&](const Cell& c) { n += c.active ? 1 : 0; });
        n = ctx.sum(n);
        self.active = (n == 3) || (self.active && n == 2);
        return true;
    }
};

}  // namespace gol
