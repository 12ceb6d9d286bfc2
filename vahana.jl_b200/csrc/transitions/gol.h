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

// A reduce transition (include/vahana_model.h): the count of active neighbours is an integer fold, exact in any order, so the
// engine may walk the implicit stencil with its grid-stencil kernel instead of enumerating the row in insertion order.
struct Life : vb::ReduceTransition<Life> {
    using State = Cell;
    using Source = Cell;
    struct Acc { int32_t n; };
    static constexpr int kAccBytes = 4;
    static constexpr int kPrimaryEdge = E_NEIGHBOR;
    static constexpr int kSourceType = T_CELL;
    template <class Ctx> VB_HD void init(const Ctx&, const Cell&, Acc& a) const { a.n = 0; }
    template <class Ctx> VB_HD void fold(const Ctx&, const Cell&, const Cell& c, Acc& a) const { a.n += c.active ? 1 : 0; }
    VB_HD void merge(Acc& a, const Acc& b) const { a.n += b.n; }
    template <class Ctx> VB_HD bool finish(const Ctx&, Cell& self, vb::AgentID, const Acc& a) const {
        self.active = (a.n == 3) || (self.active && a.n == 2);
        return true;
    }
};

}  // namespace gol
