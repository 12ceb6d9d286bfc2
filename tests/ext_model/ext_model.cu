// ext_model.cu — a model library built OUTSIDE the engine, the way a Vahana user ships transitions: include the
// device API, define functors, register them by name, compile for sm_100a into a shared object and hand it to
// vb_load_model_library().  tests/test_ext_model.py compiles this file with nvcc, loads it and (on a GPU) applies it.
//     nvcc -shared -Xcompiler -fPIC -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a \
//          -I include tests/ext_model/ext_model.cu -o libext_model.so -L vahana.jl_b200/csrc/build -lvahana_b200
#include "vahana_device.cuh"

namespace ext {

struct Foo { int64_t foo; };                                  // the 8-byte state of the test models' agent types

// apply!(sim, AMortal, [AMortal, ESLDict1], AMortal) do s, id, sim; AMortal(s.foo + 1 + num_edges(sim, id, ESLDict1)) end
struct AddOnePlusDegree : vb::TransitionBase {
    using State = Foo;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo& s, vb::AgentID id) const {
        s.foo += 1 + (int64_t)ctx.num_edges(1 /* ESLDict1: the second edge type of the core model */, id);
        return true;
    }
};
// mapreduce(sim, a -> a.foo * a.foo, +, AMortal)
struct Square : vb::MapBase {
    using Elem = Foo;
    using Result = int64_t;
    VB_HD int64_t operator()(const Foo& a) const { return a.foo * a.foo; }
};

}  // namespace ext

VB_REGISTER_TRANSITION("ext_add_one_plus_degree", "AMortal", ext::AddOnePlusDegree)
VB_REGISTER_MAP("ext_square", "AMortal", ext::Square)
