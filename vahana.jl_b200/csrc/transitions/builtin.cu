// builtin.cu — compiles the built-in single-source transitions (transitions/*.h) into CUDA launchers and registers them with
// the engine under (name, agent type).  Model authors build the same way: include vahana_device.cuh, define functors,
// VB_REGISTER_TRANSITION(...), link or vb_load_model_library() the result.  (Part 0 of registry.inc: the models and test/core.jl.)
#include "../../../include/vahana_device.cuh"
#include "all.h"

#define VB_PART 0
#define VB_TRANSITION(tname, atype, ...) VB_REGISTER_TRANSITION(tname, atype, __VA_ARGS__)
#define VB_MAP(mname, tname, ...) VB_REGISTER_MAP(mname, tname, __VA_ARGS__)
#include "registry.inc"
#undef VB_MAP
#undef VB_TRANSITION
