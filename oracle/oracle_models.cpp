// oracle_models.cpp — CPU ORACLE (test infrastructure): instantiates the single-source
// transition functors of vahana.jl_b200/csrc/transitions with the sequential oracle context.
#include <cmath>
#include "vahana_oracle.hpp"
#include "../vahana.jl_b200/csrc/transitions/all.h"

#define VB_TRANSITION(tname, atype, ...) VO_REGISTER_TRANSITION(tname, atype, __VA_ARGS__)
#include "../vahana.jl_b200/csrc/transitions/registry.inc"
#undef VB_TRANSITION
