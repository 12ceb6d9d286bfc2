// builtin_tests.cu — part 1 of registry.inc: the closures of the reference's edge / lifecycle / raster tests (testkit.h).
#include "../../../include/vahana_device.cuh"
#include "all.h"

#define VB_PART 1
#define VB_TRANSITION(tname, atype, ...) VB_REGISTER_TRANSITION(tname, atype, __VA_ARGS__)
#include "registry.inc"
#undef VB_TRANSITION
