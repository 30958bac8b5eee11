// hans_walk.cuh -- the walker kernels (sm_100a), instantiated for two state spaces (see hans_walk_body.inc):
// namespace hb (state in shared memory) and namespace hbg (state in global memory, walkers beyond 227 KB).
#pragma once

#include "hans_device.cuh"

#define HB_NS hb
#include "hans_walk_body.inc"
#undef HB_NS

#define HB_NS hbg
#define HB_STATE_GLOBAL 1
#include "hans_walk_body.inc"
#undef HB_STATE_GLOBAL
#undef HB_NS
