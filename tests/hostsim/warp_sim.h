// Test scaffolding: single-lane stand-ins for the CUDA warp intrinsics so that the decoder's
// device functions (written against a warp of GQW_LANES lanes) compile with g++ and run with one
// lane that owns every slot.  Not part of the product.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <algorithm>
#define GQ_HOST_SIM 1
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
using std::min;
using std::max;
static inline int __shfl_sync(unsigned, int v, int) { return v; }
static inline int __shfl_up_sync(unsigned, int v, int) { return v; }
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline int __any_sync(unsigned, bool p) { return p ? 1 : 0; }
static inline unsigned __reduce_min_sync(unsigned, unsigned v) { return v; }
static inline int __reduce_add_sync(unsigned, int v) { return v; }
static inline unsigned __reduce_xor_sync(unsigned, unsigned v) { return v; }
static inline void __syncwarp() {}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
struct uint4 { unsigned x, y, z, w; };
