// Shared device/host helpers for the ndjir_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define NDJIR_OK 0
#define NDJIR_ERR_ARG -1

#define NDJIR_NUM_SMS 148
#define NDJIR_BLOCK 256

// Every C-ABI entry point returns 0 or the CUDA error code of its last launch (the reference only
// printf's launch errors, csrc/cuda_common.cuh:24-32; we hand them back so the shim can raise).
#define NDJIR_RETURN_LAST_ERROR()                     \
  do {                                                \
    cudaError_t e__ = cudaGetLastError();             \
    return e__ == cudaSuccess ? NDJIR_OK : (int)e__;  \
  } while (0)

namespace ndjir {

// Grid sizing: a multiple of the SM count when the work is large (many short waves of 256-thread CTAs, so the
// tail of the last wave is negligible whatever the per-kernel occupancy is), grid-stride inside.
static inline int grid_for(long long work_items, int block = NDJIR_BLOCK, int ctas_per_sm = 64) {
  long long need = (work_items + block - 1) / block;
  long long cap = (long long)NDJIR_NUM_SMS * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

template <int V> struct Vec { float v[V]; };

template <int V> __device__ __forceinline__ Vec<V> ldg_vec(const float* p);
template <> __device__ __forceinline__ Vec<1> ldg_vec<1>(const float* p) {
  Vec<1> r; r.v[0] = __ldg(p); return r;
}
template <> __device__ __forceinline__ Vec<2> ldg_vec<2>(const float* p) {
  float2 t = __ldg(reinterpret_cast<const float2*>(p));
  Vec<2> r; r.v[0] = t.x; r.v[1] = t.y; return r;
}
template <> __device__ __forceinline__ Vec<4> ldg_vec<4>(const float* p) {
  float4 t = __ldg(reinterpret_cast<const float4*>(p));
  Vec<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}

template <int V> __device__ __forceinline__ Vec<V> ld_vec(const float* p);
template <> __device__ __forceinline__ Vec<1> ld_vec<1>(const float* p) { Vec<1> r; r.v[0] = *p; return r; }
template <> __device__ __forceinline__ Vec<2> ld_vec<2>(const float* p) {
  float2 t = *reinterpret_cast<const float2*>(p);
  Vec<2> r; r.v[0] = t.x; r.v[1] = t.y; return r;
}
template <> __device__ __forceinline__ Vec<4> ld_vec<4>(const float* p) {
  float4 t = *reinterpret_cast<const float4*>(p);
  Vec<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}

template <int V> __device__ __forceinline__ void st_vec(float* p, const Vec<V>& a);
template <> __device__ __forceinline__ void st_vec<1>(float* p, const Vec<1>& a) { *p = a.v[0]; }
template <> __device__ __forceinline__ void st_vec<2>(float* p, const Vec<2>& a) {
  *reinterpret_cast<float2*>(p) = make_float2(a.v[0], a.v[1]);
}
template <> __device__ __forceinline__ void st_vec<4>(float* p, const Vec<4>& a) {
  *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}

// Vector reductions to global memory (REDG.E.ADD.F32x{2,4}): one L2 atomic transaction per cell
// instead of D scalar atomics (reference: voxel_feature_cuda.cu:277-287 issues 8*D scalar atomicAdd).
template <int V> __device__ __forceinline__ void red_vec(float* p, const Vec<V>& a);
template <> __device__ __forceinline__ void red_vec<1>(float* p, const Vec<1>& a) { atomicAdd(p, a.v[0]); }
template <> __device__ __forceinline__ void red_vec<2>(float* p, const Vec<2>& a) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a.v[0]), "f"(a.v[1]) : "memory");
}
template <> __device__ __forceinline__ void red_vec<4>(float* p, const Vec<4>& a) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a.v[0]), "f"(a.v[1]), "f"(a.v[2]),
               "f"(a.v[3])
               : "memory");
}

// Warp-aggregated scatter: lanes that target the same cell are summed with shuffles and one lane issues
// the reduction.  `key` identifies the cell, `active` says whether this lane has a contribution.  All 32
// lanes of the warp must call this (converged).
template <int V>
__device__ __forceinline__ void warp_agg_red(float* base, unsigned long long key, Vec<V> val, bool active) {
  unsigned mask_active = __ballot_sync(0xffffffffu, active);
  if (!active) key = ~0ull - (threadIdx.x & 31);  // unique keys for idle lanes
  unsigned peers = __match_any_sync(0xffffffffu, key);
  peers &= mask_active;
  int lane = threadIdx.x & 31;
  int leader = __ffs(peers) - 1;
  // Uniform trip count = largest peer group in the warp (1 when nobody shares a cell: no shuffles).
  unsigned maxpop = __reduce_max_sync(0xffffffffu, (unsigned)__popc(peers));
  if (maxpop > 1) {
    // Each lane walks its own peer set in ascending lane order (deterministic summation order).
    Vec<V> acc;
#pragma unroll
    for (int c = 0; c < V; ++c) acc.v[c] = 0.f;
    unsigned rem = peers;
    for (unsigned k = 0; k < maxpop; ++k) {
      bool take = rem != 0;
      int src = take ? (__ffs(rem) - 1) : lane;
      rem &= rem - 1;
#pragma unroll
      for (int c = 0; c < V; ++c) {
        float other = __shfl_sync(0xffffffffu, val.v[c], src);
        if (take) acc.v[c] += other;
      }
    }
    val = acc;
  }
  if (active && lane == leader) red_vec<V>(base, val);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ndjir
