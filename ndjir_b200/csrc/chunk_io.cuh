// (B, D*3) rows with channel index c = d*3 + plane (reference common.cuh:29-35): vector access to the 3*V contiguous
// values of a channel chunk.  Shared by triplaneline.cu and lanczos_triplaneline.cu.
#pragma once
#include "common.cuh"

namespace ndjir {
namespace tpl {

// Write / read the 3*V contiguous values [d*3 .. (d+V)*3) of a (B, D*3) row as three Vec<V>.
template <int V>
__device__ __forceinline__ void store_chunk(float* row, int d, const float (&o)[3][V], bool accum) {
  float flat[3 * V];
#pragma unroll
  for (int j = 0; j < V; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i) flat[j * 3 + i] = o[i][j];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    Vec<V> t;
    float* dst = row + d * 3 + k * V;
    if (accum) {
      Vec<V> prev = ld_vec<V>(dst);
#pragma unroll
      for (int j = 0; j < V; ++j) t.v[j] = prev.v[j] + flat[k * V + j];
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) t.v[j] = flat[k * V + j];
    }
    st_vec<V>(dst, t);
  }
}

template <int V>
__device__ __forceinline__ void load_chunk(const float* row, int d, float (&o)[3][V]) {
  float flat[3 * V];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    Vec<V> t = ldg_vec<V>(row + d * 3 + k * V);
#pragma unroll
    for (int j = 0; j < V; ++j) flat[k * V + j] = t.v[j];
  }
#pragma unroll
  for (int j = 0; j < V; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i][j] = flat[j * 3 + i];
}

}  // namespace tpl
}  // namespace ndjir
