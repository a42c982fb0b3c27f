// Lanczos-2 window functions and per-axis taps shared by the Lanczos grid families (lanczos_voxel.cu,
// lanczos_triplaneline.cu): csrc/grid_feature/common.cuh:54-97 of the reference; tap coordinates clamp(x0 + i, 0, G-1)
// with x0 = floor(x) NOT clamped (lanczos_voxel_feature_cuda.cu:61-76, lanczos_triplane_feature_cuda.cu:68-75).
#pragma once
#include "grid_common.cuh"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ndjir {
namespace lanczos {

constexpr int W = 2;       // window size a
constexpr int K = 2 * W;   // taps per axis

__device__ __forceinline__ float sinc(float x) {  // common.cuh:54-59
  if (x == 0.f) return 1.0f;
  return sinf(x) / x;
}
__device__ __forceinline__ float lanczos_w(float x, int a) {  // common.cuh:62-69
  auto z = M_PI * x;
  auto u = sinc(z);
  auto v = sinc(z / a);
  return u * v;
}
__device__ __forceinline__ float grad_coefficient(float x, int a) {  // common.cuh:82-97
  if (x == 0.f) return 0.0f;
  auto z0 = M_PI * x;
  auto z1 = M_PI * x / a;
  auto sinc_z0 = sinc(z0);
  auto sinc_z1 = sinc(z1);
  auto t0 = (cosf(z0) - sinc_z0) * sinc_z1;
  auto t1 = (cosf(z1) - sinc_z1) * sinc_z0;
  return (t0 + t1) / x;
}

struct Taps {
  unsigned ix[K], iy[K], iz[K];
  float cx[K], cy[K], cz[K];
  float gx[K], gy[K], gz[K];
};

template <bool DERIV>
__device__ __forceinline__ void axis_taps(float q, float mn, float s, float g1, unsigned (&idx)[K], float (&c)[K],
                                          float (&gc)[K]) {
  float x = __fmul_rn(__fsub_rn(q, mn), s);
  float x0 = floorf(x);
#pragma unroll
  for (int t = 0; t < K; ++t) {
    float xi = fminf(fmaxf(x0 + (float)(t - W + 1), 0.f), g1);  // clamp(x0 + i, 0, G-1)
    float dx = __fsub_rn(x, xi);
    c[t] = lanczos_w(dx, W);
    if (DERIV) gc[t] = grad_coefficient(dx, W);
    idx[t] = (unsigned)xi;
  }
}

template <bool DERIV>
__device__ __forceinline__ Taps make_taps(const GridFrame& g, const float* q) {
  Taps t;
  axis_taps<DERIV>(__ldg(q), g.mnx, g.sx, g.gx1, t.ix, t.cx, t.gx);
  axis_taps<DERIV>(__ldg(q + 1), g.mny, g.sy, g.gy1, t.iy, t.cy, t.gy);
  axis_taps<DERIV>(__ldg(q + 2), g.mnz, g.sz, g.gz1, t.iz, t.cz, t.gz);
  return t;
}

}  // namespace lanczos
}  // namespace ndjir
