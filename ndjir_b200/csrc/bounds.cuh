// Ray-AABB / ray-sphere bounds as device functions, shared by the C-ABI exports (misc.cu) and the fused
// sampler (render.cu).  Arithmetic pinned with explicit fp32 intrinsics to what nvcc emits for the reference
// source (csrc/intersection/ray_aabb_intersection_cuda.cu:27-142, ray_sphere_intersection_cuda.cu:27-78):
// IEEE reciprocal, FADD->FMUL for (bound-o)*inv, FFMA for o + t*d.
#pragma once
#include "common.cuh"

namespace ndjir {
namespace bounds {

struct Box { float mnx, mny, mnz, mxx, mxy, mxz; };

__device__ __forceinline__ bool inside(float x, float y, float z, float t, const Box& b) {
  bool cond = (t >= 0.f);
  cond &= (x >= b.mnx) && (x <= b.mxx);
  cond &= (y >= b.mny) && (y <= b.mxy);
  cond &= (z >= b.mnz) && (z <= b.mxz);
  return cond;
}

__device__ __forceinline__ void ray_aabb(float ox, float oy, float oz, float dx, float dy, float dz, const Box& b,
                                         float& t_near, float& t_far, int& n_hits) {
  float ix = __fdiv_rn(1.f, dx), iy = __fdiv_rn(1.f, dy), iz = __fdiv_rn(1.f, dz);
  float t[6];
  t[0] = __fmul_rn(__fsub_rn(b.mxx, ox), ix);
  t[1] = __fmul_rn(__fsub_rn(b.mxy, oy), iy);
  t[2] = __fmul_rn(__fsub_rn(b.mxz, oz), iz);
  t[3] = __fmul_rn(__fsub_rn(b.mnx, ox), ix);
  t[4] = __fmul_rn(__fsub_rn(b.mny, oy), iy);
  t[5] = __fmul_rn(__fsub_rn(b.mnz, oz), iz);
  n_hits = 0;
  int first = 0, last = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    float ti = t[i];
    float x = __fmaf_rn(ti, dx, ox), y = __fmaf_rn(ti, dy, oy), z = __fmaf_rn(ti, dz, oz);
    // snap the tested axis onto the plane (:60-66)
    if (i == 0) x = b.mxx;
    if (i == 1) y = b.mxy;
    if (i == 2) z = b.mxz;
    if (i == 3) x = b.mnx;
    if (i == 4) y = b.mny;
    if (i == 5) z = b.mnz;
    if (isinf(ti)) continue;
    if (!inside(x, y, z, ti, b)) continue;
    if (n_hits == 0) first = i; else last = i;
    n_hits++;
  }
  t_near = 0.f; t_far = 0.f;
  if (n_hits >= 2) {
    float a = t[0], c = t[0];
#pragma unroll
    for (int i = 0; i < 6; ++i) { if (i == first) a = t[i]; if (i == last) c = t[i]; }
    if (a <= c) { t_near = a; t_far = c; } else { t_near = c; t_far = a; }
  } else if (n_hits == 1) {
    float a = t[0];
#pragma unroll
    for (int i = 0; i < 6; ++i) if (i == first) a = t[i];
    t_far = a;
  }
}

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
  // helper_math.h dot(): a.x*b.x + a.y*b.y + a.z*b.z.  nvcc fuses the FIRST product of each sum into the FMA and
  // leaves the second as the FMUL: fma(a.z, b.z, fma(a.x, b.x, a.y*b.y))
  return __fmaf_rn(az, bz, __fmaf_rn(ax, bx, __fmul_rn(ay, by)));
}

__device__ __forceinline__ void ray_sphere(float ox, float oy, float oz, float dx, float dy, float dz, float radius,
                                           float& tn, float& tf, int& nh) {
  float r2 = __fmul_rn(radius, radius);   // the reference build computes cc - r*r as fma(-r, r, cc)
  float cv = dot3(ox, oy, oz, dx, dy, dz);
  float vv = dot3(dx, dy, dz, dx, dy, dz);
  float cc = dot3(ox, oy, oz, ox, oy, oz);
  float X = -cv;
  (void)r2;
  float Y = __fmaf_rn(cv, cv, -__fmul_rn(vv, __fmaf_rn(-radius, radius, cc)));  // cv*cv - vv*(cc-r2), contracted
  float Zi = __fdiv_rn(1.f, vv);
  nh = 0; tn = 0.f; tf = 0.f;
  if (Y > 0) {
    float Ys = sqrtf(Y);
    tn = __fmul_rn(__fsub_rn(X, Ys), Zi);
    tf = __fmul_rn(__fadd_rn(X, Ys), Zi);
    int pos = int(tn >= 0);
    tn = pos * tn;
    nh = 2 - (1 - pos);
  } else if (Y == 0) {
    nh = 1; tn = __fmul_rn(X, Zi); tf = tn;
  }
}

}  // namespace bounds
}  // namespace ndjir
