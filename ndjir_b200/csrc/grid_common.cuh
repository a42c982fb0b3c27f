// Cell arithmetic shared by every linear grid family.  The operation order is pinned with explicit
// round-to-nearest intrinsics so the integer cell indices are bit-identical to the reference kernels
// (csrc/grid_feature/voxel_feature_cuda.cu:52-70: subtract, multiply by (G-1)/(max-min), floorf,
// clamp, float->uint truncation).  `scales` is computed on the host with the same IEEE fp32 division the
// reference performs per thread.
#pragma once
#include "common.cuh"

namespace ndjir {

enum Interp { INTERP_LINEAR = 0, INTERP_COSINE = 1 };

struct GridFrame {
  float mnx, mny, mnz;   // min
  float sx, sy, sz;      // (G-1)/(max-min)
  float gx1, gy1, gz1;   // G-1
  int interp;            // INTERP_LINEAR (voxel / triplane / triline) or INTERP_COSINE (the cosine_* families)
};

static inline GridFrame make_frame(int Gx, int Gy, int Gz, const float* mn, const float* mx) {
  GridFrame f;
  f.mnx = mn[0]; f.mny = mn[1]; f.mnz = mn[2];
  f.gx1 = (float)Gx - 1.f; f.gy1 = (float)Gy - 1.f; f.gz1 = (float)Gz - 1.f;
  volatile float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
  volatile float sx = f.gx1 / dx, sy = f.gy1 / dy, sz = f.gz1 / dz;
  f.sx = sx; f.sy = sy; f.sz = sz;
  f.interp = INTERP_LINEAR;
  return f;
}

struct Cell {
  unsigned x0, y0, z0, x1, y1, z1;
  float p0, q0, r0, p1, q1, r1;
  // d(weight of the upper corner)/d(query) per axis: the frame scale for the linear families; for the cosine families
  // scale * 0.5 pi sin(pi frac) (cosine_voxel_feature_cuda.cu:163, :188-196: every derivative is the linear one with
  // cosine weights and this extra per-axis factor)
  float sx, sy, sz;
};

__device__ __forceinline__ void cell_axis(float q, float mn, float s, float g1, unsigned& i0, unsigned& i1,
                                          float& w0, float& w1) {
  float x = __fmul_rn(__fsub_rn(q, mn), s);
  float f0 = floorf(x);
  f0 = fmaxf(f0, 0.f);
  f0 = fminf(f0, g1);
  float f1 = fminf(__fadd_rn(f0, 1.f), g1);
  w0 = __fsub_rn(f1, x);
  w1 = __fsub_rn(1.f, w0);
  i0 = (unsigned)f0;
  i1 = (unsigned)f1;
}

// cosine_voxel_feature_cuda.cu:52-67: same cell, weights 0.5 cos(pi (x - x0)) + 0.5 (fp32 pi, cosf / sinf)
__device__ __forceinline__ void cell_axis_cosine(float q, float mn, float s, float g1, unsigned& i0, unsigned& i1,
                                                 float& w0, float& w1, float& ds) {
  float x = __fmul_rn(__fsub_rn(q, mn), s);
  float f0 = floorf(x);
  f0 = fmaxf(f0, 0.f);
  f0 = fminf(f0, g1);
  float f1 = fminf(__fadd_rn(f0, 1.f), g1);
  float ang = __fmul_rn(3.14159274f, __fsub_rn(x, f0));
  w0 = 0.5f * cosf(ang) + 0.5f;
  w1 = __fsub_rn(1.f, w0);
  ds = s * (1.57079637f * sinf(ang));
  i0 = (unsigned)f0;
  i1 = (unsigned)f1;
}

__device__ __forceinline__ Cell make_cell_linear(const GridFrame& g, float qx, float qy, float qz) {
  Cell c;
  cell_axis(qx, g.mnx, g.sx, g.gx1, c.x0, c.x1, c.p0, c.p1);
  cell_axis(qy, g.mny, g.sy, g.gy1, c.y0, c.y1, c.q0, c.q1);
  cell_axis(qz, g.mnz, g.sz, g.gz1, c.z0, c.z1, c.r0, c.r1);
  c.sx = g.sx; c.sy = g.sy; c.sz = g.sz;
  return c;
}

__device__ __forceinline__ Cell make_cell(const GridFrame& g, float qx, float qy, float qz) {
  Cell c;
  if (g.interp == INTERP_COSINE) {
    cell_axis_cosine(qx, g.mnx, g.sx, g.gx1, c.x0, c.x1, c.p0, c.p1, c.sx);
    cell_axis_cosine(qy, g.mny, g.sy, g.gy1, c.y0, c.y1, c.q0, c.q1, c.sy);
    cell_axis_cosine(qz, g.mnz, g.sz, g.gz1, c.z0, c.z1, c.r0, c.r1, c.sz);
    return c;
  }
  cell_axis(qx, g.mnx, g.sx, g.gx1, c.x0, c.x1, c.p0, c.p1);
  cell_axis(qy, g.mny, g.sy, g.gy1, c.y0, c.y1, c.q0, c.q1);
  cell_axis(qz, g.mnz, g.sz, g.gz1, c.z0, c.z1, c.r0, c.r1);
  c.sx = g.sx; c.sy = g.sy; c.sz = g.sz;
  return c;
}

static __global__ void fill_zero_kernel(float* __restrict__ p, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  long long n4 = n >> 2;
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    float4* p4 = reinterpret_cast<float4*>(p);
    for (long long k = i; k < n4; k += stride) p4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long k = (n4 << 2) + i; k < n; k += stride) p[k] = 0.f;
  } else {
    for (long long k = i; k < n; k += stride) p[k] = 0.f;
  }
}

static inline void fill_zero(float* p, long long n, cudaStream_t st) {
  if (n <= 0) return;
  fill_zero_kernel<<<grid_for((n + 3) / 4), NDJIR_BLOCK, 0, st>>>(p, n);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }

// Channel vector width usable for a table/row of D channels at pointer p.
static inline int pick_vec(int D, const void* a, const void* b = nullptr, const void* c = nullptr) {
  auto ok16 = [&](const void* p) { return p == nullptr || aligned16(p); };
  auto ok8 = [&](const void* p) { return p == nullptr || aligned8(p); };
  if (D % 4 == 0 && ok16(a) && ok16(b) && ok16(c)) return 4;
  if (D % 2 == 0 && ok8(a) && ok8(b) && ok8(c)) return 2;
  return 1;
}

extern int g_scatter_aggregate;  // 0: plain vector reductions, 1: warp-aggregated (set by ndjir_set_option)

}  // namespace ndjir
