// Ray bounds (AABB / sphere), light-direction sampling, squareplus (sm_100a).
//
// Replaces csrc/intersection/ray_aabb_intersection_cuda.cu (:71-142), ray_sphere_intersection_cuda.cu
// (:27-78), csrc/sampling/inverse_transform_cuda.cu (:31-69, :94-136) and csrc/activation/squareplus_cuda.cu
// (:30-60).  All are tiny, latency-bound kernels (36 B in / 12 B out per ray); they matter for parity
// (hit masks must be bit-exact), not for bandwidth.  The arithmetic is pinned with explicit fp32
// intrinsics to what nvcc emits for the reference source (IEEE reciprocal, FADD->FMUL for (bound-o)*inv,
// FFMA for o + t*d; SURVEY.md section 7 "hard parts" item 3).
#include "common.cuh"
#include "../../include/ndjir_b200.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ndjir {
namespace misc {

struct Box { float mnx, mny, mnz, mxx, mxy, mxz; };

__device__ __forceinline__ bool inside(float x, float y, float z, float t, const Box& b) {
  bool cond = (t >= 0.f);
  cond &= (x >= b.mnx) && (x <= b.mxx);
  cond &= (y >= b.mny) && (y <= b.mxy);
  cond &= (z >= b.mnz) && (z <= b.mxz);
  return cond;
}

// Shared by the C-ABI export and the fused sampler.
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
    if (i == 0) x = b.mxx; if (i == 1) y = b.mxy; if (i == 2) z = b.mxz;
    if (i == 3) x = b.mnx; if (i == 4) y = b.mny; if (i == 5) z = b.mnz;
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

__global__ void __launch_bounds__(NDJIR_BLOCK)
ray_aabb_kernel(int N, float* __restrict__ t_near, float* __restrict__ t_far, float* __restrict__ n_hits,
                const float* __restrict__ camloc, const float* __restrict__ raydir, int R, Box b) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    const float* o = camloc + (n / R) * 3;
    const float* d = raydir + (long long)n * 3;
    float tn, tf; int nh;
    ray_aabb(__ldg(o), __ldg(o + 1), __ldg(o + 2), __ldg(d), __ldg(d + 1), __ldg(d + 2), b, tn, tf, nh);
    t_near[n] = tn; t_far[n] = tf; n_hits[n] = (float)nh;
  }
}

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
  // helper_math.h dot(): a.x*b.x + a.y*b.y + a.z*b.z, which nvcc contracts to FMUL, FFMA, FFMA
  return __fmaf_rn(az, bz, __fmaf_rn(ay, by, __fmul_rn(ax, bx)));
}

__global__ void __launch_bounds__(NDJIR_BLOCK)
ray_sphere_kernel(int N, float* __restrict__ t_near, float* __restrict__ t_far, float* __restrict__ n_hits,
                  const float* __restrict__ camloc, const float* __restrict__ raydir, int R, float radius) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    const float* o = camloc + (n / R) * 3;
    const float* d = raydir + (long long)n * 3;
    float ox = __ldg(o), oy = __ldg(o + 1), oz = __ldg(o + 2);
    float dx = __ldg(d), dy = __ldg(d + 1), dz = __ldg(d + 2);
    float r2 = __fmul_rn(radius, radius);
    float cv = dot3(ox, oy, oz, dx, dy, dz);
    float vv = dot3(dx, dy, dz, dx, dy, dz);
    float cc = dot3(ox, oy, oz, ox, oy, oz);
    float X = -cv;
    float Y = __fmaf_rn(cv, cv, -__fmul_rn(vv, __fsub_rn(cc, r2)));  // cv*cv - vv*(cc-r2), contracted
    float Zi = __fdiv_rn(1.f, vv);
    int nh = 0; float tn = 0.f, tf = 0.f;
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
    n_hits[n] = (float)nh; t_near[n] = tn; t_far[n] = tf;
  }
}

// (B*R, M) light directions around the pixel normal; all (theta_i, phi_j) pairs, m_the = m / n_phis.
template <bool IMPORTANCE>
__global__ void __launch_bounds__(NDJIR_BLOCK)
directions_kernel(long long size, float* __restrict__ light_dirs, const float* __restrict__ normal,
                  const float* __restrict__ cdf_the, const float* __restrict__ cdf_phi,
                  const float* __restrict__ alpha, int n_lights, int n_thes, int n_phis, float eps) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < size; s += stride) {
    long long b = s / n_lights;
    int m = (int)(s - b * n_lights);
    int m_the = m / n_phis;
    int m_phi = m - m_the * n_phis;
    float u_the = __ldg(cdf_the + b * n_thes + m_the);
    float u_phi = __ldg(cdf_phi + b * n_phis + m_phi);
    auto phi = 2.f * M_PI * u_phi;  // double product, narrowed by cosf/sinf (:48)
    float cos_the;
    if (IMPORTANCE) {
      float a = __ldg(alpha + b);
      float a2 = a * a;
      cos_the = sqrtf((1.f - u_the) / ((a2 - 1.f) * u_the + 1.f));  // GGX inverse CDF (:116)
    } else {
      cos_the = u_the;
    }
    float sin_the = sqrtf(1.f - cos_the * cos_the);
    float x = sin_the * cosf(phi);
    float y = sin_the * sinf(phi);
    float z = cos_the;
    float nx = __ldg(normal + b * 3) + eps, ny = __ldg(normal + b * 3 + 1) + eps, nz = __ldg(normal + b * 3 + 2) + eps;
    // helper_math.h normalize(): v * rsqrtf(dot(v,v))
    float inz = rsqrtf(nx * nx + ny * ny + nz * nz);
    float zx = nx * inz, zy = ny * inz, zz = nz * inz;
    float xx0 = -ny, xy0 = nx;
    float inx = rsqrtf(xx0 * xx0 + xy0 * xy0 + 0.f * 0.f);
    float xx = xx0 * inx, xy = xy0 * inx, xz = 0.f * inx;
    // y = cross(z, x)
    float yx = zy * xz - zz * xy, yy = zz * xx - zx * xz, yz = zx * xy - zy * xx;
    float ox = x * xx + y * yx + z * zx;
    float oy = x * xy + y * yy + z * zy;
    float oz = x * xz + y * yz + z * zz;
    light_dirs[s * 3] = ox; light_dirs[s * 3 + 1] = oy; light_dirs[s * 3 + 2] = oz;
  }
}

template <bool ACCUM>
__global__ void __launch_bounds__(NDJIR_BLOCK)
squareplus_bwd_kernel(long long n, float* __restrict__ dx, const float* __restrict__ dy, const float* __restrict__ x,
                      float b) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += stride) {
    float xv = x[s];
    float g = dy[s] * 0.5f * (1.f + xv * rsqrtf(xv * xv + b));
    dx[s] = ACCUM ? dx[s] + g : g;
  }
}

__global__ void __launch_bounds__(NDJIR_BLOCK)
squareplus_fwd_kernel(long long n, float* __restrict__ y, const float* __restrict__ x, float b) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += stride) {
    float xv = x[s];
    y[s] = 0.5f * (xv + sqrtf(xv * xv + b));
  }
}

}  // namespace misc
}  // namespace ndjir

using namespace ndjir;
using namespace ndjir::misc;

extern "C" {

int ndjir_ray_aabb_intersection(int n_rays, float* t_near, float* t_far, float* n_hits, const float* camloc,
                                const float* raydir, int B, int R, const float* min3, const float* max3,
                                cudaStream_t stream) {
  if (n_rays == 0) return NDJIR_OK;
  if (n_rays < 0 || R <= 0 || n_rays != B * R || !t_near || !t_far || !n_hits || !camloc || !raydir || !min3 || !max3)
    return NDJIR_ERR_ARG;
  Box b = {min3[0], min3[1], min3[2], max3[0], max3[1], max3[2]};
  ray_aabb_kernel<<<grid_for(n_rays), NDJIR_BLOCK, 0, stream>>>(n_rays, t_near, t_far, n_hits, camloc, raydir, R, b);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_ray_sphere_intersection(int n_rays, float* t_near, float* t_far, float* n_hits, const float* camloc,
                                  const float* raydir, int B, int R, float radius, cudaStream_t stream) {
  if (n_rays == 0) return NDJIR_OK;
  if (n_rays < 0 || R <= 0 || n_rays != B * R || !t_near || !t_far || !n_hits || !camloc || !raydir)
    return NDJIR_ERR_ARG;
  ray_sphere_kernel<<<grid_for(n_rays), NDJIR_BLOCK, 0, stream>>>(n_rays, t_near, t_far, n_hits, camloc, raydir, R,
                                                                  radius);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_sample_uniform_directions(long long size, float* light_dirs, const float* normal, const float* cdf_the,
                                    const float* cdf_phi, int batch_size, int n_lights, int n_thes, int n_phis,
                                    float eps, cudaStream_t stream) {
  if (size == 0) return NDJIR_OK;
  if (size < 0 || n_lights != n_thes * n_phis || size != (long long)batch_size * n_lights || !light_dirs ||
      !normal || !cdf_the || !cdf_phi)
    return NDJIR_ERR_ARG;
  directions_kernel<false><<<grid_for(size), NDJIR_BLOCK, 0, stream>>>(size, light_dirs, normal, cdf_the, cdf_phi,
                                                                       nullptr, n_lights, n_thes, n_phis, eps);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_sample_importance_directions(long long size, float* light_dirs, const float* normal,
                                       const float* cdf_the, const float* cdf_phi, const float* alpha,
                                       int batch_size, int n_lights, int n_thes, int n_phis, float eps,
                                       cudaStream_t stream) {
  if (size == 0) return NDJIR_OK;
  if (size < 0 || n_lights != n_thes * n_phis || size != (long long)batch_size * n_lights || !light_dirs ||
      !normal || !cdf_the || !cdf_phi || !alpha)
    return NDJIR_ERR_ARG;
  directions_kernel<true><<<grid_for(size), NDJIR_BLOCK, 0, stream>>>(size, light_dirs, normal, cdf_the, cdf_phi,
                                                                      alpha, n_lights, n_thes, n_phis, eps);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_squareplus_forward(long long size, float* output, const float* input, float b, cudaStream_t stream) {
  if (size == 0) return NDJIR_OK;
  if (size < 0 || !output || !input) return NDJIR_ERR_ARG;
  squareplus_fwd_kernel<<<grid_for(size), NDJIR_BLOCK, 0, stream>>>(size, output, input, b);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_squareplus_backward(long long size, float* dinput, const float* doutput, const float* input, float b,
                              int accum, cudaStream_t stream) {
  if (size == 0) return NDJIR_OK;
  if (size < 0 || !dinput || !doutput || !input) return NDJIR_ERR_ARG;
  if (accum) squareplus_bwd_kernel<true><<<grid_for(size), NDJIR_BLOCK, 0, stream>>>(size, dinput, doutput, input, b);
  else squareplus_bwd_kernel<false><<<grid_for(size), NDJIR_BLOCK, 0, stream>>>(size, dinput, doutput, input, b);
  NDJIR_RETURN_LAST_ERROR();
}

}  // extern "C"
