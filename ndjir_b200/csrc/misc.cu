// Ray bounds (AABB / sphere), light-direction sampling, squareplus (sm_100a).
//
// Replaces csrc/intersection/ray_aabb_intersection_cuda.cu (:71-142), ray_sphere_intersection_cuda.cu
// (:27-78), csrc/sampling/inverse_transform_cuda.cu (:31-69, :94-136) and csrc/activation/squareplus_cuda.cu
// (:30-60).  All are tiny, latency-bound kernels (36 B in / 12 B out per ray); they matter for parity
// (hit masks must be bit-exact), not for bandwidth.  The arithmetic is pinned with explicit fp32
// intrinsics to what nvcc emits for the reference source (IEEE reciprocal, FADD->FMUL for (bound-o)*inv,
// FFMA for o + t*d; SURVEY.md section 7 "hard parts" item 3).
#include "common.cuh"
#include "bounds.cuh"
#include "../../include/ndjir_b200.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ndjir {
namespace misc {

using bounds::Box;
using bounds::ray_aabb;

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

__global__ void __launch_bounds__(NDJIR_BLOCK)
ray_sphere_kernel(int N, float* __restrict__ t_near, float* __restrict__ t_far, float* __restrict__ n_hits,
                  const float* __restrict__ camloc, const float* __restrict__ raydir, int R, float radius) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    const float* o = camloc + (n / R) * 3;
    const float* d = raydir + (long long)n * 3;
    float ox = __ldg(o), oy = __ldg(o + 1), oz = __ldg(o + 2);
    float dx = __ldg(d), dy = __ldg(d + 1), dz = __ldg(d + 2);
    int nh; float tn, tf;
    bounds::ray_sphere(ox, oy, oz, dx, dy, dz, radius, tn, tf, nh);
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

// ---- random-gather throughput probe (include/ndjir_b200.h: ndjir_bench_gather) ----
namespace ndjir {
namespace probe {
__device__ __forceinline__ unsigned mix(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
template <typename T>
__device__ __forceinline__ float fold(const T& v);
template <> __device__ __forceinline__ float fold<float>(const float& v) { return v; }
template <> __device__ __forceinline__ float fold<float2>(const float2& v) { return v.x + v.y; }
template <> __device__ __forceinline__ float fold<float4>(const float4& v) { return v.x + v.y + v.z + v.w; }

template <typename T, int UNROLL>
__global__ void __launch_bounds__(NDJIR_BLOCK) gather_kernel(long long n_threads, int per_thread, const T* table,
                                                             unsigned n_elems, int coherent, unsigned seed, float* sink) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n_threads) return;
  const unsigned lane = threadIdx.x & 31;
  const unsigned key = coherent ? (unsigned)(tid >> 5) : (unsigned)tid;
  float acc = 0.f;
  for (int g = 0; g < per_thread; g += UNROLL) {
    T v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      unsigned h = mix(key * 0x9e3779b9u + (unsigned)(g + u) * 0x85ebca6bu + seed);
      unsigned idx = coherent ? ((h % (n_elems / 32u)) * 32u + lane) : (h % n_elems);
      v[u] = __ldg(table + idx);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += fold<T>(v[u]);
  }
  if (acc == 1.2345e-30f) sink[0] = acc;     // keeps the loads alive without a store per thread
}
}  // namespace probe
}  // namespace ndjir

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

int ndjir_bench_gather(long long n_threads, int per_thread, int elem_bytes, const void* table, long long table_bytes,
                       int coherent, int seed, float* sink, cudaStream_t stream) {
  using namespace ndjir::probe;
  if (n_threads <= 0 || per_thread <= 0) return NDJIR_OK;
  if (!table || !sink || table_bytes < 32LL * elem_bytes || table_bytes % elem_bytes) return NDJIR_ERR_ARG;
  const long long n_elems = table_bytes / elem_bytes;
  if (n_elems > 0xffffffffLL) return NDJIR_ERR_ARG;
  const int grid = (int)((n_threads + NDJIR_BLOCK - 1) / NDJIR_BLOCK);
  if (elem_bytes == 4)
    gather_kernel<float, 8><<<grid, NDJIR_BLOCK, 0, stream>>>(n_threads, per_thread, (const float*)table, (unsigned)n_elems,
                                                             coherent, (unsigned)seed, sink);
  else if (elem_bytes == 8)
    gather_kernel<float2, 8><<<grid, NDJIR_BLOCK, 0, stream>>>(n_threads, per_thread, (const float2*)table,
                                                              (unsigned)n_elems, coherent, (unsigned)seed, sink);
  else if (elem_bytes == 16)
    gather_kernel<float4, 8><<<grid, NDJIR_BLOCK, 0, stream>>>(n_threads, per_thread, (const float4*)table,
                                                              (unsigned)n_elems, coherent, (unsigned)seed, sink);
  else
    return NDJIR_ERR_ARG;
  NDJIR_RETURN_LAST_ERROR();
}

}  // extern "C"
