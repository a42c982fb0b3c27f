// Device-side pieces of the full-frame inference driver: the reference renders an image in
// 480 chunks of 4000 rays, generating every chunk's rays with numpy on the host and copying the colours back per chunk
// (python/renderer.py:212-272, python/helper.py:44-81).  Here the whole chunk - ray generation, the random inputs,
// the forward path, the store into the image - is device work driven by a DEVICE chunk counter, so one captured CUDA
// graph is replayed per chunk with no host round trip.
//
//   ndjir_generate_rays     pixel -> world ray, the arithmetic of helper.generate_raydir_camloc in float64 like numpy:
//                           raydir = normalize(R_c2w (K^-1 [x, y, 1]))
//   ndjir_uniform           counter-based uniform numbers (what F.rand supplies in the reference; nnabla's generator
//                           stream cannot be reproduced, the distribution is what matters for stratified sampling)
//   ndjir_store_chunk       image[chunk * n + i] = clip(colour[i], 0, 1)            (renderer.py:266-268)
//   ndjir_counter_add       chunk counter += step (the rank stride of a multi-GPU render)
#include "common.cuh"
#include "../../include/ndjir_b200.h"

namespace ndjir {
namespace inference {

__global__ void __launch_bounds__(NDJIR_BLOCK)
generate_rays_kernel(int n, int W, long long n_pixels, long long pixel0, const int* __restrict__ chunk,
                     const double* __restrict__ kinv, const double* __restrict__ rot, float* __restrict__ raydir) {
  const long long base = pixel0 + (chunk ? (long long)__ldg(chunk) * n : 0ll);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    long long p = base + i;
    if (p >= n_pixels) p = n_pixels - 1;        // the tail of the last chunk re-renders the last pixel (never stored)
    const double x = (double)(p % W), y = (double)(p / W);
    double c[3], w[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) c[r] = kinv[3 * r] * x + kinv[3 * r + 1] * y + kinv[3 * r + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) w[r] = rot[3 * r] * c[0] + rot[3 * r + 1] * c[1] + rot[3 * r + 2] * c[2];
    const double inv = 1.0 / sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    raydir[3 * i] = (float)(w[0] * inv);
    raydir[3 * i + 1] = (float)(w[1] * inv);
    raydir[3 * i + 2] = (float)(w[2] * inv);
  }
}

// three rounds of a 64-bit mix (splitmix64 finaliser) over (index, seed, counter): independent streams per call site
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(NDJIR_BLOCK)
uniform_kernel(long long n, float lo, float hi, unsigned long long seed, const int* __restrict__ counter,
               float* __restrict__ out) {
  const unsigned long long c = counter ? (unsigned long long)__ldg(counter) : 0ull;
  const unsigned long long key = mix64(seed * 0x9e3779b97f4a7c15ull + c + 0x632be59bd9b4e019ull);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long r = mix64(key + 0x9e3779b97f4a7c15ull * (unsigned long long)(i + 1));
    const float u = (float)(r >> 40) * (1.0f / 16777216.0f);          // 24 random bits -> [0, 1)
    out[i] = lo + (hi - lo) * u;
  }
}

__global__ void __launch_bounds__(NDJIR_BLOCK)
store_chunk_kernel(int n, long long n_pixels, long long pixel0, const int* __restrict__ chunk,
                   const float* __restrict__ color, float* __restrict__ image) {
  const long long base = pixel0 + (chunk ? (long long)__ldg(chunk) * n : 0ll);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const long long p = base + i;
    if (p < n_pixels) {
#pragma unroll
      for (int c = 0; c < 3; ++c) image[3 * p + c] = fminf(fmaxf(color[3 * i + c], 0.f), 1.f);
    }
  }
}

__global__ void counter_add_kernel(int* counter, int step) { *counter += step; }

// One training batch assembled on the device (python/dataset.py:33-55 IDRDataSource._get_data, default branch:
// colour = image[pixel_idx], mask = mask[pixel_idx], xy = xy[pixel_idx]; python/train.py:126-130 turns xy into rays with
// helper.generate_raydir_camloc).  One thread per ray; the pixel is given (the reference's host-drawn indices) or drawn
// here from the counter-based generator.
__global__ void __launch_bounds__(NDJIR_BLOCK)
train_batch_kernel(int B, int R, int W, long long n_pixels, const int* __restrict__ view_ids,
                   const int* __restrict__ pixel_idx, unsigned long long seed, const int* __restrict__ counter,
                   const float* __restrict__ images, const float* __restrict__ masks, const double* __restrict__ kinv9,
                   const double* __restrict__ rot9, const float* __restrict__ camloc_all, float* __restrict__ raydir,
                   float* __restrict__ camloc, float* __restrict__ color_gt, float* __restrict__ obj_mask,
                   int* __restrict__ pixel_out) {
  const unsigned long long c = counter ? (unsigned long long)__ldg(counter) : 0ull;
  const unsigned long long key = mix64(seed * 0x9e3779b97f4a7c15ull + c + 0x2545f4914f6cdd1dull);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * R; i += gridDim.x * blockDim.x) {
    const int b = i / R, v = __ldg(view_ids + b);
    long long p;
    if (pixel_idx) p = __ldg(pixel_idx + i);
    else p = (long long)(mix64(key + 0x9e3779b97f4a7c15ull * (unsigned long long)(i + 1)) % (unsigned long long)n_pixels);
    if (pixel_out) pixel_out[i] = (int)p;
    const double x = (double)(p % W), y = (double)(p / W);
    const double* kinv = kinv9 + 9 * v;
    const double* rot = rot9 + 9 * v;
    double cam[3], w[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) cam[r] = kinv[3 * r] * x + kinv[3 * r + 1] * y + kinv[3 * r + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) w[r] = rot[3 * r] * cam[0] + rot[3 * r + 1] * cam[1] + rot[3 * r + 2] * cam[2];
    const double inv = 1.0 / sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const float* px = images + ((long long)v * n_pixels + p) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      raydir[3 * i + k] = (float)(w[k] * inv);
      color_gt[3 * i + k] = __ldg(px + k);
    }
    if (obj_mask) obj_mask[i] = masks ? __ldg(masks + (long long)v * n_pixels + p) : 1.f;
    if (i % R == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) camloc[3 * b + k] = __ldg(camloc_all + 3 * v + k);
    }
  }
}

// lattice of extract_by_mc.compute_pts_vol (python/extract_by_mc.py:47-73): np.linspace(-r, r, G) on every axis, x the
// slowest axis; point i of the batch lies on x-plane ix0 + (i / G^2) * ix_stride
__global__ void __launch_bounds__(NDJIR_BLOCK)
lattice_kernel(long long n, int G, int ix0, int ix_stride, float radius, float* __restrict__ pts) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const double step = G > 1 ? 2.0 * (double)radius / (double)(G - 1) : 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long plane = i / ((long long)G * G), rem = i - plane * (long long)G * G;
    const int ix = ix0 + (int)plane * ix_stride, iy = (int)(rem / G), iz = (int)(rem % G);
    // numpy.linspace: start + k * step, the last sample pinned to `stop`
    pts[3 * i] = ix == G - 1 ? radius : (float)(-(double)radius + ix * step);
    pts[3 * i + 1] = iy == G - 1 ? radius : (float)(-(double)radius + iy * step);
    pts[3 * i + 2] = iz == G - 1 ? radius : (float)(-(double)radius + iz * step);
  }
}

}  // namespace inference
}  // namespace ndjir

using namespace ndjir;

extern "C" {

int ndjir_generate_rays(int n_rays, int W, long long n_pixels, long long pixel0, const int* chunk_dev,
                        const double* kinv9_dev, const double* rot9_dev, float* raydir, cudaStream_t stream) {
  if (n_rays == 0) return NDJIR_OK;
  if (n_rays < 0 || W <= 0 || n_pixels <= 0 || !kinv9_dev || !rot9_dev || !raydir) return NDJIR_ERR_ARG;
  inference::generate_rays_kernel<<<grid_for(n_rays), NDJIR_BLOCK, 0, stream>>>(n_rays, W, n_pixels, pixel0, chunk_dev,
                                                                               kinv9_dev, rot9_dev, raydir);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_train_batch(int B, int R, int W, long long n_pixels, const int* view_ids, const int* pixel_idx, long long seed,
                      const int* counter_dev, const float* images, const float* masks, const double* kinv9,
                      const double* rot9, const float* camloc_all, float* raydir, float* camloc, float* color_gt,
                      float* obj_mask, int* pixel_out, cudaStream_t stream) {
  if (B == 0 || R == 0) return NDJIR_OK;
  if (B < 0 || R < 0 || W <= 0 || n_pixels <= 0 || n_pixels > 0x7fffffffll || !view_ids || !images || !kinv9 || !rot9 ||
      !camloc_all || !raydir || !camloc || !color_gt)
    return NDJIR_ERR_ARG;
  inference::train_batch_kernel<<<grid_for((long long)B * R), NDJIR_BLOCK, 0, stream>>>(
      B, R, W, n_pixels, view_ids, pixel_idx, (unsigned long long)seed, counter_dev, images, masks, kinv9, rot9, camloc_all,
      raydir, camloc, color_gt, obj_mask, pixel_out);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_uniform(long long n, float lo, float hi, long long seed, const int* counter_dev, float* out,
                  cudaStream_t stream) {
  if (n == 0) return NDJIR_OK;
  if (n < 0 || !out) return NDJIR_ERR_ARG;
  inference::uniform_kernel<<<grid_for(n), NDJIR_BLOCK, 0, stream>>>(n, lo, hi, (unsigned long long)seed, counter_dev, out);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_store_chunk(int n_rays, long long n_pixels, long long pixel0, const int* chunk_dev, const float* color,
                      float* image, cudaStream_t stream) {
  if (n_rays == 0) return NDJIR_OK;
  if (n_rays < 0 || !color || !image) return NDJIR_ERR_ARG;
  inference::store_chunk_kernel<<<grid_for(n_rays), NDJIR_BLOCK, 0, stream>>>(n_rays, n_pixels, pixel0, chunk_dev, color,
                                                                             image);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_lattice_points(long long n, int G, int ix0, int ix_stride, float radius, float* pts, cudaStream_t stream) {
  if (n == 0) return NDJIR_OK;
  if (n < 0 || G <= 0 || !pts) return NDJIR_ERR_ARG;
  inference::lattice_kernel<<<grid_for(n), NDJIR_BLOCK, 0, stream>>>(n, G, ix0, ix_stride, radius, pts);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_counter_add(int* counter_dev, int step, cudaStream_t stream) {
  if (!counter_dev) return NDJIR_ERR_ARG;
  inference::counter_add_kernel<<<1, 1, 0, stream>>>(counter_dev, step);
  NDJIR_RETURN_LAST_ERROR();
}

}  // extern "C"
