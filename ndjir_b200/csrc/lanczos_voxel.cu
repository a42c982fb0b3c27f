// Separable Lanczos-2 (4x4x4 taps) voxel feature query + backward (sm_100a).
//
// Replaces csrc/grid_feature/lanczos_voxel_feature_cuda.cu (5 exports, :822-834); window functions from
// csrc/grid_feature/common.cuh:54-97.  xyz0 = floor(xyz) is NOT clamped, the tap coordinates are
// (duplicated border taps, :61-76); M_PI*x is a double product narrowed at sinc(float) (q6).
// Thread mapping: one thread per point.  The 12 window weights (and 12 derivative weights) are evaluated
// ONCE per point - the reference re-evaluates them per channel and inside the j/k loops (192 sinf per
// thread) - then the 64 taps are gathered with one 16/8/4-byte load per tap for all channels.
// Roofline: L2-gather-bound at bench size (G=256,D=4: 268 MB table): 1 052 B/pt gathered, 44 B/pt
// compulsory HBM at B=2^24.
#include "grid_common.cuh"
#include "../../include/ndjir_b200.h"
#include "lanczos_common.cuh"
#include "voxel_binned.cuh"

namespace ndjir {
namespace lanczos {

struct Strides { unsigned sx, sy, sz; };

enum Mode { FWD = 0, GRAD_QUERY = 1, GGO = 2 };

template <int MODE, int V, bool ACCUM>
__global__ void __launch_bounds__(NDJIR_BLOCK)
gather_kernel(long long B, float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ gg,
              const float* __restrict__ query, const float* __restrict__ feat, GridFrame g, Strides s, int D) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < B; p += stride) {
    Taps t = make_taps<MODE != FWD>(g, query + p * 3);
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
    if (MODE == GGO) { ggx = __ldg(gg + p * 3); ggy = __ldg(gg + p * 3 + 1); ggz = __ldg(gg + p * 3 + 2); }
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int d = 0; d < D; d += V) {
      float f[V], gx[V], gy[V], gz[V];
#pragma unroll
      for (int j = 0; j < V; ++j) { f[j] = 0.f; gx[j] = 0.f; gy[j] = 0.f; gz[j] = 0.f; }
#pragma unroll
      for (int i = 0; i < K; ++i) {
#pragma unroll
        for (int jj = 0; jj < K; ++jj) {
          const float* base = feat + t.ix[i] * s.sx + t.iy[jj] * s.sy + d;
          Vec<V> v[K];
#pragma unroll
          for (int k = 0; k < K; ++k) v[k] = ldg_vec<V>(base + t.iz[k] * s.sz);
#pragma unroll
          for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
              if (MODE == FWD) {
                f[j] += t.cx[i] * t.cy[jj] * t.cz[k] * v[k].v[j];            // :78-79
              } else {
                gx[j] += g.sx * t.gx[i] * t.cy[jj] * t.cz[k] * v[k].v[j];     // :168-170
                gy[j] += g.sy * t.cx[i] * t.gy[jj] * t.cz[k] * v[k].v[j];
                gz[j] += g.sz * t.cx[i] * t.cy[jj] * t.gz[k] * v[k].v[j];
              }
            }
          }
        }
      }
      if (MODE == FWD || MODE == GGO) {
        Vec<V> o;
#pragma unroll
        for (int j = 0; j < V; ++j) o.v[j] = MODE == FWD ? f[j] : (ggx * gx[j] + ggy * gy[j] + ggz * gz[j]);
        float* op = out + p * D + d;
        if (ACCUM) {
          Vec<V> prev = ld_vec<V>(op);
#pragma unroll
          for (int j = 0; j < V; ++j) o.v[j] += prev.v[j];
        }
        st_vec<V>(op, o);
      } else {
        Vec<V> go = ldg_vec<V>(a + p * D + d);
#pragma unroll
        for (int j = 0; j < V; ++j) { ax += go.v[j] * gx[j]; ay += go.v[j] * gy[j]; az += go.v[j] * gz[j]; }
      }
    }
    if (MODE == GRAD_QUERY) {
      float* op = out + p * 3;
      if (ACCUM) { ax += op[0]; ay += op[1]; az += op[2]; }
      op[0] = ax; op[1] = ay; op[2] = az;
    }
  }
}

template <bool SECOND, int V, bool AGG>
__global__ void __launch_bounds__(NDJIR_BLOCK)
scatter_kernel(long long B, float* __restrict__ gf, const float* __restrict__ go_, const float* __restrict__ gg,
               const float* __restrict__ query, GridFrame g, Strides s, int D) {
  long long stride = (long long)gridDim.x * blockDim.x;
  long long start = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long rounds = (B + stride - 1) / stride;
  for (long long r = 0; r < rounds; ++r) {
    long long p = start + r * stride;
    bool active = p < B;
    long long pc = active ? p : (B - 1);
    Taps t = make_taps<SECOND>(g, query + pc * 3);
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
    if (SECOND) {
      ggx = __ldg(gg + pc * 3) * g.sx; ggy = __ldg(gg + pc * 3 + 1) * g.sy; ggz = __ldg(gg + pc * 3 + 2) * g.sz;
    }
    for (int d = 0; d < D; d += V) {
      Vec<V> o = ldg_vec<V>(go_ + pc * D + d);
#pragma unroll
      for (int i = 0; i < K; ++i) {
#pragma unroll
        for (int jj = 0; jj < K; ++jj) {
#pragma unroll
          for (int k = 0; k < K; ++k) {
            float coef = SECOND ? (ggx * (t.gx[i] * t.cy[jj] * t.cz[k]) + ggy * (t.cx[i] * t.gy[jj] * t.cz[k]) +
                                   ggz * (t.cx[i] * t.cy[jj] * t.gz[k]))
                                : t.cx[i] * t.cy[jj] * t.cz[k];
            Vec<V> val;
#pragma unroll
            for (int j = 0; j < V; ++j) val.v[j] = o.v[j] * coef;
            unsigned idx = t.ix[i] * s.sx + t.iy[jj] * s.sy + t.iz[k] * s.sz + d;
            if (AGG) warp_agg_red<V>(gf + idx, (unsigned long long)idx, val, active);
            else if (active) red_vec<V>(gf + idx, val);
          }
        }
      }
    }
  }
}

// ---- brick-ordered variants for large batches on tables far larger than L2 (bench shape: 2^24 points, 256^3 x 4 =
// 268 MB): the direct kernels are DRAM-bound (64 taps = 16 z-runs of 64 B per point, ~35 % L2 hits).  The points are
// counting-sorted by table brick (voxel_binned.cu: build_records) and visited in that order, ONE CTA PER 256 RECORDS
// WITHOUT A GRID-STRIDE LOOP so that the records in flight stay inside a narrow window of the table.  Records: forward
// {q, point index}; grad_feature with D = 4 {q, index | grad_output row}. ------------------------------------------------
// FOUR LANES PER POINT, lane k owns z tap k: the four lanes of a point read the four z-neighbours of every (x, y)
// column - one contiguous 64-byte run - in ONE instruction, so a warp-wide load touches 8-16 cache lines instead of 32
// (the one-thread-per-point kernels are bound by the L1 tag stage: 64 fully divergent 16-byte loads per point).  Each lane
// evaluates the window weight of tap k on the three axes, the other taps arrive by shuffle.  REC: brick-ordered records
// instead of the caller's point order (one CTA per 64 records, no grid-stride loop).  The sum over z taps is taken last
// (across lanes), so results differ from the one-thread kernels by fp32 summation order only.
__device__ __forceinline__ void lane_tap(float q, float mn, float sc, float g1, int k, unsigned& idx, float& c) {
  float x = __fmul_rn(__fsub_rn(q, mn), sc);
  float x0 = floorf(x);
  float xi = fminf(fmaxf(x0 + (float)(k - W + 1), 0.f), g1);
  c = lanczos_w(__fsub_rn(x, xi), W);
  idx = (unsigned)xi;
}

template <bool REC, bool ACCUM>
__global__ void __launch_bounds__(NDJIR_BLOCK)
gather4_kernel(long long B, float* __restrict__ out, const float* __restrict__ query, const float4* __restrict__ rec,
               const float* __restrict__ feat, GridFrame g, Strides s, int D) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long i = w >> 2;
  const int k = (int)(w & 3);
  const bool active = i < B;
  if (!active) i = B - 1;
  float qx, qy, qz;
  long long p = i;
  if (REC) {
    float4 rc = __ldg(rec + i);
    qx = rc.x; qy = rc.y; qz = rc.z; p = (long long)__float_as_uint(rc.w);
  } else {
    qx = __ldg(query + i * 3); qy = __ldg(query + i * 3 + 1); qz = __ldg(query + i * 3 + 2);
  }
  unsigned mx, my, mz;
  float wx, wy, wz;
  lane_tap(qx, g.mnx, g.sx, g.gx1, k, mx, wx);
  lane_tap(qy, g.mny, g.sy, g.gy1, k, my, wy);
  lane_tap(qz, g.mnz, g.sz, g.gz1, k, mz, wz);
  unsigned ix[K], iy[K];
  float cx[K], cy[K];
#pragma unroll
  for (int a = 0; a < K; ++a) {
    ix[a] = __shfl_sync(0xffffffffu, mx, a, 4); cx[a] = __shfl_sync(0xffffffffu, wx, a, 4);
    iy[a] = __shfl_sync(0xffffffffu, my, a, 4); cy[a] = __shfl_sync(0xffffffffu, wy, a, 4);
  }
  const float* fz = feat + mz * s.sz;
  for (int d = 0; d < D; d += 4) {
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ii = 0; ii < K; ++ii) {
      float4 v[K];
#pragma unroll
      for (int jj = 0; jj < K; ++jj) v[jj] = __ldg(reinterpret_cast<const float4*>(fz + ix[ii] * s.sx + iy[jj] * s.sy + d));
#pragma unroll
      for (int jj = 0; jj < K; ++jj) {
        float c = cx[ii] * cy[jj];
        f.x += c * v[jj].x; f.y += c * v[jj].y; f.z += c * v[jj].z; f.w += c * v[jj].w;
      }
    }
    f.x *= wz; f.y *= wz; f.z *= wz; f.w *= wz;
#pragma unroll
    for (int m = 1; m < 4; m <<= 1) {
      f.x += __shfl_xor_sync(0xffffffffu, f.x, m); f.y += __shfl_xor_sync(0xffffffffu, f.y, m);
      f.z += __shfl_xor_sync(0xffffffffu, f.z, m); f.w += __shfl_xor_sync(0xffffffffu, f.w, m);
    }
    if (active && k == 0) {
      float4* op = reinterpret_cast<float4*>(out + p * D + d);
      if (ACCUM) { float4 pv = *op; f.x += pv.x; f.y += pv.y; f.z += pv.z; f.w += pv.w; }
      *op = f;
    }
  }
}

// first-order scatter over wide records (D = 4), four lanes per point, lane k owns z tap k (see gather4_kernel)
__global__ void __launch_bounds__(NDJIR_BLOCK)
scatter_rec_kernel(long long B, float* __restrict__ gf, const float4* __restrict__ rec, GridFrame g, Strides s) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long i = w >> 2;
  const int k = (int)(w & 3);
  const bool active = i < B;
  if (!active) i = B - 1;
  float4 rc = __ldg(rec + 2 * i), go = __ldg(rec + 2 * i + 1);
  unsigned mx, my, mz;
  float wx, wy, wz;
  lane_tap(rc.x, g.mnx, g.sx, g.gx1, k, mx, wx);
  lane_tap(rc.y, g.mny, g.sy, g.gy1, k, my, wy);
  lane_tap(rc.z, g.mnz, g.sz, g.gz1, k, mz, wz);
  unsigned ix[K], iy[K];
  float cx[K], cy[K];
#pragma unroll
  for (int a = 0; a < K; ++a) {
    ix[a] = __shfl_sync(0xffffffffu, mx, a, 4); cx[a] = __shfl_sync(0xffffffffu, wx, a, 4);
    iy[a] = __shfl_sync(0xffffffffu, my, a, 4); cy[a] = __shfl_sync(0xffffffffu, wy, a, 4);
  }
  if (!active) return;
  float* gz = gf + mz * s.sz;
#pragma unroll
  for (int ii = 0; ii < K; ++ii) {
#pragma unroll
    for (int jj = 0; jj < K; ++jj) {
      float coef = cx[ii] * cy[jj] * wz;
      Vec<4> val;
      val.v[0] = go.x * coef; val.v[1] = go.y * coef; val.v[2] = go.z * coef; val.v[3] = go.w * coef;
      red_vec<4>(gz + ix[ii] * s.sx + iy[jj] * s.sy, val);
    }
  }
}

static bool bad_grid(const int* G, int D) {
  if (!G || D <= 0 || G[0] <= 0 || G[1] <= 0 || G[2] <= 0) return true;
  return (long long)G[0] * G[1] * G[2] * D >= (1ll << 32);
}

static Strides make_strides(const int* G, int D) {
  Strides s;
  s.sx = (unsigned)G[1] * (unsigned)G[2] * (unsigned)D;
  s.sy = (unsigned)G[2] * (unsigned)D;
  s.sz = (unsigned)D;
  return s;
}

template <int MODE>
static int launch_gather(long long B, float* out, const float* a, const float* gg, const float* query,
                         const float* feat, const int* G, int D, const float* mn, const float* mx, bool accum,
                         cudaStream_t st) {
  if (B == 0) return NDJIR_OK;
  if (B < 0 || bad_grid(G, D) || !out || !query || !feat || !mn || !mx) return NDJIR_ERR_ARG;
  GridFrame g = make_frame(G[0], G[1], G[2], mn, mx);
  Strides s = make_strides(G, D);
  int V = pick_vec(D, feat, MODE == GRAD_QUERY ? (const void*)a : (const void*)out);
  int grid = grid_for(B);
  if (MODE == FWD && V == 4) {     // four lanes per point; brick-ordered for large batches on large tables
    unsigned grid4 = (unsigned)((B * 4 + NDJIR_BLOCK - 1) / NDJIR_BLOCK);
    if (voxel_binned::worthwhile(B, G, D)) {
      long long wsb = voxel_binned::workspace_bytes(B);
      if (void* ws = voxel_binned::scratch_alloc(wsb, st)) {
        const float4* rec = nullptr;
        int rc = voxel_binned::build_records(B, query, nullptr, g, G, D, ws, wsb, st, &rec);
        if (rc == NDJIR_OK) {
          if (accum) gather4_kernel<true, true><<<grid4, NDJIR_BLOCK, 0, st>>>(B, out, query, rec, feat, g, s, D);
          else gather4_kernel<true, false><<<grid4, NDJIR_BLOCK, 0, st>>>(B, out, query, rec, feat, g, s, D);
          cudaError_t e = cudaGetLastError();
          rc = e == cudaSuccess ? NDJIR_OK : (int)e;
        }
        voxel_binned::scratch_free(ws, st);
        return rc;
      }
    }
    if (accum) gather4_kernel<false, true><<<grid4, NDJIR_BLOCK, 0, st>>>(B, out, query, nullptr, feat, g, s, D);
    else gather4_kernel<false, false><<<grid4, NDJIR_BLOCK, 0, st>>>(B, out, query, nullptr, feat, g, s, D);
    NDJIR_RETURN_LAST_ERROR();
  }
#define NDJIR_LAUNCH(VV)                                                                                   \
  if (accum) gather_kernel<MODE, VV, true><<<grid, NDJIR_BLOCK, 0, st>>>(B, out, a, gg, query, feat, g, s, D); \
  else gather_kernel<MODE, VV, false><<<grid, NDJIR_BLOCK, 0, st>>>(B, out, a, gg, query, feat, g, s, D);
  if (V == 4) { NDJIR_LAUNCH(4) } else if (V == 2) { NDJIR_LAUNCH(2) } else { NDJIR_LAUNCH(1) }
#undef NDJIR_LAUNCH
  NDJIR_RETURN_LAST_ERROR();
}

template <bool SECOND>
static int launch_scatter(long long B, float* gf, const float* go, const float* gg, const float* query,
                          const int* G, int D, const float* mn, const float* mx, cudaStream_t st) {
  if (B == 0) return NDJIR_OK;
  if (B < 0 || bad_grid(G, D) || !gf || !go || !query || !mn || !mx) return NDJIR_ERR_ARG;
  GridFrame g = make_frame(G[0], G[1], G[2], mn, mx);
  Strides s = make_strides(G, D);
  int V = pick_vec(D, gf, go);
  int grid = grid_for(B);
  bool agg = g_scatter_aggregate != 0;
  if (!SECOND && D == 4 && V == 4 && voxel_binned::worthwhile(B, G, D)) {
    long long wsb = voxel_binned::workspace_bytes(B);
    if (void* ws = voxel_binned::scratch_alloc(wsb, st)) {
      const float4* rec = nullptr;
      int rc = voxel_binned::build_records(B, query, go, g, G, D, ws, wsb, st, &rec);
      if (rc == NDJIR_OK) {
        scatter_rec_kernel<<<(unsigned)((B * 4 + NDJIR_BLOCK - 1) / NDJIR_BLOCK), NDJIR_BLOCK, 0, st>>>(B, gf, rec, g, s);
        cudaError_t e = cudaGetLastError();
        rc = e == cudaSuccess ? NDJIR_OK : (int)e;
      }
      voxel_binned::scratch_free(ws, st);
      return rc;
    }
  }
#define NDJIR_LAUNCH(VV)                                                                                  \
  if (agg) scatter_kernel<SECOND, VV, true><<<grid, NDJIR_BLOCK, 0, st>>>(B, gf, go, gg, query, g, s, D); \
  else scatter_kernel<SECOND, VV, false><<<grid, NDJIR_BLOCK, 0, st>>>(B, gf, go, gg, query, g, s, D);
  if (V == 4) { NDJIR_LAUNCH(4) } else if (V == 2) { NDJIR_LAUNCH(2) } else { NDJIR_LAUNCH(1) }
#undef NDJIR_LAUNCH
  NDJIR_RETURN_LAST_ERROR();
}

}  // namespace lanczos
}  // namespace ndjir

using namespace ndjir;
using namespace ndjir::lanczos;

extern "C" {

int ndjir_lanczos_voxel_query_on_voxel(long long n_points, float* output, const float* query,
                                       const float* feature, const int* grid_sizes, int D, const float* min3,
                                       const float* max3, int accum, cudaStream_t stream) {
  return launch_gather<FWD>(n_points, output, nullptr, nullptr, query, feature, grid_sizes, D, min3, max3,
                            accum != 0, stream);
}

int ndjir_lanczos_voxel_grad_query(long long n_points, float* grad_query, const float* grad_output,
                                   const float* query, const float* feature, const int* grid_sizes, int D,
                                   const float* min3, const float* max3, int accum, cudaStream_t stream) {
  if (n_points > 0 && !grad_output) return NDJIR_ERR_ARG;
  return launch_gather<GRAD_QUERY>(n_points, grad_query, grad_output, nullptr, query, feature, grid_sizes, D,
                                   min3, max3, accum != 0, stream);
}

int ndjir_lanczos_voxel_grad_feature(long long n_points, float* grad_feature, const float* grad_output,
                                     const float* query, const int* grid_sizes, int D, const float* min3,
                                     const float* max3, int accum, cudaStream_t stream) {
  if (bad_grid(grid_sizes, D) || !grad_feature) return NDJIR_ERR_ARG;
  if (!accum) fill_zero(grad_feature, (long long)grid_sizes[0] * grid_sizes[1] * grid_sizes[2] * D, stream);
  return launch_scatter<false>(n_points, grad_feature, grad_output, nullptr, query, grid_sizes, D, min3, max3,
                               stream);
}

int ndjir_lanczos_voxel_grad_query_grad_grad_output(long long n_points, float* grad_grad_output,
                                                    const float* grad_grad_query, const float* query,
                                                    const float* feature, const int* grid_sizes, int D,
                                                    const float* min3, const float* max3, int accum,
                                                    cudaStream_t stream) {
  if (n_points > 0 && !grad_grad_query) return NDJIR_ERR_ARG;
  return launch_gather<GGO>(n_points, grad_grad_output, nullptr, grad_grad_query, query, feature, grid_sizes, D,
                            min3, max3, accum != 0, stream);
}

// Always accumulates (lanczos_voxel_feature_cuda.cu:585-607 has no zero-fill).
int ndjir_lanczos_voxel_grad_query_grad_feature(long long n_points, float* grad_feature,
                                                const float* grad_grad_query, const float* grad_output,
                                                const float* query, const int* grid_sizes, int D,
                                                const float* min3, const float* max3, cudaStream_t stream) {
  if (n_points > 0 && !grad_grad_query) return NDJIR_ERR_ARG;
  return launch_scatter<true>(n_points, grad_feature, grad_output, grad_grad_query, query, grid_sizes, D, min3,
                              max3, stream);
}

}  // extern "C"
