// Triplane (bilinear on 3 planes) and triline (linear on 3 lines) feature query + backward (sm_100a).
//
// Replaces csrc/grid_feature/triplane_feature_cuda.cu and triline_feature_cuda.cu (5 exports each).
// Layout: triplane feature (3,G,G,D), triline feature (3,G,D), channel-last; output (B, D*3) with channel
// index c = d*3 + i (common.cuh:29-35), plane 0=(x,y), 1=(y,z), 2=(z,x) (common_triplane.cuh:24-35),
// line i <-> axis i (common_triline.cuh).
// Thread mapping: one thread per point does all 3 planes/lines (the reference: one thread per
// (point,channel,plane), i.e. 3*D threads re-deriving the same cell).  Channels are processed V at a time
// with 16/8/4-byte loads; the 3*V outputs of a chunk are contiguous in the (B, D*3) row and are
// written as 3 vector stores.
// Roofline: HBM gather; triplane G=2048,D=8 fwd = 12+96+3*4*32 = 492 B/pt; triline = 12+96 = 108 B/pt
// (its 192 KiB table is L2-resident).
#include "grid_common.cuh"
#include "chunk_io.cuh"
#include "../../include/ndjir_b200.h"

namespace ndjir {
namespace tpl {

enum Mode { FWD = 0, GRAD_QUERY = 1, GGO = 2 };

struct Axis2 {  // the two axes of a plane, as seen from the cell
  unsigned u0, u1, v0, v1;
  float a0, a1, b0, b1, su, sv, ggu, ggv;
  int au, av;
};

__device__ __forceinline__ Axis2 plane_axes(int i, const Cell& c, const GridFrame& g, float ggx, float ggy,
                                            float ggz) {
  Axis2 r;
  if (i == 0) {
    r.u0 = c.x0; r.u1 = c.x1; r.v0 = c.y0; r.v1 = c.y1; r.a0 = c.p0; r.a1 = c.p1; r.b0 = c.q0; r.b1 = c.q1;
    r.su = c.sx; r.sv = c.sy; r.ggu = ggx; r.ggv = ggy; r.au = 0; r.av = 1;
  } else if (i == 1) {
    r.u0 = c.y0; r.u1 = c.y1; r.v0 = c.z0; r.v1 = c.z1; r.a0 = c.q0; r.a1 = c.q1; r.b0 = c.r0; r.b1 = c.r1;
    r.su = c.sy; r.sv = c.sz; r.ggu = ggy; r.ggv = ggz; r.au = 1; r.av = 2;
  } else {
    r.u0 = c.z0; r.u1 = c.z1; r.v0 = c.x0; r.v1 = c.x1; r.a0 = c.r0; r.a1 = c.r1; r.b0 = c.p0; r.b1 = c.p1;
    r.su = c.sz; r.sv = c.sx; r.ggu = ggz; r.ggv = ggx; r.au = 2; r.av = 0;
  }
  return r;
}

// PLANE=true: triplane, PLANE=false: triline.
template <bool PLANE, int MODE, int V, bool ACCUM>
__global__ void __launch_bounds__(NDJIR_BLOCK)
gather_kernel(long long B, float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b,
              const float* __restrict__ query, const float* __restrict__ feat, GridFrame g, int G, int D) {
  long long stride = (long long)gridDim.x * blockDim.x;
  const long long plane_elems = PLANE ? (long long)G * G * D : (long long)G * D;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < B; p += stride) {
    const float* q = query + p * 3;
    Cell c = make_cell(g, __ldg(q), __ldg(q + 1), __ldg(q + 2));
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
    if (MODE == GGO) { ggx = __ldg(b + p * 3); ggy = __ldg(b + p * 3 + 1); ggz = __ldg(b + p * 3 + 2); }
    float acc[3] = {0.f, 0.f, 0.f};
    for (int d = 0; d < D; d += V) {
      float o[3][V];
      float go[3][V];
      if (MODE == GRAD_QUERY) load_chunk<V>(a + p * 3 * D, d, go);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        Axis2 x = plane_axes(i, c, g, ggx, ggy, ggz);
        const float* fi = feat + i * plane_elems + d;
        if (PLANE) {
          Vec<V> f00 = ldg_vec<V>(fi + ((long long)x.u0 * G + x.v0) * D);
          Vec<V> f01 = ldg_vec<V>(fi + ((long long)x.u0 * G + x.v1) * D);
          Vec<V> f10 = ldg_vec<V>(fi + ((long long)x.u1 * G + x.v0) * D);
          Vec<V> f11 = ldg_vec<V>(fi + ((long long)x.u1 * G + x.v1) * D);
#pragma unroll
          for (int j = 0; j < V; ++j) {
            if (MODE == FWD) {
              // triplane_feature_cuda.cu:86
              o[i][j] = x.a0 * x.b0 * f00.v[j] + x.a0 * x.b1 * f01.v[j] + x.a1 * x.b0 * f10.v[j] +
                        x.a1 * x.b1 * f11.v[j];
            } else {
              // triplane_feature_cuda.cu:159-167
              float du = x.su * (x.b0 * (f10.v[j] - f00.v[j]) + x.b1 * (f11.v[j] - f01.v[j]));
              float dv = x.sv * (x.a0 * (f01.v[j] - f00.v[j]) + x.a1 * (f11.v[j] - f10.v[j]));
              if (MODE == GRAD_QUERY) {
                acc[x.au] += go[i][j] * du;
                acc[x.av] += go[i][j] * dv;
              } else {
                o[i][j] = x.ggu * du + x.ggv * dv;
              }
            }
          }
        } else {
          Vec<V> f0 = ldg_vec<V>(fi + (long long)x.u0 * D);
          Vec<V> f1 = ldg_vec<V>(fi + (long long)x.u1 * D);
#pragma unroll
          for (int j = 0; j < V; ++j) {
            if (MODE == FWD) {
              o[i][j] = x.a0 * f0.v[j] + x.a1 * f1.v[j];  // triline_feature_cuda.cu:79
            } else {
              float du = x.su * (f1.v[j] - f0.v[j]);      // triline_feature_cuda.cu:152-158
              if (MODE == GRAD_QUERY) acc[x.au] += go[i][j] * du;
              else o[i][j] = x.ggu * du;
            }
          }
        }
      }
      if (MODE != GRAD_QUERY) store_chunk<V>(out + p * 3 * D, d, o, ACCUM);
    }
    if (MODE == GRAD_QUERY) {
      float* op = out + p * 3;
      if (ACCUM) { acc[0] += op[0]; acc[1] += op[1]; acc[2] += op[2]; }
      op[0] = acc[0]; op[1] = acc[1]; op[2] = acc[2];
    }
  }
}

// SECOND=false: kernel_grad_feature (triplane :203-255, triline :187-232)
// SECOND=true : kernel_grad_query_grad_feature (triplane :497-560, triline :468-520)
template <bool PLANE, bool SECOND, int V, bool AGG>
__global__ void __launch_bounds__(NDJIR_BLOCK)
scatter_kernel(long long B, float* __restrict__ gf, const float* __restrict__ go_, const float* __restrict__ gg,
               const float* __restrict__ query, GridFrame g, int G, int D) {
  long long stride = (long long)gridDim.x * blockDim.x;
  long long start = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long rounds = (B + stride - 1) / stride;
  const long long plane_elems = PLANE ? (long long)G * G * D : (long long)G * D;
  for (long long r = 0; r < rounds; ++r) {
    long long p = start + r * stride;
    bool active = p < B;
    long long pc = active ? p : (B - 1);
    const float* q = query + pc * 3;
    Cell c = make_cell(g, __ldg(q), __ldg(q + 1), __ldg(q + 2));
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
    if (SECOND) { ggx = __ldg(gg + pc * 3); ggy = __ldg(gg + pc * 3 + 1); ggz = __ldg(gg + pc * 3 + 2); }
    for (int d = 0; d < D; d += V) {
      float go[3][V];
      load_chunk<V>(go_ + pc * 3 * D, d, go);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        Axis2 x = plane_axes(i, c, g, ggx, ggy, ggz);
        float* gi = gf + i * plane_elems + d;
        if (PLANE) {
          const unsigned us[2] = {x.u0, x.u1}, vs[2] = {x.v0, x.v1};
          const float as[2] = {x.a0, x.a1}, bs[2] = {x.b0, x.b1};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            int cu = k >> 1, cv = k & 1;
            float coef = SECOND ? (x.ggu * x.su * ((cu ? 1.f : -1.f) * bs[cv]) +
                                   x.ggv * x.sv * ((cv ? 1.f : -1.f) * as[cu]))
                                : as[cu] * bs[cv];
            Vec<V> val;
#pragma unroll
            for (int j = 0; j < V; ++j) val.v[j] = go[i][j] * coef;
            long long off = ((long long)us[cu] * G + vs[cv]) * D;
            if (AGG) warp_agg_red<V>(gi + off, (unsigned long long)(i * plane_elems + off + d), val, active);
            else if (active) red_vec<V>(gi + off, val);
          }
        } else {
          const unsigned us[2] = {x.u0, x.u1};
          const float as[2] = {x.a0, x.a1};
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            float coef = SECOND ? (x.ggu * x.su * (k ? 1.f : -1.f)) : as[k];
            Vec<V> val;
#pragma unroll
            for (int j = 0; j < V; ++j) val.v[j] = go[i][j] * coef;
            long long off = (long long)us[k] * D;
            if (AGG) warp_agg_red<V>(gi + off, (unsigned long long)(i * plane_elems + off + d), val, active);
            else if (active) red_vec<V>(gi + off, val);
          }
        }
      }
    }
  }
}

// Forward with one thread per (point, channel chunk of V): D/V times as many independent gathers in flight as the
// one-thread-per-point mapping, and the 3*V outputs of a chunk are still contiguous (three vector stores).
// (A thread per (point, plane) was tried and lost to its stride-3 scalar stores: 4.18 ms vs 2.90 ms at 2^24 points.)
template <bool PLANE, int V, bool ACCUM>
__global__ void __launch_bounds__(NDJIR_BLOCK)
gather_fwd_chunk_kernel(long long B, float* __restrict__ out, const float* __restrict__ query,
                        const float* __restrict__ feat, GridFrame g, int G, int D) {
  long long stride = (long long)gridDim.x * blockDim.x;
  const long long plane_elems = PLANE ? (long long)G * G * D : (long long)G * D;
  const int nchunk = D / V;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < B * nchunk; w += stride) {
    long long p = w / nchunk;
    int d = (int)(w - p * nchunk) * V;
    const float* q = query + p * 3;
    Cell c = make_cell(g, __ldg(q), __ldg(q + 1), __ldg(q + 2));
    float o[3][V];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      Axis2 x = plane_axes(i, c, g, 0.f, 0.f, 0.f);
      const float* fi = feat + i * plane_elems + d;
      if (PLANE) {
        Vec<V> f00 = ldg_vec<V>(fi + ((long long)x.u0 * G + x.v0) * D);
        Vec<V> f01 = ldg_vec<V>(fi + ((long long)x.u0 * G + x.v1) * D);
        Vec<V> f10 = ldg_vec<V>(fi + ((long long)x.u1 * G + x.v0) * D);
        Vec<V> f11 = ldg_vec<V>(fi + ((long long)x.u1 * G + x.v1) * D);
#pragma unroll
        for (int j = 0; j < V; ++j)
          o[i][j] = x.a0 * x.b0 * f00.v[j] + x.a0 * x.b1 * f01.v[j] + x.a1 * x.b0 * f10.v[j] + x.a1 * x.b1 * f11.v[j];
      } else {
        Vec<V> f0 = ldg_vec<V>(fi + (long long)x.u0 * D);
        Vec<V> f1 = ldg_vec<V>(fi + (long long)x.u1 * D);
#pragma unroll
        for (int j = 0; j < V; ++j) o[i][j] = x.a0 * f0.v[j] + x.a1 * f1.v[j];
      }
    }
    store_chunk<V>(out + p * 3 * D, d, o, ACCUM);
  }
}

// Scatter with one thread per (point, plane|line, corner, channel chunk of V): the chunks of a cell sit in ADJACENT
// lanes, so one warp-wide reduction instruction carries whole 32-byte sectors (D = 8: two 16-byte halves side by side)
// instead of issuing the halves of a cell from two instructions of the same lane - half as many L2 transactions on
// tables whose reductions are contention-bound (triline: 49 152 cells for 2^24 points).
template <bool PLANE, bool SECOND, int V>
__global__ void __launch_bounds__(NDJIR_BLOCK)
scatter_split_kernel(long long B, float* __restrict__ gf, const float* __restrict__ go_, const float* __restrict__ gg,
                     const float* __restrict__ query, GridFrame g, int G, int D) {
  constexpr int NC = PLANE ? 4 : 2;
  const int nchunk = D / V;
  long long stride = (long long)gridDim.x * blockDim.x;
  const long long plane_elems = PLANE ? (long long)G * G * D : (long long)G * D;
  const long long per_point = 3ll * NC * nchunk;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < B * per_point; w += stride) {
    long long p = w / per_point;
    int rem = (int)(w - p * per_point);
    int i = rem / (NC * nchunk);
    rem -= i * NC * nchunk;
    int k = rem / nchunk;
    int d = (rem - k * nchunk) * V;
    const float* q = query + p * 3;
    Cell c = make_cell(g, __ldg(q), __ldg(q + 1), __ldg(q + 2));
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
    if (SECOND) { ggx = __ldg(gg + p * 3); ggy = __ldg(gg + p * 3 + 1); ggz = __ldg(gg + p * 3 + 2); }
    Axis2 x = plane_axes(i, c, g, ggx, ggy, ggz);
    float coef;
    long long off;
    if (PLANE) {
      int cu = k >> 1, cv = k & 1;
      float au = cu ? x.a1 : x.a0, bv = cv ? x.b1 : x.b0;
      coef = SECOND ? (x.ggu * x.su * ((cu ? 1.f : -1.f) * bv) + x.ggv * x.sv * ((cv ? 1.f : -1.f) * au)) : au * bv;
      off = ((long long)(cu ? x.u1 : x.u0) * G + (cv ? x.v1 : x.v0)) * D;
    } else {
      coef = SECOND ? (x.ggu * x.su * (k ? 1.f : -1.f)) : (k ? x.a1 : x.a0);
      off = (long long)(k ? x.u1 : x.u0) * D;
    }
    const float* grow = go_ + p * 3 * D + i;
    Vec<V> val;
#pragma unroll
    for (int j = 0; j < V; ++j) val.v[j] = __ldg(grow + (d + j) * 3) * coef;
    red_vec<V>(gf + i * plane_elems + off + d, val);
  }
}

static bool bad(int G, int D) {
  return G <= 0 || D <= 0 || (long long)G * G * D * 3 >= (1ll << 40);
}

template <bool PLANE, int MODE>
static int launch_gather(long long B, float* out, const float* a, const float* b, const float* query,
                         const float* feat, int G, int D, const float* mn, const float* mx, bool accum,
                         cudaStream_t st, int interp = INTERP_LINEAR) {
  if (B == 0) return NDJIR_OK;
  if (B < 0 || bad(G, D) || !out || !query || !feat || !mn || !mx) return NDJIR_ERR_ARG;
  GridFrame g = make_frame(G, G, G, mn, mx);
  g.interp = interp;
  int V = pick_vec(D, feat, MODE == GRAD_QUERY ? (const void*)a : (const void*)out);
  int grid = grid_for(B);
  if (MODE == FWD && D / V > 1) {
    int gridc = grid_for(B * (D / V));
#define NDJIR_LAUNCH_S(VV)                                                                                           \
  if (accum) gather_fwd_chunk_kernel<PLANE, VV, true><<<gridc, NDJIR_BLOCK, 0, st>>>(B, out, query, feat, g, G, D);   \
  else gather_fwd_chunk_kernel<PLANE, VV, false><<<gridc, NDJIR_BLOCK, 0, st>>>(B, out, query, feat, g, G, D);
    if (V == 4) { NDJIR_LAUNCH_S(4) } else if (V == 2) { NDJIR_LAUNCH_S(2) } else { NDJIR_LAUNCH_S(1) }
#undef NDJIR_LAUNCH_S
    NDJIR_RETURN_LAST_ERROR();
  }
#define NDJIR_LAUNCH(VV)                                                                                          \
  if (accum) gather_kernel<PLANE, MODE, VV, true><<<grid, NDJIR_BLOCK, 0, st>>>(B, out, a, b, query, feat, g, G, D); \
  else gather_kernel<PLANE, MODE, VV, false><<<grid, NDJIR_BLOCK, 0, st>>>(B, out, a, b, query, feat, g, G, D);
  if (V == 4) { NDJIR_LAUNCH(4) } else if (V == 2) { NDJIR_LAUNCH(2) } else { NDJIR_LAUNCH(1) }
#undef NDJIR_LAUNCH
  NDJIR_RETURN_LAST_ERROR();
}

template <bool PLANE, bool SECOND>
static int launch_scatter(long long B, float* gf, const float* go, const float* gg, const float* query, int G,
                          int D, const float* mn, const float* mx, cudaStream_t st, int interp = INTERP_LINEAR) {
  if (B == 0) return NDJIR_OK;
  if (B < 0 || bad(G, D) || !gf || !go || !query || !mn || !mx) return NDJIR_ERR_ARG;
  GridFrame g = make_frame(G, G, G, mn, mx);
  g.interp = interp;
  int V = pick_vec(D, gf, go);
  int grid = grid_for(B);
  bool agg = g_scatter_aggregate != 0;
  if (!agg) {
    // triline: the table is tiny (49 152 cells at G=2048) and the reductions are contention-bound (1.6 ms at 2^24
    // points; tried and rejected: scalar reductions 6.0 ms, a shared-memory private copy of the table 3.1 ms)
    int V2 = pick_vec(D, gf);
    int gridc = grid_for(B * 3 * (PLANE ? 4 : 2) * (D / V2));
    if (V2 == 4) scatter_split_kernel<PLANE, SECOND, 4><<<gridc, NDJIR_BLOCK, 0, st>>>(B, gf, go, gg, query, g, G, D);
    else if (V2 == 2) scatter_split_kernel<PLANE, SECOND, 2><<<gridc, NDJIR_BLOCK, 0, st>>>(B, gf, go, gg, query, g, G, D);
    else scatter_split_kernel<PLANE, SECOND, 1><<<gridc, NDJIR_BLOCK, 0, st>>>(B, gf, go, gg, query, g, G, D);
    NDJIR_RETURN_LAST_ERROR();
  }
#define NDJIR_LAUNCH(VV)                                                                                       \
  if (agg) scatter_kernel<PLANE, SECOND, VV, true><<<grid, NDJIR_BLOCK, 0, st>>>(B, gf, go, gg, query, g, G, D); \
  else scatter_kernel<PLANE, SECOND, VV, false><<<grid, NDJIR_BLOCK, 0, st>>>(B, gf, go, gg, query, g, G, D);
  if (V == 4) { NDJIR_LAUNCH(4) } else if (V == 2) { NDJIR_LAUNCH(2) } else { NDJIR_LAUNCH(1) }
#undef NDJIR_LAUNCH
  NDJIR_RETURN_LAST_ERROR();
}

}  // namespace tpl
}  // namespace ndjir

using namespace ndjir;
using namespace ndjir::tpl;

#define NDJIR_DEFINE_FAMILY(NAME, FWDNAME, PLANE, TABLE_ELEMS, INTERP)                                            \
  int ndjir_##NAME##_##FWDNAME(long long n, float* output, const float* query, const float* feature, int G,      \
                               int D, const float* min3, const float* max3, int accum, cudaStream_t st) {        \
    return launch_gather<PLANE, FWD>(n, output, nullptr, nullptr, query, feature, G, D, min3, max3, accum != 0,  \
                                     st, INTERP);                                                                \
  }                                                                                                              \
  int ndjir_##NAME##_grad_query(long long n, float* grad_query, const float* grad_output, const float* query,    \
                                const float* feature, int G, int D, const float* min3, const float* max3,        \
                                int accum, cudaStream_t st) {                                                    \
    if (n > 0 && !grad_output) return NDJIR_ERR_ARG;                                                             \
    return launch_gather<PLANE, GRAD_QUERY>(n, grad_query, grad_output, nullptr, query, feature, G, D, min3,     \
                                            max3, accum != 0, st, INTERP);                                       \
  }                                                                                                              \
  int ndjir_##NAME##_grad_feature(long long n, float* grad_feature, const float* grad_output,                    \
                                  const float* query, int G, int D, const float* min3, const float* max3,        \
                                  int accum, cudaStream_t st) {                                                  \
    if (bad(G, D) || !grad_feature) return NDJIR_ERR_ARG;                                                        \
    if (!accum) fill_zero(grad_feature, (long long)(TABLE_ELEMS), st);                                           \
    return launch_scatter<PLANE, false>(n, grad_feature, grad_output, nullptr, query, G, D, min3, max3, st,      \
                                        INTERP);                                                                 \
  }                                                                                                              \
  int ndjir_##NAME##_grad_query_grad_grad_output(long long n, float* grad_grad_output,                           \
                                                 const float* grad_grad_query, const float* query,               \
                                                 const float* feature, int G, int D, const float* min3,          \
                                                 const float* max3, int accum, cudaStream_t st) {                \
    if (n > 0 && !grad_grad_query) return NDJIR_ERR_ARG;                                                         \
    return launch_gather<PLANE, GGO>(n, grad_grad_output, nullptr, grad_grad_query, query, feature, G, D, min3,  \
                                     max3, accum != 0, st, INTERP);                                              \
  }                                                                                                              \
  int ndjir_##NAME##_grad_query_grad_feature(long long n, float* grad_feature, const float* grad_grad_query,     \
                                             const float* grad_output, const float* query, int G, int D,         \
                                             const float* min3, const float* max3, cudaStream_t st) {            \
    if (n > 0 && !grad_grad_query) return NDJIR_ERR_ARG;                                                         \
    return launch_scatter<PLANE, true>(n, grad_feature, grad_output, grad_grad_query, query, G, D, min3, max3,   \
                                       st, INTERP);                                                              \
  }

extern "C" {
// triline grad_feature zero-fills 3*G*D floats: the reference zeroes 3*G*G*D there, out of bounds
// (triline_feature_cuda.cu:247, cosine_triline_feature_cuda.cu:248-250) - deliberately not
// reproduced.
NDJIR_DEFINE_FAMILY(triplane, query_on_triplane, true, 3ll * G * G * D, INTERP_LINEAR)
NDJIR_DEFINE_FAMILY(triline, query_on_triline, false, 3ll * G * D, INTERP_LINEAR)
// cosine_triplane_feature_cuda.cu / cosine_triline_feature_cuda.cu (5 exports each): the same kernels with the
// cosine cell (weights 0.5 cos(pi frac) + 0.5, derivative factor 0.5 pi sin(pi frac), grid_common.cuh)
NDJIR_DEFINE_FAMILY(cosine_triplane, query_on_triplane, true, 3ll * G * G * D, INTERP_COSINE)
NDJIR_DEFINE_FAMILY(cosine_triline, query_on_triline, false, 3ll * G * D, INTERP_COSINE)
}
