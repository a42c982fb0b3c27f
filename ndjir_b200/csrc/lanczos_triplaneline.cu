// Lanczos-2 triplane (4x4 taps on 3 planes) and triline (4 taps on 3 lines) feature query + backward (sm_100a).
//
// Replaces csrc/grid_feature/lanczos_triplane_feature_cuda.cu (:795-808) and lanczos_triline_feature_cuda.cu
// (:735-748), 5 exports each.  Layouts as the linear families (triplaneline.cu): feature (3,G,G,D) / (3,G,D), output
// (B, D*3) with c = d*3 + plane, plane 0=(x,y), 1=(y,z), 2=(z,x) (common_triplane.cuh), line i <-> axis i.
// Thread mapping: one thread per point evaluates the 12 window weights (and derivative weights) ONCE - the reference
// runs one thread per (point, channel, plane) and re-evaluates them per thread and inside the tap loops - then visits
// the 3 x 16 (3 x 4) taps with one 16/8/4-byte load per tap for V channels.
#include "grid_common.cuh"
#include "chunk_io.cuh"
#include "lanczos_common.cuh"
#include "../../include/ndjir_b200.h"

namespace ndjir {
namespace lanczos_tpl {

using namespace ndjir::lanczos;
using ndjir::tpl::load_chunk;
using ndjir::tpl::store_chunk;

enum Mode { FWD = 0, GRAD_QUERY = 1, GGO = 2 };

// the two axes of plane i (or the axis of line i) as views into the per-axis taps
struct AxisView {
  const unsigned* iu; const float* cu; const float* gu; float su, ggu; int au;
  const unsigned* iv; const float* cv; const float* gv; float sv, ggv; int av;
};

__device__ __forceinline__ AxisView view(int i, const Taps& t, const GridFrame& g, float ggx, float ggy, float ggz) {
  AxisView r;
  if (i == 0) {
    r.iu = t.ix; r.cu = t.cx; r.gu = t.gx; r.su = g.sx; r.ggu = ggx; r.au = 0;
    r.iv = t.iy; r.cv = t.cy; r.gv = t.gy; r.sv = g.sy; r.ggv = ggy; r.av = 1;
  } else if (i == 1) {
    r.iu = t.iy; r.cu = t.cy; r.gu = t.gy; r.su = g.sy; r.ggu = ggy; r.au = 1;
    r.iv = t.iz; r.cv = t.cz; r.gv = t.gz; r.sv = g.sz; r.ggv = ggz; r.av = 2;
  } else {
    r.iu = t.iz; r.cu = t.cz; r.gu = t.gz; r.su = g.sz; r.ggu = ggz; r.au = 2;
    r.iv = t.ix; r.cv = t.cx; r.gv = t.gx; r.sv = g.sx; r.ggv = ggx; r.av = 0;
  }
  return r;
}

// FWD: out (B, D*3) = interp;  GRAD_QUERY: out (B,3) = sum_c a[b,c] d f_c/dq;  GGO: out (B, D*3) = gg . d f_c/dq
template <bool PLANE, int MODE, int V, bool ACCUM>
__global__ void __launch_bounds__(NDJIR_BLOCK)
gather_kernel(long long B, float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ gg,
              const float* __restrict__ query, const float* __restrict__ feat, GridFrame g, int G, int D) {
  long long stride = (long long)gridDim.x * blockDim.x;
  const long long plane_elems = PLANE ? (long long)G * G * D : (long long)G * D;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < B; p += stride) {
    Taps t = make_taps<MODE != FWD>(g, query + p * 3);
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
    if (MODE == GGO) { ggx = __ldg(gg + p * 3); ggy = __ldg(gg + p * 3 + 1); ggz = __ldg(gg + p * 3 + 2); }
    float acc[3] = {0.f, 0.f, 0.f};
    for (int d = 0; d < D; d += V) {
      float o[3][V], go[3][V];
      if (MODE == GRAD_QUERY) load_chunk<V>(a + p * 3 * D, d, go);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        AxisView x = view(i, t, g, ggx, ggy, ggz);
        const float* fi = feat + i * plane_elems + d;
        float f[V], du[V], dv[V];
#pragma unroll
        for (int j = 0; j < V; ++j) { f[j] = 0.f; du[j] = 0.f; dv[j] = 0.f; }
        if (PLANE) {
#pragma unroll
          for (int ti = 0; ti < K; ++ti) {
            Vec<V> v[K];
#pragma unroll
            for (int tj = 0; tj < K; ++tj) v[tj] = ldg_vec<V>(fi + ((long long)x.iu[ti] * G + x.iv[tj]) * D);
#pragma unroll
            for (int tj = 0; tj < K; ++tj) {
#pragma unroll
              for (int j = 0; j < V; ++j) {
                if (MODE == FWD) {
                  f[j] += x.cu[ti] * x.cv[tj] * v[tj].v[j];                // lanczos_triplane_feature_cuda.cu:76-79
                } else {
                  du[j] += x.su * x.gu[ti] * x.cv[tj] * v[tj].v[j];        // :159-160
                  dv[j] += x.sv * x.cu[ti] * x.gv[tj] * v[tj].v[j];
                }
              }
            }
          }
        } else {
#pragma unroll
          for (int ti = 0; ti < K; ++ti) {
            Vec<V> v = ldg_vec<V>(fi + (long long)x.iu[ti] * D);
#pragma unroll
            for (int j = 0; j < V; ++j) {
              if (MODE == FWD) f[j] += x.cu[ti] * v.v[j];                    // lanczos_triline_feature_cuda.cu:70-74
              else du[j] += x.su * x.gu[ti] * v.v[j];                        // :143-147
            }
          }
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
          if (MODE == FWD) o[i][j] = f[j];
          else if (MODE == GGO) o[i][j] = PLANE ? (x.ggu * du[j] + x.ggv * dv[j]) : x.ggu * du[j];
          else {
            acc[x.au] += go[i][j] * du[j];
            if (PLANE) acc[x.av] += go[i][j] * dv[j];
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

// One thread per (point, plane | line): 16 (4) vector reductions per channel chunk.
//   SECOND=false: kernel_grad_feature (triplane :208-262, triline :182-228)
//   SECOND=true : kernel_grad_query_grad_feature (triplane :503-560, triline :455-505)
template <bool PLANE, bool SECOND, int V>
__global__ void __launch_bounds__(NDJIR_BLOCK)
scatter_kernel(long long B, float* __restrict__ gf, const float* __restrict__ go_, const float* __restrict__ gg,
               const float* __restrict__ query, GridFrame g, int G, int D) {
  long long stride = (long long)gridDim.x * blockDim.x;
  const long long plane_elems = PLANE ? (long long)G * G * D : (long long)G * D;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < B * 3; w += stride) {
    long long p = w / 3;
    int i = (int)(w - p * 3);
    Taps t = make_taps<SECOND>(g, query + p * 3);
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
    if (SECOND) { ggx = __ldg(gg + p * 3); ggy = __ldg(gg + p * 3 + 1); ggz = __ldg(gg + p * 3 + 2); }
    AxisView x = view(i, t, g, ggx, ggy, ggz);
    const float* grow = go_ + p * 3 * D + i;
    float* gi = gf + i * plane_elems;
    for (int d = 0; d < D; d += V) {
      float o[V];
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = __ldg(grow + (d + j) * 3);
#pragma unroll
      for (int ti = 0; ti < K; ++ti) {
        if (PLANE) {
#pragma unroll
          for (int tj = 0; tj < K; ++tj) {
            float coef = SECOND ? (x.ggu * x.su * (x.gu[ti] * x.cv[tj]) + x.ggv * x.sv * (x.cu[ti] * x.gv[tj]))
                                : x.cu[ti] * x.cv[tj];
            Vec<V> val;
#pragma unroll
            for (int j = 0; j < V; ++j) val.v[j] = o[j] * coef;
            red_vec<V>(gi + ((long long)x.iu[ti] * G + x.iv[tj]) * D + d, val);
          }
        } else {
          float coef = SECOND ? (x.ggu * x.su * x.gu[ti]) : x.cu[ti];
          Vec<V> val;
#pragma unroll
          for (int j = 0; j < V; ++j) val.v[j] = o[j] * coef;
          red_vec<V>(gi + (long long)x.iu[ti] * D + d, val);
        }
      }
    }
  }
}

static bool bad(int G, int D) { return G <= 0 || D <= 0 || (long long)G * G * D * 3 >= (1ll << 40); }

template <bool PLANE, int MODE>
static int launch_gather(long long B, float* out, const float* a, const float* gg, const float* query,
                         const float* feat, int G, int D, const float* mn, const float* mx, bool accum,
                         cudaStream_t st) {
  if (B == 0) return NDJIR_OK;
  if (B < 0 || bad(G, D) || !out || !query || !feat || !mn || !mx) return NDJIR_ERR_ARG;
  GridFrame g = make_frame(G, G, G, mn, mx);
  int V = pick_vec(D, feat, MODE == GRAD_QUERY ? (const void*)a : (const void*)out);
  int grid = grid_for(B);
#define NDJIR_LAUNCH(VV)                                                                                            \
  if (accum) gather_kernel<PLANE, MODE, VV, true><<<grid, NDJIR_BLOCK, 0, st>>>(B, out, a, gg, query, feat, g, G, D); \
  else gather_kernel<PLANE, MODE, VV, false><<<grid, NDJIR_BLOCK, 0, st>>>(B, out, a, gg, query, feat, g, G, D);
  if (V == 4) { NDJIR_LAUNCH(4) } else if (V == 2) { NDJIR_LAUNCH(2) } else { NDJIR_LAUNCH(1) }
#undef NDJIR_LAUNCH
  NDJIR_RETURN_LAST_ERROR();
}

template <bool PLANE, bool SECOND>
static int launch_scatter(long long B, float* gf, const float* go, const float* gg, const float* query, int G,
                          int D, const float* mn, const float* mx, cudaStream_t st) {
  if (B == 0) return NDJIR_OK;
  if (B < 0 || bad(G, D) || !gf || !go || !query || !mn || !mx) return NDJIR_ERR_ARG;
  GridFrame g = make_frame(G, G, G, mn, mx);
  int V = pick_vec(D, gf);
  int grid = grid_for(B * 3);
  if (V == 4) scatter_kernel<PLANE, SECOND, 4><<<grid, NDJIR_BLOCK, 0, st>>>(B, gf, go, gg, query, g, G, D);
  else if (V == 2) scatter_kernel<PLANE, SECOND, 2><<<grid, NDJIR_BLOCK, 0, st>>>(B, gf, go, gg, query, g, G, D);
  else scatter_kernel<PLANE, SECOND, 1><<<grid, NDJIR_BLOCK, 0, st>>>(B, gf, go, gg, query, g, G, D);
  NDJIR_RETURN_LAST_ERROR();
}

}  // namespace lanczos_tpl
}  // namespace ndjir

using namespace ndjir;
using namespace ndjir::lanczos_tpl;

#define NDJIR_DEFINE_LANCZOS_FAMILY(NAME, FWDNAME, PLANE, TABLE_ELEMS)                                            \
  int ndjir_##NAME##_##FWDNAME(long long n, float* output, const float* query, const float* feature, int G,      \
                               int D, const float* min3, const float* max3, int accum, cudaStream_t st) {        \
    return launch_gather<PLANE, FWD>(n, output, nullptr, nullptr, query, feature, G, D, min3, max3, accum != 0,  \
                                     st);                                                                        \
  }                                                                                                              \
  int ndjir_##NAME##_grad_query(long long n, float* grad_query, const float* grad_output, const float* query,    \
                                const float* feature, int G, int D, const float* min3, const float* max3,        \
                                int accum, cudaStream_t st) {                                                    \
    if (n > 0 && !grad_output) return NDJIR_ERR_ARG;                                                             \
    return launch_gather<PLANE, GRAD_QUERY>(n, grad_query, grad_output, nullptr, query, feature, G, D, min3,     \
                                            max3, accum != 0, st);                                               \
  }                                                                                                              \
  int ndjir_##NAME##_grad_feature(long long n, float* grad_feature, const float* grad_output,                    \
                                  const float* query, int G, int D, const float* min3, const float* max3,        \
                                  int accum, cudaStream_t st) {                                                  \
    if (bad(G, D) || !grad_feature) return NDJIR_ERR_ARG;                                                        \
    if (!accum) fill_zero(grad_feature, (long long)(TABLE_ELEMS), st);                                           \
    return launch_scatter<PLANE, false>(n, grad_feature, grad_output, nullptr, query, G, D, min3, max3, st);     \
  }                                                                                                              \
  int ndjir_##NAME##_grad_query_grad_grad_output(long long n, float* grad_grad_output,                           \
                                                 const float* grad_grad_query, const float* query,               \
                                                 const float* feature, int G, int D, const float* min3,          \
                                                 const float* max3, int accum, cudaStream_t st) {                \
    if (n > 0 && !grad_grad_query) return NDJIR_ERR_ARG;                                                         \
    return launch_gather<PLANE, GGO>(n, grad_grad_output, nullptr, grad_grad_query, query, feature, G, D, min3,  \
                                     max3, accum != 0, st);                                                      \
  }                                                                                                              \
  int ndjir_##NAME##_grad_query_grad_feature(long long n, float* grad_feature, const float* grad_grad_query,     \
                                             const float* grad_output, const float* query, int G, int D,         \
                                             const float* min3, const float* max3, cudaStream_t st) {            \
    if (n > 0 && !grad_grad_query) return NDJIR_ERR_ARG;                                                         \
    return launch_scatter<PLANE, true>(n, grad_feature, grad_output, grad_grad_query, query, G, D, min3, max3,   \
                                       st);                                                                      \
  }

extern "C" {
NDJIR_DEFINE_LANCZOS_FAMILY(lanczos_triplane, query_on_triplane, true, 3ll * G * G * D)
NDJIR_DEFINE_LANCZOS_FAMILY(lanczos_triline, query_on_triline, false, 3ll * G * D)
}
