// Per-sample total-variation loss on voxel / triplane / triline grids, forward + scatter backward (sm_100a).
//
// Replaces csrc/grid_feature/total_variation_loss_cuda.cu (:33,:111),
// total_variation_loss_on_triplane_cuda.cu (:31,:101), total_variation_loss_on_triline_cuda.cu (:31,:98).
// forward : sqrt(sum of squared forward differences at the sample's cell corner)
// backward: gf[neighbour] += ograd * delta * rsqrt(sum delta^2 + 1e-12) (DOUBLE intermediate, :161);
//           sym_backward also adds -(sum) at the corner itself.  Always accumulates (q7).
// Thread mapping: one thread per point, all channels, vector loads/reductions (reference: one thread per
// (point,channel[,plane])).  HBM-bound: voxel fwd 12 + 4D + 4*4D bytes/pt = 92 B/pt at D=4.
#include "grid_common.cuh"
#include "../../include/ndjir_b200.h"

namespace ndjir {
namespace tv {

struct Strides { unsigned sx, sy, sz; };

template <int V, bool BACKWARD, bool SYM>
__global__ void __launch_bounds__(NDJIR_BLOCK)
voxel_kernel(long long B, float* __restrict__ out, const float* __restrict__ go_, const float* __restrict__ query,
             const float* __restrict__ feat, GridFrame g, Strides s, int D) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < B; p += stride) {
    const float* q = query + p * 3;
    Cell c = make_cell(g, __ldg(q), __ldg(q + 1), __ldg(q + 2));
    unsigned i000 = c.x0 * s.sx + c.y0 * s.sy + c.z0 * s.sz;
    unsigned i001 = c.x0 * s.sx + c.y0 * s.sy + c.z1 * s.sz;
    unsigned i010 = c.x0 * s.sx + c.y1 * s.sy + c.z0 * s.sz;
    unsigned i100 = c.x1 * s.sx + c.y0 * s.sy + c.z0 * s.sz;
    for (int d = 0; d < D; d += V) {
      Vec<V> f000 = ldg_vec<V>(feat + i000 + d), f001 = ldg_vec<V>(feat + i001 + d);
      Vec<V> f010 = ldg_vec<V>(feat + i010 + d), f100 = ldg_vec<V>(feat + i100 + d);
      if (!BACKWARD) {
        Vec<V> o;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float dx = f100.v[j] - f000.v[j], dy = f010.v[j] - f000.v[j], dz = f001.v[j] - f000.v[j];
          float dx2 = dx * dx, dy2 = dy * dy, dz2 = dz * dz;
          o.v[j] = sqrtf(dx2 + dy2 + dz2);
        }
        st_vec<V>(out + p * D + d, o);
      } else {
        Vec<V> go = ldg_vec<V>(go_ + p * D + d);
        Vec<V> g100, g010, g001, g000;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float dx = f100.v[j] - f000.v[j], dy = f010.v[j] - f000.v[j], dz = f001.v[j] - f000.v[j];
          float dx2 = dx * dx, dy2 = dy * dy, dz2 = dz * dz;
          double common = go.v[j] * rsqrt(dx2 + dy2 + dz2 + 1e-12);
          double a = common * dx, b = common * dy, cc = common * dz;
          g100.v[j] = (float)a; g010.v[j] = (float)b; g001.v[j] = (float)cc;
          g000.v[j] = (float)(-(a + b + cc));
        }
        red_vec<V>(out + i100 + d, g100);
        red_vec<V>(out + i010 + d, g010);
        red_vec<V>(out + i001 + d, g001);
        if (SYM) red_vec<V>(out + i000 + d, g000);
      }
    }
  }
}

// Triplane (PLANE=true) / triline (PLANE=false); values in the (B, D*3) layout, c = d*3 + i.
template <bool PLANE, bool BACKWARD, bool SYM>
__global__ void __launch_bounds__(NDJIR_BLOCK)
tpl_kernel(long long B, float* __restrict__ out, const float* __restrict__ go_, const float* __restrict__ query,
           const float* __restrict__ feat, GridFrame g, int G, int D) {
  long long stride = (long long)gridDim.x * blockDim.x;
  const long long plane_elems = PLANE ? (long long)G * G * D : (long long)G * D;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < B; p += stride) {
    const float* q = query + p * 3;
    Cell c = make_cell(g, __ldg(q), __ldg(q + 1), __ldg(q + 2));
    const unsigned lo[3] = {c.x0, c.y0, c.z0}, hi[3] = {c.x1, c.y1, c.z1};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int au = i, av = (i + 1) % 3;
      long long i00, i10, i01;
      if (PLANE) {
        i00 = ((long long)lo[au] * G + lo[av]) * D;
        i10 = ((long long)hi[au] * G + lo[av]) * D;
        i01 = ((long long)lo[au] * G + hi[av]) * D;
      } else {
        i00 = (long long)lo[au] * D; i10 = (long long)hi[au] * D; i01 = i00;
      }
      const float* fi = feat + i * plane_elems;
      float* gi = out + i * plane_elems;  // only used in backward
      for (int d = 0; d < D; ++d) {
        float f00 = __ldg(fi + i00 + d), f10 = __ldg(fi + i10 + d);
        float du = f10 - f00, du2 = du * du;
        float dv = 0.f, dv2 = 0.f;
        if (PLANE) { dv = __ldg(fi + i01 + d) - f00; dv2 = dv * dv; }
        long long oi = p * 3 * D + (long long)d * 3 + i;
        if (!BACKWARD) {
          out[oi] = PLANE ? sqrtf(du2 + dv2) : sqrtf(du2);
        } else {
          float go = __ldg(go_ + oi);
          double common = PLANE ? go * rsqrt(du2 + dv2 + 1e-12) : go * rsqrt(du2 + 1e-12);
          double a = common * du, b = common * dv;
          atomicAdd(gi + i10 + d, (float)a);
          if (PLANE) atomicAdd(gi + i01 + d, (float)b);
          if (SYM) atomicAdd(gi + i00 + d, PLANE ? (float)(-(a + b)) : (float)(-a));
        }
      }
    }
  }
}

static bool bad_grid(const int* G, int D) {
  if (!G || D <= 0 || G[0] <= 0 || G[1] <= 0 || G[2] <= 0) return true;
  return (long long)G[0] * G[1] * G[2] * D >= (1ll << 32);
}

template <bool BACKWARD>
static int launch_voxel(long long B, float* out, const float* go, const float* query, const float* feat,
                        const int* G, int D, const float* mn, const float* mx, bool sym, cudaStream_t st) {
  if (B == 0) return NDJIR_OK;
  if (B < 0 || bad_grid(G, D) || !out || !query || !feat || !mn || !mx || (BACKWARD && !go)) return NDJIR_ERR_ARG;
  GridFrame g = make_frame(G[0], G[1], G[2], mn, mx);
  Strides s;
  s.sx = (unsigned)G[1] * G[2] * D; s.sy = (unsigned)G[2] * D; s.sz = (unsigned)D;
  int V = pick_vec(D, feat, out, go);
  int grid = grid_for(B);
#define NDJIR_LAUNCH(VV)                                                                                      \
  if (!BACKWARD) voxel_kernel<VV, false, false><<<grid, NDJIR_BLOCK, 0, st>>>(B, out, go, query, feat, g, s, D); \
  else if (sym) voxel_kernel<VV, true, true><<<grid, NDJIR_BLOCK, 0, st>>>(B, out, go, query, feat, g, s, D);   \
  else voxel_kernel<VV, true, false><<<grid, NDJIR_BLOCK, 0, st>>>(B, out, go, query, feat, g, s, D);
  if (V == 4) { NDJIR_LAUNCH(4) } else if (V == 2) { NDJIR_LAUNCH(2) } else { NDJIR_LAUNCH(1) }
#undef NDJIR_LAUNCH
  NDJIR_RETURN_LAST_ERROR();
}

template <bool PLANE, bool BACKWARD>
static int launch_tpl(long long B, float* out, const float* go, const float* query, const float* feat, int G,
                      int D, const float* mn, const float* mx, bool sym, cudaStream_t st) {
  if (B == 0) return NDJIR_OK;
  if (B < 0 || G <= 0 || D <= 0 || !out || !query || !feat || !mn || !mx || (BACKWARD && !go)) return NDJIR_ERR_ARG;
  GridFrame g = make_frame(G, G, G, mn, mx);
  int grid = grid_for(B);
  if (!BACKWARD) tpl_kernel<PLANE, false, false><<<grid, NDJIR_BLOCK, 0, st>>>(B, out, go, query, feat, g, G, D);
  else if (sym) tpl_kernel<PLANE, true, true><<<grid, NDJIR_BLOCK, 0, st>>>(B, out, go, query, feat, g, G, D);
  else tpl_kernel<PLANE, true, false><<<grid, NDJIR_BLOCK, 0, st>>>(B, out, go, query, feat, g, G, D);
  NDJIR_RETURN_LAST_ERROR();
}

}  // namespace tv
}  // namespace ndjir

using namespace ndjir;
using namespace ndjir::tv;

extern "C" {

int ndjir_tv_loss_on_voxel(long long n_points, float* output, const float* query, const float* feature,
                           const int* grid_sizes, int D, const float* min3, const float* max3,
                           cudaStream_t stream) {
  return launch_voxel<false>(n_points, output, nullptr, query, feature, grid_sizes, D, min3, max3, false, stream);
}

int ndjir_tv_loss_on_voxel_backward(long long n_points, float* grad_feature, const float* grad_output,
                                    const float* query, const float* feature, const int* grid_sizes, int D,
                                    const float* min3, const float* max3, int sym_backward,
                                    cudaStream_t stream) {
  return launch_voxel<true>(n_points, grad_feature, grad_output, query, feature, grid_sizes, D, min3, max3,
                            sym_backward != 0, stream);
}

int ndjir_tv_loss_on_triplane(long long n_points, float* output, const float* query, const float* feature, int G,
                              int D, const float* min3, const float* max3, cudaStream_t stream) {
  return launch_tpl<true, false>(n_points, output, nullptr, query, feature, G, D, min3, max3, false, stream);
}

int ndjir_tv_loss_on_triplane_backward(long long n_points, float* grad_feature, const float* grad_output,
                                       const float* query, const float* feature, int G, int D,
                                       const float* min3, const float* max3, int sym_backward,
                                       cudaStream_t stream) {
  return launch_tpl<true, true>(n_points, grad_feature, grad_output, query, feature, G, D, min3, max3,
                                sym_backward != 0, stream);
}

int ndjir_tv_loss_on_triline(long long n_points, float* output, const float* query, const float* feature, int G,
                             int D, const float* min3, const float* max3, cudaStream_t stream) {
  return launch_tpl<false, false>(n_points, output, nullptr, query, feature, G, D, min3, max3, false, stream);
}

int ndjir_tv_loss_on_triline_backward(long long n_points, float* grad_feature, const float* grad_output,
                                      const float* query, const float* feature, int G, int D, const float* min3,
                                      const float* max3, int sym_backward, cudaStream_t stream) {
  return launch_tpl<false, true>(n_points, grad_feature, grad_output, query, feature, G, D, min3, max3,
                                 sym_backward != 0, stream);
}

}  // extern "C"
