// Trilinear voxel-grid feature query, first- and second-order backward (sm_100a).
//
// Replaces the reference module csrc/grid_feature/voxel_feature_cuda.cu (8 exports, :844-863).
// Layout: query (B,3) fp32; feature (Gx,Gy,Gz,D) fp32 channel-last; output (B,D).
// Thread mapping: ONE THREAD PER POINT (the reference uses one thread per (point,channel) and
// re-derives the cell D times).  The D channels of a corner are fetched with one 16-byte (D%4==0),
// 8-byte (D%2==0) or 4-byte read-only load; 8 independent gathers are in flight per thread.
// Scatter kernels issue one vector reduction (REDG.E.ADD.F32x4) per corner, optionally warp-aggregated.
// Roofline: HBM-bound gather/scatter; algorithmic bytes per point at D=4: fwd 12+16+8*16 = 156 B,
// grad_feature 12+16+2*8*16 = 284 B.
#include "grid_common.cuh"
#include "gemm.cuh"
#include "gemm_h.cuh"
#include "voxel_binned.cuh"
#include "../../include/ndjir_b200.h"

namespace ndjir {

extern int g_hash_coarse_private;
int g_scatter_aggregate = 0;

namespace voxel {

struct Strides { unsigned sx, sy, sz; };

template <int V>
struct Corners {
  Vec<V> f000, f001, f010, f011, f100, f101, f110, f111;
};

__device__ __forceinline__ unsigned fidx(const Strides& s, unsigned x, unsigned y, unsigned z) {
  return x * s.sx + y * s.sy + z * s.sz;
}

template <int V>
__device__ __forceinline__ Corners<V> gather(const float* __restrict__ feat, const Strides& s, const Cell& c,
                                             int d) {
  Corners<V> k;
  k.f000 = ldg_vec<V>(feat + fidx(s, c.x0, c.y0, c.z0) + d);
  k.f001 = ldg_vec<V>(feat + fidx(s, c.x0, c.y0, c.z1) + d);
  k.f010 = ldg_vec<V>(feat + fidx(s, c.x0, c.y1, c.z0) + d);
  k.f011 = ldg_vec<V>(feat + fidx(s, c.x0, c.y1, c.z1) + d);
  k.f100 = ldg_vec<V>(feat + fidx(s, c.x1, c.y0, c.z0) + d);
  k.f101 = ldg_vec<V>(feat + fidx(s, c.x1, c.y0, c.z1) + d);
  k.f110 = ldg_vec<V>(feat + fidx(s, c.x1, c.y1, c.z0) + d);
  k.f111 = ldg_vec<V>(feat + fidx(s, c.x1, c.y1, c.z1) + d);
  return k;
}

// d f/d(q) for one channel, voxel_feature_cuda.cu:181-199 (compute_grad without the ograd factor).
__device__ __forceinline__ float dterm(float scale, float a0, float a1, float b0, float b1, float d00, float d01,
                                       float d10, float d11) {
  return scale * (a0 * b0 * d00 + a0 * b1 * d01 + a1 * b0 * d10 + a1 * b1 * d11);
}

enum Mode { FWD = 0, GRAD_QUERY = 1, GGO = 2, GQ_GQ = 3 };

// Gather-type kernels.  `a` and `b` are mode-dependent per-point inputs:
//   FWD        : out (B,D)  = interp(feature)                                    [accum optional]
//   GRAD_QUERY : out (B,3)  = sum_d a[b,d] * d f_d/dq            a = grad_output [accum optional]
//   GGO        : out (B,D)  = b[b,:] . d f_d/dq                  b = grad_grad_query [accum optional]
//   GQ_GQ      : out (B,3) += cross second derivatives           a = grad_output, b = grad_grad_query
template <int MODE, int V, bool ACCUM>
__global__ void __launch_bounds__(NDJIR_BLOCK)
gather_kernel(long long B, float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b,
              const float* __restrict__ query, const float* __restrict__ feat, GridFrame g, Strides s, int D) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < B; p += stride) {
    const float* q = query + p * 3;
    Cell c = make_cell(g, __ldg(q), __ldg(q + 1), __ldg(q + 2));
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
    if (MODE == GGO || MODE == GQ_GQ) {
      ggx = __ldg(b + p * 3); ggy = __ldg(b + p * 3 + 1); ggz = __ldg(b + p * 3 + 2);
    }
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int d = 0; d < D; d += V) {
      Corners<V> k = gather<V>(feat, s, c, d);
      if (MODE == FWD) {
        Vec<V> o;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          // same expression shape as voxel_feature_cuda.cu:87-94
          o.v[j] = c.p0 * c.q0 * c.r0 * k.f000.v[j] + c.p0 * c.q0 * c.r1 * k.f001.v[j] +
                   c.p0 * c.q1 * c.r0 * k.f010.v[j] + c.p0 * c.q1 * c.r1 * k.f011.v[j] +
                   c.p1 * c.q0 * c.r0 * k.f100.v[j] + c.p1 * c.q0 * c.r1 * k.f101.v[j] +
                   c.p1 * c.q1 * c.r0 * k.f110.v[j] + c.p1 * c.q1 * c.r1 * k.f111.v[j];
        }
        float* op = out + p * D + d;
        if (ACCUM) {
          Vec<V> prev = ld_vec<V>(op);
#pragma unroll
          for (int j = 0; j < V; ++j) o.v[j] += prev.v[j];
        }
        st_vec<V>(op, o);
      } else {
        Vec<V> av;
        if (MODE == GRAD_QUERY || MODE == GQ_GQ) av = ldg_vec<V>(a + p * D + d);
        Vec<V> o;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float f000 = k.f000.v[j], f001 = k.f001.v[j], f010 = k.f010.v[j], f011 = k.f011.v[j];
          float f100 = k.f100.v[j], f101 = k.f101.v[j], f110 = k.f110.v[j], f111 = k.f111.v[j];
          if (MODE == GRAD_QUERY || MODE == GGO) {
            float gx = dterm(c.sx, c.q0, c.q1, c.r0, c.r1, f100 - f000, f101 - f001, f110 - f010, f111 - f011);
            float gy = dterm(c.sy, c.p0, c.p1, c.r0, c.r1, f010 - f000, f011 - f001, f110 - f100, f111 - f101);
            float gz = dterm(c.sz, c.p0, c.p1, c.q0, c.q1, f001 - f000, f011 - f010, f101 - f100, f111 - f110);
            if (MODE == GRAD_QUERY) {
              ax += av.v[j] * gx; ay += av.v[j] * gy; az += av.v[j] * gz;
            } else {
              o.v[j] = ggx * gx + ggy * gy + ggz * gz;
            }
          } else {  // GQ_GQ, voxel_feature_cuda.cu:505-516
            float go = av.v[j];
            float ti = go * g.sy * g.sz * (c.p0 * (f000 - f001 - f010 + f011) + c.p1 * (f100 - f101 - f110 + f111));
            float tj = go * g.sx * g.sz * (c.q0 * (f000 - f001 - f100 + f101) + c.q1 * (f010 - f011 - f110 + f111));
            float tk = go * g.sx * g.sy * (c.r0 * (f000 - f010 - f100 + f110) + c.r1 * (f001 - f011 - f101 + f111));
            ax += ggy * tk + ggz * tj;
            ay += ggz * ti + ggx * tk;
            az += ggx * tj + ggy * ti;
          }
        }
        if (MODE == GGO) {
          float* op = out + p * D + d;
          if (ACCUM) {
            Vec<V> prev = ld_vec<V>(op);
#pragma unroll
            for (int j = 0; j < V; ++j) o.v[j] += prev.v[j];
          }
          st_vec<V>(op, o);
        }
      }
    }
    if (MODE == GRAD_QUERY || MODE == GQ_GQ) {
      float* op = out + p * 3;
      if (ACCUM) { ax += op[0]; ay += op[1]; az += op[2]; }
      op[0] = ax; op[1] = ay; op[2] = az;
    }
  }
}

// Forward gather with FOUR LANES PER POINT (D % 4 == 0): lane (cx, cy) fetches the two z-neighbour cells of its
// (x, y) column - 32 contiguous bytes - and the partial sums are combined with two shuffles.  Measured on the
// micro-benchmark shape (2^24 uniform points, 512^3 x 4): 1.57 ms vs 2.08 ms for one thread per point and 1.66 ms for
// the reference's thread per (point, channel) mapping (profiles/r1_exp_voxel_variants.txt): four times as many
// independent 16-byte requests per point are in flight and each warp request touches 8 points instead of 32.
template <bool ACCUM>
__global__ void __launch_bounds__(NDJIR_BLOCK)
gather4_kernel(long long B, float* __restrict__ out, const float* __restrict__ query, const float* __restrict__ feat,
               GridFrame g, Strides s, int D) {
  const int sub = threadIdx.x & 3;
  const int cx = sub >> 1, cy = sub & 1;
  long long stride = (long long)gridDim.x * blockDim.x / 4;
  long long rounds = (B + stride - 1) / stride;
  long long p0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 4;
  for (long long r = 0; r < rounds; ++r) {
    long long p = p0 + r * stride;
    bool active = p < B;
    long long pc = active ? p : B - 1;
    const float* q = query + pc * 3;
    Cell c = make_cell(g, __ldg(q), __ldg(q + 1), __ldg(q + 2));
    unsigned base = (cx ? c.x1 : c.x0) * s.sx + (cy ? c.y1 : c.y0) * s.sy;
    float wxy = (cx ? c.p1 : c.p0) * (cy ? c.q1 : c.q0);
    float w0 = wxy * c.r0, w1 = wxy * c.r1;
    for (int d = 0; d < D; d += 4) {
      float4 f0 = __ldg(reinterpret_cast<const float4*>(feat + base + c.z0 * s.sz + d));
      float4 f1 = __ldg(reinterpret_cast<const float4*>(feat + base + c.z1 * s.sz + d));
      float4 o = make_float4(w0 * f0.x + w1 * f1.x, w0 * f0.y + w1 * f1.y, w0 * f0.z + w1 * f1.z, w0 * f0.w + w1 * f1.w);
#pragma unroll
      for (int m = 1; m < 4; m <<= 1) {
        o.x += __shfl_xor_sync(0xffffffffu, o.x, m); o.y += __shfl_xor_sync(0xffffffffu, o.y, m);
        o.z += __shfl_xor_sync(0xffffffffu, o.z, m); o.w += __shfl_xor_sync(0xffffffffu, o.w, m);
      }
      if (active && sub == 0) {
        float4* op = reinterpret_cast<float4*>(out + p * D + d);
        if (ACCUM) { float4 pv = *op; o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w; }
        *op = o;
      }
    }
  }
}

// Scatter-type kernels into the feature gradient (always accumulate; zero-fill is the host's job).
//   SECOND=false : gf[corner] += go * p*q*r                               (kernel_grad_feature :231-288)
//   SECOND=true  : gf[corner] += go * (ggx sx a + ggy sy b + ggz sz c)    (kernel_grad_query_grad_feature :549-614)
template <bool SECOND, int V, bool AGG>
__global__ void __launch_bounds__(NDJIR_BLOCK)
scatter_kernel(long long B, float* __restrict__ gf, const float* __restrict__ go, const float* __restrict__ gg,
               const float* __restrict__ query, GridFrame g, Strides s, int D) {
  long long stride = (long long)gridDim.x * blockDim.x;
  long long start = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // all lanes of a warp iterate together so the warp-aggregated path stays converged
  long long rounds = (B + stride - 1) / stride;
  for (long long r = 0; r < rounds; ++r) {
    long long p = start + r * stride;
    bool active = p < B;
    long long pc = active ? p : (B - 1);
    const float* q = query + pc * 3;
    Cell c = make_cell(g, __ldg(q), __ldg(q + 1), __ldg(q + 2));
    float w[8][3];  // per corner: weight (first order) or the three signed partial weights
    unsigned idx[8];
    {
      const unsigned xs[2] = {c.x0, c.x1}, ys[2] = {c.y0, c.y1}, zs[2] = {c.z0, c.z1};
      const float ps[2] = {c.p0, c.p1}, qs[2] = {c.q0, c.q1}, rs[2] = {c.r0, c.r1};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int cx = (k >> 2) & 1, cy = (k >> 1) & 1, cz = k & 1;
        idx[k] = fidx(s, xs[cx], ys[cy], zs[cz]);
        if (!SECOND) {
          w[k][0] = ps[cx] * qs[cy] * rs[cz];
        } else {
          w[k][0] = (cx ? 1.f : -1.f) * qs[cy] * rs[cz];
          w[k][1] = (cy ? 1.f : -1.f) * ps[cx] * rs[cz];
          w[k][2] = (cz ? 1.f : -1.f) * ps[cx] * qs[cy];
        }
      }
    }
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
    if (SECOND) {
      ggx = __ldg(gg + pc * 3) * c.sx; ggy = __ldg(gg + pc * 3 + 1) * c.sy; ggz = __ldg(gg + pc * 3 + 2) * c.sz;
    }
    for (int d = 0; d < D; d += V) {
      Vec<V> o = ldg_vec<V>(go + pc * D + d);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float coef = SECOND ? (ggx * w[k][0] + ggy * w[k][1] + ggz * w[k][2]) : w[k][0];
        Vec<V> val;
#pragma unroll
        for (int j = 0; j < V; ++j) val.v[j] = o.v[j] * coef;
        float* dst = gf + idx[k] + d;
        if (AGG) {
          warp_agg_red<V>(dst, (unsigned long long)(idx[k] + d), val, active);
        } else if (active) {
          red_vec<V>(dst, val);
        }
      }
    }
  }
}

// Scatter with EIGHT LANES PER POINT (one corner each): measured 3.29 ms vs 3.56 ms (one thread per point) and
// 3.42 ms (the reference's thread per (point, channel) with scalar atomics) at 2^24 uniform points.
template <bool SECOND, int V>
__global__ void __launch_bounds__(NDJIR_BLOCK)
scatter8_kernel(long long B, float* __restrict__ gf, const float* __restrict__ go, const float* __restrict__ gg,
                const float* __restrict__ query, GridFrame g, Strides s, int D) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < B * 8; w += stride) {
    long long p = w >> 3;
    int k = (int)(w & 7);
    int cx = (k >> 2) & 1, cy = (k >> 1) & 1, cz = k & 1;
    const float* q = query + p * 3;
    Cell c = make_cell(g, __ldg(q), __ldg(q + 1), __ldg(q + 2));
    float pw = cx ? c.p1 : c.p0, qw = cy ? c.q1 : c.q0, rw = cz ? c.r1 : c.r0;
    float coef;
    if (!SECOND) {
      coef = pw * qw * rw;
    } else {
      float ggx = __ldg(gg + p * 3) * c.sx, ggy = __ldg(gg + p * 3 + 1) * c.sy, ggz = __ldg(gg + p * 3 + 2) * c.sz;
      coef = ggx * ((cx ? 1.f : -1.f) * qw * rw) + ggy * ((cy ? 1.f : -1.f) * pw * rw) + ggz * ((cz ? 1.f : -1.f) * pw * qw);
    }
    float* dst = gf + fidx(s, cx ? c.x1 : c.x0, cy ? c.y1 : c.y0, cz ? c.z1 : c.z0);
    for (int d = 0; d < D; d += V) {
      Vec<V> o = ldg_vec<V>(go + p * D + d);
#pragma unroll
      for (int j = 0; j < V; ++j) o.v[j] *= coef;
      red_vec<V>(dst + d, o);
    }
  }
}

static Strides make_strides(const int* G, int D) {
  Strides s;
  s.sx = (unsigned)G[1] * (unsigned)G[2] * (unsigned)D;
  s.sy = (unsigned)G[2] * (unsigned)D;
  s.sz = (unsigned)D;
  return s;
}

static bool bad_grid(const int* G, int D) {
  if (!G || D <= 0 || G[0] <= 0 || G[1] <= 0 || G[2] <= 0) return true;
  return (long long)G[0] * G[1] * G[2] * D >= (1ll << 32);  // uint32 flat index like the reference
}

template <int MODE>
static int launch_gather(long long B, float* out, const float* a, const float* b, const float* query,
                         const float* feat, const int* G, int D, const float* mn, const float* mx, bool accum,
                         cudaStream_t st, int interp = INTERP_LINEAR) {
  if (B == 0) return NDJIR_OK;
  if (B < 0 || bad_grid(G, D) || !out || !query || !feat || !mn || !mx) return NDJIR_ERR_ARG;
  GridFrame g = make_frame(G[0], G[1], G[2], mn, mx);
  g.interp = interp;
  Strides s = make_strides(G, D);
  const void* vec_out = (MODE == FWD || MODE == GGO) ? out : nullptr;
  const void* vec_a = (MODE == GRAD_QUERY || MODE == GQ_GQ) ? a : nullptr;
  int V = pick_vec(D, feat, vec_out, vec_a);
  int grid = grid_for(B);
  if (MODE == FWD && interp == INTERP_LINEAR && voxel_binned::worthwhile(B, G, D)) {
    long long wsb = voxel_binned::workspace_bytes(B);
    if (void* ws = voxel_binned::scratch_alloc(wsb, st)) {
      int rc = voxel_binned::query(B, out, query, feat, G, D, mn, mx, accum, ws, wsb, st);
      voxel_binned::scratch_free(ws, st);
      return rc;
    }
  }
  if (MODE == FWD && V == 4) {
    int grid4 = grid_for(B * 4);
    if (accum) gather4_kernel<true><<<grid4, NDJIR_BLOCK, 0, st>>>(B, out, query, feat, g, s, D);
    else gather4_kernel<false><<<grid4, NDJIR_BLOCK, 0, st>>>(B, out, query, feat, g, s, D);
    NDJIR_RETURN_LAST_ERROR();
  }
#define NDJIR_LAUNCH(VV)                                                                               \
  if (accum) gather_kernel<MODE, VV, true><<<grid, NDJIR_BLOCK, 0, st>>>(B, out, a, b, query, feat, g, s, D); \
  else gather_kernel<MODE, VV, false><<<grid, NDJIR_BLOCK, 0, st>>>(B, out, a, b, query, feat, g, s, D);
  if (V == 4) { NDJIR_LAUNCH(4) } else if (V == 2) { NDJIR_LAUNCH(2) } else { NDJIR_LAUNCH(1) }
#undef NDJIR_LAUNCH
  NDJIR_RETURN_LAST_ERROR();
}

template <bool SECOND>
static int launch_scatter(long long B, float* gf, const float* go, const float* gg, const float* query,
                          const int* G, int D, const float* mn, const float* mx, cudaStream_t st,
                          int interp = INTERP_LINEAR) {
  if (B == 0) return NDJIR_OK;
  if (B < 0 || bad_grid(G, D) || !gf || !go || !query || !mn || !mx) return NDJIR_ERR_ARG;
  GridFrame g = make_frame(G[0], G[1], G[2], mn, mx);
  g.interp = interp;
  Strides s = make_strides(G, D);
  if (interp == INTERP_LINEAR && voxel_binned::worthwhile(B, G, D)) {
    long long wsb = voxel_binned::workspace_bytes(B);
    if (void* ws = voxel_binned::scratch_alloc(wsb, st)) {
      int rc = voxel_binned::scatter(SECOND, B, gf, go, gg, query, G, D, mn, mx, ws, wsb, st);
      voxel_binned::scratch_free(ws, st);
      return rc;
    }
  }
  int V = pick_vec(D, gf, go);
  int grid = grid_for(B);
  bool agg = g_scatter_aggregate != 0;
  if (!agg) {
    int grid8 = grid_for(B * 8);
    if (V == 4) scatter8_kernel<SECOND, 4><<<grid8, NDJIR_BLOCK, 0, st>>>(B, gf, go, gg, query, g, s, D);
    else if (V == 2) scatter8_kernel<SECOND, 2><<<grid8, NDJIR_BLOCK, 0, st>>>(B, gf, go, gg, query, g, s, D);
    else scatter8_kernel<SECOND, 1><<<grid8, NDJIR_BLOCK, 0, st>>>(B, gf, go, gg, query, g, s, D);
    NDJIR_RETURN_LAST_ERROR();
  }
#define NDJIR_LAUNCH(VV)                                                                                     \
  if (agg) scatter_kernel<SECOND, VV, true><<<grid, NDJIR_BLOCK, 0, st>>>(B, gf, go, gg, query, g, s, D);    \
  else scatter_kernel<SECOND, VV, false><<<grid, NDJIR_BLOCK, 0, st>>>(B, gf, go, gg, query, g, s, D);
  if (V == 4) { NDJIR_LAUNCH(4) } else if (V == 2) { NDJIR_LAUNCH(2) } else { NDJIR_LAUNCH(1) }
#undef NDJIR_LAUNCH
  NDJIR_RETURN_LAST_ERROR();
}

}  // namespace voxel
}  // namespace ndjir

using namespace ndjir;
using namespace ndjir::voxel;

extern "C" {

int ndjir_set_option(const char* key, int value) {
  if (!key) return NDJIR_ERR_ARG;
  struct Opt { const char* name; int* target; };
  const Opt opts[] = {
      {"scatter_aggregate", &g_scatter_aggregate},          // 1: warp-aggregated scatter reductions (coherent rays)
      {"mlp_tensor_cores", &ndjir::gemm::g_mlp_tensor_cores},  // 0: fp32 FFMA parity path for every MLP product
      {"mlp_cta_pair", &ndjir::gemm::g_mlp_cta_pair},       // 1: tcgen05 cta_group::2 product kernel (measured slower)
      {"mlp_presplit", &ndjir::gemm::g_mlp_presplit},       // 0: ignore caller-supplied lo parts of the weight operand
      {"mlp_fused_colsum", &ndjir::gemm::g_mlp_fused_colsum},  // 0: bias gradients by a separate column-sum pass
      {"mlp_dbg", &ndjir::gemm::g_mlp_dbg},                 // profiling switches of the tcgen05 kernel
      {"mlp_mask_hi", &ndjir::gemm::g_mlp_mask_hi},
      {"mlp_h_dbg", &ndjir::gemmh::g_h_dbg},                // profiling switches of the split-fp16 tcgen05 kernel
      {"mlp_h_chain", &ndjir::gemmh::g_h_chain},            // 1: the SDF-only network evaluation as one on-chip kernel
      {"mlp_h_tall", &ndjir::gemmh::g_h_tall},              // 0: 128-row work items in the weight-gradient products
      {"mlp_h_tma_epi", &ndjir::gemmh::g_h_tma_epi},        // 0: row-per-lane global accesses in every epilogue
      {"mlp_h_resident", &ndjir::gemmh::g_h_resident},      // 1 (default): resident-weight kernel (csrc/gemm_h3.cu)
      {"mlp_h_pair", &ndjir::gemmh::g_h_pair},              // 1: CTA-pair (cta_group::2) kernel for activation-row products
      {"voxel_binned", &ndjir::g_voxel_binned},             // -1 auto, 0 never, 1 whenever possible (brick-ordered sweeps)
      {"voxel_bin_mb", &ndjir::g_voxel_bin_mb},             // brick size in MiB
      {"hash_coarse_private", &ndjir::g_hash_coarse_private},  // 0 off, 1 batches >= 2^20 points, 2 always
      {"voxel_tma_bx", &ndjir::g_voxel_tma_bx},             // brick extent along x of the TMA sweep (8 or 16)
      {"voxel_tma_l2", &ndjir::g_voxel_tma_l2},             // L2 promotion of its requests (0 none, 1 128 B, 2 256 B)
      {"voxel_tma_dbg", &ndjir::g_voxel_tma_dbg},           // experiment switch of the TMA sweep
      {"voxel_tma", &ndjir::g_voxel_tma},                   // 1: fine-brick TMA sweep for the D = 4 forward gather
      {"voxel_pair256", &ndjir::g_voxel_pair256},           // 256-bit z-pair loads in the binned gather (measured slower)
  };
  for (const Opt& o : opts) {
    int i = 0;
    while (o.name[i] && key[i] == o.name[i]) ++i;
    if (o.name[i] == 0 && key[i] == 0) { *o.target = value; return NDJIR_OK; }
  }
  return NDJIR_ERR_ARG;
}

int ndjir_voxel_query_on_voxel(long long n_points, float* output, const float* query, const float* feature,
                               const int* grid_sizes, int D, const float* min3, const float* max3, int accum,
                               cudaStream_t stream) {
  return launch_gather<FWD>(n_points, output, nullptr, nullptr, query, feature, grid_sizes, D, min3, max3,
                            accum != 0, stream);
}

int ndjir_voxel_grad_query(long long n_points, float* grad_query, const float* grad_output, const float* query,
                           const float* feature, const int* grid_sizes, int D, const float* min3,
                           const float* max3, int accum, cudaStream_t stream) {
  if (n_points > 0 && !grad_output) return NDJIR_ERR_ARG;
  return launch_gather<GRAD_QUERY>(n_points, grad_query, grad_output, nullptr, query, feature, grid_sizes, D,
                                   min3, max3, accum != 0, stream);
}

int ndjir_voxel_grad_feature(long long n_points, float* grad_feature, const float* grad_output,
                             const float* query, const int* grid_sizes, int D, const float* min3,
                             const float* max3, int accum, cudaStream_t stream) {
  if (bad_grid(grid_sizes, D) || !grad_feature) return NDJIR_ERR_ARG;
  if (!accum) fill_zero(grad_feature, (long long)grid_sizes[0] * grid_sizes[1] * grid_sizes[2] * D, stream);
  return launch_scatter<false>(n_points, grad_feature, grad_output, nullptr, query, grid_sizes, D, min3, max3,
                               stream);
}

int ndjir_voxel_grad_query_grad_grad_output(long long n_points, float* grad_grad_output,
                                            const float* grad_grad_query, const float* query,
                                            const float* feature, const int* grid_sizes, int D,
                                            const float* min3, const float* max3, int accum,
                                            cudaStream_t stream) {
  if (n_points > 0 && !grad_grad_query) return NDJIR_ERR_ARG;
  return launch_gather<GGO>(n_points, grad_grad_output, nullptr, grad_grad_query, query, feature, grid_sizes, D,
                            min3, max3, accum != 0, stream);
}

// Always accumulates (the reference ignores `accum` here, voxel_feature_cuda.cu:523-545).
int ndjir_voxel_grad_query_grad_query(long long n_points, float* grad_query, const float* grad_grad_query,
                                      const float* grad_output, const float* query, const float* feature,
                                      const int* grid_sizes, int D, const float* min3, const float* max3,
                                      cudaStream_t stream) {
  if (n_points > 0 && (!grad_grad_query || !grad_output)) return NDJIR_ERR_ARG;
  return launch_gather<GQ_GQ>(n_points, grad_query, grad_output, grad_grad_query, query, feature, grid_sizes, D,
                              min3, max3, true, stream);
}

// Always accumulates (voxel_feature_cuda.cu:616-637).
int ndjir_voxel_grad_query_grad_feature(long long n_points, float* grad_feature, const float* grad_grad_query,
                                        const float* grad_output, const float* query, const int* grid_sizes,
                                        int D, const float* min3, const float* max3, cudaStream_t stream) {
  if (n_points > 0 && !grad_grad_query) return NDJIR_ERR_ARG;
  return launch_scatter<true>(n_points, grad_feature, grad_output, grad_grad_query, query, grid_sizes, D, min3,
                              max3, stream);
}

// grad_feature wrt grad_output: the interpolation of grad_grad_feature (voxel_feature_cuda.cu:642-731).
int ndjir_voxel_grad_feature_grad_grad_output(long long n_points, float* grad_grad_output,
                                              const float* grad_grad_feature, const float* query,
                                              const int* grid_sizes, int D, const float* min3,
                                              const float* max3, int accum, cudaStream_t stream) {
  return launch_gather<FWD>(n_points, grad_grad_output, nullptr, nullptr, query, grad_grad_feature, grid_sizes, D,
                            min3, max3, accum != 0, stream);
}

// grad_feature wrt query: grad_query with grad_grad_feature as the table; always accumulates (:734-838).
int ndjir_voxel_grad_feature_grad_query(long long n_points, float* grad_query, const float* grad_grad_feature,
                                        const float* grad_output, const float* query, const int* grid_sizes,
                                        int D, const float* min3, const float* max3, cudaStream_t stream) {
  if (n_points > 0 && !grad_output) return NDJIR_ERR_ARG;
  return launch_gather<GRAD_QUERY>(n_points, grad_query, grad_output, nullptr, query, grad_grad_feature,
                                   grid_sizes, D, min3, max3, true, stream);
}

// ---- cosine_voxel_feature_cuda (csrc/grid_feature/cosine_voxel_feature_cuda.cu:855-866): same cells and corner
// order, weights 0.5 cos(pi frac) + 0.5 and derivative factor 0.5 pi sin(pi frac) per axis ------------------------------
int ndjir_cosine_voxel_query_on_voxel(long long n_points, float* output, const float* query, const float* feature,
                                      const int* grid_sizes, int D, const float* min3, const float* max3, int accum,
                                      cudaStream_t stream) {
  return launch_gather<FWD>(n_points, output, nullptr, nullptr, query, feature, grid_sizes, D, min3, max3,
                            accum != 0, stream, INTERP_COSINE);
}

int ndjir_cosine_voxel_grad_query(long long n_points, float* grad_query, const float* grad_output, const float* query,
                                  const float* feature, const int* grid_sizes, int D, const float* min3,
                                  const float* max3, int accum, cudaStream_t stream) {
  if (n_points > 0 && !grad_output) return NDJIR_ERR_ARG;
  return launch_gather<GRAD_QUERY>(n_points, grad_query, grad_output, nullptr, query, feature, grid_sizes, D,
                                   min3, max3, accum != 0, stream, INTERP_COSINE);
}

int ndjir_cosine_voxel_grad_feature(long long n_points, float* grad_feature, const float* grad_output,
                                    const float* query, const int* grid_sizes, int D, const float* min3,
                                    const float* max3, int accum, cudaStream_t stream) {
  if (bad_grid(grid_sizes, D) || !grad_feature) return NDJIR_ERR_ARG;
  if (!accum) fill_zero(grad_feature, (long long)grid_sizes[0] * grid_sizes[1] * grid_sizes[2] * D, stream);
  return launch_scatter<false>(n_points, grad_feature, grad_output, nullptr, query, grid_sizes, D, min3, max3,
                               stream, INTERP_COSINE);
}

int ndjir_cosine_voxel_grad_query_grad_grad_output(long long n_points, float* grad_grad_output,
                                                   const float* grad_grad_query, const float* query,
                                                   const float* feature, const int* grid_sizes, int D,
                                                   const float* min3, const float* max3, int accum,
                                                   cudaStream_t stream) {
  if (n_points > 0 && !grad_grad_query) return NDJIR_ERR_ARG;
  return launch_gather<GGO>(n_points, grad_grad_output, nullptr, grad_grad_query, query, feature, grid_sizes, D,
                            min3, max3, accum != 0, stream, INTERP_COSINE);
}

// Always accumulates (cosine_voxel_feature_cuda.cu:626-650 never reads `accum`).
int ndjir_cosine_voxel_grad_query_grad_feature(long long n_points, float* grad_feature, const float* grad_grad_query,
                                               const float* grad_output, const float* query, const int* grid_sizes,
                                               int D, const float* min3, const float* max3, cudaStream_t stream) {
  if (n_points > 0 && !grad_grad_query) return NDJIR_ERR_ARG;
  return launch_scatter<true>(n_points, grad_feature, grad_output, grad_grad_query, query, grid_sizes, D, min3,
                              max3, stream, INTERP_COSINE);
}

}  // extern "C"
