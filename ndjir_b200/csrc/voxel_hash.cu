// Multiresolution hashed voxel features (Instant-NGP style) query + backward (sm_100a).
//
// Replaces csrc/grid_feature/voxel_hash_feature_cuda.cu (6 exports, :985-1001) and
// common_voxel_hash.cuh:24-55.  Per level l: G_l = floor(G0*gf^l), T_l = min(G_l^3, T0), table offset =
// sum_{l'<l} (T*D + (T*D)%8) (sic, q1); every level is hashed with the tiny-cuda-nn primes (:38-48).
// The reference re-evaluates the level table with pow() per thread per level (O(L^2) per point); here the
// table is evaluated ONCE PER CTA into shared memory with the same device expression (same pow(float,int)
// overload, so identical values by construction).
// Thread mapping: one thread per (level, point), n = l*B + b as in the reference, D channels per
// thread fetched with one 8/16-byte load per corner.  Output layout 0 = the reference's (D,L,B)
// (:190, coalesced over b), layout 1 = (B, D*L) with channel c = d*L + l (what the reference's Python
// wrapper produces after its in-place transpose, voxel_hash_feature.py:152-155).
// Roofline: bench table (3.8 MB) is L2-resident -> L2-gather-bound; algorithmic HBM bytes 12 + 4*D*L per
// point (+ table once), L2 gather bytes 8*4*D*L per point.
#include "grid_common.cuh"
#include "lanczos_common.cuh"
#include "../../include/ndjir_b200.h"

namespace ndjir {
int g_hash_coarse_private = 1;   // grad_feature of large batches: coarse levels accumulate in shared memory (2: always)
namespace vhash {

#define NDJIR_HASH_MAX_LEVELS 32

struct HashSpec {
  int G0; float gf; int T0; int L; int D;
  float mnx, mny, mnz, dx, dy, dz;  // min and (max - min)
};

struct LevelTable {
  int G[NDJIR_HASH_MAX_LEVELS];
  int T[NDJIR_HASH_MAX_LEVELS];
  long long off[NDJIR_HASH_MAX_LEVELS + 1];
};

__host__ __device__ inline int force_align(int size, int mod = 8) { return size + size % mod; }

// common_voxel_hash.cuh:31-43, same expression text so host and device agree with the reference build.
__host__ __device__ inline int level_grid_size(int G0, float growth_factor, int level) {
  auto Gf = floor(G0 * pow(growth_factor, level));
  return int(Gf);
}
__host__ __device__ inline int level_table_size(int G, int T0) {
  float Gf = G;
  float t = Gf * Gf * Gf;
  float T = t < float(T0) ? t : float(T0);
  int Ti = int(T);
  return Ti < T0 ? Ti : T0;
}

__device__ __forceinline__ void build_table(LevelTable& tab, const HashSpec& h) {
  if (threadIdx.x < h.L) {
    int G = level_grid_size(h.G0, h.gf, threadIdx.x);
    tab.G[threadIdx.x] = G;
    tab.T[threadIdx.x] = level_table_size(G, h.T0);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long o = 0;
    for (int l = 0; l < h.L; ++l) { tab.off[l] = o; o += force_align(tab.T[l] * h.D); }
    tab.off[h.L] = o;
  }
  __syncthreads();
}

__device__ __forceinline__ unsigned hash3(unsigned x, unsigned y, unsigned z, unsigned T) {
  unsigned r = (x * 1u) ^ (y * 2654435761u) ^ (z * 805459861u);
  return r % T;
}

__device__ __forceinline__ GridFrame level_frame(const HashSpec& h, int G) {
  GridFrame g;
  g.mnx = h.mnx; g.mny = h.mny; g.mnz = h.mnz;
  float g1 = (float)G - 1.f;
  g.gx1 = g.gy1 = g.gz1 = g1;
  g.sx = __fdiv_rn(g1, h.dx); g.sy = __fdiv_rn(g1, h.dy); g.sz = __fdiv_rn(g1, h.dz);
  g.interp = INTERP_LINEAR;
  return g;
}

__device__ __forceinline__ long long out_index(int layout, int d, int l, long long b, int L, long long B, int D) {
  return layout == 0 ? ((long long)d * L + l) * B + b : b * ((long long)D * L) + (long long)d * L + l;
}

__device__ __forceinline__ float dterm(float scale, float a0, float a1, float b0, float b1, float d00, float d01,
                                       float d10, float d11) {
  return scale * (a0 * b0 * d00 + a0 * b1 * d01 + a1 * b0 * d10 + a1 * b1 * d11);
}

// kernel_hash_index (:55-101): the 8 hashed corner indices of one level, written as floats (B,8).
__global__ void __launch_bounds__(NDJIR_BLOCK)
hash_index_kernel(long long B, float* __restrict__ out, const float* __restrict__ query, int G, int T, HashSpec h) {
  GridFrame g = level_frame(h, G);
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += stride) {
    const float* q = query + b * 3;
    Cell c = make_cell(g, __ldg(q), __ldg(q + 1), __ldg(q + 2));
    const unsigned xs[2] = {c.x0, c.x1}, ys[2] = {c.y0, c.y1}, zs[2] = {c.z0, c.z1};
#pragma unroll
    for (int k = 0; k < 8; ++k)
      out[b * 8 + k] = (float)hash3(xs[(k >> 2) & 1], ys[(k >> 1) & 1], zs[k & 1], (unsigned)T);
  }
}

// The level table as the DEVICE evaluates it (pow(float,int) on the device is not guaranteed to be exactly
// rounded, so e.g. 16 * 1.5^4 may floor to 80 where the host's double pow gives 81 - a reference quirk, q2).
__global__ void level_table_kernel(HashSpec h, int* __restrict__ G_out, int* __restrict__ T_out) {
  __shared__ LevelTable tab;
  build_table(tab, h);
  if (threadIdx.x < h.L) { G_out[threadIdx.x] = tab.G[threadIdx.x]; T_out[threadIdx.x] = tab.T[threadIdx.x]; }
}

enum Mode { FWD = 0, GRAD_QUERY = 1, GGO = 2 };

template <int MODE, int V, bool ACCUM>
__global__ void __launch_bounds__(NDJIR_BLOCK)
gather_kernel(long long B, float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ gg,
              const float* __restrict__ query, const float* __restrict__ feat, HashSpec h, int layout, int l_begin) {
  __shared__ LevelTable tab;
  build_table(tab, h);
  const long long N = B * (h.L - l_begin);    // levels below l_begin were handled by fwd_coarse_kernel
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += stride) {
    int l = l_begin + (int)(n / B);
    long long b = n - (long long)(l - l_begin) * B;
    GridFrame g = level_frame(h, tab.G[l]);
    unsigned T = (unsigned)tab.T[l];
    const float* fl = feat + tab.off[l];
    const float* q = query + b * 3;
    Cell c = make_cell(g, __ldg(q), __ldg(q + 1), __ldg(q + 2));
    unsigned i000 = hash3(c.x0, c.y0, c.z0, T) * h.D, i001 = hash3(c.x0, c.y0, c.z1, T) * h.D;
    unsigned i010 = hash3(c.x0, c.y1, c.z0, T) * h.D, i011 = hash3(c.x0, c.y1, c.z1, T) * h.D;
    unsigned i100 = hash3(c.x1, c.y0, c.z0, T) * h.D, i101 = hash3(c.x1, c.y0, c.z1, T) * h.D;
    unsigned i110 = hash3(c.x1, c.y1, c.z0, T) * h.D, i111 = hash3(c.x1, c.y1, c.z1, T) * h.D;
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
    if (MODE == GGO) { ggx = __ldg(gg + b * 3); ggy = __ldg(gg + b * 3 + 1); ggz = __ldg(gg + b * 3 + 2); }
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int d = 0; d < h.D; d += V) {
      Vec<V> f000 = ldg_vec<V>(fl + i000 + d), f001 = ldg_vec<V>(fl + i001 + d);
      Vec<V> f010 = ldg_vec<V>(fl + i010 + d), f011 = ldg_vec<V>(fl + i011 + d);
      Vec<V> f100 = ldg_vec<V>(fl + i100 + d), f101 = ldg_vec<V>(fl + i101 + d);
      Vec<V> f110 = ldg_vec<V>(fl + i110 + d), f111 = ldg_vec<V>(fl + i111 + d);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        long long oi = out_index(layout, d + j, l, b, h.L, B, h.D);
        if (MODE == FWD) {
          float f = c.p0 * c.q0 * c.r0 * f000.v[j] + c.p0 * c.q0 * c.r1 * f001.v[j] +
                    c.p0 * c.q1 * c.r0 * f010.v[j] + c.p0 * c.q1 * c.r1 * f011.v[j] +
                    c.p1 * c.q0 * c.r0 * f100.v[j] + c.p1 * c.q0 * c.r1 * f101.v[j] +
                    c.p1 * c.q1 * c.r0 * f110.v[j] + c.p1 * c.q1 * c.r1 * f111.v[j];
          out[oi] = ACCUM ? out[oi] + f : f;
        } else {
          float gx = dterm(g.sx, c.q0, c.q1, c.r0, c.r1, f100.v[j] - f000.v[j], f101.v[j] - f001.v[j],
                           f110.v[j] - f010.v[j], f111.v[j] - f011.v[j]);
          float gy = dterm(g.sy, c.p0, c.p1, c.r0, c.r1, f010.v[j] - f000.v[j], f011.v[j] - f001.v[j],
                           f110.v[j] - f100.v[j], f111.v[j] - f101.v[j]);
          float gz = dterm(g.sz, c.p0, c.p1, c.q0, c.q1, f001.v[j] - f000.v[j], f011.v[j] - f010.v[j],
                           f101.v[j] - f100.v[j], f111.v[j] - f110.v[j]);
          if (MODE == GRAD_QUERY) {
            float go = __ldg(a + oi);
            ax += go * gx; ay += go * gy; az += go * gz;
          } else {
            float v = ggx * gx + ggy * gy + ggz * gz;
            out[oi] = ACCUM ? out[oi] + v : v;
          }
        }
      }
    }
    if (MODE == GRAD_QUERY) {  // L threads contribute to one point: atomics (zero-filled by the host if !accum)
      atomicAdd(out + b * 3, ax); atomicAdd(out + b * 3 + 1, ay); atomicAdd(out + b * 3 + 2, az);
    }
  }
}

template <bool SECOND, int V>
__global__ void __launch_bounds__(NDJIR_BLOCK)
scatter_kernel(long long B, float* __restrict__ gf, const float* __restrict__ go_, const float* __restrict__ gg,
               const float* __restrict__ query, HashSpec h, int layout, int l_begin) {
  __shared__ LevelTable tab;
  build_table(tab, h);
  const long long N = B * (h.L - l_begin);    // levels below l_begin were handled by scatter_coarse_kernel
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += stride) {
    int l = l_begin + (int)(n / B);
    long long b = n - (long long)(l - l_begin) * B;
    GridFrame g = level_frame(h, tab.G[l]);
    unsigned T = (unsigned)tab.T[l];
    float* gl = gf + tab.off[l];
    const float* q = query + b * 3;
    Cell c = make_cell(g, __ldg(q), __ldg(q + 1), __ldg(q + 2));
    const unsigned xs[2] = {c.x0, c.x1}, ys[2] = {c.y0, c.y1}, zs[2] = {c.z0, c.z1};
    const float ps[2] = {c.p0, c.p1}, qs[2] = {c.q0, c.q1}, rs[2] = {c.r0, c.r1};
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
    if (SECOND) {
      ggx = __ldg(gg + b * 3) * g.sx; ggy = __ldg(gg + b * 3 + 1) * g.sy; ggz = __ldg(gg + b * 3 + 2) * g.sz;
    }
    for (int d = 0; d < h.D; d += V) {
      Vec<V> o;
#pragma unroll
      for (int j = 0; j < V; ++j) o.v[j] = __ldg(go_ + out_index(layout, d + j, l, b, h.L, B, h.D));
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int cx = (k >> 2) & 1, cy = (k >> 1) & 1, cz = k & 1;
        float coef = SECOND ? (ggx * ((cx ? 1.f : -1.f) * qs[cy] * rs[cz]) + ggy * ((cy ? 1.f : -1.f) * ps[cx] * rs[cz]) +
                               ggz * ((cz ? 1.f : -1.f) * ps[cx] * qs[cy]))
                            : ps[cx] * qs[cy] * rs[cz];
        Vec<V> val;
#pragma unroll
        for (int j = 0; j < V; ++j) val.v[j] = o.v[j] * coef;
        red_vec<V>(gl + hash3(xs[cx], ys[cy], zs[cz], T) * h.D + d, val);
      }
    }
  }
}

// Forward over the COARSE levels of a large batch with the level table staged in shared memory: the gathers of the
// per-thread kernel above are bound by the L1 tag stage (one look-up per lane and clock, 0.26 ms per level at 2^24
// points); a level that fits 160 KB is served from shared memory by a persistent grid instead.
template <bool ACCUM>
__global__ void __launch_bounds__(1024, 1)
fwd_coarse_kernel(long long B, float* __restrict__ out, const float* __restrict__ query, const float* __restrict__ feat,
                  HashSpec h, int layout, int n_coarse) {
  extern __shared__ float tabs[];
  __shared__ LevelTable tab;
  build_table(tab, h);
  for (int l = 0; l < n_coarse; ++l) {
    const unsigned T = (unsigned)tab.T[l];
    const unsigned n_fl = T * (unsigned)h.D;
    __syncthreads();
    for (unsigned i = threadIdx.x; i < n_fl; i += blockDim.x) tabs[i] = __ldg(feat + tab.off[l] + i);
    __syncthreads();
    GridFrame g = level_frame(h, tab.G[l]);
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
      const float* q = query + b * 3;
      Cell c = make_cell_linear(g, __ldg(q), __ldg(q + 1), __ldg(q + 2));
      const float* e000 = tabs + hash3(c.x0, c.y0, c.z0, T) * h.D; const float* e001 = tabs + hash3(c.x0, c.y0, c.z1, T) * h.D;
      const float* e010 = tabs + hash3(c.x0, c.y1, c.z0, T) * h.D; const float* e011 = tabs + hash3(c.x0, c.y1, c.z1, T) * h.D;
      const float* e100 = tabs + hash3(c.x1, c.y0, c.z0, T) * h.D; const float* e101 = tabs + hash3(c.x1, c.y0, c.z1, T) * h.D;
      const float* e110 = tabs + hash3(c.x1, c.y1, c.z0, T) * h.D; const float* e111 = tabs + hash3(c.x1, c.y1, c.z1, T) * h.D;
      for (int d = 0; d < h.D; ++d) {
        float f = c.p0 * c.q0 * c.r0 * e000[d] + c.p0 * c.q0 * c.r1 * e001[d] + c.p0 * c.q1 * c.r0 * e010[d] +
                  c.p0 * c.q1 * c.r1 * e011[d] + c.p1 * c.q0 * c.r0 * e100[d] + c.p1 * c.q0 * c.r1 * e101[d] +
                  c.p1 * c.q1 * c.r0 * e110[d] + c.p1 * c.q1 * c.r1 * e111[d];      // gather_kernel<FWD>'s expression
        long long oi = out_index(layout, d, l, b, h.L, B, h.D);
        out[oi] = ACCUM ? out[oi] + f : f;
      }
    }
  }
}

// First-order scatter into the COARSE levels of a large batch.  At the bench shape 2^24 points send 134 M reductions
// into the 4096 entries of level 0 (13 824 at level 1): same-address reductions serialise in the L2 atomic units and
// these two levels cost 4.1 + 1.8 ms of a 13 ms call (tools/exp/hash_levels.py).  Here every CTA of a persistent
// grid accumulates a level into a private copy in shared memory and flushes it with one reduction per non-zero entry.
__global__ void __launch_bounds__(1024, 1)
scatter_coarse_kernel(long long B, float* __restrict__ gf, const float* __restrict__ go_,
                      const float* __restrict__ query, HashSpec h, int layout, int n_coarse) {
  extern __shared__ float priv[];
  __shared__ LevelTable tab;
  build_table(tab, h);
  for (int l = 0; l < n_coarse; ++l) {
    const unsigned T = (unsigned)tab.T[l];
    const unsigned n_fl = T * (unsigned)h.D;
    for (unsigned i = threadIdx.x; i < n_fl; i += blockDim.x) priv[i] = 0.f;
    __syncthreads();
    GridFrame g = level_frame(h, tab.G[l]);
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
      const float* q = query + b * 3;
      Cell c = make_cell_linear(g, __ldg(q), __ldg(q + 1), __ldg(q + 2));
      const unsigned xs[2] = {c.x0, c.x1}, ys[2] = {c.y0, c.y1}, zs[2] = {c.z0, c.z1};
      const float ps[2] = {c.p0, c.p1}, qs[2] = {c.q0, c.q1}, rs[2] = {c.r0, c.r1};
      for (int d = 0; d < h.D; ++d) {
        float go = __ldg(go_ + out_index(layout, d, l, b, h.L, B, h.D));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          int cx = (k >> 2) & 1, cy = (k >> 1) & 1, cz = k & 1;
          atomicAdd(&priv[hash3(xs[cx], ys[cy], zs[cz], T) * h.D + d], go * (ps[cx] * qs[cy] * rs[cz]));
        }
      }
    }
    __syncthreads();
    float* gl = gf + tab.off[l];
    for (unsigned i = threadIdx.x; i < n_fl; i += blockDim.x) {
      float v = priv[i];
      if (v != 0.f) atomicAdd(gl + i, v);
    }
    __syncthreads();
  }
}

// ---- Lanczos-2 variant (csrc/grid_feature/lanczos_voxel_hash_feature_cuda.cu, 6 exports :959-976): per level the
// 4x4x4 clamped taps of lanczos_voxel, each hashed into the level's table; same level table, offsets and output
// layouts as the trilinear family.  The window weights are evaluated once per (level, point), the 64 hashed indices
// once per tap (the reference recomputes both per channel). ---------------------------------------------------------
__global__ void __launch_bounds__(NDJIR_BLOCK)
lz_hash_index_kernel(long long B, float* __restrict__ out, const float* __restrict__ query, unsigned T) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += stride) {
    const float* q = query + b * 3;   // integer cell coordinates stored as floats (:54-66)
    out[b] = (float)hash3((unsigned)__ldg(q), (unsigned)__ldg(q + 1), (unsigned)__ldg(q + 2), T);
  }
}

template <int MODE, int V, bool ACCUM>
__global__ void __launch_bounds__(NDJIR_BLOCK)
lz_gather_kernel(long long B, float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ gg,
                 const float* __restrict__ query, const float* __restrict__ feat, HashSpec h, int layout) {
  using namespace ndjir::lanczos;
  __shared__ LevelTable tab;
  build_table(tab, h);
  const long long N = B * h.L;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += stride) {
    int l = (int)(n / B);
    long long b = n - (long long)l * B;
    GridFrame g = level_frame(h, tab.G[l]);
    unsigned T = (unsigned)tab.T[l];
    const float* fl = feat + tab.off[l];
    Taps t = make_taps<MODE != FWD>(g, query + b * 3);
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
    if (MODE == GGO) { ggx = __ldg(gg + b * 3); ggy = __ldg(gg + b * 3 + 1); ggz = __ldg(gg + b * 3 + 2); }
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int d = 0; d < h.D; d += V) {
      float f[V], gx[V], gy[V], gz[V];
#pragma unroll
      for (int j = 0; j < V; ++j) { f[j] = 0.f; gx[j] = 0.f; gy[j] = 0.f; gz[j] = 0.f; }
#pragma unroll
      for (int i = 0; i < K; ++i) {
#pragma unroll
        for (int jj = 0; jj < K; ++jj) {
          Vec<V> v[K];
#pragma unroll
          for (int k = 0; k < K; ++k) v[k] = ldg_vec<V>(fl + hash3(t.ix[i], t.iy[jj], t.iz[k], T) * h.D + d);
#pragma unroll
          for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
              if (MODE == FWD) {
                f[j] += t.cx[i] * t.cy[jj] * t.cz[k] * v[k].v[j];                 // :184-186
              } else {
                gx[j] += g.sx * t.gx[i] * t.cy[jj] * t.cz[k] * v[k].v[j];          // :283-285
                gy[j] += g.sy * t.cx[i] * t.gy[jj] * t.cz[k] * v[k].v[j];
                gz[j] += g.sz * t.cx[i] * t.cy[jj] * t.gz[k] * v[k].v[j];
              }
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < V; ++j) {
        long long oi = out_index(layout, d + j, l, b, h.L, B, h.D);
        if (MODE == FWD) {
          out[oi] = ACCUM ? out[oi] + f[j] : f[j];
        } else if (MODE == GRAD_QUERY) {
          float go = __ldg(a + oi);
          ax += go * gx[j]; ay += go * gy[j]; az += go * gz[j];
        } else {
          float v = ggx * gx[j] + ggy * gy[j] + ggz * gz[j];
          out[oi] = ACCUM ? out[oi] + v : v;
        }
      }
    }
    if (MODE == GRAD_QUERY) {
      atomicAdd(out + b * 3, ax); atomicAdd(out + b * 3 + 1, ay); atomicAdd(out + b * 3 + 2, az);
    }
  }
}

template <bool SECOND, int V>
__global__ void __launch_bounds__(NDJIR_BLOCK)
lz_scatter_kernel(long long B, float* __restrict__ gf, const float* __restrict__ go_, const float* __restrict__ gg,
                  const float* __restrict__ query, HashSpec h, int layout) {
  using namespace ndjir::lanczos;
  __shared__ LevelTable tab;
  build_table(tab, h);
  const long long N = B * h.L;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += stride) {
    int l = (int)(n / B);
    long long b = n - (long long)l * B;
    GridFrame g = level_frame(h, tab.G[l]);
    unsigned T = (unsigned)tab.T[l];
    float* gl = gf + tab.off[l];
    Taps t = make_taps<SECOND>(g, query + b * 3);
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
    if (SECOND) {
      ggx = __ldg(gg + b * 3) * g.sx; ggy = __ldg(gg + b * 3 + 1) * g.sy; ggz = __ldg(gg + b * 3 + 2) * g.sz;
    }
    for (int d = 0; d < h.D; d += V) {
      Vec<V> o;
#pragma unroll
      for (int j = 0; j < V; ++j) o.v[j] = __ldg(go_ + out_index(layout, d + j, l, b, h.L, B, h.D));
#pragma unroll
      for (int i = 0; i < K; ++i) {
#pragma unroll
        for (int jj = 0; jj < K; ++jj) {
#pragma unroll
          for (int k = 0; k < K; ++k) {
            float coef = SECOND ? (ggx * (t.gx[i] * t.cy[jj] * t.cz[k]) + ggy * (t.cx[i] * t.gy[jj] * t.cz[k]) +
                                   ggz * (t.cx[i] * t.cy[jj] * t.gz[k]))                       // :690-697
                                : t.cx[i] * t.cy[jj] * t.cz[k];                                  // :362-369
            Vec<V> val;
#pragma unroll
            for (int j = 0; j < V; ++j) val.v[j] = o.v[j] * coef;
            red_vec<V>(gl + hash3(t.ix[i], t.iy[jj], t.iz[k], T) * h.D + d, val);
          }
        }
      }
    }
  }
}

// ---- total variation on the hashed levels (csrc/grid_feature/total_variation_loss_on_voxel_hash_cuda.cu:50-106 forward,
// :132-196 backward): sqrt of the squared forward differences at the lower corner; the backward sends
// ograd * delta * rsqrt(sum + 1e-12) (double intermediate, like the reference's literal) to the three upper
// neighbours only - the lower corner itself gets no gradient in this family (no sym_backward argument). --------------
template <bool BWD, int V>
__global__ void __launch_bounds__(NDJIR_BLOCK)
tv_hash_kernel(long long B, float* __restrict__ out, const float* __restrict__ go_, const float* __restrict__ query,
               const float* __restrict__ feat, HashSpec h, int layout) {
  __shared__ LevelTable tab;
  build_table(tab, h);
  const long long N = B * h.L;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += stride) {
    int l = (int)(n / B);
    long long b = n - (long long)l * B;
    GridFrame g = level_frame(h, tab.G[l]);
    unsigned T = (unsigned)tab.T[l];
    const float* fl = feat + tab.off[l];
    const float* q = query + b * 3;
    Cell c = make_cell(g, __ldg(q), __ldg(q + 1), __ldg(q + 2));
    unsigned i000 = hash3(c.x0, c.y0, c.z0, T) * h.D, i001 = hash3(c.x0, c.y0, c.z1, T) * h.D;
    unsigned i010 = hash3(c.x0, c.y1, c.z0, T) * h.D, i100 = hash3(c.x1, c.y0, c.z0, T) * h.D;
    for (int d = 0; d < h.D; d += V) {
      Vec<V> f000 = ldg_vec<V>(fl + i000 + d), f001 = ldg_vec<V>(fl + i001 + d);
      Vec<V> f010 = ldg_vec<V>(fl + i010 + d), f100 = ldg_vec<V>(fl + i100 + d);
      Vec<V> gx, gy, gz;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float dx = f100.v[j] - f000.v[j], dy = f010.v[j] - f000.v[j], dz = f001.v[j] - f000.v[j];
        float s2 = dx * dx + dy * dy + dz * dz;
        long long oi = out_index(layout, d + j, l, b, h.L, B, h.D);
        if (!BWD) {
          out[oi] = sqrtf(s2);
        } else {
          double common = (double)__ldg(go_ + oi) * rsqrt((double)s2 + 1e-12);
          gx.v[j] = (float)(common * dx); gy.v[j] = (float)(common * dy); gz.v[j] = (float)(common * dz);
        }
      }
      if (BWD) {
        float* gl = out + tab.off[l];
        red_vec<V>(gl + i100 + d, gx); red_vec<V>(gl + i010 + d, gy); red_vec<V>(gl + i001 + d, gz);
      }
    }
  }
}

static int make_spec(HashSpec& h, int G0, float gf, int T0, int L, int D, const float* mn, const float* mx) {
  if (G0 <= 0 || T0 <= 0 || L <= 0 || L > NDJIR_HASH_MAX_LEVELS || D <= 0 || !mn || !mx) return NDJIR_ERR_ARG;
  h.G0 = G0; h.gf = gf; h.T0 = T0; h.L = L; h.D = D;
  h.mnx = mn[0]; h.mny = mn[1]; h.mnz = mn[2];
  volatile float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
  h.dx = dx; h.dy = dy; h.dz = dz;
  return NDJIR_OK;
}

// Level offsets are multiples of 2 floats only when (T*D)%8 is even; pick the widest safe vector.
static int hash_vec(const HashSpec& h, const void* table) {
  bool all_even = true, all_quad = true;
  long long o = 0;
  for (int l = 0; l < h.L; ++l) {
    if (o % 2) all_even = false;
    if (o % 4) all_quad = false;
    int G = level_grid_size(h.G0, h.gf, l);
    o += force_align(level_table_size(G, h.T0) * h.D);
  }
  if (h.D % 4 == 0 && all_quad && aligned16(table)) return 4;
  if (h.D % 2 == 0 && all_even && aligned8(table)) return 2;
  return 1;
}

}  // namespace vhash
}  // namespace ndjir

using namespace ndjir;
using namespace ndjir::vhash;

extern "C" {

long long ndjir_voxel_hash_num_params(int G0, float growth_factor, int T0, int L, int D) {
  long long o = 0;
  for (int l = 0; l < L; ++l) o += force_align(level_table_size(level_grid_size(G0, growth_factor, l), T0) * D);
  return o;
}

int ndjir_voxel_hash_level_table(int G0, float growth_factor, int T0, int L, int D, int* G_out, int* T_out,
                                 long long* offset_out) {
  if (L <= 0 || L > NDJIR_HASH_MAX_LEVELS) return NDJIR_ERR_ARG;
  long long o = 0;
  for (int l = 0; l < L; ++l) {
    int G = level_grid_size(G0, growth_factor, l);
    int T = level_table_size(G, T0);
    if (G_out) G_out[l] = G;
    if (T_out) T_out[l] = T;
    if (offset_out) offset_out[l] = o;
    o += force_align(T * D);
  }
  return NDJIR_OK;
}

int ndjir_voxel_hash_level_table_device(int G0, float growth_factor, int T0, int L, int D, int* G_dev, int* T_dev,
                                        cudaStream_t st) {
  HashSpec h;
  const float mn[3] = {-1.f, -1.f, -1.f}, mx[3] = {1.f, 1.f, 1.f};
  if (make_spec(h, G0, growth_factor, T0, L, D, mn, mx) || !G_dev || !T_dev) return NDJIR_ERR_ARG;
  level_table_kernel<<<1, 64, 0, st>>>(h, G_dev, T_dev);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_voxel_hash_hash_index(long long n_points, float* output, const float* query, int G, int T,
                                const float* min3, const float* max3, cudaStream_t st) {
  if (n_points == 0) return NDJIR_OK;
  HashSpec h;
  if (make_spec(h, G, 1.f, T, 1, 1, min3, max3) || !output || !query || n_points < 0) return NDJIR_ERR_ARG;
  hash_index_kernel<<<grid_for(n_points), NDJIR_BLOCK, 0, st>>>(n_points, output, query, G, T, h);
  NDJIR_RETURN_LAST_ERROR();
}

#define NDJIR_HASH_GATHER(MODE, accum_)                                                                          \
  do {                                                                                                           \
    int V = hash_vec(h, feature);                                                                                \
    int grid = grid_for(n_points * (L - l_begin));                                                               \
    if (V == 4) {                                                                                                \
      if (accum_) gather_kernel<MODE, 4, true><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, out, a, gg, query, feature, h, layout, l_begin); \
      else gather_kernel<MODE, 4, false><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, out, a, gg, query, feature, h, layout, l_begin);       \
    } else if (V == 2) {                                                                                         \
      if (accum_) gather_kernel<MODE, 2, true><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, out, a, gg, query, feature, h, layout, l_begin); \
      else gather_kernel<MODE, 2, false><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, out, a, gg, query, feature, h, layout, l_begin);       \
    } else {                                                                                                     \
      if (accum_) gather_kernel<MODE, 1, true><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, out, a, gg, query, feature, h, layout, l_begin); \
      else gather_kernel<MODE, 1, false><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, out, a, gg, query, feature, h, layout, l_begin);       \
    }                                                                                                            \
  } while (0)

int ndjir_voxel_hash_voxel_hash_feature(long long n_points, float* output, const float* query,
                                        const float* feature, int G0, float growth_factor, int T0, int L, int D,
                                        const float* min3, const float* max3, int layout, int accum,
                                        cudaStream_t st) {
  if (n_points == 0) return NDJIR_OK;
  HashSpec h;
  if (make_spec(h, G0, growth_factor, T0, L, D, min3, max3) || !output || !query || !feature || n_points < 0)
    return NDJIR_ERR_ARG;
  float* out = output; const float* a = nullptr; const float* gg = nullptr;
  int l_begin = 0;
  if (g_hash_coarse_private && n_points >= (g_hash_coarse_private == 2 ? 1 : (1ll << 20))) {
    long long max_fl = 0;
    while (l_begin < L) {       // the dense prefix of levels whose table fits 160 KB of shared memory
      // sized for a grid one cell larger than the host evaluates: the kernels index shared memory with the table the
      // DEVICE builds, and host pow / device powf may floor G0 * gf^l differently (q2)
      long long fl = (long long)level_table_size(level_grid_size(G0, growth_factor, l_begin) + 1, T0) * D;
      if (fl * 4 > 160 * 1024) break;
      if (fl > max_fl) max_fl = fl;
      ++l_begin;
    }
    if (l_begin > 0) {
      size_t smem = (size_t)max_fl * 4;
      cudaError_t e = accum ? cudaFuncSetAttribute(fwd_coarse_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                            : cudaFuncSetAttribute(fwd_coarse_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { (void)cudaGetLastError(); l_begin = 0; }
      else if (accum) fwd_coarse_kernel<true><<<NDJIR_NUM_SMS, 1024, smem, st>>>(n_points, output, query, feature, h, layout, l_begin);
      else fwd_coarse_kernel<false><<<NDJIR_NUM_SMS, 1024, smem, st>>>(n_points, output, query, feature, h, layout, l_begin);
    }
  }
  if (l_begin < L) { NDJIR_HASH_GATHER(FWD, accum); }
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_voxel_hash_grad_query(long long n_points, float* grad_query, const float* grad_output,
                                const float* query, const float* feature, int G0, float growth_factor, int T0,
                                int L, int D, const float* min3, const float* max3, int layout, int accum,
                                cudaStream_t st) {
  if (n_points == 0) return NDJIR_OK;
  HashSpec h;
  if (make_spec(h, G0, growth_factor, T0, L, D, min3, max3) || !grad_query || !grad_output || !query || !feature ||
      n_points < 0)
    return NDJIR_ERR_ARG;
  if (!accum) fill_zero(grad_query, n_points * 3, st);
  float* out = grad_query; const float* a = grad_output; const float* gg = nullptr;
  const int l_begin = 0;
  NDJIR_HASH_GATHER(GRAD_QUERY, true);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_voxel_hash_grad_query_grad_grad_output(long long n_points, float* grad_grad_output,
                                                 const float* grad_grad_query, const float* query,
                                                 const float* feature, int G0, float growth_factor, int T0,
                                                 int L, int D, const float* min3, const float* max3, int layout,
                                                 int accum, cudaStream_t st) {
  if (n_points == 0) return NDJIR_OK;
  HashSpec h;
  if (make_spec(h, G0, growth_factor, T0, L, D, min3, max3) || !grad_grad_output || !grad_grad_query || !query ||
      !feature || n_points < 0)
    return NDJIR_ERR_ARG;
  float* out = grad_grad_output; const float* a = nullptr; const float* gg = grad_grad_query;
  const int l_begin = 0;
  NDJIR_HASH_GATHER(GGO, accum);
  NDJIR_RETURN_LAST_ERROR();
}

#define NDJIR_HASH_SCATTER(SECOND)                                                                             \
  do {                                                                                                         \
    int V = hash_vec(h, grad_feature);                                                                         \
    int grid = grid_for(n_points * (L - l_begin));                                                             \
    if (V == 4) scatter_kernel<SECOND, 4><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, grad_feature, grad_output, gg, query, h, layout, l_begin);      \
    else if (V == 2) scatter_kernel<SECOND, 2><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, grad_feature, grad_output, gg, query, h, layout, l_begin); \
    else scatter_kernel<SECOND, 1><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, grad_feature, grad_output, gg, query, h, layout, l_begin);             \
  } while (0)

int ndjir_voxel_hash_grad_feature(long long n_points, float* grad_feature, const float* grad_output,
                                  const float* query, int G0, float growth_factor, int T0, int L, int D,
                                  const float* min3, const float* max3, int layout, int accum, cudaStream_t st) {
  HashSpec h;
  if (make_spec(h, G0, growth_factor, T0, L, D, min3, max3) || !grad_feature || n_points < 0) return NDJIR_ERR_ARG;
  if (!accum) fill_zero(grad_feature, ndjir_voxel_hash_num_params(G0, growth_factor, T0, L, D), st);
  if (n_points == 0) NDJIR_RETURN_LAST_ERROR();
  if (!grad_output || !query) return NDJIR_ERR_ARG;
  const float* gg = nullptr;
  int l_begin = 0;
  if (g_hash_coarse_private && n_points >= (g_hash_coarse_private == 2 ? 1 : (1ll << 20))) {
    // coarse levels (private copy <= 160 KB of shared memory, a dense prefix of the level list) go through the
    // shared-memory privatised kernel
    long long max_fl = 0;
    while (l_begin < L) {
      // sized for a grid one cell larger than the host evaluates: the kernels index shared memory with the table the
      // DEVICE builds, and host pow / device powf may floor G0 * gf^l differently (q2)
      long long fl = (long long)level_table_size(level_grid_size(G0, growth_factor, l_begin) + 1, T0) * D;
      if (fl * 4 > 160 * 1024 || fl / D >= n_points) break;              // too large, or too few points to contend
      if (fl > max_fl) max_fl = fl;
      ++l_begin;
    }
    if (l_begin > 0) {
      size_t smem = (size_t)max_fl * 4;
      cudaError_t e = cudaFuncSetAttribute(scatter_coarse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { (void)cudaGetLastError(); l_begin = 0; }
      else scatter_coarse_kernel<<<NDJIR_NUM_SMS, 1024, smem, st>>>(n_points, grad_feature, grad_output, query, h, layout, l_begin);
    }
  }
  if (l_begin < L) { NDJIR_HASH_SCATTER(false); }
  NDJIR_RETURN_LAST_ERROR();
}

// Always accumulates (voxel_hash_feature_cuda.cu:750-775 has no zero-fill).
int ndjir_voxel_hash_grad_query_grad_feature(long long n_points, float* grad_feature,
                                             const float* grad_grad_query, const float* grad_output,
                                             const float* query, int G0, float growth_factor, int T0, int L,
                                             int D, const float* min3, const float* max3, int layout,
                                             cudaStream_t st) {
  if (n_points == 0) return NDJIR_OK;
  HashSpec h;
  if (make_spec(h, G0, growth_factor, T0, L, D, min3, max3) || !grad_feature || !grad_grad_query || !grad_output ||
      !query || n_points < 0)
    return NDJIR_ERR_ARG;
  const float* gg = grad_grad_query;
  const int l_begin = 0;
  NDJIR_HASH_SCATTER(true);
  NDJIR_RETURN_LAST_ERROR();
}

// ---- lanczos_voxel_hash_feature_cuda (csrc/grid_feature/lanczos_voxel_hash_feature_cuda.cu:959-976) -------------------
// hash_index here hashes ONE integer cell per row (:54-66), not the 8 corners of a query.
int ndjir_lanczos_voxel_hash_hash_index(long long n_points, float* output, const float* query, int T,
                                        cudaStream_t st) {
  if (n_points == 0) return NDJIR_OK;
  if (n_points < 0 || T <= 0 || !output || !query) return NDJIR_ERR_ARG;
  lz_hash_index_kernel<<<grid_for(n_points), NDJIR_BLOCK, 0, st>>>(n_points, output, query, (unsigned)T);
  NDJIR_RETURN_LAST_ERROR();
}

#define NDJIR_LZ_HASH_GATHER(MODE, accum_)                                                                       \
  do {                                                                                                           \
    int V = hash_vec(h, feature);                                                                                \
    int grid = grid_for(n_points * L);                                                                           \
    if (V == 4) {                                                                                                \
      if (accum_) lz_gather_kernel<MODE, 4, true><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, out, a, gg, query, feature, h, layout); \
      else lz_gather_kernel<MODE, 4, false><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, out, a, gg, query, feature, h, layout);       \
    } else if (V == 2) {                                                                                         \
      if (accum_) lz_gather_kernel<MODE, 2, true><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, out, a, gg, query, feature, h, layout); \
      else lz_gather_kernel<MODE, 2, false><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, out, a, gg, query, feature, h, layout);       \
    } else {                                                                                                     \
      if (accum_) lz_gather_kernel<MODE, 1, true><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, out, a, gg, query, feature, h, layout); \
      else lz_gather_kernel<MODE, 1, false><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, out, a, gg, query, feature, h, layout);       \
    }                                                                                                            \
  } while (0)

int ndjir_lanczos_voxel_hash_voxel_hash_feature(long long n_points, float* output, const float* query,
                                                const float* feature, int G0, float growth_factor, int T0, int L,
                                                int D, const float* min3, const float* max3, int layout, int accum,
                                                cudaStream_t st) {
  if (n_points == 0) return NDJIR_OK;
  HashSpec h;
  if (make_spec(h, G0, growth_factor, T0, L, D, min3, max3) || !output || !query || !feature || n_points < 0)
    return NDJIR_ERR_ARG;
  float* out = output; const float* a = nullptr; const float* gg = nullptr;
  NDJIR_LZ_HASH_GATHER(FWD, accum);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_lanczos_voxel_hash_grad_query(long long n_points, float* grad_query, const float* grad_output,
                                        const float* query, const float* feature, int G0, float growth_factor,
                                        int T0, int L, int D, const float* min3, const float* max3, int layout,
                                        int accum, cudaStream_t st) {
  if (n_points == 0) return NDJIR_OK;
  HashSpec h;
  if (make_spec(h, G0, growth_factor, T0, L, D, min3, max3) || !grad_query || !grad_output || !query || !feature ||
      n_points < 0)
    return NDJIR_ERR_ARG;
  if (!accum) fill_zero(grad_query, n_points * 3, st);
  float* out = grad_query; const float* a = grad_output; const float* gg = nullptr;
  NDJIR_LZ_HASH_GATHER(GRAD_QUERY, true);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_lanczos_voxel_hash_grad_query_grad_grad_output(long long n_points, float* grad_grad_output,
                                                         const float* grad_grad_query, const float* query,
                                                         const float* feature, int G0, float growth_factor, int T0,
                                                         int L, int D, const float* min3, const float* max3,
                                                         int layout, int accum, cudaStream_t st) {
  if (n_points == 0) return NDJIR_OK;
  HashSpec h;
  if (make_spec(h, G0, growth_factor, T0, L, D, min3, max3) || !grad_grad_output || !grad_grad_query || !query ||
      !feature || n_points < 0)
    return NDJIR_ERR_ARG;
  float* out = grad_grad_output; const float* a = nullptr; const float* gg = grad_grad_query;
  NDJIR_LZ_HASH_GATHER(GGO, accum);
  NDJIR_RETURN_LAST_ERROR();
}

#define NDJIR_LZ_HASH_SCATTER(SECOND)                                                                          \
  do {                                                                                                         \
    int V = hash_vec(h, grad_feature);                                                                         \
    int grid = grid_for(n_points * L);                                                                         \
    if (V == 4) lz_scatter_kernel<SECOND, 4><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, grad_feature, grad_output, gg, query, h, layout);      \
    else if (V == 2) lz_scatter_kernel<SECOND, 2><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, grad_feature, grad_output, gg, query, h, layout); \
    else lz_scatter_kernel<SECOND, 1><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, grad_feature, grad_output, gg, query, h, layout);             \
  } while (0)

int ndjir_lanczos_voxel_hash_grad_feature(long long n_points, float* grad_feature, const float* grad_output,
                                          const float* query, int G0, float growth_factor, int T0, int L, int D,
                                          const float* min3, const float* max3, int layout, int accum,
                                          cudaStream_t st) {
  HashSpec h;
  if (make_spec(h, G0, growth_factor, T0, L, D, min3, max3) || !grad_feature || n_points < 0) return NDJIR_ERR_ARG;
  if (!accum) fill_zero(grad_feature, ndjir_voxel_hash_num_params(G0, growth_factor, T0, L, D), st);
  if (n_points == 0) NDJIR_RETURN_LAST_ERROR();
  if (!grad_output || !query) return NDJIR_ERR_ARG;
  const float* gg = nullptr;
  NDJIR_LZ_HASH_SCATTER(false);
  NDJIR_RETURN_LAST_ERROR();
}

// Always accumulates (lanczos_voxel_hash_feature_cuda.cu:708-735 has no zero-fill).
int ndjir_lanczos_voxel_hash_grad_query_grad_feature(long long n_points, float* grad_feature,
                                                     const float* grad_grad_query, const float* grad_output,
                                                     const float* query, int G0, float growth_factor, int T0, int L,
                                                     int D, const float* min3, const float* max3, int layout,
                                                     cudaStream_t st) {
  if (n_points == 0) return NDJIR_OK;
  HashSpec h;
  if (make_spec(h, G0, growth_factor, T0, L, D, min3, max3) || !grad_feature || !grad_grad_query || !grad_output ||
      !query || n_points < 0)
    return NDJIR_ERR_ARG;
  const float* gg = grad_grad_query;
  NDJIR_LZ_HASH_SCATTER(true);
  NDJIR_RETURN_LAST_ERROR();
}

// ---- total_variation_loss_on_voxel_hash_cuda (csrc/grid_feature/total_variation_loss_on_voxel_hash_cuda.cu:229-234) ----
int ndjir_tv_loss_on_voxel_hash(long long n_points, float* output, const float* query, const float* feature, int G0,
                                float growth_factor, int T0, int L, int D, const float* min3, const float* max3,
                                int layout, cudaStream_t st) {
  if (n_points == 0) return NDJIR_OK;
  HashSpec h;
  if (make_spec(h, G0, growth_factor, T0, L, D, min3, max3) || !output || !query || !feature || n_points < 0)
    return NDJIR_ERR_ARG;
  int V = hash_vec(h, feature);
  int grid = grid_for(n_points * L);
  if (V == 4) tv_hash_kernel<false, 4><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, output, nullptr, query, feature, h, layout);
  else if (V == 2) tv_hash_kernel<false, 2><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, output, nullptr, query, feature, h, layout);
  else tv_hash_kernel<false, 1><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, output, nullptr, query, feature, h, layout);
  NDJIR_RETURN_LAST_ERROR();
}

// Always accumulates (:198-216 has no zero-fill).
int ndjir_tv_loss_on_voxel_hash_backward(long long n_points, float* grad_feature, const float* grad_output,
                                         const float* query, const float* feature, int G0, float growth_factor,
                                         int T0, int L, int D, const float* min3, const float* max3, int layout,
                                         cudaStream_t st) {
  if (n_points == 0) return NDJIR_OK;
  HashSpec h;
  if (make_spec(h, G0, growth_factor, T0, L, D, min3, max3) || !grad_feature || !grad_output || !query || !feature ||
      n_points < 0)
    return NDJIR_ERR_ARG;
  int V = hash_vec(h, feature);
  if (V > 1 && hash_vec(h, grad_feature) < V) V = hash_vec(h, grad_feature);
  int grid = grid_for(n_points * L);
  if (V == 4) tv_hash_kernel<true, 4><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, grad_feature, grad_output, query, feature, h, layout);
  else if (V == 2) tv_hash_kernel<true, 2><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, grad_feature, grad_output, query, feature, h, layout);
  else tv_hash_kernel<true, 1><<<grid, NDJIR_BLOCK, 0, st>>>(n_points, grad_feature, grad_output, query, feature, h, layout);
  NDJIR_RETURN_LAST_ERROR();
}

}  // extern "C"
