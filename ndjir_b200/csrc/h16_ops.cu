// Split-fp16 MLP engine: everything that is not a tensor-core product (sm_100a).
//   * the memory-bound corner shapes of the MLP passes with split-fp16 activations (one pass over the large operand):
//       skinny_n_h  C (M x N<=8)  = A(split) B32          last layers (sdf, colours ...), grid-feature gradient
//       skinny_k_h  C (M x N)     = A32 (M x K<=8) B32    rank-1..8 updates with the full fused epilogue in either form
//       skinny_w_h  C (M x N<=8) += A(split, K x M)^T B32 weight gradients of those last layers
//   * format conversion: pack (fp32 -> split), unpack, plane copies, column sums (bias gradients)
//   * the per-tensor power-of-two scales: running max -> scale (delayed by one use of the tensor)
#include "gemm_h.cuh"

namespace ndjir {
namespace gemmh {

constexpr int SK_WARPS = 8;

// ---------------------------------------------------------------------------------------------------------------
// C (M x N <= 8) = A(split) B32.  Eight lanes share a row (each owns 8 of every 64 columns); a lane works on SK_R rows
// at once so that one read of the staged B serves SK_R x 4 rows of the warp: ncu showed the first version (one row per
// lane group) bound by its shared-memory wavefronts (L1 data pipe 77-87 %, DRAM 18-40 %), not by the operand stream.
// B is staged as [n][64-column step][half][lane][4 floats]: the eight lanes of a row read 128 contiguous bytes per
// instruction, the four row groups of a warp the same bytes (broadcast).  The row sums are folded across the eight
// lanes by recursive halving (7 shuffles for 8 outputs instead of 24), which leaves output n in lane n.
template <int NT>
struct SkR {
  static constexpr int R = NT >= 8 ? 2 : 4;     // rows per lane: bounded by the NT x R accumulators a lane can hold
};

template <int NT>
__device__ __forceinline__ float fold8(float (&acc)[NT], int sub) {
  int cnt = NT;
#pragma unroll
  for (int d = 4; d >= 1; d >>= 1) {
    if (cnt > d) {                 // halving stage: the lanes with bit d set keep the upper half of the outputs
      const int half = cnt / 2;
      const bool up = (sub & d) != 0;
#pragma unroll
      for (int t = 0; t < half; ++t) {
        const float send = up ? acc[t] : acc[t + half];
        const float keep = up ? acc[t + half] : acc[t];
        acc[t] = keep + __shfl_xor_sync(0xffffffffu, send, d);
      }
      cnt = half;
    } else {
#pragma unroll
      for (int t = 0; t < NT; ++t)
        if (t < cnt) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], d);
    }
  }
  return acc[0];
}

template <int NT, int EPI>
__global__ void __launch_bounds__(SK_WARPS * 32, 2) skinny_n_h_kernel(HArgs a, int vec) {
  constexpr int SK_R = SkR<NT>::R;
  extern __shared__ __align__(16) float Bs[];
  const int steps = (a.K + 63) / 64;
  for (int i = threadIdx.x; i < steps * 64 * NT; i += blockDim.x) {
    // i = ((n * steps + j) * 2 + half) * 32 + lane8 * 4 + e   <->   k = j * 64 + lane8 * 8 + half * 4 + e
    const int e = i & 3, l8 = (i >> 2) & 7, hf = (i >> 5) & 1, nj = i >> 6;
    const int j = nj % steps, n = nj / steps;
    const int k = j * 64 + l8 * 8 + hf * 4 + e;
    Bs[i] = (n < a.N && k < a.K) ? __ldg(a.B32 + (long long)k * a.b_rs + (long long)n * a.b_cs) : 0.f;
  }
  __syncthreads();
  const float inv_a = 1.f / dev_scalar(a.a_scale);
  const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
  const long long rows_per_block = (long long)SK_WARPS * 4 * SK_R;
  const int K8 = vec ? (a.K & ~7) : 0;
  for (long long blk = blockIdx.x; blk * rows_per_block < a.M; blk += gridDim.x) {
    const long long m0 = blk * rows_per_block + ((threadIdx.x >> 5) * 4 + grp) * SK_R;
    const __half* rh[SK_R];
    const __half* rl[SK_R];
#pragma unroll
    for (int r = 0; r < SK_R; ++r) {
      const long long mm = m0 + r < a.M ? m0 + r : a.M - 1;
      rh[r] = a.Ahi + mm * a.lda;
      rl[r] = a.Alo + mm * a.lda;
    }
    float acc[SK_R][NT];
#pragma unroll
    for (int r = 0; r < SK_R; ++r)
#pragma unroll
      for (int n = 0; n < NT; ++n) acc[r][n] = 0.f;
    for (int j = 0; j < steps; ++j) {
      const int k = j * 64 + sub * 8;
      if (k < K8) {
        uint4 xh[SK_R], xl[SK_R];
#pragma unroll
        for (int r = 0; r < SK_R; ++r) {
          xh[r] = __ldg(reinterpret_cast<const uint4*>(rh[r] + k));
          xl[r] = __ldg(reinterpret_cast<const uint4*>(rl[r] + k));
        }
        float x[SK_R][8];
#pragma unroll
        for (int r = 0; r < SK_R; ++r) {
          const float2 x0 = join2(xh[r].x, xl[r].x), x1 = join2(xh[r].y, xl[r].y), x2 = join2(xh[r].z, xl[r].z),
                       x3 = join2(xh[r].w, xl[r].w);
          x[r][0] = x0.x; x[r][1] = x0.y; x[r][2] = x1.x; x[r][3] = x1.y;
          x[r][4] = x2.x; x[r][5] = x2.y; x[r][6] = x3.x; x[r][7] = x3.y;
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const float* bp = Bs + ((n * steps + j) * 2) * 32 + sub * 4;
          const float4 b0 = *reinterpret_cast<const float4*>(bp);
          const float4 b1 = *reinterpret_cast<const float4*>(bp + 32);
#pragma unroll
          for (int r = 0; r < SK_R; ++r)
            acc[r][n] += x[r][0] * b0.x + x[r][1] * b0.y + x[r][2] * b0.z + x[r][3] * b0.w + x[r][4] * b1.x +
                         x[r][5] * b1.y + x[r][6] * b1.z + x[r][7] * b1.w;
        }
      }
    }
    // columns the 16-byte path does not cover (K % 8 != 0 or unaligned planes)
    for (int k = K8 + sub; k < a.K; k += 8) {
      const int j = k >> 6, l8 = (k >> 3) & 7, hf = (k >> 2) & 1, e = k & 3;
#pragma unroll
      for (int r = 0; r < SK_R; ++r) {
        const float xv = __half2float(rh[r][k]) + __half2float(rl[r][k]);
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[r][n] += xv * Bs[((n * steps + j) * 2 + hf) * 32 + l8 * 4 + e];
      }
    }
#pragma unroll
    for (int r = 0; r < SK_R; ++r) {
      const float v = fold8<NT>(acc[r], sub) * inv_a;
      const long long m = m0 + r;
      if (m < a.M && sub < a.N && sub < NT) {
        float* cp = a.C.f + m * a.C.ldf + sub;
        if (EPI == EPI_BIAS) *cp = a.alpha * v + (a.bias ? __ldg(a.bias + sub) : 0.f);
        else *cp += a.alpha * v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Rank-K update (K <= 8) with the fused epilogue.  Vector form: a thread owns ONE group of 8 adjacent columns for the
// whole kernel (its K x 8 block of B lives in registers) and walks the rows with the block: every access to the
// epilogue operands is a 16-byte piece of a 512-byte contiguous row segment.  Needs N % 8 == 0, blockDim % (N/8) == 0.
template <int EPI, int KT>
__global__ void __launch_bounds__(NDJIR_BLOCK) skinny_k_h_vec_kernel(HArgs a) {
  const int ncol = a.N / 8;
  const int g = threadIdx.x % ncol;
  const int rows_per_block = blockDim.x / ncol;
  const int n = g * 8;
  const float sc = dev_scalar(a.C.scale);
  const float inv_h = 1.f / dev_scalar(a.H.scale), inv_u = 1.f / dev_scalar(a.U.scale);
  const bool need_h = (EPI == EPI_MUL_S);
  const bool need_u = (EPI == EPI_MUL_S) && (a.U.f || a.U.hi);
  float bw[KT][8];
#pragma unroll
  for (int k = 0; k < KT; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e)
      bw[k][e] = k < a.K ? __ldg(a.B32 + (long long)k * a.b_rs + (long long)(n + e) * a.b_cs) : 0.f;
  float b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) b[e] = ((EPI == EPI_BIAS || EPI == EPI_SOFTPLUS) && a.bias) ? __ldg(a.bias + n + e) : 0.f;
  float mx = 0.f;
  const long long row_stride = (long long)gridDim.x * rows_per_block;
  for (long long m = (long long)blockIdx.x * rows_per_block + threadIdx.x / ncol; m < a.M; m += row_stride) {
    float xa[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) xa[k] = k < a.K ? __ldg(a.A32 + m * a.a_rs + (long long)k * a.a_cs) : 0.f;
    float h[8], u[8], cp[8], o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) h[e] = u[e] = cp[e] = 0.f;
    if (need_h) {
      if (a.H.hi) {
        uint4 xh = __ldg(reinterpret_cast<const uint4*>(a.H.hi + m * a.H.ldh + n));
        uint4 xl = __ldg(reinterpret_cast<const uint4*>(a.H.lo + m * a.H.ldh + n));
        float2 t0 = join2(xh.x, xl.x), t1 = join2(xh.y, xl.y), t2 = join2(xh.z, xl.z), t3 = join2(xh.w, xl.w);
        h[0] = t0.x; h[1] = t0.y; h[2] = t1.x; h[3] = t1.y; h[4] = t2.x; h[5] = t2.y; h[6] = t3.x; h[7] = t3.y;
#pragma unroll
        for (int e = 0; e < 8; ++e) h[e] *= inv_h;
      } else {
        float4 t0 = __ldg(reinterpret_cast<const float4*>(a.H.f + m * a.H.ldf + n));
        float4 t1 = __ldg(reinterpret_cast<const float4*>(a.H.f + m * a.H.ldf + n + 4));
        h[0] = t0.x; h[1] = t0.y; h[2] = t0.z; h[3] = t0.w; h[4] = t1.x; h[5] = t1.y; h[6] = t1.z; h[7] = t1.w;
      }
    }
    if (need_u) {
      if (a.U.hi) {
        uint4 xh = *reinterpret_cast<const uint4*>(a.U.hi + m * a.U.ldh + n);
        uint4 xl = *reinterpret_cast<const uint4*>(a.U.lo + m * a.U.ldh + n);
        float2 t0 = join2(xh.x, xl.x), t1 = join2(xh.y, xl.y), t2 = join2(xh.z, xl.z), t3 = join2(xh.w, xl.w);
        u[0] = t0.x; u[1] = t0.y; u[2] = t1.x; u[3] = t1.y; u[4] = t2.x; u[5] = t2.y; u[6] = t3.x; u[7] = t3.y;
#pragma unroll
        for (int e = 0; e < 8; ++e) u[e] *= inv_u;
      } else {
        float4 t0 = *reinterpret_cast<const float4*>(a.U.f + m * a.U.ldf + n);
        float4 t1 = *reinterpret_cast<const float4*>(a.U.f + m * a.U.ldf + n + 4);
        u[0] = t0.x; u[1] = t0.y; u[2] = t0.z; u[3] = t0.w; u[4] = t1.x; u[5] = t1.y; u[6] = t1.z; u[7] = t1.w;
      }
    }
    if (EPI == EPI_ACCUM) {
      float4 t0 = *reinterpret_cast<const float4*>(a.C.f + m * a.C.ldf + n);
      float4 t1 = *reinterpret_cast<const float4*>(a.C.f + m * a.C.ldf + n + 4);
      cp[0] = t0.x; cp[1] = t0.y; cp[2] = t0.z; cp[3] = t0.w; cp[4] = t1.x; cp[5] = t1.y; cp[6] = t1.z; cp[7] = t1.w;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float acc = 0.f, o2;
#pragma unroll
      for (int k = 0; k < KT; ++k) acc += xa[k] * bw[k][e];
      epi_math<EPI>(a, acc, h[e], u[e], cp[e], b[e], o[e], o2);
      mx = fmaxf(mx, fabsf(o[e]));
    }
    if (a.C.hi) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split2(o[2 * e] * sc, o[2 * e + 1] * sc, hi[e], lo[e]);
      *reinterpret_cast<uint4*>(a.C.hi + m * a.C.ldh + n) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(a.C.lo + m * a.C.ldh + n) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    } else {
      *reinterpret_cast<float4*>(a.C.f + m * a.C.ldf + n) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(a.C.f + m * a.C.ldf + n + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
  amax_commit(a.C.amax, mx);
}

// scalar form: any N, unaligned operands
template <int EPI>
__global__ void __launch_bounds__(NDJIR_BLOCK) skinny_k_h_kernel(HArgs a) {
  const long long total = (long long)a.M * a.N;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float sc = dev_scalar(a.C.scale);
  const float inv_h = 1.f / dev_scalar(a.H.scale), inv_u = 1.f / dev_scalar(a.U.scale);
  const bool need_u = (EPI == EPI_MUL_S) && (a.U.f || a.U.hi);
  float mx = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long m = i / a.N;
    const int n = (int)(i - m * a.N);
    float acc = 0.f;
    for (int k = 0; k < a.K; ++k)
      acc += __ldg(a.A32 + m * a.a_rs + (long long)k * a.a_cs) * __ldg(a.B32 + (long long)k * a.b_rs + (long long)n * a.b_cs);
    float h = 0.f, u = 0.f, cp = 0.f, b = 0.f, o, o2;
    if (EPI == EPI_MUL_S) h = op_load(a.H, inv_h, m, n);
    if (need_u) u = op_load(a.U, inv_u, m, n);
    if (EPI == EPI_ACCUM) cp = a.C.f[m * a.C.ldf + n];
    if ((EPI == EPI_BIAS || EPI == EPI_SOFTPLUS) && a.bias) b = __ldg(a.bias + n);
    epi_math<EPI>(a, acc, h, u, cp, b, o, o2);
    mx = fmaxf(mx, fabsf(o));
    op_store(a.C, sc, m, n, o);
  }
  amax_commit(a.C.amax, mx);
}

// ---------------------------------------------------------------------------------------------------------------
// C[m, n] += alpha * sum_k A(k, m) * B32(k, n), N <= 8.  A block walks a slab of the k rows in steps of RB rows; its
// 256 threads are (row lane, pair of adjacent m): a warp reads 128 contiguous bytes of each plane per row.  The narrow
// B rows of a step are staged in shared memory once (instead of NT broadcast loads per thread and row) and all A loads
// of a step are issued before any is used.  Row lanes are folded through shared memory, one atomic per (m, n) and block.
template <int NT>
__global__ void __launch_bounds__(NDJIR_BLOCK) skinny_w_h_kernel(HArgs a) {
  constexpr int RB = 64;                               // rows per step
  __shared__ float red[NDJIR_BLOCK * 2 * NT];
  __shared__ float sb[RB * NT];
  const int pairs = (a.M + 1) / 2;                     // <= blockDim.x (checked by the launcher)
  const int lanes = blockDim.x / pairs;
  const int pr = threadIdx.x % pairs, rl = threadIdx.x / pairs;
  const int m = pr * 2;
  const long long per = ((a.K + gridDim.x - 1) / gridDim.x + RB - 1) / RB * RB;
  const long long k0 = (long long)blockIdx.x * per, k1 = k0 + per < a.K ? k0 + per : a.K;
  const float inv_a = 1.f / dev_scalar(a.a_scale);
  float acc0[NT], acc1[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) acc0[n] = acc1[n] = 0.f;
  const bool pair = (m + 1 < a.M) && (a.lda % 2 == 0);
  constexpr int UNR = 8;
  for (long long kb = k0; kb < k1; kb += RB) {
    __syncthreads();
    for (int i = threadIdx.x; i < RB * NT; i += blockDim.x) {
      const long long k = kb + i / NT;
      const int n = i % NT;
      sb[i] = (k < k1 && n < a.N) ? __ldg(a.B32 + k * a.b_rs + (long long)n * a.b_cs) : 0.f;
    }
    __syncthreads();
    if (rl < lanes) {
      for (int r0 = rl; r0 < RB; r0 += lanes * UNR) {
        uint32_t xh[UNR], xl[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int rr = r0 + u * lanes;
          const long long k = kb + rr;
          xh[u] = xl[u] = 0u;
          if (rr < RB && k < k1) {
            if (pair) {
              xh[u] = __ldg(reinterpret_cast<const uint32_t*>(a.Ahi + k * a.lda + m));
              xl[u] = __ldg(reinterpret_cast<const uint32_t*>(a.Alo + k * a.lda + m));
            } else {
              const unsigned short h0 = __half_as_ushort(a.Ahi[k * a.lda + m]), l0 = __half_as_ushort(a.Alo[k * a.lda + m]);
              unsigned short h1 = 0, l1 = 0;
              if (m + 1 < a.M) {
                h1 = __half_as_ushort(a.Ahi[k * a.lda + m + 1]);
                l1 = __half_as_ushort(a.Alo[k * a.lda + m + 1]);
              }
              xh[u] = (uint32_t)h0 | ((uint32_t)h1 << 16);
              xl[u] = (uint32_t)l0 | ((uint32_t)l1 << 16);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int rr = r0 + u * lanes;
          if (rr < RB) {
            const float2 t = join2(xh[u], xl[u]);
#pragma unroll
            for (int n = 0; n < NT; ++n) {
              const float bv = sb[rr * NT + n];
              acc0[n] += t.x * bv;
              acc1[n] += t.y * bv;
            }
          }
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    red[(threadIdx.x * 2) * NT + n] = acc0[n];
    red[(threadIdx.x * 2 + 1) * NT + n] = acc1[n];
  }
  __syncthreads();
  if (rl == 0) {
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      if (n < a.N) {
        float s0 = 0.f, s1 = 0.f;
        for (int l = 0; l < lanes; ++l) {
          s0 += red[((l * pairs + pr) * 2) * NT + n];
          s1 += red[((l * pairs + pr) * 2 + 1) * NT + n];
        }
        if (s0 != 0.f) atomicAdd(a.C.f + (long long)m * a.C.ldf + n, a.alpha * inv_a * s0);
        if (m + 1 < a.M && s1 != 0.f) atomicAdd(a.C.f + (long long)(m + 1) * a.C.ldf + n, a.alpha * inv_a * s1);
      }
    }
  }
}

static inline bool al16s(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int which_corner(const HArgs& a) {
  if (!a.mn && a.N <= 8 && a.K >= 8 && a.K <= 2048 && a.Ahi && a.B32 && a.C.f && (a.epi == EPI_BIAS || a.epi == EPI_ACCUM))
    return 1;
  if (!a.mn && a.K <= 8 && a.A32 && a.B32 && a.epi != EPI_ATOMIC && a.epi != EPI_ADJ) return 2;
  if (a.mn && a.N <= 8 && a.M <= 2 * NDJIR_BLOCK && a.Ahi && a.B32 && a.C.f && a.epi == EPI_ATOMIC) return 3;
  return 0;
}

bool corner_shape(const HArgs& a) { return which_corner(a) != 0; }

template <int NT>
static void launch_skinny_n(const HArgs& a, cudaStream_t st) {
  int vec = al16s(a.Ahi) && al16s(a.Alo) && a.lda % 8 == 0;
  const long long rows_per_block = (long long)SK_WARPS * 4 * SkR<NT>::R;
  long long blocks = (a.M + rows_per_block - 1) / rows_per_block;
  long long cap = (long long)NDJIR_NUM_SMS * 8;
  int grid = (int)(blocks < cap ? blocks : cap);
  size_t smem = (size_t)((a.K + 63) / 64) * 64 * NT * sizeof(float);
  if (a.epi == EPI_BIAS) skinny_n_h_kernel<NT, EPI_BIAS><<<grid, SK_WARPS * 32, smem, st>>>(a, vec);
  else skinny_n_h_kernel<NT, EPI_ACCUM><<<grid, SK_WARPS * 32, smem, st>>>(a, vec);
}

int launch_corner(const HArgs& a, cudaStream_t st) {
  const int w = which_corner(a);
  if (w == 1) {
    if (a.N == 1) launch_skinny_n<1>(a, st);
    else if (a.N == 2) launch_skinny_n<2>(a, st);
    else if (a.N <= 4) launch_skinny_n<4>(a, st);
    else launch_skinny_n<8>(a, st);
  } else if (w == 2) {
    auto ok_op = [](const Op& o) {
      if (o.hi) return al16s(o.hi) && al16s(o.lo) && o.ldh % 8 == 0;
      if (o.f) return al16s(o.f) && o.ldf % 4 == 0;
      return true;
    };
    if (a.epi == EPI_ACCUM && !a.C.f) return NDJIR_ERR_ARG;
    if (a.epi == EPI_MUL_S && !a.H.f && !a.H.hi) return NDJIR_ERR_ARG;
    const bool vec = a.N % 8 == 0 && NDJIR_BLOCK % (a.N / 8) == 0 && ok_op(a.C) && ok_op(a.H) && ok_op(a.U);
    if (vec) {
      const int rows_per_block = NDJIR_BLOCK / (a.N / 8);
      long long blocks = (a.M + rows_per_block - 1) / rows_per_block;
      long long cap = (long long)NDJIR_NUM_SMS * 16;
      int grid = (int)(blocks < cap ? blocks : cap);
#define NDJIR_SK(E, KT) skinny_k_h_vec_kernel<E, KT><<<grid, NDJIR_BLOCK, 0, st>>>(a)
#define NDJIR_SK_K(E)                                     \
  if (a.K <= 1) NDJIR_SK(E, 1); else if (a.K <= 2) NDJIR_SK(E, 2); else if (a.K <= 3) NDJIR_SK(E, 3); \
  else if (a.K <= 4) NDJIR_SK(E, 4); else if (a.K <= 6) NDJIR_SK(E, 6); else NDJIR_SK(E, 8)
      switch (a.epi) {
        case EPI_BIAS: NDJIR_SK_K(EPI_BIAS); break;
        case EPI_SOFTPLUS: NDJIR_SK_K(EPI_SOFTPLUS); break;
        case EPI_ACCUM: NDJIR_SK_K(EPI_ACCUM); break;
        case EPI_MUL_S: NDJIR_SK_K(EPI_MUL_S); break;
        default: return NDJIR_ERR_ARG;
      }
#undef NDJIR_SK_K
#undef NDJIR_SK
    } else {
      int grid = grid_for((long long)a.M * a.N);
      switch (a.epi) {
        case EPI_BIAS: skinny_k_h_kernel<EPI_BIAS><<<grid, NDJIR_BLOCK, 0, st>>>(a); break;
        case EPI_SOFTPLUS: skinny_k_h_kernel<EPI_SOFTPLUS><<<grid, NDJIR_BLOCK, 0, st>>>(a); break;
        case EPI_ACCUM: skinny_k_h_kernel<EPI_ACCUM><<<grid, NDJIR_BLOCK, 0, st>>>(a); break;
        case EPI_MUL_S: skinny_k_h_kernel<EPI_MUL_S><<<grid, NDJIR_BLOCK, 0, st>>>(a); break;
        default: return NDJIR_ERR_ARG;
      }
    }
  } else if (w == 3) {
    if ((a.M + 1) / 2 > NDJIR_BLOCK) return NDJIR_ERR_ARG;
    long long maxg = (a.K + 63) / 64;
    long long want = (long long)NDJIR_NUM_SMS * 8;
    int grid = (int)(want > maxg ? maxg : want);
    if (grid < 1) grid = 1;
    if (a.N == 1) skinny_w_h_kernel<1><<<grid, NDJIR_BLOCK, 0, st>>>(a);
    else if (a.N == 2) skinny_w_h_kernel<2><<<grid, NDJIR_BLOCK, 0, st>>>(a);
    else if (a.N <= 4) skinny_w_h_kernel<4><<<grid, NDJIR_BLOCK, 0, st>>>(a);
    else skinny_w_h_kernel<8><<<grid, NDJIR_BLOCK, 0, st>>>(a);
  } else {
    return NDJIR_ERR_ARG;
  }
  NDJIR_RETURN_LAST_ERROR();
}

}  // namespace gemmh
}  // namespace ndjir

// ---------------------------------------------------------------------------------------------------------------
// format conversion and scales
// ---------------------------------------------------------------------------------------------------------------
namespace {
using namespace ndjir::gemmh;

struct HM {
  __half* hi; __half* lo; long long ld; const float* scale; float* amax;
};
HM to_hm(const ndjir_hmat* h) {
  HM m;
  m.hi = reinterpret_cast<__half*>(h->hi); m.lo = reinterpret_cast<__half*>(h->lo); m.ld = h->ld;
  m.scale = h->scale; m.amax = h->amax;
  return m;
}

__global__ void __launch_bounds__(NDJIR_BLOCK)
pack_h_kernel(long long rows, int cols, const float* __restrict__ src, long long ld_src, int rep, float alpha, HM d,
              int vec) {
  const int CPT = vec ? 8 : 1;
  const int ncol = (cols + CPT - 1) / CPT;
  const long long total = rows * ncol;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float sc = dev_scalar(d.scale);
  float mx = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / ncol;
    const int c = (int)(i - r * ncol) * CPT;
    const float* sp = src + (r / rep) * ld_src + c;
    if (vec && c + 8 > cols) {
      // ragged last group of a row (cols % 8 != 0): element by element, nothing written beyond the view's width
      for (int e = 0; c + e < cols; ++e) {
        const float x = alpha * __ldg(sp + e);
        mx = fmaxf(mx, fabsf(x));
        __half h, l;
        split1(x * sc, h, l);
        d.hi[r * d.ld + c + e] = h;
        d.lo[r * d.ld + c + e] = l;
      }
    } else if (vec) {
      float4 t0 = __ldg(reinterpret_cast<const float4*>(sp)), t1 = __ldg(reinterpret_cast<const float4*>(sp + 4));
      float x[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 8; ++e) { x[e] *= alpha; mx = fmaxf(mx, fabsf(x[e])); }
#pragma unroll
      for (int e = 0; e < 4; ++e) split2(x[2 * e] * sc, x[2 * e + 1] * sc, hi[e], lo[e]);
      *reinterpret_cast<uint4*>(d.hi + r * d.ld + c) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(d.lo + r * d.ld + c) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    } else {
      float x = alpha * __ldg(sp);
      mx = fmaxf(mx, fabsf(x));
      __half h, l;
      split1(x * sc, h, l);
      d.hi[r * d.ld + c] = h;
      d.lo[r * d.ld + c] = l;
    }
  }
  amax_commit(d.amax, mx);
}

__global__ void __launch_bounds__(NDJIR_BLOCK)
unpack_h_kernel(long long rows, int cols, HM s, float* __restrict__ dst, long long ld_dst) {
  const long long total = rows * cols;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float inv = 1.f / dev_scalar(s.scale);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    dst[r * ld_dst + c] = (__half2float(s.hi[r * s.ld + c]) + __half2float(s.lo[r * s.ld + c])) * inv;
  }
}

__global__ void __launch_bounds__(NDJIR_BLOCK)
copy2d_h_kernel(long long rows, int cols, HM d, HM s, int rep, int vec) {
  const int CPT = vec ? 8 : 1;
  const int ncol = (cols + CPT - 1) / CPT;
  const long long total = rows * ncol;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / ncol;
    const int c = (int)(i - r * ncol) * CPT;
    const long long so = (r / rep) * s.ld + c, dof = r * d.ld + c;
    if (vec) {
      *reinterpret_cast<uint4*>(d.hi + dof) = __ldg(reinterpret_cast<const uint4*>(s.hi + so));
      *reinterpret_cast<uint4*>(d.lo + dof) = __ldg(reinterpret_cast<const uint4*>(s.lo + so));
    } else {
      d.hi[dof] = s.hi[so];
      d.lo[dof] = s.lo[so];
    }
  }
}

// out[c] += alpha * sum_r src[r, c]: a thread owns two adjacent columns, blockIdx.y a slab of rows
__global__ void __launch_bounds__(128)
colsum_h_kernel(long long rows, int cols, float* __restrict__ out, HM s, float alpha) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = (long long)blockIdx.y * per, r1 = r0 + per < rows ? r0 + per : rows;
  if (c >= cols) return;
  const bool pair = c + 1 < cols;
  float a0 = 0.f, a1 = 0.f;
  if (pair && s.ld % 2 == 0) {
#pragma unroll 8
    for (long long r = r0; r < r1; ++r) {
      uint32_t xh = __ldg(reinterpret_cast<const uint32_t*>(s.hi + r * s.ld + c));
      uint32_t xl = __ldg(reinterpret_cast<const uint32_t*>(s.lo + r * s.ld + c));
      float2 t = join2(xh, xl);
      a0 += t.x; a1 += t.y;
    }
  } else {
    for (long long r = r0; r < r1; ++r) {
      a0 += __half2float(s.hi[r * s.ld + c]) + __half2float(s.lo[r * s.ld + c]);
      if (pair) a1 += __half2float(s.hi[r * s.ld + c + 1]) + __half2float(s.lo[r * s.ld + c + 1]);
    }
  }
  const float f = alpha / dev_scalar(s.scale);
  if (a0 != 0.f) atomicAdd(out + c, f * a0);
  if (pair && a1 != 0.f) atomicAdd(out + c + 1, f * a1);
}

__global__ void __launch_bounds__(NDJIR_BLOCK) amax_kernel(long long n, const float* __restrict__ x, float* amax) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  float mx = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) mx = fmaxf(mx, fabsf(__ldg(x + i)));
  amax_commit(amax, mx);
}

__global__ void scale_update_kernel(int n, float* scales, float* amax, int* flags, int target_log2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float m = amax[i];
  amax[i] = 0.f;
  if (!(m > 0.f) || !isfinite(m)) return;
  float old = scales[i];
  int fl = 0;
  if (m * old > H16_MAX) fl |= 1;
  int e;
  frexpf(m, &e);                   // m = f * 2^e, f in [0.5, 1)  ->  m * 2^(target - e) in [2^(target-1), 2^target)
  float s = ldexpf(1.f, target_log2 - e);
  if (s != old) fl |= 2;
  scales[i] = s;
  if (fl && flags) atomicOr(flags, fl);
}
}  // namespace

extern "C" {

int ndjir_pack_h(long long rows, int cols, const float* src, long long ld_src, int rep, float alpha,
                 const ndjir_hmat* dst, cudaStream_t stream) {
  if (rows == 0 || cols == 0) return NDJIR_OK;
  if (rows < 0 || cols < 0 || !src || !dst || !dst->hi || !dst->lo || rep < 1) return NDJIR_ERR_ARG;
  HM d = to_hm(dst);
  // groups of 8 columns move as 16-byte accesses; a ragged last group (cols % 8 != 0) is handled inside the same launch
  int vec = cols >= 8 && al16s(src) && ld_src % 4 == 0 && al16s(d.hi) && al16s(d.lo) && d.ld % 8 == 0;
  pack_h_kernel<<<ndjir::grid_for(rows * (vec ? (cols + 7) / 8 : cols)), NDJIR_BLOCK, 0, stream>>>(rows, cols, src, ld_src,
                                                                                                  rep, alpha, d, vec);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_unpack_h(long long rows, int cols, const ndjir_hmat* src, float* dst, long long ld_dst, cudaStream_t stream) {
  if (rows == 0 || cols == 0) return NDJIR_OK;
  if (rows < 0 || cols < 0 || !src || !src->hi || !src->lo || !dst) return NDJIR_ERR_ARG;
  unpack_h_kernel<<<ndjir::grid_for(rows * cols), NDJIR_BLOCK, 0, stream>>>(rows, cols, to_hm(src), dst, ld_dst);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_copy2d_h(long long rows, int cols, const ndjir_hmat* dst, const ndjir_hmat* src, int rep,
                   cudaStream_t stream) {
  if (rows == 0 || cols == 0) return NDJIR_OK;
  if (rows < 0 || cols < 0 || !src || !dst || !src->hi || !dst->hi || rep < 1) return NDJIR_ERR_ARG;
  HM d = to_hm(dst), s = to_hm(src);
  int vec = cols % 8 == 0 && al16s(d.hi) && al16s(d.lo) && al16s(s.hi) && al16s(s.lo) && d.ld % 8 == 0 && s.ld % 8 == 0;
  copy2d_h_kernel<<<ndjir::grid_for(rows * (vec ? cols / 8 : cols)), NDJIR_BLOCK, 0, stream>>>(rows, cols, d, s, rep, vec);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_colsum_h(long long rows, int cols, float* out, const ndjir_hmat* src, float alpha, cudaStream_t stream) {
  if (rows == 0 || cols == 0) return NDJIR_OK;
  if (rows < 0 || cols < 0 || !src || !src->hi || !out) return NDJIR_ERR_ARG;
  int gx = ((cols + 1) / 2 + 127) / 128;
  long long want = (long long)NDJIR_NUM_SMS * 16 / gx;
  long long maxy = (rows + 63) / 64;
  int gy = (int)(want < 1 ? 1 : (want > maxy ? maxy : want));
  colsum_h_kernel<<<dim3(gx, gy), 128, 0, stream>>>(rows, cols, out, to_hm(src), alpha);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_amax(long long n, const float* x, float* amax, cudaStream_t stream) {
  if (n == 0) return NDJIR_OK;
  if (n < 0 || !x || !amax) return NDJIR_ERR_ARG;
  amax_kernel<<<ndjir::grid_for(n, NDJIR_BLOCK, 8), NDJIR_BLOCK, 0, stream>>>(n, x, amax);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_scale_update(int n_slots, float* scales, float* amax, int* flags, int target_log2, cudaStream_t stream) {
  if (n_slots == 0) return NDJIR_OK;
  if (n_slots < 0 || !scales || !amax) return NDJIR_ERR_ARG;
  scale_update_kernel<<<(n_slots + 255) / 256, 256, 0, stream>>>(n_slots, scales, amax, flags, target_log2);
  NDJIR_RETURN_LAST_ERROR();
}

}  // extern "C"
