// Tiled fp32 GEMM with fused MLP epilogues (see gemm.cuh).  128 x {128,32} x 16 CTA tile, 256 threads,
// 8 x {8,2} register tile per thread, operands staged through shared memory with 16-byte global loads when
// the operand is aligned (activations always are; nnabla-layout weights with odd widths fall back to scalar
// loads - they are small and L2-resident).
#include "gemm.cuh"
#include "../../include/ndjir_b200.h"

namespace ndjir {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int TM = 8;
constexpr int APAD = 4;

// A_KC: A is contiguous along k (row-major activations); otherwise contiguous along m (transposed use, X^T dY).
// B_NC: B is contiguous along n (nnabla (in,out) weights used as-is); otherwise contiguous along k (W^T use).
template <int BN, int TN, int EPI, bool A_KC, bool B_NC>
__global__ void __launch_bounds__(256) gemm_kernel(Args a, int vecA, int vecB) {
  __shared__ __align__(16) float As[BK][BM + APAD];
  __shared__ __align__(16) float Bs[BK][BN + APAD];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int k_begin = 0, k_end = a.K;
  if (a.split_k > 1) {
    int kc = ((a.K + a.split_k - 1) / a.split_k + BK - 1) / BK * BK;
    k_begin = blockIdx.z * kc;
    k_end = min(a.K, k_begin + kc);
    if (k_begin >= k_end) return;
  }
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    // ---- stage A tile (BM x BK) as As[k][m] ----
    if (A_KC) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int r = (tid >> 2) + h * 64, kv = (tid & 3) * 4;
        int m = m0 + r, k = k0 + kv;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (m < a.M) {
          const float* p = a.A + (long long)m * a.a_rs + (long long)k * a.a_cs;
          if (vecA && k + 3 < k_end) {
            float4 t = __ldg(reinterpret_cast<const float4*>(p));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (k + j < k_end) v[j] = __ldg(p + (long long)j * a.a_cs);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) As[kv + j][r] = v[j];
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int kk = (tid >> 5) + h * 8, mv = (tid & 31) * 4;
        int k = k0 + kk, m = m0 + mv;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (k < k_end) {
          const float* p = a.A + (long long)m * a.a_rs + (long long)k * a.a_cs;
          if (vecA && m + 3 < a.M) {
            float4 t = __ldg(reinterpret_cast<const float4*>(p));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (m + j < a.M) v[j] = __ldg(p + (long long)j * a.a_rs);
          }
        }
        *reinterpret_cast<float4*>(&As[kk][mv]) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    // ---- stage B tile (BK x BN) as Bs[k][n] ----
    if (B_NC) {
      constexpr int VPR = BN / 4;               // float4 per k-row
      constexpr int ROWS_PER_PASS = 256 / VPR;  // 8 (BN=128) or 32 (BN=32)
#pragma unroll
      for (int h = 0; h < (BK + ROWS_PER_PASS - 1) / ROWS_PER_PASS; ++h) {
        int kk = tid / VPR + h * ROWS_PER_PASS, nv = (tid % VPR) * 4;
        if (kk < BK) {
          int k = k0 + kk, n = n0 + nv;
          float v[4] = {0.f, 0.f, 0.f, 0.f};
          if (k < k_end) {
            const float* p = a.B + (long long)k * a.b_rs + (long long)n * a.b_cs;
            if (vecB && n + 3 < a.N) {
              float4 t = __ldg(reinterpret_cast<const float4*>(p));
              v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) if (n + j < a.N) v[j] = __ldg(p + (long long)j * a.b_cs);
            }
          }
          *reinterpret_cast<float4*>(&Bs[kk][nv]) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    } else {
      constexpr int PASSES = (BN * BK / 4 + 255) / 256;  // 2 (BN=128) or 1 (BN=32, half the threads)
#pragma unroll
      for (int h = 0; h < PASSES; ++h) {
        int c = (tid >> 2) + h * 64, kv = (tid & 3) * 4;
        if (c < BN) {
          int n = n0 + c, k = k0 + kv;
          float v[4] = {0.f, 0.f, 0.f, 0.f};
          if (n < a.N) {
            const float* p = a.B + (long long)k * a.b_rs + (long long)n * a.b_cs;
            if (vecB && k + 3 < k_end) {
              float4 t = __ldg(reinterpret_cast<const float4*>(p));
              v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) if (k + j < k_end) v[j] = __ldg(p + (long long)j * a.b_rs);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) Bs[kv + j][c] = v[j];
        }
      }
    }
    __syncthreads();
    // ---- 8 x TN outer products per k ----
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
      if constexpr (TN == 8) {
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][(BN / 2) + tx * 4]);
        bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
        bv[TN - 4] = b1.x; bv[TN - 3] = b1.y; bv[TN - 2] = b1.z; bv[TN - 1] = b1.w;
      } else {
        float2 b0 = *reinterpret_cast<const float2*>(&Bs[kk][tx * 2]);
        bv[0] = b0.x; bv[TN - 1] = b0.y;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- fused epilogue ----
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n;
      if constexpr (TN == 8) n = n0 + (j < 4 ? tx * 4 + j : (BN / 2) + tx * 4 + (j - 4));
      else n = n0 + tx * 2 + j;
      if (n < a.N) epilogue_store<EPI>(a, m, n, acc[i][j]);
    }
  }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int BN, int TN, int EPI>
static void dispatch_layout(const Args& a, dim3 grid, cudaStream_t st) {
  bool a_kc = (a.a_cs == 1), b_nc = (a.b_cs == 1);
  // vector-load eligibility of each operand along its contiguous dimension
  int vecA = a_kc ? (al16(a.A) && a.a_rs % 4 == 0) : (a.a_rs == 1 && al16(a.A) && a.a_cs % 4 == 0);
  int vecB = b_nc ? (al16(a.B) && a.b_rs % 4 == 0) : (a.b_rs == 1 && al16(a.B) && a.b_cs % 4 == 0);
  if (a_kc && b_nc) gemm_kernel<BN, TN, EPI, true, true><<<grid, 256, 0, st>>>(a, vecA, vecB);
  else if (a_kc && !b_nc) gemm_kernel<BN, TN, EPI, true, false><<<grid, 256, 0, st>>>(a, vecA, vecB);
  else if (!a_kc && b_nc) gemm_kernel<BN, TN, EPI, false, true><<<grid, 256, 0, st>>>(a, vecA, vecB);
  else gemm_kernel<BN, TN, EPI, false, false><<<grid, 256, 0, st>>>(a, vecA, vecB);
}

template <int EPI>
static void dispatch_tile(const Args& a, cudaStream_t st) {
  int split = a.split_k > 1 ? a.split_k : 1;
  if (a.N > 32) {
    dim3 grid((a.N + 127) / 128, (a.M + BM - 1) / BM, split);
    dispatch_layout<128, 8, EPI>(a, grid, st);
  } else {
    dim3 grid((a.N + 31) / 32, (a.M + BM - 1) / BM, split);
    dispatch_layout<32, 2, EPI>(a, grid, st);
  }
}

int g_mlp_fused_colsum = 1;   // ndjir_wgrad_bias: bias gradient carried by the weight-gradient product's transform warps
int g_mlp_tensor_cores = 1;   // 1: tcgen05 3xTF32 path where shapes allow (gemm_tc.cu), 0: fp32 FFMA everywhere

int launch(const Args& a, int epi, cudaStream_t st) {
  if (a.M <= 0 || a.N <= 0) return NDJIR_OK;
  if (a.K < 0 || !a.A || !a.B || !a.C) return NDJIR_ERR_ARG;
  if (a.a_cs != 1 && a.a_rs != 1) return NDJIR_ERR_ARG;
  if (a.b_cs != 1 && a.b_rs != 1) return NDJIR_ERR_ARG;
  if (a.split_k > 1 && epi != EPI_ATOMIC) return NDJIR_ERR_ARG;
  if (g_mlp_tensor_cores && skinny_eligible(a, epi)) return launch_skinny(a, epi, st);
  if (g_mlp_tensor_cores && tc_eligible(a, epi)) return launch_tc(a, epi, st);
  switch (epi) {
    case EPI_BIAS: dispatch_tile<EPI_BIAS>(a, st); break;
    case EPI_SOFTPLUS: dispatch_tile<EPI_SOFTPLUS>(a, st); break;
    case EPI_ACCUM: dispatch_tile<EPI_ACCUM>(a, st); break;
    case EPI_MUL_S: if (!a.H) return NDJIR_ERR_ARG; dispatch_tile<EPI_MUL_S>(a, st); break;
    case EPI_ADJ: if (!a.H || !a.U || !a.C2) return NDJIR_ERR_ARG; dispatch_tile<EPI_ADJ>(a, st); break;
    case EPI_ATOMIC: dispatch_tile<EPI_ATOMIC>(a, st); break;
    default: return NDJIR_ERR_ARG;
  }
  NDJIR_RETURN_LAST_ERROR();
}

}  // namespace gemm
}  // namespace ndjir

// Building block of the MLP passes: C = epilogue(A(MxK) * B(KxN)) with explicit element strides.
extern "C" int ndjir_gemm(int M, int N, int K, const float* A, long long a_rs, long long a_cs, const float* B,
                          long long b_rs, long long b_cs, float* C, long long ldc, const float* bias, float alpha,
                          float out_scale, float beta, const float* H, long long ldh, float hscale, const float* U,
                          long long ldu, float* C2, long long ldc2, int split_k, int epilogue,
                          cudaStream_t stream) {
  ndjir::gemm::Args a = ndjir::gemm::make_args(M, N, K);
  a.A = A; a.a_rs = a_rs; a.a_cs = a_cs; a.B = B; a.b_rs = b_rs; a.b_cs = b_cs; a.C = C; a.ldc = ldc;
  a.bias = bias; a.alpha = alpha; a.out_scale = out_scale; a.beta = beta; a.H = H; a.ldh = ldh; a.hscale = hscale;
  a.U = U; a.ldu = ldu; a.C2 = C2; a.ldc2 = ldc2; a.split_k = split_k;
  return ndjir::gemm::launch(a, epilogue, stream);
}

// Same product with the pre-split lo part of the weight operand supplied by the caller (B_lo may be NULL).
extern "C" int ndjir_gemm_presplit(int M, int N, int K, const float* A, long long a_rs, long long a_cs, const float* B,
                                   const float* B_lo, long long b_rs, long long b_cs, float* C, long long ldc,
                                   const float* bias, float alpha, float out_scale, float beta, const float* H,
                                   long long ldh, float hscale, const float* U, long long ldu, float* C2,
                                   long long ldc2, int split_k, int epilogue, cudaStream_t stream) {
  ndjir::gemm::Args a = ndjir::gemm::make_args(M, N, K);
  a.A = A; a.a_rs = a_rs; a.a_cs = a_cs; a.B = B; a.B_lo = B_lo; a.b_rs = b_rs; a.b_cs = b_cs; a.C = C; a.ldc = ldc;
  a.bias = bias; a.alpha = alpha; a.out_scale = out_scale; a.beta = beta; a.H = H; a.ldh = ldh; a.hscale = hscale;
  a.U = U; a.ldu = ldu; a.C2 = C2; a.ldc2 = ldc2; a.split_k = split_k;
  return ndjir::gemm::launch(a, epilogue, stream);
}

// Weight gradient and bias gradient of one layer in one call: gW (K_in, N) += A(rows, K_in)^T dZ(rows, N) (split-K with
// atomic accumulation) and gb (N) += column sums of dZ.  On the tcgen05 path the transform warps, which read every dZ
// tile anyway, carry the column sums (no second pass over dZ); otherwise the product is followed by ndjir_colsum.
extern "C" int ndjir_colsum(long long rows, int cols, float* out, const float* src, long long ld_src, float alpha,
                            cudaStream_t stream);
extern "C" int ndjir_wgrad_bias(long long rows, int K_in, int N, const float* A, long long lda, const float* dZ,
                                long long ldz, float* gW, long long ldw, float* gb, int split_k, cudaStream_t stream) {
  if (rows == 0 || K_in <= 0 || N <= 0) return NDJIR_OK;
  if (rows < 0 || rows > 0x7fffffffll || !A || !dZ || !gW) return NDJIR_ERR_ARG;
  using namespace ndjir::gemm;
  Args a = make_args(K_in, N, (int)rows);
  a.A = A; a.a_rs = 1; a.a_cs = lda; a.B = dZ; a.b_rs = ldz; a.b_cs = 1; a.C = gW; a.ldc = ldw;
  a.split_k = split_k > 1 ? split_k : 1;
  const bool fused = gb && g_mlp_tensor_cores && g_mlp_fused_colsum && !skinny_eligible(a, EPI_ATOMIC) &&
                     tc_eligible(a, EPI_ATOMIC);
  if (fused) a.colsum = gb;
  int rc = launch(a, EPI_ATOMIC, stream);
  if (rc != NDJIR_OK || !gb || fused) return rc;
  return ndjir_colsum(rows, N, gb, dZ, ldz, 1.0f, stream);
}

