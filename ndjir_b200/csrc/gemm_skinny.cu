// Memory-bound corner shapes of the MLP passes (sm_100a).  The tiled kernels (gemm.cu / gemm_tc.cu) waste a 32- or
// 256-wide tile on them; these three kernels stream the one large operand exactly once:
//   skinny_n   : C (M x N<=8)  = A (M x K) B          last layers (sdf, colours, roughness ...) and the 4-wide grid
//                                                      feature gradient: one warp per row, shuffle reduction
//   skinny_k   : C (M x N)     = A (M x K<=8) B        rank-1..8 updates with the fused epilogue: element-wise
//   skinny_w   : C (M x N<=8) += A^T (M x P) B (P x N) weight gradients of those last layers: one thread per weight
//                                                      row, rows of A read coalesced, split over the sample axis
#include "gemm.cuh"
#include "../../include/ndjir_b200.h"

namespace ndjir {
namespace gemm {

constexpr int SK_WARPS = 8;

template <int NT, int EPI>
__global__ void __launch_bounds__(SK_WARPS * 32) skinny_n_kernel(Args a, int vec) {
  // B staged as [NT][Kp] (k contiguous, Kp = K rounded up to 4): lane l reads the float4 at k = 4 l of every output's
  // row, consecutive lanes consecutive 16 bytes - conflict-free.  (The first version staged [K][NT]: the lanes of a
  // 128-bit shared load were 16 NT bytes apart, a 4-way bank conflict that made this one-pass kernel run at 1.6 TB/s.)
  extern __shared__ __align__(16) float Bs[];
  const int Kp = (a.K + 3) & ~3;
  for (int i = threadIdx.x; i < Kp * NT; i += blockDim.x) {
    int n = i / Kp, k = i - n * Kp;
    Bs[i] = (n < a.N && k < a.K) ? __ldg(a.B + (long long)k * a.b_rs + (long long)n * a.b_cs) : 0.f;
  }
  __syncthreads();
  // FOUR ROWS PER WARP, eight lanes per row: every lane has K/32 independent 16-byte loads in flight (one warp per row
  // had two) and the reduction is three shuffle steps per output instead of five.
  const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
  const long long row0 = ((long long)blockIdx.x * SK_WARPS + (threadIdx.x >> 5)) * 4 + grp;
  const long long row_stride = (long long)gridDim.x * SK_WARPS * 4;
  const int K4 = vec ? (a.K & ~3) : 0;
  const long long rounds = (a.M + row_stride - 1) / row_stride;
  for (long long r = 0; r < rounds; ++r) {
    const long long m = row0 + r * row_stride;
    const bool active = m < a.M;
    const float* row = a.A + (active ? m : a.M - 1) * a.a_rs;
    float acc[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n] = 0.f;
#pragma unroll 4
    for (int k = sub * 4; k < K4; k += 32) {
      float4 x = __ldg(reinterpret_cast<const float4*>(row + k));
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        float4 b = *reinterpret_cast<const float4*>(Bs + n * Kp + k);
        acc[n] += x.x * b.x + x.y * b.y + x.z * b.z + x.w * b.w;
      }
    }
    for (int k = K4 + sub; k < a.K; k += 8) {
      float x = __ldg(row + k);
#pragma unroll
      for (int n = 0; n < NT; ++n) acc[n] += x * Bs[n * Kp + k];
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 1);
      acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 2);
      acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 4);
    }
    if (active && sub < a.N) {
      float v = 0.f;
#pragma unroll
      for (int n = 0; n < NT; ++n) if (sub == n) v = acc[n];
      epilogue_store<EPI>(a, (int)m, sub, v);
    }
  }
}

template <int EPI>
__global__ void __launch_bounds__(NDJIR_BLOCK) skinny_k_kernel(Args a) {
  long long total = (long long)a.M * a.N;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    long long m = i / a.N;
    int n = (int)(i - m * a.N);
    float acc = 0.f;
    for (int k = 0; k < a.K; ++k)
      acc += __ldg(a.A + m * a.a_rs + (long long)k * a.a_cs) * __ldg(a.B + (long long)k * a.b_rs + (long long)n * a.b_cs);
    epilogue_store<EPI>(a, (int)m, n, acc);
  }
}

// 4 columns per thread, 16-byte loads/stores of the epilogue operands (needs N % 4 == 0 and 16-byte aligned rows)
template <int EPI>
__global__ void __launch_bounds__(NDJIR_BLOCK) skinny_k_vec4_kernel(Args a) {
  const int n4 = a.N / 4;
  long long total = (long long)a.M * n4;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    long long m = i / n4;
    int n = (int)(i - m * n4) * 4;
    float4 hv = make_float4(0.f, 0.f, 0.f, 0.f), uv = hv, cv = hv, bv = hv;
    if (EPI == EPI_MUL_S) {
      hv = __ldg(reinterpret_cast<const float4*>(a.H + m * a.ldh + n));
      if (a.U) uv = *reinterpret_cast<const float4*>(a.U + m * a.ldu + n);
    }
    if (EPI == EPI_ACCUM) cv = *reinterpret_cast<const float4*>(a.C + m * a.ldc + n);
    if ((EPI == EPI_BIAS || EPI == EPI_SOFTPLUS) && a.bias) bv = __ldg(reinterpret_cast<const float4*>(a.bias + n));
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < a.K; ++k) {
      float x = __ldg(a.A + m * a.a_rs + (long long)k * a.a_cs);
      const float* bp = a.B + (long long)k * a.b_rs + (long long)n * a.b_cs;   // small operand: L1-resident
      acc[0] += x * __ldg(bp); acc[1] += x * __ldg(bp + a.b_cs); acc[2] += x * __ldg(bp + 2 * a.b_cs);
      acc[3] += x * __ldg(bp + 3 * a.b_cs);
    }
    float h[4] = {hv.x, hv.y, hv.z, hv.w}, u[4] = {uv.x, uv.y, uv.z, uv.w}, co[4] = {cv.x, cv.y, cv.z, cv.w};
    float bb[4] = {bv.x, bv.y, bv.z, bv.w}, o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (EPI == EPI_BIAS) o[e] = a.alpha * acc[e] + bb[e];
      // fast-math activations as in the tensor-core epilogue (these kernels only run on that path): the exact expm1f /
      // log1pf made this memory-bound kernel compute-bound (0.23 ms for 268 M elements against a 0.08-0.12 ms HBM floor)
      else if (EPI == EPI_SOFTPLUS) o[e] = a.out_scale * softplus_beta_fast(acc[e] + bb[e], a.beta);
      else if (EPI == EPI_ACCUM) o[e] = co[e] + a.alpha * acc[e];
      else o[e] = a.alpha * acc[e] * sig_from_softplus_fast(h[e] * a.hscale, a.beta) + u[e];
    }
    *reinterpret_cast<float4*>(a.C + m * a.ldc + n) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// C[m, n] += alpha * sum_k A[m + k*a_cs] * B(k, n);  thread = one m, block walks a slice of k
template <int NT>
__global__ void __launch_bounds__(NDJIR_BLOCK) skinny_w_kernel(Args a) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  long long per = (a.K + gridDim.y - 1) / gridDim.y;
  long long k0 = (long long)blockIdx.y * per, k1 = k0 + per < a.K ? k0 + per : a.K;
  float acc[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) acc[n] = 0.f;
  if (m < a.M) {
    long long k = k0;
    for (; k + 3 < k1; k += 4) {
      float x0 = __ldg(a.A + m + k * a.a_cs), x1 = __ldg(a.A + m + (k + 1) * a.a_cs);
      float x2 = __ldg(a.A + m + (k + 2) * a.a_cs), x3 = __ldg(a.A + m + (k + 3) * a.a_cs);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        if (n < a.N) {
          acc[n] += x0 * __ldg(a.B + k * a.b_rs + (long long)n * a.b_cs) + x1 * __ldg(a.B + (k + 1) * a.b_rs + (long long)n * a.b_cs) +
                    x2 * __ldg(a.B + (k + 2) * a.b_rs + (long long)n * a.b_cs) + x3 * __ldg(a.B + (k + 3) * a.b_rs + (long long)n * a.b_cs);
        }
      }
    }
    for (; k < k1; ++k) {
      float x = __ldg(a.A + m + k * a.a_cs);
#pragma unroll
      for (int n = 0; n < NT; ++n) if (n < a.N) acc[n] += x * __ldg(a.B + k * a.b_rs + (long long)n * a.b_cs);
    }
#pragma unroll
    for (int n = 0; n < NT; ++n)
      if (n < a.N && acc[n] != 0.f) atomicAdd(a.C + (long long)m * a.ldc + n, a.alpha * acc[n]);
  }
}

static inline bool al16s(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool skinny_eligible(const Args& a, int epi) {
  if (a.N <= 8 && a.a_cs == 1 && a.K >= 32 && a.K <= 1024 && a.split_k <= 1 && (epi == EPI_BIAS || epi == EPI_ACCUM))
    return true;
  if (a.K <= 8 && a.split_k <= 1 && epi != EPI_ATOMIC && epi != EPI_ADJ) return true;
  if (a.N <= 8 && a.a_rs == 1 && a.a_cs != 1 && epi == EPI_ATOMIC) return true;
  return false;
}

template <int NT>
static void launch_skinny_n(const Args& a, int epi, cudaStream_t st) {
  int vec = al16s(a.A) && a.a_rs % 4 == 0;
  long long blocks = (a.M + SK_WARPS * 4 - 1) / (SK_WARPS * 4);
  long long cap = (long long)NDJIR_NUM_SMS * 16;
  int grid = (int)(blocks < cap ? blocks : cap);
  size_t smem = (size_t)((a.K + 3) & ~3) * NT * sizeof(float);
  if (epi == EPI_BIAS) skinny_n_kernel<NT, EPI_BIAS><<<grid, SK_WARPS * 32, smem, st>>>(a, vec);
  else skinny_n_kernel<NT, EPI_ACCUM><<<grid, SK_WARPS * 32, smem, st>>>(a, vec);
}

int launch_skinny(const Args& a, int epi, cudaStream_t st) {
  if (a.N <= 8 && a.a_cs == 1 && a.K >= 32 && a.K <= 1024 && a.split_k <= 1 && (epi == EPI_BIAS || epi == EPI_ACCUM)) {
    if (a.N == 1) launch_skinny_n<1>(a, epi, st);
    else if (a.N == 2) launch_skinny_n<2>(a, epi, st);
    else if (a.N <= 4) launch_skinny_n<4>(a, epi, st);
    else launch_skinny_n<8>(a, epi, st);
  } else if (a.K <= 8 && epi != EPI_ATOMIC && epi != EPI_ADJ) {
    auto ok16 = [](const void* q, long long ld) { return q == nullptr || (al16s(q) && ld % 4 == 0); };
    bool vec = a.N % 4 == 0 && ok16(a.C, a.ldc) && ok16(a.H, a.ldh) &&
               ok16(a.U, a.ldu) && ok16(a.bias, 0);
    if (vec) {
      int grid = grid_for((long long)a.M * (a.N / 4));
      switch (epi) {
        case EPI_BIAS: skinny_k_vec4_kernel<EPI_BIAS><<<grid, NDJIR_BLOCK, 0, st>>>(a); break;
        case EPI_SOFTPLUS: skinny_k_vec4_kernel<EPI_SOFTPLUS><<<grid, NDJIR_BLOCK, 0, st>>>(a); break;
        case EPI_ACCUM: skinny_k_vec4_kernel<EPI_ACCUM><<<grid, NDJIR_BLOCK, 0, st>>>(a); break;
        case EPI_MUL_S: if (!a.H) return NDJIR_ERR_ARG; skinny_k_vec4_kernel<EPI_MUL_S><<<grid, NDJIR_BLOCK, 0, st>>>(a); break;
        default: return NDJIR_ERR_ARG;
      }
      NDJIR_RETURN_LAST_ERROR();
    }
    int grid = grid_for((long long)a.M * a.N);
    switch (epi) {
      case EPI_BIAS: skinny_k_kernel<EPI_BIAS><<<grid, NDJIR_BLOCK, 0, st>>>(a); break;
      case EPI_SOFTPLUS: skinny_k_kernel<EPI_SOFTPLUS><<<grid, NDJIR_BLOCK, 0, st>>>(a); break;
      case EPI_ACCUM: skinny_k_kernel<EPI_ACCUM><<<grid, NDJIR_BLOCK, 0, st>>>(a); break;
      case EPI_MUL_S: if (!a.H) return NDJIR_ERR_ARG; skinny_k_kernel<EPI_MUL_S><<<grid, NDJIR_BLOCK, 0, st>>>(a); break;
      default: return NDJIR_ERR_ARG;
    }
  } else {
    int gx = (a.M + NDJIR_BLOCK - 1) / NDJIR_BLOCK;
    long long want = (long long)NDJIR_NUM_SMS * 8 / gx;
    long long maxy = (a.K + 63) / 64;
    int gy = (int)(want < 1 ? 1 : (want > maxy ? maxy : want));
    dim3 grid(gx, gy);
    if (a.N == 1) skinny_w_kernel<1><<<grid, NDJIR_BLOCK, 0, st>>>(a);
    else if (a.N == 2) skinny_w_kernel<2><<<grid, NDJIR_BLOCK, 0, st>>>(a);
    else if (a.N <= 4) skinny_w_kernel<4><<<grid, NDJIR_BLOCK, 0, st>>>(a);
    else skinny_w_kernel<8><<<grid, NDJIR_BLOCK, 0, st>>>(a);
  }
  NDJIR_RETURN_LAST_ERROR();
}

}  // namespace gemm
}  // namespace ndjir
