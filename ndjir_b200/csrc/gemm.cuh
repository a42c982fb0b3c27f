// Dense linear-algebra engine of the fused MLP pipeline (sm_100a).
//
// The reference composes its nine MLPs from nnabla `PF.affine` + `F.softplus` ops (python/network.py:84-93,
// 154-232), which nnabla lowers to cuBLAS SGEMM + one elementwise kernel per op and a separate kernel for every
// backward product.  Here every product of the forward, input-gradient, tangent and weight-gradient passes goes
// through ONE tiled kernel with a fused prologue-free / epilogue-rich design: bias, softplus(beta=100), the
// sigmoid factor of the activation derivative, the second-order (double-backward) term and accumulation are
// applied while the accumulator tile is still in registers, so an activation matrix is written once and read
// once per pass.
//
// This file is the exact-fp32 FFMA path (parity mode, 1e-5 forward tolerance of BASELINE.json).
#pragma once
#include "common.cuh"

namespace ndjir {
namespace gemm {

enum Epi {
  EPI_BIAS = 0,      // C = alpha*acc + bias
  EPI_SOFTPLUS = 1,  // C = out_scale * softplus_beta(acc + bias)
  EPI_ACCUM = 2,     // C += alpha*acc
  EPI_MUL_S = 3,     // C = alpha*acc * s(H) [+ U]                           s(h) = 1 - exp(-beta*hscale*h)
  EPI_ADJ = 4,       // C = acc * U * beta * (1 - s(H)) ; C2 = out_scale * acc * s(H)   (adjoint of the normal pass)
  EPI_ATOMIC = 5     // atomicAdd(C, alpha*acc)                              (split-K weight gradients)
};

struct Args {
  int M, N, K;
  const float* A; long long a_rs, a_cs;  // A(m,k) = A[m*a_rs + k*a_cs]
  const float* B; long long b_rs, b_cs;  // B(k,n) = B[k*b_rs + n*b_cs]
  const float* B_lo;                     // optional pre-split lo part of B (same strides): x - tf32(x), or nullptr
  float* C; long long ldc;               // C(m,n) = C[m*ldc + n]
  const float* bias;                     // [N] or nullptr
  float alpha;                           // scale on the accumulator
  float out_scale;                       // scale on the stored activation (1/sqrt(2) at the skip layer)
  float beta;                            // softplus beta
  const float* H; long long ldh; float hscale;  // activation the sigmoid factor is derived from
  const float* U; long long ldu;         // EPI_MUL_S addend / EPI_ADJ factor
  float* C2; long long ldc2;             // EPI_ADJ second output
  int split_k;                           // >1: K is split over gridDim.z (EPI_ATOMIC only)
  float* colsum;                         // optional [N]: += column sums of B over k (bias gradient of a wgrad product)
};

static inline Args make_args(int M, int N, int K) {
  Args a;
  a.M = M; a.N = N; a.K = K;
  a.A = nullptr; a.a_rs = 0; a.a_cs = 1;
  a.B = nullptr; a.b_rs = 0; a.b_cs = 1; a.B_lo = nullptr;
  a.C = nullptr; a.ldc = 0; a.bias = nullptr;
  a.alpha = 1.f; a.out_scale = 1.f; a.beta = 100.f;
  a.H = nullptr; a.ldh = 0; a.hscale = 1.f;
  a.U = nullptr; a.ldu = 0;
  a.C2 = nullptr; a.ldc2 = 0; a.split_k = 1; a.colsum = nullptr;
  return a;
}

// softplus(x, beta) = log(1 + exp(beta x)) / beta, evaluated without overflow
__device__ __forceinline__ float softplus_beta(float x, float beta) {
  float z = beta * x;
  float r = fmaxf(z, 0.f) + log1pf(expf(-fabsf(z)));
  return r / beta;
}
// sigmoid(beta * a) recovered from h = softplus(a, beta): 1 - exp(-beta h)
__device__ __forceinline__ float sig_from_softplus(float h, float beta) { return -expm1f(-beta * h); }

template <int EPI>
__device__ __forceinline__ void epilogue_store(const Args& a, int m, int n, float acc) {
  long long ci = (long long)m * a.ldc + n;
  if (EPI == EPI_BIAS) {
    a.C[ci] = a.alpha * acc + (a.bias ? __ldg(a.bias + n) : 0.f);
  } else if (EPI == EPI_SOFTPLUS) {
    a.C[ci] = a.out_scale * softplus_beta(acc + (a.bias ? __ldg(a.bias + n) : 0.f), a.beta);
  } else if (EPI == EPI_ACCUM) {
    a.C[ci] += a.alpha * acc;
  } else if (EPI == EPI_MUL_S) {
    float s = sig_from_softplus(__ldg(a.H + (long long)m * a.ldh + n) * a.hscale, a.beta);
    float v = a.alpha * acc * s;
    if (a.U) v += __ldg(a.U + (long long)m * a.ldu + n);
    a.C[ci] = v;
  } else if (EPI == EPI_ADJ) {
    float s = sig_from_softplus(__ldg(a.H + (long long)m * a.ldh + n) * a.hscale, a.beta);
    float u = __ldg(a.U + (long long)m * a.ldu + n);
    a.C[ci] = acc * u * a.beta * (1.f - s);
    a.C2[(long long)m * a.ldc2 + n] = a.out_scale * acc * s;
  } else if (EPI == EPI_ATOMIC) {
    atomicAdd(a.C + ci, a.alpha * acc);
  }
}

// fast-math variants for the tensor-core epilogue (MUFU ex2 / lg2): absolute error ~1e-6 on beta*x, i.e. ~1e-8 on
// the activation (beta = 100), far inside the 1e-5 forward bar; the FFMA parity path keeps the exact versions
__device__ __forceinline__ float softplus_beta_fast(float x, float beta) {
  float z = beta * x;
  float r = fmaxf(z, 0.f) + __logf(1.f + __expf(-fabsf(z)));
  return __fdividef(r, beta);
}
__device__ __forceinline__ float sig_from_softplus_fast(float h, float beta) { return 1.f - __expf(-beta * h); }

int launch(const Args& a, int epi, cudaStream_t st);

// tcgen05 path (gemm_tc.cu): 3xTF32 error-compensated products on the 5th-generation tensor cores
extern int g_mlp_tensor_cores;
extern int g_mlp_cta_pair;
extern int g_mlp_dbg;
extern int g_mlp_fused_colsum;
extern int g_mlp_presplit;    // 1: a B operand that comes with its lo part is fetched as raw + lo TMA tiles
extern int g_mlp_mask_hi;     // 1: the transform warps also clear the low 13 mantissa bits of the raw tiles
bool tc_eligible(const Args& a, int epi);
int launch_tc(const Args& a, int epi, cudaStream_t st);
// memory-bound corner shapes (gemm_skinny.cu): N <= 8 outputs, K <= 8 rank updates, weight gradients with N <= 8
bool skinny_eligible(const Args& a, int epi);
int launch_skinny(const Args& a, int epi, cudaStream_t st);

}  // namespace gemm
}  // namespace ndjir
