// Split-fp16 MLP engine (sm_100a): shared argument block and epilogue helpers of csrc/gemm_h.cu (tcgen05 products)
// and csrc/h16_ops.cu (memory-bound corner shapes, format conversion).  See include/ndjir_b200.h, "split-fp16 MLP
// engine", for the storage format.
#pragma once
#include <cuda_fp16.h>
#include "gemm.cuh"
#include "../../include/ndjir_b200.h"

namespace ndjir {
namespace gemmh {

using gemm::EPI_BIAS;
using gemm::EPI_SOFTPLUS;
using gemm::EPI_ACCUM;
using gemm::EPI_MUL_S;
using gemm::EPI_ADJ;
using gemm::EPI_ATOMIC;

constexpr float H16_MAX = 65504.f;

// one matrix operand of an epilogue in either form (fp32 rows or two fp16 planes)
struct Op {
  float* f;            // fp32 form (nullptr when the split form is used)
  long long ldf;
  __half* hi;          // split form
  __half* lo;
  long long ldh;
  const float* scale;  // device scalar
  float* amax;         // device scalar (outputs only)
};

struct HArgs {
  int M, N, K;
  int mn, epi, precise, split_k;
  float alpha, out_scale, beta, hscale;
  const float* a_scale;
  const float* b_scale;
  // split operands: K-major mode A (M x K), B (N x K); MN mode A (K x M), B (K x N); row strides in halfs
  const __half* Ahi; const __half* Alo; long long lda;
  const __half* Bhi; const __half* Blo; long long ldb;
  // fp32 operands of the memory-bound corner shapes (h16_ops.cu)
  const float* A32; long long a_rs, a_cs;
  const float* B32; long long b_rs, b_cs;
  Op C, C2, H, U;
  const float* bias;
  float* colsum;       // MN-major mode: [N] += column sums of B (bias gradient), or nullptr
};

__device__ __forceinline__ float dev_scalar(const float* p) { return p ? __ldg(p) : 1.f; }

__device__ __forceinline__ float clamp_h16(float v) { return fminf(fmaxf(v, -H16_MAX), H16_MAX); }

// x (already multiplied by the tensor's scale) -> hi, lo
__device__ __forceinline__ void split1(float xs, __half& hi, __half& lo) {
  xs = clamp_h16(xs);
  hi = __float2half_rn(xs);
  lo = __float2half_rn(xs - __half2float(hi));
}
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  x0 = clamp_h16(x0); x1 = clamp_h16(x1);
  __half2 h = __floats2half2_rn(x0, x1);
  float2 b = __half22float2(h);
  __half2 l = __floats2half2_rn(x0 - b.x, x1 - b.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ float2 join2(uint32_t hi, uint32_t lo) {
  float2 a = __half22float2(*reinterpret_cast<__half2*>(&hi));
  float2 b = __half22float2(*reinterpret_cast<__half2*>(&lo));
  return make_float2(a.x + b.x, a.y + b.y);
}

__device__ __forceinline__ float op_load(const Op& o, float inv, long long m, int n) {
  if (o.hi) return (__half2float(o.hi[m * o.ldh + n]) + __half2float(o.lo[m * o.ldh + n])) * inv;
  return o.f[m * o.ldf + n];
}
__device__ __forceinline__ void op_store(const Op& o, float sc, long long m, int n, float v) {
  if (o.hi) {
    __half h, l;
    split1(v * sc, h, l);
    o.hi[m * o.ldh + n] = h;
    o.lo[m * o.ldh + n] = l;
  } else {
    o.f[m * o.ldf + n] = v;
  }
}

// fused epilogue arithmetic on one element (acc already carries 1 / (scale_A * scale_B))
template <int EPI>
__device__ __forceinline__ void epi_math(const HArgs& a, float acc, float h, float u, float cprev, float b, float& o,
                                         float& o2) {
  o2 = 0.f;
  if (EPI == EPI_BIAS) o = a.alpha * acc + b;
  else if (EPI == EPI_SOFTPLUS) o = a.out_scale * gemm::softplus_beta_fast(acc + b, a.beta);
  else if (EPI == EPI_ACCUM) o = cprev + a.alpha * acc;
  else if (EPI == EPI_MUL_S) {
    float sg = gemm::sig_from_softplus_fast(h * a.hscale, a.beta);
    o = a.alpha * acc * sg + u;
  } else if (EPI == EPI_ADJ) {
    float sg = gemm::sig_from_softplus_fast(h * a.hscale, a.beta);
    o = acc * u * a.beta * (1.f - sg);
    o2 = a.out_scale * acc * sg;
  } else o = a.alpha * acc;
}

// fold a thread's running max into the tensor's device scalar (non-negative floats order like their bit patterns)
__device__ __forceinline__ void amax_commit(float* amax, float mx) {
  unsigned bits = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
  if (amax && (threadIdx.x & 31) == 0 && bits != 0u) atomicMax(reinterpret_cast<unsigned*>(amax), bits);
}

extern int g_h_dbg;                                   // gemm_h.cu: profiling switches
extern int g_h_tall;                                  // gemm_h.cu: 1 = 256-row work items for the weight-gradient products
extern int g_h_tma_epi;                               // gemm_h.cu: 1 = TMA-staged epilogue where the operands allow it
int launch_tc(const HArgs& a, cudaStream_t st);       // gemm_h.cu
extern int g_h_pair;                                  // gemm_h2.cu: 1 = CTA-pair kernel for the activation-row products
bool pair_eligible(const HArgs& a);                   // gemm_h2.cu
int launch_pair(const HArgs& a, cudaStream_t st);     // gemm_h2.cu
extern int g_h_resident;                              // gemm_h3.cu: 1 = resident-weight kernel for K-major products, K <= 256
bool resident_eligible(const HArgs& a);               // gemm_h3.cu
int launch_resident(const HArgs& a, cudaStream_t st); // gemm_h3.cu
extern int g_h_chain;                                 // gemm_h_chain.cu: 1 = the SDF network as one kernel, activations on chip
int launch_geo_chain(const ndjir_geo_net* net, long long rows, int din, const ndjir_hmat* ench, const float* enc32,
                     long long ld_enc32, const ndjir_hmat* act, float* sdf, cudaStream_t st);   // gemm_h_chain.cu
bool corner_shape(const HArgs& a);                    // h16_ops.cu
int launch_corner(const HArgs& a, cudaStream_t st);   // h16_ops.cu

}  // namespace gemmh
}  // namespace ndjir
