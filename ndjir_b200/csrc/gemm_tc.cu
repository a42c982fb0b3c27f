// MLP products on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::tf32 with TMEM accumulators, operands
// staged by TMA, 3xTF32 error compensation so the result keeps fp32 accuracy (parity bar 1e-5 forward):
//
//     x = hi + lo,  hi = x with the low 13 mantissa bits cleared (what a TF32 operand keeps),  lo = x - hi  (exact)
//     A*B ~= A_hi*B_hi + A_hi*B_lo + A_lo*B_hi            (the dropped lo*lo term is 2^-22 relative)
//
// One CTA computes one 128 x (<=256) output tile (optionally one K split of it):
//   warp 0      TMA producer: raw fp32 tiles of A and B, swizzled, 4-stage ring of 16-wide K blocks (UTMALDG)
//   warps 2-5   transform: lo = x - hi written to a second tile with the SAME swizzled addresses (pure element-wise
//               pass over shared memory), fence.proxy.async, then signal the MMA warp
//   warp 1      one elected thread issues 3 x 4 tcgen05.mma (128 x N x 8) per 32-wide K block        (UTCHMMA..)
//   warps 2-5   epilogue: tcgen05.ld -> padded shared staging -> coalesced fused epilogue (bias / softplus /
//               sigmoid factor / second-order term / atomics), shared with the FFMA path (gemm.cuh)
// Operand layouts come straight from the MLP passes, no transposed copies:
//   forward  C = A W      : A K-major,  B = W (in,out) MN-major
//   dgrad    dA = dZ W^T  : A K-major,  B = W          K-major
//   wgrad    gW = A^T dZ  : A MN-major, B MN-major, contraction over the sample rows, split over gridDim.z
// Shapes that do not fit (N < 32, K < 16, unaligned views) stay on the FFMA kernel (gemm.cu).
#include <cuda.h>
#include <cudaTypedefs.h>
#include "gemm.cuh"
#include "../../include/ndjir_b200.h"

namespace ndjir {
namespace gemm {

int g_mlp_mask_hi = 0;
int g_mlp_cta_pair = 0;   // 1: activation-row products run on CTA pairs (tcgen05 cta_group::2)
int g_mlp_dbg = 0;   // experiment switches (profiling only): 1 = skip the lo transform, 2 = issue only the hi*hi product

constexpr int TC_BM = 128;
constexpr int TC_BN = 256;
constexpr int TC_BK = 16;                       // fp32 elements per K block: 64-byte rows (K-major tiles use SWIZZLE_64B);
                                                // 32-wide blocks with 2 stages measured the same
constexpr int TC_STAGES = 4;                    // (tried: 6 raw + 2 lo stages in the same 192 KB: no gain, the ring is
                                                //  not latency bound; L2 prefetch of the next item: no gain either)
constexpr int KM_ROW_BYTES = TC_BK * 4;         // K-major tile row
constexpr uint32_t KM_LAYOUT = TC_BK == 32 ? 2u : 4u;   // SWIZZLE_128B : SWIZZLE_64B
constexpr uint32_t KM_SBO = 8 * KM_ROW_BYTES;   // 8-row swizzle groups
constexpr int A_TILE_BYTES = TC_BM * TC_BK * 4;  // 8 KB
constexpr int B_TILE_BYTES = TC_BN * TC_BK * 4;  // 16 KB
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;   // raw + lo of both operands = 48 KB
constexpr int TC_SMEM_BYTES = TC_STAGES * STAGE_BYTES + 1024;      // + alignment slack
constexpr int TC_THREADS = 576;                 // 1 TMA + 1 MMA + 8 transform + 8 epilogue warps (<= 112 registers)
constexpr int TMEM_COLS = 512;                  // two 256-column fp32 accumulators

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded spin: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; spin < (1u << 28); ++spin)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// L2 prefetch of one TMA box / of a contiguous byte range: no shared memory is reserved, so HBM requests for the NEXT
// work item are in flight while the current one occupies the whole operand ring
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// explicit shared-space accesses for the transform pass: a pointer rebuilt from an aligned integer loses its address
// space and nvcc emits generic LD.E/ST.E (seen in the ncu source page: the transform stalled on them)
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (tcgen05), version 1 (cute/arch/mma_sm100_desc.hpp:SmemDescriptor).
// layout 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B: the only layout tcgen05 accepts for MN-major
// 32-bit operands (32-byte chunks swizzled inside a 128-byte span over 4 rows; TMA mode SWIZZLE_128B_ATOM_32B).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;
  return d;
}

struct TcParams {
  Args a;
  int a_mn, b_mn;        // 1: operand is MN-major (contiguous along m / n), 0: K-major
  int mask_hi;
  int dbg;
  int a_3d, b_3d;        // MN-major operand loaded with ONE 3-D TMA request per K block (extent % 32 == 0)
  int do_colsum;         // the transform warps also accumulate the column sums of the MN-major B tiles (a.colsum)
  int b_lo_tma;          // the lo part of B comes pre-split from global memory (registered weights): no B transform
  int vec_epi;           // every epilogue operand is 16-byte aligned with a row stride that is a multiple of 4
  int m_tiles, n_tiles, splits, kb_per_split, nkb_total;
};

// fused epilogue on 4 consecutive columns (vector path of the tensor-core kernel)
template <int EPI>
__device__ __forceinline__ void epilogue_vec4(const Args& a, long long m, int n, const float* acc, float4 hv, float4 uv,
                                              float4 cv, float4 bv) {
  float h[4] = {hv.x, hv.y, hv.z, hv.w}, u[4] = {uv.x, uv.y, uv.z, uv.w}, co[4] = {cv.x, cv.y, cv.z, cv.w};
  float bb[4] = {bv.x, bv.y, bv.z, bv.w};
  float o[4], o2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (EPI == EPI_BIAS) o[e] = a.alpha * acc[e] + bb[e];
    else if (EPI == EPI_SOFTPLUS) o[e] = a.out_scale * softplus_beta_fast(acc[e] + bb[e], a.beta);
    else if (EPI == EPI_ACCUM) o[e] = co[e] + a.alpha * acc[e];
    else if (EPI == EPI_MUL_S) {
      float sg = sig_from_softplus_fast(h[e] * a.hscale, a.beta);
      o[e] = a.alpha * acc[e] * sg + u[e];
    } else if (EPI == EPI_ADJ) {
      float sg = sig_from_softplus_fast(h[e] * a.hscale, a.beta);
      o[e] = acc[e] * u[e] * a.beta * (1.f - sg);
      o2[e] = a.out_scale * acc[e] * sg;
    } else o[e] = a.alpha * acc[e];
  }
  float* cp = a.C + m * a.ldc + n;
  if (EPI == EPI_ATOMIC) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(cp), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3])
                 : "memory");
  } else {
    *reinterpret_cast<float4*>(cp) = make_float4(o[0], o[1], o[2], o[3]);
    if (EPI == EPI_ADJ) *reinterpret_cast<float4*>(a.C2 + m * a.ldc2 + n) = make_float4(o2[0], o2[1], o2[2], o2[3]);
  }
}

// Persistent kernel: one CTA per SM walks a static list of work items (m tile, n tile, K split).
//   warp 0        TMA producer                                  warps 2-9    hi/lo transform (256 threads)
//   warp 1        MMA issuer (+ TMEM allocation, 512 columns)   warps 10-17  epilogue straight from TMEM (2 warps per
//                                                                            sub-partition, alternate 16-column chunks)
// Two 256-column TMEM accumulators: the epilogue of item i overlaps the main loop of item i+1.
constexpr int TC_XFORM_THREADS = 256;
constexpr int TC_EPI_WARP0 = 2 + TC_XFORM_THREADS / 32;
constexpr int TC_EPI_THREADS = 256;            // two warps per TMEM sub-partition, each takes half of the columns

template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapBlo, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * TC_STAGES + 4];
  __shared__ uint32_t tmem_base_sh;
  __shared__ float colsum_sh[TC_BN];

  const Args& a = p.a;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  auto bar_full = [&](int s) { return smem_u32(&bars[s]); };
  auto bar_ready = [&](int s) { return smem_u32(&bars[TC_STAGES + s]); };
  auto bar_empty = [&](int s) { return smem_u32(&bars[2 * TC_STAGES + s]); };
  auto bar_acc_full = [&](int b) { return smem_u32(&bars[3 * TC_STAGES + b]); };
  auto bar_acc_empty = [&](int b) { return smem_u32(&bars[3 * TC_STAGES + 2 + b]); };
  auto a_raw = [&](int s) { return smem_base + s * STAGE_BYTES; };
  auto a_lo = [&](int s) { return smem_base + s * STAGE_BYTES + A_TILE_BYTES; };
  auto b_raw = [&](int s) { return smem_base + s * STAGE_BYTES + 2 * A_TILE_BYTES; };
  auto b_lo = [&](int s) { return smem_base + s * STAGE_BYTES + 2 * A_TILE_BYTES + B_TILE_BYTES; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_ready(s), TC_XFORM_THREADS);
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), TC_EPI_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_sh;

  const int n_items = p.m_tiles * p.n_tiles * p.splits;
  // work item -> (m0, n0, kb0, nkb); m tiles vary fastest so co-running CTAs share the same weight tile
  auto item_info = [&](int item, int& m0, int& n0, int& kb0, int& nkb) {
    int mt = item % p.m_tiles;
    int rest = item / p.m_tiles;
    int ntile = rest % p.n_tiles;
    int sp = rest / p.n_tiles;
    m0 = mt * TC_BM;
    n0 = ntile * TC_BN;
    kb0 = sp * p.kb_per_split;
    int kb1 = min(p.nkb_total, kb0 + p.kb_per_split);
    nkb = max(0, kb1 - kb0);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int m0, n0, kb0, nkb;
        item_info(item, m0, n0, kb0, nkb);
        const int umma_n = (min(TC_BN, a.N - n0) + 15) & ~15;
        const int b_chunks = (umma_n + 31) / 32;
        const uint32_t b_tx = (p.b_mn && !p.b_3d) ? b_chunks * TC_BK * 128 : B_TILE_BYTES;
        const uint32_t tx_bytes = A_TILE_BYTES + (p.b_lo_tma ? 2 * b_tx : b_tx);
        for (int i = 0; i < nkb; ++i, ++it) {
          int s = it % TC_STAGES;
          uint32_t ph = (it / TC_STAGES) & 1;
          mbar_wait(bar_empty(s), ph ^ 1);
          mbar_expect_tx(bar_full(s), tx_bytes);
          int k0 = (kb0 + i) * TC_BK;
          if (p.a_3d) {
            tma_load_3d(a_raw(s), &mapA, 0, k0, m0 / 32, bar_full(s));
          } else if (p.a_mn) {
#pragma unroll
            for (int c = 0; c < TC_BM / 32; ++c) tma_load_2d(a_raw(s) + c * TC_BK * 128, &mapA, m0 + c * 32, k0, bar_full(s));
          } else {
            tma_load_2d(a_raw(s), &mapA, k0, m0, bar_full(s));
          }
          if (p.b_3d) {
            tma_load_3d(b_raw(s), &mapB, 0, k0, n0 / 32, bar_full(s));
          } else if (p.b_mn) {
            for (int c = 0; c < b_chunks; ++c) tma_load_2d(b_raw(s) + c * TC_BK * 128, &mapB, n0 + c * 32, k0, bar_full(s));
          } else {
            tma_load_2d(b_raw(s), &mapB, k0, n0, bar_full(s));
          }
          if (p.b_lo_tma) {   // same box, same swizzle, from the pre-split copy of the weights
            if (p.b_3d) {
              tma_load_3d(b_lo(s), &mapBlo, 0, k0, n0 / 32, bar_full(s));
            } else if (p.b_mn) {
              for (int c = 0; c < b_chunks; ++c) tma_load_2d(b_lo(s) + c * TC_BK * 128, &mapBlo, n0 + c * 32, k0, bar_full(s));
            } else {
              tma_load_2d(b_lo(s), &mapBlo, k0, n0, bar_full(s));
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // K-major (SWIZZLE_64B for 16-wide K blocks): rows of TC_BK*4 B, 8-row groups KM_SBO apart, K step of 8
      // elements = +32 B inside the row.
      // MN-major (SWIZZLE_128B_BASE32B): k rows of 128 B (32 mn elements), 4-row k groups 512 B apart (SBO), 32-wide
      // mn chunks TC_BK*128 B apart (LBO), K step of 8 rows = +1024 B.
      const uint32_t a_lbo = p.a_mn ? TC_BK * 128 : 16, a_sbo = p.a_mn ? 512 : KM_SBO, a_step = p.a_mn ? 1024 : 32;
      const uint32_t b_lbo = p.b_mn ? TC_BK * 128 : 16, b_sbo = p.b_mn ? 512 : KM_SBO, b_step = p.b_mn ? 1024 : 32;
      const uint32_t a_lay = p.a_mn ? 1 : KM_LAYOUT, b_lay = p.b_mn ? 1 : KM_LAYOUT;
      const uint64_t da_hi0 = make_desc(a_raw(0), a_lbo, a_sbo, a_lay), da_lo0 = make_desc(a_lo(0), a_lbo, a_sbo, a_lay);
      const uint64_t db_hi0 = make_desc(b_raw(0), b_lbo, b_sbo, b_lay), db_lo0 = make_desc(b_lo(0), b_lbo, b_sbo, b_lay);
      uint32_t it = 0, tile_it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tile_it) {
        int m0, n0, kb0, nkb;
        item_info(item, m0, n0, kb0, nkb);
        if (nkb == 0) continue;
        const int umma_n = (min(TC_BN, a.N - n0) + 15) & ~15;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                               ((uint32_t)(umma_n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        const int buf = tile_it & 1;
        mbar_wait(bar_acc_empty(buf), ((tile_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * TC_BN;
        for (int i = 0; i < nkb; ++i, ++it) {
          int s = it % TC_STAGES;
          uint32_t ph = (it / TC_STAGES) & 1;
          mbar_wait(bar_ready(s), ph);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < TC_BK / 8; ++ks) {
            // descriptors differ from the stage-0 ones only in the 14-bit start-address field (units of 16 bytes):
            // one 64-bit add per operand instead of rebuilding them (the single issuing thread is on the critical path)
            const uint64_t oa = (uint64_t)((s * STAGE_BYTES + ks * a_step) >> 4);
            const uint64_t ob = (uint64_t)((s * STAGE_BYTES + ks * b_step) >> 4);
            uint64_t da_hi = da_hi0 + oa, da_lo = da_lo0 + oa, db_hi = db_hi0 + ob, db_lo = db_lo0 + ob;
            if (p.dbg & 2) {
              umma_tf32(tacc, da_hi, db_hi, idesc, (i > 0 || ks > 0) ? 1u : 0u);
            } else {
              umma_tf32(tacc, da_lo, db_hi, idesc, (i > 0 || ks > 0) ? 1u : 0u);   // small terms first
              umma_tf32(tacc, da_hi, db_lo, idesc, 1u);
              umma_tf32(tacc, da_hi, db_hi, idesc, 1u);
            }
          }
          umma_commit(bar_empty(s));     // frees the stage once these MMAs have read it
        }
        umma_commit(bar_acc_full(buf));  // accumulator complete
      }
    }
  } else if (warp < TC_EPI_WARP0) {
    // ===================== transform: lo = x - hi for both operand tiles =====================
    const int t = threadIdx.x - 64;   // 0..TC_XFORM_THREADS-1
    uint32_t it = 0;
    // Column sums of B (the bias gradient that goes with a weight-gradient product): every thread meets the same
    // four 16-byte pieces of the [chunk][k][32 n] tile in every K block, so it keeps four private float4 sums and
    // folds them into the CTA's column vector once per work item.  Physical -> logical offset of a piece: the
    // 128-byte rows are swizzled in 32-byte units, unit ^= (row & 3) (Swizzle<2,5,2>, the layout TMA wrote).
    float4 cs[B_TILE_BYTES / (TC_XFORM_THREADS * 16)];
#pragma unroll
    for (int j = 0; j < B_TILE_BYTES / (TC_XFORM_THREADS * 16); ++j) cs[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.do_colsum) {
      for (int c = t; c < TC_BN; c += TC_XFORM_THREADS) colsum_sh[c] = 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(TC_XFORM_THREADS) : "memory");
    }
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int m0, n0, kb0, nkb;
      item_info(item, m0, n0, kb0, nkb);
      const int umma_n = (min(TC_BN, a.N - n0) + 15) & ~15;
      const int b_bytes = (p.b_mn && !p.b_3d) ? ((umma_n + 31) / 32) * TC_BK * 128 : B_TILE_BYTES;
      const bool sum_b = p.do_colsum && m0 == 0;     // one m tile per (n tile, K split) carries the sums
      for (int i = 0; i < nkb; ++i, ++it) {
        int s = it % TC_STAGES;
        uint32_t ph = (it / TC_STAGES) & 1;
        mbar_wait(bar_full(s), ph);
        const uint32_t st = smem_base + s * STAGE_BYTES;
        if (p.dbg & 1) { fence_proxy_async(); mbar_arrive(bar_ready(s)); continue; }
        // A: raw at 0, lo at A_TILE_BYTES;  B: raw at 2*A_TILE_BYTES, lo at 2*A_TILE_BYTES + B_TILE_BYTES
#pragma unroll
        for (int off = t * 16; off < A_TILE_BYTES; off += TC_XFORM_THREADS * 16) {
          float4 x = lds128(st + off);
          float4 h;
          h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
          h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
          h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
          h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
          sts128(st + A_TILE_BYTES + off, make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w));
          if (p.mask_hi) sts128(st + off, h);
        }
        if (!p.b_lo_tma) {
#pragma unroll
          for (int j = 0; j < B_TILE_BYTES / (TC_XFORM_THREADS * 16); ++j) {
            const int off = t * 16 + j * TC_XFORM_THREADS * 16;
            if (off < b_bytes) {
              float4 x = lds128(st + 2 * A_TILE_BYTES + off);
              float4 h;
              h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
              h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
              h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
              h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
              sts128(st + 2 * A_TILE_BYTES + B_TILE_BYTES + off, make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w));
              if (p.mask_hi) sts128(st + 2 * A_TILE_BYTES + off, h);
              if (sum_b) { cs[j].x += x.x; cs[j].y += x.y; cs[j].z += x.z; cs[j].w += x.w; }
            }
          }
        }
        fence_proxy_async();          // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(bar_ready(s));
      }
      if (sum_b) {
#pragma unroll
        for (int j = 0; j < B_TILE_BYTES / (TC_XFORM_THREADS * 16); ++j) {
          const int off = t * 16 + j * TC_XFORM_THREADS * 16;
          if (off < b_bytes) {
            const int lof = off ^ (((off >> 7) & 3) << 5);                 // undo the 32-byte-unit swizzle
            const int col = (lof >> 11) * 32 + ((lof & 127) >> 2);          // chunk * 32 + n inside the 128-byte row
            atomicAdd(&colsum_sh[col], cs[j].x); atomicAdd(&colsum_sh[col + 1], cs[j].y);
            atomicAdd(&colsum_sh[col + 2], cs[j].z); atomicAdd(&colsum_sh[col + 3], cs[j].w);
          }
          cs[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(TC_XFORM_THREADS) : "memory");
        for (int c = t; c < TC_BN; c += TC_XFORM_THREADS) {
          float v = colsum_sh[c];
          colsum_sh[c] = 0.f;
          if (v != 0.f && n0 + c < a.N) atomicAdd(a.colsum + n0 + c, v);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(TC_XFORM_THREADS) : "memory");
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global, one output row per thread =====================
    const int q = warp & 3;           // TMEM sub-partition of this warp: lanes 32q .. 32q+31
    const int chalf = (warp - TC_EPI_WARP0) >> 2;   // 0: even 16-column chunks, 1: odd chunks
    uint32_t tile_it = 0;
    constexpr bool NEED_H = (EPI == EPI_MUL_S || EPI == EPI_ADJ);
    constexpr bool NEED_C = (EPI == EPI_ACCUM);
    const bool need_u = (EPI == EPI_ADJ) || (EPI == EPI_MUL_S && a.U != nullptr);
    const bool need_b = (EPI == EPI_BIAS || EPI == EPI_SOFTPLUS) && a.bias != nullptr;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tile_it) {
      int m0, n0, kb0, nkb;
      item_info(item, m0, n0, kb0, nkb);
      if (nkb == 0) continue;
      const int n_valid = min(TC_BN, a.N - n0);
      const int buf = tile_it & 1;
      const long long m = m0 + q * 32 + lane;
      const bool row_ok = m < a.M;
      const int n_vec = p.vec_epi ? (n_valid & ~3) : 0;
      mbar_wait(bar_acc_full(buf), (tile_it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * TC_BN + ((uint32_t)(q * 32) << 16);
      for (int c0 = chalf * 16; c0 < n_valid; c0 += 32) {
        // operands of the fused epilogue for this thread's 16 columns: issued before the accumulator is read so
        // up to 8 independent 16-byte loads per thread (x 256 epilogue threads) are in flight
        float4 hv[4], uv[4], cv[4], bv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int c = c0 + 4 * j;
          hv[j] = uv[j] = cv[j] = bv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && c < n_vec) {
            if (NEED_H) hv[j] = __ldg(reinterpret_cast<const float4*>(a.H + m * a.ldh + n0 + c));
            if (need_u) uv[j] = *reinterpret_cast<const float4*>(a.U + m * a.ldu + n0 + c);
            if (NEED_C) cv[j] = *reinterpret_cast<const float4*>(a.C + m * a.ldc + n0 + c);
            if (need_b) bv[j] = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + c));
          }
        }
        float v[16];
        tmem_ld16(tacc + (uint32_t)c0, v);
        if (p.dbg & 4) { if (v[0] == 1.2345e-30f && row_ok) a.C[m * a.ldc + n0 + c0] = v[1]; continue; }
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            int c = c0 + 4 * j;
            if (c < n_vec) {
              epilogue_vec4<EPI>(a, m, n0 + c, v + 4 * j, hv[j], uv[j], cv[j], bv[j]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (c + e < n_valid) epilogue_store<EPI>(a, (int)m, n0 + c + e, v[4 * j + e]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_acc_empty(buf));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-PAIR variant (tcgen05 cta_group::2) for the activation-row products (A K-major, M large).
// Two CTAs of a cluster - the two SMs of a TPC - compute one 256 x (<=256) tile: each stages ITS 128 rows of A and
// ITS HALF of the B columns, the leader's elected thread issues 256 x N x 8 MMAs that read both SMs' shared memory,
// and every CTA finds the accumulator rows of its own 128 rows in its own TMEM.  Per SM and K block this halves the B
// traffic through shared memory (TMA fill, hi/lo transform, MMA operand reads: 144 KB -> 96 KB), which is what bounds
// the single-CTA kernel (l1tex data pipe at ~90 %).
// Synchronisation: the peer's transform threads arrive on the LEADER's `ready` barrier (mapa + remote arrive), the
// leader's tcgen05.commit multicasts to the `empty` / `acc_full` barriers of both CTAs, and both CTAs' epilogue
// threads release the accumulator on the leader's `acc_empty` barrier.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // acquire at cluster scope
  for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

constexpr int P2_BNH = 128;                                  // B columns staged per CTA
constexpr int P2_B_TILE_BYTES = P2_BNH * TC_BK * 4;          // 8 KB
constexpr int P2_STAGE_BYTES = 2 * A_TILE_BYTES + 2 * P2_B_TILE_BYTES;   // 32 KB
constexpr int P2_STAGES = 6;
constexpr int P2_SMEM_BYTES = P2_STAGES * P2_STAGE_BYTES + 1024;

template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * P2_STAGES + 4];
  __shared__ uint32_t tmem_base_sh;

  const Args& a = p.a;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  auto bar_full = [&](int s) { return smem_u32(&bars[s]); };
  auto bar_ready = [&](int s) { return smem_u32(&bars[P2_STAGES + s]); };
  auto bar_empty = [&](int s) { return smem_u32(&bars[2 * P2_STAGES + s]); };
  auto bar_acc_full = [&](int b) { return smem_u32(&bars[3 * P2_STAGES + b]); };
  auto bar_acc_empty = [&](int b) { return smem_u32(&bars[3 * P2_STAGES + 2 + b]); };
  auto a_raw = [&](int s) { return smem_base + s * P2_STAGE_BYTES; };
  auto a_lo = [&](int s) { return smem_base + s * P2_STAGE_BYTES + A_TILE_BYTES; };
  auto b_raw = [&](int s) { return smem_base + s * P2_STAGE_BYTES + 2 * A_TILE_BYTES; };
  auto b_lo = [&](int s) { return smem_base + s * P2_STAGE_BYTES + 2 * A_TILE_BYTES + P2_B_TILE_BYTES; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < P2_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_ready(s), 2 * TC_XFORM_THREADS);     // only the leader's copy is used
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), 2 * TC_EPI_THREADS);   // only the leader's copy is used
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_sh;

  // work items: (pair of m tiles, n tile); clusters walk them round-robin
  const int pair_tiles = (p.m_tiles + 1) / 2;
  const int n_items = pair_tiles * p.n_tiles;
  const int n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  auto item_info = [&](int item, int& m0, int& n0, int& umma_n) {
    int pt = item % pair_tiles;
    int ntile = item / pair_tiles;
    m0 = pt * 2 * TC_BM;
    n0 = ntile * TC_BN;
    umma_n = (min(TC_BN, a.N - n0) + 63) & ~63;   // both halves are multiples of 32 columns
  };
  const int nkb = p.nkb_total;

  if (warp == 0) {
    // ===================== TMA producer (each CTA: its A rows, its half of B) =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        int m0, n0, umma_n;
        item_info(item, m0, n0, umma_n);
        const int half = umma_n / 2, chunks = half / 32;
        const int mr = m0 + rank * TC_BM, nr = n0 + rank * half;
        const uint32_t tx_bytes = A_TILE_BYTES + ((p.b_mn && !p.b_3d) ? chunks * TC_BK * 128 : P2_B_TILE_BYTES);
        for (int i = 0; i < nkb; ++i, ++it) {
          int s = it % P2_STAGES;
          uint32_t ph = (it / P2_STAGES) & 1;
          mbar_wait(bar_empty(s), ph ^ 1);
          mbar_expect_tx(bar_full(s), tx_bytes);
          int k0 = i * TC_BK;
          tma_load_2d(a_raw(s), &mapA, k0, mr, bar_full(s));
          if (p.b_3d) {
            tma_load_3d(b_raw(s), &mapB, 0, k0, nr / 32, bar_full(s));
          } else if (p.b_mn) {
            for (int c = 0; c < chunks; ++c) tma_load_2d(b_raw(s) + c * TC_BK * 128, &mapB, nr + c * 32, k0, bar_full(s));
          } else {
            tma_load_2d(b_raw(s), &mapB, k0, nr, bar_full(s));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the leader CTA only =====================
    if (leader && lane == 0) {
      const uint32_t a_lbo = 16, a_sbo = KM_SBO, a_step = 32, a_lay = KM_LAYOUT;
      const uint32_t b_lbo = p.b_mn ? TC_BK * 128 : 16, b_sbo = p.b_mn ? 512 : KM_SBO, b_step = p.b_mn ? 1024 : 32;
      const uint32_t b_lay = p.b_mn ? 1 : KM_LAYOUT;
      const uint64_t da_hi0 = make_desc(a_raw(0), a_lbo, a_sbo, a_lay), da_lo0 = make_desc(a_lo(0), a_lbo, a_sbo, a_lay);
      const uint64_t db_hi0 = make_desc(b_raw(0), b_lbo, b_sbo, b_lay), db_lo0 = make_desc(b_lo(0), b_lbo, b_sbo, b_lay);
      uint32_t it = 0, tile_it = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters, ++tile_it) {
        int m0, n0, umma_n;
        item_info(item, m0, n0, umma_n);
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.b_mn << 16) |
                               ((uint32_t)(umma_n >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);
        const int buf = tile_it & 1;
        mbar_wait_cluster(bar_acc_empty(buf), ((tile_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * TC_BN;
        for (int i = 0; i < nkb; ++i, ++it) {
          int s = it % P2_STAGES;
          uint32_t ph = (it / P2_STAGES) & 1;
          mbar_wait_cluster(bar_ready(s), ph);       // 512 arrivals: the transform threads of both CTAs
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < TC_BK / 8; ++ks) {
            const uint64_t oa = (uint64_t)((s * P2_STAGE_BYTES + ks * a_step) >> 4);
            const uint64_t ob = (uint64_t)((s * P2_STAGE_BYTES + ks * b_step) >> 4);
            uint64_t da_hi = da_hi0 + oa, da_lo = da_lo0 + oa, db_hi = db_hi0 + ob, db_lo = db_lo0 + ob;
            umma_tf32_pair(tacc, da_lo, db_hi, idesc, (i > 0 || ks > 0) ? 1u : 0u);
            umma_tf32_pair(tacc, da_hi, db_lo, idesc, 1u);
            umma_tf32_pair(tacc, da_hi, db_hi, idesc, 1u);
          }
          umma_commit_pair(bar_empty(s));       // frees the stage in both CTAs
        }
        umma_commit_pair(bar_acc_full(buf));    // accumulator complete, both CTAs
      }
    }
  } else if (warp < TC_EPI_WARP0) {
    // ===================== transform: lo = x - hi for this CTA's tiles; arrive on the LEADER's barrier ==========
    const int t = threadIdx.x - 64;
    uint32_t it = 0;
    for (int item = cluster_id; item < n_items; item += n_clusters) {
      int m0, n0, umma_n;
      item_info(item, m0, n0, umma_n);
      const int half = umma_n / 2;
      const int b_bytes = (p.b_mn && !p.b_3d) ? (half / 32) * TC_BK * 128 : P2_B_TILE_BYTES;
      for (int i = 0; i < nkb; ++i, ++it) {
        int s = it % P2_STAGES;
        uint32_t ph = (it / P2_STAGES) & 1;
        mbar_wait(bar_full(s), ph);
        const uint32_t st = smem_base + s * P2_STAGE_BYTES;
#pragma unroll
        for (int off = t * 16; off < A_TILE_BYTES; off += TC_XFORM_THREADS * 16) {
          float4 x = lds128(st + off);
          float4 h;
          h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
          h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
          h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
          h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
          sts128(st + A_TILE_BYTES + off, make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w));
        }
#pragma unroll
        for (int off = t * 16; off < b_bytes; off += TC_XFORM_THREADS * 16) {
          float4 x = lds128(st + 2 * A_TILE_BYTES + off);
          float4 h;
          h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
          h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
          h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
          h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
          sts128(st + 2 * A_TILE_BYTES + P2_B_TILE_BYTES + off, make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w));
        }
        fence_proxy_async();
        if (leader) mbar_arrive(bar_ready(s));
        else mbar_arrive_cluster(mapa_cluster(bar_ready(s), 0));
      }
    }
  } else {
    // ===================== epilogue: this CTA's 128 accumulator rows, all columns =====================
    const int q = warp & 3;
    const int chalf = (warp - TC_EPI_WARP0) >> 2;
    uint32_t tile_it = 0;
    constexpr bool NEED_H = (EPI == EPI_MUL_S || EPI == EPI_ADJ);
    constexpr bool NEED_C = (EPI == EPI_ACCUM);
    const bool need_u = (EPI == EPI_ADJ) || (EPI == EPI_MUL_S && a.U != nullptr);
    const bool need_b = (EPI == EPI_BIAS || EPI == EPI_SOFTPLUS) && a.bias != nullptr;
    for (int item = cluster_id; item < n_items; item += n_clusters, ++tile_it) {
      int m0, n0, umma_n;
      item_info(item, m0, n0, umma_n);
      const int n_valid = min(TC_BN, a.N - n0);
      const int buf = tile_it & 1;
      const long long m = m0 + rank * TC_BM + q * 32 + lane;
      const bool row_ok = m < a.M;
      const int n_vec = p.vec_epi ? (n_valid & ~3) : 0;
      mbar_wait(bar_acc_full(buf), (tile_it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * TC_BN + ((uint32_t)(q * 32) << 16);
      for (int c0 = chalf * 16; c0 < n_valid; c0 += 32) {
        float4 hv[4], uv[4], cv[4], bv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int c = c0 + 4 * j;
          hv[j] = uv[j] = cv[j] = bv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && c < n_vec) {
            if (NEED_H) hv[j] = __ldg(reinterpret_cast<const float4*>(a.H + m * a.ldh + n0 + c));
            if (need_u) uv[j] = *reinterpret_cast<const float4*>(a.U + m * a.ldu + n0 + c);
            if (NEED_C) cv[j] = *reinterpret_cast<const float4*>(a.C + m * a.ldc + n0 + c);
            if (need_b) bv[j] = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + c));
          }
        }
        float v[16];
        tmem_ld16(tacc + (uint32_t)c0, v);
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            int c = c0 + 4 * j;
            if (c < n_vec) {
              epilogue_vec4<EPI>(a, m, n0 + c, v + 4 * j, hv[j], uv[j], cv[j], bv[j]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (c + e < n_valid) epilogue_store<EPI>(a, (int)m, n0 + c + e, v[4 * j + e]);
            }
          }
        }
      }
      tc_fence_before();
      if (leader) mbar_arrive(bar_acc_empty(buf));
      else mbar_arrive_cluster(mapa_cluster(bar_acc_empty(buf), 0));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // no CTA leaves (or frees TMEM) while its peer may still touch its barriers / smem
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled get_encode() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(f);
  }
  return fn;
}

static inline bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// inner = contiguous extent (elements), outer = number of rows, ld = row stride (elements)
static bool make_map(CUtensorMap* map, const float* base, long long inner, long long outer, long long ld, int box_inner,
                     int box_outer, bool mn_major) {
  PFN_cuTensorMapEncodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                            : (TC_BK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// MN-major operand as a 3-D tensor (32 contiguous elements, K rows, 32-wide chunks): one request fills the whole
// [chunk][k][32] tile.  Only used when the MN extent is a multiple of 32 (no partial chunk).
static bool make_map3(CUtensorMap* map, const float* base, long long mn, long long rows, long long ld, int box_chunks) {
  PFN_cuTensorMapEncodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)(mn / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
  cuuint32_t box[3] = {32, (cuuint32_t)TC_BK, (cuuint32_t)box_chunks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Pre-split weights: when the caller supplies the lo part of B (Args::B_lo = x - tf32(x), same strides - the engine
// keeps such copies of its weight buffers and refreshes them when the parameters change, ndjir_split_lo), the B
// operand is fetched as two TMA tiles (raw + lo) and the in-kernel hi/lo transform only handles A.
int g_mlp_presplit = 1;

bool tc_eligible(const Args& a, int epi) {
  if (a.N < 32 || a.K < 16 || a.M < 32) return false;
  if (!al16p(a.A) || !al16p(a.B)) return false;
  // operand row strides must be multiples of 16 bytes; one of the two element strides is 1 (checked by launch)
  long long lda = a.a_cs == 1 ? a.a_rs : a.a_cs;
  long long ldb = a.b_cs == 1 ? a.b_rs : a.b_cs;
  if (a.a_cs == 1 && a.a_rs == 1) return false;   // degenerate views stay on the generic kernel
  if (lda % 4 != 0 || ldb % 4 != 0 || lda <= 0 || ldb <= 0) return false;
  if (a.a_cs == 1 && a.a_rs == 0) return false;
  if (a.b_cs == 1 && a.b_rs == 0) return false;
  (void)epi;
  return get_encode() != nullptr;
}

template <int EPI>
static int launch_tc2_epi(const Args& a, cudaStream_t st) {
  TcParams p;
  p.a = a;
  p.a_mn = 0;
  p.b_mn = (a.b_cs == 1);
  p.mask_hi = 0; p.dbg = 0; p.a_3d = 0; p.b_lo_tma = 0; p.do_colsum = 0;
  auto ok16 = [](const void* q, long long ld) { return q == nullptr || (al16p(q) && ld % 4 == 0); };
  p.vec_epi = ok16(a.C, a.ldc) && ok16(a.H, a.ldh) && ok16(a.U, a.ldu) && ok16(a.C2, a.ldc2) && ok16(a.bias, 0);
  p.b_3d = p.b_mn && a.N % 64 == 0;
  CUtensorMap mapA, mapB;
  bool ok = make_map(&mapA, a.A, a.K, a.M, a.a_rs, TC_BK, TC_BM, false);
  if (p.b_3d) ok = ok && make_map3(&mapB, a.B, a.N, a.K, a.b_rs, P2_BNH / 32);
  else if (p.b_mn) ok = ok && make_map(&mapB, a.B, a.N, a.K, a.b_rs, 32, TC_BK, true);
  else ok = ok && make_map(&mapB, a.B, a.K, a.N, a.b_cs, TC_BK, P2_BNH, false);
  if (!ok) return NDJIR_ERR_ARG;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  p.m_tiles = (a.M + TC_BM - 1) / TC_BM;
  p.n_tiles = (a.N + TC_BN - 1) / TC_BN;
  p.nkb_total = (a.K + TC_BK - 1) / TC_BK;
  p.splits = 1; p.kb_per_split = p.nkb_total;
  int n_items = ((p.m_tiles + 1) / 2) * p.n_tiles;
  int clusters = n_items < NDJIR_NUM_SMS / 2 ? n_items : NDJIR_NUM_SMS / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = P2_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<EPI>, mapA, mapB, p);
  if (e != cudaSuccess) return (int)e;
  NDJIR_RETURN_LAST_ERROR();
}

static bool tc2_eligible(const Args& a, int epi) {
  // activation-row products only: A row-major (K-major), many rows, no split-K
  return g_mlp_cta_pair && a.a_cs == 1 && a.M >= 4096 && a.split_k <= 1 && epi != EPI_ATOMIC && a.N >= 64;
}

template <int EPI>
static int launch_tc_epi(const Args& a, cudaStream_t st) {
  if (tc2_eligible(a, EPI)) return launch_tc2_epi<EPI>(a, st);
  TcParams p;
  p.a = a;
  p.a_mn = (a.a_cs != 1);        // A(m,k) contiguous along m
  p.b_mn = (a.b_cs == 1);        // B(k,n) contiguous along n
  p.mask_hi = g_mlp_mask_hi;
  p.dbg = g_mlp_dbg;

  if (a.b_cs == 1 && a.b_rs == 1) p.b_mn = 1;
  auto ok16 = [](const void* q, long long ld) { return q == nullptr || (al16p(q) && ld % 4 == 0); };
  p.vec_epi = ok16(a.C, a.ldc) && ok16(a.H, a.ldh) && ok16(a.U, a.ldu) && ok16(a.C2, a.ldc2) && ok16(a.bias, 0);
  CUtensorMap mapA, mapB;
  bool ok;
  p.a_3d = p.a_mn && a.M % 32 == 0;
  p.b_3d = p.b_mn && a.N % 32 == 0;
  if (p.a_3d) ok = make_map3(&mapA, a.A, a.M, a.K, a.a_cs, TC_BM / 32);
  else if (p.a_mn) ok = make_map(&mapA, a.A, a.M, a.K, a.a_cs, 32, TC_BK, true);
  else ok = make_map(&mapA, a.A, a.K, a.M, a.a_rs, TC_BK, TC_BM, false);
  if (p.b_3d) ok = ok && make_map3(&mapB, a.B, a.N, a.K, a.b_rs, TC_BN / 32);
  else if (p.b_mn) ok = ok && make_map(&mapB, a.B, a.N, a.K, a.b_rs, 32, TC_BK, true);
  else ok = ok && make_map(&mapB, a.B, a.K, a.N, a.b_cs, TC_BK, TC_BN, false);
  CUtensorMap mapBlo = mapB;
  const float* blo = g_mlp_presplit ? a.B_lo : nullptr;
  p.b_lo_tma = (blo != nullptr && al16p(blo)) ? 1 : 0;
  p.do_colsum = (a.colsum != nullptr && p.b_mn && EPI == EPI_ATOMIC && !p.b_lo_tma) ? 1 : 0;
  if (p.b_lo_tma) {
    if (p.b_3d) ok = ok && make_map3(&mapBlo, blo, a.N, a.K, a.b_rs, TC_BN / 32);
    else if (p.b_mn) ok = ok && make_map(&mapBlo, blo, a.N, a.K, a.b_rs, 32, TC_BK, true);
    else ok = ok && make_map(&mapBlo, blo, a.K, a.N, a.b_cs, TC_BK, TC_BN, false);
  }
  if (!ok) return NDJIR_ERR_ARG;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  p.m_tiles = (a.M + TC_BM - 1) / TC_BM;
  p.n_tiles = (a.N + TC_BN - 1) / TC_BN;
  p.nkb_total = (a.K + TC_BK - 1) / TC_BK;
  int splits = a.split_k > 1 ? a.split_k : 1;
  p.kb_per_split = (p.nkb_total + splits - 1) / splits;
  p.splits = (p.nkb_total + p.kb_per_split - 1) / p.kb_per_split;   // no empty splits
  int n_items = p.m_tiles * p.n_tiles * p.splits;
  int grid = n_items < NDJIR_NUM_SMS ? n_items : NDJIR_NUM_SMS;
  gemm_tc_kernel<EPI><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(mapA, mapB, mapBlo, p);
  NDJIR_RETURN_LAST_ERROR();
}

int launch_tc(const Args& a, int epi, cudaStream_t st) {
  switch (epi) {
    case EPI_BIAS: return launch_tc_epi<EPI_BIAS>(a, st);
    case EPI_SOFTPLUS: return launch_tc_epi<EPI_SOFTPLUS>(a, st);
    case EPI_ACCUM: return launch_tc_epi<EPI_ACCUM>(a, st);
    case EPI_MUL_S: if (!a.H) return NDJIR_ERR_ARG; return launch_tc_epi<EPI_MUL_S>(a, st);
    case EPI_ADJ: if (!a.H || !a.U || !a.C2) return NDJIR_ERR_ARG; return launch_tc_epi<EPI_ADJ>(a, st);
    case EPI_ATOMIC: return launch_tc_epi<EPI_ATOMIC>(a, st);
    default: return NDJIR_ERR_ARG;
  }
}

}  // namespace gemm
}  // namespace ndjir

namespace {
// lo = x - trunc_tf32(x): what the in-kernel transform computes (the tensor core itself truncates the raw operand)
__global__ void __launch_bounds__(NDJIR_BLOCK)
split_lo_kernel(long long n, float* __restrict__ lo, const float* __restrict__ x) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = x[i];
    lo[i] = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  }
}
}  // namespace

extern "C" {

int ndjir_split_lo(long long n, float* lo, const float* x, cudaStream_t stream) {
  if (n == 0) return NDJIR_OK;
  if (n < 0 || !lo || !x) return NDJIR_ERR_ARG;
  split_lo_kernel<<<ndjir::grid_for(n), NDJIR_BLOCK, 0, stream>>>(n, lo, x);
  NDJIR_RETURN_LAST_ERROR();
}

}  // extern "C"

