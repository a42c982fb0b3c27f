// Epilogue of the split-fp16 tcgen05 kernels (csrc/gemm_h.cu, csrc/gemm_h2.cu): one accumulator row per thread, read
// from TMEM in 16-column chunks (tcgen05.ld), fused arithmetic, operands and results as fp32 rows or fp16 plane pairs.
#pragma once
#include "gemm_h.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>

namespace ndjir {
namespace gemmh {

// 256-bit global accesses (sm_100): one output row per lane means one L1 wavefront per lane and instruction, so the
// 32 bytes a lane owns per 16-column chunk of an fp16 plane move in one instruction instead of two
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void ld256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p)
               : "memory");
}
__device__ __forceinline__ void st256(void* p, uint4 a, uint4 b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}


// the 16-column chunks c_begin, c_begin + c_stride, ... of one 128 x n_valid accumulator tile; `m` is this thread's
// output row
template <int EPI>
__device__ __forceinline__ void epilogue_tile(const HArgs& a, int vec_epi, int dbg, long long m, bool row_ok, int n0,
                                              int n_valid, uint32_t tacc, int c_begin, int c_stride, bool need_u,
                                              bool need_b,
                                              float inv_ab, float sc, float sc2, float inv_h, float inv_u, float& mx,
                                              float& mx2, bool two_acc = false, uint32_t tacc2 = 0u) {
  using namespace tcp;
  constexpr bool NEED_H = (EPI == EPI_MUL_S || EPI == EPI_ADJ);
  constexpr bool NEED_C = (EPI == EPI_ACCUM);
  const int n_vec = vec_epi ? (n_valid & ~15) : 0;     // whole 16-column chunks take the vector path
  for (int c0 = c_begin; c0 < n_valid; c0 += c_stride) {
    const bool vec = row_ok && c0 < n_vec;
    const int n = n0 + c0;
    // operands of the fused epilogue for this thread's 16 columns, issued before the accumulator is read
    uint4 hr[4], ur[4], cr[4];
    float4 bv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      hr[j] = ur[j] = cr[j] = make_uint4(0u, 0u, 0u, 0u);
      bv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (vec && !(dbg & 1)) {
      if (NEED_H) {
        if (a.H.hi) {
          ldg256(a.H.hi + m * a.H.ldh + n, hr[0], hr[1]);
          ldg256(a.H.lo + m * a.H.ldh + n, hr[2], hr[3]);
        } else {
          ldg256(a.H.f + m * a.H.ldf + n, hr[0], hr[1]);
          ldg256(a.H.f + m * a.H.ldf + n + 8, hr[2], hr[3]);
        }
      }
      if (need_u) {
        if (a.U.hi) {
          ld256(a.U.hi + m * a.U.ldh + n, ur[0], ur[1]);
          ld256(a.U.lo + m * a.U.ldh + n, ur[2], ur[3]);
        } else {
          ld256(a.U.f + m * a.U.ldf + n, ur[0], ur[1]);
          ld256(a.U.f + m * a.U.ldf + n + 8, ur[2], ur[3]);
        }
      }
      if (NEED_C) {
        ld256(a.C.f + m * a.C.ldf + n, cr[0], cr[1]);
        ld256(a.C.f + m * a.C.ldf + n + 8, cr[2], cr[3]);
      }
      if (need_b) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = __ldg(reinterpret_cast<const float4*>(a.bias + n) + j);
      }
    }
    float v[16];
    if (two_acc) {
      // two accumulators per tile (csrc/gemm_h3.cu): the correction products and the hi*hi products, summed here
      float w[16];
      tmem_ld16_nowait(tacc + (uint32_t)c0, v);
      tmem_ld16_nowait(tacc2 + (uint32_t)c0, w);
      tmem_ld_wait();
      tmem_ld_fence16(v);      // the sums below must not be scheduled above the wait
      tmem_ld_fence16(w);
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] += w[e];
    } else {
      tmem_ld16(tacc + (uint32_t)c0, v);
    }
    if (dbg & 1) { if (v[0] == 1.2345e-30f && row_ok) a.C.f[0] = v[1]; continue; }
    if (vec) {
      float h[16], u[16], cp[16], o[16], o2[16];
      const float* bb = reinterpret_cast<const float*>(bv);
      // decode the raw operand registers
      if (NEED_H) {
        if (a.H.hi) {
          const uint32_t* hh = reinterpret_cast<const uint32_t*>(hr);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float2 t = join2(hh[e], hh[8 + e]);
            h[2 * e] = t.x * inv_h; h[2 * e + 1] = t.y * inv_h;
          }
        } else {
          const float* hf = reinterpret_cast<const float*>(hr);
#pragma unroll
          for (int e = 0; e < 16; ++e) h[e] = hf[e];
        }
      }
      if (need_u) {
        if (a.U.hi) {
          const uint32_t* uu = reinterpret_cast<const uint32_t*>(ur);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float2 t = join2(uu[e], uu[8 + e]);
            u[2 * e] = t.x * inv_u; u[2 * e + 1] = t.y * inv_u;
          }
        } else {
          const float* uf = reinterpret_cast<const float*>(ur);
#pragma unroll
          for (int e = 0; e < 16; ++e) u[e] = uf[e];
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) u[e] = 0.f;
      }
      if (NEED_C) {
        const float* cf = reinterpret_cast<const float*>(cr);
#pragma unroll
        for (int e = 0; e < 16; ++e) cp[e] = cf[e];
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        epi_math<EPI>(a, v[e] * inv_ab, NEED_H ? h[e] : 0.f, u[e], NEED_C ? cp[e] : 0.f, bb[e], o[e], o2[e]);
        mx = fmaxf(mx, fabsf(o[e]));
        if (EPI == EPI_ADJ) mx2 = fmaxf(mx2, fabsf(o2[e]));
      }
      // store
      if (EPI == EPI_ATOMIC) {
        float* cp_ = a.C.f + m * a.C.ldf + n;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(cp_ + 4 * j), "f"(o[4 * j]),
                       "f"(o[4 * j + 1]), "f"(o[4 * j + 2]), "f"(o[4 * j + 3])
                       : "memory");
      } else if (a.C.hi) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split2(o[2 * e] * sc, o[2 * e + 1] * sc, hi[e], lo[e]);
        st256(a.C.hi + m * a.C.ldh + n, make_uint4(hi[0], hi[1], hi[2], hi[3]), make_uint4(hi[4], hi[5], hi[6], hi[7]));
        st256(a.C.lo + m * a.C.ldh + n, make_uint4(lo[0], lo[1], lo[2], lo[3]), make_uint4(lo[4], lo[5], lo[6], lo[7]));
      } else {
        const uint4* oq = reinterpret_cast<const uint4*>(o);
        st256(a.C.f + m * a.C.ldf + n, oq[0], oq[1]);
        st256(a.C.f + m * a.C.ldf + n + 8, oq[2], oq[3]);
      }
      if (EPI == EPI_ADJ) {
        if (a.C2.hi) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) split2(o2[2 * e] * sc2, o2[2 * e + 1] * sc2, hi[e], lo[e]);
          st256(a.C2.hi + m * a.C2.ldh + n, make_uint4(hi[0], hi[1], hi[2], hi[3]), make_uint4(hi[4], hi[5], hi[6], hi[7]));
          st256(a.C2.lo + m * a.C2.ldh + n, make_uint4(lo[0], lo[1], lo[2], lo[3]), make_uint4(lo[4], lo[5], lo[6], lo[7]));
        } else {
          const uint4* oq = reinterpret_cast<const uint4*>(o2);
          st256(a.C2.f + m * a.C2.ldf + n, oq[0], oq[1]);
          st256(a.C2.f + m * a.C2.ldf + n + 8, oq[2], oq[3]);
        }
      }
    } else if (row_ok) {
      // ragged chunk / unaligned operands: one element at a time.  The loop indexes the accumulator values
      // dynamically, which puts them in local memory: a private copy made on THIS path only - indexing v itself made
      // every chunk of every tile store its 16 values to the stack (268 MB of L2 write traffic per 262144 x 256 product)
      float vr[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) vr[e] = v[e];
#pragma unroll 1
      for (int e = 0; e < 16; ++e) {
        const int c = c0 + e;
        if (c >= n_valid) break;
        const int nn = n0 + c;
        float h = 0.f, u = 0.f, cpv = 0.f, b = 0.f, o, o2;
        if (NEED_H) h = op_load(a.H, inv_h, m, nn);
        if (need_u) u = op_load(a.U, inv_u, m, nn);
        if (NEED_C) cpv = a.C.f[m * a.C.ldf + nn];
        if (need_b) b = __ldg(a.bias + nn);
        epi_math<EPI>(a, vr[e] * inv_ab, h, u, cpv, b, o, o2);
        mx = fmaxf(mx, fabsf(o));
        if (EPI == EPI_ATOMIC) atomicAdd(a.C.f + m * a.C.ldf + nn, o);
        else op_store(a.C, sc, m, nn, o);
        if (EPI == EPI_ADJ) {
          mx2 = fmaxf(mx2, fabsf(o2));
          op_store(a.C2, sc2, m, nn, o2);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The same epilogue with every operand and result moved by TMA through a per-warp staging buffer in shared memory
// (plane-form operands only; csrc/gemm_h.cu decides per launch).
//
// Why.  With one output row per lane, a 256-bit global access touches 32 different 128-byte lines per warp
// instruction and moves 32 bytes of each: ncu shows the heavy epilogues (sigmoid factor, adjoint: 3-5 plane pairs per
// tile) bound by the number of such sector requests the LSU / L1 keeps in flight (lg_throttle, long scoreboard), not
// by bytes and not by the main loop (a slower main loop does not change their time).  Here a warp owns 32 rows x 32
// columns at a time: lane 0 asks TMA for the 64-byte rows of every operand plane (whole-line requests, no LSU
// involvement), the lanes read their 2 x 16 bytes per plane from the swizzled buffer, compute, write the results IN
// PLACE over the operands they consumed (result C over H, C2 over U) and lane 0 hands the planes to TMA stores.
// Staging layout per warp: plane slots of 2 KB = 32 rows x 64 B, CU_TENSOR_MAP_SWIZZLE_64B (16-byte unit u of row r
// lives at unit u ^ ((r >> 1) & 3): conflict-free for quarter-warps).
// single-instruction forms for the staged epilogue: the epilogue warps are bound by instruction issue (ncu: two warps
// per scheduler issuing every other cycle), so every per-element constant is folded on the host side of the loop
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_fast(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// (x0, x1), already multiplied by the tensor's scale -> packed hi, lo halves; out-of-range values saturate to +-65504
__device__ __forceinline__ void split2_sat(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - b.y), "f"(x0 - b.x));
}
// per-launch constants of the fused arithmetic (results come out already multiplied by the result tensor's scale)
struct EpiK {
  float k0, k1, k2, k3;
};
template <int EPI>
__device__ __forceinline__ EpiK epi_consts(const HArgs& a, float inv_ab, float sc, float sc2, float inv_h, float inv_u) {
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  EpiK k{0.f, 0.f, 0.f, 0.f};
  if (EPI == EPI_BIAS) { k.k0 = a.alpha * inv_ab * sc; k.k1 = sc; }
  else if (EPI == EPI_SOFTPLUS) { k.k0 = inv_ab; k.k1 = a.beta * LOG2E; k.k2 = a.out_scale * LN2 / a.beta * sc; }
  else if (EPI == EPI_MUL_S) { k.k0 = -a.beta * a.hscale * inv_h * LOG2E; k.k1 = a.alpha * inv_ab * sc; k.k2 = inv_u * sc; }
  else if (EPI == EPI_ADJ) {
    k.k0 = -a.beta * a.hscale * inv_h * LOG2E;
    k.k1 = inv_ab * inv_u * a.beta * sc;
    k.k2 = a.out_scale * inv_ab * sc2;
  }
  return k;
}
// one element: accumulator v, raw (hi + lo) operands h, u, bias b -> scaled results os (and os2 for the adjoint)
template <int EPI>
__device__ __forceinline__ void epi_math_scaled(const EpiK& k, float v, float h, float u, float b, float& os, float& os2) {
  os2 = 0.f;
  if (EPI == EPI_BIAS) os = fmaf(v, k.k0, b * k.k1);
  else if (EPI == EPI_SOFTPLUS) {
    const float z = fmaf(v, k.k0, b) * k.k1;               // beta x in base-2 units
    const float l = lg2_fast(1.f + ex2_fast(-fabsf(z)));
    os = (fmaxf(z, 0.f) + l) * k.k2;
  } else if (EPI == EPI_MUL_S) {
    const float e = ex2_fast(h * k.k0);                    // 1 - sigmoid(beta a) = exp(-beta softplus)
    const float t = v * k.k1;
    os = fmaf(u, k.k2, fmaf(-t, e, t));
  } else if (EPI == EPI_ADJ) {
    const float e = ex2_fast(h * k.k0);
    os = (v * u) * (e * k.k1);
    const float t = v * k.k2;
    os2 = fmaf(-t, e, t);
  } else os = v;
}

struct EpiMaps {
  CUtensorMap m[8];      // H.hi, H.lo, U.hi, U.lo, C.hi, C.lo, C2.hi, C2.lo (unused entries: copies of a valid map)
};
constexpr int STG_PLANE = 2048;
constexpr int STG_COLS = 32;

template <int EPI>
__device__ __forceinline__ void epilogue_tile_tma(const HArgs& a, const EpiMaps& em, int dbg, uint32_t stg, uint32_t ld_bar,
                                                  uint32_t& ld_phase, uint32_t acc_bar, uint32_t acc_parity, int mrow,
                                                  bool row_ok, int n0, int n_valid, uint32_t tacc, int chalf, bool need_u,
                                                  bool need_b, float inv_ab, float sc, float sc2, float inv_h,
                                                  float inv_u, float& mx, float& mx2) {
  using namespace tcp;
  constexpr bool NEED_H = (EPI == EPI_MUL_S || EPI == EPI_ADJ);
  const int lane = threadIdx.x & 31;
  const uint32_t sw = (uint32_t)((lane >> 1) & 3);
  const uint32_t row_base = stg + (uint32_t)lane * 64u;
  auto unit = [&](int plane, int u) { return row_base + (uint32_t)plane * STG_PLANE + ((((uint32_t)u) ^ sw) << 4); };
  const bool traffic = !(dbg & 1);
  const EpiK kc = epi_consts<EPI>(a, inv_ab, sc, sc2, inv_h, inv_u);
  float mxs = 0.f, mxs2 = 0.f;
  bool waited_acc = false;
  for (int hg = chalf; hg * STG_COLS < n_valid; hg += 2) {
    const int ncol = n0 + hg * STG_COLS;
    if (lane == 0) bulk_wait_read0();            // the previous group's stores have left the buffer
    __syncwarp();
    if (NEED_H && traffic && lane == 0) {
      mbar_expect_tx(ld_bar, (need_u ? 4u : 2u) * STG_PLANE);
      tma_load_2d(stg, &em.m[0], ncol, mrow, ld_bar);
      tma_load_2d(stg + STG_PLANE, &em.m[1], ncol, mrow, ld_bar);
      if (need_u) {
        tma_load_2d(stg + 2 * STG_PLANE, &em.m[2], ncol, mrow, ld_bar);
        tma_load_2d(stg + 3 * STG_PLANE, &em.m[3], ncol, mrow, ld_bar);
      }
    }
    if (!waited_acc) {                           // the first operands are on their way before the accumulator is awaited
      mbar_wait(acc_bar, acc_parity);
      tc_fence_after();
      waited_acc = true;
    }
    if (NEED_H && traffic) {
      mbar_wait(ld_bar, ld_phase);
      ld_phase ^= 1u;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c0 = hg * STG_COLS + j * 16;
      if (c0 >= n_valid) break;
      const int n = n0 + c0;
      uint4 hr[4], ur[4];
      float4 bv[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        hr[t] = ur[t] = make_uint4(0u, 0u, 0u, 0u);
        bv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (traffic) {
        if (NEED_H) {
          hr[0] = lds128(unit(0, 2 * j)); hr[1] = lds128(unit(0, 2 * j + 1));
          hr[2] = lds128(unit(1, 2 * j)); hr[3] = lds128(unit(1, 2 * j + 1));
        }
        if (need_u) {
          ur[0] = lds128(unit(2, 2 * j)); ur[1] = lds128(unit(2, 2 * j + 1));
          ur[2] = lds128(unit(3, 2 * j)); ur[3] = lds128(unit(3, 2 * j + 1));
        }
        if (need_b) {
#pragma unroll
          for (int t = 0; t < 4; ++t) bv[t] = __ldg(reinterpret_cast<const float4*>(a.bias + n) + t);
        }
      }
      float v[16];
      tmem_ld16(tacc + (uint32_t)c0, v);
      if (!traffic) { if (v[0] == 1.2345e-30f && row_ok) a.C.hi[0] = __float2half(v[1]); continue; }
      const float* bb = reinterpret_cast<const float*>(bv);
      const uint32_t* hh = reinterpret_cast<const uint32_t*>(hr);
      const uint32_t* uu = reinterpret_cast<const uint32_t*>(ur);
      uint32_t hi[8], lo[8], hi2[8], lo2[8];
      float cmx = 0.f, cmx2 = 0.f;                 // running max of the SCALED results of this chunk
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float2 h = make_float2(0.f, 0.f), u = make_float2(0.f, 0.f);
        if (NEED_H) h = join2(hh[e], hh[8 + e]);
        if (need_u) u = join2(uu[e], uu[8 + e]);
        float o0, o1, p0, p1;
        epi_math_scaled<EPI>(kc, v[2 * e], h.x, u.x, bb[2 * e], o0, p0);
        epi_math_scaled<EPI>(kc, v[2 * e + 1], h.y, u.y, bb[2 * e + 1], o1, p1);
        cmx = fmaxf(cmx, fmaxf(fabsf(o0), fabsf(o1)));
        split2_sat(o0, o1, hi[e], lo[e]);
        if (EPI == EPI_ADJ) {
          cmx2 = fmaxf(cmx2, fmaxf(fabsf(p0), fabsf(p1)));
          split2_sat(p0, p1, hi2[e], lo2[e]);
        }
      }
      if (row_ok) {
        mxs = fmaxf(mxs, cmx);
        mxs2 = fmaxf(mxs2, cmx2);
      }
      sts128(unit(0, 2 * j), make_uint4(hi[0], hi[1], hi[2], hi[3]));
      sts128(unit(0, 2 * j + 1), make_uint4(hi[4], hi[5], hi[6], hi[7]));
      sts128(unit(1, 2 * j), make_uint4(lo[0], lo[1], lo[2], lo[3]));
      sts128(unit(1, 2 * j + 1), make_uint4(lo[4], lo[5], lo[6], lo[7]));
      if (EPI == EPI_ADJ) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { hi[e] = hi2[e]; lo[e] = lo2[e]; }
        sts128(unit(2, 2 * j), make_uint4(hi[0], hi[1], hi[2], hi[3]));
        sts128(unit(2, 2 * j + 1), make_uint4(hi[4], hi[5], hi[6], hi[7]));
        sts128(unit(3, 2 * j), make_uint4(lo[0], lo[1], lo[2], lo[3]));
        sts128(unit(3, 2 * j + 1), make_uint4(lo[4], lo[5], lo[6], lo[7]));
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (traffic && lane == 0) {
      tma_store_2d(&em.m[4], ncol, mrow, stg);
      tma_store_2d(&em.m[5], ncol, mrow, stg + STG_PLANE);
      if (EPI == EPI_ADJ) {
        tma_store_2d(&em.m[6], ncol, mrow, stg + 2 * STG_PLANE);
        tma_store_2d(&em.m[7], ncol, mrow, stg + 3 * STG_PLANE);
      }
      bulk_commit();
    }
  }
  if (!waited_acc) {      // a tile without columns for this warp still takes part in the accumulator hand-over
    mbar_wait(acc_bar, acc_parity);
    tc_fence_after();
  }
  mx = fmaxf(mx, mxs / sc);
  if (EPI == EPI_ADJ) mx2 = fmaxf(mx2, mxs2 / sc2);
}

// C (fp32 rows) += alpha * acc through the same staging: a warp's 32 rows x 32 columns of C are ONE 4 KB box of
// 128-byte rows (CU_TENSOR_MAP_SWIZZLE_128B: 16-byte unit u of row r lives at u ^ (r & 7)), loaded, updated in place and
// stored by TMA (em.m[4] is the fp32 map of C).  The input-gradient products that collect the heads' contributions
// into dO use this epilogue.
__device__ __forceinline__ void epilogue_tile_tma_accum(const HArgs& a, const EpiMaps& em, int dbg, uint32_t stg,
                                                        uint32_t ld_bar, uint32_t& ld_phase, uint32_t acc_bar,
                                                        uint32_t acc_parity, int mrow, int n0, int n_valid, uint32_t tacc,
                                                        int chalf, float inv_ab) {
  using namespace tcp;
  const int lane = threadIdx.x & 31;
  const uint32_t sw = (uint32_t)(lane & 7);
  const uint32_t row_base = stg + (uint32_t)lane * 128u;
  auto unit = [&](int u) { return row_base + ((((uint32_t)u) ^ sw) << 4); };
  const bool traffic = !(dbg & 1);
  const float k = a.alpha * inv_ab;
  bool waited_acc = false;
  for (int hg = chalf; hg * STG_COLS < n_valid; hg += 2) {
    const int ncol = n0 + hg * STG_COLS;
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
    if (traffic && lane == 0) {
      mbar_expect_tx(ld_bar, 2u * STG_PLANE);
      tma_load_2d(stg, &em.m[4], ncol, mrow, ld_bar);
    }
    if (!waited_acc) {
      mbar_wait(acc_bar, acc_parity);
      tc_fence_after();
      waited_acc = true;
    }
    if (traffic) {
      mbar_wait(ld_bar, ld_phase);
      ld_phase ^= 1u;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c0 = hg * STG_COLS + j * 16;
      if (c0 >= n_valid) break;
      float v[16];
      tmem_ld16(tacc + (uint32_t)c0, v);
      if (!traffic) { if (v[0] == 1.2345e-30f) a.C.f[0] = v[1]; continue; }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        uint4 c = lds128(unit(4 * j + t));
        c.x = __float_as_uint(fmaf(v[4 * t], k, __uint_as_float(c.x)));
        c.y = __float_as_uint(fmaf(v[4 * t + 1], k, __uint_as_float(c.y)));
        c.z = __float_as_uint(fmaf(v[4 * t + 2], k, __uint_as_float(c.z)));
        c.w = __float_as_uint(fmaf(v[4 * t + 3], k, __uint_as_float(c.w)));
        sts128(unit(4 * j + t), c);
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (traffic && lane == 0) {
      tma_store_2d(&em.m[4], ncol, mrow, stg);
      bulk_commit();
    }
  }
  if (!waited_acc) {
    mbar_wait(acc_bar, acc_parity);
    tc_fence_after();
  }
}

}  // namespace gemmh
}  // namespace ndjir
