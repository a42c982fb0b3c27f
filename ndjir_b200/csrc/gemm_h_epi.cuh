// Epilogue of the split-fp16 tcgen05 kernels (csrc/gemm_h.cu, csrc/gemm_h2.cu): one accumulator row per thread, read
// from TMEM in 16-column chunks (tcgen05.ld), fused arithmetic, operands and results as fp32 rows or fp16 plane pairs.
#pragma once
#include "gemm_h.cuh"
#include "tc_ptx.cuh"

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


// all 16-column chunks of parity `chalf` of one 128 x n_valid accumulator tile; `m` is this thread's output row
template <int EPI>
__device__ __forceinline__ void epilogue_tile(const HArgs& a, int vec_epi, int dbg, long long m, bool row_ok, int n0,
                                              int n_valid, uint32_t tacc, int chalf, bool need_u, bool need_b,
                                              float inv_ab, float sc, float sc2, float inv_h, float inv_u, float& mx,
                                              float& mx2) {
  using namespace tcp;
  constexpr bool NEED_H = (EPI == EPI_MUL_S || EPI == EPI_ADJ);
  constexpr bool NEED_C = (EPI == EPI_ACCUM);
  const int n_vec = vec_epi ? (n_valid & ~15) : 0;     // whole 16-column chunks take the vector path
  for (int c0 = chalf * 16; c0 < n_valid; c0 += 32) {
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
    tmem_ld16(tacc + (uint32_t)c0, v);
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
      // ragged chunk / unaligned operands: one element at a time
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
        epi_math<EPI>(a, v[e] * inv_ab, h, u, cpv, b, o, o2);
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

}  // namespace gemmh
}  // namespace ndjir
