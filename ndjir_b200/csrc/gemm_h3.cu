// Resident-weight variant of the split-fp16 tcgen05 product kernel (sm_100a): K-major products with K <= 256, i.e. the
// forward, input-gradient and adjoint products of the MLP passes (activation rows x a weight matrix).
//
// Why.  csrc/gemm_h.cu streams a (128 x 64 A piece, 256 x 64 B piece) pair per ring slot, so every 128-row tile
// re-fetches the whole weight matrix from L2: 384 KB per tile (576 KB with the `precise` second walk) of which only
// 128 KB are activations.  Measured, the main loop runs at the L2 -> SM fill rate of the chip (~37 B/clk/SM of the
// ~42 B/clk/SM slice throughput cap), not at the tensor-pipe rate.  Here the weights stay put:
//   * a CTA owns ONE 128-column block of the weight matrix for the whole launch: both planes of all K blocks
//     (<= 4 x 2 x 16 KB = 128 KB) are fetched once and stay in shared memory;
//   * only activations stream: a ring of 3 slots of 32 KB (the hi and lo piece of one 64-wide K block of a 128-row
//     tile); the two CTAs that own the two column blocks of a 256-wide layer walk the same row tiles side by side,
//     so the second fetch of a tile is an L2 hit: L2 -> SM traffic per 128 x 256 output drops from 384 / 576 KB to
//     256 KB, HBM traffic is unchanged;
//   * FOUR 128-column TMEM accumulators: per tile one collects the two correction products (A_lo B_hi + A_hi B_lo),
//     the other the hi*hi products, and the epilogue adds them.  That is the accumulation order `precise` obtains in
//     gemm_h.cu by walking the K blocks twice (only K/16 truncating accumulations at full magnitude) without fetching
//     anything twice, so every product of this kernel is `precise`; the two pairs double-buffer epilogue and main loop.
// Warps: 0 = TMA producer, 1 = MMA issuer (one elected thread, 128 x N x 16 instructions), 2-17 = epilogue (code shared
// with gemm_h.cu: gemm_h_epi.cuh).  Sixteen epilogue warps because the epilogue, not the main loop, bounds these products:
// one output row per thread is a dependent chain (operand loads -> tcgen05.ld -> arithmetic -> stores) and the bytes in
// flight per SM are what the threads hold.
#include <cudaTypedefs.h>
#include "gemm_h.cuh"
#include "tc_ptx.cuh"
#include "gemm_h_epi.cuh"

namespace ndjir {
namespace gemmh {

using namespace tcp;

PFN_cuTensorMapEncodeTiled get_encode();                                                   // gemm_h.cu
bool map_kmajor(CUtensorMap* map, const __half* base, long long k, long long rows, long long ld, int box_rows);

namespace res {
constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 64;
constexpr int PIECE = BM * BK * 2;              // 16 KB: 128 rows x 128 B
constexpr int A_SLOT = 2 * PIECE;               // hi | lo
constexpr int NSLOT = 3;
constexpr int MAX_KB = 4;                       // K <= 256
constexpr int B_OFF = NSLOT * A_SLOT;           // resident weights: [kb][plane] pieces
constexpr int SMEM_BYTES = B_OFF + MAX_KB * 2 * PIECE + 1024;
constexpr int EPI_WARP0 = 2;
constexpr int EPI_THREADS = 512;          // 16 warps: (TMEM sub-partition, column quarter)
constexpr int THREADS = 64 + EPI_THREADS;
constexpr int TMEM_COLS = 512;
}  // namespace res

int g_h_resident = 0;   // measured: the 128-column instructions saturate the shared-memory read port; off

struct RParams {
  HArgs a;
  int m_tiles, n_tiles, ctas_per_n, nkb;
  int b_box_rows;
  int vec_epi;
  int dbg;
};

template <int EPI>
__global__ void __launch_bounds__(res::THREADS, 1)
gemm_h_res_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                  const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, RParams p) {
  using namespace res;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * NSLOT + 5];
  __shared__ uint32_t tmem_base_sh;

  const HArgs& a = p.a;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  auto bar_full = [&](int s) { return smem_u32(&bars[s]); };
  auto bar_empty = [&](int s) { return smem_u32(&bars[NSLOT + s]); };
  auto bar_acc_full = [&](int b) { return smem_u32(&bars[2 * NSLOT + b]); };
  auto bar_acc_empty = [&](int b) { return smem_u32(&bars[2 * NSLOT + 2 + b]); };
  const uint32_t bar_b = smem_u32(&bars[2 * NSLOT + 4]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), EPI_THREADS);
    }
    mbar_init(bar_b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapAhi); prefetch_tmap(&mapAlo); prefetch_tmap(&mapBhi); prefetch_tmap(&mapBlo);
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

  // this CTA's column block and its share of the row tiles; the CTAs of one row-tile index sit next to each other
  const int nt = blockIdx.x % p.n_tiles;
  const int mt0 = blockIdx.x / p.n_tiles;
  const int n0 = nt * BN;
  const int nkb = p.nkb;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && mt0 < p.m_tiles) {
      const uint32_t b_piece = (uint32_t)p.b_box_rows * 128u;
      mbar_expect_tx(bar_b, (uint32_t)nkb * 2u * b_piece);
      for (int kb = 0; kb < nkb; ++kb) {
        tma_load_2d(smem_base + B_OFF + (kb * 2 + 0) * PIECE, &mapBhi, kb * BK, n0, bar_b);
        tma_load_2d(smem_base + B_OFF + (kb * 2 + 1) * PIECE, &mapBlo, kb * BK, n0, bar_b);
      }
      uint32_t it = 0;
      const int pf = p.dbg >> 4;      // experiment: L2 prefetch distance in row tiles (0 = off)
      for (int mt = mt0; mt < p.m_tiles; mt += p.ctas_per_n) {
        if (pf && nt == 0) {
          const int mp = mt + pf * p.ctas_per_n;
          if (mp < p.m_tiles)
            for (int kb = 0; kb < nkb; ++kb) {
              tma_prefetch_l2_2d(&mapAhi, kb * BK, mp * BM);
              tma_prefetch_l2_2d(&mapAlo, kb * BK, mp * BM);
            }
        }
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % NSLOT;
          mbar_wait(bar_empty(s), ((it / NSLOT) & 1) ^ 1);
          mbar_expect_tx(bar_full(s), (uint32_t)A_SLOT);
          tma_load_2d(smem_base + s * A_SLOT, &mapAhi, kb * BK, mt * BM, bar_full(s));
          tma_load_2d(smem_base + s * A_SLOT + PIECE, &mapAlo, kb * BK, mt * BM, bar_full(s));
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && mt0 < p.m_tiles) {
      // K-major SWIZZLE_128B pieces: rows of 128 B, 8-row groups 1024 B apart (SBO), a K step of 16 halfs = +32 B
      const uint64_t d0 = make_desc(smem_base, 16, 1024, 2);
      const int umma_n = (min(BN, a.N - n0) + 15) & ~15;
      const uint32_t idesc = (1u << 4) | ((uint32_t)(umma_n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      mbar_wait(bar_b, 0);
      uint32_t it = 0, tile_it = 0;
      for (int mt = mt0; mt < p.m_tiles; mt += p.ctas_per_n, ++tile_it) {
        const int buf = tile_it & 1;
        mbar_wait(bar_acc_empty(buf), ((tile_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t t_small = tmem_base + buf * 2 * BN, t_big = t_small + BN;
        uint32_t acc = 0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % NSLOT;
          mbar_wait(bar_full(s), (it / NSLOT) & 1);
          tc_fence_after();
          const int ksteps = min(BK / 16, (a.K - kb * BK + 15) / 16);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t ah = d0 + (uint64_t)((s * A_SLOT + ks * 32) >> 4);
            const uint64_t al = d0 + (uint64_t)((s * A_SLOT + PIECE + ks * 32) >> 4);
            const uint64_t bh = d0 + (uint64_t)((B_OFF + (kb * 2 + 0) * PIECE + ks * 32) >> 4);
            const uint64_t bl = d0 + (uint64_t)((B_OFF + (kb * 2 + 1) * PIECE + ks * 32) >> 4);
            if (!(p.dbg & 2)) {
              umma_f16(t_small, al, bh, idesc, acc);   // lo * hi
              umma_f16(t_small, ah, bl, idesc, 1u);    // hi * lo
            }
            umma_f16(t_big, ah, bh, idesc, acc);       // hi * hi in its own accumulator
            acc = 1;
          }
          umma_commit(bar_empty(s));
        }
        umma_commit(bar_acc_full(buf));
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global, one output row per thread =====================
    const int q = warp & 3;                       // TMEM sub-partition of this warp: lanes 32q .. 32q+31
    const int cq = (warp - EPI_WARP0) >> 2;       // 16-column chunks cq, cq + 4, ... of the tile
    const bool need_u = (EPI == EPI_ADJ) || (EPI == EPI_MUL_S && (a.U.f != nullptr || a.U.hi != nullptr));
    const bool need_b = (EPI == EPI_BIAS || EPI == EPI_SOFTPLUS) && a.bias != nullptr;
    const float inv_ab = 1.f / (dev_scalar(a.a_scale) * dev_scalar(a.b_scale));
    const float sc = dev_scalar(a.C.scale), sc2 = dev_scalar(a.C2.scale);
    const float inv_h = 1.f / dev_scalar(a.H.scale), inv_u = 1.f / dev_scalar(a.U.scale);
    const int n_valid = min(BN, a.N - n0);
    float mx = 0.f, mx2 = 0.f;
    uint32_t tile_it = 0;
    for (int mt = mt0; mt < p.m_tiles; mt += p.ctas_per_n, ++tile_it) {
      const int buf = tile_it & 1;
      const long long m = (long long)mt * BM + q * 32 + lane;
      const bool row_ok = m < a.M;
      mbar_wait(bar_acc_full(buf), (tile_it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_small = tmem_base + buf * 2 * BN + ((uint32_t)(q * 32) << 16);
      epilogue_tile<EPI>(a, p.vec_epi, p.dbg, m, row_ok, n0, n_valid, t_small + BN, cq * 16, 64, need_u, need_b, inv_ab, sc, sc2,
                         inv_h, inv_u, mx, mx2, !(p.dbg & 2), t_small);
      tc_fence_before();
      mbar_arrive(bar_acc_empty(buf));
    }
    if (EPI != EPI_ATOMIC) {
      amax_commit(a.C.amax, mx);
      if (EPI == EPI_ADJ) amax_commit(a.C2.amax, mx2);
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
bool resident_eligible(const HArgs& a) {
  using namespace res;
  if (!g_h_resident || a.mn || a.epi == EPI_ATOMIC) return false;
  if (a.K > MAX_KB * BK || a.K < 1) return false;
  const int n_tiles = (a.N + BN - 1) / BN;
  if (n_tiles > NDJIR_NUM_SMS) return false;
  return a.M >= 2 * BM;
}

template <int EPI>
static int launch_res(const HArgs& a, cudaStream_t st) {
  using namespace res;
  RParams p;
  p.a = a;
  p.dbg = g_h_dbg;
  const int ncap = a.N < BN ? a.N : BN;
  p.b_box_rows = ncap <= 64 ? 64 : 128;
  auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  auto ok_op = [&](const Op& o) {     // 32-byte accesses at 16-column granularity
    if (o.hi) return al32(o.hi) && al32(o.lo) && o.ldh % 16 == 0;
    if (o.f) return al32(o.f) && o.ldf % 8 == 0;
    return true;
  };
  p.vec_epi = ok_op(a.C) && ok_op(a.C2) && ok_op(a.H) && ok_op(a.U) && (a.bias == nullptr || al16(a.bias));
  CUtensorMap mAh, mAl, mBh, mBl;
  bool ok = map_kmajor(&mAh, a.Ahi, a.K, a.M, a.lda, BM) && map_kmajor(&mAl, a.Alo, a.K, a.M, a.lda, BM) &&
            map_kmajor(&mBh, a.Bhi, a.K, a.N, a.ldb, p.b_box_rows) && map_kmajor(&mBl, a.Blo, a.K, a.N, a.ldb, p.b_box_rows);
  if (!ok) return NDJIR_ERR_ARG;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_h_res_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  p.m_tiles = (a.M + BM - 1) / BM;
  p.n_tiles = (a.N + BN - 1) / BN;
  p.nkb = (a.K + BK - 1) / BK;
  const int per_n = NDJIR_NUM_SMS / p.n_tiles;
  p.ctas_per_n = per_n < p.m_tiles ? per_n : p.m_tiles;
  const int grid = p.ctas_per_n * p.n_tiles;
  gemm_h_res_kernel<EPI><<<grid, THREADS, SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, p);
  NDJIR_RETURN_LAST_ERROR();
}

int launch_resident(const HArgs& a, cudaStream_t st) {
  switch (a.epi) {
    case EPI_BIAS: return launch_res<EPI_BIAS>(a, st);
    case EPI_SOFTPLUS: return launch_res<EPI_SOFTPLUS>(a, st);
    case EPI_ACCUM: return launch_res<EPI_ACCUM>(a, st);
    case EPI_MUL_S: return launch_res<EPI_MUL_S>(a, st);
    case EPI_ADJ: return launch_res<EPI_ADJ>(a, st);
    default: return NDJIR_ERR_ARG;
  }
}

}  // namespace gemmh
}  // namespace ndjir
