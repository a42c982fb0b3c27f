// MLP products of the split-fp16 engine on the 5th-generation tensor cores (sm_100a).
//
// Operands are stored in HBM as two fp16 planes per matrix (hi, lo; include/ndjir_b200.h), so a 128 x 64 operand tile
// of either plane is fetched by ONE TMA request straight into the 128-byte-swizzled layout tcgen05.mma reads: there is
// no in-kernel conversion pass (the 3xTF32 kernel of csrc/gemm_tc.cu spends a third of its shared-memory bandwidth on
// one).  A product is  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (kind::f16, fp32 accumulation in TMEM; the dropped lo*lo term
// is 2^-22 relative) at twice the tf32 rate and half the operand bytes.
//
// Persistent kernel, one CTA per SM, 10 warps:
//   warp 0      TMA producer: ring of 4 slots of 48 KB, a slot = one (A piece, B piece) pair of one 64-wide K block
//   warp 1      one elected thread issues tcgen05.mma 128 x N x 16 into one of two 256-column TMEM accumulators
//   warps 2-9   epilogue straight from TMEM (tcgen05.ld, one output row per thread): bias / softplus / sigmoid factor /
//               second-order term / accumulate / atomics, reading and writing fp32 rows or split-fp16 planes; the
//               epilogue of item i overlaps the main loop of item i+1
// Accumulation order.  The tensor core adds every 128 x N x 16 partial product into the fp32 accumulator with
// truncation, ~2^-24 relative per instruction; 96 instructions per output (3xTF32, K = 256) measured 1.6e-5 on the
// SDF after eight layers, above the 1e-5 forward bar.  `precise` walks the K blocks twice: first the two correction
// products of ALL blocks (their partial sums are 2^-11 of the result, so their truncation is invisible), then the
// hi*hi products: only K/16 accumulations happen at full magnitude.  The hi tiles are fetched twice (L2 hits).
//   K-major mode : A (M x K), B (N x K) row-major planes  -> forward and input-gradient products (B = W^T resp. W)
//   MN-major mode: A (K x M), B (K x N) row-major planes  -> weight gradients A^T dZ, contraction over the sample
//                  rows, split over work items, atomic fp32 epilogue
#include <cudaTypedefs.h>
#include "gemm_h.cuh"
#include "tc_ptx.cuh"
#include "gemm_h_epi.cuh"

namespace ndjir {
namespace gemmh {

using namespace tcp;

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;                          // halfs per K block: 128-byte rows, SWIZZLE_128B
constexpr int A_BYTES = BM * BK * 2;            // 16 KB
constexpr int B_BYTES = BN * BK * 2;            // 32 KB
constexpr int SLOT_BYTES = A_BYTES + B_BYTES;   // 48 KB
constexpr int MAX_NSLOT = 4;
constexpr int EPI_WARPS = 8;
// Per-epilogue configuration.  The products with operand-reading epilogues (sigmoid factor, adjoint) are bound by their
// epilogue, not by the main loop (a 3-slot ring changes their time by < 2 %), so they trade one ring slot for 64 KB of
// TMA staging (4 plane slots of 2 KB per epilogue warp: WIDE, whenever a second operand U or a second result C2 is
// staged); the others keep 4 slots and stage 2 planes (result C, in place over the operand H if there is one).
template <int EPI, bool WIDE>
struct Cfg {
  static constexpr bool HEAVY = (EPI == EPI_MUL_S || EPI == EPI_ADJ);
  static constexpr bool STAGED = (EPI == EPI_BIAS || EPI == EPI_SOFTPLUS || EPI == EPI_ACCUM || HEAVY);
  // weight-gradient products (MN-major, atomic epilogue), WIDE: one work item covers 256 rows of A^T (two 128-row
  // halves, one 256-column accumulator each - all of TMEM), so the dZ tile it stages serves both halves: L2 -> SM
  // traffic per product 805 -> 537 MB.  The accumulators are single-buffered then; with one or two long K ranges per
  // CTA nothing is lost.
  static constexpr bool TALL = (EPI == EPI_ATOMIC) && WIDE;
  static constexpr int TM = TALL ? 2 * BM : BM;
  static constexpr int A_PIECE = TM * BK * 2;
  static constexpr int SLOT = A_PIECE + B_BYTES;
  static constexpr int NSLOT = WIDE ? 3 : 4;
  static constexpr int STG_WARP = STAGED ? (WIDE ? 4 : 2) * STG_PLANE : 0;
  static constexpr int STG_OFF = NSLOT * SLOT;
  static constexpr int SMEM = STG_OFF + EPI_WARPS * STG_WARP + 1024;
};
constexpr int EPI_WARP0 = 2;
constexpr int EPI_THREADS = 256;
constexpr int CS_WARP0 = EPI_WARP0 + EPI_THREADS / 32;   // column-sum warps (weight-gradient products with a bias gradient)
constexpr int CS_WARPS = 4;
constexpr int THREADS = 64 + EPI_THREADS + CS_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int CHUNK_BYTES = BK * 128;           // MN-major tiles: [chunk of 64 mn][k row][128 B]

int g_h_dbg = 0;   // profiling switches: 1 = skip the epilogue's global traffic, 2 = one MMA per K step

struct HParams {
  HArgs a;
  int m_tiles, n_tiles, splits, kb_per_split, nkb_total;
  int a3d, b3d;          // MN-major operand fetched as one 3-D box per slot
  int b_box_rows;        // rows (K-major) / columns (MN-major) of B staged per slot: 64, 128 or 256
  int vec_epi;           // every epilogue operand allows 32-byte accesses at 16-column granularity
  int do_colsum;         // MN-major mode: a.colsum[n] += column sums of B (the bias gradient of the layer)
  int tma_epi;           // operands and results of the epilogue move by TMA through shared memory (gemm_h_epi.cuh)
  int dbg;
};

template <int EPI, bool WIDE>
__global__ void __launch_bounds__(THREADS, 1)
gemm_h_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
              const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
              const __grid_constant__ EpiMaps em, HParams p) {
  constexpr int NSLOT = Cfg<EPI, WIDE>::NSLOT;
  constexpr bool TALL = Cfg<EPI, WIDE>::TALL;
  constexpr int TM = Cfg<EPI, WIDE>::TM;               // rows of the A side per work item
  constexpr int A_PIECE = Cfg<EPI, WIDE>::A_PIECE;
  constexpr int SLOT = Cfg<EPI, WIDE>::SLOT;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_NSLOT + 4 + EPI_WARPS];
  __shared__ uint32_t tmem_base_sh;
  __shared__ float colsum_sh[EPI == EPI_ATOMIC ? BN : 4];   // only the weight-gradient products carry column sums

  const HArgs& a = p.a;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  auto bar_full = [&](int s) { return smem_u32(&bars[s]); };
  auto bar_empty = [&](int s) { return smem_u32(&bars[MAX_NSLOT + s]); };
  auto bar_acc_full = [&](int b) { return smem_u32(&bars[2 * MAX_NSLOT + b]); };
  auto bar_acc_empty = [&](int b) { return smem_u32(&bars[2 * MAX_NSLOT + 2 + b]); };
  auto a_dst = [&](int s) { return smem_base + s * SLOT; };
  auto b_dst = [&](int s) { return smem_base + s * SLOT + A_PIECE; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1 + (p.do_colsum ? CS_WARPS : 0));   // tcgen05.commit (+ one arrival per column-sum warp)
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), EPI_THREADS);
    }
    for (int w = 0; w < EPI_WARPS; ++w) mbar_init(smem_u32(&bars[2 * MAX_NSLOT + 4 + w]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapAhi); prefetch_tmap(&mapAlo); prefetch_tmap(&mapBhi); prefetch_tmap(&mapBlo);
    if (p.tma_epi)
      for (int i = 0; i < 8; ++i) prefetch_tmap(&em.m[i]);
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
  // work item -> (m0, n0, kb0, nkb); m tiles vary fastest so co-running CTAs share the same B tile in L2
  auto item_info = [&](int item, int& m0, int& n0, int& kb0, int& nkb) {
    int mt = item % p.m_tiles;
    int rest = item / p.m_tiles;
    int ntile = rest % p.n_tiles;
    int sp = rest / p.n_tiles;
    // weight gradients: only the m-tile-0 items carry the bias-gradient column sums; rotating the row tile by the round
    // gives every CTA its share of them (co-running neighbours still share one K range, i.e. the same dZ tiles in L2)
    if (a.mn && gridDim.x % p.m_tiles == 0) mt = (mt + item / (int)gridDim.x) % p.m_tiles;
    m0 = mt * TM;
    n0 = ntile * BN;
    kb0 = sp * p.kb_per_split;
    int kb1 = min(p.nkb_total, kb0 + p.kb_per_split);
    nkb = max(0, kb1 - kb0);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int b_chunks = p.b_box_rows / 64;
      const uint32_t tx_bytes = A_PIECE + (uint32_t)p.b_box_rows * 128u;
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int m0, n0, kb0, nkb;
        item_info(item, m0, n0, kb0, nkb);
        auto load = [&](int piece, int kb) {
          const int s = it % NSLOT;
          const uint32_t ph = (it / NSLOT) & 1;
          mbar_wait(bar_empty(s), ph ^ 1);
          mbar_expect_tx(bar_full(s), tx_bytes);
          const CUtensorMap* ma = piece ? &mapAlo : &mapAhi;
          const CUtensorMap* mb = piece ? &mapBlo : &mapBhi;
          const int k0 = kb * BK;
          if (!a.mn) {
            tma_load_2d(a_dst(s), ma, k0, m0, bar_full(s));
            tma_load_2d(b_dst(s), mb, k0, n0, bar_full(s));
          } else {
            if (p.a3d) {
              tma_load_3d(a_dst(s), ma, 0, k0, m0 / 64, bar_full(s));
            } else {
#pragma unroll
              for (int c = 0; c < TM / 64; ++c) tma_load_2d(a_dst(s) + c * CHUNK_BYTES, ma, m0 + c * 64, k0, bar_full(s));
            }
            if (p.b3d) {
              tma_load_3d(b_dst(s), mb, 0, k0, n0 / 64, bar_full(s));
            } else {
              for (int c = 0; c < b_chunks; ++c) tma_load_2d(b_dst(s) + c * CHUNK_BYTES, mb, n0 + c * 64, k0, bar_full(s));
            }
          }
          ++it;
        };
        for (int i = 0; i < nkb; ++i) {
          load(0, kb0 + i);
          load(1, kb0 + i);
        }
        if (a.precise)
          for (int i = 0; i < nkb; ++i) load(0, kb0 + i);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // K-major (SWIZZLE_128B): rows of 128 B, 8-row groups 1024 B apart (SBO), a K step of 16 halfs = +32 B.
      // MN-major (SWIZZLE_128B): k rows of 128 B (64 mn halfs), 8-row k groups 1024 B apart (SBO), 64-wide mn chunks
      // CHUNK_BYTES apart (LBO), a K step of 16 rows = +2048 B.
      const uint32_t lbo = a.mn ? CHUNK_BYTES : 16, sbo = 1024, kstep = a.mn ? 2048 : 32;
      const uint64_t da0 = make_desc(a_dst(0), lbo, sbo, 2), db0 = make_desc(b_dst(0), lbo, sbo, 2);
      const int k_total = a.K;
      uint32_t it = 0, tile_it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int m0, n0, kb0, nkb;
        item_info(item, m0, n0, kb0, nkb);
        if (nkb == 0) continue;
        const int umma_n = (min(BN, a.N - n0) + 15) & ~15;
        const uint32_t idesc = (1u << 4) | ((uint32_t)a.mn << 15) | ((uint32_t)a.mn << 16) |
                               ((uint32_t)(umma_n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const int buf = TALL ? 0 : (tile_it & 1);
        mbar_wait(bar_acc_empty(buf), (TALL ? (tile_it & 1) : ((tile_it >> 1) & 1)) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * BN;
        const bool second = TALL && (m0 + BM < a.M);      // the lower 128 rows of a tall item hold valid outputs
        const uint64_t oh2 = (uint64_t)((2 * CHUNK_BYTES) >> 4);   // descriptor offset of the second 128-row half of A
        uint32_t first = 0;   // accumulate flag of the next instruction
        for (int i = 0; i < nkb; ++i) {
          const int sx = it % NSLOT, sy = (it + 1) % NSLOT;
          mbar_wait(bar_full(sx), (it / NSLOT) & 1);
          mbar_wait(bar_full(sy), ((it + 1) / NSLOT) & 1);
          tc_fence_after();
          const int ksteps = min(BK / 16, (k_total - (kb0 + i) * BK + 15) / 16);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t ox = (uint64_t)((sx * SLOT + ks * kstep) >> 4);
            const uint64_t oy = (uint64_t)((sy * SLOT + ks * kstep) >> 4);
            if (p.dbg & 2) {
              umma_f16(tacc, da0 + ox, db0 + ox, idesc, first);
              first = 1;
            } else {
              umma_f16(tacc, da0 + oy, db0 + ox, idesc, first);   // lo * hi   (small terms first)
              umma_f16(tacc, da0 + ox, db0 + oy, idesc, 1u);      // hi * lo
              if (!a.precise) umma_f16(tacc, da0 + ox, db0 + ox, idesc, 1u);   // hi * hi
              if (second) {
                umma_f16(tacc + BN, da0 + oy + oh2, db0 + ox, idesc, first);
                umma_f16(tacc + BN, da0 + ox + oh2, db0 + oy, idesc, 1u);
                if (!a.precise) umma_f16(tacc + BN, da0 + ox + oh2, db0 + ox, idesc, 1u);
              }
              first = 1;
            }
          }
          umma_commit(bar_empty(sx));
          umma_commit(bar_empty(sy));
          it += 2;
        }
        if (a.precise) {
          for (int i = 0; i < nkb; ++i, ++it) {
            const int s = it % NSLOT;
            mbar_wait(bar_full(s), (it / NSLOT) & 1);
            tc_fence_after();
            const int ksteps = min(BK / 16, (k_total - (kb0 + i) * BK + 15) / 16);
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t o = (uint64_t)((s * SLOT + ks * kstep) >> 4);
              umma_f16(tacc, da0 + o, db0 + o, idesc, 1u);         // hi * hi at full magnitude, last
              if (second) umma_f16(tacc + BN, da0 + o + oh2, db0 + o, idesc, 1u);
            }
            umma_commit(bar_empty(s));
          }
        }
        umma_commit(bar_acc_full(buf));
        ++tile_it;
      }
    }
  } else if (warp >= CS_WARP0) {
    // ===================== column sums of B (MN-major mode): the bias gradient rides on the weight gradient ==========
    // B slots hold [chunk of 64 n][k row][128 B], 16-byte pieces XOR-swizzled with (row & 7).  A thread reads the same
    // PHYSICAL piece of every row of two row classes (row & 7 == cl), so the logical columns it meets never change and
    // its 2 x 8 sums stay in registers for a whole work item; items of m tile 0 carry the sums.
    if (p.do_colsum) {
      const int t = threadIdx.x - CS_WARP0 * 32;         // 0 .. 127
      const int q = t & 31, chunk = q >> 3, pp = q & 7, rg = t >> 5;
      const int b_chunks = p.b_box_rows / 64;
      const float inv_b = 1.f / dev_scalar(a.b_scale);
      for (int c = t; c < BN; c += CS_WARPS * 32) colsum_sh[c] = 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(CS_WARPS * 32) : "memory");
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int m0, n0, kb0, nkb;
        item_info(item, m0, n0, kb0, nkb);
        const bool mine = (m0 == 0) && chunk < b_chunks;
        float cs[2][8];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 8; ++e) cs[j][e] = 0.f;
        for (int i = 0; i < 2 * nkb; ++i, ++it) {          // both slots of every K block (hi and lo planes)
          const int s = it % NSLOT;
          mbar_wait(bar_full(s), (it / NSLOT) & 1);
          if (mine) {
            const uint8_t* bt = smem + s * SLOT + A_PIECE + chunk * CHUNK_BYTES + pp * 16;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int cl = rg + 4 * j;
#pragma unroll
              for (int r8 = 0; r8 < BK / 8; ++r8) {
                const uint4 x = *reinterpret_cast<const uint4*>(bt + (r8 * 8 + cl) * 128);
                const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                  cs[j][2 * e] += f.x;
                  cs[j][2 * e + 1] += f.y;
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_empty(s));
        }
        if (m0 == 0) {
          if (mine) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int col = chunk * 64 + ((pp ^ (rg + 4 * j)) << 3);     // logical piece = physical ^ (row & 7)
#pragma unroll
              for (int e = 0; e < 8; ++e) atomicAdd(&colsum_sh[col + e], cs[j][e]);
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(CS_WARPS * 32) : "memory");
          for (int c = t; c < BN; c += CS_WARPS * 32) {
            const float v = colsum_sh[c];
            colsum_sh[c] = 0.f;
            if (v != 0.f && n0 + c < a.N) atomicAdd(a.colsum + n0 + c, v * inv_b);
          }
          asm volatile("bar.sync 1, %0;" ::"n"(CS_WARPS * 32) : "memory");
        }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global, one output row per thread =====================
    const int q = warp & 3;                       // TMEM sub-partition of this warp: lanes 32q .. 32q+31
    const int chalf = (warp - EPI_WARP0) >> 2;    // 0: even 16-column chunks, 1: odd chunks
    const bool need_u = (EPI == EPI_ADJ) || (EPI == EPI_MUL_S && (a.U.f != nullptr || a.U.hi != nullptr));
    const bool need_b = (EPI == EPI_BIAS || EPI == EPI_SOFTPLUS) && a.bias != nullptr;
    const float inv_ab = 1.f / (dev_scalar(a.a_scale) * dev_scalar(a.b_scale));
    const float sc = dev_scalar(a.C.scale), sc2 = dev_scalar(a.C2.scale);
    const float inv_h = 1.f / dev_scalar(a.H.scale), inv_u = 1.f / dev_scalar(a.U.scale);
    float mx = 0.f, mx2 = 0.f;
    uint32_t tile_it = 0;
    const int ew = warp - EPI_WARP0;
    const uint32_t stg = smem_base + Cfg<EPI, WIDE>::STG_OFF + ew * Cfg<EPI, WIDE>::STG_WARP;
    const uint32_t ld_bar = smem_u32(&bars[2 * MAX_NSLOT + 4 + ew]);
    uint32_t ld_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int m0, n0, kb0, nkb;
      item_info(item, m0, n0, kb0, nkb);
      if (nkb == 0) continue;
      const int n_valid = min(BN, a.N - n0);
      const int buf = TALL ? 0 : (tile_it & 1);
      const uint32_t acc_parity = TALL ? (tile_it & 1) : ((tile_it >> 1) & 1);
      const long long m = m0 + q * 32 + lane;
      const bool row_ok = m < a.M;
      const uint32_t tacc = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
      if (Cfg<EPI, WIDE>::STAGED && p.tma_epi) {
        // whole 32-column groups move by TMA; a ragged remainder (the 213- / 43-wide matrices of the skip layer) takes
        // the direct path: TMA clips a store at 16-byte granularity, so a partial last group would clobber up to 7
        // halves of whatever follows the view in its rows
        const int n_tma = n_valid & ~(STG_COLS - 1);
        if (EPI == EPI_ACCUM)
          epilogue_tile_tma_accum(a, em, p.dbg, stg, ld_bar, ld_phase, bar_acc_full(buf), acc_parity, m0 + q * 32, n0,
                                  n_tma, tacc, chalf, inv_ab);
        else
          epilogue_tile_tma<EPI>(a, em, p.dbg, stg, ld_bar, ld_phase, bar_acc_full(buf), acc_parity, m0 + q * 32, row_ok,
                                 n0, n_tma, tacc, chalf, need_u, need_b, inv_ab, sc, sc2, inv_h, inv_u, mx, mx2);
        if (n_tma < n_valid)
          epilogue_tile<EPI>(a, p.vec_epi, p.dbg, m, row_ok, n0, n_valid, tacc, n_tma + chalf * 16, 32, need_u, need_b,
                             inv_ab, sc, sc2, inv_h, inv_u, mx, mx2);
      } else {
        mbar_wait(bar_acc_full(buf), acc_parity);
        tc_fence_after();
        epilogue_tile<EPI>(a, p.vec_epi, p.dbg, m, row_ok, n0, n_valid, tacc, chalf * 16, 32, need_u, need_b, inv_ab, sc,
                           sc2, inv_h, inv_u, mx, mx2);
        if (TALL && m0 + BM < a.M)
          epilogue_tile<EPI>(a, p.vec_epi, p.dbg, m + BM, m + BM < a.M, n0, n_valid, tacc + BN, chalf * 16, 32, need_u,
                             need_b, inv_ab, sc, sc2, inv_h, inv_u, mx, mx2);
      }
      tc_fence_before();
      mbar_arrive(bar_acc_empty(buf));
      ++tile_it;
    }
    if (Cfg<EPI, WIDE>::STAGED && p.tma_epi && lane == 0) bulk_wait0();     // shared memory stays valid until TMA has read it
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
// host side
// ---------------------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled get_encode() {
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

// K-major plane: `rows` rows of `k` halfs (row stride ld); box = 64 halfs x box_rows rows
bool map_kmajor(CUtensorMap* map, const __half* base, long long k, long long rows, long long ld, int box_rows) {
  PFN_cuTensorMapEncodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// MN-major plane: `krows` rows (the contraction index) of `mn` halfs.  3-D form (64 contiguous halfs, k rows, 64-wide
// chunks): one request fills the whole [chunk][k][64] tile; needs ld % 64 == 0 (columns beyond mn up to ld are padding
// inside the allocation and only reach outputs that are never stored).
static bool map_mnmajor3(CUtensorMap* map, const __half* base, long long krows, long long ld, int box_chunks) {
  PFN_cuTensorMapEncodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[3] = {64, (cuuint64_t)krows, (cuuint64_t)(ld / 64)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, 128};
  cuuint32_t box[3] = {64, (cuuint32_t)BK, (cuuint32_t)box_chunks};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool map_mnmajor2(CUtensorMap* map, const __half* base, long long mn, long long krows, long long ld) {
  PFN_cuTensorMapEncodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)mn, (cuuint64_t)krows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)BK};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// fp32 rows of an accumulated result: boxes of 32 columns (128 bytes) x 32 rows, 128-byte swizzle
static bool map_epi_f32(CUtensorMap* map, const float* base, long long cols, long long rows, long long ld) {
  PFN_cuTensorMapEncodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)STG_COLS, 32u};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// plane of an epilogue operand / result: N columns x M rows of halfs, boxes of 32 columns x 32 rows, 64-byte swizzle
static bool map_epi(CUtensorMap* map, const __half* base, long long cols, long long rows, long long ld) {
  PFN_cuTensorMapEncodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)STG_COLS, 32u};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int g_h_tall = 0;      // 1: weight-gradient work items of 256 rows (measured slower: 3 slots of 64 KB starve the ring)
int g_h_tma_epi = 1;   // 0: every epilogue on the direct (row-per-lane global access) path

template <int EPI, bool WIDE>
static int launch_epi_w(const HArgs& a, cudaStream_t st) {
  constexpr int SMEM_BYTES = Cfg<EPI, WIDE>::SMEM;
  constexpr int TM = Cfg<EPI, WIDE>::TM;
  HParams p;
  p.a = a;
  p.dbg = g_h_dbg;
  const int ncap = a.N < BN ? a.N : BN;
  p.b_box_rows = ncap <= 64 ? 64 : (ncap <= 128 ? 128 : 256);
  auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
  auto ok_op = [&](const Op& o) {     // 32-byte accesses at 16-column granularity
    if (o.hi) return al32(o.hi) && al32(o.lo) && o.ldh % 16 == 0;
    if (o.f) return al32(o.f) && o.ldf % 8 == 0;
    return true;
  };
  p.vec_epi = ok_op(a.C) && ok_op(a.C2) && ok_op(a.H) && ok_op(a.U) && (a.bias == nullptr || al16p(a.bias));
  CUtensorMap mAh, mAl, mBh, mBl;
  bool ok = true;
  p.a3d = p.b3d = 0;
  if (!a.mn) {
    ok = ok && map_kmajor(&mAh, a.Ahi, a.K, a.M, a.lda, BM) && map_kmajor(&mAl, a.Alo, a.K, a.M, a.lda, BM);
    ok = ok && map_kmajor(&mBh, a.Bhi, a.K, a.N, a.ldb, p.b_box_rows) && map_kmajor(&mBl, a.Blo, a.K, a.N, a.ldb, p.b_box_rows);
  } else {
    p.a3d = a.lda % 64 == 0;
    p.b3d = a.ldb % 64 == 0;
    if (p.a3d) ok = ok && map_mnmajor3(&mAh, a.Ahi, a.K, a.lda, TM / 64) && map_mnmajor3(&mAl, a.Alo, a.K, a.lda, TM / 64);
    else ok = ok && map_mnmajor2(&mAh, a.Ahi, a.M, a.K, a.lda) && map_mnmajor2(&mAl, a.Alo, a.M, a.K, a.lda);
    if (p.b3d) ok = ok && map_mnmajor3(&mBh, a.Bhi, a.K, a.ldb, p.b_box_rows / 64) && map_mnmajor3(&mBl, a.Blo, a.K, a.ldb, p.b_box_rows / 64);
    else ok = ok && map_mnmajor2(&mBh, a.Bhi, a.N, a.K, a.ldb) && map_mnmajor2(&mBl, a.Blo, a.N, a.K, a.ldb);
  }
  if (!ok) return NDJIR_ERR_ARG;
  // TMA epilogue: K-major products whose epilogue operands and results are all split-fp16 planes
  EpiMaps em;
  for (int i = 0; i < 8; ++i) em.m[i] = mAh;
  p.tma_epi = 0;
  if (EPI == EPI_ACCUM) {
    if (g_h_tma_epi && !a.mn && a.N >= STG_COLS && a.C.f && al16p(a.C.f) && a.C.ldf % 4 == 0)
      p.tma_epi = map_epi_f32(&em.m[4], a.C.f, a.N, a.M, a.C.ldf) ? 1 : 0;
  } else if (Cfg<EPI, WIDE>::STAGED && g_h_tma_epi && !a.mn && a.N >= STG_COLS) {
    auto plane_ok = [&](const Op& o) { return o.hi && o.lo && al16p(o.hi) && al16p(o.lo) && o.ldh % 8 == 0; };
    auto absent = [&](const Op& o) { return !o.hi && !o.f; };
    bool eligible = plane_ok(a.C) && (a.bias == nullptr || al16p(a.bias));
    if (Cfg<EPI, WIDE>::HEAVY) eligible = eligible && plane_ok(a.H) && (absent(a.U) || plane_ok(a.U));
    if (EPI == EPI_ADJ) eligible = eligible && plane_ok(a.U) && plane_ok(a.C2);
    if (eligible) {
      bool mok = map_epi(&em.m[4], a.C.hi, a.N, a.M, a.C.ldh) && map_epi(&em.m[5], a.C.lo, a.N, a.M, a.C.ldh);
      if (Cfg<EPI, WIDE>::HEAVY) mok = mok && map_epi(&em.m[0], a.H.hi, a.N, a.M, a.H.ldh) && map_epi(&em.m[1], a.H.lo, a.N, a.M, a.H.ldh);
      if (Cfg<EPI, WIDE>::HEAVY && a.U.hi)
        mok = mok && map_epi(&em.m[2], a.U.hi, a.N, a.M, a.U.ldh) && map_epi(&em.m[3], a.U.lo, a.N, a.M, a.U.ldh);
      if (EPI == EPI_ADJ) mok = mok && map_epi(&em.m[6], a.C2.hi, a.N, a.M, a.C2.ldh) && map_epi(&em.m[7], a.C2.lo, a.N, a.M, a.C2.ldh);
      p.tma_epi = mok ? 1 : 0;
    }
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_h_kernel<EPI, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  p.do_colsum = (a.mn && a.colsum != nullptr) ? 1 : 0;
  p.m_tiles = (a.M + TM - 1) / TM;
  p.n_tiles = (a.N + BN - 1) / BN;
  p.nkb_total = (a.K + BK - 1) / BK;
  int splits = (a.mn && a.split_k > 1) ? a.split_k : 1;
  p.kb_per_split = (p.nkb_total + splits - 1) / splits;
  p.splits = (p.nkb_total + p.kb_per_split - 1) / p.kb_per_split;   // no empty splits
  int n_items = p.m_tiles * p.n_tiles * p.splits;
  int grid = n_items < NDJIR_NUM_SMS ? n_items : NDJIR_NUM_SMS;
  gemm_h_kernel<EPI, WIDE><<<grid, THREADS, SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, em, p);
  NDJIR_RETURN_LAST_ERROR();
}

template <int EPI>
static int launch_epi(const HArgs& a, cudaStream_t st) {
  const bool wide = (EPI == EPI_ADJ) || (EPI == EPI_MUL_S && (a.U.hi || a.U.f)) ||
                    (EPI == EPI_ATOMIC && a.mn && a.M > BM && g_h_tall);
  return wide ? launch_epi_w<EPI, true>(a, st) : launch_epi_w<EPI, false>(a, st);
}

int launch_tc(const HArgs& a, cudaStream_t st) {
  if (!get_encode()) return NDJIR_ERR_ARG;
  if (!a.Ahi || !a.Alo || !a.Bhi || !a.Blo) return NDJIR_ERR_ARG;
  if (!al16p(a.Ahi) || !al16p(a.Alo) || !al16p(a.Bhi) || !al16p(a.Blo) || a.lda % 8 != 0 || a.ldb % 8 != 0)
    return NDJIR_ERR_ARG;
  if (a.mn && a.epi != EPI_ATOMIC) return NDJIR_ERR_ARG;
  if (a.epi == EPI_ATOMIC && (a.C.hi || !a.C.f)) return NDJIR_ERR_ARG;
  if (a.epi == EPI_ACCUM && (a.C.hi || !a.C.f)) return NDJIR_ERR_ARG;
  if (a.epi == EPI_MUL_S && !a.H.f && !a.H.hi) return NDJIR_ERR_ARG;
  if (a.epi == EPI_ADJ && ((!a.H.f && !a.H.hi) || (!a.U.f && !a.U.hi) || (!a.C2.f && !a.C2.hi))) return NDJIR_ERR_ARG;
  if (pair_eligible(a)) return launch_pair(a, st);
  if (resident_eligible(a)) return launch_resident(a, st);
  switch (a.epi) {
    case EPI_BIAS: return launch_epi<EPI_BIAS>(a, st);
    case EPI_SOFTPLUS: return launch_epi<EPI_SOFTPLUS>(a, st);
    case EPI_ACCUM: return launch_epi<EPI_ACCUM>(a, st);
    case EPI_MUL_S: if (!a.H.f && !a.H.hi) return NDJIR_ERR_ARG; return launch_epi<EPI_MUL_S>(a, st);
    case EPI_ADJ:
      if ((!a.H.f && !a.H.hi) || (!a.U.f && !a.U.hi) || (!a.C2.f && !a.C2.hi)) return NDJIR_ERR_ARG;
      return launch_epi<EPI_ADJ>(a, st);
    case EPI_ATOMIC: return launch_epi<EPI_ATOMIC>(a, st);
    default: return NDJIR_ERR_ARG;
  }
}

}  // namespace gemmh
}  // namespace ndjir

// ---------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------
namespace {
using ndjir::gemmh::Op;
Op make_op(float* f, long long ldf, const ndjir_hmat& h) {
  Op o;
  o.f = h.hi ? nullptr : f;
  o.ldf = ldf;
  o.hi = reinterpret_cast<__half*>(h.hi);
  o.lo = reinterpret_cast<__half*>(h.lo);
  o.ldh = h.ld;
  o.scale = h.hi ? h.scale : nullptr;
  o.amax = h.hi ? h.amax : nullptr;
  return o;
}
}  // namespace

extern "C" int ndjir_gemm_h(const ndjir_gemm_h_desc* d, cudaStream_t stream) {
  using namespace ndjir::gemmh;
  if (!d) return NDJIR_ERR_ARG;
  if (d->M <= 0 || d->N <= 0) return NDJIR_OK;
  if (d->K < 0) return NDJIR_ERR_ARG;
  HArgs a;
  a.M = d->M; a.N = d->N; a.K = d->K;
  a.mn = d->mn_major ? 1 : 0; a.epi = d->epilogue; a.precise = d->precise ? 1 : 0;
  a.split_k = d->split_k > 1 ? d->split_k : 1;
  a.alpha = d->alpha; a.out_scale = d->out_scale; a.beta = d->beta; a.hscale = d->hscale;
  a.a_scale = d->A.hi ? d->A.scale : nullptr;
  a.b_scale = d->B.hi ? d->B.scale : nullptr;
  a.Ahi = reinterpret_cast<const __half*>(d->A.hi); a.Alo = reinterpret_cast<const __half*>(d->A.lo); a.lda = d->A.ld;
  a.Bhi = reinterpret_cast<const __half*>(d->B.hi); a.Blo = reinterpret_cast<const __half*>(d->B.lo); a.ldb = d->B.ld;
  a.A32 = d->A32; a.a_rs = d->a_rs; a.a_cs = d->a_cs;
  a.B32 = d->B32; a.b_rs = d->b_rs; a.b_cs = d->b_cs;
  a.C = make_op(d->C, d->ldc, d->Ch);
  a.C2 = make_op(d->C2, d->ldc2, d->C2h);
  a.H = make_op(const_cast<float*>(d->H), d->ldh, d->Hh);
  a.U = make_op(const_cast<float*>(d->U), d->ldu, d->Uh);
  a.bias = d->bias;
  a.colsum = d->colsum;
  if (!a.C.f && !a.C.hi) return NDJIR_ERR_ARG;
  if (corner_shape(a)) return launch_corner(a, stream);
  return launch_tc(a, stream);
}
