// CTA-pair variant of the split-fp16 product kernel (csrc/gemm_h.cu) for the activation-row products (K-major mode,
// many rows): tcgen05 cta_group::2.
//
// ncu on the single-CTA kernel (profiles/r2_gemm_h_ncu.md): no unit is saturated - tensor pipe 36 %, L1/shared data
// pipe 60 %, L2 48 %, DRAM 39 % - because every SM re-fetches the whole 256-column weight tile (2/3 of the bytes it
// stages) from L2 for each 128-row tile, and only two 96 KB K-block pairs fit in flight.  Here the two SMs of a TPC
// form a cluster and compute one 256 x N tile: each CTA stages ITS 128 rows of A and ITS HALF of the weight rows, the
// leader's elected thread issues 256 x N x 16 MMAs that read both SMs' shared memory, and every CTA finds the
// accumulator rows of its own 128 rows in its own TMEM.  Per SM and K block: 64 KB staged instead of 96 KB, 8 KB read
// per MMA instead of 12 KB, six 32 KB slots (three K-block pairs) in flight instead of four 48 KB slots (two).
//
// Synchronisation (no data crosses the cluster through memory):
//   full[s]      each CTA's own TMA transaction barrier
//   pfull[s]     LEADER only: the peer's relay thread (its otherwise idle MMA warp) arrives here remotely once the
//                peer's slot has landed, so the leader waits on full[s] and pfull[s] before issuing
//   empty[s]     both CTAs: the leader's tcgen05.commit multicast frees the slot in both
//   acc_full[b]  both CTAs: commit multicast; acc_empty[b]: LEADER only, 2 x 256 epilogue threads (peer: remote arrive)
#include <cudaTypedefs.h>
#include "gemm_h.cuh"
#include "tc_ptx.cuh"
#include "gemm_h_epi.cuh"

namespace ndjir {
namespace gemmh {

using namespace tcp;

namespace {

constexpr int BM = 128;                         // rows per CTA (256 per pair)
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;            // 16 KB
constexpr int BH_BYTES = (BN / 2) * BK * 2;     // 16 KB: this CTA's half of the weight rows
constexpr int SLOT_BYTES = A_BYTES + BH_BYTES;  // 32 KB
constexpr int NSLOT = 6;
constexpr int SMEM_BYTES = NSLOT * SLOT_BYTES + 1024;
constexpr int EPI_WARP0 = 2;
constexpr int EPI_THREADS = 256;
constexpr int THREADS = 64 + EPI_THREADS;
constexpr int TMEM_COLS = 512;

struct H2Params {
  HArgs a;
  int pair_tiles, n_tiles, nkb;
  int bh_box_rows;       // weight rows staged per CTA and slot: 32, 64 or 128
  int vec_epi;
  int dbg;
};

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
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1)
gemm_h2_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, H2Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * NSLOT + 4];
  __shared__ uint32_t tmem_base_sh;

  const HArgs& a = p.a;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  auto bar_full = [&](int s) { return smem_u32(&bars[s]); };
  auto bar_empty = [&](int s) { return smem_u32(&bars[NSLOT + s]); };
  auto bar_pfull = [&](int s) { return smem_u32(&bars[2 * NSLOT + s]); };
  auto bar_acc_full = [&](int b) { return smem_u32(&bars[3 * NSLOT + b]); };
  auto bar_acc_empty = [&](int b) { return smem_u32(&bars[3 * NSLOT + 2 + b]); };
  auto a_dst = [&](int s) { return smem_base + s * SLOT_BYTES; };
  auto b_dst = [&](int s) { return smem_base + s * SLOT_BYTES + A_BYTES; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
      mbar_init(bar_pfull(s), 1);                      // only the leader's copy is used
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), 2 * EPI_THREADS);    // only the leader's copy is used
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapAhi); prefetch_tmap(&mapAlo); prefetch_tmap(&mapBhi); prefetch_tmap(&mapBlo);
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

  // work items: (pair of m tiles, n tile); clusters walk them round-robin, m pairs fastest
  const int n_items = p.pair_tiles * p.n_tiles;
  const int n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  auto item_info = [&](int item, int& m0, int& n0, int& umma_n) {
    int pt = item % p.pair_tiles;
    int ntile = item / p.pair_tiles;
    m0 = pt * 2 * BM;
    n0 = ntile * BN;
    umma_n = (min(BN, a.N - n0) + 15) & ~15;          // both halves are multiples of 8 weight rows
  };
  const int nkb = p.nkb;
  const int slots_per_item = (a.precise ? 3 : 2) * nkb;

  if (warp == 0) {
    // ===================== TMA producer (each CTA: its A rows, its half of the weight rows) =====================
    if (lane == 0) {
      const uint32_t tx_bytes = A_BYTES + (uint32_t)p.bh_box_rows * 128u;
      uint32_t it = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        int m0, n0, umma_n;
        item_info(item, m0, n0, umma_n);
        const int mr = m0 + (int)rank * BM, nr = n0 + (int)rank * (umma_n / 2);
        auto load = [&](int piece, int kb) {
          const int s = it % NSLOT;
          mbar_wait(bar_empty(s), ((it / NSLOT) & 1) ^ 1);
          mbar_expect_tx(bar_full(s), tx_bytes);
          tma_load_2d(a_dst(s), piece ? &mapAlo : &mapAhi, kb * BK, mr, bar_full(s));
          tma_load_2d(b_dst(s), piece ? &mapBlo : &mapBhi, kb * BK, nr, bar_full(s));
          ++it;
        };
        for (int i = 0; i < nkb; ++i) {
          load(0, i);
          load(1, i);
        }
        if (a.precise)
          for (int i = 0; i < nkb; ++i) load(0, i);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      if (!leader) {
        // ===================== peer: relay "my slot has landed" to the leader =====================
        const uint32_t remote0 = mapa_cluster(bar_pfull(0), 0);
        uint32_t it = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters)
          for (int i = 0; i < slots_per_item; ++i, ++it) {
            const int s = it % NSLOT;
            mbar_wait(bar_full(s), (it / NSLOT) & 1);
            mbar_arrive_cluster(remote0 + s * 8);
          }
      } else {
        // ===================== leader: MMA issuer for the pair =====================
        const uint64_t da0 = make_desc(a_dst(0), 16, 1024, 2), db0 = make_desc(b_dst(0), 16, 1024, 2);
        uint32_t it = 0, tile_it = 0;
        auto wait_slot = [&](uint32_t i_) {
          const int s = i_ % NSLOT;
          const uint32_t ph = (i_ / NSLOT) & 1;
          mbar_wait(bar_full(s), ph);
          mbar_wait_cluster(bar_pfull(s), ph);
        };
        for (int item = cluster_id; item < n_items; item += n_clusters, ++tile_it) {
          int m0, n0, umma_n;
          item_info(item, m0, n0, umma_n);
          const uint32_t idesc = (1u << 4) | ((uint32_t)(umma_n >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
          const int buf = tile_it & 1;
          mbar_wait_cluster(bar_acc_empty(buf), ((tile_it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tacc = tmem_base + buf * BN;
          uint32_t first = 0;
          for (int i = 0; i < nkb; ++i) {
            const int sx = it % NSLOT, sy = (it + 1) % NSLOT;
            wait_slot(it);
            wait_slot(it + 1);
            tc_fence_after();
            const int ksteps = min(BK / 16, (a.K - i * BK + 15) / 16);
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t ox = (uint64_t)((sx * SLOT_BYTES + ks * 32) >> 4);
              const uint64_t oy = (uint64_t)((sy * SLOT_BYTES + ks * 32) >> 4);
              umma_f16_pair(tacc, da0 + oy, db0 + ox, idesc, first);       // lo * hi   (small terms first)
              first = 1;
              umma_f16_pair(tacc, da0 + ox, db0 + oy, idesc, 1u);          // hi * lo
              if (!a.precise) umma_f16_pair(tacc, da0 + ox, db0 + ox, idesc, 1u);   // hi * hi
            }
            umma_commit_pair(bar_empty(sx));
            umma_commit_pair(bar_empty(sy));
            it += 2;
          }
          if (a.precise) {
            for (int i = 0; i < nkb; ++i, ++it) {
              const int s = it % NSLOT;
              wait_slot(it);
              tc_fence_after();
              const int ksteps = min(BK / 16, (a.K - i * BK + 15) / 16);
              for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t o = (uint64_t)((s * SLOT_BYTES + ks * 32) >> 4);
                umma_f16_pair(tacc, da0 + o, db0 + o, idesc, 1u);          // hi * hi at full magnitude, last
              }
              umma_commit_pair(bar_empty(s));
            }
          }
          umma_commit_pair(bar_acc_full(buf));
        }
      }
    }
  } else {
    // ===================== epilogue: this CTA's 128 accumulator rows, all columns =====================
    const int q = warp & 3;
    const int chalf = (warp - EPI_WARP0) >> 2;
    const bool need_u = (EPI == EPI_ADJ) || (EPI == EPI_MUL_S && (a.U.f != nullptr || a.U.hi != nullptr));
    const bool need_b = (EPI == EPI_BIAS || EPI == EPI_SOFTPLUS) && a.bias != nullptr;
    const float inv_ab = 1.f / (dev_scalar(a.a_scale) * dev_scalar(a.b_scale));
    const float sc = dev_scalar(a.C.scale), sc2 = dev_scalar(a.C2.scale);
    const float inv_h = 1.f / dev_scalar(a.H.scale), inv_u = 1.f / dev_scalar(a.U.scale);
    const uint32_t remote_acc_empty0 = mapa_cluster(bar_acc_empty(0), 0);
    float mx = 0.f, mx2 = 0.f;
    uint32_t tile_it = 0;
    for (int item = cluster_id; item < n_items; item += n_clusters, ++tile_it) {
      int m0, n0, umma_n;
      item_info(item, m0, n0, umma_n);
      const int n_valid = min(BN, a.N - n0);
      const int buf = tile_it & 1;
      const long long m = m0 + (int)rank * BM + q * 32 + lane;
      const bool row_ok = m < a.M;
      mbar_wait(bar_acc_full(buf), (tile_it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
      epilogue_tile<EPI>(a, p.vec_epi, p.dbg, m, row_ok, n0, n_valid, tacc, chalf * 16, 32, need_u, need_b, inv_ab, sc, sc2, inv_h,
                         inv_u, mx, mx2);
      tc_fence_before();
      if (leader) mbar_arrive(bar_acc_empty(buf));
      else mbar_arrive_cluster(remote_acc_empty0 + buf * 8);
    }
    amax_commit(a.C.amax, mx);
    if (EPI == EPI_ADJ) amax_commit(a.C2.amax, mx2);
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

PFN_cuTensorMapEncodeTiled get_encode2() {
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

bool map_kmajor2(CUtensorMap* map, const __half* base, long long k, long long rows, long long ld, int box_rows) {
  PFN_cuTensorMapEncodeTiled enc = get_encode2();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int EPI>
int launch_pair_epi(const HArgs& a, cudaStream_t st) {
  H2Params p;
  p.a = a;
  p.dbg = g_h_dbg;
  const int ncap = a.N < BN ? a.N : BN;
  const int umma_n = (ncap + 15) & ~15;
  const int half = umma_n / 2;
  p.bh_box_rows = half <= 32 ? 32 : (half <= 64 ? 64 : 128);
  auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  auto ok_op = [&](const Op& o) {
    if (o.hi) return al32(o.hi) && al32(o.lo) && o.ldh % 16 == 0;
    if (o.f) return al32(o.f) && o.ldf % 8 == 0;
    return true;
  };
  p.vec_epi = ok_op(a.C) && ok_op(a.C2) && ok_op(a.H) && ok_op(a.U) && (a.bias == nullptr || al16(a.bias));
  CUtensorMap mAh, mAl, mBh, mBl;
  bool ok = map_kmajor2(&mAh, a.Ahi, a.K, a.M, a.lda, BM) && map_kmajor2(&mAl, a.Alo, a.K, a.M, a.lda, BM) &&
            map_kmajor2(&mBh, a.Bhi, a.K, a.N, a.ldb, p.bh_box_rows) &&
            map_kmajor2(&mBl, a.Blo, a.K, a.N, a.ldb, p.bh_box_rows);
  if (!ok) return NDJIR_ERR_ARG;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_h2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int m_tiles = (a.M + BM - 1) / BM;
  p.pair_tiles = (m_tiles + 1) / 2;
  p.n_tiles = (a.N + BN - 1) / BN;
  p.nkb = (a.K + BK - 1) / BK;
  const int n_items = p.pair_tiles * p.n_tiles;
  const int clusters = n_items < NDJIR_NUM_SMS / 2 ? n_items : NDJIR_NUM_SMS / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_h2_kernel<EPI>, mAh, mAl, mBh, mBl, p);
  if (e != cudaSuccess) return (int)e;
  NDJIR_RETURN_LAST_ERROR();
}

}  // namespace

int g_h_pair = 0;   // 1: activation-row products with >= 4096 rows run on CTA pairs (measured slower)

bool pair_eligible(const HArgs& a) {
  return g_h_pair && !a.mn && a.M >= 4096 && a.epi != EPI_ATOMIC && get_encode2() != nullptr;
}

int launch_pair(const HArgs& a, cudaStream_t st) {
  switch (a.epi) {
    case EPI_BIAS: return launch_pair_epi<EPI_BIAS>(a, st);
    case EPI_SOFTPLUS: return launch_pair_epi<EPI_SOFTPLUS>(a, st);
    case EPI_ACCUM: return launch_pair_epi<EPI_ACCUM>(a, st);
    case EPI_MUL_S: return launch_pair_epi<EPI_MUL_S>(a, st);
    case EPI_ADJ: return launch_pair_epi<EPI_ADJ>(a, st);
    default: return NDJIR_ERR_ARG;
  }
}

}  // namespace gemmh
}  // namespace ndjir
