// The SDF evaluation of the geometric network as ONE kernel with the activations kept on chip (sm_100a): what the
// forward-only passes of the path need (the SDF-guided sampling rounds of python/sampler.py:190-192, the lattice query
// of python/extract_by_mc.py:47-73), where no layer output has to survive for a backward pass.
//
// The reference evaluates python/network.py:154-232 as eight PF.affine + F.softplus(beta = 100) ops; the layer-wise
// engine (csrc/gemm_h.cu) runs them as eight products, each of which writes its 128 x 256 activation tile to HBM as two
// fp16 planes and reads it back.  Here a CTA owns a 128-row tile for ALL layers:
//   * the tile's current activation lives in shared memory as the A operand of the next product: 4 K blocks x 2 planes
//     (hi, lo) x 16 KB = 128 KB, K-major 128-byte-swizzled rows exactly as TMA would have written them;
//   * the weights stream: a ring of 3 slots of 32 KB (one 256 x 64 piece of one plane of W^T), fetched by TMA while the
//     previous layer's epilogue runs.  CTAs run as clusters of two that walk their tiles in lockstep: each fetches HALF
//     of every weight piece and multicasts it into both CTAs' slots, so a tile-layer pulls 128-192 KB instead of
//     256-384 KB through the L2 -> SM path (which bounded the first version: 27.7 k cycles per tile-layer);
//   * the epilogue reads the 128 x 256 accumulator from TMEM (one row per thread), applies bias + softplus_100 (+ the
//     skip layer's 1/sqrt2 and the concatenation with the encoded input, network.py:171-176), splits the result into
//     hi / lo halves with the layer's power-of-two scale and writes them IN PLACE over the operand tile the finished
//     products have consumed; the last hidden layer's epilogue takes the dot product with the sdf column of the output
//     layer instead (network.py:190-214), so the only global traffic of a tile is its encoded input (128 x 64 halfs x 2)
//     and 128 floats of SDF.
// Products are the three-term split products of gemm_h.cu (A_lo B_hi + A_hi B_lo + A_hi B_hi, optionally in the
// `precise` two-walk order).  Warps: 0 = TMA producer, 1 = MMA issuer, 2-9 = epilogue.
#include <cudaTypedefs.h>
#include "gemm_h.cuh"
#include "tc_ptx.cuh"
#include "gemm_h_epi.cuh"

namespace ndjir {
namespace gemmh {

using namespace tcp;

PFN_cuTensorMapEncodeTiled get_encode();                                                   // gemm_h.cu
bool map_kmajor(CUtensorMap* map, const __half* base, long long k, long long rows, long long ld, int box_rows);

namespace chain {
constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int PIECE = BM * BK * 2;               // 16 KB: one plane of one K block of the activation tile
constexpr int A_BYTES = 4 * 2 * PIECE;           // 128 KB
constexpr int B_SLOT = BN * BK * 2;              // 32 KB
constexpr int NSLOT = 3;
constexpr int SMEM_BYTES = A_BYTES + NSLOT * B_SLOT + 1024;
constexpr int EPI_WARP0 = 2;
constexpr int EPI_THREADS = 256;
constexpr int THREADS = 64 + EPI_THREADS;
constexpr int MAX_LAYERS = 8;

struct LayerP {
  int K, N;                 // inputs (<= 256), outputs (<= 256)
  const float* bias;
  float out_scale;          // 1, or the skip factor when this layer feeds the skip layer
  int append_enc;           // the encoded input follows the N outputs (columns N .. N + din) in the next layer's input
  const float* w_scale;     // scale of the weight planes
  const float* o_scale;     // scale of this layer's output planes (read), and where their running maximum goes
  float* o_amax;
};

struct ChainParams {
  int n_layers, din;
  long long rows;
  LayerP L[MAX_LAYERS];
  const float* enc_scale;   // scale of the encoded-input planes
  const float* enc32;       // the encoded input as fp32 rows (the skip layer's copy is re-split with that layer's scale)
  long long ld_enc32;
  const float* w_sdf;       // sdf column of the output layer: K_last fp32 weights (stride ldw_sdf) and its bias
  long long ldw_sdf;
  const float* b_sdf;
  float* sdf;
  int precise;
};

struct ChainMaps {
  CUtensorMap enc[2];                 // encoded input planes: box 64 x 128 rows
  CUtensorMap w[MAX_LAYERS][2];       // W^T planes per layer: box 64 x 128 rows (one CTA's half of a piece)
};

__global__ void __launch_bounds__(THREADS, 1)
geo_chain_kernel(const __grid_constant__ ChainMaps maps, const ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * NSLOT + 4];
  __shared__ uint32_t tmem_base_sh;
  __shared__ float dot_sh[BM];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar_full = [&](int s) { return smem_u32(&bars[s]); };
  auto bar_empty = [&](int s) { return smem_u32(&bars[NSLOT + s]); };
  const uint32_t bar_a0 = smem_u32(&bars[2 * NSLOT]);          // the tile's encoded input has landed (TMA)
  const uint32_t bar_layer = smem_u32(&bars[2 * NSLOT + 1]);   // an epilogue is done: accumulator free, next operand written
  const uint32_t bar_acc = smem_u32(&bars[2 * NSLOT + 2]);     // a layer's products are complete
  const uint32_t bar_tile = smem_u32(&bars[2 * NSLOT + 3]);    // the last layer's products have consumed the operand tile
  auto a_piece = [&](int kb, int plane) { return smem_base + (uint32_t)((kb * 2 + plane) * PIECE); };
  auto b_slot = [&](int s) { return smem_base + (uint32_t)(A_BYTES + s * B_SLOT); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 2);          // both CTAs of the cluster have consumed the slot (multicast commits)
    }
    mbar_init(bar_a0, 1);
    mbar_init(bar_layer, EPI_THREADS);
    mbar_init(bar_acc, 1);
    mbar_init(bar_tile, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.enc[0]); prefetch_tmap(&maps.enc[1]);
    for (int l = 0; l < p.n_layers; ++l) { prefetch_tmap(&maps.w[l][0]); prefetch_tmap(&maps.w[l][1]); }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                 "r"((uint32_t)BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_barrier();           // both CTAs' barriers exist before any multicast write / remote arrival
  tc_fence_after();
  const uint32_t tacc0 = tmem_base_sh;
  const int n_tiles = (int)((p.rows + BM - 1) / BM);
  const int nkb0 = (p.din + BK - 1) / BK;
  // the two CTAs of a cluster take adjacent tiles and the same number of them (a tile index past the end is a dummy:
  // zero-filled input, nothing stored), so their weight rings advance together
  const uint32_t crank = cluster_rank();
  const int n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  const int iters = (n_tiles + 2 * n_clusters - 1) / (2 * n_clusters);
  auto tile_of = [&](int k) { return (k * n_clusters + cluster_id) * 2 + (int)crank; };

  if (warp == 0) {
    // ===================== TMA producer: the tile's encoded input, then every layer's weight pieces =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile_it = 0; tile_it < iters; ++tile_it) {
        const int tile = tile_of(tile_it);
        if (tile_it) mbar_wait(bar_tile, (tile_it - 1) & 1);      // the previous tile's last products have read the operand tile
        mbar_expect_tx(bar_a0, (uint32_t)(nkb0 * 2 * PIECE));
        for (int kb = 0; kb < nkb0; ++kb) {
          tma_load_2d(a_piece(kb, 0), &maps.enc[0], kb * BK, tile * BM, bar_a0);
          tma_load_2d(a_piece(kb, 1), &maps.enc[1], kb * BK, tile * BM, bar_a0);
        }
        for (int l = 0; l < p.n_layers; ++l) {
          const int nkb = (p.L[l].K + BK - 1) / BK;
          auto load = [&](int plane, int kb) {
            const int s = it % NSLOT;
            mbar_wait(bar_empty(s), ((it / NSLOT) & 1) ^ 1);
            mbar_expect_tx(bar_full(s), (uint32_t)B_SLOT);       // this CTA's half + the peer's half
            tma_load_2d_mc(b_slot(s) + crank * (B_SLOT / 2), &maps.w[l][plane], kb * BK, (int)crank * (BN / 2), bar_full(s),
                           (uint16_t)3);
            ++it;
          };
          for (int kb = 0; kb < nkb; ++kb) { load(0, kb); load(1, kb); }
          if (p.precise)
            for (int kb = 0; kb < nkb; ++kb) load(0, kb);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint64_t d0 = make_desc(smem_base, 16, 1024, 2);      // K-major SWIZZLE_128B, a K step of 16 halfs = +32 B
      uint32_t it = 0, layer_it = 0;
      for (int tile_it = 0; tile_it < iters; ++tile_it) {
        for (int l = 0; l < p.n_layers; ++l, ++layer_it) {
          const LayerP& L = p.L[l];
          const int nkb = (L.K + BK - 1) / BK;
          const int umma_n = (L.N + 15) & ~15;
          const uint32_t idesc = (1u << 4) | ((uint32_t)(umma_n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
          // operand tile ready and accumulator free: layer 0 waits for TMA (and for the previous tile's last epilogue),
          // the others for the previous layer's epilogue
          if (l == 0) mbar_wait(bar_a0, tile_it & 1);
          if (layer_it) mbar_wait(bar_layer, (layer_it - 1) & 1);
          tc_fence_after();
          uint32_t first = 0;
          for (int kb = 0; kb < nkb; ++kb) {
            const int sx = it % NSLOT, sy = (it + 1) % NSLOT;
            mbar_wait(bar_full(sx), (it / NSLOT) & 1);
            mbar_wait(bar_full(sy), ((it + 1) / NSLOT) & 1);
            tc_fence_after();
            const int ksteps = min(BK / 16, (L.K - kb * BK + 15) / 16);
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t ah = d0 + (uint64_t)(((kb * 2 + 0) * PIECE + ks * 32) >> 4);
              const uint64_t al = d0 + (uint64_t)(((kb * 2 + 1) * PIECE + ks * 32) >> 4);
              const uint64_t bh = d0 + (uint64_t)((A_BYTES + sx * B_SLOT + ks * 32) >> 4);
              const uint64_t bl = d0 + (uint64_t)((A_BYTES + sy * B_SLOT + ks * 32) >> 4);
              umma_f16(tacc0, al, bh, idesc, first);     // lo * hi (small terms first)
              first = 1;
              umma_f16(tacc0, ah, bl, idesc, 1u);        // hi * lo
              if (!p.precise) umma_f16(tacc0, ah, bh, idesc, 1u);
            }
            umma_commit_mc(bar_empty(sx), (uint16_t)3);
            umma_commit_mc(bar_empty(sy), (uint16_t)3);
            it += 2;
          }
          if (p.precise) {
            for (int kb = 0; kb < nkb; ++kb, ++it) {
              const int s = it % NSLOT;
              mbar_wait(bar_full(s), (it / NSLOT) & 1);
              tc_fence_after();
              const int ksteps = min(BK / 16, (L.K - kb * BK + 15) / 16);
              for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t ah = d0 + (uint64_t)(((kb * 2 + 0) * PIECE + ks * 32) >> 4);
                const uint64_t bh = d0 + (uint64_t)((A_BYTES + s * B_SLOT + ks * 32) >> 4);
                umma_f16(tacc0, ah, bh, idesc, 1u);      // hi * hi at full magnitude, last
              }
              umma_commit_mc(bar_empty(s), (uint16_t)3);
            }
          }
          umma_commit(bar_acc);
          if (l == p.n_layers - 1) umma_commit(bar_tile);
        }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> the next operand tile in shared memory =====================
    constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    const int q = warp & 3;
    const int chalf = (warp - EPI_WARP0) >> 2;
    const int row = q * 32 + lane;                         // row of the tile = TMEM lane
    const uint32_t tacc = tacc0 + ((uint32_t)(q * 32) << 16);
    const uint32_t sw = (uint32_t)(row & 7);
    uint32_t layer_it = 0;
    for (int tile_it = 0; tile_it < iters; ++tile_it) {
      const long long m = (long long)tile_of(tile_it) * BM + row;
      const bool row_ok = m < p.rows;
#pragma unroll 1
      for (int l = 0; l < p.n_layers; ++l, ++layer_it) {
        const LayerP& L = p.L[l];
        const bool last = l == p.n_layers - 1;
        const float a_scale = dev_scalar(l ? p.L[l - 1].o_scale : p.enc_scale);
        const float inv_ab = 1.f / (a_scale * dev_scalar(L.w_scale));
        const float sc = last ? 1.f : dev_scalar(L.o_scale);
        const float kz = 100.f * LOG2E;                              // beta x in base-2 units
        const float ko = L.out_scale * LN2 / 100.f;                  // softplus_100 back from base 2, times the skip factor
        mbar_wait(bar_acc, layer_it & 1);
        tc_fence_after();
        float dot = 0.f, lmx = 0.f;
#pragma unroll 1
        for (int c = chalf; c < BN / 16; c += 2) {
          const int c0 = c * 16;
          float o[16];
          if (c0 < L.N) {
            float v[16];
            tmem_ld16(tacc + (uint32_t)c0, v);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int n = c0 + e;
              const float b = (L.bias && n < L.N) ? __ldg(L.bias + n) : 0.f;
              const float z = fmaf(v[e], inv_ab, b) * kz;
              const float sp = (fmaxf(z, 0.f) + lg2_fast(1.f + ex2_fast(-fabsf(z)))) * ko;
              o[e] = n < L.N ? sp : 0.f;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = 0.f;
          }
          if (last) {
            // sdf = a . w_sdf + b_sdf (network.py:190-214): the activation never leaves the registers
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (c0 + e < L.N) dot = fmaf(o[e], __ldg(p.w_sdf + (long long)(c0 + e) * p.ldw_sdf), dot);
            continue;
          }
          if (L.append_enc && c0 + 16 > L.N && row_ok) {
            // the skip layer's input is [a | encoded input] / sqrt2 (network.py:171-176)
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int j = c0 + e - L.N;
              if (j >= 0 && j < p.din) o[e] = __ldg(p.enc32 + m * p.ld_enc32 + j) * L.out_scale;
            }
          }
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float x0 = o[2 * e] * sc, x1 = o[2 * e + 1] * sc;
            lmx = fmaxf(lmx, fmaxf(fabsf(x0), fabsf(x1)));
            split2_sat(x0, x1, hi[e], lo[e]);
          }
          // columns c0 .. c0 + 15 of the next operand: K block c0 / 64, 16-byte units (c0 % 64) / 8 and the next one
          const int kb = c0 >> 6, u0 = (c0 & 63) >> 3;
          const uint32_t rbase = (uint32_t)row * 128u;
          const uint32_t ph = a_piece(kb, 0) + rbase, pl = a_piece(kb, 1) + rbase;
          sts128(ph + ((((uint32_t)u0) ^ sw) << 4), make_uint4(hi[0], hi[1], hi[2], hi[3]));
          sts128(ph + ((((uint32_t)u0 + 1) ^ sw) << 4), make_uint4(hi[4], hi[5], hi[6], hi[7]));
          sts128(pl + ((((uint32_t)u0) ^ sw) << 4), make_uint4(lo[0], lo[1], lo[2], lo[3]));
          sts128(pl + ((((uint32_t)u0 + 1) ^ sw) << 4), make_uint4(lo[4], lo[5], lo[6], lo[7]));
        }
        if (!last) amax_commit(L.o_amax, row_ok ? lmx / sc : 0.f);     // running maximum of the layer's output (its next scale)
        if (last) {
          // the two column halves of a row meet in shared memory
          if (chalf == 1) dot_sh[row] = dot;
          asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
          if (chalf == 0 && row_ok) p.sdf[m] = dot + dot_sh[row] + __ldg(p.b_sdf);
        }
        // generic-proxy writes of the operand tile -> visible to the tensor core's (async proxy) reads
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(bar_layer);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_barrier();           // no CTA leaves while its peer may still multicast into it or arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tacc0), "r"((uint32_t)BN) : "memory");
  }
}

}  // namespace chain

int g_h_chain = 0;    // 1: ndjir_geo_sdf_forward runs the whole network in this kernel (measured slower than the layer-wise
                      // products: the weights re-stream from L2 for every 128-row tile)

// The SDF network of `net` over `rows` encoded inputs (planes `ench`, fp32 rows `enc`), layer scales / running maxima
// in act[l & 1] like the layer-wise sequencing.  Returns NDJIR_ERR_ARG when the shapes do not fit the kernel (the caller
// then runs the layers one by one).
int launch_geo_chain(const ndjir_geo_net* net, long long rows, int din, const ndjir_hmat* ench, const float* enc32,
                     long long ld_enc32, const ndjir_hmat* act, float* sdf, cudaStream_t st) {
  using namespace chain;
  if (!g_h_chain || !get_encode()) return NDJIR_ERR_ARG;
  if (net->n_hidden < 1 || net->n_hidden > MAX_LAYERS || din > 2 * BK || rows < BM) return NDJIR_ERR_ARG;
  ChainParams p = {};
  static ChainMaps maps;      // host staging of the kernel's tensor maps (launches are serialised by the caller's stream)
  p.n_layers = net->n_hidden; p.din = din; p.rows = rows;
  p.enc_scale = ench->scale; p.enc32 = enc32; p.ld_enc32 = ld_enc32;
  p.w_sdf = net->sdf.W; p.ldw_sdf = net->sdf.ldw; p.b_sdf = net->sdf.bias; p.sdf = sdf; p.precise = net->precise ? 1 : 0;
  if (net->sdf.N != 1 || !net->sdf.W || !net->sdf.bias) return NDJIR_ERR_ARG;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al16(ench->hi) || !al16(ench->lo) || ench->ld % 8) return NDJIR_ERR_ARG;
  if (!map_kmajor(&maps.enc[0], (const __half*)ench->hi, din, rows, ench->ld, BM) ||
      !map_kmajor(&maps.enc[1], (const __half*)ench->lo, din, rows, ench->ld, BM))
    return NDJIR_ERR_ARG;
  int k_expect = din;
  for (int l = 0; l < net->n_hidden; ++l) {
    const ndjir_mlp_layer& L = net->hidden[l];
    const bool into_skip = (l + 1) == net->skip_layer;
    if (L.K != k_expect || L.K > 4 * BK || L.N > BN || L.N < 16 || !L.Wt.hi || !L.Wt.lo || !al16(L.Wt.hi) ||
        !al16(L.Wt.lo) || L.Wt.ld % 8)
      return NDJIR_ERR_ARG;
    LayerP& o = p.L[l];
    o.K = L.K; o.N = L.N; o.bias = L.bias;
    o.out_scale = into_skip ? net->skip_scale : 1.f;
    o.append_enc = into_skip ? 1 : 0;
    o.w_scale = L.Wt.scale;
    o.o_scale = act[l & 1].scale; o.o_amax = act[l & 1].amax;
    if (!map_kmajor(&maps.w[l][0], (const __half*)L.Wt.hi, L.K, L.N, L.Wt.ld, BN / 2) ||
        !map_kmajor(&maps.w[l][1], (const __half*)L.Wt.lo, L.K, L.N, L.Wt.ld, BN / 2))
      return NDJIR_ERR_ARG;
    k_expect = L.N + (into_skip ? din : 0);
    if (k_expect > 4 * BK) return NDJIR_ERR_ARG;
  }
  if (net->sdf.K != k_expect) return NDJIR_ERR_ARG;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(geo_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int n_tiles = (int)((rows + BM - 1) / BM);
  int n_clusters = (n_tiles + 1) / 2;
  if (n_clusters > NDJIR_NUM_SMS / 2) n_clusters = NDJIR_NUM_SMS / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * n_clusters);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, geo_chain_kernel, maps, p);
  if (e != cudaSuccess) return (int)e;
  NDJIR_RETURN_LAST_ERROR();
}

}  // namespace gemmh
}  // namespace ndjir
