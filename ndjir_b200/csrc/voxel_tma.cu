// Fine-brick TMA sweep of the trilinear voxel gather (D = 4) for LARGE batches of scattered points.
//
// The L2-window sweep of voxel_binned.cu is latency-bound (ncu: DRAM 45 %, L2 51 %, long-scoreboard stalls): every
// point is a dependent chain record -> 8 cells -> store through an L2 with ~5 us loaded latency.  Here the table is
// cut into bricks of 16 x 16 x 16 cells, the points are sorted by the brick of their lower corner in TWO counting
// passes (512 coarse bins through shared-memory staging as before, then <= 1024 fine bins inside each coarse bin's
// L2-resident run), and the sweep runs ONE CTA PER BRICK: a single 3-D TMA box (17 x 17 rows of 17 cells = 272
// contiguous bytes each, 78.6 KB) lands the brick and its upper halo in shared memory, and the brick's ~512 points
// gather their 8 corners from there.  The table is read once, fully coalesced, by the copy engine (2.6 GB incl. halo
// instead of a 3.4 GB latency-bound L2 sweep); two CTAs per SM keep one box in flight while the other computes.
// Same cell arithmetic as every linear family (grid_common.cuh), results identical to the other sweeps up to the
// association of the 8-term sum.
//
// MEASURED (2^24 uniform points, 512^3 x 4, profiles/r2_voxel_tma.md): the sweep itself takes 0.39-0.46 ms when the
// results are written in record order, against 0.95 ms for the L2-window sweep - but 0.94 ms with the results written
// to row `point index`: 16.7 M 16-byte stores scattered over a 256 MB array are read-modify-write DRAM traffic and cost
// ~0.5 ms whichever sweep issues them (an inverse-permutation pass instead: 0.33 ms to build + 0.31 ms to apply).  With
// the second counting pass (0.17 ms) the whole call is 1.32 ms against 1.16 ms for the L2-window sweep, so this path is
// OFF by default (option "voxel_tma"); it stays as the measured answer to "fetch bricks with TMA".
#include <cuda.h>
#include <cudaTypedefs.h>
#include "grid_common.cuh"
#include "voxel_binned.cuh"
#include "tc_ptx.cuh"
#include "../../include/ndjir_b200.h"

namespace ndjir {

int g_voxel_tma = 0;     // 1: fine-brick TMA sweep for the D = 4 forward gather (measured: no faster, see the header)
int g_voxel_tma_bx = 16;     // brick extent along x (8 or 16 cells): smaller boxes, more of them in flight per SM
int g_voxel_tma_l2 = 1;      // L2 promotion of the TMA requests: 0 none, 1 128 B, 2 256 B
int g_voxel_tma_dbg = 0;     // experiment: 1 = write results in record order (wrong rows; isolates the scattered store)

namespace voxel_tma {

using namespace tcp;

constexpr int FB = 16;                         // brick edge in cells of the lower corner
constexpr int BOX = FB + 1;                    // + upper halo
constexpr int ROW_F = BOX * 4;                 // floats per (x, y) row of the box
constexpr int BRICK_BYTES = BOX * BOX * ROW_F * 4;   // 78 608 for 16 x 16 x 16 bricks
constexpr int kMaxFine = 1024;                 // fine bins per coarse bin

struct Layout {
  unsigned fbx;              // brick extent along x (cells)
  unsigned nbx, nby, nbz;    // fine bricks per axis
  unsigned gy;               // fine y-bricks per coarse bin
  unsigned nby1;             // coarse bins along y
  unsigned nb2;              // fine bins per coarse bin = gy * nbz
  unsigned n1, nfine;
};

static bool make_layout(const int* G, Layout& L) {
  L.fbx = g_voxel_tma_bx == 8 ? 8 : FB;
  L.nbx = (G[0] + L.fbx - 1) / L.fbx; L.nby = (G[1] + FB - 1) / FB; L.nbz = (G[2] + FB - 1) / FB;
  L.gy = 1;
  while (L.nbx * ((L.nby + L.gy - 1) / L.gy) > (unsigned)voxel_binned::kMaxBins) L.gy *= 2;
  L.nby1 = (L.nby + L.gy - 1) / L.gy;
  L.nb2 = L.gy * L.nbz;
  L.n1 = L.nbx * L.nby1;
  L.nfine = L.n1 * L.nb2;
  return L.nb2 <= (unsigned)kMaxFine && (long long)L.nfine * 4 + 8 <= voxel_binned::kTailBytes;
}

// second counting pass: one CTA per coarse bin sorts its run of records by fine brick (y within the bin, z) into rec2
// and publishes the fine offsets.  The run (B / 512 records, 512 KB at 2^24 points) stays in L2 between the two reads.
__global__ void __launch_bounds__(1024)
fine_sort_kernel(const float4* __restrict__ rec, float4* __restrict__ rec2, const unsigned* __restrict__ cursors,
                 unsigned* __restrict__ off, GridFrame g, Layout L, long long B) {
  __shared__ unsigned hist[kMaxFine];
  __shared__ unsigned loff[kMaxFine];
  __shared__ unsigned wtot[32];
  const int grp = blockIdx.x, t = threadIdx.x;
  const unsigned start = grp ? cursors[grp - 1] : 0u, end = cursors[grp];
  const unsigned by_base = (grp % L.nby1) * L.gy;
  hist[t] = 0;
  __syncthreads();
  auto fine_of = [&](const float4& r) {
    unsigned i0, i1;
    float w0, w1;
    cell_axis(r.y, g.mny, g.sy, g.gy1, i0, i1, w0, w1);
    unsigned by = i0 / FB - by_base;
    cell_axis(r.z, g.mnz, g.sz, g.gz1, i0, i1, w0, w1);
    return by * L.nbz + i0 / FB;
  };
  for (unsigned i = start + t; i < end; i += 1024) atomicAdd(&hist[fine_of(__ldg(rec + i))], 1u);
  __syncthreads();
  // exclusive scan of 1024 counters (one per thread)
  unsigned v = hist[t], inc = v;
  const int lane = t & 31, w = t >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { unsigned u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
  if (lane == 31) wtot[w] = inc;
  __syncthreads();
  if (w == 0) {
    unsigned x = wtot[lane], xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned u = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += u; }
    wtot[lane] = xi - x;
  }
  __syncthreads();
  const unsigned ex = wtot[w] + inc - v;
  loff[t] = ex;
  if ((unsigned)t < L.nb2) off[(unsigned)grp * L.nb2 + t] = start + ex;
  if (grp == (int)gridDim.x - 1 && t == 0) off[L.nfine] = (unsigned)B;
  __syncthreads();
  for (unsigned i = start + t; i < end; i += 1024) {
    float4 r = __ldg(rec + i);
    unsigned slot = atomicAdd(&loff[fine_of(r)], 1u);
    rec2[start + slot] = r;
  }
}

template <bool ACCUM>
__global__ void __launch_bounds__(256)
brick_gather_kernel(const __grid_constant__ CUtensorMap map, float* __restrict__ out, const float4* __restrict__ rec,
                    const unsigned* __restrict__ off, GridFrame g, Layout L, int dbg) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const unsigned f = blockIdx.x;
  const unsigned start = off[f], end = off[f + 1];
  if (start == end) return;                       // nothing samples this brick: it is never fetched
  const unsigned grp = f / L.nb2, j = f % L.nb2;
  const unsigned bx = grp / L.nby1, by = (grp % L.nby1) * L.gy + j / L.nbz, bz = j % L.nbz;
  const uint32_t brick_bytes = (L.fbx + 1) * BOX * ROW_F * 4;
  float* brick = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const uint32_t bar_a = smem_u32(&bar);
  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar_a, brick_bytes);
    tma_load_3d(smem_u32(brick), &map, (int)(bz * FB * 4), (int)(by * FB), (int)(bx * L.fbx), bar_a);
  }
  __syncthreads();
  // records first (independent of the box), then wait for the brick
  const unsigned x_base = bx * L.fbx, y_base = by * FB, z_base = bz * FB;
  for (unsigned i0 = start + threadIdx.x; i0 < end; i0 += 512) {
    const unsigned i1 = i0 + 256;
    const bool two = i1 < end;
    float4 r0 = __ldg(rec + i0), r1 = __ldg(rec + (two ? i1 : i0));
    float4 p0, p1;
    if (ACCUM) {
      p0 = *reinterpret_cast<const float4*>(out + (long long)__float_as_uint(r0.w) * 4);
      p1 = *reinterpret_cast<const float4*>(out + (long long)__float_as_uint(r1.w) * 4);
    }
    mbar_wait(bar_a, 0);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      const float4 r = u ? r1 : r0;
      Cell c = make_cell_linear(g, r.x, r.y, r.z);
      const unsigned lx0 = c.x0 - x_base, lx1 = c.x1 - x_base, ly0 = c.y0 - y_base, ly1 = c.y1 - y_base;
      const unsigned lz0 = (c.z0 - z_base) * 4, lz1 = (c.z1 - z_base) * 4;
      const float* r00 = brick + (lx0 * BOX + ly0) * ROW_F;
      const float* r01 = brick + (lx0 * BOX + ly1) * ROW_F;
      const float* r10 = brick + (lx1 * BOX + ly0) * ROW_F;
      const float* r11 = brick + (lx1 * BOX + ly1) * ROW_F;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      auto add = [&](const float* row, unsigned lz, float wgt) {
        float4 v = *reinterpret_cast<const float4*>(row + lz);
        acc.x += wgt * v.x; acc.y += wgt * v.y; acc.z += wgt * v.z; acc.w += wgt * v.w;
      };
      add(r00, lz0, c.p0 * c.q0 * c.r0); add(r00, lz1, c.p0 * c.q0 * c.r1);
      add(r01, lz0, c.p0 * c.q1 * c.r0); add(r01, lz1, c.p0 * c.q1 * c.r1);
      add(r10, lz0, c.p1 * c.q0 * c.r0); add(r10, lz1, c.p1 * c.q0 * c.r1);
      add(r11, lz0, c.p1 * c.q1 * c.r0); add(r11, lz1, c.p1 * c.q1 * c.r1);
      if (ACCUM) {
        const float4 pv = u ? p1 : p0;
        acc.x += pv.x; acc.y += pv.y; acc.z += pv.z; acc.w += pv.w;
      }
      // dbg 1 (experiment): rows in record order - isolates the cost of the scattered 16-byte store below
      const long long row = dbg ? (long long)(u ? i1 : i0) : (long long)__float_as_uint(r.w);
      *reinterpret_cast<float4*>(out + row * 4) = acc;
    }
  }
}

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

bool eligible(long long B, const int* G, int D, const float* feat, const float* out) {
  if (!g_voxel_tma || D != 4 || B < (1ll << 16) || !get_encode()) return false;
  if ((reinterpret_cast<uintptr_t>(feat) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return false;
  if (G[0] < FB || G[1] < FB || G[2] < FB) return false;
  if ((G[2] * 16) % 16 != 0) return false;
  Layout L;
  return make_layout(G, L);
}

int query(long long B, float* out, const float* query_, const float* feat, const GridFrame& g, const int* G, bool accum,
          void* ws, long long ws_bytes, cudaStream_t st) {
  Layout L;
  if (!make_layout(G, L)) return NDJIR_ERR_ARG;
  voxel_binned::Bins b;
  b.px = L.fbx; b.py = FB * L.gy; b.nby = L.nby1; b.n = L.n1;
  const float4* rec = nullptr;
  int rc = voxel_binned::build_records_bins(B, query_, nullptr, g, b, ws, ws_bytes, st, &rec);
  if (rc != NDJIR_OK) return rc;
  unsigned* cursors = reinterpret_cast<unsigned*>(ws);
  float4* rec2 = const_cast<float4*>(rec) + B;
  unsigned* off = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws) + voxel_binned::kHeaderBytes + 32 * B);
  fine_sort_kernel<<<L.n1, 1024, 0, st>>>(rec, rec2, cursors, off, g, L, B);
  CUtensorMap map;
  cuuint64_t dims[3] = {(cuuint64_t)G[2] * 4, (cuuint64_t)G[1], (cuuint64_t)G[0]};
  cuuint64_t strides[2] = {(cuuint64_t)G[2] * 16, (cuuint64_t)G[1] * G[2] * 16};
  cuuint32_t box[3] = {(cuuint32_t)ROW_F, (cuuint32_t)BOX, (cuuint32_t)(L.fbx + 1)};
  cuuint32_t estr[3] = {1, 1, 1};
  if (get_encode()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(feat), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   g_voxel_tma_l2 == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                       : (g_voxel_tma_l2 == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE),
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return NDJIR_ERR_ARG;
  const int smem = (int)(L.fbx + 1) * BOX * ROW_F * 4 + 128;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(brick_gather_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BRICK_BYTES + 128);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(brick_gather_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BRICK_BYTES + 128);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  if (accum) brick_gather_kernel<true><<<L.nfine, 256, smem, st>>>(map, out, rec2, off, g, L, g_voxel_tma_dbg);
  else brick_gather_kernel<false><<<L.nfine, 256, smem, st>>>(map, out, rec2, off, g, L, g_voxel_tma_dbg);
  NDJIR_RETURN_LAST_ERROR();
}

}  // namespace voxel_tma
}  // namespace ndjir
