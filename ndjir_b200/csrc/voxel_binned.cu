// Brick-ordered (binned) trilinear voxel gather / scatter for LARGE query batches on tables far larger than L2.
//
// Same arithmetic as voxel.cu (reference csrc/grid_feature/voxel_feature_cuda.cu:34-115 forward, :231-311
// grad_feature, :549-637 grad_query_grad_feature); what changes is the ORDER in which points are visited.
// A uniformly random point costs 8 cells x 16 B of table, but DRAM moves 64-byte atoms and nothing is reused:
// the direct kernels move 580 B/point for 156 algorithmic bytes (profiles/r1_voxel_gather_ncu_full.csv).  Here the
// table is cut into bricks of ~16 MB (slabs of whole x-planes, or y-strips of one plane when a plane is larger),
// points are counting-sorted by the brick of their lower corner, and the gather / scatter sweeps the bricks in
// order with every resident CTA working in the same narrow window of the sorted list, so that each brick is pulled
// into the 126 MB L2 once and all of its cells are served (or atomically updated) from there.
//
//   1. bin_count   : per-CTA shared-memory histogram of brick ids -> global counts            (12 B/pt read)
//   2. bin_scan    : one CTA, exclusive scan of <= 1024 counts -> cursors
//   3. bin_place   : per-CTA chunk of 4096 points reserves one contiguous run per brick (one global atomic per
//                    CTA and brick) and writes 16-byte records {qx, qy, qz, point index}      (12 B read, 16 B write)
//   4. gather / scatter over the records in brick order; results go to row `point index`.
//
// Workspace: 16 B per point + 8 KB (ndjir_voxel_binned_workspace_bytes).  The order inside a brick is not
// deterministic (atomics), which does not matter: gathers are per-point independent and scatters are atomic sums.
#include <mutex>
#include "grid_common.cuh"
#include "voxel_binned.cuh"
#include "../../include/ndjir_b200.h"

namespace ndjir {

int g_voxel_binned = -1;    // -1 auto, 0 never, 1 always (when the shape allows it)
int g_voxel_bin_mb = 16;    // target brick size
int g_voxel_pair256 = 0;    // 256-bit loads for z-neighbour pairs that share a sector

namespace voxel_binned {

constexpr int kPlaceBlock = 256;

static Bins make_bins(const int* G, int D, int target_mb) {
  Bins b;
  long long target = (long long)(target_mb > 0 ? target_mb : 16) << 20;
  long long row = (long long)G[2] * D * 4, plane = row * G[1];
  b.px = 1; b.py = (unsigned)G[1];
  if (plane <= target) {
    b.px = (unsigned)(target / plane);
    if (b.px < 1) b.px = 1;
  } else {
    long long py = target / row;
    b.py = (unsigned)(py < 1 ? 1 : py);
  }
  for (;;) {
    unsigned nbx = ((unsigned)G[0] + b.px - 1) / b.px;
    b.nby = ((unsigned)G[1] + b.py - 1) / b.py;
    b.n = nbx * b.nby;
    if (b.n <= (unsigned)kMaxBins) break;
    if (b.py < (unsigned)G[1]) b.py *= 2; else b.px *= 2;
  }
  return b;
}

__device__ __forceinline__ unsigned bin_of(const GridFrame& g, const Bins& b, float qx, float qy) {
  unsigned x0, x1, y0, y1;
  float w0, w1;
  cell_axis(qx, g.mnx, g.sx, g.gx1, x0, x1, w0, w1);
  cell_axis(qy, g.mny, g.sy, g.gy1, y0, y1, w0, w1);
  return (x0 / b.px) * b.nby + (y0 / b.py);
}

__global__ void __launch_bounds__(256)
bin_count_kernel(long long B, const float* __restrict__ query, GridFrame g, Bins b, unsigned* __restrict__ counts) {
  __shared__ unsigned hist[kMaxBins];
  for (int i = threadIdx.x; i < (int)b.n; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < B; p += stride) {
    const float* q = query + p * 3;
    atomicAdd(&hist[bin_of(g, b, __ldg(q), __ldg(q + 1))], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (int)b.n; i += blockDim.x)
    if (hist[i]) atomicAdd(&counts[i], hist[i]);
}

// counts -> exclusive prefix (in place); one CTA of kMaxBins threads.
__global__ void __launch_bounds__(kMaxBins)
bin_scan_kernel(unsigned* __restrict__ counts, int n) {
  __shared__ unsigned s[kMaxBins];
  int t = threadIdx.x;
  unsigned v = t < n ? counts[t] : 0u;
  s[t] = v;
  __syncthreads();
  for (int o = 1; o < kMaxBins; o <<= 1) {
    unsigned add = t >= o ? s[t - o] : 0u;
    __syncthreads();
    s[t] += add;
    __syncthreads();
  }
  if (t < n) counts[t] = s[t] - v;
}

// Exclusive scan of n <= kMaxBins shared-memory counters by one CTA of kPlaceBlock threads (4 counters per thread).
__device__ __forceinline__ void block_exclusive_scan(const unsigned* __restrict__ in, unsigned* __restrict__ out,
                                                     int n, unsigned* warp_tot) {
  constexpr int per = kMaxBins / kPlaceBlock;
  int t = threadIdx.x, lane = t & 31, w = t >> 5;
  unsigned v[per], sum = 0;
#pragma unroll
  for (int k = 0; k < per; ++k) { int i = t * per + k; v[k] = i < n ? in[i] : 0u; sum += v[k]; }
  unsigned inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { unsigned u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  unsigned before = 0;
  for (int k = 0; k < w; ++k) before += warp_tot[k];
  unsigned run = before + inc - sum;
#pragma unroll
  for (int k = 0; k < per; ++k) { int i = t * per + k; if (i < n) out[i] = run; run += v[k]; }
  __syncthreads();
}

// Records are staged in shared memory in brick order within the CTA's chunk, then written out so that consecutive
// threads write consecutive 16/32-byte records of one run (full sectors) instead of 32 scattered partial sectors per
// store instruction.  WIDE: 32-byte records {qx,qy,qz,index | payload(4)} carrying the point's grad_output row
// (D = 4), so that the scatter sweep needs no random 16-byte read per point (a 128-byte DRAM fetch each).
template <bool WIDE>
__global__ void __launch_bounds__(kPlaceBlock)
bin_place_kernel(long long B, const float* __restrict__ query, const float* __restrict__ payload, GridFrame g, Bins b,
                 unsigned* __restrict__ cursors, float4* __restrict__ rec) {
  constexpr int kPlacePerThread = WIDE ? 4 : 8;   // 32 KB of staged records per chunk either way
  constexpr int kChunk = kPlaceBlock * kPlacePerThread;
  __shared__ unsigned hist[kMaxBins];    // counts, then running local cursors
  __shared__ unsigned loff[kMaxBins];    // chunk-local exclusive offsets
  __shared__ unsigned base[kMaxBins];    // global position of the chunk's run minus loff
  __shared__ unsigned warp_tot[kPlaceBlock / 32];
  __shared__ unsigned short sbin[kChunk];
  __shared__ float4 stage[kChunk * (WIDE ? 2 : 1)];
  long long n_chunks = (B + kChunk - 1) / kChunk;
  for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    for (int i = threadIdx.x; i < (int)b.n; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    long long p0 = ch * kChunk + threadIdx.x;
    float qx[kPlacePerThread], qy[kPlacePerThread], qz[kPlacePerThread];
    unsigned short bins[kPlacePerThread];
#pragma unroll
    for (int k = 0; k < kPlacePerThread; ++k) {
      long long p = p0 + (long long)k * kPlaceBlock;
      unsigned bi = 0;
      qx[k] = qy[k] = qz[k] = 0.f;
      if (p < B) {
        const float* q = query + p * 3;
        qx[k] = __ldg(q); qy[k] = __ldg(q + 1); qz[k] = __ldg(q + 2);
        bi = bin_of(g, b, qx[k], qy[k]);
        atomicAdd(&hist[bi], 1u);
      }
      bins[k] = (unsigned short)bi;
    }
    __syncthreads();
    block_exclusive_scan(hist, loff, (int)b.n, warp_tot);
    for (int i = threadIdx.x; i < (int)b.n; i += blockDim.x) {
      unsigned c = hist[i];
      base[i] = (c ? atomicAdd(&cursors[i], c) : 0u) - loff[i];
      hist[i] = loff[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPlacePerThread; ++k) {
      long long p = p0 + (long long)k * kPlaceBlock;
      if (p < B) {
        unsigned bi = bins[k];
        unsigned slot = atomicAdd(&hist[bi], 1u);
        sbin[slot] = (unsigned short)bi;
        float4 r = make_float4(qx[k], qy[k], qz[k], __uint_as_float((unsigned)p));
        if (WIDE) {
          stage[2 * slot] = r;
          stage[2 * slot + 1] = __ldg(reinterpret_cast<const float4*>(payload) + p);
        } else {
          stage[slot] = r;
        }
      }
    }
    __syncthreads();
    long long left = B - ch * kChunk;
    int cnt = (int)(left < kChunk ? left : kChunk);
    if (WIDE) {
      for (int j = threadIdx.x; j < 2 * cnt; j += blockDim.x) {
        int slot = j >> 1;
        rec[2 * (long long)(base[sbin[slot]] + slot) + (j & 1)] = stage[j];
      }
    } else {
      for (int j = threadIdx.x; j < cnt; j += blockDim.x) rec[base[sbin[j]] + j] = stage[j];
    }
    __syncthreads();
  }
}

struct Strides { unsigned sx, sy, sz; };

// Forward gather over the records, FOUR LANES PER POINT (lane (cx,cy) fetches the two z-neighbour cells of its
// column), like voxel::gather4_kernel; expression shape per channel as voxel_feature_cuda.cu:87-94 up to the
// association of the 8-term sum (pairs per column, then a 4-lane butterfly).
// One 32-byte request for a column whose two z-neighbour cells share a sector (D = 4, z0 even): the sweep is bound
// by L2 request throughput, not DRAM (ncu: 11.5 sector reads per point, 10 TB/s through the L2 slices).
__device__ __forceinline__ void ldg_pair256(const float* p, Vec<4>& a, Vec<4>& b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.v[0]), "=f"(a.v[1]), "=f"(a.v[2]), "=f"(a.v[3]), "=f"(b.v[0]), "=f"(b.v[1]), "=f"(b.v[2]),
                 "=f"(b.v[3])
               : "l"(p));
}

template <int V, bool ACCUM, bool PAIR256>
__global__ void __launch_bounds__(256, 6)
gather_kernel(long long B, float* __restrict__ out, const float4* __restrict__ rec, const float* __restrict__ feat,
              GridFrame g, Strides s, int D) {
  const int sub = threadIdx.x & 3;
  const int cx = sub >> 1, cy = sub & 1;
  // NO grid-stride loop: CTA k owns records [128k, 128k+128), two per lane group (both issued before either is
  // used: the sweep is latency-bound, not DRAM-bound).  The hardware hands out CTAs in blockIdx order as slots free
  // up, so the set of records in flight is a sliding window of (resident CTAs x 128) ~ 114k records = 14 MB of
  // table; a persistent grid-stride sweep lets fast CTAs run many rounds ahead and the window (and with it the
  // L2 footprint) grows without bound (measured: 8.0 GB of DRAM reads instead of 2.9 GB).
  long long i0 = (long long)blockIdx.x * 128 + (threadIdx.x >> 2);
  bool active[2];
  float4 rc[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    long long i = i0 + 64 * u;
    active[u] = i < B;
    rc[u] = __ldg(rec + (active[u] ? i : B - 1));
  }
  for (int d = 0; d < D; d += V) {
    Vec<V> f0[2], f1[2];
    float w0[2], w1[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      Cell c = make_cell_linear(g, rc[u].x, rc[u].y, rc[u].z);
      unsigned base = (cx ? c.x1 : c.x0) * s.sx + (cy ? c.y1 : c.y0) * s.sy;
      float wxy = (cx ? c.p1 : c.p0) * (cy ? c.q1 : c.q0);
      w0[u] = wxy * c.r0; w1[u] = wxy * c.r1;
      if (PAIR256 && V == 4 && (c.z0 & 1u) == 0u && c.z1 == c.z0 + 1u) {
        ldg_pair256(feat + base + c.z0 * s.sz, *reinterpret_cast<Vec<4>*>(&f0[u]), *reinterpret_cast<Vec<4>*>(&f1[u]));
      } else {
        f0[u] = ldg_vec<V>(feat + base + c.z0 * s.sz + d);
        f1[u] = ldg_vec<V>(feat + base + c.z1 * s.sz + d);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      Vec<V> o;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        o.v[j] = w0[u] * f0[u].v[j] + w1[u] * f1[u].v[j];
        o.v[j] += __shfl_xor_sync(0xffffffffu, o.v[j], 1);
        o.v[j] += __shfl_xor_sync(0xffffffffu, o.v[j], 2);
      }
      if (active[u] && sub == 0) {
        float* op = out + (long long)__float_as_uint(rc[u].w) * D + d;
        if (ACCUM) {
          Vec<V> pv = ld_vec<V>(op);
#pragma unroll
          for (int j = 0; j < V; ++j) o.v[j] += pv.v[j];
        }
        st_vec<V>(op, o);
      }
    }
  }
}

// Scatter over the records, EIGHT LANES PER POINT (one corner each), like voxel::scatter8_kernel.
// WIDE (first order, D = 4): the record carries the grad_output row, nothing is read at `point index`.
template <bool SECOND, int V, bool WIDE>
__global__ void __launch_bounds__(256, 6)
scatter_kernel(long long B, float* __restrict__ gf, const float* __restrict__ go, const float* __restrict__ gg,
               const float4* __restrict__ rec, GridFrame g, Strides s, int D) {
  const int k = threadIdx.x & 7;
  const int cx = (k >> 2) & 1, cy = (k >> 1) & 1, cz = k & 1;
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 8;  // one pass per CTA (see gather_kernel)
  if (i < B) {
    float4 rc = __ldg(rec + (WIDE ? 2 * i : i));
    float4 pay = WIDE ? __ldg(rec + 2 * i + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    long long p = (long long)__float_as_uint(rc.w);
    Cell c = make_cell_linear(g, rc.x, rc.y, rc.z);
    float pw = cx ? c.p1 : c.p0, qw = cy ? c.q1 : c.q0, rw = cz ? c.r1 : c.r0;
    float coef;
    if (!SECOND) {
      coef = pw * qw * rw;
    } else {
      float ggx = __ldg(gg + p * 3) * c.sx, ggy = __ldg(gg + p * 3 + 1) * c.sy, ggz = __ldg(gg + p * 3 + 2) * c.sz;
      coef = ggx * ((cx ? 1.f : -1.f) * qw * rw) + ggy * ((cy ? 1.f : -1.f) * pw * rw) +
             ggz * ((cz ? 1.f : -1.f) * pw * qw);
    }
    float* dst = gf + ((cx ? c.x1 : c.x0) * s.sx + (cy ? c.y1 : c.y0) * s.sy + (cz ? c.z1 : c.z0) * s.sz);
    if (WIDE) {
      Vec<4> o;
      o.v[0] = pay.x * coef; o.v[1] = pay.y * coef; o.v[2] = pay.z * coef; o.v[3] = pay.w * coef;
      red_vec<4>(dst, o);
      return;
    }
    for (int d = 0; d < D; d += V) {
      Vec<V> o = ldg_vec<V>(go + p * D + d);
#pragma unroll
      for (int j = 0; j < V; ++j) o.v[j] *= coef;
      red_vec<V>(dst + d, o);
    }
  }
}

long long workspace_bytes(long long n_points) {
  if (n_points < 0) return -1;
  return kHeaderBytes + 32 * n_points + kTailBytes;     // cursors | two record buffers | fine-brick offsets
}

bool shape_ok(long long B, const int* G, int D) {
  if (B <= 0 || B >= (1ll << 32)) return false;
  return (long long)G[0] * G[1] * G[2] * D < (1ll << 32);
}

bool worthwhile(long long B, const int* G, int D) {
  if (g_voxel_binned == 0 || !shape_ok(B, G, D)) return false;
  if (g_voxel_binned > 0) return true;
  long long table = (long long)G[0] * G[1] * G[2] * D * 4;
  return table >= (96ll << 20) && B >= (1ll << 21);
}

// Builds the brick-ordered records in `ws`: 16-byte {q, point index}, or 32-byte {q, index | payload row (4 floats)}
// when `payload` is given.  Also used by the Lanczos voxel family (lanczos_voxel.cu).
int build_records(long long B, const float* query, const float* payload, const GridFrame& g, const int* G, int D,
                  void* ws, long long ws_bytes, cudaStream_t st, const float4** rec_out) {
  return build_records_bins(B, query, payload, g, make_bins(G, D, g_voxel_bin_mb), ws, ws_bytes, st, rec_out);
}

int build_records_bins(long long B, const float* query, const float* payload, const GridFrame& g, const Bins& b,
                       void* ws, long long ws_bytes, cudaStream_t st, const float4** rec_out) {
  if (!ws || ws_bytes < workspace_bytes(B) || (reinterpret_cast<uintptr_t>(ws) & 15)) return NDJIR_ERR_ARG;
  if (b.n > (unsigned)kMaxBins) return NDJIR_ERR_ARG;
  unsigned* cursors = reinterpret_cast<unsigned*>(ws);
  float4* rec = reinterpret_cast<float4*>(reinterpret_cast<char*>(ws) + kHeaderBytes);
  cudaError_t e = cudaMemsetAsync(cursors, 0, kMaxBins * sizeof(unsigned), st);
  if (e != cudaSuccess) return (int)e;
  int grid = grid_for(B, 256, 8);
  bin_count_kernel<<<grid, 256, 0, st>>>(B, query, g, b, cursors);
  bin_scan_kernel<<<1, kMaxBins, 0, st>>>(cursors, (int)b.n);
  long long chunk = kPlaceBlock * (payload ? 4 : 8);
  long long n_chunks = (B + chunk - 1) / chunk;
  long long cap = (long long)NDJIR_NUM_SMS * 16;
  int pgrid = (int)(n_chunks < cap ? n_chunks : cap);
  if (payload) bin_place_kernel<true><<<pgrid, kPlaceBlock, 0, st>>>(B, query, payload, g, b, cursors, rec);
  else bin_place_kernel<false><<<pgrid, kPlaceBlock, 0, st>>>(B, query, nullptr, g, b, cursors, rec);
  *rec_out = rec;
  return NDJIR_OK;
}

static Strides make_strides(const int* G, int D) {
  Strides s;
  s.sx = (unsigned)G[1] * (unsigned)G[2] * (unsigned)D;
  s.sy = (unsigned)G[2] * (unsigned)D;
  s.sz = (unsigned)D;
  return s;
}

// One CTA per 256/lanes records, in record order (see gather_kernel).
static long long sweep_grid(long long items, int lanes) {
  long long need = (items * lanes + 255) / 256;
  return need < 1 ? 1 : need;
}

int query(long long B, float* out, const float* query_, const float* feat, const int* G, int D, const float* mn,
          const float* mx, bool accum, void* ws, long long ws_bytes, cudaStream_t st) {
  if (B == 0) return NDJIR_OK;
  if (!shape_ok(B, G, D) || !out || !query_ || !feat) return NDJIR_ERR_ARG;
  GridFrame g = make_frame(G[0], G[1], G[2], mn, mx);
  if (voxel_tma::eligible(B, G, D, feat, out)) return voxel_tma::query(B, out, query_, feat, g, G, accum, ws, ws_bytes, st);
  const float4* rec = nullptr;
  int rc = build_records(B, query_, nullptr, g, G, D, ws, ws_bytes, st, &rec);
  if (rc != NDJIR_OK) return rc;
  Strides s = make_strides(G, D);
  int V = pick_vec(D, feat, out);
  unsigned grid = (unsigned)((B + 127) / 128);
#define NDJIR_LAUNCH(VV)                                                                        \
  if (accum) gather_kernel<VV, true, false><<<grid, 256, 0, st>>>(B, out, rec, feat, g, s, D);  \
  else gather_kernel<VV, false, false><<<grid, 256, 0, st>>>(B, out, rec, feat, g, s, D);
  bool pair256 = V == 4 && D == 4 && (G[2] & 1) == 0 && (reinterpret_cast<uintptr_t>(feat) & 31) == 0 && g_voxel_pair256;
  if (pair256) {
    if (accum) gather_kernel<4, true, true><<<grid, 256, 0, st>>>(B, out, rec, feat, g, s, D);
    else gather_kernel<4, false, true><<<grid, 256, 0, st>>>(B, out, rec, feat, g, s, D);
  } else if (V == 4) { NDJIR_LAUNCH(4) } else if (V == 2) { NDJIR_LAUNCH(2) } else { NDJIR_LAUNCH(1) }
#undef NDJIR_LAUNCH
  NDJIR_RETURN_LAST_ERROR();
}

int scatter(bool second, long long B, float* gf, const float* go, const float* gg, const float* query_,
            const int* G, int D, const float* mn, const float* mx, void* ws, long long ws_bytes, cudaStream_t st) {
  if (B == 0) return NDJIR_OK;
  if (!shape_ok(B, G, D) || !gf || !go || !query_ || (second && !gg)) return NDJIR_ERR_ARG;
  GridFrame g = make_frame(G[0], G[1], G[2], mn, mx);
  const float4* rec = nullptr;
  int V = pick_vec(D, gf, go);
  bool wide = !second && D == 4 && V == 4;
  int rc = build_records(B, query_, wide ? go : nullptr, g, G, D, ws, ws_bytes, st, &rec);
  if (rc != NDJIR_OK) return rc;
  Strides s = make_strides(G, D);
  unsigned grid = (unsigned)sweep_grid(B, 8);
  if (wide) {
    scatter_kernel<false, 4, true><<<grid, 256, 0, st>>>(B, gf, go, gg, rec, g, s, D);
    NDJIR_RETURN_LAST_ERROR();
  }
#define NDJIR_LAUNCH(VV)                                                                          \
  if (second) scatter_kernel<true, VV, false><<<grid, 256, 0, st>>>(B, gf, go, gg, rec, g, s, D); \
  else scatter_kernel<false, VV, false><<<grid, 256, 0, st>>>(B, gf, go, gg, rec, g, s, D);
  if (V == 4) { NDJIR_LAUNCH(4) } else if (V == 2) { NDJIR_LAUNCH(2) } else { NDJIR_LAUNCH(1) }
#undef NDJIR_LAUNCH
  NDJIR_RETURN_LAST_ERROR();
}

// Stream-ordered scratch for the reference-signature entry points (no host synchronisation).  The blocks come from a
// memory pool this library owns, one per device, created on first use: its release threshold keeps freed scratch for the
// next call, and the host application's default pool (and PyTorch's allocator behind it) is left exactly as it was.
// Returns nullptr when the allocation is refused - the caller then runs the direct kernel.
void* scratch_alloc(long long bytes, cudaStream_t st) {
  constexpr int MAX_DEV = 64;
  static cudaMemPool_t pools[MAX_DEV] = {};
  static std::mutex mu;
  int devi = 0;
  if (cudaGetDevice(&devi) != cudaSuccess || devi < 0 || devi >= MAX_DEV) { (void)cudaGetLastError(); return nullptr; }
  cudaMemPool_t pool;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!pools[devi]) {
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = devi;
      cudaMemPool_t created;
      if (cudaMemPoolCreate(&created, &props) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(created, cudaMemPoolAttrReleaseThreshold, &keep);
      (void)cudaGetLastError();
      pools[devi] = created;
    }
    pool = pools[devi];
  }
  void* p = nullptr;
  cudaError_t e = cudaMallocFromPoolAsync(&p, (size_t)bytes, pool, st);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
  return p;
}

void scratch_free(void* p, cudaStream_t st) {
  if (p) cudaFreeAsync(p, st);
}

}  // namespace voxel_binned
}  // namespace ndjir

using namespace ndjir;

extern "C" {

long long ndjir_voxel_binned_workspace_bytes(long long n_points) { return voxel_binned::workspace_bytes(n_points); }

int ndjir_voxel_query_on_voxel_binned(long long n_points, float* output, const float* query, const float* feature,
                                      const int* grid_sizes, int D, const float* min3, const float* max3, int accum,
                                      void* workspace, long long workspace_bytes, cudaStream_t stream) {
  if (!grid_sizes || D <= 0 || !min3 || !max3) return NDJIR_ERR_ARG;
  return voxel_binned::query(n_points, output, query, feature, grid_sizes, D, min3, max3, accum != 0, workspace,
                             workspace_bytes, stream);
}

int ndjir_voxel_grad_feature_binned(long long n_points, float* grad_feature, const float* grad_output,
                                    const float* query, const int* grid_sizes, int D, const float* min3,
                                    const float* max3, int accum, void* workspace, long long workspace_bytes,
                                    cudaStream_t stream) {
  if (!grid_sizes || D <= 0 || !min3 || !max3 || !grad_feature) return NDJIR_ERR_ARG;
  if (!voxel_binned::shape_ok(n_points > 0 ? n_points : 1, grid_sizes, D)) return NDJIR_ERR_ARG;
  if (!accum) fill_zero(grad_feature, (long long)grid_sizes[0] * grid_sizes[1] * grid_sizes[2] * D, stream);
  return voxel_binned::scatter(false, n_points, grad_feature, grad_output, nullptr, query, grid_sizes, D, min3, max3,
                               workspace, workspace_bytes, stream);
}

int ndjir_voxel_grad_query_grad_feature_binned(long long n_points, float* grad_feature, const float* grad_grad_query,
                                               const float* grad_output, const float* query, const int* grid_sizes,
                                               int D, const float* min3, const float* max3, void* workspace,
                                               long long workspace_bytes, cudaStream_t stream) {
  if (!grid_sizes || D <= 0 || !min3 || !max3 || !grad_feature) return NDJIR_ERR_ARG;
  return voxel_binned::scatter(true, n_points, grad_feature, grad_output, grad_grad_query, query, grid_sizes, D, min3,
                               max3, workspace, workspace_bytes, stream);
}

}  // extern "C"
