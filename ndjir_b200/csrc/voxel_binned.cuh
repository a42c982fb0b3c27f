// Brick-ordered voxel gather / scatter (voxel_binned.cu): interface used by the dispatch in voxel.cu.
#pragma once
#include <cuda_runtime.h>
#include "grid_common.cuh"

namespace ndjir {

extern int g_voxel_binned;  // -1 auto (large batch on a table far larger than L2), 0 never, 1 whenever possible
extern int g_voxel_pair256;
extern int g_voxel_bin_mb;  // target brick size in MiB

extern int g_voxel_tma_bx, g_voxel_tma_l2, g_voxel_tma_dbg;
extern int g_voxel_tma;     // 1: fine-brick TMA sweep for the D = 4 forward gather (voxel_tma.cu), 0: L2-window sweep

namespace voxel_binned {

constexpr int kMaxBins = 512;
constexpr long long kHeaderBytes = 8192;  // kMaxBins cursors + padding; records start 16-byte aligned
constexpr long long kTailBytes = (1ll << 21) + 64;   // fine-brick offsets of the TMA sweep (voxel_tma.cu)

struct Bins {
  unsigned px, py;   // planes / rows per brick
  unsigned nby;      // bricks along y
  unsigned n;        // total
};

long long workspace_bytes(long long n_points);
bool shape_ok(long long B, const int* G, int D);
bool worthwhile(long long B, const int* G, int D);
int query(long long B, float* out, const float* query, const float* feat, const int* G, int D, const float* mn,
          const float* mx, bool accum, void* ws, long long ws_bytes, cudaStream_t st);
int scatter(bool second, long long B, float* gf, const float* go, const float* gg, const float* query, const int* G,
            int D, const float* mn, const float* mx, void* ws, long long ws_bytes, cudaStream_t st);
int build_records(long long B, const float* query, const float* payload, const GridFrame& g, const int* G, int D,
                  void* ws, long long ws_bytes, cudaStream_t st, const float4** rec_out);
int build_records_bins(long long B, const float* query, const float* payload, const GridFrame& g, const Bins& b,
                       void* ws, long long ws_bytes, cudaStream_t st, const float4** rec_out);
void* scratch_alloc(long long bytes, cudaStream_t st);
void scratch_free(void* p, cudaStream_t st);

}  // namespace voxel_binned

namespace voxel_tma {
bool eligible(long long B, const int* G, int D, const float* feat, const float* out);
int query(long long B, float* out, const float* query, const float* feat, const GridFrame& g, const int* G, bool accum,
          void* ws, long long ws_bytes, cudaStream_t st);
}  // namespace voxel_tma
}  // namespace ndjir
