/* ndjir_b200 - C ABI of the B200-native (sm_100a) implementation of NDJIR's per-ray rendering hot path.
 *
 * Drop-in boundary: the reference binds its native code as 19 pybind11 modules whose functions take raw
 * device addresses as int64 (reference Makefile:23; e.g. csrc/grid_feature/voxel_feature_cuda.cu:101-115,
 * :844-863).  Every function below replaces one of those exports one-to-one; the pybind-name compatible
 * Python shims live in ndjir_b200/compat/ and INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions (all entry points):
 *   - plain pointers to DEVICE memory, fp32, dense C order, channel-last; caller owns every buffer, the
 *     library allocates nothing, keeps no state between calls and never synchronises;
 *   - `n_points` counts POINTS (rows of `query`); the reference's first argument N counts its threads
 *     (B*D, B*D*3, L*B ...) - the compat shims convert;
 *   - `min3`/`max3` are HOST pointers to 3 floats (the reference passes std::vector<float>);
 *     `grid_sizes` is a HOST pointer to 3 ints;
 *   - `accum` != 0 accumulates into the output, 0 overwrites (zero-filling scatter targets first);
 *     functions without `accum` ALWAYS accumulate, exactly like the reference kernels that ignore the flag;
 *   - the reference's `boundary_check` argument is accepted by its kernels and never read; it is dropped;
 *   - `stream`: kernels are enqueued there (0 = legacy default stream, what nnabla-ext-cuda uses);
 *   - return value: 0 on success, -1 for an invalid argument, otherwise the cudaError_t of the launch
 *     (the reference only printf's launch errors, csrc/cuda_common.cuh:24-32).
 */
#ifndef NDJIR_B200_H
#define NDJIR_B200_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library options ------------------------------------------------------------------------------- */
/* "scatter_aggregate": 1 = warp-aggregated scatter reductions (coherent rays), 0 = plain vector reductions;
 * "mlp_tensor_cores" (default 1; 0 = fp32 FFMA parity path), "mlp_presplit" (1), "mlp_cta_pair" (0), "voxel_binned"
 * (-1 auto / 0 / 1), "voxel_bin_mb" (16), plus profiling switches; unknown keys return -1 */
int ndjir_set_option(const char* key, int value);

/* ---- voxel_feature_cuda (csrc/grid_feature/voxel_feature_cuda.cu:844-863) --------------------------- */
/* query_on_voxel :101 */
int ndjir_voxel_query_on_voxel(long long n_points, float* output, const float* query, const float* feature,
                               const int* grid_sizes, int D, const float* min3, const float* max3, int accum,
                               cudaStream_t stream);
/* grad_query :205 */
int ndjir_voxel_grad_query(long long n_points, float* grad_query, const float* grad_output, const float* query,
                           const float* feature, const int* grid_sizes, int D, const float* min3,
                           const float* max3, int accum, cudaStream_t stream);
/* grad_feature :289 */
int ndjir_voxel_grad_feature(long long n_points, float* grad_feature, const float* grad_output,
                             const float* query, const int* grid_sizes, int D, const float* min3,
                             const float* max3, int accum, cudaStream_t stream);
/* grad_query_grad_grad_output :414 */
int ndjir_voxel_grad_query_grad_grad_output(long long n_points, float* grad_grad_output,
                                            const float* grad_grad_query, const float* query,
                                            const float* feature, const int* grid_sizes, int D,
                                            const float* min3, const float* max3, int accum,
                                            cudaStream_t stream);
/* grad_query_grad_query :523 (always accumulates) */
int ndjir_voxel_grad_query_grad_query(long long n_points, float* grad_query, const float* grad_grad_query,
                                      const float* grad_output, const float* query, const float* feature,
                                      const int* grid_sizes, int D, const float* min3, const float* max3,
                                      cudaStream_t stream);
/* grad_query_grad_feature :616 (always accumulates) */
int ndjir_voxel_grad_query_grad_feature(long long n_points, float* grad_feature, const float* grad_grad_query,
                                        const float* grad_output, const float* query, const int* grid_sizes,
                                        int D, const float* min3, const float* max3, cudaStream_t stream);
/* grad_feature_grad_grad_output :711 */
int ndjir_voxel_grad_feature_grad_grad_output(long long n_points, float* grad_grad_output,
                                              const float* grad_grad_feature, const float* query,
                                              const int* grid_sizes, int D, const float* min3,
                                              const float* max3, int accum, cudaStream_t stream);
/* grad_feature_grad_query :816 (always accumulates) */
int ndjir_voxel_grad_feature_grad_query(long long n_points, float* grad_query, const float* grad_grad_feature,
                                        const float* grad_output, const float* query, const int* grid_sizes,
                                        int D, const float* min3, const float* max3, cudaStream_t stream);

/* Brick-ordered variants for LARGE batches on tables far larger than L2 (no reference counterpart; same results
 * as query_on_voxel :101 / grad_feature :289 / grad_query_grad_feature :616 up to fp32 summation order).  Points
 * are counting-sorted by the ~16 MB table brick of their lower corner so that every brick is pulled into L2 once.
 * `workspace`: DEVICE scratch of ndjir_voxel_binned_workspace_bytes(n_points) bytes, 16-byte aligned, caller-owned.
 * The reference-signature entry points above take this path by themselves (scratch from the stream-ordered CUDA
 * pool, cudaMallocAsync) when n_points >= 2^21 and the table is >= 96 MB; option "voxel_binned" (-1 auto, 0 off,
 * 1 whenever possible) and "voxel_bin_mb" (brick size) control it. */
long long ndjir_voxel_binned_workspace_bytes(long long n_points);
int ndjir_voxel_query_on_voxel_binned(long long n_points, float* output, const float* query, const float* feature,
                                      const int* grid_sizes, int D, const float* min3, const float* max3, int accum,
                                      void* workspace, long long workspace_bytes, cudaStream_t stream);
int ndjir_voxel_grad_feature_binned(long long n_points, float* grad_feature, const float* grad_output,
                                    const float* query, const int* grid_sizes, int D, const float* min3,
                                    const float* max3, int accum, void* workspace, long long workspace_bytes,
                                    cudaStream_t stream);
int ndjir_voxel_grad_query_grad_feature_binned(long long n_points, float* grad_feature, const float* grad_grad_query,
                                               const float* grad_output, const float* query, const int* grid_sizes,
                                               int D, const float* min3, const float* max3, void* workspace,
                                               long long workspace_bytes, cudaStream_t stream);

/* ---- cosine_voxel_feature_cuda (csrc/grid_feature/cosine_voxel_feature_cuda.cu:855-866; 5 exports: the
 *      second-order grad_query_grad_query / grad_feature_* are commented out in the reference).  Same cells and
 *      corner order as voxel_feature_cuda; weights 0.5 cos(pi frac) + 0.5 (:65-66), derivative factor
 *      0.5 pi sin(pi frac) per axis (:163). -------------------------------------------------------------------- */
int ndjir_cosine_voxel_query_on_voxel(long long n_points, float* output, const float* query, const float* feature,
                                      const int* grid_sizes, int D, const float* min3, const float* max3, int accum,
                                      cudaStream_t stream);
int ndjir_cosine_voxel_grad_query(long long n_points, float* grad_query, const float* grad_output, const float* query,
                                  const float* feature, const int* grid_sizes, int D, const float* min3,
                                  const float* max3, int accum, cudaStream_t stream);
int ndjir_cosine_voxel_grad_feature(long long n_points, float* grad_feature, const float* grad_output,
                                    const float* query, const int* grid_sizes, int D, const float* min3,
                                    const float* max3, int accum, cudaStream_t stream);
int ndjir_cosine_voxel_grad_query_grad_grad_output(long long n_points, float* grad_grad_output,
                                                   const float* grad_grad_query, const float* query,
                                                   const float* feature, const int* grid_sizes, int D,
                                                   const float* min3, const float* max3, int accum,
                                                   cudaStream_t stream);
/* always accumulates (:626-650 never reads accum) */
int ndjir_cosine_voxel_grad_query_grad_feature(long long n_points, float* grad_feature, const float* grad_grad_query,
                                               const float* grad_output, const float* query, const int* grid_sizes,
                                               int D, const float* min3, const float* max3, cudaStream_t stream);

/* ---- lanczos_voxel_feature_cuda (csrc/grid_feature/lanczos_voxel_feature_cuda.cu:822-834) ----------- */
int ndjir_lanczos_voxel_query_on_voxel(long long n_points, float* output, const float* query,
                                       const float* feature, const int* grid_sizes, int D, const float* min3,
                                       const float* max3, int accum, cudaStream_t stream);
int ndjir_lanczos_voxel_grad_query(long long n_points, float* grad_query, const float* grad_output,
                                   const float* query, const float* feature, const int* grid_sizes, int D,
                                   const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_lanczos_voxel_grad_feature(long long n_points, float* grad_feature, const float* grad_output,
                                     const float* query, const int* grid_sizes, int D, const float* min3,
                                     const float* max3, int accum, cudaStream_t stream);
int ndjir_lanczos_voxel_grad_query_grad_grad_output(long long n_points, float* grad_grad_output,
                                                    const float* grad_grad_query, const float* query,
                                                    const float* feature, const int* grid_sizes, int D,
                                                    const float* min3, const float* max3, int accum,
                                                    cudaStream_t stream);
int ndjir_lanczos_voxel_grad_query_grad_feature(long long n_points, float* grad_feature,
                                                const float* grad_grad_query, const float* grad_output,
                                                const float* query, const int* grid_sizes, int D,
                                                const float* min3, const float* max3, cudaStream_t stream);

/* ---- voxel_hash_feature_cuda (csrc/grid_feature/voxel_hash_feature_cuda.cu:985-1001) ----------------
 * `layout` 0: values in the reference kernel's (D, L, B) layout (:190); 1: (B, D*L) with c = d*L + l, what
 * the reference's Python wrapper produces after its in-place transpose (voxel_hash_feature.py:152-155). */
long long ndjir_voxel_hash_num_params(int G0, float growth_factor, int T0, int L, int D);
int ndjir_voxel_hash_level_table(int G0, float growth_factor, int T0, int L, int D, int* G_out, int* T_out,
                                 long long* offset_out); /* HOST outputs, L entries each */
/* the table as the DEVICE evaluates it (device pow(float,int) is not exactly rounded: reference quirk q2);
 * G_dev / T_dev are DEVICE int buffers of L entries */
int ndjir_voxel_hash_level_table_device(int G0, float growth_factor, int T0, int L, int D, int* G_dev, int* T_dev,
                                        cudaStream_t stream);
/* hash_index :102 - the 8 hashed corner indices of one level as floats, (B, 8) */
int ndjir_voxel_hash_hash_index(long long n_points, float* output, const float* query, int G, int T,
                                const float* min3, const float* max3, cudaStream_t stream);
int ndjir_voxel_hash_voxel_hash_feature(long long n_points, float* output, const float* query,
                                        const float* feature, int G0, float growth_factor, int T0, int L, int D,
                                        const float* min3, const float* max3, int layout, int accum,
                                        cudaStream_t stream);
int ndjir_voxel_hash_grad_query(long long n_points, float* grad_query, const float* grad_output,
                                const float* query, const float* feature, int G0, float growth_factor, int T0,
                                int L, int D, const float* min3, const float* max3, int layout, int accum,
                                cudaStream_t stream);
int ndjir_voxel_hash_grad_feature(long long n_points, float* grad_feature, const float* grad_output,
                                  const float* query, int G0, float growth_factor, int T0, int L, int D,
                                  const float* min3, const float* max3, int layout, int accum,
                                  cudaStream_t stream);
int ndjir_voxel_hash_grad_query_grad_grad_output(long long n_points, float* grad_grad_output,
                                                 const float* grad_grad_query, const float* query,
                                                 const float* feature, int G0, float growth_factor, int T0,
                                                 int L, int D, const float* min3, const float* max3, int layout,
                                                 int accum, cudaStream_t stream);
int ndjir_voxel_hash_grad_query_grad_feature(long long n_points, float* grad_feature,
                                             const float* grad_grad_query, const float* grad_output,
                                             const float* query, int G0, float growth_factor, int T0, int L,
                                             int D, const float* min3, const float* max3, int layout,
                                             cudaStream_t stream);

/* ---- lanczos_voxel_hash_feature_cuda (csrc/grid_feature/lanczos_voxel_hash_feature_cuda.cu:959-976): the same level
 * table, offsets and layouts with Lanczos-2 windows over 4x4x4 clamped, hashed taps per level.  hash_index (:54-66)
 * hashes ONE integer cell (stored as 3 floats) per row into `output` (n_points floats). */
int ndjir_lanczos_voxel_hash_hash_index(long long n_points, float* output, const float* query, int T,
                                        cudaStream_t stream);
int ndjir_lanczos_voxel_hash_voxel_hash_feature(long long n_points, float* output, const float* query,
                                        const float* feature, int G0, float growth_factor, int T0, int L, int D,
                                        const float* min3, const float* max3, int layout, int accum,
                                        cudaStream_t stream);
int ndjir_lanczos_voxel_hash_grad_query(long long n_points, float* grad_query, const float* grad_output,
                                const float* query, const float* feature, int G0, float growth_factor, int T0,
                                int L, int D, const float* min3, const float* max3, int layout, int accum,
                                cudaStream_t stream);
int ndjir_lanczos_voxel_hash_grad_feature(long long n_points, float* grad_feature, const float* grad_output,
                                  const float* query, int G0, float growth_factor, int T0, int L, int D,
                                  const float* min3, const float* max3, int layout, int accum,
                                  cudaStream_t stream);
int ndjir_lanczos_voxel_hash_grad_query_grad_grad_output(long long n_points, float* grad_grad_output,
                                                 const float* grad_grad_query, const float* query,
                                                 const float* feature, int G0, float growth_factor, int T0,
                                                 int L, int D, const float* min3, const float* max3, int layout,
                                                 int accum, cudaStream_t stream);
int ndjir_lanczos_voxel_hash_grad_query_grad_feature(long long n_points, float* grad_feature,
                                             const float* grad_grad_query, const float* grad_output,
                                             const float* query, int G0, float growth_factor, int T0, int L,
                                             int D, const float* min3, const float* max3, int layout,
                                             cudaStream_t stream);

/* ---- triplane_feature_cuda / triline_feature_cuda (csrc/grid_feature/triplane_feature_cuda.cu:793-805,
 *      triline_feature_cuda.cu:756-768).  feature (3,G,G,D) / (3,G,D); values (B, D*3), c = d*3 + plane. */
int ndjir_triplane_query_on_triplane(long long n_points, float* output, const float* query, const float* feature, int G, int D,
                   const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_triplane_grad_query(long long n_points, float* grad_query, const float* grad_output, const float* query,
                   const float* feature, int G, int D, const float* min3, const float* max3, int accum,
                   cudaStream_t stream);
int ndjir_triplane_grad_feature(long long n_points, float* grad_feature, const float* grad_output, const float* query,
                   int G, int D, const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_triplane_grad_query_grad_grad_output(long long n_points, float* grad_grad_output,
                   const float* grad_grad_query, const float* query, const float* feature, int G, int D,
                   const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_triplane_grad_query_grad_feature(long long n_points, float* grad_feature, const float* grad_grad_query,
                   const float* grad_output, const float* query, int G, int D, const float* min3,
                   const float* max3, cudaStream_t stream);
int ndjir_triline_query_on_triline(long long n_points, float* output, const float* query, const float* feature, int G, int D,
                   const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_triline_grad_query(long long n_points, float* grad_query, const float* grad_output, const float* query,
                   const float* feature, int G, int D, const float* min3, const float* max3, int accum,
                   cudaStream_t stream);
int ndjir_triline_grad_feature(long long n_points, float* grad_feature, const float* grad_output, const float* query,
                   int G, int D, const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_triline_grad_query_grad_grad_output(long long n_points, float* grad_grad_output,
                   const float* grad_grad_query, const float* query, const float* feature, int G, int D,
                   const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_triline_grad_query_grad_feature(long long n_points, float* grad_feature, const float* grad_grad_query,
                   const float* grad_output, const float* query, int G, int D, const float* min3,
                   const float* max3, cudaStream_t stream);
/* cosine_triplane_feature_cuda / cosine_triline_feature_cuda (csrc/grid_feature/cosine_triplane_feature_cuda.cu,
 * cosine_triline_feature_cuda.cu; 5 exports each): same layouts and semantics with the cosine weights. */
int ndjir_cosine_triplane_query_on_triplane(long long n_points, float* output, const float* query, const float* feature, int G, int D,
                   const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_cosine_triplane_grad_query(long long n_points, float* grad_query, const float* grad_output, const float* query,
                   const float* feature, int G, int D, const float* min3, const float* max3, int accum,
                   cudaStream_t stream);
int ndjir_cosine_triplane_grad_feature(long long n_points, float* grad_feature, const float* grad_output, const float* query,
                   int G, int D, const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_cosine_triplane_grad_query_grad_grad_output(long long n_points, float* grad_grad_output,
                   const float* grad_grad_query, const float* query, const float* feature, int G, int D,
                   const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_cosine_triplane_grad_query_grad_feature(long long n_points, float* grad_feature, const float* grad_grad_query,
                   const float* grad_output, const float* query, int G, int D, const float* min3,
                   const float* max3, cudaStream_t stream);
int ndjir_cosine_triline_query_on_triline(long long n_points, float* output, const float* query, const float* feature, int G, int D,
                   const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_cosine_triline_grad_query(long long n_points, float* grad_query, const float* grad_output, const float* query,
                   const float* feature, int G, int D, const float* min3, const float* max3, int accum,
                   cudaStream_t stream);
int ndjir_cosine_triline_grad_feature(long long n_points, float* grad_feature, const float* grad_output, const float* query,
                   int G, int D, const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_cosine_triline_grad_query_grad_grad_output(long long n_points, float* grad_grad_output,
                   const float* grad_grad_query, const float* query, const float* feature, int G, int D,
                   const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_cosine_triline_grad_query_grad_feature(long long n_points, float* grad_feature, const float* grad_grad_query,
                   const float* grad_output, const float* query, int G, int D, const float* min3,
                   const float* max3, cudaStream_t stream);
/* lanczos_triplane_feature_cuda / lanczos_triline_feature_cuda (csrc/grid_feature/lanczos_triplane_feature_cuda.cu:795-808,
 * lanczos_triline_feature_cuda.cu:735-748; 5 exports each): Lanczos-2 windows, 4x4 / 4 clamped taps per plane / line. */
int ndjir_lanczos_triplane_query_on_triplane(long long n_points, float* output, const float* query, const float* feature, int G, int D,
                   const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_lanczos_triplane_grad_query(long long n_points, float* grad_query, const float* grad_output, const float* query,
                   const float* feature, int G, int D, const float* min3, const float* max3, int accum,
                   cudaStream_t stream);
int ndjir_lanczos_triplane_grad_feature(long long n_points, float* grad_feature, const float* grad_output, const float* query,
                   int G, int D, const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_lanczos_triplane_grad_query_grad_grad_output(long long n_points, float* grad_grad_output,
                   const float* grad_grad_query, const float* query, const float* feature, int G, int D,
                   const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_lanczos_triplane_grad_query_grad_feature(long long n_points, float* grad_feature, const float* grad_grad_query,
                   const float* grad_output, const float* query, int G, int D, const float* min3,
                   const float* max3, cudaStream_t stream);
int ndjir_lanczos_triline_query_on_triline(long long n_points, float* output, const float* query, const float* feature, int G, int D,
                   const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_lanczos_triline_grad_query(long long n_points, float* grad_query, const float* grad_output, const float* query,
                   const float* feature, int G, int D, const float* min3, const float* max3, int accum,
                   cudaStream_t stream);
int ndjir_lanczos_triline_grad_feature(long long n_points, float* grad_feature, const float* grad_output, const float* query,
                   int G, int D, const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_lanczos_triline_grad_query_grad_grad_output(long long n_points, float* grad_grad_output,
                   const float* grad_grad_query, const float* query, const float* feature, int G, int D,
                   const float* min3, const float* max3, int accum, cudaStream_t stream);
int ndjir_lanczos_triline_grad_query_grad_feature(long long n_points, float* grad_feature, const float* grad_grad_query,
                   const float* grad_output, const float* query, int G, int D, const float* min3,
                   const float* max3, cudaStream_t stream);

/* ---- total_variation_loss*_cuda (csrc/grid_feature/total_variation_loss_cuda.cu:203-209,
 *      ..._on_triplane_cuda.cu:188-194, ..._on_triline_cuda.cu:180-186); backward always accumulates -------- */
int ndjir_tv_loss_on_voxel(long long n_points, float* output, const float* query, const float* feature,
                           const int* grid_sizes, int D, const float* min3, const float* max3,
                           cudaStream_t stream);
int ndjir_tv_loss_on_voxel_backward(long long n_points, float* grad_feature, const float* grad_output,
                                    const float* query, const float* feature, const int* grid_sizes, int D,
                                    const float* min3, const float* max3, int sym_backward,
                                    cudaStream_t stream);
int ndjir_tv_loss_on_triplane(long long n_points, float* output, const float* query, const float* feature, int G,
                              int D, const float* min3, const float* max3, cudaStream_t stream);
int ndjir_tv_loss_on_triplane_backward(long long n_points, float* grad_feature, const float* grad_output,
                                       const float* query, const float* feature, int G, int D,
                                       const float* min3, const float* max3, int sym_backward,
                                       cudaStream_t stream);
int ndjir_tv_loss_on_triline(long long n_points, float* output, const float* query, const float* feature, int G,
                             int D, const float* min3, const float* max3, cudaStream_t stream);
int ndjir_tv_loss_on_triline_backward(long long n_points, float* grad_feature, const float* grad_output,
                                      const float* query, const float* feature, int G, int D, const float* min3,
                                      const float* max3, int sym_backward, cudaStream_t stream);
/* total_variation_loss_on_voxel_hash_cuda (csrc/grid_feature/total_variation_loss_on_voxel_hash_cuda.cu:229-234): values in
 * the hash family's layouts (0 = (D,L,B), 1 = (B, D*L)); the backward reaches the three upper neighbours only. */
int ndjir_tv_loss_on_voxel_hash(long long n_points, float* output, const float* query, const float* feature, int G0,
                                float growth_factor, int T0, int L, int D, const float* min3, const float* max3,
                                int layout, cudaStream_t stream);
int ndjir_tv_loss_on_voxel_hash_backward(long long n_points, float* grad_feature, const float* grad_output,
                                         const float* query, const float* feature, int G0, float growth_factor,
                                         int T0, int L, int D, const float* min3, const float* max3, int layout,
                                         cudaStream_t stream);

/* ---- ray_aabb_intersection_cuda / ray_sphere_intersection_cuda (csrc/intersection/*.cu:145-169, :81-104);
 *      camloc (B,3), raydir (B,R,3) -> t_near, t_far, n_hits (B,R,1); n_rays = B*R -------------------------- */
int ndjir_ray_aabb_intersection(int n_rays, float* t_near, float* t_far, float* n_hits, const float* camloc,
                                const float* raydir, int B, int R, const float* min3, const float* max3,
                                cudaStream_t stream);
int ndjir_ray_sphere_intersection(int n_rays, float* t_near, float* t_far, float* n_hits, const float* camloc,
                                  const float* raydir, int B, int R, float radius, cudaStream_t stream);

/* ---- inverse_transform_cuda (csrc/sampling/inverse_transform_cuda.cu:72-91, :139-160);
 *      size = batch_size * n_lights, n_lights = n_thes * n_phis; light_dirs (batch, n_lights, 3) ------------ */
int ndjir_sample_uniform_directions(long long size, float* light_dirs, const float* normal, const float* cdf_the,
                                    const float* cdf_phi, int batch_size, int n_lights, int n_thes, int n_phis,
                                    float eps, cudaStream_t stream);
int ndjir_sample_importance_directions(long long size, float* light_dirs, const float* normal,
                                       const float* cdf_the, const float* cdf_phi, const float* alpha,
                                       int batch_size, int n_lights, int n_thes, int n_phis, float eps,
                                       cudaStream_t stream);

/* ---- squareplus_cuda (csrc/activation/squareplus_cuda.cu:63-99) -------------------------------------- */
int ndjir_squareplus_forward(long long size, float* output, const float* input, float b, cudaStream_t stream);
int ndjir_squareplus_backward(long long size, float* dinput, const float* doutput, const float* input, float b,
                              int accum, cudaStream_t stream);

/* ---- measurement helper: random-gather throughput of the memory hierarchy (the yardstick of the L2-resident grid
 *      families: voxel hash, triline, Lanczos voxel; tools/bench_gather.py, profiles/r2_gather_roofline.md).
 *      Every thread sums `per_thread` independent gathers of `elem_bytes` (4, 8 or 16) at counter-hashed element
 *      indices of `table` (table_bytes, a multiple of elem_bytes); sink[0] receives a value that depends on all of
 *      them.  coherent = 1: the 32 lanes of a warp read 32 consecutive elements at a hashed position instead. */
int ndjir_bench_gather(long long n_threads, int per_thread, int elem_bytes, const void* table, long long table_bytes,
                       int coherent, int seed, float* sink, cudaStream_t stream);

/* ==== fused per-ray path ==================================================================================
 * No native counterpart in the reference: there these stages are ~400 stock nnabla ops composed in
 * python/sampler.py, network.py, renderer.py, specular_brdf.py and loss.py, differentiated by nnabla's autodiff.
 * Here every stage (forward AND hand-derived backward) is one kernel; ndjir_b200/engine.py strings them into
 * sample_points / pb_render / total_loss with the reference's signatures.  `ld*` are row strides in floats. */

/* loss terms (python/loss.py:180-192); `losses` buffers hold NDJIR_N_LOSSES floats */
enum { NDJIR_LOSS_TOTAL = 0, NDJIR_LOSS_RGB, NDJIR_LOSS_EIKONAL, NDJIR_LOSS_TV, NDJIR_LOSS_MASK,
       NDJIR_LOSS_PRIOR_BASE_COLOR, NDJIR_LOSS_PRIOR_ROUGHNESS, NDJIR_LOSS_PRIOR_SPECULAR,
       NDJIR_LOSS_REG_STD_ROUGHNESS, NDJIR_LOSS_REG_STD_SPECULAR, NDJIR_N_LOSSES };

/* ---- MLP engine: C = epilogue(A(MxK) * B(KxN)); element strides; epilogues in csrc/gemm.cuh.
 *      Replaces nnabla PF.affine + F.softplus(beta=100) and their backward products (python/network.py:84-93). */
int ndjir_gemm(int M, int N, int K, const float* A, long long a_rs, long long a_cs, const float* B, long long b_rs,
               long long b_cs, float* C, long long ldc, const float* bias, float alpha, float out_scale, float beta,
               const float* H, long long ldh, float hscale, const float* U, long long ldu, float* C2,
               long long ldc2, int split_k, int epilogue, cudaStream_t stream);

/* Weight and bias gradient of one affine layer in one call (the backward of PF.affine, python/network.py:88-93):
 * gW (K_in, N) += A(rows, K_in)^T dZ(rows, N), gb (N) += column sums of dZ (gb may be NULL).  On the tcgen05 path the
 * column sums ride on the product's shared-memory pass over dZ; option "mlp_fused_colsum" = 0 runs them separately. */
int ndjir_wgrad_bias(long long rows, int K_in, int N, const float* A, long long lda, const float* dZ, long long ldz,
                     float* gW, long long ldw, float* gb, int split_k, cudaStream_t stream);

/* ---- sample placement (python/sampler.py) ---- */
/* :140-165  t = t_near + (t_far - t_near)/N0 * (i + xi) */
int ndjir_stratified_dists(int n_rays, int N0, float* t, const float* t_near, const float* t_far, const float* xi,
                           cudaStream_t stream);
/* x[r,i,:] = camloc[r / R] + t[r,i] * raydir[r]   (:193, :276) */
int ndjir_ray_points(int n_rays, int Nt, int R, float* x, const float* camloc, const float* raydir, const float* t,
                     long long ldt, cudaStream_t stream);
/* :196-240 one SDF-guided up-sampling round: Nt sorted distances + sdf -> Nt+M sorted distances */
int ndjir_importance_round(int n_rays, int Nt, int M, const float* t_in, long long ld_in, const float* sdf,
                           long long ld_sdf, const float* t_near, const float* t_far, float gain, float* t_out,
                           long long ld_out, float* t_new_out, int* idx_out, cudaStream_t stream);
/* Incremental form of the same round: (t, sdf) hold Nt sorted pairs per ray (in place, row strides ld_t / ld_sdf);
 * the Mp pending samples of the previous round (dense t_pend (n_rays, Mp) with their freshly evaluated sdf_pend) are
 * merged in first, then M new distances are placed from the merged Nt+Mp pairs and written densely to t_new_out
 * (n_rays, M) for the caller to evaluate.  M == 0 merges only.  Identical results to re-evaluating the SDF at every
 * current sample each round (sampler.py:190-192) at 112 instead of 352 network evaluations per ray. */
int ndjir_importance_round_incremental(int n_rays, int Nt, int Mp, int M, float* t, long long ld_t, float* sdf,
                                       long long ld_sdf, const float* t_pend, const float* sdf_pend,
                                       const float* t_near, const float* t_far, float gain, float* t_new_out,
                                       int* idx_out, cudaStream_t stream);
/* :244-254, :282-291 background inverse-depth samples; t_bg (n_rays, Nb+1), x_bg (n_rays, Nb, 4) */
int ndjir_background_samples(int n_rays, int Nb, int R, const float* camloc, const float* raydir,
                             const float* t_far, const float* mask, const float* xi, float radius, float* t_bg,
                             float* x_bg, cudaStream_t stream);
/* :100 mask = n_hits > 1; mask_sum[0] += sum(mask) */
int ndjir_hit_mask(int n_rays, const float* n_hits, float* mask, float* mask_sum, cudaStream_t stream);

/* ---- optimizer (python/solver.py:29-69 over nnabla S.Adam; python/train.py:135-148) --------------------
 * One fused pass per parameter buffer: g' = g + weight_decay*w; m = b1 m + (1-b1) g'; v = b2 v + (1-b2) g'^2;
 * w -= alpha_t m / (sqrt(v) + eps); g = 0 when zero_grad.  alpha_t = alpha*sqrt(1-b2^t)/(1-b1^t) with t read from
 * the DEVICE counter t_dev (ndjir_adam_tick advances it on steps that are not skipped, like nnabla's update()
 * count); t_dev == NULL: `alpha` is used as alpha_t.
 * skip_flags: DEVICE int[2] or NULL; the update is skipped (gradient still zeroed) when BOTH are non-zero - the
 * reference's `and` between its two solvers (solver.py:67-69). */
int ndjir_adam_tick(int* t_dev, const int* skip_flags, cudaStream_t stream);
int ndjir_adam_step(long long n, float* w, float* g, float* m, float* v, float alpha, float beta1, float beta2,
                    float eps, float weight_decay, const int* t_dev, const int* skip_flags, int zero_grad,
                    cudaStream_t stream);
/* flag[0] |= any(!isfinite(g)); the scan is skipped when only_if != NULL and *only_if == 0 (device) */
int ndjir_nonfinite_flag(long long n, const float* g, int* flag, const int* only_if, cudaStream_t stream);
/* the loop's second guard, `if np.any(np.isnan(loss.d)): continue` (python/train.py:144-146): a NaN among the n loss
 * values raises both skip flags (DEVICE int[2]) so that ndjir_adam_tick / ndjir_adam_step skip the iteration */
int ndjir_nan_loss_flag(int n, const float* loss, int* skip_flags, cudaStream_t stream);
/* g += rate * w: S.Adam.weight_decay as its own pass (solver.py:48-50) */
int ndjir_weight_decay(long long n, float* g, const float* w, float rate, cudaStream_t stream);

/* Pre-split weight operand of the tensor-core products: lo[i] = x[i] - (x[i] with the low 13 mantissa bits cleared).
 * ndjir_gemm_presplit is ndjir_gemm with the lo part of B supplied by the caller (same strides as B; NULL = split in
 * the kernel): B then arrives as two TMA tiles (raw + lo) and the in-kernel hi/lo transform only handles A.  The
 * result is bit-identical.  The CALLER keeps the copy current (ndjir_split_lo after every parameter change).
 * Option "mlp_presplit" (default 1) switches the use of B_lo off. */
int ndjir_split_lo(long long n, float* lo, const float* x, cudaStream_t stream);
int ndjir_gemm_presplit(int M, int N, int K, const float* A, long long a_rs, long long a_cs, const float* B,
                        const float* B_lo, long long b_rs, long long b_cs, float* C, long long ldc, const float* bias,
                        float alpha, float out_scale, float beta, const float* H, long long ldh, float hscale,
                        const float* U, long long ldu, float* C2, long long ldc2, int split_k, int epilogue,
                        cudaStream_t stream);

/* ---- data movement helpers ---- */
int ndjir_copy2d(long long rows, int cols, float* dst, long long ld_dst, const float* src, long long ld_src, int rep,
                 float alpha, int accum, cudaStream_t stream);          /* dst[r,c] (+)= alpha*src[r/rep,c] */
int ndjir_colsum(long long rows, int cols, float* out, const float* src, long long ld_src, float alpha,
                 cudaStream_t stream);                                   /* out[c] += alpha*sum_r src[r,c] */
int ndjir_group_sum(long long n_groups, int group, int cols, float* out, long long ld_out, const float* src,
                    long long ld_src, int accum, cudaStream_t stream);  /* out[g,c] (+)= sum_i src[g*group+i,c] */
int ndjir_fill(long long n, float* p, float value, cudaStream_t stream);
int ndjir_transpose(int rows, int cols, float* dst, long long ld_dst, const float* src, long long ld_src,
                    cudaStream_t stream);                                /* dst[c,r] = src[r,c] */

/* ---- positional encoding [x, cos(b), sin(b)], b[axis*bands+k] = x[axis]*2^k (python/network.py:96-117),
 *      its input gradient (the nn.grad path of renderer.py:52) and the adjoint of that gradient ---- */
int ndjir_positional_encoding(long long rows, int dim, int bands, const float* x, long long ld_x, int rep, float* out,
                              long long ld_out, cudaStream_t stream);
int ndjir_positional_encoding_grad_input(long long rows, int dim, int bands, const float* pe, long long ld_pe,
                                         const float* g, long long ld_g, float* out, long long ld_out, int accum,
                                         cudaStream_t stream);
int ndjir_positional_encoding_grad_input_adjoint(long long rows, int dim, int bands, const float* pe, long long ld_pe,
                                                 const float* nbar, long long ld_n, float* ghat, long long ld_g,
                                                 cudaStream_t stream);

/* ---- NeuS alpha and compositing (python/renderer.py:54-91) ---- */
int ndjir_neus_alpha_forward(long long n_points, int N, float* alpha, const float* sdf, const float* normal,
                             long long ld_n, const float* raydir, const float* t_fg, const float* gain_param,
                             float cos_anneal_ratio, cudaStream_t stream);
int ndjir_neus_alpha_backward(long long n_points, int N, const float* dalpha, const float* sdf, const float* normal,
                              long long ld_n, const float* raydir, const float* t_fg, const float* gain_param,
                              float cos_anneal_ratio, float* dsdf, float* dnormal, long long ld_dn,
                              float* dgain_param, cudaStream_t stream);
int ndjir_bg_alpha_forward(long long n, int Nb, float* alpha, const float* h0, long long ld_h, const float* t_bg,
                           cudaStream_t stream);                         /* python/network.py:544-545 */
int ndjir_bg_alpha_backward(long long n, int Nb, const float* dalpha, const float* h0, long long ld_h,
                            const float* t_bg, float* dh0, long long ld_dh, cudaStream_t stream);
int ndjir_composite_forward(int n_rays, int N, int Nb, const float* alpha_fg, const float* mask,
                            const float* alpha_bg, float* weights, float* trans, cudaStream_t stream);
int ndjir_composite_backward(int n_rays, int N, int Nb, const float* alpha_fg, const float* mask,
                             const float* alpha_bg, const float* trans, const float* dweights, float* dalpha_fg,
                             float* dalpha_bg, cudaStream_t stream);
int ndjir_volume_render_forward(int n_rays, int N, int C, const float* w, long long ld_w, const float* V,
                                long long ld_v, float* out, long long ld_out, cudaStream_t stream);
int ndjir_volume_render_backward(int n_rays, int N, int C, const float* w, long long ld_w, const float* V,
                                 long long ld_v, const float* dpix, long long ld_dpix, float* dV, long long ld_dv,
                                 int accum_dv, float* dw, long long ld_dw, cudaStream_t stream);

/* ---- ONE fused kernel per ray segment for the compositing stage (python/renderer.py:54-91, 179-185; network.py:544-545;
 *      csrc/render_segment.cu): forward = NeuS alpha of the N foreground samples + background alpha of the Nb background
 *      samples + exclusive-cumprod transmittance (chunked warp scan over N + Nb samples) + weights + the volume-rendering
 *      reduction of the C per-sample columns of V + the background colour; backward = weight gradients from both
 *      reductions (V, and the C2 material attributes V2) and the background colour + the division-free suffix warp scan +
 *      alpha backward (dsdf, dnormal += into the normal columns the same call wrote as part of dV, dgain += ) + background
 *      density backward.  dweights / dalpha_fg / dalpha_bg are optional copies of the intermediates (NULL = not stored). */
int ndjir_render_segment_forward(int n_rays, int N, int Nb, int C, const float* sdf, const float* normal, long long ld_n,
                                 const float* raydir, const float* t_fg, const float* gain_param,
                                 float cos_anneal_ratio, const float* mask, const float* bg_h0, long long ld_h,
                                 const float* t_bg, const float* bg_raw, long long ld_raw, const float* V,
                                 long long ld_v, float* alpha_fg, float* alpha_bg, float* weights, float* trans,
                                 float* pix, long long ld_pix, float* colbg, cudaStream_t stream);
int ndjir_render_segment_backward(int n_rays, int N, int Nb, int C, int C2, const float* sdf, const float* normal,
                                  long long ld_n, const float* raydir, const float* t_fg, const float* gain_param,
                                  float cos_anneal_ratio, const float* mask, const float* bg_h0, long long ld_h,
                                  const float* t_bg, const float* bg_raw, long long ld_raw, const float* V,
                                  long long ld_v, const float* V2, long long ld_v2, const float* weights,
                                  const float* trans, const float* dpix, long long ld_dpix, const float* dpix2,
                                  long long ld_dpix2, const float* dcolbg, float* dV, long long ld_dv, float* dV2,
                                  long long ld_dv2, float* d_bg_raw, long long ld_draw, float* d_bg_h0, long long ld_dh,
                                  float* dsdf, float* dnormal, long long ld_dn, float* dgain_param, float* dweights,
                                  float* dalpha_fg, float* dalpha_bg, cudaStream_t stream);

/* ---- per-sample material activations + priors (python/network.py:235-509, python/loss.py:70-176).
 *      raw (P,16): bc 0:3 | ii 3 | ro 4:6 | sp 6:12 | pl 12 | bc_ptb 13:16;  att (P,12): ii | rough | spec(3) |
 *      pl | bc*pl(3) | pad.  cfg10 (HOST): rough_lb, rough_prior, spec_prior, spec_scale, pl_gain, w_eik, w_bc,
 *      w_ro, w_sp, flags (bit 0: base_color_prior_sym_backward; bit 1: diffuse_brdf.entangle is FALSE, att 6:9 then
 *      holds bc instead of bc*pl, renderer.py:166-173; bit 2: implicit illumination switched off, att 0 = 0,
 *      network.py:308-309; bit 3: photogrammetric light switched off, pl = 1, renderer.py:174-176) ---- */
int ndjir_sample_attributes_forward(long long n_points, int N, const float* raw, float* att, const float* normal,
                                    long long ld_n, const float* mask, const float* cfg10, float* losses,
                                    cudaStream_t stream);
int ndjir_sample_attributes_backward(long long n_points, int N, const float* raw, const float* datt,
                                     const float* normal, long long ld_n, const float* mask, const float* cfg10,
                                     const float* inv_denorm, float* draw, float* dnormal, long long ld_dn,
                                     cudaStream_t stream);

/* ---- per-ray shading + colour loss (python/renderer.py:88-180, python/specular_brdf.py:40-118).
 *      cfg5 (HOST): eps_dot, specular weight, 1/(B*R over all ranks), flags (bit 0: diffuse_brdf.entangle, bit 1:
 *      specular_brdf.sampling is uniform: sBRDF = pi D V F, specular_brdf.py:104-108, instead of V F 4 voh / noh;
 *      bit 2: photogrammetric light switched off: colour = VR(bc) + specular, renderer.py:174-176), l2 ---- */
int ndjir_pixel_normal_forward(int n_rays, const float* npix, long long ld, float eps, float* nhat,
                               cudaStream_t stream);
int ndjir_pixel_normal_backward(int n_rays, const float* npix, long long ld, float eps, const float* dnhat,
                                float* dnpix, long long ld_d, int accum, cudaStream_t stream);
int ndjir_shade_forward(int n_rays, int M, const float* nhat, const float* attpix, const float* raydir,
                        const float* dirs_u, const float* dirs_s, const float* el_raw, long long ld_el,
                        const float* sv_raw, long long ld_sv, const float* colbg, const float* color_gt,
                        const float* cfg5, float* color, float* losses, cudaStream_t stream);
int ndjir_shade_backward(int n_rays, int M, const float* nhat, const float* attpix, const float* raydir,
                         const float* dirs_u, const float* dirs_s, const float* el_raw, long long ld_el,
                         const float* sv_raw, long long ld_sv, const float* colbg, const float* color_gt,
                         const float* cfg5, float* d_el_raw, float* d_sv_raw, float* d_attpix, float* d_nhat,
                         float* d_colbg, cudaStream_t stream);
/* the same with a per-ray weight on the colour loss (python/loss.py:63-65: with train.mask_weight > 0 the colour loss
 * covers the object's rays only); ray_weight NULL = the calls above */
int ndjir_shade_forward_weighted(int n_rays, int M, const float* nhat, const float* attpix, const float* raydir,
                                 const float* dirs_u, const float* dirs_s, const float* el_raw, long long ld_el,
                                 const float* sv_raw, long long ld_sv, const float* colbg, const float* color_gt,
                                 const float* cfg5, const float* ray_weight, float* color, float* losses,
                                 cudaStream_t stream);
int ndjir_shade_backward_weighted(int n_rays, int M, const float* nhat, const float* attpix, const float* raydir,
                                  const float* dirs_u, const float* dirs_s, const float* el_raw, long long ld_el,
                                  const float* sv_raw, long long ld_sv, const float* colbg, const float* color_gt,
                                  const float* cfg5, const float* ray_weight, float* d_el_raw, float* d_sv_raw,
                                  float* d_attpix, float* d_nhat, float* d_colbg, cudaStream_t stream);
/* out[r] = obj_mask[r] * n_rays_total / (obj_sum[0] + 1e-5): the weights that turn the 1 / n_rays_total normalisation of
 * the colour loss into loss.py:64's 1 / (sum(obj_mask) + 1e-5) */
int ndjir_ray_loss_weights(int n_rays, const float* obj_mask, const float* obj_sum, float n_rays_total, float* out,
                           cudaStream_t stream);
int ndjir_bg_color_forward(int n_rays, int Nb, const float* w_bg, long long ld_w, const float* raw, long long ld_raw,
                           float* colbg, cudaStream_t stream);
int ndjir_bg_color_backward(int n_rays, int Nb, const float* w_bg, long long ld_w, const float* raw, long long ld_raw,
                            const float* dcolbg, float* dw_bg, long long ld_dw, float* draw, long long ld_draw,
                            cudaStream_t stream);

/* inverse squared camera distance fed to the photogrammetric-light network (python/network.py:396-400) */
int ndjir_inv_sq_dist(long long n_points, long long points_per_view, const float* x, const float* camloc, float* out,
                      long long ld_out, cudaStream_t stream);

/* ---- mask loss (python/loss.py:108-116) on obj_mask_pred = sum_i alpha_fg_i T_i (python/renderer.py:183-185; alpha_fg
 *      there is the UNMASKED foreground alpha while T comes from alpha_fg * mask, so a ray that misses the bounds has
 *      obj_mask_pred = sum_i alpha_fg_i, a ray that hits has the sum of its foreground weights).
 *      p = clip(obj_mask_pred, 1e-3, 1 - 1e-3), term = sum_r BCE(p, obj_mask[r]) / (mask_sum + 1e-5).
 *      acc (n_rays, ld): column `col` holds sum_i w_i (the volume-rendering reduction of a constant-1 attribute column);
 *      alpha_fg (n_rays, N) the stored foreground alphas; mask (n_rays) the hit mask.
 *      losses (may be NULL): [NDJIR_LOSS_MASK] += term, [NDJIR_LOSS_TOTAL] += weight * term.
 *      d_acc / dalpha_missed (both or neither): d_acc[r, col] = weight * d term / d obj_mask_pred for rays that hit (else 0);
 *      dalpha_missed (n_rays, N) = the same factor for every sample of a ray that misses (else 0), to be fed to
 *      ndjir_neus_alpha_backward, which accumulates.  The gradient is 0 where the clip is active. */
int ndjir_mask_loss(int n_rays, int N, const float* acc, long long ld, int col, const float* alpha_fg, const float* mask,
                    const float* obj_mask, const float* mask_sum, float weight, float* losses, float* d_acc,
                    long long ld_d, float* dalpha_missed, cudaStream_t stream);

/* ---- loss plumbing (python/loss.py:59-192) ---- */
int ndjir_ray_mask_fill(long long n_points, int N, int C, float* out, const float* mask, const float* dev_scalar,
                        float scale, cudaStream_t stream);              /* out[p,c] = scale*mask[p/N]*(*dev_scalar) */
int ndjir_masked_sum(long long n_points, int N, int C, const float* v, const float* mask, float* out,
                     cudaStream_t stream);                               /* out[0] += sum v[p,c]*mask[p/N] */
int ndjir_loss_inv_denorm(const float* mask_sum, int N, float* inv_denorm, cudaStream_t stream);
int ndjir_finalize_losses(float* losses, const float* mask_sum, int N, float inv_rays, float w_eik, float w_tv,
                          float w_bc, float w_ro, float w_sp, cudaStream_t stream);

/* ==== full-frame inference driver, device side (python/renderer.py:212-272, python/helper.py:44-81) ==
 * The chunk loop of render_image as device work driven by a DEVICE chunk counter (chunk_dev may be NULL = chunk 0), so
 * that one captured CUDA graph is replayed per chunk without host round trips.  Pixel p = pixel0 + chunk * n_rays + i,
 * x = p % W, y = p / W. */
/* raydir[i] = normalize(R_c2w (K^-1 [x, y, 1])) in float64 like helper.generate_raydir_camloc; kinv9 / rot9: DEVICE
 * row-major 3x3 doubles (inverse intrinsic, camera-to-world rotation) */
int ndjir_generate_rays(int n_rays, int W, long long n_pixels, long long pixel0, const int* chunk_dev,
                        const double* kinv9_dev, const double* rot9_dev, float* raydir, cudaStream_t stream);
/* out[i] = lo + (hi - lo) * u_i, u_i in [0, 1) from a counter-based generator keyed by (seed, *counter_dev, i): the
 * explicit random inputs of sample_points / pb_render (F.rand in the reference) without a host generator */
int ndjir_uniform(long long n, float lo, float hi, long long seed, const int* counter_dev, float* out,
                  cudaStream_t stream);
/* image[p] = clip(color[i], 0, 1) for p < n_pixels (renderer.py:266-268) */
int ndjir_store_chunk(int n_rays, long long n_pixels, long long pixel0, const int* chunk_dev, const float* color,
                      float* image, cudaStream_t stream);
int ndjir_counter_add(int* counter_dev, int step, cudaStream_t stream);
/* One TRAINING batch assembled on the device (python/dataset.py:33-55 IDRDataSource._get_data, default
 * branch, + python/train.py:126-130 -> helper.generate_raydir_camloc): with the dataset resident in HBM (images
 * (n_views, H*W, 3) fp32 in [0, 1], masks (n_views, H*W) fp32 or NULL, per view the float64 inverse intrinsic kinv9 and
 * camera-to-world rotation rot9 and the fp32 camera position) a step's inputs are one launch and no host->device copy.
 *   view_ids   B ints (device): the dataset rows of this step (the reference's rng.permutation order)
 *   pixel_idx  B x R ints (device): flat pixel y*W+x per ray (the reference's rng.randint draw), or NULL: drawn here,
 *              uniform over H*W, from the counter-based generator keyed by (seed, *counter_dev, ray)
 *   outputs    raydir (B,R,3), camloc (B,3), color_gt (B,R,3), obj_mask (B,R,1; may be NULL; 1 without masks),
 *              pixel_out (B x R ints, may be NULL): the pixels used */
int ndjir_train_batch(int B, int R, int W, long long n_pixels, const int* view_ids, const int* pixel_idx, long long seed,
                      const int* counter_dev, const float* images, const float* masks, const double* kinv9,
                      const double* rot9, const float* camloc_all, float* raydir, float* camloc, float* color_gt,
                      float* obj_mask, int* pixel_out, cudaStream_t stream);
/* points of the marching-cubes lattice (python/extract_by_mc.py:47-73: np.linspace(-radius, radius, G) per axis, x the
 * slowest): point i lies on x-plane ix0 + (i / G^2) * ix_stride (the rank stride of a sharded extraction) */
int ndjir_lattice_points(long long n, int G, int ix0, int ix_stride, float radius, float* pts, cudaStream_t stream);

/* ==== split-fp16 MLP engine (csrc/gemm_h.cu, csrc/h16_ops.cu) ==============================================
 * The hidden activations and gradients of the MLPs (nnabla PF.affine + F.softplus chains, python/network.py:84-93,
 * 154-232) are STORED as two fp16 planes, value = (hi + lo) / scale with hi = fp16(x*scale), lo = fp16(x*scale - hi):
 * 22 mantissa bits in the same 4 bytes per element as fp32, and both planes are ready-made tcgen05 kind::f16 operands
 * (TMA-loaded, no in-kernel conversion).  A product is three tensor-core products  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi
 * at the fp16 rate; with `precise` the two correction products of the WHOLE contraction are accumulated first and the
 * main product last, so that the tensor core's truncating fp32 accumulation only sees 16 full-magnitude partial sums
 * per 256-long contraction.  `scale` is a per-tensor power of two kept on the device (ndjir_scale_update derives it from
 * the running max of the tensor's previous use, ndjir_hmat.amax); writers clamp to the fp16 range. */
typedef struct ndjir_hmat {
  void* hi;            /* fp16 plane, rows x ld */
  void* lo;            /* fp16 plane, same shape */
  long long ld;        /* row stride in halfs; a multiple of 8 */
  const float* scale;  /* DEVICE scalar, a power of two; NULL = 1 */
  float* amax;         /* DEVICE scalar: writers fold max|value| into it; NULL = not tracked */
} ndjir_hmat;

typedef struct ndjir_gemm_h_desc {
  int M, N, K;
  int mn_major;        /* 0: A (M x K) and B (N x K) row-major planes (both K-major);
                          1: A (K x M) and B (K x N) row-major planes: C (M x N) += A^T B, the weight gradient */
  int epilogue;        /* csrc/gemm.cuh Epi */
  int precise;         /* 1: corrections first, main product last (forward-accuracy passes) */
  int split_k;         /* mn_major only: K (the sample rows) split over this many work items, atomic epilogue */
  float alpha, out_scale, beta, hscale;
  ndjir_hmat A, B;     /* tensor-core operands */
  const float* A32; long long a_rs, a_cs;   /* fp32 A for the memory-bound corner shapes (K <= 8) */
  const float* B32; long long b_rs, b_cs;   /* fp32 B(k,n) = B32[k*b_rs + n*b_cs] for the corner shapes (N <= 8, K <= 8) */
  float* C; long long ldc;                  /* fp32 output, used when Ch.hi == NULL */
  ndjir_hmat Ch;                            /* split-fp16 output */
  float* C2; long long ldc2; ndjir_hmat C2h;        /* EPI_ADJ second output */
  const float* bias;
  const float* H; long long ldh; ndjir_hmat Hh;     /* activation the sigmoid factor is derived from (either form) */
  const float* U; long long ldu; ndjir_hmat Uh;     /* EPI_MUL_S addend / EPI_ADJ factor (either form) */
  float* colsum;       /* mn_major only: colsum[n] += sum over the K rows of B(k, n): the bias gradient that goes with a
                          weight gradient, carried by spare warps that read the B tiles the product stages anyway */
} ndjir_gemm_h_desc;

int ndjir_gemm_h(const ndjir_gemm_h_desc* d, cudaStream_t stream);
/* dst[r,c] = split(alpha * src[r / rep, c]) for c < cols (fp32 -> split fp16, dst->scale applied, dst->amax updated) */
int ndjir_pack_h(long long rows, int cols, const float* src, long long ld_src, int rep, float alpha,
                 const ndjir_hmat* dst, cudaStream_t stream);
/* dst[r,c] = value of src[r,c] as fp32 */
int ndjir_unpack_h(long long rows, int cols, const ndjir_hmat* src, float* dst, long long ld_dst, cudaStream_t stream);
/* plane copy dst[r,c] = src[r / rep, c]; both must carry the same scale */
int ndjir_copy2d_h(long long rows, int cols, const ndjir_hmat* dst, const ndjir_hmat* src, int rep,
                   cudaStream_t stream);
/* out[c] += alpha * sum_r src[r,c] */
int ndjir_colsum_h(long long rows, int cols, float* out, const ndjir_hmat* src, float alpha, cudaStream_t stream);
/* amax[0] = max(amax[0], max|x|) */
int ndjir_amax(long long n, const float* x, float* amax, cudaStream_t stream);
/* per slot i: if amax[i] is finite and > 0: scales[i] = 2^(target_log2 - 1 - floor(log2(amax[i]))), i.e. the tensor's
 * largest value lands in [2^(target_log2-1), 2^target_log2); amax[i] = 0 afterwards.  flags[0] |= 1 if a slot's
 * amax * old scale exceeded the fp16 range (values were clamped), |= 2 if a scale changed. */
int ndjir_scale_update(int n_slots, float* scales, float* amax, int* flags, int target_log2, cudaStream_t stream);

/* ==== fused path behind the C ABI (last row): host-side sequencing in C++ ======================
 * The reference composes these stages from ~40 nnabla calls per round in Python; here the whole stage is ONE entry
 * point that enqueues the library's own kernels on `stream`, so a host that is not Python can run it.  Everything the
 * call needs is a POD description plus caller-owned device buffers: nothing is allocated, nothing synchronises.
 *
 * ndjir_geo_sdf_forward   python/network.py:154-232 geometric_network (positional encoding + grid features ->
 *                         hidden affine + softplus_100 layers with the skip connection -> sdf column)
 * ndjir_sample_points_fwd python/sampler.py:140-314 SamplePoints._forward_impl / sample_points: ray bounds,
 *                         stratified distances, n_upsamples SDF-guided rounds (incremental: only the new samples of
 *                         a round are evaluated), background samples
 * Split-fp16 engine only (the weights are given as the planes of W^T the tensor-core products read). */
#define NDJIR_MAX_MLP_LAYERS 16
typedef struct ndjir_mlp_layer {
  int K, N;               /* inputs, outputs */
  const float* W;         /* fp32 weights, K rows of N (row stride ldw): read by the <= 8-output products */
  long long ldw;
  const float* bias;      /* N */
  ndjir_hmat Wt;          /* split-fp16 planes of W^T (N rows of K halfs, scale = the weight scale); unused when N <= 8 */
  ndjir_hmat Wp;          /* split-fp16 planes of W itself (K rows of N halfs): the operand of the input-gradient products
                             (ndjir_geo_normal); may be empty for forward-only use */
} ndjir_mlp_layer;

/* One of the reference's head networks (python/network.py:235-561: base colour, implicit illumination, roughness,
 * specular reflectance, photogrammetric / environment light, soft visibility, the two background networks) evaluated
 * as ONE call: n_hidden affine + softplus_100 layers whose outputs are kept in acts[l] (the backward reads them),
 * then n_out output layers (the reference's last layer, split by column group where the engine splits it) with plain
 * affine epilogues, each into fp32 rows (out32[i], row stride ld_out[i]) or, when out32[i] is NULL, into planes outh[i]. */
typedef struct ndjir_mlp_desc {
  int n_hidden, n_out;
  ndjir_mlp_layer hidden[NDJIR_MAX_MLP_LAYERS];
  ndjir_mlp_layer out[4];
  int precise;
} ndjir_mlp_desc;
int ndjir_mlp_forward(const ndjir_mlp_desc* net, long long rows, const ndjir_hmat* x, const ndjir_hmat* acts,
                      float* const* out32, const long long* ld_out, const ndjir_hmat* outh, cudaStream_t stream);

/* Reverse of ndjir_mlp_forward (the backward nnabla derives for python/network.py:235-561 heads): weight and bias
 * gradients ACCUMULATE into g_hidden[l] / g_out[i] (fp32, the layout of W / bias); the gradient with respect to the
 * input, if wanted, goes to dx (columns [0, dx_cols)), overwritten or, with accum_dx, added to.
 *   douts[i]   dL/d(output block i): fp32 rows (d32, row stride ld) or planes (dh) when d32 is NULL
 *   dz[0..1]   two plane pairs of rows x max hidden width (ping-pong of the hidden-layer gradients; their scale
 *              slots are tracked like any gradient tensor)
 * Products issued: per output block one weight-gradient product (+ a column sum for fp32 blocks) and one
 * input-gradient product with the sigmoid factor of the last hidden activation; per hidden layer the same pair. */
typedef struct ndjir_mlp_grad {
  float* gW;              /* K rows of N, row stride = the layer's ldw */
  float* gb;              /* N; may be NULL (no bias gradient) */
} ndjir_mlp_grad;
typedef struct ndjir_mlp_dmat {
  float* d32;             /* fp32 rows ... */
  long long ld;
  ndjir_hmat dh;          /* ... or planes when d32 is NULL */
} ndjir_mlp_dmat;
int ndjir_mlp_backward(const ndjir_mlp_desc* net, const ndjir_mlp_grad* g_hidden, const ndjir_mlp_grad* g_out,
                       long long rows, const ndjir_hmat* x, const ndjir_hmat* acts, const ndjir_mlp_dmat* douts,
                       const ndjir_hmat* dz, const ndjir_mlp_dmat* dx, int dx_cols, int accum_dx, cudaStream_t stream);

typedef struct ndjir_geo_net {
  int n_hidden;                                   /* hidden layers (affine + softplus_100) */
  ndjir_mlp_layer hidden[NDJIR_MAX_MLP_LAYERS];
  ndjir_mlp_layer sdf;                            /* sdf column of the last layer (N = 1) */
  int skip_layer;         /* hidden layer whose INPUT is [a | encoded input] * skip_scale (network.py:171-176); -1 none */
  float skip_scale;       /* 1 / sqrt(2) */
  int pe_bands;           /* encoded input = [x, cos, sin] (3 + 6 pe_bands columns) | grid features */
  int grid_kind;          /* 0 none, 1 voxel (G^3 x D), 2 triplane + triline (3 G^2 D + 3 G D) */
  int grid_size, grid_channels;
  const float* grid0;     /* voxel / triplane table */
  const float* grid1;     /* triline table */
  int precise;            /* accumulation order of the forward products (ndjir_gemm_h_desc.precise) */
  ndjir_mlp_layer feat;   /* feature block of the output layer (N = feature_size); used by ndjir_geo_forward only */
  int use_ste;            /* geometric_network.voxel.use_ste: the normal (ndjir_geo_normal) does not differentiate the
                             grid features with respect to the point (voxel_feature.py:390-391) */
} ndjir_geo_net;

/* caller-owned scratch of one network evaluation over up to `rows` points */
typedef struct ndjir_geo_scratch {
  float* enc;             /* rows x ld_enc fp32: the encoded input */
  long long ld_enc;       /* >= 3 + 6 pe_bands + grid width, multiple of 4 */
  float* grid_tmp;        /* rows x grid width fp32 (unused without a grid) */
  ndjir_hmat ench;        /* planes of the encoded input */
  ndjir_hmat act[2];      /* two activation plane buffers, rows x (widest hidden input), used alternately */
} ndjir_geo_scratch;

int ndjir_geo_sdf_forward(const ndjir_geo_net* net, long long rows, const float* x, float* sdf,
                          const ndjir_geo_scratch* ws, cudaStream_t stream);

/* The geometric network with every layer input kept (python/network.py:154-232 as python/renderer.py:46-52 uses it):
 * acts[0] = planes of the encoded input, acts[l] = input of hidden layer l, acts[n_hidden] = the last activation; sdf
 * (rows) and, when feat32 is given, the feature block (rows x feature_size fp32, row stride ld_feat). */
typedef struct ndjir_geo_store {
  float* enc;             /* rows x ld_enc fp32: the encoded input [PE(x) | grid features | 0] */
  long long ld_enc;
  float* grid_tmp;        /* rows x grid width fp32 scratch (unused without a grid) */
  ndjir_hmat acts[NDJIR_MAX_MLP_LAYERS + 1];
} ndjir_geo_store;
int ndjir_geo_forward(const ndjir_geo_net* net, long long rows, const float* x, float* sdf, float* feat32,
                      long long ld_feat, const ndjir_geo_store* ws, cudaStream_t stream);

/* normal = d sdf / d x (nn.grad([sdf], [x]), python/renderer.py:52) by a reverse sweep over the stored activations:
 * gz[l] = d sdf / d z_l (kept: the adjoint pass of the backward reads them), g_in (rows x ld_enc fp32) = d sdf / d
 * (encoded input), then through the positional encoding and the grid's grad_query.  `ones`: device float(s) = 1.0. */
typedef struct ndjir_geo_normal_ws {
  ndjir_hmat gz[NDJIR_MAX_MLP_LAYERS];
  float* g_in;
  float* grid_tmp;        /* rows x grid width fp32 scratch */
  const float* ones;
} ndjir_geo_normal_ws;
int ndjir_geo_normal(const ndjir_geo_net* net, long long rows, const float* x, const ndjir_geo_store* fwd,
                     const ndjir_geo_normal_ws* ws, float* normal, long long ld_n, cudaStream_t stream);

/* Adjoint of ndjir_geo_normal's sweep (double backward of nn.grad, renderer.py:52): given the seed gh0 = dL/d(g_in)
 * (the adjoint of the encoding's input gradient, which the caller forms with
 * ndjir_positional_encoding_grad_input_adjoint and the grids' *_grad_query_grad_grad_output calls; fp32 rows AND
 * their planes), walks the layers upwards: per layer ONE product with two results (EPI_ADJ: the second-order addend
 * z2[l] to dL/dz_l and the next Ghat), the weight gradient Ghat_l^T gz[l], and at the top the sdf column's gradient.
 *   gz       plane pairs kept by ndjir_geo_normal (ws->gz)
 *   ghat     n_hidden plane pairs (scratch): ghat[l] = input of layer l + 1 of the walk
 *   z2       n_hidden plane pairs (result; ndjir_geo_backward adds them)
 *   ones     one float 1.0f on the device */
int ndjir_geo_normal_adjoint(const ndjir_geo_net* net, const ndjir_mlp_grad* g_hidden, const ndjir_mlp_grad* g_sdf,
                             long long rows, const ndjir_geo_store* fwd, const ndjir_hmat* gz, const float* gh0,
                             long long ld_gh0, const ndjir_hmat* gh0h, const ndjir_hmat* ghat, const ndjir_hmat* z2,
                             const float* ones, cudaStream_t stream);

/* Reverse sweep of the geometric network (the backward of python/network.py:154-232 incl. the second-order terms of the
 * normal): weight / bias gradients accumulate into g_hidden[l], g_sdf, g_feat; the gradient with respect to the grid
 * features of the encoded input is written to dgrid (rows x grid width, row stride ld_dgrid; the caller scatters it
 * into the grid with ndjir_*_grad_feature).
 *   fwd      activations kept by ndjir_geo_forward
 *   dfeat    planes of dL/d(feature output) (rows x feat.N)
 *   dsdf     dL/d(sdf) (rows x 1 fp32) or NULL
 *   z2       NULL, or n_hidden plane pairs: addends to dL/dz_l from the adjoint of ndjir_geo_normal
 *   dz[0..1] two plane pairs of rows x hidden width (ping-pong) */
int ndjir_geo_backward(const ndjir_geo_net* net, const ndjir_mlp_grad* g_hidden, const ndjir_mlp_grad* g_sdf,
                       const ndjir_mlp_grad* g_feat, long long rows, const ndjir_geo_store* fwd, const ndjir_hmat* dfeat,
                       const float* dsdf, const ndjir_hmat* z2, const ndjir_hmat* dz, float* dgrid, long long ld_dgrid,
                       cudaStream_t stream);

/* SDF on the marching-cubes lattice (python/extract_by_mc.py:47-73 compute_pts_vol: linspace(-radius, radius, G)^3, x the
 * slowest axis): `n_planes` x-planes ix0, ix0 + ix_stride, ... (the rank stride of a sharded extraction), evaluated in
 * batches of whole planes of at most batch_points points.  pts: scratch for batch_points x 3 floats; ws: sized for
 * batch_points rows; sdf_out (n_planes, G, G). */
int ndjir_sdf_lattice(const ndjir_geo_net* net, int G, int ix0, int ix_stride, int n_planes, float radius,
                      long long batch_points, float* pts, const ndjir_geo_scratch* ws, float* sdf_out,
                      cudaStream_t stream);

typedef struct ndjir_sampler_config {
  int n_samples0, n_samples1, n_upsamples, n_bg_samples;    /* renderer.n_samples0 / n_samples1 / n_upsamples / n_bg_samples */
  float sampling_sigmoid_gain;                               /* doubled every round (sampler.py:188) */
  int bounds;                                                /* 0: intersect_with_aabb, 1: intersect_with_r_sphere */
  float radius;                                              /* renderer.bounding_sphere_radius */
} ndjir_sampler_config;

/* caller-owned scratch; n = B R rays, N = n_samples0 + n_upsamples n_samples1, Mx = max(n_samples0, n_samples1) */
typedef struct ndjir_sampler_workspace {
  float *t_near, *t_far, *n_hits;   /* n each */
  float* sdf_cur;                   /* n x N: SDF of the current sorted samples, carried from round to round */
  float* t_pend;                    /* n x Mx: the stratified distances */
  float* t_new[2];                  /* n x n_samples1 each: the new distances of a round, alternating */
  float* x;                         /* n Mx x 3: points of the pending samples */
  float* sdf_pend;                  /* n Mx */
  ndjir_geo_scratch geo;            /* sized for n Mx rows */
} ndjir_sampler_workspace;

/* camloc (B,3), raydir (B,R,3), stratified (B,R,n_samples0), background (B,R,n_bg_samples+1): device fp32.
 * Outputs: x_fg (n,N,3), t_fg (n,N+1) [= sorted distances | t_far], x_bg (n,Nb,4), t_bg (n,Nb+1), mask (n);
 * mask_sum[0] += sum(mask) when mask_sum is given. */
int ndjir_sample_points_fwd(const ndjir_sampler_config* cfg, const ndjir_geo_net* net, int B, int R, const float* camloc,
                            const float* raydir, const float* stratified, const float* background,
                            const ndjir_sampler_workspace* ws, float* x_fg, float* t_fg, float* x_bg, float* t_bg,
                            float* mask, float* mask_sum, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NDJIR_B200_H */
