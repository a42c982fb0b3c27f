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
/* "scatter_aggregate": 1 = warp-aggregated scatter reductions (coherent rays), 0 = plain vector reductions */
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

#ifdef __cplusplus
}
#endif
#endif /* NDJIR_B200_H */
