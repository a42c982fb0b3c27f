// Host-side sequencing of the fused per-ray path behind the C ABI: the stages the reference
// composes from nnabla calls in Python (python/sampler.py:140-314, python/network.py:154-232) as single entry points
// that enqueue this library's kernels.  No allocation, no synchronisation, no Python: a C / C++ host can run them.
#include "common.cuh"
#include "gemm.cuh"
#include "gemm_h.cuh"
#include "../../include/ndjir_b200.h"

namespace {

#define NDJIR_TRY(call)                 \
  do {                                  \
    int rc_ = (call);                   \
    if (rc_ != NDJIR_OK) return rc_;    \
  } while (0)

inline ndjir_hmat view(const ndjir_hmat& m, long long col, bool track) {
  ndjir_hmat v = m;
  v.hi = reinterpret_cast<char*>(m.hi) + 2 * col;
  v.lo = reinterpret_cast<char*>(m.lo) + 2 * col;
  if (!track) v.amax = nullptr;
  return v;
}

// fp32 columns -> planes at column dcol of dst
int pack_cols(long long rows, int ncols, const float* src, long long ld_src, float alpha, const ndjir_hmat& dst,
              long long dcol, cudaStream_t st) {
  ndjir_hmat d = view(dst, dcol, true);
  return ndjir_pack_h(rows, ncols, src, ld_src, 1, alpha, &d, st);
}

int grid_width(const ndjir_geo_net* net) {
  if (net->grid_kind == 1) return net->grid_channels;
  if (net->grid_kind == 2) return 6 * net->grid_channels;
  return 0;
}

}  // namespace

namespace {
// one layer of a head: Y = epi(X W + b), Y as planes (tracked) or fp32 rows; <= 8 outputs take the memory-bound kernels
int head_layer(const ndjir_mlp_layer& L, long long rows, const ndjir_hmat& X, int epi, int precise, float* y32,
               long long ld_y, const ndjir_hmat* yh, cudaStream_t st) {
  ndjir_gemm_h_desc d = {};
  d.M = (int)rows; d.N = L.N; d.K = L.K;
  d.epilogue = epi; d.split_k = 1;
  d.alpha = 1.f; d.out_scale = 1.f; d.beta = 100.f; d.hscale = 1.f;
  d.A = view(X, 0, false);
  d.a_cs = 1; d.b_cs = 1;
  d.bias = L.bias;
  if (L.N <= 8) {
    if (!y32) return NDJIR_ERR_ARG;
    d.B32 = L.W; d.b_rs = L.ldw; d.b_cs = 1;
    d.C = y32; d.ldc = ld_y;
  } else {
    d.B = L.Wt;
    d.precise = precise;
    if (y32) { d.C = y32; d.ldc = ld_y; } else if (yh) { d.Ch = view(*yh, 0, true); } else return NDJIR_ERR_ARG;
  }
  return ndjir_gemm_h(&d, st);
}
}  // namespace

extern "C" int ndjir_mlp_forward(const ndjir_mlp_desc* net, long long rows, const ndjir_hmat* x, const ndjir_hmat* acts,
                                 float* const* out32, const long long* ld_out, const ndjir_hmat* outh, cudaStream_t st) {
  if (!net || !x || rows < 0 || net->n_hidden < 0 || net->n_hidden > NDJIR_MAX_MLP_LAYERS || net->n_out < 1 ||
      net->n_out > 4 || (net->n_hidden && !acts) || !out32 || !ld_out)
    return NDJIR_ERR_ARG;
  if (rows == 0) return NDJIR_OK;
  const ndjir_hmat* cur = x;
  for (int l = 0; l < net->n_hidden; ++l) {
    NDJIR_TRY(head_layer(net->hidden[l], rows, *cur, ndjir::gemm::EPI_SOFTPLUS, net->precise, nullptr, 0, &acts[l], st));
    cur = &acts[l];
  }
  for (int i = 0; i < net->n_out; ++i)
    NDJIR_TRY(head_layer(net->out[i], rows, *cur, ndjir::gemm::EPI_BIAS, net->precise, out32[i], ld_out[i],
                         outh ? &outh[i] : nullptr, st));
  return NDJIR_OK;
}

namespace {
// encoded input [PE(x) | grid features | zero padding]   (network.py:96-151)
int encode_input(const ndjir_geo_net* net, long long rows, const float* x, float* enc, long long ld_enc, float* grid_tmp,
                 cudaStream_t st) {
  const int npe = 3 + 6 * net->pe_bands, gw = grid_width(net), din = npe + gw;
  NDJIR_TRY(ndjir_positional_encoding(rows, 3, net->pe_bands, x, 3, 1, enc, ld_enc, st));
  const float mn[3] = {-1.f, -1.f, -1.f}, mx[3] = {1.f, 1.f, 1.f};     // PF defaults (voxel_feature.py:147-148)
  const int G = net->grid_size, D = net->grid_channels;
  if (net->grid_kind == 1) {
    const int gs[3] = {G, G, G};
    NDJIR_TRY(ndjir_voxel_query_on_voxel(rows, grid_tmp, x, net->grid0, gs, D, mn, mx, 0, st));
    NDJIR_TRY(ndjir_copy2d(rows, D, enc + npe, ld_enc, grid_tmp, D, 1, 1.f, 0, st));
  } else if (net->grid_kind == 2) {
    NDJIR_TRY(ndjir_triplane_query_on_triplane(rows, grid_tmp, x, net->grid0, G, D, mn, mx, 0, st));
    NDJIR_TRY(ndjir_copy2d(rows, 3 * D, enc + npe, ld_enc, grid_tmp, 3 * D, 1, 1.f, 0, st));
    NDJIR_TRY(ndjir_triline_query_on_triline(rows, grid_tmp, x, net->grid1, G, D, mn, mx, 0, st));
    NDJIR_TRY(ndjir_copy2d(rows, 3 * D, enc + npe + 3 * D, ld_enc, grid_tmp, 3 * D, 1, 1.f, 0, st));
  }
  if (ld_enc > din) {
    cudaError_t e = cudaMemset2DAsync(enc + din, ld_enc * sizeof(float), 0, (ld_enc - din) * sizeof(float), rows, st);
    if (e != cudaSuccess) return (int)e;
  }
  return NDJIR_OK;
}

inline ndjir_hmat rows_from(const ndjir_hmat& m, long long row0) {      // rows [row0, ...) of a plane pair
  ndjir_hmat v = m;
  v.hi = reinterpret_cast<char*>(m.hi) + 2 * row0 * m.ld;
  v.lo = reinterpret_cast<char*>(m.lo) + 2 * row0 * m.ld;
  return v;
}
}  // namespace

extern "C" int ndjir_geo_forward(const ndjir_geo_net* net, long long rows, const float* x, float* sdf, float* feat32,
                                 long long ld_feat, const ndjir_geo_store* ws, cudaStream_t st) {
  if (!net || !ws || !x || !sdf || rows < 0 || net->n_hidden < 1 || net->n_hidden > NDJIR_MAX_MLP_LAYERS ||
      net->grid_kind < 0 || net->grid_kind > 2 || !ws->enc)
    return NDJIR_ERR_ARG;
  if (rows == 0) return NDJIR_OK;
  const int din = 3 + 6 * net->pe_bands + grid_width(net);
  if (ws->ld_enc < din || net->hidden[0].K != din || (grid_width(net) && !ws->grid_tmp)) return NDJIR_ERR_ARG;
  for (int l = 0; l <= net->n_hidden; ++l)
    if (!ws->acts[l].hi || !ws->acts[l].lo) return NDJIR_ERR_ARG;
  NDJIR_TRY(encode_input(net, rows, x, ws->enc, ws->ld_enc, ws->grid_tmp, st));
  NDJIR_TRY(pack_cols(rows, din, ws->enc, ws->ld_enc, 1.f, ws->acts[0], 0, st));
  for (int l = 0; l < net->n_hidden; ++l) {
    const ndjir_mlp_layer& L = net->hidden[l];
    const bool into_skip = (l + 1) == net->skip_layer;
    ndjir_gemm_h_desc d = {};
    d.M = (int)rows; d.N = L.N; d.K = L.K;
    d.epilogue = ndjir::gemm::EPI_SOFTPLUS; d.precise = net->precise; d.split_k = 1;
    d.alpha = 1.f; d.out_scale = into_skip ? net->skip_scale : 1.f; d.beta = 100.f; d.hscale = 1.f;
    d.A = view(ws->acts[l], 0, false);
    d.B = L.Wt;
    d.a_cs = 1; d.b_cs = 1;
    d.Ch = view(ws->acts[l + 1], 0, true);
    d.bias = L.bias;
    NDJIR_TRY(ndjir_gemm_h(&d, st));
    if (into_skip) NDJIR_TRY(pack_cols(rows, din, ws->enc, ws->ld_enc, net->skip_scale, ws->acts[l + 1], L.N, st));
  }
  const ndjir_hmat& last = ws->acts[net->n_hidden];
  NDJIR_TRY(head_layer(net->sdf, rows, last, ndjir::gemm::EPI_BIAS, net->precise, sdf, 1, nullptr, st));
  if (feat32) NDJIR_TRY(head_layer(net->feat, rows, last, ndjir::gemm::EPI_BIAS, net->precise, feat32, ld_feat, nullptr, st));
  return NDJIR_OK;
}

extern "C" int ndjir_geo_normal(const ndjir_geo_net* net, long long rows, const float* x, const ndjir_geo_store* fwd,
                                const ndjir_geo_normal_ws* ws, float* normal, long long ld_n, cudaStream_t st) {
  if (!net || !fwd || !ws || !x || !normal || rows < 0 || net->n_hidden < 1 || net->n_hidden > NDJIR_MAX_MLP_LAYERS ||
      !ws->g_in || !ws->ones || !fwd->enc || ld_n != 3)      // (the grid's grad_query writes packed (rows, 3) normals)
    return NDJIR_ERR_ARG;
  if (rows == 0) return NDJIR_OK;
  const int nl = net->n_hidden, npe = 3 + 6 * net->pe_bands, gw = grid_width(net), din = npe + gw;
  if (gw && !ws->grid_tmp) return NDJIR_ERR_ARG;
  const float c = net->skip_scale;
  cudaError_t e = cudaMemsetAsync(ws->g_in, 0, (size_t)rows * fwd->ld_enc * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  auto base = [&](int M, int N, int K, int epi) {
    ndjir_gemm_h_desc d = {};
    d.M = M; d.N = N; d.K = K; d.epilogue = epi; d.split_k = 1;
    d.alpha = 1.f; d.out_scale = 1.f; d.beta = 100.f; d.hscale = 1.f;
    d.a_cs = 1; d.b_cs = 1;
    return d;
  };
  {
    // top: gz[nl-1] = w_sdf (x) s(a_nl): a rank-1 update with the sigmoid factor (ones (rows x 1) times the sdf column)
    ndjir_gemm_h_desc d = base((int)rows, net->sdf.K, 1, ndjir::gemm::EPI_MUL_S);
    d.A32 = ws->ones; d.a_rs = 0; d.a_cs = 1;
    d.B32 = net->sdf.W; d.b_rs = 1; d.b_cs = net->sdf.ldw;
    d.Ch = view(ws->gz[nl - 1], 0, true);
    d.Hh = view(fwd->acts[nl], 0, false);
    NDJIR_TRY(ndjir_gemm_h(&d, st));
  }
  bool skip_wrote = false;
  for (int l = nl - 1; l >= 1; --l) {
    const ndjir_mlp_layer& L = net->hidden[l];
    const bool is_skip = l == net->skip_layer;
    const int n_prev = net->hidden[l - 1].N;
    ndjir_gemm_h_desc d = base((int)rows, n_prev, L.N, ndjir::gemm::EPI_MUL_S);
    d.precise = net->precise;
    d.alpha = is_skip ? c : 1.f; d.hscale = is_skip ? 1.f / c : 1.f;
    d.A = view(ws->gz[l], 0, false);
    d.B = L.Wp;
    d.Ch = view(ws->gz[l - 1], 0, true);
    d.Hh = view(fwd->acts[l], 0, false);
    NDJIR_TRY(ndjir_gemm_h(&d, st));
    if (is_skip) {     // the encoded-input rows of the skip layer's weights
      ndjir_gemm_h_desc s = base((int)rows, din, L.N, ndjir::gemm::EPI_BIAS);
      s.precise = net->precise; s.alpha = c;
      s.A = view(ws->gz[l], 0, false);
      s.B = rows_from(L.Wp, n_prev);
      s.C = ws->g_in; s.ldc = fwd->ld_enc;
      NDJIR_TRY(ndjir_gemm_h(&s, st));
      skip_wrote = true;
    }
  }
  {
    const ndjir_mlp_layer& L = net->hidden[0];
    ndjir_gemm_h_desc d = base((int)rows, din, L.N, skip_wrote ? ndjir::gemm::EPI_ACCUM : ndjir::gemm::EPI_BIAS);
    d.precise = net->precise;
    d.A = view(ws->gz[0], 0, false);
    d.B = L.Wp;
    d.C = ws->g_in; d.ldc = fwd->ld_enc;
    NDJIR_TRY(ndjir_gemm_h(&d, st));
  }
  NDJIR_TRY(ndjir_positional_encoding_grad_input(rows, 3, net->pe_bands, fwd->enc, fwd->ld_enc, ws->g_in, fwd->ld_enc, normal,
                                                 ld_n, 0, st));
  if (net->use_ste) return NDJIR_OK;      // straight-through: no d(grid feature) / d(point) in the normal
  const float mn[3] = {-1.f, -1.f, -1.f}, mx[3] = {1.f, 1.f, 1.f};
  const int G = net->grid_size, D = net->grid_channels;
  if (net->grid_kind == 1) {
    const int gs[3] = {G, G, G};
    NDJIR_TRY(ndjir_copy2d(rows, D, ws->grid_tmp, D, ws->g_in + npe, fwd->ld_enc, 1, 1.f, 0, st));
    NDJIR_TRY(ndjir_voxel_grad_query(rows, normal, ws->grid_tmp, x, net->grid0, gs, D, mn, mx, 1, st));
  } else if (net->grid_kind == 2) {
    NDJIR_TRY(ndjir_copy2d(rows, 3 * D, ws->grid_tmp, 3 * D, ws->g_in + npe, fwd->ld_enc, 1, 1.f, 0, st));
    NDJIR_TRY(ndjir_triplane_grad_query(rows, normal, ws->grid_tmp, x, net->grid0, G, D, mn, mx, 1, st));
    NDJIR_TRY(ndjir_copy2d(rows, 3 * D, ws->grid_tmp, 3 * D, ws->g_in + npe + 3 * D, fwd->ld_enc, 1, 1.f, 0, st));
    NDJIR_TRY(ndjir_triline_grad_query(rows, normal, ws->grid_tmp, x, net->grid1, G, D, mn, mx, 1, st));
  }
  return NDJIR_OK;
}

extern "C" int ndjir_geo_sdf_forward(const ndjir_geo_net* net, long long rows, const float* x, float* sdf,
                                     const ndjir_geo_scratch* ws, cudaStream_t st) {
  if (!net || !ws || !x || !sdf || rows < 0) return NDJIR_ERR_ARG;
  if (rows == 0) return NDJIR_OK;
  if (net->n_hidden < 1 || net->n_hidden > NDJIR_MAX_MLP_LAYERS || net->grid_kind < 0 || net->grid_kind > 2)
    return NDJIR_ERR_ARG;
  const int npe = 3 + 6 * net->pe_bands, gw = grid_width(net), din = npe + gw;
  if (ws->ld_enc < din || !ws->enc || !ws->ench.hi || !ws->act[0].hi || !ws->act[1].hi || (gw && !ws->grid_tmp))
    return NDJIR_ERR_ARG;
  if (net->hidden[0].K != din) return NDJIR_ERR_ARG;
  // encoded input [PE(x) | grid features | zero padding]   (network.py:96-151)
  NDJIR_TRY(ndjir_positional_encoding(rows, 3, net->pe_bands, x, 3, 1, ws->enc, ws->ld_enc, st));
  const float mn[3] = {-1.f, -1.f, -1.f}, mx[3] = {1.f, 1.f, 1.f};     // PF defaults (voxel_feature.py:147-148)
  const int G = net->grid_size, D = net->grid_channels;
  if (net->grid_kind == 1) {
    const int gs[3] = {G, G, G};
    NDJIR_TRY(ndjir_voxel_query_on_voxel(rows, ws->grid_tmp, x, net->grid0, gs, D, mn, mx, 0, st));
    NDJIR_TRY(ndjir_copy2d(rows, D, ws->enc + npe, ws->ld_enc, ws->grid_tmp, D, 1, 1.f, 0, st));
  } else if (net->grid_kind == 2) {
    NDJIR_TRY(ndjir_triplane_query_on_triplane(rows, ws->grid_tmp, x, net->grid0, G, D, mn, mx, 0, st));
    NDJIR_TRY(ndjir_copy2d(rows, 3 * D, ws->enc + npe, ws->ld_enc, ws->grid_tmp, 3 * D, 1, 1.f, 0, st));
    NDJIR_TRY(ndjir_triline_query_on_triline(rows, ws->grid_tmp, x, net->grid1, G, D, mn, mx, 0, st));
    NDJIR_TRY(ndjir_copy2d(rows, 3 * D, ws->enc + npe + 3 * D, ws->ld_enc, ws->grid_tmp, 3 * D, 1, 1.f, 0, st));
  }
  if (ws->ld_enc > din) {
    cudaError_t e = cudaMemset2DAsync(ws->enc + din, ws->ld_enc * sizeof(float), 0, (ws->ld_enc - din) * sizeof(float),
                                      rows, st);
    if (e != cudaSuccess) return (int)e;
  }
  NDJIR_TRY(pack_cols(rows, din, ws->enc, ws->ld_enc, 1.f, ws->ench, 0, st));
  // all layers in ONE kernel with the activations on chip (csrc/gemm_h_chain.cu) when the shapes fit it
  {
    const int rc = ndjir::gemmh::launch_geo_chain(net, rows, din, &ws->ench, ws->enc, ws->ld_enc, ws->act, sdf, st);
    if (rc == NDJIR_OK) return NDJIR_OK;
    if (rc != NDJIR_ERR_ARG) return rc;
  }
  // hidden layers: affine + softplus_100, the skip layer's input is [a | encoded input] / sqrt2   (network.py:160-188)
  ndjir_hmat cur = ws->ench;
  for (int l = 0; l < net->n_hidden; ++l) {
    const ndjir_mlp_layer& L = net->hidden[l];
    const ndjir_hmat& nxt = ws->act[l & 1];
    const bool into_skip = (l + 1) == net->skip_layer;
    ndjir_gemm_h_desc d = {};
    d.M = (int)rows; d.N = L.N; d.K = L.K;
    d.epilogue = ndjir::gemm::EPI_SOFTPLUS; d.precise = net->precise; d.split_k = 1;
    d.alpha = 1.f; d.out_scale = into_skip ? net->skip_scale : 1.f; d.beta = 100.f; d.hscale = 1.f;
    d.A = view(cur, 0, false);
    d.B = L.Wt;
    d.a_cs = 1; d.b_cs = 1;
    d.Ch = view(nxt, 0, true);
    d.bias = L.bias;
    NDJIR_TRY(ndjir_gemm_h(&d, st));
    if (into_skip) NDJIR_TRY(pack_cols(rows, din, ws->enc, ws->ld_enc, net->skip_scale, nxt, L.N, st));
    cur = nxt;
  }
  // sdf column (network.py:190-214): one pass over the last activations, fp32 weights, fp32 output
  ndjir_gemm_h_desc d = {};
  d.M = (int)rows; d.N = 1; d.K = net->sdf.K;
  d.epilogue = ndjir::gemm::EPI_BIAS; d.split_k = 1;
  d.alpha = 1.f; d.out_scale = 1.f; d.beta = 100.f; d.hscale = 1.f;
  d.A = view(cur, 0, false);
  d.a_cs = 1;
  d.B32 = net->sdf.W; d.b_rs = net->sdf.ldw; d.b_cs = 1;
  d.bias = net->sdf.bias;
  d.C = sdf; d.ldc = 1;
  return ndjir_gemm_h(&d, st);
}

extern "C" int ndjir_sdf_lattice(const ndjir_geo_net* net, int G, int ix0, int ix_stride, int n_planes, float radius,
                                 long long batch_points, float* pts, const ndjir_geo_scratch* ws, float* sdf_out,
                                 cudaStream_t st) {
  if (!net || !ws || !pts || !sdf_out || G < 2 || ix0 < 0 || ix_stride < 1 || n_planes < 0) return NDJIR_ERR_ARG;
  const long long plane = (long long)G * G;
  const long long per = batch_points / plane;
  if (per < 1) return NDJIR_ERR_ARG;
  for (long long p0 = 0; p0 < n_planes; p0 += per) {
    const long long cnt = (n_planes - p0) < per ? (n_planes - p0) : per;
    const long long n = cnt * plane;
    NDJIR_TRY(ndjir_lattice_points(n, G, ix0 + (int)p0 * ix_stride, ix_stride, radius, pts, st));
    NDJIR_TRY(ndjir_geo_sdf_forward(net, n, pts, sdf_out + p0 * plane, ws, st));
  }
  return NDJIR_OK;
}

extern "C" int ndjir_sample_points_fwd(const ndjir_sampler_config* cfg, const ndjir_geo_net* net, int B, int R,
                                       const float* camloc, const float* raydir, const float* stratified,
                                       const float* background, const ndjir_sampler_workspace* ws, float* x_fg,
                                       float* t_fg, float* x_bg, float* t_bg, float* mask, float* mask_sum,
                                       cudaStream_t st) {
  if (!cfg || !net || !ws || !camloc || !raydir || !stratified || !background || !x_fg || !t_fg || !x_bg || !t_bg || !mask)
    return NDJIR_ERR_ARG;
  if (B <= 0 || R <= 0 || cfg->n_samples0 < 1 || cfg->n_samples1 < 1 || cfg->n_upsamples < 0 || cfg->n_bg_samples < 1)
    return NDJIR_ERR_ARG;
  const int n = B * R, N0 = cfg->n_samples0, M = cfg->n_samples1, U = cfg->n_upsamples, Nb = cfg->n_bg_samples;
  const int N = N0 + U * M;
  const long long ld = N + 1;
  // ray bounds and the hit mask (sampler.py:59-104)
  if (cfg->bounds == 0) {
    const float lo[3] = {-cfg->radius, -cfg->radius, -cfg->radius}, hi[3] = {cfg->radius, cfg->radius, cfg->radius};
    NDJIR_TRY(ndjir_ray_aabb_intersection(n, ws->t_near, ws->t_far, ws->n_hits, camloc, raydir, B, R, lo, hi, st));
  } else if (cfg->bounds == 1) {
    NDJIR_TRY(ndjir_ray_sphere_intersection(n, ws->t_near, ws->t_far, ws->n_hits, camloc, raydir, B, R, cfg->radius, st));
  } else {
    return NDJIR_ERR_ARG;
  }
  NDJIR_TRY(ndjir_hit_mask(n, ws->n_hits, mask, mask_sum, st));
  // stratified distances, then U SDF-guided rounds.  t_fg holds the sorted distances, merged in place round by round;
  // only the pending samples (the stratified ones, then the M new ones of each round) are evaluated: a sample's SDF
  // does not change between rounds (the reference re-evaluates all of them: sampler.py:190-192)
  NDJIR_TRY(ndjir_stratified_dists(n, N0, ws->t_pend, ws->t_near, ws->t_far, stratified, st));
  const float* pend = ws->t_pend;
  int Nt = 0, Mp = N0;
  for (int u = 0; u <= U; ++u) {
    const bool last = u == U;
    if (!last) {
      NDJIR_TRY(ndjir_ray_points(n, Mp, R, ws->x, camloc, raydir, pend, Mp, st));
      NDJIR_TRY(ndjir_geo_sdf_forward(net, (long long)n * Mp, ws->x, ws->sdf_pend, &ws->geo, st));
    }
    float gain = cfg->sampling_sigmoid_gain;
    for (int i = 0; i < u; ++i) gain *= 2.f;
    float* tnew = ws->t_new[u & 1];
    // the last call only merges the final M samples (their SDF is not needed)
    NDJIR_TRY(ndjir_importance_round_incremental(n, Nt, Mp, last ? 0 : M, t_fg, ld, ws->sdf_cur, N, pend, ws->sdf_pend,
                                                 ws->t_near, ws->t_far, gain, tnew, nullptr, st));
    Nt += Mp;
    pend = tnew;
    Mp = M;
  }
  // t_fg = concat(t, t_far); points of both segments (sampler.py:276-299)
  NDJIR_TRY(ndjir_copy2d(n, 1, t_fg + N, ld, ws->t_far, 1, 1, 1.f, 0, st));
  NDJIR_TRY(ndjir_ray_points(n, N, R, x_fg, camloc, raydir, t_fg, ld, st));
  return ndjir_background_samples(n, Nb, R, camloc, raydir, ws->t_far, mask, background, cfg->radius, t_bg, x_bg, st);
}

// ---------------------------------------------------------------------------------------------------------------------
// reverse sweeps: weight gradients (split-K atomic products over the rows) and input gradients with the sigmoid factor
// ---------------------------------------------------------------------------------------------------------------------
namespace {

inline ndjir_gemm_h_desc product(long long M, int N, long long K, int epi) {
  ndjir_gemm_h_desc d = {};
  d.M = (int)M; d.N = N; d.K = (int)K; d.epilogue = epi; d.split_k = 1;
  d.alpha = 1.f; d.out_scale = 1.f; d.beta = 100.f; d.hscale = 1.f;
  d.a_cs = 1; d.b_cs = 1;
  return d;
}

inline ndjir_mlp_dmat planes(const ndjir_hmat& h) {
  ndjir_mlp_dmat m = {};
  m.dh = h;
  return m;
}

// gW[:kin, :] += X[:, :kin]^T dY;  gb += column sums of dY
int weight_grad(const ndjir_mlp_layer& L, const ndjir_mlp_grad& g, long long rows, const ndjir_hmat& X,
                const ndjir_mlp_dmat& dY, int kin, bool bias, cudaStream_t st) {
  if (!g.gW) return NDJIR_ERR_ARG;
  ndjir_gemm_h_desc d = product(kin, L.N, rows, ndjir::gemm::EPI_ATOMIC);
  d.mn_major = 1;
  d.A = view(X, 0, false);
  d.C = g.gW; d.ldc = L.ldw;
  if (dY.d32) {      // a few fp32 columns: the memory-bound product, the bias gradient as its own column sum
    d.B32 = dY.d32; d.b_rs = dY.ld; d.b_cs = 1;
    NDJIR_TRY(ndjir_gemm_h(&d, st));
    if (bias && g.gb) NDJIR_TRY(ndjir_colsum(rows, L.N, g.gb, dY.d32, dY.ld, 1.f, st));
    return NDJIR_OK;
  }
  const long long tiles = (long long)((kin + 127) / 128) * ((L.N + 255) / 256);
  long long split = (148 * 2) / tiles;
  if (split > rows / 256) split = rows / 256;
  if (split < 1) split = 1;
  d.split_k = (int)split;
  d.B = view(dY.dh, 0, false);
  d.colsum = bias ? g.gb : nullptr;
  return ndjir_gemm_h(&d, st);
}

// dX[:, :ncols] = epi(alpha * dY W[row0 : row0 + ncols, :]^T)   (H: activation of the sigmoid factor, U: addend)
int input_grad(const ndjir_mlp_layer& L, long long rows, const ndjir_mlp_dmat& dY, const ndjir_mlp_dmat& dX, long long dxcol,
               int epi, int row0, int ncols, float alpha, const ndjir_hmat* H, float hscale, const ndjir_hmat* U,
               cudaStream_t st) {
  const bool plain = (epi == ndjir::gemm::EPI_BIAS || epi == ndjir::gemm::EPI_ACCUM) && !H && !U;
  if (!dY.d32 && ncols > 256 && ncols <= 264 && dX.d32 && plain) {
    // [feature (256) | x | normal]: one 256-wide tensor-core tile + a memory-bound pass for the few remaining columns
    NDJIR_TRY(input_grad(L, rows, dY, dX, dxcol, epi, row0, 256, alpha, nullptr, 1.f, nullptr, st));
    return input_grad(L, rows, dY, dX, dxcol + 256, epi, row0 + 256, ncols - 256, alpha, nullptr, 1.f, nullptr, st);
  }
  ndjir_gemm_h_desc d = product(rows, ncols, L.N, epi);
  d.alpha = alpha; d.hscale = hscale;
  if (dX.d32) { d.C = dX.d32 + dxcol; d.ldc = dX.ld; } else { d.Ch = view(dX.dh, dxcol, true); }
  if (H) d.Hh = view(*H, 0, false);
  if (U) d.Uh = view(*U, 0, false);
  const float* w32 = L.W + (long long)row0 * L.ldw;      // B(k, j) = W[row0 + j, k]
  if (dY.d32) {
    d.A32 = dY.d32; d.a_rs = dY.ld; d.a_cs = 1;
    d.B32 = w32; d.b_rs = 1; d.b_cs = L.ldw;
  } else if (ncols <= 8) {
    d.A = view(dY.dh, 0, false);
    d.B32 = w32; d.b_rs = 1; d.b_cs = L.ldw;
  } else {
    if (!L.Wp.hi) return NDJIR_ERR_ARG;
    d.A = view(dY.dh, 0, false);
    d.B = rows_from(L.Wp, row0);
  }
  return ndjir_gemm_h(&d, st);
}

}  // namespace

extern "C" int ndjir_mlp_backward(const ndjir_mlp_desc* net, const ndjir_mlp_grad* g_hidden, const ndjir_mlp_grad* g_out,
                                  long long rows, const ndjir_hmat* x, const ndjir_hmat* acts, const ndjir_mlp_dmat* douts,
                                  const ndjir_hmat* dz, const ndjir_mlp_dmat* dx, int dx_cols, int accum_dx,
                                  cudaStream_t st) {
  if (!net || !g_out || !x || !douts || rows < 0 || net->n_hidden < 0 || net->n_hidden > NDJIR_MAX_MLP_LAYERS ||
      net->n_out < 1 || net->n_out > 4 || (net->n_hidden && (!acts || !dz || !g_hidden)))
    return NDJIR_ERR_ARG;
  if (rows == 0) return NDJIR_OK;
  const int nh = net->n_hidden;
  const ndjir_hmat& last = nh ? acts[nh - 1] : *x;
  int cur = 0;
  for (int i = 0; i < net->n_out; ++i) {
    const ndjir_mlp_layer& L = net->out[i];
    NDJIR_TRY(weight_grad(L, g_out[i], rows, last, douts[i], L.K, true, st));
    if (nh)
      NDJIR_TRY(input_grad(L, rows, douts[i], planes(dz[cur]), 0, ndjir::gemm::EPI_MUL_S, 0, L.K, 1.f, &last, 1.f,
                           i ? &dz[cur] : nullptr, st));
    else if (dx)
      NDJIR_TRY(input_grad(L, rows, douts[i], *dx, 0, (accum_dx || i) ? ndjir::gemm::EPI_ACCUM : ndjir::gemm::EPI_BIAS, 0,
                           dx_cols > 0 ? dx_cols : L.K, 1.f, nullptr, 1.f, nullptr, st));
  }
  for (int l = nh - 1; l >= 0; --l) {
    const ndjir_mlp_layer& L = net->hidden[l];
    const ndjir_hmat& in = l > 0 ? acts[l - 1] : *x;
    NDJIR_TRY(weight_grad(L, g_hidden[l], rows, in, planes(dz[cur]), L.K, true, st));
    if (l > 0) {
      NDJIR_TRY(input_grad(L, rows, planes(dz[cur]), planes(dz[cur ^ 1]), 0, ndjir::gemm::EPI_MUL_S, 0, L.K, 1.f, &in, 1.f,
                           nullptr, st));
      cur ^= 1;
    } else if (dx) {
      NDJIR_TRY(input_grad(L, rows, planes(dz[cur]), *dx, 0, accum_dx ? ndjir::gemm::EPI_ACCUM : ndjir::gemm::EPI_BIAS, 0,
                           dx_cols > 0 ? dx_cols : L.K, 1.f, nullptr, 1.f, nullptr, st));
    }
  }
  return NDJIR_OK;
}

extern "C" int ndjir_geo_backward(const ndjir_geo_net* net, const ndjir_mlp_grad* g_hidden, const ndjir_mlp_grad* g_sdf,
                                  const ndjir_mlp_grad* g_feat, long long rows, const ndjir_geo_store* fwd,
                                  const ndjir_hmat* dfeat, const float* dsdf, const ndjir_hmat* z2, const ndjir_hmat* dz,
                                  float* dgrid, long long ld_dgrid, cudaStream_t st) {
  if (!net || !g_hidden || !g_feat || !fwd || !dfeat || !dz || rows < 0 || net->n_hidden < 1 ||
      net->n_hidden > NDJIR_MAX_MLP_LAYERS || (dsdf && !g_sdf))
    return NDJIR_ERR_ARG;
  if (rows == 0) return NDJIR_OK;
  const int nl = net->n_hidden, npe = 3 + 6 * net->pe_bands, gw = grid_width(net);
  if (gw && (!dgrid || ld_dgrid < gw)) return NDJIR_ERR_ARG;
  const float c = net->skip_scale;
  const ndjir_hmat& last = fwd->acts[nl];
  int cur = 0;
  // output layer: the feature block, then the sdf column on top of it
  NDJIR_TRY(weight_grad(net->feat, *g_feat, rows, last, planes(*dfeat), net->feat.K, true, st));
  NDJIR_TRY(input_grad(net->feat, rows, planes(*dfeat), planes(dz[cur]), 0, ndjir::gemm::EPI_MUL_S, 0, net->feat.K, 1.f,
                       &last, 1.f, z2 ? &z2[nl - 1] : nullptr, st));
  if (dsdf) {
    ndjir_mlp_dmat ds = {};
    ds.d32 = const_cast<float*>(dsdf); ds.ld = 1;
    NDJIR_TRY(weight_grad(net->sdf, *g_sdf, rows, last, ds, net->sdf.K, true, st));
    NDJIR_TRY(input_grad(net->sdf, rows, ds, planes(dz[cur]), 0, ndjir::gemm::EPI_MUL_S, 0, net->sdf.K, 1.f, &last, 1.f,
                         &dz[cur], st));
  }
  ndjir_mlp_dmat dg = {};
  dg.d32 = dgrid; dg.ld = ld_dgrid;
  bool skip_wrote = false;
  for (int l = nl - 1; l >= 0; --l) {
    const ndjir_mlp_layer& L = net->hidden[l];
    NDJIR_TRY(weight_grad(L, g_hidden[l], rows, fwd->acts[l], planes(dz[cur]), L.K, true, st));
    if (l > 0) {
      const bool is_skip = l == net->skip_layer;
      const int n_prev = net->hidden[l - 1].N;
      NDJIR_TRY(input_grad(L, rows, planes(dz[cur]), planes(dz[cur ^ 1]), 0, ndjir::gemm::EPI_MUL_S, 0, n_prev,
                           is_skip ? c : 1.f, &fwd->acts[l], is_skip ? 1.f / c : 1.f, z2 ? &z2[l - 1] : nullptr, st));
      if (is_skip && gw) {      // the grid-feature rows of the skip layer's weights
        NDJIR_TRY(input_grad(L, rows, planes(dz[cur]), dg, 0, ndjir::gemm::EPI_BIAS, n_prev + npe, gw, c, nullptr, 1.f,
                             nullptr, st));
        skip_wrote = true;
      }
      cur ^= 1;
    } else if (gw) {
      NDJIR_TRY(input_grad(L, rows, planes(dz[cur]), dg, 0, skip_wrote ? ndjir::gemm::EPI_ACCUM : ndjir::gemm::EPI_BIAS, npe,
                           gw, 1.f, nullptr, 1.f, nullptr, st));
    }
  }
  return NDJIR_OK;
}

extern "C" int ndjir_geo_normal_adjoint(const ndjir_geo_net* net, const ndjir_mlp_grad* g_hidden, const ndjir_mlp_grad* g_sdf,
                                        long long rows, const ndjir_geo_store* fwd, const ndjir_hmat* gz, const float* gh0,
                                        long long ld_gh0, const ndjir_hmat* gh0h, const ndjir_hmat* ghat,
                                        const ndjir_hmat* z2, const float* ones, cudaStream_t st) {
  if (!net || !g_hidden || !g_sdf || !fwd || !gz || !gh0 || !gh0h || !ghat || !z2 || !ones || rows < 0 ||
      net->n_hidden < 1 || net->n_hidden > NDJIR_MAX_MLP_LAYERS)
    return NDJIR_ERR_ARG;
  if (rows == 0) return NDJIR_OK;
  const int nl = net->n_hidden, din = 3 + 6 * net->pe_bands + grid_width(net);
  const float c = net->skip_scale;
  const ndjir_hmat* in = gh0h;
  for (int l = 0; l < nl; ++l) {
    const ndjir_mlp_layer& L = net->hidden[l];
    const bool into_skip = (l + 1) == net->skip_layer;
    ndjir_gemm_h_desc d = product(rows, L.N, L.K, ndjir::gemm::EPI_ADJ);
    d.out_scale = into_skip ? c : 1.f; d.hscale = into_skip ? 1.f / c : 1.f;
    d.A = view(*in, 0, false);
    d.B = L.Wt;
    d.Hh = view(fwd->acts[l + 1], 0, false);
    d.Uh = view(gz[l], 0, false);
    d.Ch = view(z2[l], 0, true);
    d.C2h = view(ghat[l], 0, true);
    NDJIR_TRY(ndjir_gemm_h(&d, st));
    if (into_skip) NDJIR_TRY(pack_cols(rows, din, gh0, ld_gh0, c, ghat[l], L.N, st));
    NDJIR_TRY(weight_grad(L, g_hidden[l], rows, *in, planes(gz[l]), L.K, false, st));
    in = &ghat[l];
  }
  ndjir_mlp_dmat one = {};      // g w_sdf[k] += sum_p Ghat_top[p, k]: a column of ones (row stride 0)
  one.d32 = const_cast<float*>(ones); one.ld = 0;
  return weight_grad(net->sdf, *g_sdf, rows, *in, one, net->sdf.K, false, st);
}
