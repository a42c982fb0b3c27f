// A host that is NOT Python running the geometric network of the TRAIN step through the C ABI alone
// (include/ndjir_b200.h): forward with kept activations (ndjir_geo_forward), normal = d sdf / d x (ndjir_geo_normal),
// and for the loss  L = 1/2 sum sdf^2 + 1/2 * 0.01 sum feature^2 + 1/2 * 0.1 sum |normal|^2  the full reverse pass:
// adjoint of the normal (ndjir_positional_encoding_grad_input_adjoint, the grids' *_grad_query_grad_grad_output,
// ndjir_geo_normal_adjoint) and the standard sweep with the second-order addends (ndjir_geo_backward).
// tests/test_capi_host_gpu.py writes the input file from an Engine, runs this program and compares sdf, normal, every
// weight / bias gradient and the grid-feature gradient with the engine's.
//   g++ -O2 -I include -I /usr/local/cuda/include tests/host/geo_train_host.cpp -o build_host/geo_train_host \
//       -L ndjir_b200 -lndjir_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/ndjir_b200
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ndjir_b200.h"

#define CK(x)                                                                  \
  do {                                                                         \
    int rc_ = (int)(x);                                                        \
    if (rc_ != 0) { std::fprintf(stderr, "%s failed: %d (line %d)\n", #x, rc_, __LINE__); std::exit(2); } \
  } while (0)

static std::vector<float> read_f(FILE* f, size_t n) {
  std::vector<float> v(n);
  if (n && std::fread(v.data(), 4, n, f) != n) { std::fprintf(stderr, "short read\n"); std::exit(3); }
  return v;
}
template <class T>
static T* dev_alloc(size_t n) {
  void* p = nullptr;
  CK(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
  CK(cudaMemset(p, 0, (n ? n : 1) * sizeof(T)));
  return static_cast<T*>(p);
}
static float* upload(const std::vector<float>& v) {
  float* p = dev_alloc<float>(v.size());
  if (!v.empty()) CK(cudaMemcpy(p, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  return p;
}
static long long r4(long long x) { return (x + 3) / 4 * 4; }
static long long r8(long long x) { return (x + 7) / 8 * 8; }
static long long r64(long long x) { return (x + 63) / 64 * 64; }

// scale slots: scale[i], amax[i]; activations start at 2^4, gradient tensors at 2^24 (delayed scaling settles them)
struct Slots {
  float *scale, *amax;
  int n = 0, cap;
  std::vector<float> init;
  explicit Slots(int cap_) : cap(cap_) { scale = dev_alloc<float>(cap_); amax = dev_alloc<float>(cap_); }
  int add(bool grad) { init.push_back(grad ? 16777216.f : 16.f); return n++; }
  void commit() { CK(cudaMemcpy(scale, init.data(), n * 4, cudaMemcpyHostToDevice)); }
};
static ndjir_hmat planes(long long rows, long long cols, Slots& s, bool grad) {
  ndjir_hmat m;
  const int i = s.add(grad);
  m.ld = r64(cols);
  m.hi = dev_alloc<unsigned short>(2 * rows * m.ld);
  m.lo = static_cast<unsigned short*>(m.hi) + rows * m.ld;
  m.scale = s.scale + i;
  m.amax = s.amax + i;
  return m;
}
static ndjir_hmat untracked(ndjir_hmat m) { m.amax = nullptr; return m; }

struct Layer {               // one affine layer on the device: parameters, their planes and the gradient buffers
  ndjir_mlp_layer L;
  ndjir_mlp_grad g;
  float *Wt32 = nullptr;     // W^T as fp32 rows (staging for the planes)
  long long ldt = 0;
};

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 1; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::perror(argv[1]); return 1; }
  int h[16];
  if (std::fread(h, 4, 16, f) != 16) return 3;
  const long long rows = h[0];
  const int pe_bands = h[1], nl = h[2], skip_layer = h[3], grid_kind = h[4], G = h[5], D = h[6], passes = h[7], Df = h[8];
  const float skip_scale = read_f(f, 1)[0];
  cudaStream_t st;
  CK(cudaStreamCreate(&st));

  // ---- parameters: hidden layers, the sdf column and the feature block of the last reference layer ----
  float* wscale = dev_alloc<float>(2);            // [scale, amax]: one power-of-two scale for all weight planes
  const float one = 1.f;
  CK(cudaMemcpy(wscale, &one, 4, cudaMemcpyHostToDevice));
  int* wflags = dev_alloc<int>(4);
  std::vector<Layer> layers(nl + 2);
  for (int l = 0; l < nl + 2; ++l) {
    int kn[2];
    if (std::fread(kn, 4, 2, f) != 2) return 3;
    const int K = kn[0], N = kn[1];
    const long long ldw = N > 8 ? r8(N) : r4(N);
    std::vector<float> W = read_f(f, (size_t)K * N), b = read_f(f, N), Wp((size_t)K * ldw, 0.f);
    for (int k = 0; k < K; ++k) std::memcpy(&Wp[k * ldw], &W[(size_t)k * N], N * 4);
    Layer& y = layers[l];
    std::memset(&y.L, 0, sizeof(y.L));
    y.L.K = K; y.L.N = N; y.L.W = upload(Wp); y.L.ldw = ldw; y.L.bias = upload(b);
    y.g.gW = dev_alloc<float>((size_t)K * ldw);
    y.g.gb = dev_alloc<float>(N);
    CK(ndjir_amax((long long)K * ldw, y.L.W, wscale + 1, st));
    if (N > 8) {
      y.ldt = r8(K);
      y.Wt32 = dev_alloc<float>((size_t)r8(N) * y.ldt);
      CK(ndjir_transpose(K, N, y.Wt32, y.ldt, y.L.W, ldw, st));
    }
  }
  CK(ndjir_scale_update(1, wscale, wscale + 1, wflags, 10, st));
  auto weight_planes = [&](const float* src, long long r, long long ld) {
    ndjir_hmat m;
    m.ld = ld;
    m.hi = dev_alloc<unsigned short>(2 * r * ld);
    m.lo = static_cast<unsigned short*>(m.hi) + r * ld;
    m.scale = wscale; m.amax = nullptr;
    CK(ndjir_pack_h(1, (int)(r * ld), src, r * ld, 1, 1.f, &m, st));
    return m;
  };
  for (Layer& y : layers)
    if (y.L.N > 8) {
      y.L.Wt = weight_planes(y.Wt32, r8(y.L.N), y.ldt);
      y.L.Wp = weight_planes(y.L.W, y.L.K, y.L.ldw);
    }
  ndjir_geo_net net;
  std::memset(&net, 0, sizeof(net));
  net.n_hidden = nl; net.skip_layer = skip_layer; net.skip_scale = skip_scale; net.pe_bands = pe_bands;
  net.grid_kind = grid_kind; net.grid_size = G; net.grid_channels = D; net.precise = 1;
  for (int l = 0; l < nl; ++l) net.hidden[l] = layers[l].L;
  net.sdf = layers[nl].L;
  net.feat = layers[nl + 1].L;
  const int gw = grid_kind == 1 ? D : (grid_kind == 2 ? 6 * D : 0);
  if (grid_kind == 1) net.grid0 = upload(read_f(f, (size_t)G * G * G * D));
  if (grid_kind == 2) {
    net.grid0 = upload(read_f(f, (size_t)3 * G * G * D));
    net.grid1 = upload(read_f(f, (size_t)3 * G * D));
  }
  float* x = upload(read_f(f, (size_t)rows * 3));
  std::fclose(f);

  // ---- caller-owned buffers ----
  const int npe = 3 + 6 * pe_bands, din = npe + gw;
  const long long ld_enc = r4(din);
  Slots slots(8 * (nl + 2));
  ndjir_geo_store fwd;
  std::memset(&fwd, 0, sizeof(fwd));
  fwd.ld_enc = ld_enc;
  fwd.enc = dev_alloc<float>((size_t)rows * ld_enc);
  fwd.grid_tmp = gw ? dev_alloc<float>((size_t)rows * gw) : nullptr;
  fwd.acts[0] = planes(rows, din, slots, false);
  for (int l = 1; l < nl; ++l) fwd.acts[l] = planes(rows, net.hidden[l].K, slots, false);
  fwd.acts[nl] = planes(rows, net.sdf.K, slots, false);
  ndjir_geo_normal_ws nws;
  std::memset(&nws, 0, sizeof(nws));
  for (int l = 0; l < nl; ++l) nws.gz[l] = planes(rows, net.hidden[l].N, slots, true);
  nws.g_in = dev_alloc<float>((size_t)rows * ld_enc);
  nws.grid_tmp = gw ? dev_alloc<float>((size_t)rows * gw) : nullptr;
  nws.ones = upload(std::vector<float>(4, 1.f));
  std::vector<ndjir_hmat> ghat(nl), z2(nl), gz_in(nl), z2_in(nl);
  for (int l = 0; l < nl; ++l) {
    ghat[l] = planes(rows, l + 1 < nl ? net.hidden[l + 1].K : net.sdf.K, slots, true);
    z2[l] = planes(rows, net.hidden[l].N, slots, true);
  }
  ndjir_hmat dz[2] = {planes(rows, net.sdf.K, slots, true), planes(rows, net.sdf.K, slots, true)};
  ndjir_hmat dfeat = planes(rows, Df, slots, true), gh0h = planes(rows, din, slots, true);
  slots.commit();
  int* flags = dev_alloc<int>(4);
  float* sdf = dev_alloc<float>(rows);
  float* feat = dev_alloc<float>((size_t)rows * Df);
  float* normal = dev_alloc<float>((size_t)rows * 3);
  float* dsdf = dev_alloc<float>(rows);
  float* dfeat32 = dev_alloc<float>((size_t)rows * Df);
  float* nbar = dev_alloc<float>((size_t)rows * 3);
  float* gh0 = dev_alloc<float>((size_t)rows * ld_enc);
  float* ggo = gw ? dev_alloc<float>((size_t)rows * gw) : nullptr;
  float* dgrid = gw ? dev_alloc<float>((size_t)rows * gw) : nullptr;
  std::vector<ndjir_mlp_grad> g_hidden(nl);
  for (int l = 0; l < nl; ++l) g_hidden[l] = layers[l].g;
  const float mn[3] = {-1.f, -1.f, -1.f}, mx[3] = {1.f, 1.f, 1.f};

  for (int p = 0; p < passes; ++p) {
    if (p) CK(ndjir_scale_update(slots.n, slots.scale, slots.amax, flags, 10, st));
    for (Layer& y : layers) {      // gradients accumulate: start every pass from zero
      CK(cudaMemsetAsync(y.g.gW, 0, (size_t)y.L.K * y.L.ldw * 4, st));
      CK(cudaMemsetAsync(y.g.gb, 0, (size_t)y.L.N * 4, st));
    }
    CK(ndjir_geo_forward(&net, rows, x, sdf, feat, Df, &fwd, st));
    CK(ndjir_geo_normal(&net, rows, x, &fwd, &nws, normal, 3, st));
    // upstream gradients of L
    CK(ndjir_copy2d(rows, 1, dsdf, 1, sdf, 1, 1, 1.f, 0, st));
    CK(ndjir_copy2d(rows, Df, dfeat32, Df, feat, Df, 1, 0.01f, 0, st));
    CK(ndjir_copy2d(rows, 3, nbar, 3, normal, 3, 1, 0.1f, 0, st));
    CK(ndjir_pack_h(rows, Df, dfeat32, Df, 1, 1.f, &dfeat, st));
    // seed of the normal's adjoint: d L / d g_in = adjoint of the encoding's (and the grids') input gradient
    CK(cudaMemsetAsync(gh0, 0, (size_t)rows * ld_enc * 4, st));
    CK(ndjir_positional_encoding_grad_input_adjoint(rows, 3, pe_bands, fwd.enc, ld_enc, nbar, 3, gh0, ld_enc, st));
    if (grid_kind == 1) {
      const int gs[3] = {G, G, G};
      CK(ndjir_voxel_grad_query_grad_grad_output(rows, ggo, nbar, x, net.grid0, gs, D, mn, mx, 0, st));
      CK(ndjir_copy2d(rows, D, gh0 + npe, ld_enc, ggo, D, 1, 1.f, 0, st));
    } else if (grid_kind == 2) {
      CK(ndjir_triplane_grad_query_grad_grad_output(rows, ggo, nbar, x, net.grid0, G, D, mn, mx, 0, st));
      CK(ndjir_copy2d(rows, 3 * D, gh0 + npe, ld_enc, ggo, 3 * D, 1, 1.f, 0, st));
      CK(ndjir_triline_grad_query_grad_grad_output(rows, ggo, nbar, x, net.grid1, G, D, mn, mx, 0, st));
      CK(ndjir_copy2d(rows, 3 * D, gh0 + npe + 3 * D, ld_enc, ggo, 3 * D, 1, 1.f, 0, st));
    }
    CK(ndjir_pack_h(rows, din, gh0, ld_enc, 1, 1.f, &gh0h, st));
    for (int l = 0; l < nl; ++l) gz_in[l] = untracked(nws.gz[l]);
    ndjir_hmat gh0_in = untracked(gh0h);
    CK(ndjir_geo_normal_adjoint(&net, g_hidden.data(), &layers[nl].g, rows, &fwd, gz_in.data(), gh0, ld_enc, &gh0_in,
                                ghat.data(), z2.data(), nws.ones, st));
    for (int l = 0; l < nl; ++l) z2_in[l] = untracked(z2[l]);
    ndjir_hmat dfeat_in = untracked(dfeat);
    CK(ndjir_geo_backward(&net, g_hidden.data(), &layers[nl].g, &layers[nl + 1].g, rows, &fwd, &dfeat_in, dsdf,
                          z2_in.data(), dz, dgrid, gw, st));
  }
  CK(cudaStreamSynchronize(st));

  FILE* o = std::fopen(argv[2], "wb");
  if (!o) { std::perror(argv[2]); return 1; }
  auto dump = [&](const float* d, size_t cnt) {
    std::vector<float> v(cnt);
    CK(cudaMemcpy(v.data(), d, cnt * 4, cudaMemcpyDeviceToHost));
    std::fwrite(v.data(), 4, cnt, o);
  };
  dump(sdf, rows);
  dump(normal, (size_t)rows * 3);
  for (Layer& y : layers) {       // gradients in the (K, N) layout of the parameters
    std::vector<float> gw_pad((size_t)y.L.K * y.L.ldw);
    CK(cudaMemcpy(gw_pad.data(), y.g.gW, gw_pad.size() * 4, cudaMemcpyDeviceToHost));
    for (int k = 0; k < y.L.K; ++k) std::fwrite(&gw_pad[(size_t)k * y.L.ldw], 4, y.L.N, o);
    dump(y.g.gb, y.L.N);
  }
  if (gw) dump(dgrid, (size_t)rows * gw);
  std::fclose(o);
  std::printf("geo_train_host: %lld points, %d hidden layers, gradients written\n", rows, nl);
  return 0;
}
