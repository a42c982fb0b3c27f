// A host that is NOT Python running the sampling stage of the fused path through the C ABI alone (include/ndjir_b200.h):
// it reads a parameter / ray file, builds the split-fp16 weight planes with the library's own entry points, describes
// the geometric network and the sampler as PODs, hands over caller-owned scratch and calls ndjir_sample_points_fwd.
// tests/test_capi_host_gpu.py writes the input file from an Engine, runs this program and compares the outputs.
//   g++ -O2 -I include -I /usr/local/cuda/include tests/host/sample_points_host.cpp -o build/sample_points_host \
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

// planes [2][rows][ld] in one allocation, one scale slot
static ndjir_hmat planes(long long rows, long long cols, float* scale, float* amax) {
  ndjir_hmat m;
  m.ld = r64(cols);
  m.hi = dev_alloc<unsigned short>(2 * rows * m.ld);
  m.lo = static_cast<unsigned short*>(m.hi) + rows * m.ld;
  m.scale = scale;
  m.amax = amax;
  return m;
}

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 1; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::perror(argv[1]); return 1; }
  int h[16];
  if (std::fread(h, 4, 16, f) != 16) return 3;
  const int B = h[0], R = h[1], N0 = h[2], M = h[3], U = h[4], Nb = h[5], bounds = h[6], pe_bands = h[7], n_hidden = h[8],
            skip_layer = h[9], grid_kind = h[10], G = h[11], D = h[12], passes = h[13];
  std::vector<float> fl = read_f(f, 4);
  const float gain = fl[0], radius = fl[1], skip_scale = fl[2];
  cudaStream_t st;
  CK(cudaStreamCreate(&st));

  // ---- parameters: fp32 W (K x N, row stride r4(N)), W^T (N x r8(K)), one power-of-two scale for all weight planes ----
  ndjir_geo_net net;
  std::memset(&net, 0, sizeof(net));
  net.n_hidden = n_hidden; net.skip_layer = skip_layer; net.skip_scale = skip_scale; net.pe_bands = pe_bands;
  net.grid_kind = grid_kind; net.grid_size = G; net.grid_channels = D; net.precise = 1;
  float* wscale = dev_alloc<float>(2);            // [scale, amax]
  const float one = 1.f;
  CK(cudaMemcpy(wscale, &one, 4, cudaMemcpyHostToDevice));
  int* wflags = dev_alloc<int>(4);
  struct Staged { float* Wt; long long rows, ldt; ndjir_mlp_layer* L; };
  std::vector<Staged> staged;
  for (int l = 0; l <= n_hidden; ++l) {
    int kn[2];
    if (std::fread(kn, 4, 2, f) != 2) return 3;
    const int K = kn[0], N = kn[1];
    const long long ldw = r4(N);
    std::vector<float> W = read_f(f, (size_t)K * N), b = read_f(f, N), Wp((size_t)K * ldw, 0.f);
    for (int k = 0; k < K; ++k) std::memcpy(&Wp[k * ldw], &W[(size_t)k * N], N * 4);
    ndjir_mlp_layer* L = l < n_hidden ? &net.hidden[l] : &net.sdf;
    L->K = K; L->N = N; L->W = upload(Wp); L->ldw = ldw; L->bias = upload(b);
    CK(ndjir_amax((long long)K * ldw, L->W, wscale + 1, st));
    if (l < n_hidden) {
      const long long ldt = r8(K);
      float* Wt = dev_alloc<float>((size_t)r8(N) * ldt);
      CK(ndjir_transpose(K, N, Wt, ldt, L->W, ldw, st));
      staged.push_back({Wt, r8(N), ldt, L});
    }
  }
  CK(ndjir_scale_update(1, wscale, wscale + 1, wflags, 10, st));
  for (Staged& s : staged) {
    ndjir_hmat m;
    m.ld = s.ldt;
    m.hi = dev_alloc<unsigned short>(2 * s.rows * s.ldt);
    m.lo = static_cast<unsigned short*>(m.hi) + s.rows * s.ldt;
    m.scale = wscale; m.amax = nullptr;
    CK(ndjir_pack_h(1, (int)(s.rows * s.ldt), s.Wt, s.rows * s.ldt, 1, 1.f, &m, st));
    s.L->Wt = m;
  }
  const int gw = grid_kind == 1 ? D : (grid_kind == 2 ? 6 * D : 0);
  if (grid_kind == 1) net.grid0 = upload(read_f(f, (size_t)G * G * G * D));
  if (grid_kind == 2) {
    net.grid0 = upload(read_f(f, (size_t)3 * G * G * D));
    net.grid1 = upload(read_f(f, (size_t)3 * G * D));
  }
  // ---- rays and random inputs ----
  const int n = B * R, N = N0 + U * M, Mx = N0 > M ? N0 : M;
  float* camloc = upload(read_f(f, (size_t)B * 3));
  float* raydir = upload(read_f(f, (size_t)n * 3));
  float* strat = upload(read_f(f, (size_t)n * N0));
  float* backg = upload(read_f(f, (size_t)n * (Nb + 1)));
  std::fclose(f);

  // ---- caller-owned workspace and the three activation scale slots (delayed scaling: settle over `passes` runs) ----
  const int din = 3 + 6 * pe_bands + gw;
  int widest = 0;
  for (int l = 1; l < n_hidden; ++l) widest = net.hidden[l].K > widest ? net.hidden[l].K : widest;
  widest = net.sdf.K > widest ? net.sdf.K : widest;
  float* scales = dev_alloc<float>(3);
  float* amax = dev_alloc<float>(3);
  const float init[3] = {16.f, 16.f, 16.f};
  CK(cudaMemcpy(scales, init, 12, cudaMemcpyHostToDevice));
  int* flags = dev_alloc<int>(4);
  const long long rows = (long long)n * Mx;
  ndjir_sampler_workspace ws;
  std::memset(&ws, 0, sizeof(ws));
  ws.t_near = dev_alloc<float>(n); ws.t_far = dev_alloc<float>(n); ws.n_hits = dev_alloc<float>(n);
  ws.sdf_cur = dev_alloc<float>((size_t)n * N);
  ws.t_pend = dev_alloc<float>((size_t)n * Mx);
  ws.t_new[0] = dev_alloc<float>((size_t)n * M); ws.t_new[1] = dev_alloc<float>((size_t)n * M);
  ws.x = dev_alloc<float>((size_t)rows * 3);
  ws.sdf_pend = dev_alloc<float>(rows);
  ws.geo.ld_enc = r4(din);
  ws.geo.enc = dev_alloc<float>((size_t)rows * ws.geo.ld_enc);
  ws.geo.grid_tmp = gw ? dev_alloc<float>((size_t)rows * gw) : nullptr;
  ws.geo.ench = planes(rows, din, scales + 0, amax + 0);
  ws.geo.act[0] = planes(rows, widest, scales + 1, amax + 1);
  ws.geo.act[1] = planes(rows, widest, scales + 2, amax + 2);
  float* x_fg = dev_alloc<float>((size_t)n * N * 3);
  float* t_fg = dev_alloc<float>((size_t)n * (N + 1));
  float* x_bg = dev_alloc<float>((size_t)n * Nb * 4);
  float* t_bg = dev_alloc<float>((size_t)n * (Nb + 1));
  float* mask = dev_alloc<float>(n);
  float* mask_sum = dev_alloc<float>(1);
  ndjir_sampler_config cfg = {N0, M, U, Nb, gain, bounds, radius};
  for (int p = 0; p < passes; ++p) {
    if (p) CK(ndjir_scale_update(3, scales, amax, flags, 10, st));
    CK(cudaMemsetAsync(mask_sum, 0, 4, st));
    CK(ndjir_sample_points_fwd(&cfg, &net, B, R, camloc, raydir, strat, backg, &ws, x_fg, t_fg, x_bg, t_bg, mask, mask_sum, st));
  }
  CK(cudaStreamSynchronize(st));

  FILE* o = std::fopen(argv[2], "wb");
  if (!o) { std::perror(argv[2]); return 1; }
  auto dump = [&](const float* d, size_t cnt) {
    std::vector<float> v(cnt);
    CK(cudaMemcpy(v.data(), d, cnt * 4, cudaMemcpyDeviceToHost));
    std::fwrite(v.data(), 4, cnt, o);
  };
  dump(x_fg, (size_t)n * N * 3); dump(t_fg, (size_t)n * (N + 1)); dump(x_bg, (size_t)n * Nb * 4);
  dump(t_bg, (size_t)n * (Nb + 1)); dump(mask, n); dump(mask_sum, 1);
  std::fclose(o);
  std::printf("sample_points_host: %d rays, %d + %d samples, mask sum written\n", n, N, Nb);
  return 0;
}
