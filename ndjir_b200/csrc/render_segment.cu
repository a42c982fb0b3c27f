// ONE fused kernel per ray segment for the compositing stage of pb_render (python/renderer.py:54-91, 179-185;
// python/network.py:544-545), forward and backward.  The reference composes ~30 nnabla ops forward and ~60 backward
// (SURVEY.md a20); the first version here used five / six separate kernels (alpha, background alpha, scan, weighted
// reductions, background colour).  One CTA per ray:
//   forward   alpha of the N foreground samples (NeuS, cos-annealed) and the Nb background samples -> exclusive-cumprod
//             transmittance by a chunked multiplicative WARP SCAN over all N + Nb samples -> weights -> the volume-rendering
//             reduction of the C per-sample columns [feature | x | normal] and the background colour;
//   backward  weight gradients from the three consumers (VR of the C columns, VR of the material attributes, background
//             colour) -> division-free suffix WARP SCAN of the affine recurrence S_{i-1} = wbar_i a_i + (1 - a_i) S_i,
//             abar_i = T_i (wbar_i - S_i) -> NeuS alpha backward (dsdf, dnormal, dgain) and background density backward,
//             plus dV = w dpix for both reductions.
// Arithmetic and operation order are those of the separate kernels in csrc/render.cu (kept as the stage-level C ABI and
// used by the tests as the comparison), so both paths agree to the last bit except for atomics.
#include "common.cuh"
#include "../../include/ndjir_b200.h"

namespace ndjir {
namespace segment {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float softplus100(float x) {
  float z = 100.f * x;
  return (fmaxf(z, 0.f) + log1pf(expf(-fabsf(z)))) / 100.f;
}

constexpr int SEG_THREADS = 256;
constexpr int SEG_MAX_S = 1024;     // samples per ray the shared buffers hold

struct SegIn {
  int N, Nb, C;
  const float* sdf; const float* nrm; long long ld_n;
  const float* raydir; const float* t_fg; const float* gain_p; float c;
  const float* mask;
  const float* h0; long long ld_h; const float* t_bg;
  const float* raw; long long ld_raw;
  const float* V; long long ld_v;
};

// effective alpha of sample j of ray r (foreground: clipped NeuS alpha times the hit mask; background: density alpha);
// *plain receives the unmasked foreground value (what alpha_fg stores)
__device__ __forceinline__ float alpha_of(const SegIn& in, long long r, int j, float g, float m, float* plain) {
  if (j < in.N) {
    const long long p = r * in.N + j;
    const float* d = in.raydir + r * 3;
    const float* nn = in.nrm + p * in.ld_n;
    float cosv = __ldg(d) * __ldg(nn) + __ldg(d + 1) * __ldg(nn + 1) + __ldg(d + 2) * __ldg(nn + 2);
    float ic = -(fmaxf(-cosv * 0.5f + 0.5f, 0.f) * (1.f - in.c) + fmaxf(-cosv, 0.f) * in.c);
    float delta = __ldg(in.t_fg + r * (in.N + 1) + j + 1) - __ldg(in.t_fg + r * (in.N + 1) + j);
    float s = __ldg(in.sdf + p);
    float s1 = s + ic * delta * 0.5f, s0 = s - ic * delta * 0.5f;
    float c0 = sigmoidf_(g * s0), c1 = sigmoidf_(g * s1);
    float a = fminf(fmaxf((c0 - c1 + 1e-5f) / (c0 + 1e-5f), 0.f), 1.f);
    *plain = a;
    return a * m;
  }
  const int jb = j - in.N;
  const long long s = r * in.Nb + jb;
  float delta = __ldg(in.t_bg + r * (in.Nb + 1) + jb + 1) - __ldg(in.t_bg + r * (in.Nb + 1) + jb);
  float a = 1.f - expf(-softplus100(__ldg(in.h0 + s * in.ld_h)) * delta);
  *plain = a;
  return a;
}

__global__ void __launch_bounds__(SEG_THREADS)
segment_fwd_kernel(SegIn in, float* __restrict__ alpha_fg, float* __restrict__ alpha_bg, float* __restrict__ w,
                   float* __restrict__ T, float* __restrict__ pix, long long ld_pix, float* __restrict__ colbg) {
  __shared__ float sa[SEG_MAX_S];
  __shared__ float sw[SEG_MAX_S];
  const long long r = blockIdx.x;
  const int N = in.N, Nb = in.Nb, S = N + Nb;
  const float g = fminf(fmaxf(expf(10.f * __ldg(in.gain_p)), 1e-6f), 5e4f);     // network.py:227-231
  const float m = __ldg(in.mask + r);
  for (int j = threadIdx.x; j < S; j += SEG_THREADS) {
    float plain;
    sa[j] = alpha_of(in, r, j, g, m, &plain);
    if (j < N) alpha_fg[r * N + j] = plain;
    else alpha_bg[r * Nb + (j - N)] = plain;
  }
  __syncthreads();
  if (threadIdx.x < 32) {       // T = exclusive cumprod(1 - alpha): chunked multiplicative warp scan with a carry
    const int lane = threadIdx.x;
    float carry = 1.f;
    for (int base = 0; base < S; base += 32) {
      const int j = base + lane;
      const float a = j < S ? sa[j] : 0.f;
      float incl = 1.f - a;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= up;
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.f;
      const float t = carry * excl;
      if (j < S) { T[r * S + j] = t; w[r * S + j] = a * t; sw[j] = a * t; }
      carry *= __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < in.C; c += SEG_THREADS) {           // VR(v) = sum_i w_i v_i
    const float* p = in.V + r * N * in.ld_v + c;
    float acc = 0.f;
    for (int i = 0; i < N; ++i) acc += sw[i] * __ldg(p + (long long)i * in.ld_v);
    pix[r * ld_pix + c] = acc;
  }
  if (threadIdx.x >= SEG_THREADS - 3 && Nb > 0) {                    // background colour, one thread per channel
    const int k = threadIdx.x - (SEG_THREADS - 3);
    float acc = 0.f;
    for (int j = 0; j < Nb; ++j) acc += sw[N + j] * sigmoidf_(__ldg(in.raw + (r * Nb + j) * in.ld_raw + k));
    colbg[r * 3 + k] = acc;
  }
}

struct SegBwd {
  const float* w; const float* T;                       // from the forward
  const float* dpix; long long ld_dpix;                 // gradient of VR(V)
  const float* V2; long long ld_v2; int C2;             // second reduction (material attributes) and its gradient
  const float* dpix2; long long ld_dpix2;
  const float* dcolbg;
  float* dV; long long ld_dv;
  float* dV2; long long ld_dv2;
  float* draw; long long ld_draw;
  float* dh0; long long ld_dh;
  float* dsdf; float* dnrm; long long ld_dn;
  float* dgain;
  float* dw_out; float* dalpha_fg; float* dalpha_bg;    // optional copies of the intermediates (may be null)
};

__global__ void __launch_bounds__(SEG_THREADS)
segment_bwd_kernel(SegIn in, SegBwd b) {
  __shared__ float sa[SEG_MAX_S];      // effective alpha
  __shared__ float sdw[SEG_MAX_S];     // wbar
  __shared__ float sda[SEG_MAX_S];     // alphabar (w.r.t. the effective alpha)
  extern __shared__ float sd[];        // dpix row (C) | dpix2 row (C2)
  const long long r = blockIdx.x;
  const int N = in.N, Nb = in.Nb, S = N + Nb, C = in.C, C2 = b.C2;
  const float gp = __ldg(in.gain_p);
  const float g_raw = expf(10.f * gp);
  const float g = fminf(fmaxf(g_raw, 1e-6f), 5e4f);
  const bool g_pass = (g_raw >= 1e-6f) && (g_raw <= 5e4f);
  const float m = __ldg(in.mask + r);
  for (int c = threadIdx.x; c < C; c += SEG_THREADS) sd[c] = __ldg(b.dpix + r * b.ld_dpix + c);
  for (int c = threadIdx.x; c < C2; c += SEG_THREADS) sd[C + c] = __ldg(b.dpix2 + r * b.ld_dpix2 + c);
  for (int j = threadIdx.x; j < S; j += SEG_THREADS) {
    float plain;
    sa[j] = alpha_of(in, r, j, g, m, &plain);
    sdw[j] = 0.f;
  }
  __syncthreads();
  // ---- wbar: one warp per foreground sample row (both reductions), dV = w dpix on the way ----
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = SEG_THREADS >> 5;
  for (int i = warp; i < N; i += nw) {
    const long long p = r * N + i;
    const float wi = __ldg(b.w + r * S + i);
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float d = sd[c];
      acc += d * __ldg(in.V + p * in.ld_v + c);
      b.dV[p * b.ld_dv + c] = wi * d;
    }
    for (int c = lane; c < C2; c += 32) {
      const float d = sd[C + c];
      acc += d * __ldg(b.V2 + p * b.ld_v2 + c);
      b.dV2[p * b.ld_dv2 + c] = wi * d;
    }
    acc = warp_sum(acc);
    if (lane == 0) sdw[i] = acc;
  }
  for (int j = threadIdx.x; j < Nb; j += SEG_THREADS) {          // background colour: wbar and d raw
    const long long s = r * Nb + j;
    const float wv = __ldg(b.w + r * S + N + j);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float cc = sigmoidf_(__ldg(in.raw + s * in.ld_raw + k));
      const float gk = __ldg(b.dcolbg + r * 3 + k);
      acc += gk * cc;
      b.draw[s * b.ld_draw + k] = gk * wv * cc * (1.f - cc);
    }
    sdw[N + j] = acc;
  }
  __syncthreads();
  // ---- suffix scan of the affine maps f_j(x) = (1 - a_j) x + wbar_j a_j, abar_j = T_j (wbar_j - S_j) ----
  if (threadIdx.x < 32) {
    const int nchunk = (S + 31) / 32;
    float carryS = 0.f;
    for (int ch = nchunk - 1; ch >= 0; --ch) {
      const int j = ch * 32 + lane;
      const float a = j < S ? sa[j] : 0.f, wb = j < S ? sdw[j] : 0.f;
      float GA = (j < S) ? 1.f - a : 1.f, GB = (j < S) ? wb * a : 0.f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float nA = __shfl_down_sync(0xffffffffu, GA, o);
        float nB = __shfl_down_sync(0xffffffffu, GB, o);
        if (lane + o < 32) { GB = GA * nB + GB; GA = GA * nA; }
      }
      float EA = __shfl_down_sync(0xffffffffu, GA, 1);
      float EB = __shfl_down_sync(0xffffffffu, GB, 1);
      if (lane == 31) { EA = 1.f; EB = 0.f; }
      const float Sj = EA * carryS + EB;
      if (j < S) sda[j] = __ldg(b.T + r * S + j) * (wb - Sj);
      const float GA0 = __shfl_sync(0xffffffffu, GA, 0), GB0 = __shfl_sync(0xffffffffu, GB, 0);
      carryS = GA0 * carryS + GB0;
    }
  }
  __syncthreads();
  // ---- alpha backward: NeuS alpha (dsdf, dnormal, dgain) and background density ----
  float dg_local = 0.f;
  for (int j = threadIdx.x; j < S; j += SEG_THREADS) {
    if (b.dw_out) b.dw_out[r * S + j] = sdw[j];
    if (j < N) {
      const long long p = r * N + j;
      const float da_eff = sda[j];
      float da = da_eff * m;                       // effective alpha = clip(alpha) * mask
      if (b.dalpha_fg) b.dalpha_fg[p] = da;
      const float* d = in.raydir + r * 3;
      const float* nn = in.nrm + p * in.ld_n;
      const float dx = __ldg(d), dy = __ldg(d + 1), dz = __ldg(d + 2);
      const float cosv = dx * __ldg(nn) + dy * __ldg(nn + 1) + dz * __ldg(nn + 2);
      const float ic = -(fmaxf(-cosv * 0.5f + 0.5f, 0.f) * (1.f - in.c) + fmaxf(-cosv, 0.f) * in.c);
      const float delta = __ldg(in.t_fg + r * (N + 1) + j + 1) - __ldg(in.t_fg + r * (N + 1) + j);
      const float s = __ldg(in.sdf + p);
      const float s1 = s + ic * delta * 0.5f, s0 = s - ic * delta * 0.5f;
      const float c0 = sigmoidf_(g * s0), c1 = sigmoidf_(g * s1);
      const float den = c0 + 1e-5f;
      const float a = (c0 - c1 + 1e-5f) / den;
      if (!(a >= 0.f && a <= 1.f)) da = 0.f;       // clip passes gradient inside the range only
      const float dc0 = da * (c1 / (den * den));
      const float dc1 = -da / den;
      const float e0 = dc0 * c0 * (1.f - c0), e1 = dc1 * c1 * (1.f - c1);
      const float ds0 = e0 * g, ds1 = e1 * g;
      dg_local += e0 * s0 + e1 * s1;
      b.dsdf[p] = ds0 + ds1;
      const float dic = (ds1 - ds0) * delta * 0.5f;
      const float dcos = dic * ((-cosv * 0.5f + 0.5f > 0.f ? 0.5f * (1.f - in.c) : 0.f) + (-cosv > 0.f ? in.c : 0.f));
      float* o = b.dnrm + p * b.ld_dn;             // written by the VR phase above (dV's normal columns): add
      o[0] += dcos * dx; o[1] += dcos * dy; o[2] += dcos * dz;
    } else {
      const int jb = j - N;
      const long long s = r * Nb + jb;
      const float da = sda[j];
      if (b.dalpha_bg) b.dalpha_bg[s] = da;
      const float delta = __ldg(in.t_bg + r * (Nb + 1) + jb + 1) - __ldg(in.t_bg + r * (Nb + 1) + jb);
      const float h = __ldg(in.h0 + s * in.ld_h);
      const float dens = softplus100(h);
      b.dh0[s * b.ld_dh] = da * expf(-dens * delta) * delta * sigmoidf_(100.f * h);
    }
  }
  dg_local = warp_sum(dg_local);
  if (lane == 0 && g_pass && dg_local != 0.f) atomicAdd(b.dgain, dg_local * 10.f * g);
}

}  // namespace segment
}  // namespace ndjir

using namespace ndjir;

extern "C" {

int ndjir_render_segment_forward(int n_rays, int N, int Nb, int C, const float* sdf, const float* normal, long long ld_n,
                                 const float* raydir, const float* t_fg, const float* gain_param,
                                 float cos_anneal_ratio, const float* mask, const float* bg_h0, long long ld_h,
                                 const float* t_bg, const float* bg_raw, long long ld_raw, const float* V,
                                 long long ld_v, float* alpha_fg, float* alpha_bg, float* weights, float* trans,
                                 float* pix, long long ld_pix, float* colbg, cudaStream_t stream) {
  if (n_rays == 0) return NDJIR_OK;
  if (n_rays < 0 || N <= 0 || Nb < 0 || C < 0 || N + Nb > segment::SEG_MAX_S || !sdf || !normal || !raydir || !t_fg ||
      !gain_param || !mask || !alpha_fg || !weights || !trans || (C > 0 && (!V || !pix)) ||
      (Nb > 0 && (!bg_h0 || !t_bg || !bg_raw || !alpha_bg || !colbg)))
    return NDJIR_ERR_ARG;
  segment::SegIn in{N, Nb, C, sdf, normal, ld_n, raydir, t_fg, gain_param, cos_anneal_ratio, mask, bg_h0, ld_h, t_bg,
                    bg_raw, ld_raw, V, ld_v};
  segment::segment_fwd_kernel<<<n_rays, segment::SEG_THREADS, 0, stream>>>(in, alpha_fg, alpha_bg, weights, trans, pix,
                                                                         ld_pix, colbg);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_render_segment_backward(int n_rays, int N, int Nb, int C, int C2, const float* sdf, const float* normal,
                                  long long ld_n, const float* raydir, const float* t_fg, const float* gain_param,
                                  float cos_anneal_ratio, const float* mask, const float* bg_h0, long long ld_h,
                                  const float* t_bg, const float* bg_raw, long long ld_raw, const float* V,
                                  long long ld_v, const float* V2, long long ld_v2, const float* weights,
                                  const float* trans, const float* dpix, long long ld_dpix, const float* dpix2,
                                  long long ld_dpix2, const float* dcolbg, float* dV, long long ld_dv, float* dV2,
                                  long long ld_dv2, float* d_bg_raw, long long ld_draw, float* d_bg_h0, long long ld_dh,
                                  float* dsdf, float* dnormal, long long ld_dn, float* dgain_param, float* dweights,
                                  float* dalpha_fg, float* dalpha_bg, cudaStream_t stream) {
  if (n_rays == 0) return NDJIR_OK;
  if (n_rays < 0 || N <= 0 || Nb < 0 || C <= 0 || C2 < 0 || N + Nb > segment::SEG_MAX_S || !sdf || !normal || !raydir ||
      !t_fg || !gain_param || !mask || !V || !weights || !trans || !dpix || !dV || !dsdf || !dnormal || !dgain_param ||
      (C2 > 0 && (!V2 || !dpix2 || !dV2)) || (Nb > 0 && (!bg_h0 || !t_bg || !bg_raw || !dcolbg || !d_bg_raw || !d_bg_h0)))
    return NDJIR_ERR_ARG;
  segment::SegIn in{N, Nb, C, sdf, normal, ld_n, raydir, t_fg, gain_param, cos_anneal_ratio, mask, bg_h0, ld_h, t_bg,
                    bg_raw, ld_raw, V, ld_v};
  segment::SegBwd b{weights, trans, dpix, ld_dpix, V2, ld_v2, C2, dpix2, ld_dpix2, dcolbg, dV, ld_dv, dV2, ld_dv2,
                    d_bg_raw, ld_draw, d_bg_h0, ld_dh, dsdf, dnormal, ld_dn, dgain_param, dweights, dalpha_fg, dalpha_bg};
  segment::segment_bwd_kernel<<<n_rays, segment::SEG_THREADS, (size_t)(C + C2) * sizeof(float), stream>>>(in, b);
  NDJIR_RETURN_LAST_ERROR();
}

}  // extern "C"
