// Fused dense optimizer step (sm_100a) for the flat MLP buffer and the feature grids.
//
// Replaces, per training iteration, the reference's sequence solvers.zero_grad() -> solvers.weight_decay(1e-3) ->
// [loss.backward()] -> solvers.check_inf_or_nan_grad() -> solvers.update() (python/train.py:135-148,
// python/solver.py:29-69) over nnabla's S.Adam(alpha, beta1 = 0.9, beta2 = 0.999, eps = 1e-8):
//     g      = dL/dw + decay * w              (weight decay is added to the gradient buffer, solver.py:48-50)
//     m      = beta1 m + (1 - beta1) g
//     v      = beta2 v + (1 - beta2) g^2
//     w      = w - alpha_t m / (sqrt(v) + eps),   alpha_t = alpha sqrt(1 - beta2^t) / (1 - beta1^t)   (host)
// nnabla runs these as separate passes over every parameter (zero 1 stream, decay 3, check 1, update 7); with the
// default 2 GiB voxel grid that is ~25 GB of HBM traffic per iteration.  Here ONE pass reads w, g, m, v and writes w,
// m, v and the zeroed g: 32 B per parameter, HBM-bound.
// Skipping on non-finite gradients keeps the reference's semantics, including its `and` between the two solvers
// (solver.py:67-69, q-list in SURVEY.md section 5): the update is skipped only when BOTH flags are set, so the 2 GiB
// grid gradient only has to be scanned when the (tiny) MLP gradient already is non-finite (`only_if`).
#include "common.cuh"
#include "../../include/ndjir_b200.h"

namespace ndjir {
namespace optimizer {

struct Hyper { float alpha_t, beta1, beta2, eps, decay; };

__device__ __forceinline__ void adam1(float& w, float& g, float& m, float& v, const Hyper& h) {
  float gg = g + h.decay * w;
  m = h.beta1 * m + (1.f - h.beta1) * gg;
  v = h.beta2 * v + (1.f - h.beta2) * gg * gg;
  w = w - h.alpha_t * m / (sqrtf(v) + h.eps);
}

template <bool ZERO>
__global__ void __launch_bounds__(NDJIR_BLOCK)
adam_kernel(long long n, float* __restrict__ w, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            Hyper h, const int* __restrict__ t_dev, const int* __restrict__ skip_flags, bool vec) {
  bool skip = skip_flags && skip_flags[0] != 0 && skip_flags[1] != 0;
  if (t_dev) {   // bias correction from the device-side step counter (it only advances on steps that were not skipped)
    double t = (double)*t_dev;
    h.alpha_t = (float)((double)h.alpha_t * sqrt(1.0 - pow((double)h.beta2, t)) / (1.0 - pow((double)h.beta1, t)));
  }
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  long long n4 = vec ? (n >> 2) : 0;
  if (skip) {                      // train.py:141-146 `continue`: no update; the next iteration zeroes the gradient
    if (ZERO) {
      float4* g4 = reinterpret_cast<float4*>(g);
      for (long long k = i; k < n4; k += stride) __stcs(g4 + k, make_float4(0.f, 0.f, 0.f, 0.f));
      for (long long k = (n4 << 2) + i; k < n; k += stride) g[k] = 0.f;
    }
    return;
  }
  float4 *w4 = reinterpret_cast<float4*>(w), *g4 = reinterpret_cast<float4*>(g), *m4 = reinterpret_cast<float4*>(m),
         *v4 = reinterpret_cast<float4*>(v);
  for (long long k = i; k < n4; k += stride) {
    float4 ww = __ldcs(w4 + k), gg = __ldcs(g4 + k), mm = __ldcs(m4 + k), vv = __ldcs(v4 + k);
    adam1(ww.x, gg.x, mm.x, vv.x, h); adam1(ww.y, gg.y, mm.y, vv.y, h);
    adam1(ww.z, gg.z, mm.z, vv.z, h); adam1(ww.w, gg.w, mm.w, vv.w, h);
    __stcs(w4 + k, ww); __stcs(m4 + k, mm); __stcs(v4 + k, vv);
    if (ZERO) __stcs(g4 + k, make_float4(0.f, 0.f, 0.f, 0.f));
  }
  for (long long k = (n4 << 2) + i; k < n; k += stride) {
    float ww = w[k], gg = g[k], mm = m[k], vv = v[k];
    adam1(ww, gg, mm, vv, h);
    w[k] = ww; m[k] = mm; v[k] = vv;
    if (ZERO) g[k] = 0.f;
  }
}

// flag[0] = 1 if any g is inf or nan (S.Adam.check_inf_or_nan_grad); scanned only if !only_if || *only_if != 0.
__global__ void __launch_bounds__(NDJIR_BLOCK)
nonfinite_kernel(long long n, const float* __restrict__ g, int* __restrict__ flag, const int* __restrict__ only_if,
                 bool vec) {
  if (only_if && *only_if == 0) return;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  long long n4 = vec ? (n >> 2) : 0;
  bool bad = false;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long k = i; k < n4; k += stride) {
    float4 x = __ldcs(g4 + k);
    bad |= !(isfinite(x.x) && isfinite(x.y) && isfinite(x.z) && isfinite(x.w));
  }
  for (long long k = (n4 << 2) + i; k < n; k += stride) bad |= !isfinite(g[k]);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// The reference's second guard (python/train.py:144-146): `if np.any(np.isnan(loss.d)): continue`.  A NaN in `loss`
// raises BOTH skip flags, so the fused update is skipped whatever the gradient scans found.
__global__ void nan_loss_kernel(int n, const float* __restrict__ loss, int* __restrict__ skip_flags) {
  bool bad = false;
  for (int i = threadIdx.x; i < n; i += blockDim.x) bad |= isnan(loss[i]);
  if (__any_sync(0xffffffffu, bad) && threadIdx.x == 0) { skip_flags[0] = 1; skip_flags[1] = 1; }
}

// t += 1 unless the step is skipped (nnabla's Adam counts update() calls; a skipped iteration does not call it)
__global__ void tick_kernel(int* __restrict__ t_dev, const int* __restrict__ skip_flags) {
  bool skip = skip_flags && skip_flags[0] != 0 && skip_flags[1] != 0;
  if (!skip) *t_dev += 1;
}

// g += rate * w (S.Adam.weight_decay as its own pass, for callers that keep the reference's call order)
__global__ void __launch_bounds__(NDJIR_BLOCK)
decay_kernel(long long n, float* __restrict__ g, const float* __restrict__ w, float rate) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) g[k] += rate * w[k];
}

static bool all16(const void* a, const void* b, const void* c, const void* d) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d)) & 15) == 0;
}

}  // namespace optimizer
}  // namespace ndjir

using namespace ndjir;
using namespace ndjir::optimizer;

extern "C" {

int ndjir_adam_tick(int* t_dev, const int* skip_flags, cudaStream_t stream) {
  if (!t_dev) return NDJIR_ERR_ARG;
  tick_kernel<<<1, 1, 0, stream>>>(t_dev, skip_flags);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_adam_step(long long n, float* w, float* g, float* m, float* v, float alpha, float beta1, float beta2,
                    float eps, float weight_decay, const int* t_dev, const int* skip_flags, int zero_grad,
                    cudaStream_t stream) {
  if (n == 0) return NDJIR_OK;
  if (n < 0 || !w || !g || !m || !v) return NDJIR_ERR_ARG;
  Hyper h{alpha, beta1, beta2, eps, weight_decay};
  bool vec = all16(w, g, m, v);
  // two float4 groups per thread in flight per iteration; a grid of whole waves (8 CTAs x 148 SMs x 4)
  int grid = grid_for((n + 3) / 4, NDJIR_BLOCK, 32);
  if (zero_grad) adam_kernel<true><<<grid, NDJIR_BLOCK, 0, stream>>>(n, w, g, m, v, h, t_dev, skip_flags, vec);
  else adam_kernel<false><<<grid, NDJIR_BLOCK, 0, stream>>>(n, w, g, m, v, h, t_dev, skip_flags, vec);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_nonfinite_flag(long long n, const float* g, int* flag, const int* only_if, cudaStream_t stream) {
  if (n == 0) return NDJIR_OK;
  if (n < 0 || !g || !flag) return NDJIR_ERR_ARG;
  bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
  nonfinite_kernel<<<grid_for((n + 3) / 4, NDJIR_BLOCK, 32), NDJIR_BLOCK, 0, stream>>>(n, g, flag, only_if, vec);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_nan_loss_flag(int n, const float* loss, int* skip_flags, cudaStream_t stream) {
  if (n == 0) return NDJIR_OK;
  if (n < 0 || !loss || !skip_flags) return NDJIR_ERR_ARG;
  nan_loss_kernel<<<1, 32, 0, stream>>>(n, loss, skip_flags);
  NDJIR_RETURN_LAST_ERROR();
}

int ndjir_weight_decay(long long n, float* g, const float* w, float rate, cudaStream_t stream) {
  if (n == 0) return NDJIR_OK;
  if (n < 0 || !g || !w) return NDJIR_ERR_ARG;
  decay_kernel<<<grid_for(n, NDJIR_BLOCK, 32), NDJIR_BLOCK, 0, stream>>>(n, g, w, rate);
  NDJIR_RETURN_LAST_ERROR();
}

}  // extern "C"
