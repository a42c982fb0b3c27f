// Experiment harness: voxel gather / scatter variants at the micro-benchmark shape (2^24 random points, 512^3 x 4).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <curand_kernel.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct Frame { float mn, s, g1; unsigned sx, sy, sz; };

__device__ __forceinline__ void axis(float q, const Frame& f, unsigned& i0, unsigned& i1, float& w0, float& w1) {
  float x = __fmul_rn(__fsub_rn(q, f.mn), f.s);
  float f0 = fminf(fmaxf(floorf(x), 0.f), f.g1);
  float f1 = fminf(f0 + 1.f, f.g1);
  w0 = f1 - x; w1 = 1.f - w0; i0 = (unsigned)f0; i1 = (unsigned)f1;
}

__global__ void init_q(float* q, long long n, unsigned long long seed) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  curandStatePhilox4_32_10_t st; curand_init(seed, i, 0, &st);
  q[i] = curand_uniform(&st) * 2.f - 1.f;
}
__global__ void init_f(float* f, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) f[i] = (float)((i * 2654435761u) & 0xffff) * 1e-6f;
}

// ---- V0: thread per point, ld.global.nc float4, grid-stride
template <int LD>  // 0: __ldg, 1: __ldcg, 2: ld.global.nc.L1::no_allocate
__device__ __forceinline__ float4 ldv(const float* p) {
  if (LD == 0) return __ldg(reinterpret_cast<const float4*>(p));
  if (LD == 1) return __ldcg(reinterpret_cast<const float4*>(p));
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

template <int LD>
__global__ void __launch_bounds__(256) gather_v0(long long B, float* __restrict__ out, const float* __restrict__ query,
                                                const float* __restrict__ feat, Frame f) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < B; p += stride) {
    unsigned x0, x1, y0, y1, z0, z1; float p0, p1, q0, q1, r0, r1;
    axis(__ldg(query + p * 3), f, x0, x1, p0, p1);
    axis(__ldg(query + p * 3 + 1), f, y0, y1, q0, q1);
    axis(__ldg(query + p * 3 + 2), f, z0, z1, r0, r1);
    float4 c000 = ldv<LD>(feat + x0 * f.sx + y0 * f.sy + z0 * f.sz), c001 = ldv<LD>(feat + x0 * f.sx + y0 * f.sy + z1 * f.sz);
    float4 c010 = ldv<LD>(feat + x0 * f.sx + y1 * f.sy + z0 * f.sz), c011 = ldv<LD>(feat + x0 * f.sx + y1 * f.sy + z1 * f.sz);
    float4 c100 = ldv<LD>(feat + x1 * f.sx + y0 * f.sy + z0 * f.sz), c101 = ldv<LD>(feat + x1 * f.sx + y0 * f.sy + z1 * f.sz);
    float4 c110 = ldv<LD>(feat + x1 * f.sx + y1 * f.sy + z0 * f.sz), c111 = ldv<LD>(feat + x1 * f.sx + y1 * f.sy + z1 * f.sz);
    float w000 = p0 * q0 * r0, w001 = p0 * q0 * r1, w010 = p0 * q1 * r0, w011 = p0 * q1 * r1;
    float w100 = p1 * q0 * r0, w101 = p1 * q0 * r1, w110 = p1 * q1 * r0, w111 = p1 * q1 * r1;
    float4 o;
    o.x = w000 * c000.x + w001 * c001.x + w010 * c010.x + w011 * c011.x + w100 * c100.x + w101 * c101.x + w110 * c110.x + w111 * c111.x;
    o.y = w000 * c000.y + w001 * c001.y + w010 * c010.y + w011 * c011.y + w100 * c100.y + w101 * c101.y + w110 * c110.y + w111 * c111.y;
    o.z = w000 * c000.z + w001 * c001.z + w010 * c010.z + w011 * c011.z + w100 * c100.z + w101 * c101.z + w110 * c110.z + w111 * c111.z;
    o.w = w000 * c000.w + w001 * c001.w + w010 * c010.w + w011 * c011.w + w100 * c100.w + w101 * c101.w + w110 * c110.w + w111 * c111.w;
    *reinterpret_cast<float4*>(out + p * 4) = o;
  }
}

// ---- V2/V3/V4: LPP lanes per point (2, 4, 8); each lane loads 8/LPP corners, shuffle-reduce
template <int LPP, int LD>
__global__ void __launch_bounds__(256) gather_lpp(long long B, float* __restrict__ out, const float* __restrict__ query,
                                                  const float* __restrict__ feat, Frame f) {
  long long stride = (long long)gridDim.x * blockDim.x / LPP;
  int sub = threadIdx.x % LPP;
  for (long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPP; p < B; p += stride) {   // B % (32/LPP) == 0 assumed
    unsigned x0, x1, y0, y1, z0, z1; float p0, p1, q0, q1, r0, r1;
    axis(__ldg(query + p * 3), f, x0, x1, p0, p1);
    axis(__ldg(query + p * 3 + 1), f, y0, y1, q0, q1);
    axis(__ldg(query + p * 3 + 2), f, z0, z1, r0, r1);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8 / LPP; ++k) {
      int corner = sub * (8 / LPP) + k;          // bits: x y z
      int cx = (corner >> 2) & 1, cy = (corner >> 1) & 1, cz = corner & 1;
      unsigned xi = cx ? x1 : x0, yi = cy ? y1 : y0, zi = cz ? z1 : z0;
      float w = (cx ? p1 : p0) * (cy ? q1 : q0) * (cz ? r1 : r0);
      float4 c = ldv<LD>(feat + xi * f.sx + yi * f.sy + zi * f.sz);
      o.x += w * c.x; o.y += w * c.y; o.z += w * c.z; o.w += w * c.w;
    }
#pragma unroll
    for (int m = 1; m < LPP; m <<= 1) {
      o.x += __shfl_xor_sync(0xffffffffu, o.x, m); o.y += __shfl_xor_sync(0xffffffffu, o.y, m);
      o.z += __shfl_xor_sync(0xffffffffu, o.z, m); o.w += __shfl_xor_sync(0xffffffffu, o.w, m);
    }
    if (sub == 0) *reinterpret_cast<float4*>(out + p * 4) = o;
  }
}

// ---- V5: reference style, thread per (point, channel), scalar loads
__global__ void __launch_bounds__(256) gather_ref(long long N, float* __restrict__ out, const float* __restrict__ query,
                                                  const float* __restrict__ feat, Frame f) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += stride) {
    long long p = n >> 2; int d = n & 3;
    unsigned x0, x1, y0, y1, z0, z1; float p0, p1, q0, q1, r0, r1;
    axis(query[p * 3], f, x0, x1, p0, p1);
    axis(query[p * 3 + 1], f, y0, y1, q0, q1);
    axis(query[p * 3 + 2], f, z0, z1, r0, r1);
    float acc = p0 * q0 * r0 * feat[x0 * f.sx + y0 * f.sy + z0 * f.sz + d] + p0 * q0 * r1 * feat[x0 * f.sx + y0 * f.sy + z1 * f.sz + d] +
                p0 * q1 * r0 * feat[x0 * f.sx + y1 * f.sy + z0 * f.sz + d] + p0 * q1 * r1 * feat[x0 * f.sx + y1 * f.sy + z1 * f.sz + d] +
                p1 * q0 * r0 * feat[x1 * f.sx + y0 * f.sy + z0 * f.sz + d] + p1 * q0 * r1 * feat[x1 * f.sx + y0 * f.sy + z1 * f.sz + d] +
                p1 * q1 * r0 * feat[x1 * f.sx + y1 * f.sy + z0 * f.sz + d] + p1 * q1 * r1 * feat[x1 * f.sx + y1 * f.sy + z1 * f.sz + d];
    out[n] = acc;
  }
}

// ---- scatter variants: thread per point red.v4 (S0); LPP lanes per point (S1: 2, S2: 4, S3: 8); reference style scalar atomics
template <int LPP>
__global__ void __launch_bounds__(256) scatter_lpp(long long B, float* __restrict__ gf, const float* __restrict__ go,
                                                   const float* __restrict__ query, Frame f) {
  long long stride = (long long)gridDim.x * blockDim.x / LPP;
  int sub = threadIdx.x % LPP;
  for (long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPP; p < B; p += stride) {
    unsigned x0, x1, y0, y1, z0, z1; float p0, p1, q0, q1, r0, r1;
    axis(__ldg(query + p * 3), f, x0, x1, p0, p1);
    axis(__ldg(query + p * 3 + 1), f, y0, y1, q0, q1);
    axis(__ldg(query + p * 3 + 2), f, z0, z1, r0, r1);
    float4 g = __ldg(reinterpret_cast<const float4*>(go + p * 4));
#pragma unroll
    for (int k = 0; k < 8 / LPP; ++k) {
      int corner = sub * (8 / LPP) + k;
      int cx = (corner >> 2) & 1, cy = (corner >> 1) & 1, cz = corner & 1;
      unsigned xi = cx ? x1 : x0, yi = cy ? y1 : y0, zi = cz ? z1 : z0;
      float w = (cx ? p1 : p0) * (cy ? q1 : q0) * (cz ? r1 : r0);
      float* dst = gf + xi * f.sx + yi * f.sy + zi * f.sz;
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(g.x * w), "f"(g.y * w), "f"(g.z * w), "f"(g.w * w) : "memory");
    }
  }
}
__global__ void __launch_bounds__(256) scatter_ref(long long N, float* __restrict__ gf, const float* __restrict__ go,
                                                   const float* __restrict__ query, Frame f) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += stride) {
    long long p = n >> 2; int d = n & 3;
    unsigned x0, x1, y0, y1, z0, z1; float p0, p1, q0, q1, r0, r1;
    axis(query[p * 3], f, x0, x1, p0, p1);
    axis(query[p * 3 + 1], f, y0, y1, q0, q1);
    axis(query[p * 3 + 2], f, z0, z1, r0, r1);
    float g = go[n];
    atomicAdd(gf + x0 * f.sx + y0 * f.sy + z0 * f.sz + d, g * p0 * q0 * r0); atomicAdd(gf + x0 * f.sx + y0 * f.sy + z1 * f.sz + d, g * p0 * q0 * r1);
    atomicAdd(gf + x0 * f.sx + y1 * f.sy + z0 * f.sz + d, g * p0 * q1 * r0); atomicAdd(gf + x0 * f.sx + y1 * f.sy + z1 * f.sz + d, g * p0 * q1 * r1);
    atomicAdd(gf + x1 * f.sx + y0 * f.sy + z0 * f.sz + d, g * p1 * q0 * r0); atomicAdd(gf + x1 * f.sx + y0 * f.sy + z1 * f.sz + d, g * p1 * q0 * r1);
    atomicAdd(gf + x1 * f.sx + y1 * f.sy + z0 * f.sz + d, g * p1 * q1 * r0); atomicAdd(gf + x1 * f.sx + y1 * f.sy + z1 * f.sz + d, g * p1 * q1 * r1);
  }
}

template <class F> float timeit(F fn, int iters = 10) {
  for (int i = 0; i < 3; ++i) fn();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) fn();
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / iters;
}

int main() {
  const int G = 512; const long long B = 1ll << 24;
  long long nf = (long long)G * G * G * 4;
  float *feat, *gf, *q, *out, *go;
  CK(cudaMalloc(&feat, nf * 4)); CK(cudaMalloc(&gf, nf * 4)); CK(cudaMalloc(&q, B * 12)); CK(cudaMalloc(&out, B * 16)); CK(cudaMalloc(&go, B * 16));
  init_f<<<148 * 16, 256>>>(feat, nf);
  init_q<<<(unsigned)((B * 3 + 255) / 256), 256>>>(q, B * 3, 412);
  CK(cudaMemset(gf, 0, nf * 4)); CK(cudaMemset(go, 0x3f, B * 16));
  CK(cudaDeviceSynchronize());
  Frame f; f.mn = -1.f; f.g1 = G - 1.f; f.s = f.g1 / 2.f; f.sx = G * G * 4; f.sy = G * 4; f.sz = 4;
  auto rep = [&](const char* name, float ms, int bytes) { printf("%-44s %8.3f ms  %7.1f GB/s algorithmic (%.3f of 6545.6)\n", name, ms, bytes * (double)B / ms / 1e6, bytes * (double)B / ms / 1e6 / 6545.6); };
  int grids[3] = {148 * 8, 148 * 64, (int)(B / 256)};
  for (int gi = 0; gi < 3; ++gi) {
    int g = grids[gi]; char nm[128];
    snprintf(nm, 128, "gather v0 ldg      grid=%d", g); rep(nm, timeit([&] { gather_v0<0><<<g, 256>>>(B, out, q, feat, f); }), 156);
    snprintf(nm, 128, "gather v0 ldcg     grid=%d", g); rep(nm, timeit([&] { gather_v0<1><<<g, 256>>>(B, out, q, feat, f); }), 156);
    snprintf(nm, 128, "gather v0 noalloc  grid=%d", g); rep(nm, timeit([&] { gather_v0<2><<<g, 256>>>(B, out, q, feat, f); }), 156);
  }
  for (int mult = 1; mult <= 8; mult *= 2) {
    int g = 148 * 64; char nm[128];
    if (mult == 2) { snprintf(nm, 128, "gather 2 lanes/pt ldg"); rep(nm, timeit([&] { gather_lpp<2, 0><<<g, 256>>>(B, out, q, feat, f); }), 156);
                     snprintf(nm, 128, "gather 2 lanes/pt noalloc"); rep(nm, timeit([&] { gather_lpp<2, 2><<<g, 256>>>(B, out, q, feat, f); }), 156); }
    if (mult == 4) { snprintf(nm, 128, "gather 4 lanes/pt ldg"); rep(nm, timeit([&] { gather_lpp<4, 0><<<g, 256>>>(B, out, q, feat, f); }), 156);
                     snprintf(nm, 128, "gather 4 lanes/pt noalloc"); rep(nm, timeit([&] { gather_lpp<4, 2><<<g, 256>>>(B, out, q, feat, f); }), 156); }
    if (mult == 8) { snprintf(nm, 128, "gather 8 lanes/pt ldg"); rep(nm, timeit([&] { gather_lpp<8, 0><<<g, 256>>>(B, out, q, feat, f); }), 156);
                     snprintf(nm, 128, "gather 8 lanes/pt noalloc"); rep(nm, timeit([&] { gather_lpp<8, 2><<<g, 256>>>(B, out, q, feat, f); }), 156); }
  }
  rep("gather reference-style (pt,channel)", timeit([&] { gather_ref<<<65535, 256>>>(B * 4, out, q, feat, f); }), 156);
  rep("scatter 1 lane/pt red.v4", timeit([&] { scatter_lpp<1><<<148 * 64, 256>>>(B, gf, go, q, f); }), 284);
  rep("scatter 2 lanes/pt red.v4", timeit([&] { scatter_lpp<2><<<148 * 64, 256>>>(B, gf, go, q, f); }), 284);
  rep("scatter 4 lanes/pt red.v4", timeit([&] { scatter_lpp<4><<<148 * 64, 256>>>(B, gf, go, q, f); }), 284);
  rep("scatter 8 lanes/pt red.v4", timeit([&] { scatter_lpp<8><<<148 * 64, 256>>>(B, gf, go, q, f); }), 284);
  rep("scatter 8 lanes/pt red.v4 full grid", timeit([&] { scatter_lpp<8><<<(unsigned)(B * 8 / 256), 256>>>(B, gf, go, q, f); }), 284);
  rep("scatter reference-style scalar atomics", timeit([&] { scatter_ref<<<65535, 256>>>(B * 4, gf, go, q, f); }), 284);
  return 0;
}
