// The census loop's instruction mix around 8 MUFU.RSQ: 20 packed fp32 (FFMA2), 5 scalar fp32, 4 LOP3, 2 FMNMX and a
// variable number / width of conflict-free shared-memory loads -- how much XU throughput do the loads cost?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float rsq(float x) { float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// NL loads of LW floats per 8 MUFU; PACK: packed (1) or scalar (0: 40 FFMA instead of 20 FFMA2)
template <int NL, int LW, int PACK, int CTAS>
__global__ void __launch_bounds__(256, CTAS) k(float* out, int iters, float seed) {
  __shared__ __align__(16) float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += 256) sm[i] = seed + i;
  __syncthreads();
  float x[8];
  float2 a[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = seed + threadIdx.x + j;
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = make_float2(seed + j, seed - j);
  const float2 c = make_float2(1.0001f, 0.9999f);
  unsigned l = threadIdx.x;
  float mn = 1.f;
  const float* base = sm + LW * threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = rsq(x[j] + a[j & 3].x);  // 8 MUFU, 8 FADD
    if (PACK) {
#pragma unroll
      for (int f = 0; f < 20; ++f) a[f & 3] = __ffma2_rn(a[f & 3], c, c);
    } else {
#pragma unroll
      for (int f = 0; f < 20; ++f) { a[f & 3].x = fmaf(a[f & 3].x, 1.0001f, 0.9999f); a[f & 3].y = fmaf(a[f & 3].y, 0.9999f, 1.0001f); }
    }
#pragma unroll
    for (int f = 0; f < 4; ++f) l = (l | 0x80u) ^ (l >> 3);
    mn = fminf(mn, fminf(fabsf(a[0].x), fabsf(a[1].y)));
    mn = fminf(mn, fminf(fabsf(a[2].x), fabsf(a[3].y)));
#pragma unroll
    for (int f = 0; f < NL; ++f) {
      const float* p = base + ((it + f) & 3) * (LW * 256);
      if (LW == 1) a[f & 3].x += *p;
      if (LW == 2) { const float2 v = *reinterpret_cast<const float2*>(p); a[f & 3].x += v.x; a[f & 3].y += v.y; }
      if (LW == 4) { const float4 v = *reinterpret_cast<const float4*>(p); a[f & 3].x += v.x + v.z; a[f & 3].y += v.y + v.w; }
    }
  }
  float s = mn;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
#pragma unroll
  for (int j = 0; j < 4; ++j) s += a[j].x + a[j].y;
  out[blockIdx.x * 256 + threadIdx.x] = s + l;
}
template <int NL, int LW, int PACK, int CTAS>
void run(const char* name, float* out) {
  const int iters = 20000, grid = 148 * CTAS;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NL, LW, PACK, CTAS><<<grid, 256>>>(out, 100, 1.f);
  cudaEventRecord(e0);
  k<NL, LW, PACK, CTAS><<<grid, 256>>>(out, iters, 1.f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double mufu_warp = (double)grid * 8 * iters * 8;
  printf("%-44s %8.3f ms  %6.2f MUFU lanes/clk/SM\n", name, ms, mufu_warp * 32 / 148 / (ms * 1e-3) / (clk * 1e3));
}
int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
  run<0, 1, 1, 3>("packed, no loads, 24 warps", out);
  run<1, 2, 1, 3>("packed, 1 LDS.64 / 8 MUFU", out);
  run<2, 2, 1, 3>("packed, 2 LDS.64 / 8 MUFU", out);
  run<3, 2, 1, 3>("packed, 3 LDS.64 / 8 MUFU", out);
  run<4, 2, 1, 3>("packed, 4 LDS.64 / 8 MUFU", out);
  run<1, 4, 1, 3>("packed, 1 LDS.128 / 8 MUFU", out);
  run<2, 4, 1, 3>("packed, 2 LDS.128 / 8 MUFU", out);
  run<3, 1, 1, 3>("packed, 3 LDS.32 / 8 MUFU", out);
  run<0, 1, 0, 3>("scalar, no loads, 24 warps", out);
  run<3, 2, 0, 3>("scalar, 3 LDS.64 / 8 MUFU", out);
  run<0, 1, 1, 2>("packed, no loads, 16 warps", out);
  run<1, 4, 1, 2>("packed, 1 LDS.128 / 8 MUFU, 16 warps", out);
  run<3, 2, 1, 2>("packed, 3 LDS.64 / 8 MUFU, 16 warps", out);
  run<3, 2, 1, 4>("packed, 3 LDS.64 / 8 MUFU, 32 warps", out);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
