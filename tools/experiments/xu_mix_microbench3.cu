// Moving a fraction of the census loop's reciprocal square roots from the XU pipe (MUFU.RSQ) to the FMA pipe (integer
// seed + three packed Newton steps): per 16 rsqrt, SW of them in software, with the loop's other instructions
// (40 FFMA2, 8 LOP3, 4 FMNMX, 6 LDS.64) unchanged.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float rsq(float x) { float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float2 rsq_sw2(float2 x) {  // two values: seed by the shift trick, three Newton steps
  float2 y = make_float2(__int_as_float(0x5f375a86 - (__float_as_int(x.x) >> 1)), __int_as_float(0x5f375a86 - (__float_as_int(x.y) >> 1)));
  const float2 h = __fmul2_rn(x, make_float2(-0.5f, -0.5f));
  const float2 c15 = make_float2(1.5f, 1.5f);
#pragma unroll
  for (int i = 0; i < 3; ++i) y = __fmul2_rn(y, __ffma2_rn(__fmul2_rn(h, y), y, c15));
  return y;
}
template <int SW>  // rsqrt pairs (of 8 pairs per iteration) done in software
__global__ void __launch_bounds__(256, 3) k(float* out, int iters, float seed) {
  __shared__ __align__(16) float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += 256) sm[i] = seed + i;
  __syncthreads();
  float2 x[8];
  float2 a[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = make_float2(seed + threadIdx.x + j, seed + j);
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = make_float2(seed + j, seed - j);
  const float2 c = make_float2(1.0001f, 0.9999f);
  unsigned l = threadIdx.x;
  float mn = 1.f;
  const float* base = sm + 2 * threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 arg = __fadd2_rn(x[j], a[j & 3]);
      if (j < SW) x[j] = rsq_sw2(arg);
      else x[j] = make_float2(rsq(arg.x), rsq(arg.y));
    }
#pragma unroll
    for (int f = 0; f < 40; ++f) a[f & 3] = __ffma2_rn(a[f & 3], c, c);
#pragma unroll
    for (int f = 0; f < 8; ++f) l = (l | 0x80u) ^ (l >> 3);
#pragma unroll
    for (int f = 0; f < 4; ++f) mn = fminf(mn, fminf(fabsf(a[f].x), fabsf(a[f].y)));
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      const float2 v = *reinterpret_cast<const float2*>(base + ((it + f) & 7) * 512);
      a[f & 3].x += v.x; a[f & 3].y += v.y;
    }
  }
  float s = mn;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j].x + x[j].y;
#pragma unroll
  for (int j = 0; j < 4; ++j) s += a[j].x + a[j].y;
  out[blockIdx.x * 256 + threadIdx.x] = s + l;
}
template <int SW>
void run(float* out) {
  const int iters = 10000, grid = 148 * 3;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<SW><<<grid, 256>>>(out, 100, 1.f);
  cudaEventRecord(e0);
  k<SW><<<grid, 256>>>(out, iters, 1.f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double rs = (double)grid * 8 * iters * 16 * 32;
  printf("%d of 16 rsqrt in software: %8.3f ms  %6.2f rsqrt/clk/SM\n", 2 * SW, ms, rs / 148 / (ms * 1e-3) / (clk * 1e3));
}
int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
  run<0>(out); run<1>(out); run<2>(out); run<3>(out); run<4>(out);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
