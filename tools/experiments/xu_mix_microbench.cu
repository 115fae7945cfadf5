// What the XU pipe (MUFU.RSQ) sustains on its own and next to the packed / scalar fp32 instructions of the census loop.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/xu_mix xu_mix_microbench.cu && /tmp/xu_mix
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float rsq(float x) { float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// MODE 0: MUFU only; 1: + NF2 packed FFMA2 per MUFU pair-group; 2: + scalar FFMA; 3: packed + LOP3; 4: packed + LOP3 + LDS
template <int MODE, int NF>
__global__ void __launch_bounds__(256, 3) k(float* out, int iters, float seed) {
  __shared__ float sm[2048];
  for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = seed + i;
  __syncthreads();
  float x[8];
  float2 a[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = seed + threadIdx.x + j;
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = make_float2(seed + j, seed - j);
  const float2 c = make_float2(1.0001f, 0.9999f);
  unsigned l = threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = rsq(x[j]) + 1.5f;  // 8 MUFU + 8 FADD (dependent per chain, 8 chains)
    if (MODE == 1 || MODE == 3 || MODE == 4) {
#pragma unroll
      for (int f = 0; f < NF; ++f) a[f & 3] = __ffma2_rn(a[f & 3], c, c);
    }
    if (MODE == 2) {
#pragma unroll
      for (int f = 0; f < NF; ++f) { a[f & 3].x = fmaf(a[f & 3].x, 1.0001f, 0.9999f); }
    }
    if (MODE == 3 || MODE == 4) {
#pragma unroll
      for (int f = 0; f < 4; ++f) l = (l | 0x80u) ^ (l >> 3);
    }
    if (MODE == 4) {
#pragma unroll
      for (int f = 0; f < 3; ++f) a[f].x += sm[(l + 64 * f) & 2047];
    }
  }
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
#pragma unroll
  for (int j = 0; j < 4; ++j) s += a[j].x + a[j].y;
  out[blockIdx.x * 256 + threadIdx.x] = s + l;
}
template <int MODE, int NF>
void run(const char* name, float* out) {
  const int iters = 20000, grid = 148 * 3;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, NF><<<grid, 256>>>(out, 100, 1.f);
  cudaEventRecord(e0);
  k<MODE, NF><<<grid, 256>>>(out, iters, 1.f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double mufu_warp = (double)grid * 8 * iters * 8;  // warp-level MUFU instructions
  const double per_sm_per_s = mufu_warp * 32 / 148 / (ms * 1e-3);
  printf("%-34s %8.3f ms  %6.2f MUFU lanes/clk/SM at %d MHz (nominal)  other instr per MUFU %.2f\n", name, ms, per_sm_per_s / (clk * 1e3), clk / 1000,
         (MODE == 0 ? 1.0 : MODE == 1 ? 1 + NF / 8.0 : MODE == 2 ? 1 + NF / 8.0 : MODE == 3 ? 1 + (NF + 8) / 8.0 : 1 + (NF + 8 + 6) / 8.0));
}
int main() {
  float* out; cudaMalloc(&out, 148 * 3 * 256 * 4);
  run<0, 0>("MUFU + FADD", out);
  run<1, 8>("+ 1 FFMA2 per MUFU", out);
  run<1, 16>("+ 2 FFMA2 per MUFU", out);
  run<1, 24>("+ 3 FFMA2 per MUFU", out);
  run<1, 32>("+ 4 FFMA2 per MUFU", out);
  run<2, 16>("+ 2 FFMA per MUFU", out);
  run<2, 32>("+ 4 FFMA per MUFU", out);
  run<3, 24>("+ 3 FFMA2 + 1 LOP per MUFU", out);
  run<4, 24>("+ 3 FFMA2 + 1 LOP + .4 LDS per MUFU", out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
