// Does packing the census tap arithmetic into f32x2 instructions relieve the issue-bound fused kernel?
// One "tap" for two pixels: scalar (30 FP/ALU issue slots + 4 MUFU) vs packed across the two pixels
// (12 x2 instructions + 2 FADD + 2 LOP3 + 1 FMNMX-pair + 4 MUFU).  Same arithmetic, same MUFU count.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tapbench tools/experiments/census_tap_packed_microbench.cu && /tmp/tapbench
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float rsq(float a) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }

constexpr int TAPS = 9;  // taps per inner step (one window row)

template <int MODE>
__global__ void __launch_bounds__(256, 3) k(const float* __restrict__ in, float* out, int iters, float eps) {
  __shared__ float row[3][64 + 16];
  for (int i = threadIdx.x; i < 3 * 80; i += 256) row[i / 80][i % 80] = in[i];
  __syncthreads();
  const int tx = threadIdx.x & 31;
  float ec[2] = {in[tx], in[tx + 1]}, tc[2] = {in[tx + 2], in[tx + 3]}, gc[2] = {in[tx + 4], in[tx + 5]};
  float acc[2] = {0, 0}, facc[2] = {0, 0}, nr[2] = {1, 1};
  u64 acc2 = pk(0, 0), facc2 = pk(0, 0);
  const u64 ec2 = pk(ec[0], ec[1]), tc2 = pk(tc[0], tc[1]), gc2 = pk(gc[0], gc[1]), eps2 = pk(eps, eps), neg1 = pk(-1.f, -1.f);
  for (int it = 0; it < iters; ++it) {
    const int o = (it & 7);
    float e[TAPS + 1], t[TAPS + 1], g[TAPS + 1];
#pragma unroll
    for (int j = 0; j < TAPS + 1; ++j) { e[j] = row[0][2 * tx + o + j]; t[j] = row[1][2 * tx + o + j]; g[j] = row[2][2 * tx + o + j]; }
    if (MODE == 0) {
#pragma unroll
      for (int dx = 0; dx < TAPS; ++dx)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const float des = ec[kk] - e[kk + dx], dta = tc[kk] - t[kk + dx];
          const float r1 = rsq(fmaf(des, des, eps)), r2 = rsq(fmaf(dta, dta, eps));
          const float dd = fmaf(des, r1, -(dta * r2));
          const float r3 = r1 * r1 * r1;
          facc[kk] += fabsf(dd);
          const float sr3 = __uint_as_float(__float_as_uint(r3) | (__float_as_uint(dd) & 0x80000000u));
          acc[kk] = fmaf(sr3, g[kk + dx] + gc[kk], acc[kk]);
          nr[kk] = fminf(nr[kk], fabsf(dd));
        }
    } else {
#pragma unroll
      for (int dx = 0; dx < TAPS; ++dx) {
        const u64 e2 = pk(e[dx], e[dx + 1]), t2 = pk(t[dx], t[dx + 1]), g2 = pk(g[dx], g[dx + 1]);  // register moves at worst
        const u64 des = sub2(ec2, e2), dta = sub2(tc2, t2);
        const u64 s1 = fma2(des, des, eps2), s2 = fma2(dta, dta, eps2);
        float s1a, s1b, s2a, s2b;
        upk(s1, s1a, s1b);
        upk(s2, s2a, s2b);
        const u64 r1 = pk(rsq(s1a), rsq(s1b)), r2 = pk(rsq(s2a), rsq(s2b));
        const u64 m = mul2(mul2(dta, r2), neg1);  // -(dta * r2): x2 has no negate modifier (one extra mul, or fold the sign into r2)
        const u64 dd = fma2(des, r1, m);
        const u64 r3 = mul2(mul2(r1, r1), r1);
        float dda, ddb, r3a, r3b;
        upk(dd, dda, ddb);
        upk(r3, r3a, r3b);
        facc[0] += fabsf(dda);
        facc[1] += fabsf(ddb);
        const u64 sr3 = pk(__uint_as_float(__float_as_uint(r3a) | (__float_as_uint(dda) & 0x80000000u)),
                           __uint_as_float(__float_as_uint(r3b) | (__float_as_uint(ddb) & 0x80000000u)));
        acc2 = fma2(sr3, add2(g2, gc2), acc2);
        nr[0] = fminf(nr[0], fabsf(dda));
        nr[1] = fminf(nr[1], fabsf(ddb));
      }
    }
  }
  float a0, a1, f0, f1;
  upk(acc2, a0, a1);
  upk(facc2, f0, f1);
  out[blockIdx.x * 256 + threadIdx.x] = acc[0] + acc[1] + facc[0] + facc[1] + nr[0] + nr[1] + a0 + a1 + f0 + f1;
}
int main() {
  float *in, *out;
  cudaMalloc(&in, 4096);
  cudaMalloc(&out, 148 * 3 * 256 * 4);
  float h[1024];
  for (int i = 0; i < 1024; ++i) h[i] = 0.37f * ((i * 2654435761u) % 1000) / 1000.f - 0.2f;
  cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode)
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 3, 256>>>(in, out, iters, 0.5f);
      else k<1><<<148 * 3, 256>>>(in, out, iters, 0.5f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      // per SMSP: 6 warps, each iters * TAPS two-pixel taps
      if (rep) printf("mode %d (%s): %.3f ms, %.2f cycles per tap and pixel per SMSP\n", mode, mode ? "packed x2" : "scalar", ms,
                      ms * 1e-3 * 1.965e9 / (6.0 * iters * TAPS * 2));
    }
  return 0;
}
