// Shared helpers of libctd_b200: status/error plumbing, launch accounting, small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ctd_b200.h"

#define CTD_API extern "C" __attribute__((visibility("default")))

namespace ctd {

// thread-local last-error message (ctd_last_error)
char* err_buf();
int fail(int code, const char* fmt, ...);
void count_launch(int n = 1);

// Stream-ordered scratch memory from a library-owned pool per device that keeps its memory between calls (no
// torch dependency, no synchronisation, usable under CUDA-graph capture).  nullptr when unavailable.
void* scratch_alloc(size_t bytes, cudaStream_t st);
void scratch_free(void* p, cudaStream_t st);

// Check the launch that was just issued (no synchronisation, like a normal async API).
int check_launch(const char* what);

#define CTD_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return ::ctd::fail(CTD_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define CTD_CUDA(call)                                                                           \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return ::ctd::fail(CTD_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                         __LINE__);                                                              \
  } while (0)

static inline cudaStream_t as_stream(ctd_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ __forceinline__ int clampi(int v, int lo, int hi) {
  return v < lo ? lo : (v > hi ? hi : v);
}

__host__ __device__ __forceinline__ int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// bare MUFU.RSQ (x is always >= eps > 0 where this is used)
__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// read-only 128-bit load
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// the same load pinned in program order (the compiler may not sink it towards its use)
__device__ __forceinline__ float4 ldg4_volatile(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace ctd
