// Masked loss reduction of the caller of PhotometricLoss (model/networks.py:377):
// val = (mask * diff).sum() / mask.sum().  One kernel produces both sums (numerator, denominator)
// deterministically: fixed grid, fixed per-block order, the last block to finish adds the block
// partials in index order.  Under batch sharding these two scalars are what crosses NVLink.
#include <algorithm>

#include "ctd_common.cuh"

namespace ctd {

constexpr int RD_THREADS = 256;
constexpr int RD_MAX_BLOCKS = 148 * 4;

__global__ void __launch_bounds__(RD_THREADS)
masked_sums_kernel(const float* __restrict__ diff, const float* __restrict__ mask, int64_t n, float* __restrict__ out2,
                   float* __restrict__ partial, unsigned int* __restrict__ ticket, int vec) {
  __shared__ float s_num[RD_THREADS / 32], s_den[RD_THREADS / 32];
  __shared__ bool last;
  float num = 0.f, den = 0.f;
  const int64_t tid = (int64_t)blockIdx.x * RD_THREADS + threadIdx.x, nthreads = (int64_t)gridDim.x * RD_THREADS;
  if (vec) {
    for (int64_t i = tid; i < n / 4; i += nthreads) {
      const float4 d = ldg4(diff + 4 * i), m = ldg4(mask + 4 * i);
      num += (m.x * d.x + m.y * d.y) + (m.z * d.z + m.w * d.w);
      den += (m.x + m.y) + (m.z + m.w);
    }
    for (int64_t i = (n / 4) * 4 + tid; i < n; i += nthreads) {
      num += __ldg(mask + i) * __ldg(diff + i);
      den += __ldg(mask + i);
    }
  } else {
    for (int64_t i = tid; i < n; i += nthreads) {
      num += __ldg(mask + i) * __ldg(diff + i);
      den += __ldg(mask + i);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    num += __shfl_xor_sync(0xffffffffu, num, o);
    den += __shfl_xor_sync(0xffffffffu, den, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_num[threadIdx.x >> 5] = num;
    s_den[threadIdx.x >> 5] = den;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < RD_THREADS / 32; ++w) {
      a += s_num[w];
      b += s_den[w];
    }
    partial[2 * blockIdx.x] = a;
    partial[2 * blockIdx.x + 1] = b;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
    double a = 0.0, b = 0.0;  // few hundred partials: fp64 keeps the final sum order-insensitive to ~1e-16
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) {
      a += (double)__ldcg(partial + 2 * i);
      b += (double)__ldcg(partial + 2 * i + 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (threadIdx.x == 0) {
      out2[0] = (float)a;
      out2[1] = (float)b;
      *ticket = 0;  // ready for the next launch on this workspace
    }
  }
}

// library-internal form: workspace from the scratch pool (zeroed here), used by the unfused fall-back of
// ctd_photometric_fwd_bwd_masked_f32
int masked_sums_launch(const float* diff, const float* mask, int64_t n, float* out2, cudaStream_t st) {
  const size_t bytes = 2 * RD_MAX_BLOCKS * sizeof(float) + 256;
  char* ws = static_cast<char*>(scratch_alloc(bytes, st));
  if (!ws) return fail(CTD_ERR_NOMEM, "masked_sums: no scratch memory for the block partials");
  if (cudaMemsetAsync(ws, 0, 256, st) != cudaSuccess) {
    cudaGetLastError();
    scratch_free(ws, st);
    return fail(CTD_ERR_CUDA, "masked_sums: cannot zero the ticket");
  }
  const int vec = ((reinterpret_cast<uintptr_t>(diff) | reinterpret_cast<uintptr_t>(mask)) & 15) == 0;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(RD_MAX_BLOCKS, cdiv(n, RD_THREADS * 8)));
  masked_sums_kernel<<<blocks, RD_THREADS, 0, st>>>(diff, mask, n, out2, reinterpret_cast<float*>(ws + 256),
                                                    reinterpret_cast<unsigned int*>(ws), vec);
  count_launch();
  scratch_free(ws, st);
  return CTD_OK;
}

}  // namespace ctd

using namespace ctd;

CTD_API int64_t ctd_masked_sums_workspace_bytes(void) { return (int64_t)(2 * RD_MAX_BLOCKS * sizeof(float) + 256); }

// out2[0] = sum(mask * diff), out2[1] = sum(mask).  `workspace` (ctd_masked_sums_workspace_bytes() bytes,
// 16-byte aligned, zero-filled once before its first use) must not be shared by concurrent launches.
CTD_API int ctd_masked_sums_f32(const float* diff, const float* mask, int64_t n, float* out2, void* workspace,
                                ctd_stream_t stream) {
  CTD_REQUIRE(n >= 0, "masked_sums: negative size");
  CTD_REQUIRE(out2 && workspace && (n == 0 || (diff && mask)), "masked_sums: null pointer");
  const int vec = ((reinterpret_cast<uintptr_t>(diff) | reinterpret_cast<uintptr_t>(mask)) & 15) == 0;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(RD_MAX_BLOCKS, cdiv(n, RD_THREADS * 8)));
  unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
  masked_sums_kernel<<<blocks, RD_THREADS, 0, as_stream(stream)>>>(diff, mask, n, out2, partial, ticket, vec);
  count_launch();
  return check_launch("masked_sums");
}
