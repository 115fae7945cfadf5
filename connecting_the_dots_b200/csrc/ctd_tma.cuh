// TMA (cp.async.bulk.tensor) + mbarrier helpers for sm_100a, and the host-side tensor-map encoder.
// Tiles are fetched by ONE thread per CTA with a single instruction; out-of-bounds elements are
// zero-filled by the hardware, completion is signalled on a shared-memory mbarrier.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ctd {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 128-byte aligned start inside a dynamic shared-memory array.  The offset is derived from the shared-window address
// and ADDED to the array pointer, so the result is still known to point to shared memory (LDS/STS, not generic LD/ST).
__device__ __forceinline__ unsigned char* align128_shared(unsigned char* smem_raw) {
  return smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the barrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// order this thread's generic-proxy shared-memory accesses before subsequent async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Blocks until the phase with the given parity has completed.  A tile that never lands (a bug) traps
// after ~2^26 polls instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1u << 26)) __trap();
}

// 3-D tiled load: coordinates are (x, y, z) = (innermost, ..., outermost) element offsets of the box
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}

// 1-D bulk copy global -> shared (size and both addresses multiples of 16 bytes), completion on `bar`
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Tile coordinates of a persistent CTA that walks tiles t, t + stride, t + 2 stride, ... of a (tiles_x, tiles_y, images)
// grid in x-fastest order.  Decoding t with two runtime divisions per tile and per warp cost the TMA kernels ~100
// instructions per tile (18 % of the LCN kernel's instructions); the walk decodes the first tile and the stride once
// and then adds with carries.
struct TileWalk {
  int tx, ty, n;     // current tile: column, row, image
  int sx, sy, sn;    // the stride in the same mixed radix
  int tiles_x, tiles_y;
  __device__ __forceinline__ void init(int t, int stride, int tiles_x_, int tiles_y_) {
    tiles_x = tiles_x_;
    tiles_y = tiles_y_;
    const int per_image = tiles_x * tiles_y;
    n = t / per_image;
    int rem = t - n * per_image;
    ty = rem / tiles_x;
    tx = rem - ty * tiles_x;
    sn = stride / per_image;
    rem = stride - sn * per_image;
    sy = rem / tiles_x;
    sx = rem - sy * tiles_x;
  }
  __device__ __forceinline__ TileWalk next() const {
    TileWalk w = *this;
    w.tx += sx;
    if (w.tx >= tiles_x) { w.tx -= tiles_x; ++w.ty; }
    w.ty += sy;
    if (w.ty >= tiles_y) { w.ty -= tiles_y; ++w.n; }
    w.n += sn;
    return w;
  }
};

// Host: tensor map of a contiguous fp32 [planes, H, W] array with box (box_w, box_h, 1), zero OOB fill.
// Returns false when TMA cannot address it (unaligned base, W not a multiple of 4, driver entry missing).
bool make_plane_tensor_map(CUtensorMap* map, const float* base, int64_t planes, int64_t H, int64_t W, int box_w,
                           int box_h);

}  // namespace ctd
