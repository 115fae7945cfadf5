// PhotometricLoss mse / sad, block size 9, C = 1, fp32: persistent TMA-fed kernels for sm_100a.
//
// Same arithmetic as photo_fwd_box9 / photo_bwd_box9 (photometric.cu) -- a separable 9x9 box filter of
// phi(es - ta), and its adjoint for the backward -- but the data movement is Blackwell's: a CTA stays
// resident, walks a strided list of 128x16 output tiles, and one thread fetches each tile's halo box
// (136x24 floats per tensor) with a single cp.async.bulk.tensor instruction into a two-stage shared
// memory ring, completion signalled on an mbarrier.  The next tile's loads are in flight while the
// current tile is filtered, so load latency, arithmetic and stores overlap inside every SM instead of
// every CTA serialising load -> sync -> compute -> store.  TMA zero-fills out-of-image elements: that is
// exactly the backward's zero padding; the forward's replicate clamp is an index remap of the staged rows, and
// in tiles that touch the left/right border the (at most 4 + 4) box columns outside the image are overwritten
// with the border column once per tile (patch_clamped_columns).
#include <algorithm>

#include "ctd_common.cuh"
#include "ctd_tma.cuh"

namespace ctd {

extern int g_force_generic;
extern int g_disable_tma;

constexpr int R9 = 4;
constexpr int TT_W = 128, TT_H = 16;             // output tile
constexpr int TB_W = TT_W + 2 * R9;              // 136: box width (544 B rows, 16-byte multiple)
constexpr int TB_H = TT_H + 2 * R9;              // 24
constexpr int TBOX_BYTES = TB_W * TB_H * 4;      // 13056 = 102 * 128
constexpr int NSTAGE = 2;
constexpr float INV81 = 1.0f / 81.0f;
static_assert(TBOX_BYTES % 128 == 0, "stages must stay 128-byte aligned");

struct alignas(128) FwdSmem {
  float es[NSTAGE][TB_H][TB_W];
  float ta[NSTAGE][TB_H][TB_W];
  float hs[TB_H][TT_W];
  uint64_t full[NSTAGE];
};
struct alignas(128) BwdSmem {
  float go[NSTAGE][TB_H][TB_W];
  float hs[TB_H][TT_W];
  uint64_t full[NSTAGE];
};

__device__ __forceinline__ float4 hsum9x4(const float* v) {  // v[0..11] -> four adjacent 9-sums
  const float mid = ((v[3] + v[4]) + (v[5] + v[6])) + (v[7] + v[8]);
  const float l12 = v[1] + v[2], r910 = v[9] + v[10];
  return make_float4(mid + (v[0] + l12), mid + (l12 + v[9]), mid + (v[2] + r910), mid + (r910 + v[11]));
}
__device__ __forceinline__ float4 add4(const float4 a, const float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 fma4s(const float s, const float4 a, const float4 b) {
  return make_float4(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z), fmaf(s, a.w, b.w));
}
__device__ __forceinline__ void ld12(float* d, const float* p) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1],
               c = reinterpret_cast<const float4*>(p)[2];
  d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
  d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
  d[8] = c.x; d[9] = c.y; d[10] = c.z; d[11] = c.w;
}
__device__ __forceinline__ float sgnf(float d) { return d < 0.f ? -1.f : (d > 0.f ? 1.f : 0.f); }

struct TileCoord {
  int x0, y0, n;
};
__device__ __forceinline__ TileCoord tile_coord(const TileWalk& w) {
  TileCoord c;
  c.x0 = w.tx * TT_W;
  c.y0 = w.ty * TT_H;
  c.n = w.n;
  return c;
}
__device__ __forceinline__ TileCoord tile_coord(int t, int tiles_x, int tiles_y) {
  TileCoord c;
  c.x0 = (t % tiles_x) * TT_W;
  c.y0 = ((t / tiles_x) % tiles_y) * TT_H;
  c.n = t / (tiles_x * tiles_y);
  return c;
}

// vertical 9-sums of two adjacent output rows from ten rows of horizontal sums
__device__ __forceinline__ void vsum2(const float4* v, float4& o0, float4& o1) {
  const float4 mid = add4(add4(add4(v[1], v[2]), add4(v[3], v[4])), add4(add4(v[5], v[6]), add4(v[7], v[8])));
  o0 = add4(mid, v[0]);
  o1 = add4(mid, v[9]);
}

// Tiles touching the left / right image border: TMA zero-fills the box columns outside the image; the forward's
// replicate clamp wants the first / last image column there.  Patching those (at most 4 + 4) columns of the staged
// es / ta boxes once per tile lets every thread take the 128-bit path below (an index remap per tap costs a border
// tile three times the shared-memory instructions, and two of five tile columns of a 640-wide image are border
// tiles).  Block-uniform; ends with a barrier.
template <typename SM>
__device__ __forceinline__ void patch_clamped_columns(SM& S, int s, int x0, int W, int tid) {
  const int lo = R9 - x0;          // box column of image column 0 (4 in the first tile column)
  const int hi = W - 1 - (x0 - R9);  // box column of image column W - 1
  for (int i = tid; i < TB_H * 2 * R9; i += 256) {
    const int r = i / (2 * R9), k = i % (2 * R9);
    if (k < R9) {
      if (k < lo) {
        S.es[s][r][k] = S.es[s][r][lo];
        S.ta[s][r][k] = S.ta[s][r][lo];
      }
    } else {
      const int cc = hi + 1 + (k - R9);
      if (cc < TB_W) {
        S.es[s][r][cc] = S.es[s][r][hi];
        S.ta[s][r][cc] = S.ta[s][r][hi];
      }
    }
  }
  __syncthreads();
}

template <int TYPE>
__global__ void __launch_bounds__(256, 3)
photo_fwd_box9_tma(const __grid_constant__ CUtensorMap map_es, const __grid_constant__ CUtensorMap map_ta,
                   float* __restrict__ out, int H, int W, int tiles_x, int tiles_y, int ntiles) {
  extern __shared__ unsigned char smem_raw[];
  FwdSmem& S = *reinterpret_cast<FwdSmem*>(align128_shared(smem_raw));
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&S.full[0], 1);
    mbar_init(&S.full[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  int t = blockIdx.x;
  TileWalk walk;
  walk.init(t, gridDim.x, tiles_x, tiles_y);
  if (tid == 0 && t < ntiles) {
    const TileCoord c = tile_coord(walk);
    mbar_expect_tx(&S.full[0], 2 * TBOX_BYTES);
    tma_load_3d(&S.es[0][0][0], &map_es, &S.full[0], c.x0 - R9, c.y0 - R9, c.n);
    tma_load_3d(&S.ta[0][0][0], &map_ta, &S.full[0], c.x0 - R9, c.y0 - R9, c.n);
  }
  for (int it = 0; t < ntiles; ++it, t += gridDim.x, walk = walk.next()) {
    const int s = it & 1;
    if (tid == 0 && t + (int)gridDim.x < ntiles) {  // prefetch the next tile into the other stage
      const TileCoord c = tile_coord(walk.next());
      fence_proxy_async();
      mbar_expect_tx(&S.full[s ^ 1], 2 * TBOX_BYTES);
      tma_load_3d(&S.es[s ^ 1][0][0], &map_es, &S.full[s ^ 1], c.x0 - R9, c.y0 - R9, c.n);
      tma_load_3d(&S.ta[s ^ 1][0][0], &map_ta, &S.full[s ^ 1], c.x0 - R9, c.y0 - R9, c.n);
    }
    const TileCoord c = tile_coord(walk);
    mbar_wait(&S.full[s], (it >> 1) & 1);
    // phase A: horizontal 9-sums of phi(es - ta), replicate clamp by index remap
    if (c.x0 == 0 || c.x0 + TT_W + R9 > W) patch_clamped_columns(S, s, c.x0, W, tid);  // block-uniform
    for (int i = tid; i < TB_H * (TT_W / 4); i += 256) {
      const int r = i / (TT_W / 4), q = i % (TT_W / 4);
      const int rr = clampi(c.y0 - R9 + r, 0, H - 1) - (c.y0 - R9);
      float e[12], v[12];
      ld12(e, &S.es[s][rr][4 * q]);
      ld12(v, &S.ta[s][rr][4 * q]);
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const float d = e[j] - v[j];
        v[j] = TYPE == 0 ? d * d : fabsf(d);
      }
      *reinterpret_cast<float4*>(&S.hs[r][4 * q]) = hsum9x4(v);
    }
    __syncthreads();
    // phase B: vertical 9-sums, two output rows per thread, 128-bit stores
    {
      const int q = tid % 32, rs = tid / 32;
      float4 v[10];
#pragma unroll
      for (int j = 0; j < 10; ++j) v[j] = *reinterpret_cast<const float4*>(&S.hs[2 * rs + j][4 * q]);
      float4 o0, o1;
      vsum2(v, o0, o1);
      const int gx = c.x0 + 4 * q, gy = c.y0 + 2 * rs;
      if (gx < W) {
        float* dst = out + ((int64_t)c.n * H + gy) * W + gx;
        if (gy < H) *reinterpret_cast<float4*>(dst) = make_float4(o0.x * INV81, o0.y * INV81, o0.z * INV81, o0.w * INV81);
        if (gy + 1 < H) *reinterpret_cast<float4*>(dst + W) = make_float4(o1.x * INV81, o1.y * INV81, o1.z * INV81, o1.w * INV81);
      }
    }
    __syncthreads();
  }
}

template <int TYPE>
__global__ void __launch_bounds__(256, 3)
photo_bwd_box9_tma(const __grid_constant__ CUtensorMap map_go, const float* __restrict__ es,
                   const float* __restrict__ ta, float* __restrict__ gi, int H, int W, int tiles_x, int tiles_y,
                   int ntiles) {
  extern __shared__ unsigned char smem_raw[];
  BwdSmem& S = *reinterpret_cast<BwdSmem*>(align128_shared(smem_raw));
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&S.full[0], 1);
    mbar_init(&S.full[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  int t = blockIdx.x;
  TileWalk walk;
  walk.init(t, gridDim.x, tiles_x, tiles_y);
  if (tid == 0 && t < ntiles) {
    const TileCoord c = tile_coord(walk);
    mbar_expect_tx(&S.full[0], TBOX_BYTES);
    tma_load_3d(&S.go[0][0][0], &map_go, &S.full[0], c.x0 - R9, c.y0 - R9, c.n);
  }
  for (int it = 0; t < ntiles; ++it, t += gridDim.x, walk = walk.next()) {
    const int s = it & 1;
    if (tid == 0 && t + (int)gridDim.x < ntiles) {
      const TileCoord c = tile_coord(walk.next());
      fence_proxy_async();
      mbar_expect_tx(&S.full[s ^ 1], TBOX_BYTES);
      tma_load_3d(&S.go[s ^ 1][0][0], &map_go, &S.full[s ^ 1], c.x0 - R9, c.y0 - R9, c.n);
    }
    const TileCoord c = tile_coord(walk);
    // this thread's es/ta pixels do not depend on the staged tile: fetch them while the TMA lands
    const int q = tid % 32, rs = tid / 32;
    const int gx = c.x0 + 4 * q, gy = c.y0 + 2 * rs;
    const int64_t off = ((int64_t)c.n * H + gy) * W + gx;
    float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), t0 = e0, e1 = e0, t1 = e0;
    if (gx < W) {
      if (gy < H) { e0 = ldg4(es + off); t0 = ldg4(ta + off); }
      if (gy + 1 < H) { e1 = ldg4(es + off + W); t1 = ldg4(ta + off + W); }
    }
    mbar_wait(&S.full[s], (it >> 1) & 1);
    // phase A: horizontal sums of the zero-padded grad_out; the first / last image column collects
    // the clamp multiplicity (weights 5,4,3,2,1 instead of 1,1,1,1,1)
    const int xr = W - 1 - c.x0;  // tile-local column of the last image column
    for (int i = tid; i < TB_H * (TT_W / 4); i += 256) {
      const int r = i / (TT_W / 4), qq = i % (TT_W / 4);
      float v[12];
      ld12(v, &S.go[s][r][4 * qq]);
      float4 o = hsum9x4(v);
      if (c.x0 == 0 && qq == 0) o.x += 4.f * v[4] + 3.f * v[5] + 2.f * v[6] + v[7];
      if (xr >= 0 && xr < TT_W && qq == xr / 4) {
        const float* g = &S.go[s][r][xr + R9];
        const float extra = 4.f * g[0] + 3.f * g[-1] + 2.f * g[-2] + g[-3];
        const int j = xr % 4;
        if (j == 0) o.x += extra;
        if (j == 1) o.y += extra;
        if (j == 2) o.z += extra;
        if (j == 3) o.w += extra;
      }
      *reinterpret_cast<float4*>(&S.hs[r][4 * qq]) = o;
    }
    __syncthreads();
    {
      float4 v[10];
#pragma unroll
      for (int j = 0; j < 10; ++j) v[j] = *reinterpret_cast<const float4*>(&S.hs[2 * rs + j][4 * q]);
      float4 o[2];
      vsum2(v, o[0], o[1]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (gy + j >= H || gx >= W) continue;
        float4 sv = o[j];  // output row gy + j sits at staged row 2*rs + j + 4 = v[j + 4]
        if (gy + j == 0) sv = fma4s(4.f, v[j + 4], fma4s(3.f, v[j + 5], fma4s(2.f, v[j + 6], add4(sv, v[j + 7]))));
        if (gy + j == H - 1) sv = fma4s(4.f, v[j + 4], fma4s(3.f, v[j + 3], fma4s(2.f, v[j + 2], add4(sv, v[j + 1]))));
        const float4 e = j ? e1 : e0, tt = j ? t1 : t0;
        float4 r;
        if (TYPE == 0) {
          r = make_float4(2.f * (e.x - tt.x) * INV81 * sv.x, 2.f * (e.y - tt.y) * INV81 * sv.y,
                          2.f * (e.z - tt.z) * INV81 * sv.z, 2.f * (e.w - tt.w) * INV81 * sv.w);
        } else {
          r = make_float4(sgnf(e.x - tt.x) * INV81 * sv.x, sgnf(e.y - tt.y) * INV81 * sv.y,
                          sgnf(e.z - tt.z) * INV81 * sv.z, sgnf(e.w - tt.w) * INV81 * sv.w);
        }
        *reinterpret_cast<float4*>(gi + off + (int64_t)j * W) = r;
      }
    }
    __syncthreads();
  }
}

// Forward and backward of one tile in a single pass: the three halo boxes (es, ta, grad_out) are staged once,
// phase A forms both horizontal 9-sums (phi(es - ta) with the replicate clamp, grad_out with the border
// re-weighting), phase B both vertical sums; phi'(es - ta) comes from the staged tiles, so the fused kernel moves
// 20 B/px (three reads, two writes) instead of 12 + 16.
struct alignas(128) FbSmem {
  float es[NSTAGE][TB_H][TB_W];
  float ta[NSTAGE][TB_H][TB_W];
  float go[NSTAGE][TB_H][TB_W];
  float hf[TB_H][TT_W];  // horizontal sums of phi(es - ta)
  float hb[TB_H][TT_W];  // horizontal sums of grad_out
  uint64_t full[NSTAGE];
};

template <int TYPE>
__global__ void __launch_bounds__(256, 2)
photo_fwd_bwd_box9_tma(const __grid_constant__ CUtensorMap map_es, const __grid_constant__ CUtensorMap map_ta,
                       const __grid_constant__ CUtensorMap map_go, float* __restrict__ out, float* __restrict__ gi, int H, int W,
                       int tiles_x, int tiles_y, int ntiles, const float* __restrict__ mask, double* __restrict__ partials,
                       unsigned* __restrict__ ticket, float* __restrict__ sums2) {
  extern __shared__ unsigned char smem_raw[];
  FbSmem& S = *reinterpret_cast<FbSmem*>(align128_shared(smem_raw));
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&S.full[0], 1);
    mbar_init(&S.full[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto fetch = [&](const TileWalk& wk, int stage) {
    const TileCoord c = tile_coord(wk);
    mbar_expect_tx(&S.full[stage], 3 * TBOX_BYTES);
    tma_load_3d(&S.es[stage][0][0], &map_es, &S.full[stage], c.x0 - R9, c.y0 - R9, c.n);
    tma_load_3d(&S.ta[stage][0][0], &map_ta, &S.full[stage], c.x0 - R9, c.y0 - R9, c.n);
    tma_load_3d(&S.go[stage][0][0], &map_go, &S.full[stage], c.x0 - R9, c.y0 - R9, c.n);
  };
  int t = blockIdx.x;
  TileWalk walk;
  walk.init(t, gridDim.x, tiles_x, tiles_y);
  if (tid == 0 && t < ntiles) fetch(walk, 0);
  double mnum = 0.0, mden = 0.0;  // optional masked-mean terms, accumulated over this CTA's tiles
  for (int it = 0; t < ntiles; ++it, t += gridDim.x, walk = walk.next()) {
    const int s = it & 1;
    if (tid == 0 && t + (int)gridDim.x < ntiles) {  // prefetch the next tile into the other stage
      fence_proxy_async();
      fetch(walk.next(), s ^ 1);
    }
    const TileCoord c = tile_coord(walk);
    mbar_wait(&S.full[s], (it >> 1) & 1);
    if (c.x0 == 0 || c.x0 + TT_W + R9 > W) patch_clamped_columns(S, s, c.x0, W, tid);  // block-uniform
    const int xr = W - 1 - c.x0;  // tile-local column of the last image column
    for (int i = tid; i < TB_H * (TT_W / 4); i += 256) {
      const int r = i / (TT_W / 4), q = i % (TT_W / 4);
      // forward: replicate clamp -- rows by index remap, columns patched into the staged boxes above
      const int rr = clampi(c.y0 - R9 + r, 0, H - 1) - (c.y0 - R9);
      float e[12], v[12];
      ld12(e, &S.es[s][rr][4 * q]);
      ld12(v, &S.ta[s][rr][4 * q]);
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const float d = e[j] - v[j];
        v[j] = TYPE == 0 ? d * d : fabsf(d);
      }
      *reinterpret_cast<float4*>(&S.hf[r][4 * q]) = hsum9x4(v);
      // backward: zero-padded grad_out, first / last image column collects the clamp multiplicity
      ld12(v, &S.go[s][r][4 * q]);
      float4 o = hsum9x4(v);
      if (c.x0 == 0 && q == 0) o.x += 4.f * v[4] + 3.f * v[5] + 2.f * v[6] + v[7];
      if (xr >= 0 && xr < TT_W && q == xr / 4) {
        const float* g = &S.go[s][r][xr + R9];
        const float extra = 4.f * g[0] + 3.f * g[-1] + 2.f * g[-2] + g[-3];
        const int j = xr % 4;
        if (j == 0) o.x += extra;
        if (j == 1) o.y += extra;
        if (j == 2) o.z += extra;
        if (j == 3) o.w += extra;
      }
      *reinterpret_cast<float4*>(&S.hb[r][4 * q]) = o;
    }
    __syncthreads();
    {
      const int q = tid % 32, rs = tid / 32;
      const int gx = c.x0 + 4 * q, gy = c.y0 + 2 * rs;
      float4 v[10];
#pragma unroll
      for (int j = 0; j < 10; ++j) v[j] = *reinterpret_cast<const float4*>(&S.hf[2 * rs + j][4 * q]);
      float4 f[2];
      vsum2(v, f[0], f[1]);
#pragma unroll
      for (int j = 0; j < 10; ++j) v[j] = *reinterpret_cast<const float4*>(&S.hb[2 * rs + j][4 * q]);
      float4 o[2];
      vsum2(v, o[0], o[1]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (gy + j >= H || gx >= W) continue;
        float4 sv = o[j];  // output row gy + j sits at staged row 2*rs + j + 4 = v[j + 4]
        if (gy + j == 0) sv = fma4s(4.f, v[j + 4], fma4s(3.f, v[j + 5], fma4s(2.f, v[j + 6], add4(sv, v[j + 7]))));
        if (gy + j == H - 1) sv = fma4s(4.f, v[j + 4], fma4s(3.f, v[j + 3], fma4s(2.f, v[j + 2], add4(sv, v[j + 1]))));
        const float4 e = *reinterpret_cast<const float4*>(&S.es[s][2 * rs + j + R9][4 * q + R9]);
        const float4 tt = *reinterpret_cast<const float4*>(&S.ta[s][2 * rs + j + R9][4 * q + R9]);
        float4 r;
        if (TYPE == 0) {
          r = make_float4(2.f * (e.x - tt.x) * INV81 * sv.x, 2.f * (e.y - tt.y) * INV81 * sv.y,
                          2.f * (e.z - tt.z) * INV81 * sv.z, 2.f * (e.w - tt.w) * INV81 * sv.w);
        } else {
          r = make_float4(sgnf(e.x - tt.x) * INV81 * sv.x, sgnf(e.y - tt.y) * INV81 * sv.y,
                          sgnf(e.z - tt.z) * INV81 * sv.z, sgnf(e.w - tt.w) * INV81 * sv.w);
        }
        const int64_t off = ((int64_t)c.n * H + gy + j) * W + gx;
        *reinterpret_cast<float4*>(gi + off) = r;
        const float4 lo = make_float4(f[j].x * INV81, f[j].y * INV81, f[j].z * INV81, f[j].w * INV81);
        *reinterpret_cast<float4*>(out + off) = lo;
        if (mask != nullptr) {
          const float4 m = ldg4(mask + off);
          mnum += (double)((m.x * lo.x + m.y * lo.y) + (m.z * lo.z + m.w * lo.w));
          mden += (double)((m.x + m.y) + (m.z + m.w));
        }
      }
    }
    __syncthreads();
  }
  if (mask != nullptr) finish_masked_sums(mnum, mden, partials, ticket, sums2);
}

static int sm_count() {
  static int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

template <typename K>
static bool set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess;
}

// Returns 1 when the call was handled by the TMA kernels, 0 when the caller should use the
// shared-memory tile kernels instead (C != 1, unaligned rows, tiny images), < 0 is not used.
int box9_tma_fwd(const float* es, const float* ta, float* out, int64_t B, int64_t C, int64_t H, int64_t W, int type,
                 cudaStream_t st) {
  if (g_disable_tma || g_force_generic || C != 1 || B < 1 || W % 4 || H < 9 || W < 9) return 0;
  if ((reinterpret_cast<uintptr_t>(out) & 15) || B * cdiv(H, TT_H) * cdiv(W, TT_W) > INT32_MAX) return 0;
  CUtensorMap m_es, m_ta;
  if (!make_plane_tensor_map(&m_es, es, B, H, W, TB_W, TB_H) || !make_plane_tensor_map(&m_ta, ta, B, H, W, TB_W, TB_H))
    return 0;
  const int tiles_x = (int)cdiv(W, TT_W), tiles_y = (int)cdiv(H, TT_H), ntiles = (int)(B * tiles_x * tiles_y);
  const int grid = std::min(ntiles, sm_count() * 3);
  if (!set_smem(type == 0 ? photo_fwd_box9_tma<0> : photo_fwd_box9_tma<1>, (sizeof(FwdSmem) + 128))) return 0;
  if (type == 0)
    photo_fwd_box9_tma<0><<<grid, 256, (sizeof(FwdSmem) + 128), st>>>(m_es, m_ta, out, (int)H, (int)W, tiles_x, tiles_y, ntiles);
  else
    photo_fwd_box9_tma<1><<<grid, 256, (sizeof(FwdSmem) + 128), st>>>(m_es, m_ta, out, (int)H, (int)W, tiles_x, tiles_y, ntiles);
  count_launch();
  return 1;
}

int box9_tma_bwd(const float* es, const float* ta, const float* go, float* gi, int64_t B, int64_t C, int64_t H,
                 int64_t W, int type, cudaStream_t st) {
  if (g_disable_tma || g_force_generic || C != 1 || B < 1 || W % 4 || H < 9 || W < 9) return 0;
  if (((reinterpret_cast<uintptr_t>(es) | reinterpret_cast<uintptr_t>(ta) | reinterpret_cast<uintptr_t>(gi)) & 15) ||
      B * cdiv(H, TT_H) * cdiv(W, TT_W) > INT32_MAX)
    return 0;
  CUtensorMap m_go;
  if (!make_plane_tensor_map(&m_go, go, B, H, W, TB_W, TB_H)) return 0;
  const int tiles_x = (int)cdiv(W, TT_W), tiles_y = (int)cdiv(H, TT_H), ntiles = (int)(B * tiles_x * tiles_y);
  const int grid = std::min(ntiles, sm_count() * 3);
  if (!set_smem(type == 0 ? photo_bwd_box9_tma<0> : photo_bwd_box9_tma<1>, (sizeof(BwdSmem) + 128))) return 0;
  if (type == 0)
    photo_bwd_box9_tma<0><<<grid, 256, (sizeof(BwdSmem) + 128), st>>>(m_go, es, ta, gi, (int)H, (int)W, tiles_x, tiles_y, ntiles);
  else
    photo_bwd_box9_tma<1><<<grid, 256, (sizeof(BwdSmem) + 128), st>>>(m_go, es, ta, gi, (int)H, (int)W, tiles_x, tiles_y, ntiles);
  count_launch();
  return 1;
}

int box9_tma_fwd_bwd_masked(const float* es, const float* ta, const float* go, const float* mask, float* out, float* gi,
                            float* sums2, int64_t B, int64_t C, int64_t H, int64_t W, int type, cudaStream_t st) {
  if (g_disable_tma || g_force_generic || C != 1 || B < 1 || W % 4 || H < 9 || W < 9) return 0;
  if (((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gi) | reinterpret_cast<uintptr_t>(mask)) & 15) ||
      B * cdiv(H, TT_H) * cdiv(W, TT_W) > INT32_MAX)
    return 0;
  CUtensorMap m_es, m_ta, m_go;
  if (!make_plane_tensor_map(&m_es, es, B, H, W, TB_W, TB_H) || !make_plane_tensor_map(&m_ta, ta, B, H, W, TB_W, TB_H) ||
      !make_plane_tensor_map(&m_go, go, B, H, W, TB_W, TB_H))
    return 0;
  const int tiles_x = (int)cdiv(W, TT_W), tiles_y = (int)cdiv(H, TT_H), ntiles = (int)(B * tiles_x * tiles_y);
  const int grid = std::min(ntiles, sm_count() * 2);
  const size_t smem = sizeof(FbSmem) + 128;
  if (!set_smem(type == 0 ? photo_fwd_bwd_box9_tma<0> : photo_fwd_bwd_box9_tma<1>, smem)) return 0;
  MsSlot ms = {nullptr, nullptr, nullptr};
  if (mask && !ms_acquire((size_t)grid, st, &ms)) return 0;
  double* partials = ms.partials;
  unsigned* ticket = ms.ticket;
  if (type == 0)
    photo_fwd_bwd_box9_tma<0><<<grid, 256, smem, st>>>(m_es, m_ta, m_go, out, gi, (int)H, (int)W, tiles_x, tiles_y, ntiles, mask,
                                                       partials, ticket, sums2);
  else
    photo_fwd_bwd_box9_tma<1><<<grid, 256, smem, st>>>(m_es, m_ta, m_go, out, gi, (int)H, (int)W, tiles_x, tiles_y, ntiles, mask,
                                                       partials, ticket, sums2);
  ms_release(&ms, st);
  count_launch();
  return 1;
}

int box9_tma_fwd_bwd(const float* es, const float* ta, const float* go, float* out, float* gi, int64_t B, int64_t C,
                     int64_t H, int64_t W, int type, cudaStream_t st) {
  return box9_tma_fwd_bwd_masked(es, ta, go, nullptr, out, gi, nullptr, B, C, H, W, type, st);
}

}  // namespace ctd
