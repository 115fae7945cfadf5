// PhotometricLoss census_mse / census_sad, block size 9, C = 1, fp32: pair-symmetric kernels for sm_100a.
//
// The soft census term of a pixel pair is (anti)symmetric: with des = es[q] - es[p], dta = ta[q] - ta[p],
//     dd(p,q) = des * rsqrt(des^2 + eps) - dta * rsqrt(dta^2 + eps) = -dd(q,p),
// so the forward's psi(dd) (|.| or square) is the same number for "q as a tap of p" and "p as a tap of q", and
// the backward's per-pair gradient flips sign.  The gather kernels of photometric.cu evaluate every pair
// twice -- 162 MUFU.RSQ per pixel, and the XU pipe (16 lanes/clk/SM) is what bounds them.  Here every pair of
// real pixels is evaluated ONCE, from the pixel that comes first in raster order, over the 40 offsets
// (0,1..4), (1..4,-4..4), and credited to both ends.
//
// Work decomposition (no shared memory, no block barrier): a warp owns a strip of 128 columns x RS rows; lane
// l owns four adjacent columns ("quad") and walks down the rows.  Contributions to rows below stay in the
// thread (a five-row ring of accumulators); contributions to the neighbouring quads travel by warp shuffle
// once per row and offset row.  The first / last lane of a strip and the four rows above it are halo: they
// compute, but another warp writes those pixels.  When the image width is not a multiple of 120 the
// leftover columns form a narrow strip, and several row ranges of it are packed side by side into one warp.
// Image tiles are read straight from global memory with 128-bit loads (every row is re-read from L1).
//
// Clamped taps: a pixel at least four pixels away from the image border has 80 distinct real neighbours, so
// the pair sum is its whole window.  The 4-pixel band along the border (3% of a 480x640 image), whose
// windows are replicate-clamped, is written by a small direct kernel instead; the pair kernel only treats
// pixels outside the image as absent (0/1 masks folded into the accumulating FMAs).
//
// Reference semantics: torchext/ext/ext.h:201-266 (forward), :268-344 (backward).
#include <algorithm>

#include "ctd_common.cuh"

namespace ctd {

extern int g_force_generic;
// ctd_set_option("census_pairs", v): 0 = always the gather kernels of photometric.cu, 1 = always the pair kernel,
// 2 (default) = pair kernel from CP_MIN_PIXELS pixels up.  Measured on B200 (tools/bench_ops.py): batch 64 x
// 480x640 forward 682 us (pairs) vs 755 us (gather); at batch 8 the 2.5 M pixels yield too few strips to hide the
// MUFU latency and the two are level (117 vs 113 us), so small calls stay on the gather kernel.
int g_census_pairs = 2;
constexpr int64_t CP_MIN_PIXELS = (int64_t)6 << 20;

namespace {

constexpr int R9 = 4;
constexpr int VQ = 30;               // output quads of a full strip (lanes 1..30); 120 columns
constexpr float INV81 = 1.0f / 81.0f;

struct Geom {
  int H, W, RS, nys;       // rows per strip, strips per image
  int nfull;               // full 120-column strips
  int rem_q;               // output quads of the leftover strip (0 = none)
  int lanes_per, pack;     // leftover strip: lanes per row range (rem_q + 2), row ranges per warp
  int64_t full_tasks, tasks;
};

static Geom make_geom(int64_t B, int H, int W) {
  Geom g;
  g.H = H;
  g.W = W;
  g.nfull = W / (4 * VQ);
  g.rem_q = (W - g.nfull * 4 * VQ) / 4;
  g.lanes_per = g.rem_q + 2;
  g.pack = g.rem_q ? 32 / g.lanes_per : 1;
  // rows per strip: four extra rows are walked per strip, but the GPU wants a few thousand warps
  int rs = 64;
  for (; rs > 16; rs -= 8) {
    const int64_t nys = cdiv(H, rs);
    const int64_t t = B * nys * g.nfull + (g.rem_q ? cdiv(B * nys, g.pack) : 0);
    if (2 * t >= 148 * 12) break;  // two warps per strip (the offset rows are split between them)
  }
  g.RS = rs;
  g.nys = (int)cdiv(H, rs);
  g.full_tasks = B * g.nys * g.nfull;
  g.tasks = g.full_tasks + (g.rem_q ? cdiv(B * g.nys, g.pack) : 0);
  return g;
}

struct Lane {
  int64_t plane;   // image offset in pixels
  int xq, ys;      // first column of the quad (may lie outside the image), first output row of the strip
  int xl, xc, xr;  // clamped quad columns for the loads (left neighbour, own, right neighbour)
  float selfm, lm, rm;
  bool writes;     // this lane's quad is an output quad of the strip and lies inside the image
};

__device__ __forceinline__ Lane decode(const Geom& g, int64_t wt, int lane, int64_t B) {
  Lane L;
  int ql, nvalid, strip;
  int64_t slot;
  bool active = true;
  if (wt < g.full_tasks) {
    strip = (int)(wt % g.nfull);
    slot = wt / g.nfull;
    ql = lane;
    nvalid = VQ;
  } else {
    const int sub = lane / g.lanes_per;
    ql = lane - sub * g.lanes_per;
    slot = (wt - g.full_tasks) * g.pack + sub;
    active = sub < g.pack && slot < B * g.nys;
    if (!active) slot = 0;
    strip = g.nfull;
    nvalid = g.rem_q;
  }
  const int64_t n = slot / g.nys;
  L.plane = n * g.H * (int64_t)g.W;
  L.ys = (int)(slot % g.nys) * g.RS;
  L.xq = strip * 4 * VQ - 4 + 4 * ql;
  auto inside = [&](int x) { return x >= 0 && x + 3 < g.W; };
  L.selfm = inside(L.xq) ? 1.f : 0.f;
  L.lm = inside(L.xq - 4) ? 1.f : 0.f;
  L.rm = inside(L.xq + 4) ? 1.f : 0.f;
  L.xl = clampi(L.xq - 4, 0, g.W - 4);
  L.xc = clampi(L.xq, 0, g.W - 4);
  L.xr = clampi(L.xq + 4, 0, g.W - 4);
  // the border band (first / last quad of the image, first / last four rows) belongs to the band kernel
  L.writes = active && ql >= 1 && ql <= nvalid && L.xq >= 4 && L.xq + 4 <= g.W - 4;
  return L;
}

__device__ __forceinline__ void unpack4(float* d, const float4 v) {
  d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
}

// psi(dd) of the pair, dd = des * rsqrt(des^2+eps) - dta * rsqrt(dta^2+eps) (= 2 (h(des) - h(dta)))
template <int TYPE>
__device__ __forceinline__ float pair_fwd(float e_tap, float e_ctr, float t_tap, float t_ctr, float eps) {
  const float des = e_tap - e_ctr, dta = t_tap - t_ctr;
  const float r1 = rsqrt_approx(fmaf(des, des, eps));
  const float r2 = rsqrt_approx(fmaf(dta, dta, eps));
  const float dd = fmaf(des, r1, -(dta * r2));
  return TYPE == 2 ? dd * dd : fabsf(dd);
}

// Four pairs at a time (the quad's pixels k = 0..3 against taps k + j0), in three stages so that a thread can
// keep two batches in flight: the eight MUFU.RSQ of batch n+1 are issued before batch n's results are consumed.
// Only ~2 warps per scheduler are resident (the strips are long), so the XU pipe is kept busy by instruction-
// level parallelism, not by warp switching.
struct Batch {
  float des[4], dta[4], r1[4], r2[4];
};
__device__ __forceinline__ void stage_a(Batch& b, const float* e, const float* t, int j0, const float* ec, const float* tc,
                                        float eps) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    b.des[k] = e[k + j0] - ec[k];
    b.dta[k] = t[k + j0] - tc[k];
    b.r1[k] = fmaf(b.des[k], b.des[k], eps);
    b.r2[k] = fmaf(b.dta[k], b.dta[k], eps);
  }
}
__device__ __forceinline__ void stage_b(Batch& b) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    b.r1[k] = rsqrt_approx(b.r1[k]);
    b.r2[k] = rsqrt_approx(b.r2[k]);
  }
}
template <int TYPE>
__device__ __forceinline__ void stage_c(const Batch& b, float* v) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float dd = fmaf(b.des[k], b.r1[k], -(b.dta[k] * b.r2[k]));
    v[k] = TYPE == 2 ? dd * dd : fabsf(dd);
  }
}

// 128-bit reduction into global memory (fire and forget).  Every output element receives exactly two of
// these onto a zeroed buffer, so the result does not depend on their order.
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// HALF 0: offset rows 0..2 (88 pairs per quad and row), HALF 1: offset rows 3..4 (72 pairs).  The two warps of a
// strip run independently and add their partial sums into `out` (zeroed by the launcher): twice the resident
// warps for the same arithmetic, which is what keeps the XU pipe fed.
template <int TYPE, int HALF>
__device__ __forceinline__ void census_fwd_pairs_body(const float* __restrict__ es, const float* __restrict__ ta,
                                                      float* __restrict__ out, int64_t B, const Geom& g, float eps, int64_t wt) {
  constexpr int DY0 = HALF == 0 ? 0 : 3, DY1 = HALF == 0 ? 2 : 4;
  const int lane = threadIdx.x & 31;
  const Lane L = decode(g, wt, lane, B);
  const int H = g.H, W = g.W;
  const float* ep = es + L.plane;
  const float* tp = ta + L.plane;
  float acc[5][4];  // acc[d][k]: pixel (row r + d, column xq + k) while row r is the base row
#pragma unroll
  for (int d = 0; d < 5; ++d)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[d][k] = 0.f;
  const float scale = (TYPE == 2 ? 0.25f : 0.5f) * INV81;
  auto load12 = [&](float* e, float* t, int rr) {
    const int64_t row = (int64_t)clampi(rr, 0, H - 1) * W;
    unpack4(e, ldg4(ep + row + L.xl));
    unpack4(e + 4, ldg4(ep + row + L.xc));
    unpack4(e + 8, ldg4(ep + row + L.xr));
    unpack4(t, ldg4(tp + row + L.xl));
    unpack4(t + 4, ldg4(tp + row + L.xc));
    unpack4(t + 8, ldg4(tp + row + L.xr));
  };
#pragma unroll 1
  for (int i = 0; i < g.RS + R9; ++i) {
    const int r = L.ys - R9 + i;  // base row
    const float mp = (r >= 0 && r < H) ? L.selfm : 0.f;  // this thread's base pixels exist
    float ec[4], tc[4];
    {
      const int64_t row = (int64_t)clampi(r, 0, H - 1) * W;
      unpack4(ec, ldg4(ep + row + L.xc));
      unpack4(tc, ldg4(tp + row + L.xc));
    }
    Batch bt[2];
#pragma unroll
    for (int dy = DY0; dy <= DY1; ++dy) {
      // offsets (0, 1..4) on the base row, (dy, -4..4) on the rows below
      float e[12], t[12];  // tile row r + dy: columns xq-4 .. xq+7
      load12(e, t, r + dy);
      const float rp = (r + dy >= 0 && r + dy < H) ? 1.f : 0.f;
      const float mo_l = rp * L.lm, mo_r = rp * L.rm;
      float sl[4] = {0.f, 0.f, 0.f, 0.f}, sr[4] = {0.f, 0.f, 0.f, 0.f};
      const int dx0 = dy == 0 ? 1 : -4;
      stage_a(bt[0], e, t, dx0 + 4, ec, tc, eps);
      stage_b(bt[0]);
#pragma unroll
      for (int dx = dx0; dx <= 4; ++dx) {
        const int cur = (dx - dx0) & 1;
        if (dx < 4) {
          stage_a(bt[cur ^ 1], e, t, dx + 1 + 4, ec, tc, eps);
          stage_b(bt[cur ^ 1]);
        }
        float v[4];
        stage_c<TYPE>(bt[cur], v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = k + dx;  // partner column relative to the quad: -4 .. 7
          if (j < 0) {
            acc[0][k] = fmaf(v[k], mo_l, acc[0][k]);
            sl[j + 4] = fmaf(v[k], mp, sl[j + 4]);
          } else if (j < 4) {
            acc[0][k] = fmaf(v[k], rp, acc[0][k]);
            acc[dy][j] = fmaf(v[k], mp, acc[dy][j]);
          } else {
            acc[0][k] = fmaf(v[k], mo_r, acc[0][k]);
            sr[j - 4] = fmaf(v[k], mp, sr[j - 4]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc[dy][k] += __shfl_up_sync(0xffffffffu, sr[k], 1);  // the left neighbour's right-going share
        if (dy > 0) acc[dy][k] += __shfl_down_sync(0xffffffffu, sl[k], 1);  // the right neighbour's left-going share
      }
    }
    // row r has now met all of its neighbours in this warp's offset rows
    if (i >= R9 && L.writes && r >= R9 && r < H - R9)
      red_add4(out + L.plane + (int64_t)r * W + L.xq, acc[0][0] * scale, acc[0][1] * scale, acc[0][2] * scale, acc[0][3] * scale);
#pragma unroll
    for (int d = 0; d < 4; ++d)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[d][k] = acc[d + 1][k];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[4][k] = 0.f;
  }
}

template <int TYPE>
__global__ void __launch_bounds__(128)
census_fwd_pairs(const float* __restrict__ es, const float* __restrict__ ta, float* __restrict__ out, int64_t B,
                 const Geom g, float eps) {
  // even blocks take offset rows 0..2, odd blocks rows 3..4 of the same four strips
  const int64_t wt = (int64_t)(blockIdx.x >> 1) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wt >= g.tasks) return;  // warp-uniform
  if (blockIdx.x & 1) census_fwd_pairs_body<TYPE, 1>(es, ta, out, B, g, eps, wt);
  else census_fwd_pairs_body<TYPE, 0>(es, ta, out, B, g, eps, wt);
}

// the 4-pixel band along the image border: replicate-clamped windows, one thread per pixel
template <int TYPE>
__global__ void __launch_bounds__(128)
census_fwd_band(const float* __restrict__ es, const float* __restrict__ ta, float* __restrict__ out, int64_t B, int H,
                int W, float eps) {
  const int per_image = 8 * W + 8 * (H - 8);
  const int64_t total = B * per_image;
  const float scale = (TYPE == 2 ? 0.25f : 0.5f) * INV81;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx / per_image;
    int j = (int)(idx - n * per_image), x, y;
    if (j < 8 * W) {
      y = j / W;
      x = j - y * W;
      if (y >= 4) y += H - 8;
    } else {
      j -= 8 * W;
      y = 4 + j / 8;
      x = j % 8;
      if (x >= 4) x += W - 8;
    }
    const float* ep = es + n * H * (int64_t)W;
    const float* tp = ta + n * H * (int64_t)W;
    const float ec = __ldg(ep + (int64_t)y * W + x), tc = __ldg(tp + (int64_t)y * W + x);
    float acc = 0.f;
    for (int dy = -R9; dy <= R9; ++dy) {
      const int64_t row = (int64_t)clampi(y + dy, 0, H - 1) * W;
#pragma unroll
      for (int dx = -R9; dx <= R9; ++dx) {
        const int c = clampi(x + dx, 0, W - 1);
        acc += pair_fwd<TYPE>(__ldg(ep + row + c), ec, __ldg(tp + row + c), tc, eps);
      }
    }
    out[n * H * (int64_t)W + (int64_t)y * W + x] = acc * scale;
  }
}

}  // namespace

// Returns 1 when the call was handled here, 0 when the caller should use the gather kernels.
int census_pairs_fwd(const float* es, const float* ta, float* out, int64_t B, int64_t C, int64_t H, int64_t W, int type,
                     float eps, cudaStream_t st) {
  if (!g_census_pairs || (g_census_pairs == 2 && B * H * W < CP_MIN_PIXELS) || g_force_generic || C != 1 || B < 1 || W % 4 || H < 16 || W < 16 || H * W >= (int64_t)1 << 30) return 0;
  if ((reinterpret_cast<uintptr_t>(es) | reinterpret_cast<uintptr_t>(ta) | reinterpret_cast<uintptr_t>(out)) & 15) return 0;
  const Geom g = make_geom(B, (int)H, (int)W);
  if (g.tasks > (int64_t)INT32_MAX) return 0;
  const unsigned grid = 2 * (unsigned)cdiv(g.tasks, 4);
  if (cudaMemsetAsync(out, 0, sizeof(float) * (size_t)(B * H * W), st) != cudaSuccess) {  // the pair kernel accumulates
    cudaGetLastError();
    return 0;
  }
  const unsigned bgrid = (unsigned)std::min<int64_t>(cdiv(B * (8 * W + 8 * (H - 8)), 128), 148 * 16);
  if (type == 2) {
    census_fwd_pairs<2><<<grid, 128, 0, st>>>(es, ta, out, B, g, eps);
    census_fwd_band<2><<<bgrid, 128, 0, st>>>(es, ta, out, B, (int)H, (int)W, eps);
  } else {
    census_fwd_pairs<3><<<grid, 128, 0, st>>>(es, ta, out, B, g, eps);
    census_fwd_band<3><<<bgrid, 128, 0, st>>>(es, ta, out, B, (int)H, (int)W, eps);
  }
  count_launch(2);
  return 1;
}

}  // namespace ctd
