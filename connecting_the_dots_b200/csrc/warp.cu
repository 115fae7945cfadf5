// Disparity warp of the reference pattern: the grid_sample call in front of PhotometricLoss
// (model/networks.py:362-371, RectifiedPatternSimilarityLoss.tforward), SURVEY section 8(f) rank 1.
//
//   gx = 2 * ((u - disp) / (W - 1) - 0.5),  gy = 2 * (v / (H - 1) - 0.5)          (networks.py:364-369)
//   pattern_proj = grid_sample(pattern, (gx, gy), mode=bilinear, padding_mode='border', align_corners=False)
//
// i.e. the grid is normalised with the (W-1) convention but sampled with torch's default align_corners=False, so the
// sample position is x * Wp/(W-1) - 0.5: a slight zoom that also makes the row interpolation bilinear.  The kernels
// restate ATen's grid_sampler_2d (GridSampler.cuh: grid_sampler_compute_source_index, clip_coordinates, the nw/ne/sw/se
// accumulation order) with the same fp32 operation sequence -- including the fused multiply-add nvcc contracts in
// ATen's `(coord + 1) * size - 1` -- because a one-ulp difference in the sample position (~6e-5 px at W = 640) moves
// the interpolated value by more than the 1e-5 parity tolerance.  Parity is checked against torch's own CUDA
// grid_sample and its autograd (tests/test_gpu_ops.py).
//
// Forward: 8 B/px (disp in, pattern_proj out; the pattern is L2-resident).  Backward w.r.t. disp: 12 B/px.
#include <algorithm>
#include <cmath>

#include "ctd_common.cuh"

namespace ctd {

struct SamplePos {
  float ix, iy;       // clipped source position
  float mult_x;       // d ix / d gx  (0 where the border clip is active)
  int x0, y0;         // north-west corner
};

// grid value -> source index, align_corners = false, padding_mode = border
__device__ __forceinline__ float unnormalize(float g, int size) { return fmaf(g + 1.f, (float)size, -1.f) / 2.f; }

// inv_w = 1.f / (W - 1), inv_h = 1.f / (H - 1), computed once on the host (IEEE single division, like ATen's)
__device__ __forceinline__ SamplePos sample_pos(float disp, int u, int v, float inv_w, float inv_h, int Hp, int Wp) {
  SamplePos s;
  // the reference's normalisation in torch's CUDA arithmetic: a tensor divided by a Python scalar is multiplied by
  // the fp32 reciprocal of the scalar (ATen BinaryDivTrueKernel), the other steps are single fp32 operations
  const float gx = 2.f * __fsub_rn(__fmul_rn(__fsub_rn((float)u, disp), inv_w), 0.5f);
  const float gy = 2.f * __fsub_rn(__fmul_rn((float)v, inv_h), 0.5f);
  float ix = unnormalize(gx, Wp), iy = unnormalize(gy, Hp);
  s.mult_x = (float)Wp / 2.f;
  // clip_coordinates: min(size - 1, max(ix, 0)); gradient 0 outside [0, size - 1]
  if (ix <= 0.f) {
    ix = 0.f;
    s.mult_x = 0.f;
  } else if (ix >= (float)(Wp - 1)) {
    ix = (float)(Wp - 1);
    s.mult_x = 0.f;
  }
  iy = fminf((float)(Hp - 1), fmaxf(iy, 0.f));
  s.ix = ix;
  s.iy = iy;
  s.x0 = (int)floorf(ix);
  s.y0 = (int)floorf(iy);
  return s;
}

// grid: (columns / 128, rows / WP_ROWS, images): no integer division in the index math, WP_ROWS independent pixels
// per thread so the dependent gather chains overlap
#ifndef CTD_WP_T
#define CTD_WP_T 128
#endif
#ifndef CTD_WP_ROWS
#define CTD_WP_ROWS 2  // measured at batch 8 (fwd / bwd us): 2 rows 11.5 / 16.0, 4 rows 11.7 / 17.6, 8 rows 12.9 / 21.7; 256 threads x 4: 12.1 / 19.0
#endif
constexpr int WP_T = CTD_WP_T, WP_ROWS = CTD_WP_ROWS;

__global__ void __launch_bounds__(WP_T)
warp_fwd_kernel(const float* __restrict__ pattern, const float* __restrict__ disp, float* __restrict__ out, int Bp, int Hp,
                int Wp, int H, int W, float inv_w, float inv_h) {
  const int u = blockIdx.x * WP_T + threadIdx.x;
  if (u >= W) return;
  const int64_t b = blockIdx.z;
  const float* p = pattern + (Bp == 1 ? 0 : b) * (int64_t)Hp * Wp;
#pragma unroll
  for (int rr = 0; rr < WP_ROWS; ++rr) {
  const int v = blockIdx.y * WP_ROWS + rr;
  if (v >= H) break;
  const int64_t i = (b * H + v) * W + u;
  const SamplePos s = sample_pos(__ldg(disp + i), u, v, inv_w, inv_h, Hp, Wp);
  const int x1 = s.x0 + 1, y1 = s.y0 + 1;
  const float nw = (x1 - s.ix) * (y1 - s.iy), ne = (s.ix - s.x0) * (y1 - s.iy);
  const float sw = (x1 - s.ix) * (s.iy - s.y0), se = (s.ix - s.x0) * (s.iy - s.y0);
  const bool xin = x1 < Wp, yin = y1 < Hp;  // x0, y0 are always inside after the clip
  const float* r0 = p + s.y0 * Wp + s.x0;
  float acc = 0.f;
  acc = fmaf(__ldg(r0), nw, acc);
  if (xin) acc = fmaf(__ldg(r0 + 1), ne, acc);
  if (yin) acc = fmaf(__ldg(r0 + Wp), sw, acc);
  if (xin && yin) acc = fmaf(__ldg(r0 + Wp + 1), se, acc);
  out[i] = acc;
  }
}

// grad_disp = - 2 / (W - 1) * (Wp / 2) * clip * d(sample)/d(ix) * grad_out
__global__ void __launch_bounds__(WP_T)
warp_bwd_kernel(const float* __restrict__ pattern, const float* __restrict__ disp, const float* __restrict__ go,
                float* __restrict__ gd, int Bp, int Hp, int Wp, int H, int W, float inv_w, float inv_h) {
  const int u = blockIdx.x * WP_T + threadIdx.x;
  if (u >= W) return;
  const int64_t b = blockIdx.z;
  const float* p = pattern + (Bp == 1 ? 0 : b) * (int64_t)Hp * Wp;
#pragma unroll
  for (int rr = 0; rr < WP_ROWS; ++rr) {
  const int v = blockIdx.y * WP_ROWS + rr;
  if (v >= H) break;
  const int64_t i = (b * H + v) * W + u;
  const SamplePos s = sample_pos(__ldg(disp + i), u, v, inv_w, inv_h, Hp, Wp);
  const int x1 = s.x0 + 1, y1 = s.y0 + 1;
  const bool xin = x1 < Wp, yin = y1 < Hp;
  const float g = __ldg(go + i);
  const float* r0 = p + s.y0 * Wp + s.x0;
  // ATen's order: gix -= nw_val * (iy_se - iy) * g; gix += ne_val * (iy_sw - iy) * g; gix -= sw_val * (iy - iy_ne) * g;
  // gix += se_val * (iy - iy_nw) * g
  float gix = 0.f;
  gix -= __ldg(r0) * (y1 - s.iy) * g;
  if (xin) gix += __ldg(r0 + 1) * (y1 - s.iy) * g;
  if (yin) gix -= __ldg(r0 + Wp) * (s.iy - s.y0) * g;
  if (xin && yin) gix += __ldg(r0 + Wp + 1) * (s.iy - s.y0) * g;
  const float ggx = s.mult_x * gix;          // gradient w.r.t. the grid's x
  gd[i] = -__fmul_rn(2.f * ggx, inv_w);      // autograd through 2 * ((u - disp) / (W - 1) - 0.5)
  }
}

}  // namespace ctd

using namespace ctd;

static int warp_check(const void* pattern, const void* disp, const void* out, int64_t B, int64_t Bp, int64_t Hp, int64_t Wp,
                      int64_t H, int64_t W) {
  CTD_REQUIRE(B >= 0 && H >= 0 && W >= 0, "warp_pattern: negative size");
  CTD_REQUIRE(Bp >= 1 && Hp >= 1 && Wp >= 1, "warp_pattern: empty pattern");
  CTD_REQUIRE(Bp == 1 || Bp == B, "warp_pattern: pattern batch %lld must be 1 or %lld", (long long)Bp, (long long)B);
  CTD_REQUIRE(H <= INT32_MAX && W <= INT32_MAX && Hp <= INT32_MAX && Wp <= INT32_MAX, "warp_pattern: dimension too large");
  if (B * H * W > 0) CTD_REQUIRE(pattern && disp && out, "warp_pattern: null pointer");
  return CTD_OK;
}

static inline float inv_extent(int64_t n) { return n > 1 ? 1.f / (float)(n - 1) : INFINITY; }  // 1/0 = inf like torch

CTD_API int ctd_warp_pattern_fwd_f32(const float* pattern, const float* disp, float* out, int64_t B, int64_t Bp, int64_t Hp,
                                        int64_t Wp, int64_t H, int64_t W, ctd_stream_t stream) {
  if (int rc = warp_check(pattern, disp, out, B, Bp, Hp, Wp, H, W)) return rc;
  if (B * H * W == 0) return CTD_OK;
  CTD_REQUIRE(H <= 65535 * (int64_t)WP_ROWS && Hp * Wp < ((int64_t)1 << 31), "warp_pattern: image too large");
  for (int64_t b0 = 0; b0 < B; b0 += 65535) {
    const int64_t nb = std::min<int64_t>(65535, B - b0);
    const dim3 grid((unsigned)cdiv(W, WP_T), (unsigned)cdiv(H, WP_ROWS), (unsigned)nb);
    warp_fwd_kernel<<<grid, WP_T, 0, as_stream(stream)>>>(pattern + (Bp == 1 ? 0 : b0) * Hp * Wp, disp + b0 * H * W, out + b0 * H * W,
                                                         (int)Bp, (int)Hp, (int)Wp, (int)H, (int)W, inv_extent(W), inv_extent(H));
    count_launch();
  }
  return check_launch("warp_pattern_fwd");
}

CTD_API int ctd_warp_pattern_bwd_f32(const float* pattern, const float* disp, const float* grad_out, float* grad_disp, int64_t B,
                                        int64_t Bp, int64_t Hp, int64_t Wp, int64_t H, int64_t W, ctd_stream_t stream) {
  if (int rc = warp_check(pattern, disp, grad_disp, B, Bp, Hp, Wp, H, W)) return rc;
  if (B * H * W == 0) return CTD_OK;
  CTD_REQUIRE(grad_out, "warp_pattern_bwd: null grad_out");
  CTD_REQUIRE(H <= 65535 * (int64_t)WP_ROWS && Hp * Wp < ((int64_t)1 << 31), "warp_pattern: image too large");
  for (int64_t b0 = 0; b0 < B; b0 += 65535) {
    const int64_t nb = std::min<int64_t>(65535, B - b0);
    const dim3 grid((unsigned)cdiv(W, WP_T), (unsigned)cdiv(H, WP_ROWS), (unsigned)nb);
    warp_bwd_kernel<<<grid, WP_T, 0, as_stream(stream)>>>(pattern + (Bp == 1 ? 0 : b0) * Hp * Wp, disp + b0 * H * W, grad_out + b0 * H * W,
                                                         grad_disp + b0 * H * W, (int)Bp, (int)Hp, (int)Wp, (int)H, (int)W, inv_extent(W),
                                                         inv_extent(H));
    count_launch();
  }
  return check_launch("warp_pattern_bwd");
}
