// Disparity (smoothness / edge) loss of the stage-1 trainer: model/networks.py:380-412 (DisparityLoss.tforward)
// over the 5x5 Sobel filter of networks.py:537-565 (SobelFilter, replicate padding).  SURVEY section 8(f) rank 4.
//
//   gx, gy = 5x5 cross-correlations of the replicate-padded disparity with kx / 240 and its transpose
//   m      = sqrt(gx^2 + gy^2 + 1e-8)
//   edge given:  pdf = (1-e)/b0 exp(-m/b0) + e/b1 exp(-m/b1);   loss = mean(-log(max(pdf, 1e-4)))      (networks.py:399-402)
//   edge absent: loss = mean(clamp(m, 0, 1))                                                            (networks.py:406-409)
//
// In torch that is a pad, two single-channel library convolutions and ~15 elementwise kernels, and as many again in
// autograd.  Here ONE kernel produces the loss sum and both gradients: a CTA owns a 32x32 output tile, stages the
// 40x40 disparity halo (replicate clamp), evaluates gx, gy, the loss term and dL/dgx, dL/dgy on the 36x36 positions
// the tile's gradient needs (zero outside the image), and gathers grad_disp as the adjoint correlation from shared
// memory -- no atomics, no intermediate tensors.  The adjoint of the replicate padding folds the gradient of the pad
// cells onto the border pixel they copy.  Algorithmic bytes: disp + edge in, grad_disp + grad_edge out = 16 B/px.
#include <algorithm>
#include <cmath>

#include "ctd_common.cuh"

namespace ctd {

constexpr int DL_T = 32;            // output tile
constexpr int DL_G = DL_T + 4;      // positions whose Sobel response the tile's gradient needs
constexpr int DL_D = DL_T + 8;      // disparity halo
constexpr float DL_B0 = 0.0503428816795f, DL_B1 = 1.07274045944f;  // networks.py:390-391

__constant__ float c_kx[25] = {-5.f / 240, -4.f / 240,  0.f, 4.f / 240,  5.f / 240,
                               -8.f / 240, -10.f / 240, 0.f, 10.f / 240, 8.f / 240,
                               -10.f / 240, -20.f / 240, 0.f, 20.f / 240, 10.f / 240,
                               -8.f / 240, -10.f / 240, 0.f, 10.f / 240, 8.f / 240,
                               -5.f / 240, -4.f / 240,  0.f, 4.f / 240,  5.f / 240};

template <bool EDGE>
__global__ void __launch_bounds__(256)
disparity_loss_kernel(const float* __restrict__ disp, const float* __restrict__ edge, float* __restrict__ gdisp,
                      float* __restrict__ gedge, int H, int W, int tiles_x, int tiles_y, int ntiles, float scale,
                      double* __restrict__ partials, unsigned* __restrict__ ticket, float* __restrict__ sums2) {
  __shared__ float D[DL_D][DL_D + 1];
  // dL/dgx, dL/dgy (times scale) at the DL_G x DL_G positions, stored with a two-cell border of zeros (and zero outside
  // the image) so that the adjoint gather needs no bounds tests: position (r, c) lives at [r + 2][c + 2]
  __shared__ float Ax[DL_G + 4][DL_G + 5], Ay[DL_G + 4][DL_G + 5];
  const int tid = threadIdx.x;
  const int64_t plane = (int64_t)H * W;
  double acc = 0.0, cnt = 0.0;
  for (int i = tid; i < (DL_G + 4) * (DL_G + 4); i += 256) {  // the zero border is written once
    Ax[i / (DL_G + 4)][i % (DL_G + 4)] = 0.f;
    Ay[i / (DL_G + 4)][i % (DL_G + 4)] = 0.f;
  }
  constexpr float IB0 = 1.f / DL_B0, IB1 = 1.f / DL_B1;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {  // block-uniform
  const int x0 = (t % tiles_x) * DL_T, y0 = ((t / tiles_x) % tiles_y) * DL_T;
  const int64_t img = (int64_t)(t / (tiles_x * tiles_y)) * plane;
  const float* dp = disp + img;
  for (int i = tid; i < DL_D * DL_D; i += 256) {
    const int r = i / DL_D, c = i % DL_D;
    D[r][c] = __ldg(dp + (int64_t)clampi(y0 - 4 + r, 0, H - 1) * W + clampi(x0 - 4 + c, 0, W - 1));
  }
  __syncthreads();
  for (int i = tid; i < DL_G * DL_G; i += 256) {
    const int r = i / DL_G, c = i % DL_G;       // position (y0 - 2 + r, x0 - 2 + c)
    const int y = y0 - 2 + r, x = x0 - 2 + c;
    float ax = 0.f, ay = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      float gx = 0.f, gy = 0.f;
#pragma unroll
      for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int b = 0; b < 5; ++b) {
          const float v = D[r + a][c + b];
          gx = fmaf(c_kx[a * 5 + b], v, gx);
          gy = fmaf(c_kx[b * 5 + a], v, gy);
        }
      const float m = sqrtf(gx * gx + gy * gy + 1e-8f);
      const bool core = r >= 2 && r < DL_T + 2 && c >= 2 && c < DL_T + 2;  // this CTA's own pixel
      float dm;  // dL/dm (times scale)
      if (EDGE) {
        const float e = __ldg(edge + img + (int64_t)y * W + x);
        const float w0 = IB0 * expf(-m * IB0), w1 = IB1 * expf(-m * IB1);  // the two Laplacian densities at m
        const float pdf = (1.f - e) * w0 + e * w1;
        const bool pass = pdf >= 1e-4f;  // clamp(min=1e-4) passes the gradient where the input is not below the bound
        const float dv = pass ? -scale / pdf : 0.f;  // d(-log pdf)/d pdf
        dm = -dv * ((1.f - e) * IB0 * w0 + e * IB1 * w1);
        if (core) {
          acc += (double)(-logf(fmaxf(pdf, 1e-4f)));
          cnt += 1.0;
          if (gedge != nullptr) gedge[img + (int64_t)y * W + x] = dv * (w1 - w0);
        }
      } else {
        dm = m <= 1.f ? scale : 0.f;  // clamp(m, 0, 1); m >= 1e-4 > 0
        if (core) {
          acc += (double)fminf(m, 1.f);
          cnt += 1.0;
        }
      }
      const float inv_m = 1.f / m;
      ax = dm * gx * inv_m;
      ay = dm * gy * inv_m;
    }
    Ax[r + 2][c + 2] = ax;
    Ay[r + 2][c + 2] = ay;
  }
  __syncthreads();
  if (gdisp != nullptr) {
    // grad w.r.t. the padded plane at cell (ry, rx): sum over positions p = r - t + 2, t = 0..4, of
    // Ax[p] kx[t] + Ay[p] ky[t]; cell -> tile-local position index: (ry - y0 + 2, rx - x0 + 2) - t + 2
    auto cell = [&](int ry, int rx) -> float {  // (ry, rx) within two cells of the tile: every index is inside the arrays
      const int pr0 = ry - y0 + 6, pc0 = rx - x0 + 6;
      float s = 0.f;
#pragma unroll
      for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int b = 0; b < 5; ++b) {
          s = fmaf(Ax[pr0 - a][pc0 - b], c_kx[a * 5 + b], s);
          s = fmaf(Ay[pr0 - a][pc0 - b], c_kx[b * 5 + a], s);
        }
      return s;
    };
    for (int i = tid; i < DL_T * DL_T; i += 256) {
      const int y = y0 + i / DL_T, x = x0 + i % DL_T;
      if (y >= H || x >= W) continue;
      // the pad cells that replicate this pixel: itself, and beyond the image border the two cells per side
      const int ry0 = y == 0 ? -2 : y, ry1 = y == H - 1 ? H + 1 : y;
      const int rx0 = x == 0 ? -2 : x, rx1 = x == W - 1 ? W + 1 : x;
      float g = 0.f;
      for (int ry = ry0; ry <= ry1; ++ry)
        for (int rx = rx0; rx <= rx1; ++rx) g += cell(ry, rx);
      gdisp[img + (int64_t)y * W + x] = g;
    }
  }
  __syncthreads();  // the tiles in shared memory are reused by the next step
  }
  finish_masked_sums(acc, cnt, partials, ticket, sums2);
}

}  // namespace ctd

using namespace ctd;

CTD_API int ctd_disparity_loss_f32(const float* disp, const float* edge, float* grad_disp, float* grad_edge, float* sums2,
                                      int64_t B, int64_t H, int64_t W, float scale, ctd_stream_t stream) {
  CTD_REQUIRE(B >= 0 && H >= 0 && W >= 0, "disparity_loss: negative size");
  CTD_REQUIRE(H <= INT32_MAX && W <= INT32_MAX, "disparity_loss: image too large");
  CTD_REQUIRE(sums2, "disparity_loss: null sums2");
  cudaStream_t st = as_stream(stream);
  if (B * H * W == 0) {
    CTD_CUDA(cudaMemsetAsync(sums2, 0, 2 * sizeof(float), st));
    return CTD_OK;
  }
  CTD_REQUIRE(disp, "disparity_loss: null disp");
  CTD_REQUIRE(edge || !grad_edge, "disparity_loss: grad_edge without edge");
  const int64_t tx = cdiv(W, DL_T), ty = cdiv(H, DL_T), ntiles = tx * ty * B;
  CTD_REQUIRE(ntiles <= INT32_MAX, "disparity_loss: too many tiles");
  // one CTA per tile while the deterministic sum's workspace holds them, else persistent over the tile list
  const unsigned grid = (unsigned)(ntiles <= MS_MAXBLK ? ntiles : 148 * 6);
  MsSlot ms;
  if (!ms_acquire(grid, st, &ms)) return fail(CTD_ERR_NOMEM, "disparity_loss: no reduction workspace");
  unsigned* ticket = ms.ticket;
  double* partials = ms.partials;
  if (edge)
    disparity_loss_kernel<true><<<grid, 256, 0, st>>>(disp, edge, grad_disp, grad_edge, (int)H, (int)W, (int)tx, (int)ty, (int)ntiles,
                                                     scale, partials, ticket, sums2);
  else
    disparity_loss_kernel<false><<<grid, 256, 0, st>>>(disp, nullptr, grad_disp, nullptr, (int)H, (int)W, (int)tx, (int)ty,
                                                      (int)ntiles, scale, partials, ticket, sums2);
  ms_release(&ms, st);
  count_launch();
  return check_launch("disparity_loss");
}
