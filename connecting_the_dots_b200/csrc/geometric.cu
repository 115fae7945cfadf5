// Geometric (projection depth similarity) loss of the stage-2 trainer, one direction per launch:
// model/networks.py:474-503 (ProjectionDepthSimilarityLoss.fwd) on top of ProjectionBaseLoss.unproject /
// transform / project (networks.py:436-472).  SURVEY section 8(f) rank 3.
//
// Reference, per pixel p of frame A (row vectors):
//   xyz = depthA[p] * ray[p]                       unproject   (networks.py:448)
//   xyz = (xyz - tA) @ RA                           transform   (networks.py:436-442)
//   xyz = xyz @ RB^T + tB                           project     (networks.py:456-457)
//   uvw = xyz @ K^T;  d = uvw[2];  uv = uvw[:2] / (relu(d) + 1e-12)          (networks.py:459-466)
//   g   = 2 * (uv / (size - 1) - 0.5)               (networks.py:485-486)
//   s   = grid_sample(depthB, g, bilinear, padding_mode='border', align_corners=False)   (networks.py:489)
//   diff = |d - s|, clamped to [0, clamp] when clamp > 0; loss = mean(diff)   (networks.py:491-497)
// which in torch is two bmm, a relu, a division, grid_sample and ~10 elementwise kernels with their autograd twins.
// Here: ONE kernel per direction computes the loss sum, the gradient w.r.t. depthA (each pixel's own, a plain
// store / read-modify-write) and the gradient w.r.t. depthB (bilinear scatter, fp32 atomicAdd into at most four
// neighbours -- the only floating-point atomics of the library, so grad_depthB is reproducible to rounding, not
// bit for bit).  Both gradients are those of sum(diff) * scale; the caller passes scale = 1 / (B*H*W).
//
// Algorithmic bytes: 4 (depthA) + 12 (ray, L2-resident: one [H*W,3] table for all images) + 4 (grad depthA)
// + gathers / atomics on depthB and grad depthB (~8 B/px when the motion is smooth) = ~20 B/px per direction.
#include <algorithm>
#include <cmath>

#include "ctd_common.cuh"

namespace ctd {

__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
  return fmaf(a2, b2, fmaf(a1, b1, a0 * b0));
}

// grid value -> source index, align_corners = false (ATen grid_sampler_unnormalize), then the border clip
__device__ __forceinline__ float geo_source_index(float g, int size, float* mult) {
  float i = fmaf(g + 1.f, (float)size, -1.f) / 2.f;
  *mult = (float)size / 2.f;
  if (!(i > 0.f)) {  // also catches NaN like ATen's clip (max(i, 0) = 0 for NaN)
    i = 0.f;
    *mult = 0.f;
  } else if (i >= (float)(size - 1)) {
    i = (float)(size - 1);
    *mult = 0.f;
  }
  return i;
}

// measured (one direction, batch 8): 2 pixels x 4 blocks/SM 36.9 us, 4 x 2 46.1 us, 2 x 3 51.8 us, 4 x 3 68.9 us (spills)
#ifndef CTD_GEO_PX
#define CTD_GEO_PX 2
#endif
#ifndef CTD_GEO_CTAS
#define CTD_GEO_CTAS 4
#endif
constexpr int GEO_T = 256, GEO_PX = CTD_GEO_PX, GEO_CTAS = CTD_GEO_CTAS;  // threads per block, pixels per thread, resident blocks per SM

// direct_accumulate: 0 = gA[p] is overwritten with this pixel's gradient, 1 = it is added to what is there (the
// second direction of tforward, networks.py:500-503: gA then already holds the scatter of the first direction).
// grid = (blocks per image, B): a block stays inside one image (the poses sit in shared memory) and walks it in
// steps of gridDim.x * 256 * GEO_PX pixels (the grid is GEO_CTAS blocks per SM: a block's set-up and its part in the final sum are
// paid once per several steps).  A thread owns GEO_PX pixels 256 apart, so the lanes of a warp touch neighbouring pixels in
// every load, store and atomic (a warp's scatter lands in one or two cache lines per instruction; four consecutive
// pixels per thread were measured 2x slower in the atomics).  Depth and rays are requested one step ahead; the
// projections are computed, then all the step's bilinear taps are fetched together, so a step waits for one round trip
// (the taps) however many pixels are in flight.
__global__ void __launch_bounds__(GEO_T, GEO_CTAS)
depth_similarity_kernel(const float* __restrict__ depthA, const float* __restrict__ depthB, const float* __restrict__ ray,
                        const float* __restrict__ K, const float* __restrict__ RA, const float* __restrict__ tA,
                        const float* __restrict__ RB, const float* __restrict__ tB, float* __restrict__ gA,
                        float* __restrict__ gB, int H, int W, float inv_w, float inv_h, float clampv,
                        float scale, int direct_accumulate, double* __restrict__ partials, unsigned* __restrict__ ticket,
                        float* __restrict__ sums2) {
  __shared__ float cst[48];  // K, RA, RB, tA, tB of this image; [36..44]: M = RA RB^T K^T (d uvw / d depth = ray M)
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.y;
  if (tid < 9) cst[tid] = __ldg(K + tid);
  else if (tid < 18) cst[tid] = __ldg(RA + b * 9 + (tid - 9));
  else if (tid < 27) cst[tid] = __ldg(RB + b * 9 + (tid - 18));
  else if (tid < 30) cst[tid] = __ldg(tA + b * 3 + (tid - 27));
  else if (tid < 33) cst[tid] = __ldg(tB + b * 3 + (tid - 30));
  __syncthreads();
  if (tid < 9) {  // M[i][j] = sum_a sum_c RA[i][a] RB[c][a] K[j][c]: the derivative chain of a pixel is three FMAs per component
    const int i = tid / 3, j = tid % 3;
    float m = 0.f;
    for (int c = 0; c < 3; ++c) {
      float rr = 0.f;  // (RA RB^T)[i][c]
      for (int a = 0; a < 3; ++a) rr = fmaf(cst[9 + 3 * i + a], cst[18 + 3 * c + a], rr);
      m = fmaf(rr, cst[3 * j + c], m);
    }
    cst[36 + tid] = m;
  }
  __syncthreads();
  const float* k = cst;
  const float* ra = cst + 9;
  const float* rb = cst + 18;
  const float* ta = cst + 27;
  const float* tb = cst + 30;
  const float* M = cst + 36;
  const unsigned hw = (unsigned)H * (unsigned)W;
  const float* dA_img = depthA + b * (int64_t)hw;
  const float* dB_img = depthB + b * (int64_t)hw;
  float* gA_img = gA ? gA + b * (int64_t)hw : nullptr;
  float* gB_img = gB ? gB + b * (int64_t)hw : nullptr;
  double acc = 0.0, cnt = 0.0;
  // inputs of a step: depth and rays of the thread's pixels; the next step's are requested before this step's
  // arithmetic so their round trip overlaps it
  auto fetch = [&](unsigned p0, float (&d)[GEO_PX], float (&q)[GEO_PX][3]) {
#pragma unroll
    for (int m = 0; m < GEO_PX; ++m) {
      const unsigned pix = p0 + m * GEO_T;
      if (pix < hw) {
        d[m] = __ldg(dA_img + pix);
#pragma unroll
        for (int j = 0; j < 3; ++j) q[m][j] = __ldg(ray + (size_t)pix * 3 + j);
      }
    }
  };
  const unsigned pstep = gridDim.x * GEO_T * GEO_PX;
  float nd[GEO_PX] = {}, nq[GEO_PX][3] = {};
  fetch(blockIdx.x * GEO_T * GEO_PX + tid, nd, nq);
  for (unsigned p0 = blockIdx.x * GEO_T * GEO_PX + tid; p0 < hw; p0 += pstep) {
    float dA[GEO_PX], r[GEO_PX][3];
#pragma unroll
    for (int m = 0; m < GEO_PX; ++m) {
      dA[m] = nd[m];
#pragma unroll
      for (int j = 0; j < 3; ++j) r[m][j] = nq[m][j];
    }
    fetch(p0 + pstep, nd, nq);
    // phase 1: projections, sample positions, tap loads
    float w[GEO_PX], u[GEO_PX], v[GEO_PX], rden[GEO_PX], A[GEO_PX][3], ix[GEO_PX], iy[GEO_PX], mx[GEO_PX], my[GEO_PX];
    float tap[GEO_PX][4];
    int xa[GEO_PX], ya[GEO_PX];
#pragma unroll
    for (int m = 0; m < GEO_PX; ++m) {
      // forward chain (value); its derivative w.r.t. depthA is the same chain on the ray alone (translations drop out),
    // folded into one 3x3 product per image
      const float x0 = dA[m] * r[m][0] - ta[0], x1 = dA[m] * r[m][1] - ta[1], x2 = dA[m] * r[m][2] - ta[2];
      float y[3], z[3], U[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) y[j] = dot3(x0, x1, x2, ra[j], ra[3 + j], ra[6 + j]);  // xyz @ RA: column j of RA
#pragma unroll
      for (int j = 0; j < 3; ++j)  // xyz @ RB^T + tB: row j of RB
        z[j] = dot3(y[0], y[1], y[2], rb[3 * j], rb[3 * j + 1], rb[3 * j + 2]) + tb[j];
#pragma unroll
      for (int j = 0; j < 3; ++j) {  // xyz @ K^T: row j of K; and d uvw / d depth = ray @ M
        U[j] = dot3(z[0], z[1], z[2], k[3 * j], k[3 * j + 1], k[3 * j + 2]);
        A[m][j] = dot3(r[m][0], r[m][1], r[m][2], M[j], M[3 + j], M[6 + j]);
      }
      w[m] = U[2];
      const float den = fmaxf(w[m], 0.f) + 1e-12f;
      rden[m] = rcp_approx(den);  // gradient path only (the sample position uses the IEEE quotients below)
      u[m] = U[0] / den;
      v[m] = U[1] / den;
      const float gx = 2.f * (u[m] * inv_w - 0.5f), gy = 2.f * (v[m] * inv_h - 0.5f);
      ix[m] = geo_source_index(gx, W, &mx[m]);
      iy[m] = geo_source_index(gy, H, &my[m]);
      xa[m] = (int)floorf(ix[m]);
      ya[m] = (int)floorf(iy[m]);
      const bool xin = xa[m] + 1 < W, yin = ya[m] + 1 < H;  // xa, ya are inside after the clip
      const float* pB = dB_img + (unsigned)ya[m] * (unsigned)W + (unsigned)xa[m];
      tap[m][0] = __ldg(pB);
      tap[m][1] = xin ? __ldg(pB + 1) : 0.f;
      tap[m][2] = yin ? __ldg(pB + W) : 0.f;
      tap[m][3] = (xin && yin) ? __ldg(pB + W + 1) : 0.f;
    }
    // phase 2: loss, scatter to grad depthB, this pixel's grad depthA
    float gd[GEO_PX];
#pragma unroll
    for (int m = 0; m < GEO_PX; ++m) {
      gd[m] = 0.f;
      if (p0 + m * GEO_T >= hw) continue;
      const int xb = xa[m] + 1, yb = ya[m] + 1;
      const bool xin = xb < W, yin = yb < H;
      const float wnw = (xb - ix[m]) * (yb - iy[m]), wne = (ix[m] - xa[m]) * (yb - iy[m]);
      const float wsw = (xb - ix[m]) * (iy[m] - ya[m]), wse = (ix[m] - xa[m]) * (iy[m] - ya[m]);
      const float vnw = tap[m][0], vne = tap[m][1], vsw = tap[m][2], vse = tap[m][3];
      float s = vnw * wnw;
      s = fmaf(vne, wne, s);
      s = fmaf(vsw, wsw, s);
      s = fmaf(vse, wse, s);
      const float e = w[m] - s;
      float diff = fabsf(e);
      bool pass = true;  // torch.clamp passes the gradient where min <= x <= max
      if (clampv > 0.f) {
        pass = diff <= clampv;
        diff = fminf(diff, clampv);
      }
      acc += (double)diff;
      cnt += 1.0;
      // backward of sum(diff) * scale
      const float g = pass ? (e > 0.f ? scale : (e < 0.f ? -scale : 0.f)) : 0.f;
      if (gB_img != nullptr && g != 0.f) {  // d/d depthB: -g times the bilinear weights
        float* qB = gB_img + (unsigned)ya[m] * (unsigned)W + (unsigned)xa[m];
        atomicAdd(qB, -g * wnw);
        if (xin) atomicAdd(qB + 1, -g * wne);
        if (yin) atomicAdd(qB + W, -g * wsw);
        if (xin && yin) atomicAdd(qB + W + 1, -g * wse);
      }
      // d s / d ix, d s / d iy (ATen grid_sampler_2d_backward), through the clip and the grid normalisation
      const float ds_dix = (vne - vnw) * (yb - iy[m]) + (vse - vsw) * (iy[m] - ya[m]);
      const float ds_diy = (vsw - vnw) * (xb - ix[m]) + (vse - vne) * (ix[m] - xa[m]);
      const float gu = -g * ds_dix * mx[m] * (2.f * inv_w);  // dL/du
      const float gv = -g * ds_diy * my[m] * (2.f * inv_h);  // dL/dv
      const float gU0 = gu * rden[m], gU1 = gv * rden[m];
      float gw = g;  // d appears directly in diff ...
      if (w[m] > 0.f) gw -= (gU0 * u[m] + gU1 * v[m]);  // ... and in the perspective division (relu: only where d > 0)
      gd[m] = fmaf(gw, A[m][2], fmaf(gU1, A[m][1], gU0 * A[m][0]));
    }
    if (gA_img != nullptr) {
#pragma unroll
      for (int m = 0; m < GEO_PX; ++m) {
        const unsigned pix = p0 + m * GEO_T;
        if (pix < hw) gA_img[pix] = direct_accumulate ? gA_img[pix] + gd[m] : gd[m];
      }
    }
  }
  finish_masked_sums(acc, cnt, partials, ticket, sums2);
}

}  // namespace ctd

using namespace ctd;

CTD_API int ctd_depth_similarity_f32(const float* depthA, const float* depthB, const float* ray, const float* K,
                                        const float* RA, const float* tA, const float* RB, const float* tB,
                                        float* grad_depthA, float* grad_depthB, float* sums2, int64_t B, int64_t H,
                                        int64_t W, float clamp, float scale, int direct_accumulate,
                                        ctd_stream_t stream) {
  CTD_REQUIRE(B >= 0 && H >= 0 && W >= 0, "depth_similarity: negative size");
  CTD_REQUIRE(H * W < ((int64_t)1 << 31) / 3, "depth_similarity: image too large");
  CTD_REQUIRE(sums2, "depth_similarity: null sums2");
  const int64_t total = B * H * W;
  cudaStream_t st = as_stream(stream);
  if (total == 0) {
    CTD_CUDA(cudaMemsetAsync(sums2, 0, 2 * sizeof(float), st));
    return CTD_OK;
  }
  CTD_REQUIRE(depthA && depthB && ray && K && RA && tA && RB && tB, "depth_similarity: null pointer");
  CTD_REQUIRE(B <= 65535, "depth_similarity: batch too large");
  // blocks per image: enough to fill the GPU, at most MS_MAXBLK blocks in all (the deterministic sum's workspace)
  // two resident blocks per SM in all (a block's set-up and its share of the final reduction are amortised over
  // many steps), at least one block per image
  const int64_t per_img = std::max<int64_t>(1, std::min<int64_t>(cdiv(H * W, GEO_T * GEO_PX), cdiv(148 * GEO_CTAS, std::max<int64_t>(B, 1))));
  CTD_REQUIRE(per_img * B <= MS_MAXBLK, "depth_similarity: batch too large for the reduction workspace");
  const dim3 grid((unsigned)per_img, (unsigned)B);
  MsSlot ms;
  if (!ms_acquire((size_t)(per_img * B), st, &ms)) return fail(CTD_ERR_NOMEM, "depth_similarity: no reduction workspace");
  unsigned* ticket = ms.ticket;
  double* partials = ms.partials;
  const float inv_w = W > 1 ? 1.f / (float)(W - 1) : INFINITY, inv_h = H > 1 ? 1.f / (float)(H - 1) : INFINITY;
  depth_similarity_kernel<<<grid, GEO_T, 0, st>>>(depthA, depthB, ray, K, RA, tA, RB, tB, grad_depthA, grad_depthB, (int)H, (int)W,
                                                  inv_w, inv_h, clamp, scale, direct_accumulate, partials, ticket, sums2);
  ms_release(&ms, st);
  count_launch();
  return check_launch("depth_similarity");
}
