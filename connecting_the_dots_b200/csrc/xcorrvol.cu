// XCorrVol (zero-mean normalised cross-correlation cost volume) for sm_100a.
//
// Reference semantics: torchext/ext/ext.h:120-191 (XCorrVolFunctor), bound at ext_cuda.cpp:73-86.
// out[d,h,w] = sum_c dot / (sqrt(sigma0 * sigma1) + 1e-8) over a bs x bs window of in0 at (h,w) and
// of in1 at (h, w-d); rows replicate-clamped, in0 columns clamped, in1 columns clamped AFTER the
// disparity shift.  The reference has no batch dimension; here B images are one launch.
#include <algorithm>

#include "ctd_common.cuh"

namespace ctd {

extern int g_force_generic;

// generic: any C, any block size, float or double; one thread per output, reference operation order
// (built with -fmad=false), so it reproduces the reference's two-pass centred statistics exactly.
template <typename T>
__global__ void __launch_bounds__(256)
xcorrvol_generic(const T* __restrict__ in0, const T* __restrict__ in1, T* __restrict__ out, int64_t B, int C,
                 int H, int W, int D, int bs) {
  const int64_t total = B * D * H * W;
  const int half = bs / 2;
  const T bs2 = (T)(bs * bs);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int w = idx % W, h = (idx / W) % H, d = (idx / ((int64_t)W * H)) % D;
    const int64_t b = idx / ((int64_t)W * H * D);
    T val = 0;
    for (int c = 0; c < C; ++c) {
      const T* p0 = in0 + (b * C + c) * H * W;
      const T* p1 = in1 + (b * C + c) * H * W;
      T mu0 = 0, mu1 = 0;
      for (int bh = 0; bh < bs; ++bh) {
        const int64_t row = (int64_t)clampi(h + bh - half, 0, H - 1) * W;
        for (int bw = 0; bw < bs; ++bw) {
          const int w0 = w + bw - half;
          mu0 += __ldg(p0 + row + clampi(w0, 0, W - 1)) / bs2;
          mu1 += __ldg(p1 + row + clampi(w0 - d, 0, W - 1)) / bs2;
        }
      }
      T s0 = 0, s1 = 0, dot = 0;
      for (int bh = 0; bh < bs; ++bh) {
        const int64_t row = (int64_t)clampi(h + bh - half, 0, H - 1) * W;
        for (int bw = 0; bw < bs; ++bw) {
          const int w0 = w + bw - half;
          const T v0 = __ldg(p0 + row + clampi(w0, 0, W - 1)) - mu0;
          const T v1 = __ldg(p1 + row + clampi(w0 - d, 0, W - 1)) - mu1;
          dot += v0 * v1;
          s0 += v0 * v0;
          s1 += v1 * v1;
        }
      }
      const T norm = (T)((double)sqrt(s0 * s1) + 1e-8);
      val += dot / norm;
    }
    out[idx] = val;
  }
}

// ------------------------------------------------------------------------------------------
// direct kernel: C = 1, odd block size BS <= 9, fp32.  Same centred two-pass arithmetic as the
// reference (so flat / low-texture windows, where normalised correlation is ill-conditioned, come out
// like ext_cpu's), but tiled: a CTA owns one image row, XD_TW pixels and XD_DC disparities.
//   * rows h-R..h+R of in0 (XD_TW + 2R columns) and of in1 (XD_TW + 2R + XD_DC - 1 columns, clamped
//     AFTER the disparity shift) are staged in shared memory once;
//   * mean and sigma of every in1 window the CTA touches depend only on x' = w - d: computed once per
//     x' in the prologue instead of once per (w, d);
//   * a thread owns one pixel, keeps its centred in0 window (BS*BS values) in registers, and sweeps
//     the disparities four at a time so each shared-memory row segment feeds 4 * BS taps.
// ------------------------------------------------------------------------------------------
constexpr int XD_TW = 128;  // pixels (threads) per CTA
constexpr int XD_DC = 32;   // disparities per CTA

template <int BS>
__global__ void __launch_bounds__(XD_TW)
xcorrvol_direct(const float* __restrict__ in0, const float* __restrict__ in1, float* __restrict__ out, int H, int W,
                int D, int nchunks) {
  constexpr int R = BS / 2, N = BS * BS;
  constexpr int AW = XD_TW + 2 * R;               // in0 tile width
  constexpr int BW = XD_TW + 2 * R + XD_DC - 1;   // in1 tile width
  constexpr int NP = XD_TW + XD_DC - 1;           // in1 window positions x' touched by this CTA
  __shared__ float As[BS][AW];
  __shared__ float Bs[BS][BW];
  __shared__ float Mu1[NP], S1[NP];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * XD_TW, h = blockIdx.y;
  const int b = blockIdx.z / nchunks, d0 = (blockIdx.z % nchunks) * XD_DC;
  const int64_t plane = (int64_t)H * W;
  const float* a = in0 + b * plane;
  const float* bp = in1 + b * plane;
  // in1 tile column j <-> unclamped image column x0 - R - (d0 + XD_DC - 1) + j
  const int bcol0 = x0 - R - (d0 + XD_DC - 1);
  for (int i = tid; i < BS * AW; i += XD_TW) {
    const int r = i / AW, j = i % AW;
    As[r][j] = __ldg(a + (int64_t)clampi(h - R + r, 0, H - 1) * W + clampi(x0 - R + j, 0, W - 1));
  }
  for (int i = tid; i < BS * BW; i += XD_TW) {
    const int r = i / BW, j = i % BW;
    Bs[r][j] = __ldg(bp + (int64_t)clampi(h - R + r, 0, H - 1) * W + clampi(bcol0 + j, 0, W - 1));
  }
  __syncthreads();
  // statistics of the in1 windows: position p <-> window whose left tap is tile column p
  for (int p = tid; p < NP; p += XD_TW) {
    float mu = 0.f;
#pragma unroll
    for (int r = 0; r < BS; ++r)
#pragma unroll
      for (int c = 0; c < BS; ++c) mu += Bs[r][p + c] * (1.0f / float(N));
    float sg = 0.f;
#pragma unroll
    for (int r = 0; r < BS; ++r)
#pragma unroll
      for (int c = 0; c < BS; ++c) {
        const float v = Bs[r][p + c] - mu;
        sg += v * v;
      }
    Mu1[p] = mu;
    S1[p] = sg;
  }
  // this thread's centred in0 window
  float ac[N];
  float mu0 = 0.f;
#pragma unroll
  for (int r = 0; r < BS; ++r)
#pragma unroll
    for (int c = 0; c < BS; ++c) {
      ac[r * BS + c] = As[r][tid + c];
      mu0 += ac[r * BS + c] * (1.0f / float(N));
    }
  float s0 = 0.f;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    ac[k] -= mu0;
    s0 += ac[k] * ac[k];
  }
  __syncthreads();
  const int w = x0 + tid;
  // disparity d = d0 + q: window left tap at tile column tid + (XD_DC - 1) - q
  for (int q = 0; q < XD_DC; q += 4) {
    const int base = tid + XD_DC - 4 - q;  // left tap of the window of disparity d0 + q + 3
    float dot[4] = {0.f, 0.f, 0.f, 0.f};
    float mu1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) mu1[j] = Mu1[base + 3 - j];
#pragma unroll
    for (int r = 0; r < BS; ++r) {
      float v[BS + 3];
#pragma unroll
      for (int c = 0; c < BS + 3; ++c) v[c] = Bs[r][base + c];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < BS; ++c) dot[j] += ac[r * BS + c] * (v[3 - j + c] - mu1[j]);
    }
    if (w < W) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = d0 + q + j;
        if (d < D) {
          const float norm = (float)((double)sqrtf(s0 * S1[base + 3 - j]) + 1e-8);
          out[((int64_t)(b * D + d) * H + h) * W + w] = dot[j] / norm;
        }
      }
    }
  }
}

template <typename T>
static int xcorrvol_impl(const T* in0, const T* in1, T* out, int64_t B, int64_t C, int64_t H, int64_t W, int64_t D,
                         int bs, cudaStream_t st) {
  CTD_REQUIRE(B >= 0 && C >= 0 && H >= 0 && W >= 0 && D >= 0, "xcorrvol: negative size");
  CTD_REQUIRE(bs >= 1 && bs <= 255, "xcorrvol: block_size %d out of range [1,255]", bs);
  CTD_REQUIRE(H <= INT32_MAX && W <= INT32_MAX && D <= INT32_MAX && C <= INT32_MAX, "xcorrvol: dimension too large");
  const int64_t total = B * D * H * W;
  if (total == 0) return CTD_OK;
  CTD_REQUIRE(out && (C == 0 || (in0 && in1)), "xcorrvol: null pointer");
  if (sizeof(T) == 4 && C == 1 && !g_force_generic && (bs == 3 || bs == 5 || bs == 7 || bs == 9) && H <= 65535) {
    const int64_t nchunks = cdiv(D, XD_DC);
    for (int64_t b0 = 0; b0 < B; b0 += 1024) {  // keep gridDim.z under 65535
      const int64_t nb = std::min<int64_t>(1024, B - b0);
      if (nb * nchunks > 65535) break;
      const dim3 grid((unsigned)cdiv(W, XD_TW), (unsigned)H, (unsigned)(nb * nchunks));
      const float* p0 = reinterpret_cast<const float*>(in0) + b0 * H * W;
      const float* p1 = reinterpret_cast<const float*>(in1) + b0 * H * W;
      float* po = reinterpret_cast<float*>(out) + b0 * D * H * W;
      if (bs == 9) xcorrvol_direct<9><<<grid, XD_TW, 0, st>>>(p0, p1, po, (int)H, (int)W, (int)D, (int)nchunks);
      else if (bs == 7) xcorrvol_direct<7><<<grid, XD_TW, 0, st>>>(p0, p1, po, (int)H, (int)W, (int)D, (int)nchunks);
      else if (bs == 5) xcorrvol_direct<5><<<grid, XD_TW, 0, st>>>(p0, p1, po, (int)H, (int)W, (int)D, (int)nchunks);
      else xcorrvol_direct<3><<<grid, XD_TW, 0, st>>>(p0, p1, po, (int)H, (int)W, (int)D, (int)nchunks);
      count_launch();
      if (b0 + nb >= B) return check_launch("xcorrvol(direct)");
    }
  }
  const int grid = (int)std::min<int64_t>(cdiv(total, 256), 148 * 256);
  xcorrvol_generic<T><<<grid, 256, 0, st>>>(in0, in1, out, B, (int)C, (int)H, (int)W, (int)D, bs);
  count_launch();
  return check_launch("xcorrvol(generic)");
}

}  // namespace ctd

CTD_API int ctd_xcorrvol_f32(const float* in0, const float* in1, float* out, int64_t B, int64_t C, int64_t H,
                                int64_t W, int64_t D, int bs, ctd_stream_t s) {
  return ctd::xcorrvol_impl<float>(in0, in1, out, B, C, H, W, D, bs, ctd::as_stream(s));
}
CTD_API int ctd_xcorrvol_f64(const double* in0, const double* in1, double* out, int64_t B, int64_t C, int64_t H,
                                int64_t W, int64_t D, int bs, ctd_stream_t s) {
  return ctd::xcorrvol_impl<double>(in0, in1, out, B, C, H, W, D, bs, ctd::as_stream(s));
}
