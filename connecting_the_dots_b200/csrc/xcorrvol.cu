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

template <typename T>
static int xcorrvol_impl(const T* in0, const T* in1, T* out, int64_t B, int64_t C, int64_t H, int64_t W, int64_t D,
                         int bs, cudaStream_t st) {
  CTD_REQUIRE(B >= 0 && C >= 0 && H >= 0 && W >= 0 && D >= 0, "xcorrvol: negative size");
  CTD_REQUIRE(bs >= 1 && bs <= 255, "xcorrvol: block_size %d out of range [1,255]", bs);
  CTD_REQUIRE(H <= INT32_MAX && W <= INT32_MAX && D <= INT32_MAX && C <= INT32_MAX, "xcorrvol: dimension too large");
  const int64_t total = B * D * H * W;
  if (total == 0) return CTD_OK;
  CTD_REQUIRE(out && (C == 0 || (in0 && in1)), "xcorrvol: null pointer");
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
