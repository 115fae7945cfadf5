// Local contrast normalisation for sm_100a, one fused kernel.
//
// Reference semantics: model/networks.py:507-533 (LCN.tforward): reflection padding, (2r+1)^2 box
// sums of x and x^2 (two library convolutions there), avg = box/n, std = sqrt(box2/n - avg^2 + 1e-6)
// + eps, lcn = (x - avg)/std.  var = E[x^2] - E[x]^2 cancels catastrophically on flat image regions,
// so the box sums are accumulated in fp64 (the "exact box sum" reading of the formula, same as
// oracle/ctd_oracle_impl.h) and everything after them follows the reference's fp32 operation order
// with IEEE division and square root and no FMA contraction (this file is built with -fmad=false).
//
// Layout: a CTA owns a strip of 128 output columns and a run of rows.  Vertical pass: one thread per
// (halo'd) column keeps running fp64 sums of the last 2r+1 rows straight from global memory
// (coalesced row reads).  Horizontal pass: the column sums of 8 rows go through shared memory and
// each thread produces 4 adjacent pixels with a sliding fp64 window, then the per-pixel epilogue and
// 128-bit stores of lcn and std.
#include <algorithm>
#include <type_traits>

#include "ctd_common.cuh"
#include "ctd_tma.cuh"

namespace ctd {

constexpr int LT_W = 128;      // output columns per CTA
constexpr int L_RMAX = 16;     // largest radius the strip kernel takes (columns needed: LT_W + 2r)
constexpr int L_COLS = LT_W + 2 * L_RMAX;
// Column c of a row of fp64 column sums lives at (c % 4) * L_SUB + c / 4: the horizontal pass reads
// columns 4q + k for the 32 lanes q of a warp, which this layout makes consecutive (no bank
// conflicts), and 2 * L_SUB % 32 == 24 keeps the vertical pass's stores conflict-free too.
constexpr int L_SUB = 44;
static_assert(4 * L_SUB >= L_COLS, "column store too small");
__device__ __forceinline__ int lcol(int c) { return (c & 3) * L_SUB + (c >> 2); }
constexpr int L_RG = 8;        // rows per shared-memory round
extern int g_force_generic;
constexpr int L_THREADS = 256;

// a / n correctly rounded for a constant n whose reciprocal inv_n = RN(1/n) is given (Markstein: one
// residual step on the rounded product is exact when 1/n is correctly rounded and n's significand is not all
// ones; checked exhaustively for n = 121 by tools/experiments/div121_exhaustive.cu: bit-identical to IEEE division
// for every finite float except -0, which comes out as +0 -- harmless, avg only enters as avg * avg and x - avg).
// Three FMA-pipe instructions instead of the IEEE division's reciprocal, refinement and range checks.
__device__ __forceinline__ float div_const_rn(float a, float n, float inv_n) {
  const float q = a * inv_n;
  const float r = fmaf(-q, n, a);
  return fmaf(r, inv_n, q);
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ int reflect(int i, int n) {  // torch ReflectionPad2d: no edge repeat
  if (i < 0) i = -i;
  if (i > n - 1) i = 2 * (n - 1) - i;
  return i;
}

template <typename T>
__global__ void __launch_bounds__(L_THREADS)
lcn_strip_kernel(const T* __restrict__ x, T* __restrict__ lcn, T* __restrict__ sd_out, int H, int W, int r,
                 T eps, int rows_per_cta, int vec) {
  __shared__ double V1[L_RG][4 * L_SUB];
  __shared__ double V2[L_RG][4 * L_SUB];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * LT_W;
  const int y0 = blockIdx.y * rows_per_cta;
  const int y1 = min(H, y0 + rows_per_cta);
  const int64_t plane = (int64_t)H * W;
  const T* xp = x + blockIdx.z * plane;
  T* lp = lcn + blockIdx.z * plane;
  T* sp = sd_out + blockIdx.z * plane;
  const int ncols = LT_W + 2 * r;
  const T n = (T)((2 * r + 1) * (2 * r + 1));

  // vertical state of this thread's column
  const bool vthread = tid < ncols;
  const int gx = reflect(min(x0 - r + tid, W - 1 + r), W);
  double s1 = 0.0, s2 = 0.0;
  if (vthread) {
    for (int yy = y0 - r; yy < y0 + r; ++yy) {
      const T v = __ldg(xp + (int64_t)reflect(yy, H) * W + gx);
      s1 += (double)v;
      s2 += (double)(v * v);
    }
  }
  for (int yb = y0; yb < y1; yb += L_RG) {
    if (vthread) {
#pragma unroll
      for (int j = 0; j < L_RG; ++j) {
        const int yy = yb + j;
        if (yy < y1) {
          const T vn = __ldg(xp + (int64_t)reflect(yy + r, H) * W + gx);
          const T vo = __ldg(xp + (int64_t)reflect(yy - r, H) * W + gx);
          s1 += (double)vn;
          s2 += (double)(vn * vn);
          V1[j][lcol(tid)] = s1;
          V2[j][lcol(tid)] = s2;
          s1 -= (double)vo;
          s2 -= (double)(vo * vo);
        }
      }
    }
    __syncthreads();
    {
      const int j = tid / 32, q = tid % 32;  // 8 rows x 32 quads
      const int yy = yb + j, xq = x0 + 4 * q;
      if (yy < y1 && xq < W) {
        const double* a1 = &V1[j][q];  // column 4q + k is a1[lcol(k)]
        const double* a2 = &V2[j][q];
        double h1 = 0.0, h2 = 0.0;
        for (int k = 0; k <= 2 * r; ++k) {
          h1 += a1[lcol(k)];
          h2 += a2[lcol(k)];
        }
        T xv[4];
        const int64_t off = (int64_t)yy * W + xq;
        if (vec) {
          if (sizeof(T) == 4) {
            const float4 t = ldg4(reinterpret_cast<const float*>(xp + off));
            xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
          } else {
#pragma unroll
            for (int m = 0; m < 4; ++m) xv[m] = __ldg(xp + off + m);
          }
        } else {
#pragma unroll
          for (int m = 0; m < 4; ++m) xv[m] = xq + m < W ? __ldg(xp + off + m) : T(0);
        }
        T ol[4], os[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const T box = (T)h1, box2 = (T)h2;
          const T avg = box / n;
          const T var = box2 / n - avg * avg + (T)1e-6;
          const T sd = sqrt(var) + eps;
          ol[m] = (xv[m] - avg) / sd;
          os[m] = sd;
          if (m < 3) {
            h1 += a1[lcol(m + 2 * r + 1)] - a1[lcol(m)];
            h2 += a2[lcol(m + 2 * r + 1)] - a2[lcol(m)];
          }
        }
        if (vec && sizeof(T) == 4) {
          *reinterpret_cast<float4*>(lp + off) = make_float4(ol[0], ol[1], ol[2], ol[3]);
          *reinterpret_cast<float4*>(sp + off) = make_float4(os[0], os[1], os[2], os[3]);
        } else {
#pragma unroll
          for (int m = 0; m < 4; ++m)
            if (xq + m < W) {
              lp[off + m] = ol[m];
              sp[off + m] = os[m];
            }
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// fp32, radius <= 5: persistent TMA-fed kernel.  A CTA walks a strided list of 128x16 output tiles; one
// thread fetches each tile's 144x26 halo box with cp.async.bulk.tensor into a two-stage ring (the next
// tile loads while this one is normalised).  Reflection padding is an index remap on the staged tile
// (the mirrored pixels always lie inside the box).  Vertical pass: one thread per column builds fp64
// prefix sums of x and x^2 down the box and emits 11-row window sums as differences of prefixes held
// in a register ring; horizontal pass and epilogue as in lcn_strip_kernel.
// ------------------------------------------------------------------------------------------
constexpr int LTM_W = 128, LTM_H = 16, LTM_R = 5;
constexpr int LTM_BW = 144;                       // box: columns x0-8 .. x0+135 (16-byte aligned origin)
constexpr int LTM_BH = LTM_H + 2 * LTM_R;         // 26 rows y0-5 .. y0+20
constexpr int LTM_XOFF = 8;                       // tile column of image column x0
constexpr int LTM_BOX_BYTES = LTM_BW * LTM_BH * 4;  // 14976 = 117 * 128
constexpr int LTM_SUB = 36;                       // interleaved fp64 column layout, see lcol()
static_assert(LTM_BOX_BYTES % 128 == 0, "stage alignment");
__device__ __forceinline__ int lcol36(int c) { return (c & 3) * LTM_SUB + (c >> 2); }

struct alignas(128) LcnSmem {
  float x[2][LTM_BH][LTM_BW];
  double v1[LTM_H][4 * LTM_SUB];
  double v2[LTM_H][4 * LTM_SUB];
  uint64_t full[2];
};

template <int R>
__global__ void __launch_bounds__(256, 3)
lcn_tma_kernel(const __grid_constant__ CUtensorMap map_x, float* __restrict__ lcn, float* __restrict__ sd_out, int H,
               int W, float eps, int tiles_x, int tiles_y, int ntiles) {
  extern __shared__ unsigned char smem_raw[];
  LcnSmem& S = *reinterpret_cast<LcnSmem*>(align128_shared(smem_raw));
  constexpr int K = 2 * R + 1;
  const float n = float(K * K), inv_n = 1.0f / float(K * K);
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
    mbar_expect_tx(&S.full[0], LTM_BOX_BYTES);
    tma_load_3d(&S.x[0][0][0], &map_x, &S.full[0], walk.tx * LTM_W - LTM_XOFF, walk.ty * LTM_H - LTM_R, walk.n);
  }
  for (int it = 0; t < ntiles; ++it, t += gridDim.x, walk = walk.next()) {
    const int s = it & 1;
    if (tid == 0 && t + (int)gridDim.x < ntiles) {
      const TileWalk nx = walk.next();
      fence_proxy_async();
      mbar_expect_tx(&S.full[s ^ 1], LTM_BOX_BYTES);
      tma_load_3d(&S.x[s ^ 1][0][0], &map_x, &S.full[s ^ 1], nx.tx * LTM_W - LTM_XOFF, nx.ty * LTM_H - LTM_R, nx.n);
    }
    const int x0 = walk.tx * LTM_W, y0 = walk.ty * LTM_H, img = walk.n;
    mbar_wait(&S.full[s], (it >> 1) & 1);
    // vertical pass: thread c owns needed column c (image column x0 - R + c), 0 <= c < 128 + 2R.  Tiles whose box
    // rows all lie inside the image (all but the first and last tile row) skip the per-row reflection remap.
    if (tid < LTM_W + 2 * R) {
      // columns right of the last needed one (partial last tile) would reflect to a negative index: pin them to it
      const int cc = reflect(min(x0 - R + tid, W - 1 + R), W) - (x0 - LTM_XOFF);
      auto vertical = [&](auto interior_tag) {
        constexpr bool INTERIOR = decltype(interior_tag)::value;
        double p1[K + 1], p2[K + 1];  // ring of prefix sums: slot j % (K+1) holds prefix through box row j-1
        p1[0] = 0.0;
        p2[0] = 0.0;
#pragma unroll
        for (int j = 0; j < LTM_H + 2 * R; ++j) {  // box rows y0 - R + j
          const int rr = INTERIOR ? j + (LTM_R - R) : reflect(y0 - R + j, H) - (y0 - LTM_R);
          const float v = S.x[s][rr][cc];
          const int cur = (j + 1) % (K + 1), prev = j % (K + 1);
          p1[cur] = p1[prev] + (double)v;
          p2[cur] = p2[prev] + (double)(v * v);
          if (j >= 2 * R) {  // rows j-2R .. j form the window of output row j - 2R
            const int old = (j + 1 + 1) % (K + 1);  // slot holding the prefix through row j - 2R - 1
            S.v1[j - 2 * R][lcol36(tid)] = p1[cur] - p1[old];
            S.v2[j - 2 * R][lcol36(tid)] = p2[cur] - p2[old];
          }
        }
      };
      if (y0 >= R && y0 + LTM_H + R <= H) vertical(std::true_type{});
      else vertical(std::false_type{});
    }
    __syncthreads();
    // horizontal pass + epilogue: 16 rows x 32 quads, two items per thread
#pragma unroll 1
    for (int item = tid; item < LTM_H * (LTM_W / 4); item += 256) {
      const int j = item / (LTM_W / 4), q = item % (LTM_W / 4);
      const int yy = y0 + j, xq = x0 + 4 * q;
      if (yy >= H || xq >= W) continue;
      const double* a1 = &S.v1[j][q];  // needed column 4q + k is a1[lcol36(k)]
      const double* a2 = &S.v2[j][q];
      double h1 = 0.0, h2 = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        h1 += a1[lcol36(k)];
        h2 += a2[lcol36(k)];
      }
      const float4 xv4 = *reinterpret_cast<const float4*>(&S.x[s][j + LTM_R][LTM_XOFF + 4 * q]);
      const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
      float ol[4], os[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const float box = (float)h1, box2 = (float)h2;
        // avg, E[x^2] and var feed a cancellation, so they carry the reference's roundings bit for bit: box / n
        // is the correctly rounded quotient (div_const_rn), var keeps the reference's operation order.  What
        // follows the cancellation is not amplified: approximate square root and reciprocal (~2 ulp).
        const float avg = div_const_rn(box, n, inv_n);
        const float var = div_const_rn(box2, n, inv_n) - avg * avg + 1e-6f;
        const float sd = sqrt_approx(var) + eps;
        ol[m] = (xv[m] - avg) * rcp_approx(sd);
        os[m] = sd;
        if (m < 3) {
          h1 += a1[lcol36(m + K)] - a1[lcol36(m)];
          h2 += a2[lcol36(m + K)] - a2[lcol36(m)];
        }
      }
      const int64_t off = ((int64_t)img * H + yy) * W + xq;
      *reinterpret_cast<float4*>(lcn + off) = make_float4(ol[0], ol[1], ol[2], ol[3]);
      *reinterpret_cast<float4*>(sd_out + off) = make_float4(os[0], os[1], os[2], os[3]);
    }
    __syncthreads();
  }
}

extern int g_disable_tma;

static bool lcn_tma_launch(const float* x, float* lcn, float* sd, int64_t N, int64_t H, int64_t W, int r, float eps,
                           cudaStream_t st) {
  if (g_disable_tma || r != LTM_R || W % 4 || H < 32 || W < 32) return false;
  if ((reinterpret_cast<uintptr_t>(lcn) | reinterpret_cast<uintptr_t>(sd)) & 15) return false;
  const int64_t tiles = N * cdiv(H, LTM_H) * cdiv(W, LTM_W);
  if (tiles > INT32_MAX) return false;
  CUtensorMap m;
  if (!make_plane_tensor_map(&m, x, N, H, W, LTM_BW, LTM_BH)) return false;
  const size_t smem = sizeof(LcnSmem) + 128;
  if (cudaFuncSetAttribute(lcn_tma_kernel<LTM_R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)sms * 3);
  lcn_tma_kernel<LTM_R><<<grid, 256, smem, st>>>(m, lcn, sd, (int)H, (int)W, eps, (int)cdiv(W, LTM_W), (int)cdiv(H, LTM_H),
                                                (int)tiles);
  return true;
}

// any radius: one thread per pixel, direct (2r+1)^2 fp64 gather
template <typename T>
__global__ void __launch_bounds__(256)
lcn_generic_kernel(const T* __restrict__ x, T* __restrict__ lcn, T* __restrict__ sd_out, int64_t N, int H, int W,
                   int r, T eps) {
  const int64_t total = N * H * W;
  const T n = (T)((2 * r + 1) * (2 * r + 1));
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int w = idx % W, h = (idx / W) % H;
    const T* xp = x + (idx / ((int64_t)H * W)) * H * W;
    double s1 = 0.0, s2 = 0.0;
    for (int dh = -r; dh <= r; ++dh) {
      const T* row = xp + (int64_t)reflect(h + dh, H) * W;
      for (int dw = -r; dw <= r; ++dw) {
        const T v = __ldg(row + reflect(w + dw, W));
        s1 += (double)v;
        s2 += (double)(v * v);
      }
    }
    const T box = (T)s1, box2 = (T)s2;
    const T avg = box / n;
    const T var = box2 / n - avg * avg + (T)1e-6;
    const T sd = sqrt(var) + eps;
    lcn[idx] = (xp[(int64_t)h * W + w] - avg) / sd;
    sd_out[idx] = sd;
  }
}

template <typename T>
static int lcn_impl(const T* x, T* lcn, T* sd, int64_t N, int64_t H, int64_t W, int r, T eps, cudaStream_t st) {
  CTD_REQUIRE(N >= 0 && H >= 0 && W >= 0, "lcn: negative size");
  CTD_REQUIRE(r >= 0, "lcn: negative radius");
  if (N * H * W == 0) return CTD_OK;
  CTD_REQUIRE(x && lcn && sd, "lcn: null pointer");
  CTD_REQUIRE(H <= INT32_MAX && W <= INT32_MAX, "lcn: image too large");
  // torch.nn.ReflectionPad2d requires the padding to be smaller than the padded dimension
  CTD_REQUIRE(r < H && r < W, "lcn: radius %d must be smaller than the image (%lld x %lld)", r, (long long)H,
              (long long)W);
  if (sizeof(T) == 4 && !g_force_generic &&
      lcn_tma_launch(reinterpret_cast<const float*>(x), reinterpret_cast<float*>(lcn), reinterpret_cast<float*>(sd), N, H, W,
                     r, (float)eps, st)) {
  } else if (g_force_generic || r > L_RMAX || N > 65535) {
    const int grid = (int)std::min<int64_t>(cdiv(N * H * W, 256), 148 * 64);
    lcn_generic_kernel<T><<<grid, 256, 0, st>>>(x, lcn, sd, N, (int)H, (int)W, r, eps);
  } else {
    const int64_t strips = cdiv(W, LT_W);
    // enough CTAs for ~3 per SM, but runs of at least 16 rows so the 2r-row warm-up stays cheap
    int64_t nruns = std::max<int64_t>(1, std::min<int64_t>(cdiv(3 * 148, strips * N), cdiv(H, 16)));
    int64_t rows = cdiv(cdiv(H, nruns), L_RG) * L_RG;
    nruns = cdiv(H, rows);
    const auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const int vec = sizeof(T) == 4 && W % 4 == 0 && al(x) && al(lcn) && al(sd);
    lcn_strip_kernel<T><<<dim3((unsigned)strips, (unsigned)nruns, (unsigned)N), L_THREADS, 0, st>>>(
        x, lcn, sd, (int)H, (int)W, r, eps, (int)rows, vec);
  }
  count_launch();
  return check_launch("lcn");
}

}  // namespace ctd

CTD_API int ctd_lcn_f32(const float* x, float* lcn, float* sd, int64_t N, int64_t H, int64_t W, int r, float eps,
                           ctd_stream_t s) {
  return ctd::lcn_impl<float>(x, lcn, sd, N, H, W, r, eps, ctd::as_stream(s));
}
CTD_API int ctd_lcn_f64(const double* x, double* lcn, double* sd, int64_t N, int64_t H, int64_t W, int r,
                           double eps, ctd_stream_t s) {
  return ctd::lcn_impl<double>(x, lcn, sd, N, H, W, r, eps, ctd::as_stream(s));
}
