// PhotometricLoss census_mse / census_sad, block 9, fp32, C = 1: backward, or forward + backward (+ the caller's masked
// mean) in one pass (reference: torchext/ext/ext.h:201-344 through model/networks.py:376-377).
//
// OPT-IN (ctd_set_option("census_stream", 1)); the default stays photo_bwd_census9 (photometric.cu).  This kernel takes
// away everything the profile of that one showed around its row loop -- and measures the same (fused census_sad, batch
// 8 x 480 x 640: 138.1 us against 139.8 us; backward only 125.7 against 121.9; profiles/r02_census_stream_ab.json),
// because the row loop alone, with no staging, no barrier, no store and no second pass, already takes 129 us
// (option bits 4 | 8 | 16): it is the loop's instruction mix that bounds both kernels, not what surrounds it.
// Same pair folding as the tile kernel: grad_in[i] = eps/(2*81) * sum over the window of psi'(dd) * r1^3 *
// (M(i,q) * go[q] + go[i]), dd = h(es_i - es_q) - h(ta_i - ta_q).  What is different:
//  * persistent CTAs (three per SM), 64 x 8 tiles; a CTA's first three tiles are fixed, the rest are claimed from a
//    counter three iterations ahead of use;
//  * the staged tiles (es, ta replicate-clamped; grad_out zero outside the image) arrive through a four-stage
//    cp.async ring, two tiles ahead of the one being computed: ONE __syncthreads per tile, no load latency in front of it;
//  * two taps per instruction: a row's ten staged values are five aligned register pairs, and FADD2 / FMUL2 / FFMA2
//    (Blackwell's packed fp32 forms) do both lanes in one instruction (199 instead of 289 instructions per 18 taps; the
//    tile kernel has the same loop since round 2);
//  * ONE loop body for every tile: pixels ON the image border (clamp multiplicities M != 1) and pixels with a sign
//    decision the fast arithmetic cannot be trusted with (|dd| < SIGN_GUARD, the exact ties of flat regions included)
//    are noted in a shared-memory list and evaluated again, a warp per pixel, one iteration later -- the tile's stage
//    is still intact then (four stages, prefetch distance two), so this needs no barrier of its own and the flat regions'
//    clusters of such pixels are spread over all eight warps.  1.7 % of the pixels of the bench frames take that path.
// What the loop is bound by (tools/experiments/xu_mix_microbench*.cu, profiles/r02_xu_mix_microbench.txt): MUFU.RSQ alone
// sustains 15.6 of the XU pipe's 16 lanes / clk / SM; next to this loop's other instructions (per 8 MUFU: 20 packed and
// 5 scalar fp32, 4 LOP3, 2 FMNMX, 3 LDS.64) 10.9 -- and 162 rsqrt per pixel at 10.9 lanes are 125 us.  A packed
// instruction holds its pipe for two issue cycles, so 88 packed + 111 other instructions per row weigh 287 slots against
// the XU's 288 cycles: the loop needs both at once and gets 65 % of either.  Moving 4 of a row's 36 rsqrt to the FMA pipe
// (seed + three Newton steps, CTD_ST_SW_SINGLES) made it slower (146 us), as did 12 of 36 (168 us).
#include <algorithm>
#include <cstdio>

#include "ctd_common.cuh"

namespace ctd {

extern int g_force_generic;
int g_census_stream = 0;  // 1 = calls with a backward take this kernel instead of the tile kernel of photometric.cu (measured equal, see above); bits for experiments: 2 print occupancy, 4 no second pass, 8 no stores, 16 no staging / barriers (results invalid)

namespace {

constexpr int R9 = 4;
constexpr float INV81 = 1.0f / 81.0f;
constexpr float SIGN_GUARD = 3e-6f;   // fast dd is within ~1e-6 of the reference's; anything closer to zero is re-decided exactly
constexpr int ST_W = 64, ST_H = 8;    // output tile: warp = row, lane = two neighbouring pixels
constexpr int SE_W = ST_W + 2 * R9, SE_H = ST_H + 2 * R9;  // 72 x 16 staged
constexpr int ST_STAGES = 4;         // tile t is read in iteration t (row loop) and t + 1 (second pass); t + 2 is in flight
constexpr int ST_LISTS = 3;
constexpr int ST_SMEM = ST_STAGES * 3 * SE_H * SE_W * (int)sizeof(float);  // 55296
constexpr int ST_THREADS = 32 * ST_H;
constexpr int ST_CTAS_PER_SM = 3;

struct StreamGeom {
  int H, W, ntx, nty, ntiles;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void tile_coords(int t, const StreamGeom& g, int& n, int& x0, int& y0) {
  const int per = g.ntx * g.nty;
  n = t / per;
  const int r = t - n * per, ty = r / g.ntx;
  x0 = (r - ty * g.ntx) * ST_W;
  y0 = ty * ST_H;
}

typedef float Tile[SE_H][SE_W];

#ifndef CTD_ST_SW_SINGLES
#define CTD_ST_SW_SINGLES 0
#endif
#ifndef CTD_ST_SW_PAIRS
#define CTD_ST_SW_PAIRS 0
#endif
constexpr bool ST_SW_SINGLES = CTD_ST_SW_SINGLES != 0;  // the ninth taps' reciprocal square roots on the FMA pipe
constexpr int ST_SW_PAIRS = CTD_ST_SW_PAIRS;            // and those of this many tap pairs per pixel and row (experiments)

// 1 / sqrt(x) for a pair, x >= eps > 0: MUFU.RSQ twice, or (SW) on the FMA pipe -- the shift-and-subtract seed (3.4 %) and
// three packed Newton steps y <- y (1.5 - 0.5 x y^2): relative error ~2e-7, like rsqrt.approx's two units in the last place.
template <bool SW>
__device__ __forceinline__ float2 rsqrt_pair(float2 x) {
  if (!SW) return make_float2(rsqrt_approx(x.x), rsqrt_approx(x.y));
  float2 y = make_float2(__int_as_float(0x5f375a86 - (__float_as_int(x.x) >> 1)), __int_as_float(0x5f375a86 - (__float_as_int(x.y) >> 1)));
  const float2 h = __fmul2_rn(x, make_float2(-0.5f, -0.5f)), c15 = make_float2(1.5f, 1.5f);
#pragma unroll
  for (int i = 0; i < 3; ++i) y = __fmul2_rn(y, __ffma2_rn(__fmul2_rn(h, y), y, c15));
  return y;
}

// Thread (r = tid / 16, cj = tid % 16) fetches 16-byte pieces cj and cj + 16 (< 18) of staged row r of the three tiles.
// W % 4 == 0 and x0 - 4 is a multiple of 4: a piece lies inside the image row or outside it, never across.
__device__ __forceinline__ void issue_tile(Tile* S, int t, const StreamGeom& g, const float* __restrict__ es,
                                           const float* __restrict__ ta, const float* __restrict__ go, int tid) {
  if (t < g.ntiles) {
    int n, x0, y0;
    tile_coords(t, g, n, x0, y0);
    const int r = tid >> 4, cj = tid & 15;
    const int gy = y0 - R9 + r;
    const bool row_in = gy >= 0 && gy < g.H;
    const int64_t rowoff = (int64_t)n * g.H * g.W + (int64_t)clampi(gy, 0, g.H - 1) * g.W;
    const float *eb = es + rowoff, *tb = ta + rowoff, *gb = go + rowoff;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c4 = cj + 16 * j;
      if (c4 >= SE_W / 4) break;
      const int gx = x0 - R9 + 4 * c4;
      float *de = &S[0][r][4 * c4], *dt = &S[1][r][4 * c4], *dg = &S[2][r][4 * c4];
      if (gx >= 0 && gx < g.W) {
        cp_async16(de, eb + gx);
        cp_async16(dt, tb + gx);
        if (row_in) cp_async16(dg, gb + gx);
        else *reinterpret_cast<float4*>(dg) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {  // left of column 0 / right of column W - 1: the border column, replicated
        const int gxc = gx < 0 ? 0 : g.W - 1;
        const float ev = __ldg(eb + gxc), tv = __ldg(tb + gxc);
        *reinterpret_cast<float4*>(de) = make_float4(ev, ev, ev, ev);
        *reinterpret_cast<float4*>(dt) = make_float4(tv, tv, tv, tv);
        *reinterpret_cast<float4*>(dg) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  cp_async_commit();  // always: the group count per iteration is what cp_async_wait counts
}

__device__ __forceinline__ float sgnf(float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }

// Second pass over the listed pixels of the tile (n, x0, y0) staged in S, a warp per pixel: the whole window sum
// again, with the clamp multiplicity M(i,q) (ext.h:287-300: how many of centre q's offsets clamp onto pixel i; != 1 only
// ON the image border) and, for census_sad, the reference's own sign decisions wherever the fast dd is closer to zero
// than SIGN_GUARD: exact ties (both differences zero: flat regions, taps clamped onto the pixel itself) are zero terms,
// the others are decided in the reference's IEEE operation order (ext.h:245-246, 318-330).  Lane l owns the taps l,
// l + 32 and l + 64 (< 81, centre excluded): their offsets are worked out once per call, not per pixel.
template <int TYPE>
__device__ __forceinline__ void second_pass(const Tile* S, const unsigned short* list, unsigned count, float* __restrict__ gi, int n,
                                            int x0, int y0, int H, int W, float eps, int tid) {
  if ((unsigned)(tid >> 5) >= count) return;
  const int lane = tid & 31;
  int off[3];
  float dxf[3], dyf[3];
  bool ok[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int t = lane + 32 * i;
    ok[i] = t < 81 && t != 40;
    const int tt = ok[i] ? t : 40, ry = tt / 9, rx = tt - 9 * ry;
    off[i] = ry * SE_W + rx;
    dxf[i] = float(rx - R9);
    dyf[i] = float(ry - R9);
  }
  const float* E = &S[0][0][0];
  const float* T = &S[1][0][0];
  const float* G = &S[2][0][0];
  for (unsigned li = tid >> 5; li < count; li += ST_H) {
    const int p = list[li], yl = p / ST_W, xl = p - yl * ST_W;
    const int x = x0 + xl, y = y0 + yl;
    const int b = yl * SE_W + xl, c = b + R9 * SE_W + R9;
    const float ei = E[c], ti = T[c], gc = G[c];
    // multiplicity of the column / row offset d: base + slope * d
    const float sx = x == 0 ? -1.f : (x == W - 1 ? 1.f : 0.f), bx = sx != 0.f ? float(R9 + 1) : 1.f;
    const float sy = y == 0 ? -1.f : (y == H - 1 ? 1.f : 0.f), by = sy != 0.f ? float(R9 + 1) : 1.f;
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float e = E[b + off[i]], tt = T[b + off[i]], g = G[b + off[i]];
      const float des = ei - e, nta = tt - ti;
      const float r1 = rsqrt_approx(fmaf(des, des, eps)), r2 = rsqrt_approx(fmaf(nta, nta, eps));
      const float dd = fmaf(des, r1, nta * r2);
      const float r3 = (r1 * r1) * r1;
      const float gq = g * (fmaf(sx, dxf[i], bx) * fmaf(sy, dyf[i], by));
      float term;
      if (TYPE == 2) {
        term = (dd * r3) * (gq + gc);
      } else if (fabsf(dd) >= SIGN_GUARD) {
        term = __uint_as_float(__float_as_uint(r3) | (__float_as_uint(dd) & 0x80000000u)) * (gq + gc);
      } else if (des != 0.f || nta != 0.f) {
        const float dta = ti - tt;
        const float q1 = __fdiv_rn(des, __fsqrt_rn(__fadd_rn(__fmul_rn(des, des), eps)));
        const float q2 = __fdiv_rn(dta, __fsqrt_rn(__fadd_rn(__fmul_rn(dta, dta), eps)));
        const float d_tap = __fsub_rn(0.5f * __fadd_rn(1.f, q1), 0.5f * __fadd_rn(1.f, q2));
        const float d_ctr = __fsub_rn(0.5f * __fadd_rn(1.f, -q1), 0.5f * __fadd_rn(1.f, -q2));
        term = r3 * (sgnf(d_tap) * gq - sgnf(d_ctr) * gc);
      } else {
        term = 0.f;
      }
      d += ok[i] ? term : 0.f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) gi[(int64_t)n * H * W + (int64_t)y * W + x] = d * (0.5f * eps * INV81);
  }
}

template <int TYPE, bool FUSE>
__global__ void __launch_bounds__(ST_THREADS, ST_CTAS_PER_SM)
census_stream_kernel(const float* __restrict__ es, const float* __restrict__ ta, const float* __restrict__ go,
                     float* __restrict__ gi, float* __restrict__ out, const float* __restrict__ mask, const StreamGeom g, float eps,
                     double* __restrict__ partials, unsigned* __restrict__ ticket, float* __restrict__ sums2, unsigned* __restrict__ sched,
                     int dbg) {
  extern __shared__ __align__(16) float smem[];
  Tile(*S)[3] = reinterpret_cast<Tile(*)[3]>(smem);  // S[stage][es / ta / go]
  __shared__ unsigned short s_list[ST_LISTS][ST_W * ST_H];  // tile-local indices of the pixels for the second pass
  __shared__ unsigned s_cnt[ST_LISTS];
  __shared__ int s_tile[4];  // the CTA's tiles of iterations it .. it + 3 (slot = iteration & 3)
  const int tid = threadIdx.x, lane = tid & 31, yl = tid >> 5;
  const int stride = gridDim.x;
  const int H = g.H, W = g.W;
  float mnum = 0.f, mden = 0.f;
  if (tid < ST_LISTS) s_cnt[tid] = 0;
  // The first three tiles of a CTA are fixed (b, b + grid, b + 2 grid); the rest are claimed from a counter as CTAs get
  // to them, three iterations ahead of use: tiles with many second-pass pixels cost up to twice the others, and with a
  // fixed assignment the slowest SM was active 8 % longer than the average one.  sched[0] = next claim, sched[1] = CTAs done.
  if (tid < 3) s_tile[tid] = blockIdx.x + tid * stride;
  issue_tile(S[0], blockIdx.x, g, es, ta, go, tid);
  issue_tile(S[1], blockIdx.x + stride, g, es, ta, go, tid);
  int stage = 0, li = 0;         // of the tile computed in this iteration
  int pn = 0, px0 = 0, py0 = -1;  // the previous iteration's tile (py0 < 0: none)
#pragma unroll 1
  for (int it = 0;; ++it) {
    if (!(dbg & 16)) {
      cp_async_wait<1>();
      __syncthreads();  // this iteration's tile is complete and visible; every warp is done with the previous iteration
    }
    const int t = (dbg & 16) ? (int)blockIdx.x + it * stride : s_tile[it & 3];
    if (t >= g.ntiles) break;  // claims only grow: nothing after it either
    unsigned claim = 0;
    if (tid == 0) claim = atomicAdd(sched, 1u);  // used at the end of the iteration: the latency is not waited for
    if (!(dbg & 16)) issue_tile(S[(stage + 2) & 3], s_tile[(it + 2) & 3], g, es, ta, go, tid);  // stage last read in the previous iteration (second pass)
    if (tid == 0) s_cnt[li == ST_LISTS - 1 ? 0 : li + 1] = 0;                 // consumed in the previous iteration, filled in the next
    const Tile &Es = S[stage][0], &Ts = S[stage][1], &Gs = S[stage][2];
    int n, x0, y0;
    tile_coords(t, g, n, x0, y0);
    const int gy = y0 + yl, gx = x0 + 2 * lane;
    if (gy < H) {  // warp-uniform
      const int64_t o = (int64_t)n * H * W + (int64_t)gy * W + gx;
      float2 mk = make_float2(0.f, 0.f);
      if (FUSE && mask != nullptr && gx < W) mk = __ldg(reinterpret_cast<const float2*>(mask + o));
      const float2 ecv = *reinterpret_cast<const float2*>(&Es[yl + R9][2 * lane + R9]);
      const float2 tcv = *reinterpret_cast<const float2*>(&Ts[yl + R9][2 * lane + R9]);
      const float2 gcv = *reinterpret_cast<const float2*>(&Gs[yl + R9][2 * lane + R9]);
      const float ec[2] = {ecv.x, ecv.y}, tc[2] = {tcv.x, tcv.y}, gc[2] = {gcv.x, gcv.y};
      float2 acc2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      float2 fac2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      float2 accS = make_float2(0.f, 0.f), facS = make_float2(0.f, 0.f);  // the ninth taps: .x pixel 0, .y pixel 1
      float near0[2] = {1.f, 1.f};
      const float2 eps2 = make_float2(eps, eps);
      // Pixel 0 takes the staged pairs (0,1) .. (6,7) as tap pairs and column 8 alone, pixel 1 takes (2,3) .. (8,9) and
      // column 1 alone; .x / .y of acc2 and fac2 are the even- and odd-tap partial sums.
#pragma unroll 1
      for (int dy = 0; dy < 9; ++dy) {
        float2 e2[5], t2[5], g2[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          e2[q] = reinterpret_cast<const float2*>(&Es[yl + dy][2 * lane])[q];
          t2[q] = reinterpret_cast<const float2*>(&Ts[yl + dy][2 * lane])[q];
          g2[q] = reinterpret_cast<const float2*>(&Gs[yl + dy][2 * lane])[q];
        }
        const bool crow = dy == R9;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float2 ec2 = make_float2(ec[k], ec[k]), tc2 = make_float2(tc[k], tc[k]), gc2 = make_float2(gc[k], gc[k]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int dx = 2 * q + k;  // the pair's first tap (0 .. 8 <-> -4 .. 4)
            const float2 ev = e2[q + k], tv = t2[q + k];
            const float2 des = __fadd2_rn(ec2, make_float2(-ev.x, -ev.y));
            const float2 nta = __fadd2_rn(tv, make_float2(-tc2.x, -tc2.y));  // -(ta_i - ta_q)
            const float2 s1 = __ffma2_rn(des, des, eps2), s2 = __ffma2_rn(nta, nta, eps2);
            const float2 r1 = q < ST_SW_PAIRS ? rsqrt_pair<true>(s1) : rsqrt_pair<false>(s1);
            const float2 r2 = q < ST_SW_PAIRS ? rsqrt_pair<true>(s2) : rsqrt_pair<false>(s2);
            const float2 dd = __ffma2_rn(des, r1, __fmul2_rn(nta, r2));
            const float2 r3 = __fmul2_rn(__fmul2_rn(r1, r1), r1);
            const float2 gs = __fadd2_rn(g2[q + k], gc2);
            if (FUSE) {
              if (TYPE == 2) fac2[k] = __ffma2_rn(dd, dd, fac2[k]);
              else fac2[k] = __fadd2_rn(fac2[k], make_float2(fabsf(dd.x), fabsf(dd.y)));
            }
            if (TYPE == 2) {
              acc2[k] = __ffma2_rn(__fmul2_rn(dd, r3), gs, acc2[k]);
            } else {
              // sign(dd) * r3 by OR-ing dd's sign bit into r3 > 0; the centre tap (dd = +0) is taken out after the loop
              const float2 sr3 = make_float2(__uint_as_float(__float_as_uint(r3.x) | (__float_as_uint(dd.x) & 0x80000000u)),
                                             __uint_as_float(__float_as_uint(r3.y) | (__float_as_uint(dd.y) & 0x80000000u)));
              acc2[k] = __ffma2_rn(sr3, gs, acc2[k]);
              float m0 = fabsf(dd.x), m1 = fabsf(dd.y);
              if (dx == R9) m0 = crow ? 1.f : m0;
              if (dx + 1 == R9) m1 = crow ? 1.f : m1;
              near0[k] = fminf(near0[k], fminf(m0, m1));
            }
          }
        }
        {  // the ninth taps as one pair: staged column 8 for pixel 0 (dx = 8) in .x, staged column 1 for pixel 1 (dx = 0) in .y.
           // Their four reciprocal square roots (of 36 per row) are done on the FMA pipe: the loop is bound by the XU pipe
           // (MUFU), whose throughput next to this loop's other instructions is ~70 % of its peak (xu_mix_microbench*.cu).
          const float2 ev = make_float2(e2[4].x, e2[0].y), tv = make_float2(t2[4].x, t2[0].y), gv = make_float2(g2[4].x, g2[0].y);
          const float2 des = __fadd2_rn(ecv, make_float2(-ev.x, -ev.y));
          const float2 nta = __fadd2_rn(tv, make_float2(-tcv.x, -tcv.y));
          const float2 r1 = rsqrt_pair<ST_SW_SINGLES>(__ffma2_rn(des, des, eps2)), r2 = rsqrt_pair<ST_SW_SINGLES>(__ffma2_rn(nta, nta, eps2));
          const float2 dd = __ffma2_rn(des, r1, __fmul2_rn(nta, r2));
          const float2 r3 = __fmul2_rn(__fmul2_rn(r1, r1), r1);
          const float2 gs = __fadd2_rn(gv, gcv);
          if (FUSE) {
            if (TYPE == 2) facS = __ffma2_rn(dd, dd, facS);
            else facS = __fadd2_rn(facS, make_float2(fabsf(dd.x), fabsf(dd.y)));
          }
          if (TYPE == 2) {
            accS = __ffma2_rn(__fmul2_rn(dd, r3), gs, accS);
          } else {
            const float2 sr3 = make_float2(__uint_as_float(__float_as_uint(r3.x) | (__float_as_uint(dd.x) & 0x80000000u)),
                                           __uint_as_float(__float_as_uint(r3.y) | (__float_as_uint(dd.y) & 0x80000000u)));
            accS = __ffma2_rn(sr3, gs, accS);
            near0[0] = fminf(near0[0], fabsf(dd.x));
            near0[1] = fminf(near0[1], fabsf(dd.y));
          }
        }
      }
      if (gx < W && (!(dbg & 8) || accS.x == 12345.f)) {  // W % 4 == 0 and gx even: both pixels or neither
        const float r0 = rsqrt_approx(eps);
        const float sb = 0.5f * eps * INV81;
        float r[2];
        const bool yb = gy == 0 || gy == H - 1;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          r[k] = acc2[k].x + acc2[k].y + (k == 0 ? accS.x : accS.y);
          if (TYPE == 3) r[k] -= (r0 * r0) * r0 * (2.f * gc[k]);  // the centre tap: +r0^3 * (go_i + go_i) above, nothing in the reference
          // ON the image border (clamp multiplicities) or a sign decision too close to call: second pass, next iteration
          if (yb || gx + k == 0 || gx + k == W - 1 || (TYPE == 3 && near0[k] < SIGN_GUARD))
            s_list[li][atomicAdd(&s_cnt[li], 1u)] = (unsigned short)(yl * ST_W + 2 * lane + k);
        }
        *reinterpret_cast<float2*>(gi + o) = make_float2(r[0] * sb, r[1] * sb);
        if (FUSE) {
          const float sf = (TYPE == 2 ? 0.25f : 0.5f) * INV81;
          const float l0 = (fac2[0].x + fac2[0].y + facS.x) * sf, l1 = (fac2[1].x + fac2[1].y + facS.y) * sf;
          *reinterpret_cast<float2*>(out + o) = make_float2(l0, l1);
          if (mask != nullptr) {
            mnum = fmaf(mk.x, l0, mnum);
            mnum = fmaf(mk.y, l1, mnum);
            mden += mk.x + mk.y;
          }
        }
      }
    }
    if (py0 >= 0 && !(dbg & 4)) {  // the previous tile's list: complete since this iteration's barrier, its stage untouched until the next
      const int pl = li == 0 ? ST_LISTS - 1 : li - 1;
      second_pass<TYPE>(S[(stage + 3) & 3], s_list[pl], s_cnt[pl], gi, pn, px0, py0, H, W, eps, tid);
    }
    pn = n;
    px0 = x0;
    py0 = y0;
    stage = (stage + 1) & 3;
    li = li == ST_LISTS - 1 ? 0 : li + 1;
    if (tid == 0) s_tile[(it + 3) & 3] = (int)min(3u * (unsigned)stride + claim, (unsigned)g.ntiles);  // slot of iteration it - 1: free
  }
  cp_async_wait<0>();
  if (tid == 0 && atomicAdd(sched + 1, 1u) == (unsigned)stride - 1) {  // last CTA out: every claim has been made
    sched[0] = 0;
    sched[1] = 0;
  }
  if (py0 >= 0 && !(dbg & 4)) {
    const int pl = li == 0 ? ST_LISTS - 1 : li - 1;
    second_pass<TYPE>(S[(stage + 3) & 3], s_list[pl], s_cnt[pl], gi, pn, px0, py0, H, W, eps, tid);
  }
  if (FUSE && mask != nullptr) finish_masked_sums((double)mnum, (double)mden, partials, ticket, sums2);  // block-uniform
}

int sm_count() {
  int dev = 0, v = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
  return v > 0 ? v : 148;
}

template <int TYPE, bool FUSE>
bool launch(const float* es, const float* ta, const float* go, float* gi, float* out, const float* mask, const StreamGeom& g,
            float eps, unsigned grid, const MsSlot& ms, float* sums2, cudaStream_t st) {
  auto kernel = census_stream_kernel<TYPE, FUSE>;
  // three CTAs x 58 KB: needs the large carve-out (set per call: the attributes are per device)
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (g_census_stream & 2) {  // diagnostics
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, ST_THREADS, ST_SMEM);
    fprintf(stderr, "census_stream: %d CTAs per SM, grid %u\n", nb, grid);
  }
  kernel<<<grid, ST_THREADS, ST_SMEM, st>>>(es, ta, go, gi, out, mask, g, eps, ms.partials, ms.ticket, sums2, ms.ticket + 1, g_census_stream);
  return true;
}

}  // namespace

// Backward (gi, needs go), optionally with the forward (out) and the masked-mean terms (mask, sums2) in the same pass.
// Returns false when this path does not take the call (nothing launched): C != 1, rows that are not 16-byte aligned,
// forward only.
bool census_stream_launch(const float* es, const float* ta, const float* go, float* out, float* gi, const float* mask, float* sums2,
                          int64_t B, int64_t C, int64_t H, int64_t W, int type, float eps, cudaStream_t st) {
  if (!g_census_stream || g_force_generic || C != 1 || B < 1 || H < 9 || W < 9 || (type != 2 && type != 3)) return false;
  if (!es || !ta || !go || !gi) return false;
  if (mask && (!out || !sums2)) return false;
  const uintptr_t al = reinterpret_cast<uintptr_t>(es) | reinterpret_cast<uintptr_t>(ta) | reinterpret_cast<uintptr_t>(go) |
                       reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gi) | reinterpret_cast<uintptr_t>(mask);
  if (W % 4 != 0 || (al & 15) != 0) return false;
  StreamGeom g;
  g.H = (int)H;
  g.W = (int)W;
  g.ntx = (int)cdiv(W, ST_W);
  g.nty = (int)cdiv(H, ST_H);
  const int64_t ntiles = B * g.ntx * g.nty;
  if (ntiles > (int64_t)INT32_MAX / 4 || H * W >= (int64_t)1 << 31) return false;
  g.ntiles = (int)ntiles;
  const unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)ST_CTAS_PER_SM * sm_count());
  MsSlot ms = {nullptr, nullptr, nullptr};
  // every variant takes a slot: its spare header words (zero between launches, like the ticket) are the tile scheduler's
  if (!ms_acquire(grid, st, &ms)) return false;
  bool ok;
  if (type == 2) ok = out ? launch<2, true>(es, ta, go, gi, out, mask, g, eps, grid, ms, sums2, st)
                          : launch<2, false>(es, ta, go, gi, out, mask, g, eps, grid, ms, sums2, st);
  else ok = out ? launch<3, true>(es, ta, go, gi, out, mask, g, eps, grid, ms, sums2, st)
                : launch<3, false>(es, ta, go, gi, out, mask, g, eps, grid, ms, sums2, st);
  ms_release(&ms, st);
  if (ok) count_launch();
  return ok;
}

}  // namespace ctd
