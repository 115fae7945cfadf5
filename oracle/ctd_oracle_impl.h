/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's torchext functors.
 *
 * This header is a "template": ctd_oracle.c includes it twice, once with REAL=float and
 * once with REAL=double, so each function exists as ctdo_<name>_f32 / ctdo_<name>_f64.
 * Every function states the reference lines it follows (paths relative to
 * /root/reference/).  It is a restatement written for this repo, not a copy: plain C,
 * explicit type promotions where the C++ source relies on implicit ones, explicit
 * x86 float->int conversion semantics where the C++ source relies on the hardware.
 *
 * Promotion notes that decide the last bit (validated bit-for-bit against the compiled
 * reference, tests/test_oracle.py):
 *   - `0.5 * expr`, `expr + 1e-8`, `u + 0.5` are evaluated in double (double literals)
 *     and rounded once on assignment to REAL;
 *   - `x / block_size2` divides by the integer converted to REAL (a true division per
 *     term, not a multiplication by a reciprocal);
 *   - no fused multiply-add anywhere (the reference build has none): compile this file
 *     with -ffp-contract=off and without -march flags.
 */

#ifndef REAL
#error "include from ctd_oracle.c"
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(CAT(ctdo_, name), SUFFIX)

static inline long FN(clampl)(long v, long lo, long hi) {
  if (v < lo) v = lo;
  if (v > hi) v = hi;
  return v;
}

/* soft Heaviside of the census transform, torchext/ext/ext.h:245-246 (and 321-322):
 * h(x) = 0.5 * (1 + x / sqrt(x*x + eps)); the 0.5 is a double literal. */
static inline REAL FN(soft_step)(REAL x, REAL eps) {
  REAL q = x / REAL_SQRT(x * x + eps);
  REAL one_plus = (REAL)1 + q;
  return (REAL)(0.5 * (double)one_plus);
}

/* torchext/ext/ext.h:201-266 PhotometricLossForward::operator(), driven by the serial loop
 * of torchext/ext/ext_cpu.cpp:7-12,110-145.  es, ta: [B,C,H,W]; out: [B,1,H,W].
 * type: 0 mse, 1 sad, 2 census_mse, 3 census_sad (ext.h:196-199).  Any other type leaves
 * `out` untouched, like the reference (which returns uninitialised memory). */
void FN(photometric_fwd)(const REAL* es, const REAL* ta, REAL* out, int B, int C, int H, int W,
                         int bs, int type, float eps_f) {
  const REAL eps = (REAL)eps_f;          /* ext_cpu.cpp:110: eps arrives as a C float */
  const int bs2 = bs * bs;
  const int half = bs / 2;
  if (type < 0 || type > 3) return;
  for (int n = 0; n < B; ++n)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        REAL acc = 0;
        for (int t = 0; t < bs2; ++t) {
          int hh = (int)FN(clampl)(h + t / bs - half, 0, H - 1);   /* ext.h:227-233 */
          int ww = (int)FN(clampl)(w + t % bs - half, 0, W - 1);
          for (int c = 0; c < C; ++c) {
            long tap = (((long)n * C + c) * H + hh) * W + ww;
            REAL d;
            if (type <= 1) {
              d = es[tap] - ta[tap];                               /* ext.h:237-238 */
            } else {
              long ctr = (((long)n * C + c) * H + h) * W + w;       /* ext.h:246-251 */
              REAL des = es[tap] - es[ctr];
              REAL dta = ta[tap] - ta[ctr];
              d = FN(soft_step)(des, eps) - FN(soft_step)(dta, eps);
            }
            if (type == 0 || type == 2) acc += d * d / (REAL)bs2;  /* ext.h:240,255 */
            else acc += REAL_FABS(d) / (REAL)bs2;                  /* ext.h:243,258 */
          }
        }
        out[((long)n * H + h) * W + w] = acc;
      }
}

/* torchext/ext/ext.h:268-344 PhotometricLossBackward::operator(), serial order of
 * ext_cpu.cpp:7-12 (pixel ascending, tap ascending, channel ascending), zero-initialised
 * grad_in as ext_cpu.cpp:158.  grad_out: [B,1,H,W]; grad_in: [B,C,H,W]. */
void FN(photometric_bwd)(const REAL* es, const REAL* ta, const REAL* grad_out, REAL* grad_in,
                         int B, int C, int H, int W, int bs, int type, float eps_f) {
  const REAL eps = (REAL)eps_f;
  const int bs2 = bs * bs;
  const int half = bs / 2;
  for (long i = 0; i < (long)B * C * H * W; ++i) grad_in[i] = 0;
  if (type < 0 || type > 3) return;
  for (int n = 0; n < B; ++n)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        const REAL go = grad_out[((long)n * H + h) * W + w];
        for (int t = 0; t < bs2; ++t) {
          int hh = (int)FN(clampl)(h + t / bs - half, 0, H - 1);
          int ww = (int)FN(clampl)(w + t % bs - half, 0, W - 1);
          for (int c = 0; c < C; ++c) {
            long tap = (((long)n * C + c) * H + hh) * W + ww;
            if (type <= 1) {
              REAL d = es[tap] - ta[tap];
              REAL g;
              if (type == 0) g = (REAL)2 * d;                       /* ext.h:309 */
              else g = d < 0 ? (REAL)-1 : (d > 0 ? (REAL)1 : (REAL)0); /* ext.h:312 */
              g = g / (REAL)bs2 * go;                               /* ext.h:314 */
              grad_in[tap] += g;
            } else {
              long ctr = (((long)n * C + c) * H + h) * W + w;
              REAL des = es[tap] - es[ctr];
              REAL dta = ta[tap] - ta[ctr];
              REAL d = FN(soft_step)(des, eps) - FN(soft_step)(dta, eps);
              REAL gl;
              if (type == 2) gl = (REAL)2 * d;                      /* ext.h:327 */
              else gl = d < 0 ? (REAL)-1 : (d > 0 ? (REAL)1 : (REAL)0);
              gl = gl / (REAL)bs2;                                  /* ext.h:332 */
              REAL s = des * des + eps;                             /* ext.h:334-335 */
              REAL gh = (REAL)((0.5 * (double)eps) / (double)REAL_SQRT(s * s * s));
              REAL g = go * gl * gh;                                /* ext.h:337 */
              grad_in[tap] += g;
              grad_in[ctr] += -g;
            }
          }
        }
      }
}

/* torchext/ext/ext.h:120-191 XCorrVolFunctor (ext_cpu.cpp:88-105).  in0, in1: [C,H,W];
 * out: [D,H,W].  Two passes per channel: window means, then centred dot / sigmas. */
void FN(xcorrvol)(const REAL* in0, const REAL* in1, REAL* out, long C, long H, long W, long D,
                  long bs) {
  const long bs2 = bs * bs;
  const long half = bs / 2;
  for (long d = 0; d < D; ++d)
    for (long h = 0; h < H; ++h)
      for (long w = 0; w < W; ++w) {
        REAL val = 0;
        for (long c = 0; c < C; ++c) {
          REAL mu0 = 0, mu1 = 0;
          for (long bh = 0; bh < bs; ++bh) {
            long hh = FN(clampl)(h + bh - half, 0, H - 1);
            for (long bw = 0; bw < bs; ++bw) {
              long w0 = w + bw - half;
              long w1 = w0 - d;            /* ext.h:150-151: shift BEFORE clamping */
              w0 = FN(clampl)(w0, 0, W - 1);
              w1 = FN(clampl)(w1, 0, W - 1);
              mu0 += in0[(c * H + hh) * W + w0] / (REAL)bs2;
              mu1 += in1[(c * H + hh) * W + w1] / (REAL)bs2;
            }
          }
          REAL s0 = 0, s1 = 0, dot = 0;
          for (long bh = 0; bh < bs; ++bh) {
            long hh = FN(clampl)(h + bh - half, 0, H - 1);
            for (long bw = 0; bw < bs; ++bw) {
              long w0 = w + bw - half;
              long w1 = w0 - d;
              w0 = FN(clampl)(w0, 0, W - 1);
              w1 = FN(clampl)(w1, 0, W - 1);
              REAL v0 = in0[(c * H + hh) * W + w0] - mu0;
              REAL v1 = in1[(c * H + hh) * W + w1] - mu1;
              dot += v0 * v1;
              s0 += v0 * v0;
              s1 += v1 * v1;
            }
          }
          REAL norm = (REAL)((double)REAL_SQRT(s0 * s1) + 1e-8);   /* ext.h:185 */
          val += dot / norm;
        }
        out[(d * H + h) * W + w] = val;
      }
}

/* x86-64 cvttsd2si semantics for `int u0 = <double>` (ext.h:90-91): truncate toward zero;
 * NaN and anything outside int range give INT_MIN ("integer indefinite"). */
static inline int FN(x86_double_to_int)(double v) {
  if (!(v > -2147483649.0 && v < 2147483648.0)) return (int)(-2147483647 - 1);
  return (int)v;
}

/* torchext/ext/ext.h:65-117 ProjNNFunctor (ext_cpu.cpp:59-85).  xyz0, xyz1: [B,H,W,3];
 * K: 3x3 row-major; out: int64 [B,H,W] global flat index into xyz1's pixels, or -1. */
void FN(proj_nn)(const REAL* xyz0, const REAL* xyz1, const REAL* K, long long* out, long B,
                 long H, long W, long ps) {
  for (long i = 0; i < B * H * W; ++i) {
    const long b = i / (H * W);
    const REAL x = xyz0[i * 3 + 0], y = xyz0[i * 3 + 1], z = xyz0[i * 3 + 2];
    const REAL den = K[6] * x + K[7] * y + K[8] * z;               /* ext.h:87-89 */
    const REAL u = (K[0] * x + K[1] * y + K[2] * z) / den;
    const REAL v = (K[3] * x + K[4] * y + K[5] * z) / den;
    const int u0 = FN(x86_double_to_int)((double)u + 0.5);          /* ext.h:90-91 */
    const int v0 = FN(x86_double_to_int)((double)v + 0.5);
    long long best = -1;
    REAL best_d = (REAL)1e9;
    for (int p = 0; p < ps * ps; ++p) {
      int pu = (int)(p % ps), pv = (int)(p / ps);
      /* ext.h:98-99: int + int - long evaluated in long, then narrowed to int */
      int u1 = (int)((long)u0 + pu - ps / 2);
      int v1 = (int)((long)v0 + pv - ps / 2);
      if (u1 >= 0 && v1 >= 0 && u1 < W && v1 < H) {
        long j = (b * H + v1) * W + u1;
        const REAL* q = xyz1 + j * 3;
        REAL dd = (x - q[0]) * (x - q[0]) + (y - q[1]) * (y - q[1]) + (z - q[2]) * (z - q[2]);
        if (dd < best_d) { best_d = dd; best = j; }
      }
    }
    out[i] = best;
  }
}

/* torchext/ext/ext.h:13-46 NNFunctor<T,3> (ext_cpu.cpp:14-36).  in0: [N0,3], in1: [N1,3];
 * out: int64 [N0], first strict minimum of the squared distance below 1e9, else -1. */
void FN(nn)(const REAL* in0, const REAL* in1, long long* out, long N0, long N1) {
  for (long i = 0; i < N0; ++i) {
    const REAL* a = in0 + i * 3;
    REAL best_d = (REAL)1e9;
    long long best = -1;
    for (long j = 0; j < N1; ++j) {
      const REAL* b = in1 + j * 3;
      REAL dist = 0;
      for (int k = 0; k < 3; ++k) {
        REAL df = a[k] - b[k];
        dist += df * df;
      }
      if (dist < best_d) { best_d = dist; best = j; }
    }
    out[i] = best;
  }
}

/* model/networks.py:507-533 LCN.tforward.  x: [N,H,W] (the single channel squeezed);
 * lcn, std: [N,H,W].  Reflection padding without edge repeat (torch ReflectionPad2d),
 * box sums over (2r+1)^2 taps.  The reference's box sums come out of a library
 * convolution whose summation order is unspecified, so they are accumulated in double
 * here and rounded to REAL once (the "exact box sum" reading of networks.py:524-528);
 * everything after the box sums follows the reference's per-element formula in REAL. */
void FN(lcn)(const REAL* x, REAL* lcn, REAL* std, long N, long H, long W, long r, REAL eps) {
  const REAL n = (REAL)((2 * r + 1) * (2 * r + 1));
  for (long i = 0; i < N; ++i)
    for (long h = 0; h < H; ++h)
      for (long w = 0; w < W; ++w) {
        double s1 = 0, s2 = 0;
        for (long dh = -r; dh <= r; ++dh) {
          long hh = h + dh;
          if (hh < 0) hh = -hh;
          if (hh > H - 1) hh = 2 * (H - 1) - hh;
          for (long dw = -r; dw <= r; ++dw) {
            long ww = w + dw;
            if (ww < 0) ww = -ww;
            if (ww > W - 1) ww = 2 * (W - 1) - ww;
            REAL v = x[(i * H + hh) * W + ww];
            REAL v2 = v * v;                     /* data**2 is formed in REAL, networks.py:528 */
            s1 += (double)v;
            s2 += (double)v2;
          }
        }
        REAL box = (REAL)s1, box2 = (REAL)s2;
        REAL avg = box / n;                                           /* networks.py:526 */
        REAL var = box2 / n - avg * avg + (REAL)1e-6;                 /* networks.py:530 */
        REAL sd = REAL_SQRT(var) + eps;                               /* networks.py:530-531 */
        long o = (i * H + h) * W + w;
        lcn[o] = (x[o] - avg) / sd;                                   /* networks.py:533 */
        std[o] = sd;
      }
}

#undef FN
#undef CAT
#undef CAT_
