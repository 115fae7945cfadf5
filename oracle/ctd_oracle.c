/*
 * TEST INFRASTRUCTURE ONLY -- the CPU oracle of connecting_the_dots_b200.
 *
 * A plain-C restatement of the reference's torchext functors (torchext/ext/ext.h,
 * driven as torchext/ext/ext_cpu.cpp drives them) and of model/networks.py's LCN.  Only
 * tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this
 * library; the product (connecting_the_dots_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here bit-for-bit
 * against the unmodified reference extension compiled into oracle/_ref/ (when present)
 * and against the golden vectors in tests/golden/ that were generated from that same
 * reference build by tests/golden/make_golden.py.  The LCN restatement is pinned against
 * the reference's own torch LCN module to the tolerance stated in that test (the
 * reference's fp32 convolution order is unspecified, see ctd_oracle_impl.h).
 *
 * Build: `make -C oracle` (gcc -O3 -ffp-contract=off, no -march: the reference build has
 * no FMA instructions and neither may this).
 */
#include <math.h>
#include <stdint.h>

#define REAL float
#define SUFFIX _f32
#define REAL_SQRT sqrtf
#define REAL_FABS fabsf
#include "ctd_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef REAL_SQRT
#undef REAL_FABS

#define REAL double
#define SUFFIX _f64
#define REAL_SQRT sqrt
#define REAL_FABS fabs
#include "ctd_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef REAL_SQRT
#undef REAL_FABS

/* torchext/ext/ext.h:48-63 CrossCheckFunctor (ext_cpu.cpp:39-56).  in0: int64 [N0],
 * in1: int64 [N1] (N1 is never consulted, as in the reference); out: uint8 [N0].
 * `int idx1 = in0[i]` narrows int64 -> int32 by truncation (ext.h:59). */
void ctdo_crosscheck(const long long* in0, const long long* in1, uint8_t* out, long N0, long N1) {
  (void)N1;
  for (long i = 0; i < N0; ++i) {
    int j = (int)(uint32_t)(uint64_t)in0[i];
    out[i] = (uint8_t)(j >= 0 && in1[j] >= 0 && (long long)i == in1[j]);
  }
}

/* data/lcn/lcn.pyx:16-58 `normalize(img, kernel_size, epsilon)`: the OFFLINE local contrast normalisation of the data
 * generator (create_syn_data.py:182).  img [M,N] fp32 -> lcn, std [M,N]; a border of ks pixels stays 0 (lcn.pyx:22-23,
 * 36-37).  All arithmetic in float, sums in the reference's order (rows outer, columns inner, lcn.pyx:40-50); the
 * reference's `sqrt` is C's double sqrt applied to a float and rounded back to float (lcn.pyx:5-6), which equals sqrtf.
 * Pinned bit-for-bit against the Cython build by tests/golden/lcn_cython.npz (make_golden_lcn_cython.py). */
void ctdo_lcn_cython_f32(const float* img, float* lcn, float* sd, long M, long N, long ks, float eps) {
  const float num = (float)((ks * 2 + 1) * (ks * 2 + 1));          /* lcn.pyx:33 */
  for (long i = 0; i < M * N; ++i) lcn[i] = sd[i] = 0.0f;            /* lcn.pyx:22-23 */
  for (long m = ks; m < M - ks; ++m)
    for (long n = ks; n < N - ks; ++n) {
      float mean = 0.0f;
      for (long i = -ks; i <= ks; ++i)
        for (long j = -ks; j <= ks; ++j) mean += img[(m + i) * N + n + j];      /* lcn.pyx:40-43 */
      mean = mean / num;
      float stddev = 0.0f;
      for (long i = -ks; i <= ks; ++i)
        for (long j = -ks; j <= ks; ++j) {
          const float d = img[(m + i) * N + n + j] - mean;
          stddev = stddev + d * d;                                                 /* lcn.pyx:47-50 */
        }
      stddev = (float)sqrt((double)(stddev / num));                                /* lcn.pyx:51 */
      lcn[m * N + n] = (img[m * N + n] - mean) / (stddev + eps);                   /* lcn.pyx:54 */
      sd[m * N + n] = stddev;                                                      /* lcn.pyx:55 */
    }
}
