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
