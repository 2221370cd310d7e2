/*
 * synth_oracle.c -- CPU twin of the synthetic-workload generator (test infrastructure).
 *
 * The reference has no benchmark harness; the synthetic workload is this
 * repo's measurement fixture.  The product generates it on the GPU
 * (spark-agd_b200/csrc/synth.cu); this file restates the SAME SPEC independently on the CPU so
 * that tests can check the two agree bit-for-bit and so that the oracle can consume identical
 * inputs.  Spec (all integer arithmetic until the final scaling, hence exactly reproducible):
 *
 *   Philox4x32-10, key = (seed_lo, seed_hi), counter = (c0, c1, c2, stream)
 *   X[i][j]   : counter (i_lo, i_hi, j/2, 1) -> r[0..3];  pair = (r0,r1) for even j, (r2,r3) for odd j
 *               t = lo16(a)+hi16(a)+lo16(b)+hi16(b) - 131070   (Irwin-Hall-4, integer, zero mean)
 *               X = (float)t * (float)(sqrt(3)/65536)          (unit variance, |X| < 3.47)
 *   w_true[j] : counter (j, 0, 0, 2) -> t from (r0,r1);  w = ((double)t * (sqrt(3)/65536)) / sqrt(d)
 *   u_i       : counter (i_lo, i_hi, 0, 3);  u = ((r0>>5)*2^26 + (r1>>6) + 0.5) * 2^-53  in (0,1)
 *   e_i       : counter (i_lo, i_hi, 0, 4) -> t from (r0,r1);  e = (double)t * (sqrt(3)/65536)
 *   labels    : logistic      y = 1[ x_i.w_true + log(u) - log(1-u) > 0 ]
 *               least squares y = x_i.w_true + 0.1 * e
 *               hinge         y = 1[ x_i.w_true > 0 ], flipped when u < 0.05
 *   (x_i.w_true is accumulated in fp64; its summation order is implementation-defined, so a label
 *    can differ between implementations only when the margin is within ~1e-13 of zero.)
 */
#include <math.h>
#include <stdint.h>

#include "agd_oracle.h"

#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

static void philox4x32_10(uint32_t k0, uint32_t k1, const uint32_t c_in[4], uint32_t out[4]) {
  uint32_t c0 = c_in[0], c1 = c_in[1], c2 = c_in[2], c3 = c_in[3];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)PHILOX_M0 * c0, p1 = (uint64_t)PHILOX_M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += PHILOX_W0; k1 += PHILOX_W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static int32_t irwin_hall4(uint32_t a, uint32_t b) {
  return (int32_t)((a & 0xffffu) + (a >> 16) + (b & 0xffffu) + (b >> 16)) - 131070;
}

static const double kScale64 = 1.7320508075688772 / 65536.0; /* sqrt(3)/65536 */

static void synth_row_f32(uint32_t k0, uint32_t k1, float scale, uint64_t i, int32_t d, float *x) {
  for (int32_t j = 0; j < d; j += 2) {
    uint32_t c[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)(j >> 1), 1u}, o[4];
    philox4x32_10(k0, k1, c, o);
    x[j] = (float)irwin_hall4(o[0], o[1]) * scale;
    if (j + 1 < d) x[j + 1] = (float)irwin_hall4(o[2], o[3]) * scale;
  }
}

void oracle_synth_dense_f32(uint64_t seed, int64_t row0, int64_t rows, int32_t d, float *X) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const float scale = (float)kScale64;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int64_t r = 0; r < rows; ++r) synth_row_f32(k0, k1, scale, (uint64_t)(row0 + r), d, X + r * (int64_t)d);
}

/* Same values, written partition by partition in the thread mapping oracle_smooth folds them in (partition p on
 * thread p % threads): on a NUMA host every page of X is first touched by the thread that will read it. */
void oracle_synth_dense_f32_placed(uint64_t seed, int64_t row0, int64_t rows, int32_t d, float *X, int partitions,
                                   int threads) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const float scale = (float)kScale64;
  const int P = partitions < 1 ? 1 : partitions;
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) num_threads(threads > 0 ? threads : 1)
#endif
  for (int p = 0; p < P; ++p) {
    const int64_t lo = (int64_t)(((__int128)p * rows) / P), hi = (int64_t)(((__int128)(p + 1) * rows) / P);
    for (int64_t r = lo; r < hi; ++r) synth_row_f32(k0, k1, scale, (uint64_t)(row0 + r), d, X + r * (int64_t)d);
  }
}

void oracle_synth_wtrue(uint64_t seed, int32_t d, double *w) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const double inv = sqrt((double)d);
  for (int32_t j = 0; j < d; ++j) {
    uint32_t c[4] = {(uint32_t)j, 0u, 0u, 2u}, o[4];
    philox4x32_10(k0, k1, c, o);
    w[j] = ((double)irwin_hall4(o[0], o[1]) * kScale64) / inv;
  }
}

void oracle_synth_labels(uint64_t seed, int kind, int64_t row0, int64_t rows, int32_t d, const float *X,
                         const double *w_true, double *labels) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int64_t r = 0; r < rows; ++r) {
    const uint64_t i = (uint64_t)(row0 + r);
    const float *x = X + r * (int64_t)d;
    double m = 0.0;
    for (int32_t j = 0; j < d; ++j) m += (double)x[j] * w_true[j];
    uint32_t cu[4] = {(uint32_t)i, (uint32_t)(i >> 32), 0u, 3u}, ou[4];
    philox4x32_10(k0, k1, cu, ou);
    const double u = ((double)(ou[0] >> 5) * 67108864.0 + (double)(ou[1] >> 6) + 0.5) * 0x1.0p-53;
    if (kind == ORACLE_GRAD_LOGISTIC) {
      labels[r] = (m + log(u) - log(1.0 - u) > 0) ? 1.0 : 0.0;
    } else if (kind == ORACLE_GRAD_HINGE) {
      double y = (m > 0) ? 1.0 : 0.0;
      if (u < 0.05) y = 1.0 - y;
      labels[r] = y;
    } else {
      uint32_t ce[4] = {(uint32_t)i, (uint32_t)(i >> 32), 0u, 4u}, oe[4];
      philox4x32_10(k0, k1, ce, oe);
      labels[r] = m + 0.1 * ((double)irwin_hall4(oe[0], oe[1]) * kScale64);
    }
  }
}

/* CSR flavour (spec in spark-agd_b200/csrc/synth.cu): exactly k entries per row, entry t of row i from
 * counter (i_lo, i_hi, t, 5): col = t*(d/k) + r0 % (d/k), value = irwin_hall4(r1, r2) * (float)(sqrt(3)/65536). */
void oracle_synth_csr_f32(uint64_t seed, int64_t row0, int64_t rows, int32_t d, int32_t k, int64_t *rowptr,
                          int32_t *idx, float *val) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint32_t stride = (uint32_t)(d / k);
  const float scale = (float)kScale64;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int64_t r = 0; r < rows; ++r) {
    const uint64_t i = (uint64_t)(row0 + r);
    rowptr[r] = r * (int64_t)k;
    for (int32_t t = 0; t < k; ++t) {
      uint32_t c[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)t, 5u}, o[4];
      philox4x32_10(k0, k1, c, o);
      idx[r * (int64_t)k + t] = (int32_t)((uint32_t)t * stride + o[0] % stride);
      val[r * (int64_t)k + t] = (float)irwin_hall4(o[1], o[2]) * scale;
    }
  }
  rowptr[rows] = rows * (int64_t)k;
}

/* Bernoulli row mask of the mini-batch GD comparator (spec: row_selected() in spark-agd_b200/csrc/k1_device.cuh):
 * keep row `grow` iff the 64-bit Philox4x32-10 draw keyed by `seed`, counter (row_lo, row_hi, 0, 6) is < thresh. */
int oracle_row_selected(uint64_t seed, uint64_t thresh, int64_t grow) {
  if (thresh == 0) return 1;
  uint32_t c[4] = {(uint32_t)grow, (uint32_t)((uint64_t)grow >> 32), 0u, 6u}, o[4];
  philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), c, o);
  return ((((uint64_t)o[0]) << 32) | o[1]) < thresh;
}
