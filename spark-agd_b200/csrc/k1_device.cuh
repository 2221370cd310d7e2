// k1_device.cuh -- device helpers shared by the K1 gradient kernels (dense ring, generic, CSR).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/agd_b200.h"

namespace agd {

// MLUtils.log1pExp [mllib-1.3.0]
__device__ __forceinline__ double log1p_exp(double x) { return x > 0 ? x + log1p(exp(-x)) : log1p(exp(x)); }

// loss'(margin) and loss for one example; `m` = x.w.  Formulas: Gradient.scala of spark-mllib 1.3.0.
__device__ __forceinline__ void loss_eval(int kind, double m, double y, double &mult, double &loss) {
  if (kind == AGD_GRAD_LOGISTIC) {
    const double margin = -1.0 * m;
    mult = (1.0 / (1.0 + exp(margin))) - y;
    const double l = log1p_exp(margin);
    loss = (y > 0) ? l : l - margin;
  } else if (kind == AGD_GRAD_LEAST_SQUARES) {
    const double diff = m - y;
    mult = 2.0 * diff;
    loss = diff * diff;
  } else if (kind == AGD_GRAD_LEAST_SQUARES_HALF) {
    const double diff = m - y;
    mult = diff;
    loss = diff * diff / 2.0;
  } else {  // hinge
    const double s = 2 * y - 1.0;
    if (1.0 > s * m) {
      mult = -s;
      loss = 1.0 - s * m;
    } else {
      mult = 0.0;
      loss = 0.0;
    }
  }
}

// Transpose-reduce R per-lane partials across a warp: afterwards every lane holds the warp total of
// row (lane / (32/R)).  R/2 + R/4 + ... + 1 + log2(32/R) 64-bit shuffles instead of 5R.
template <int R>
__device__ __forceinline__ double warp_rows_reduce(double (&p)[R], int lane) {
  int bit = 16;
#pragma unroll
  for (int width = R / 2; width >= 1; width >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < width; ++i) {
      const double send = up ? p[i] : p[i + width];
      const double keep = up ? p[i + width] : p[i];
      p[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  double tot = p[0];
  for (; bit >= 1; bit >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, bit);
  return tot;
}


}  // namespace agd
