// Deferred BatchNorm finalisation ("fold"): instead of a tiny kernel per BatchNorm layer that turns the float64
// sum / sum-of-squares of layer i into (scale, shift), the kernel that CONSUMES layer i's raw output does it in its own
// prologue - every CTA folds the (<= 1280) channels it needs, CTA 0 also writes scale / shift / saved mean / invstd
// and updates the running statistics.  Same arithmetic as bn_finalize_kernel (linear_simt.cu), bit for bit.
#pragma once
#include "common.cuh"

struct BnFoldDev {
  const double* stats; double count;
  const float* gamma; const float* beta;
  float eps, momentum;
  float* running_mean; float* running_var;
  float* scale_out; float* shift_out; float* mean_out; float* invstd_out;
  int C; int active;
};

static inline BnFoldDev p2c_bn_fold_dev(const p2c_bn_fold* f) {
  BnFoldDev d{};
  if (!f) return d;
  d.stats = f->stats; d.count = (double)f->count; d.gamma = f->gamma; d.beta = f->beta; d.eps = f->eps;
  d.momentum = f->momentum; d.running_mean = f->running_mean; d.running_var = f->running_var;
  d.scale_out = f->scale_out; d.shift_out = f->shift_out; d.mean_out = f->mean_out; d.invstd_out = f->invstd_out;
  d.C = f->C; d.active = 1;
  return d;
}

// 0 when the descriptor is usable for a consumer that reads `C` channels
static inline int p2c_bn_fold_check(const p2c_bn_fold* f, int C) {
  if (!f) return 0;
  if (f->C != C || C <= 0 || !f->scale_out || !f->shift_out) return P2C_EINVAL;
  if (f->stats ? f->count <= 0 : (!f->running_mean || !f->running_var)) return P2C_EINVAL;
  return 0;
}

// (sum, sumsq) = the float64 sum / sum of squares of channel c over f.count rows (ignored when f.stats is NULL)
__device__ __forceinline__ void p2c_bn_fold_sums(const BnFoldDev& f, int c, bool writer, double sum, double sumsq,
                                                 float& sc, float& sh) {
  double mean, var;
  if (f.stats) {
    mean = sum / f.count;
    var = sumsq / f.count - mean * mean;
    if (var < 0.0) var = 0.0;
    if (writer && f.running_mean) {
      const double unbiased = f.count > 1.0 ? var * f.count / (f.count - 1.0) : var;
      f.running_mean[c] = (float)((1.0 - (double)f.momentum) * (double)f.running_mean[c] + (double)f.momentum * mean);
      f.running_var[c] = (float)((1.0 - (double)f.momentum) * (double)f.running_var[c] + (double)f.momentum * unbiased);
    }
  } else {
    mean = (double)f.running_mean[c];
    var = (double)f.running_var[c];
  }
  const double invstd = 1.0 / sqrt(var + (double)f.eps);
  const double g = f.gamma ? (double)f.gamma[c] : 1.0;
  const double bta = f.beta ? (double)f.beta[c] : 0.0;
  sc = (float)(g * invstd);
  sh = (float)(bta - mean * g * invstd);
  if (writer) {
    f.scale_out[c] = sc;
    f.shift_out[c] = sh;
    if (f.mean_out) f.mean_out[c] = (float)mean;
    if (f.invstd_out) f.invstd_out[c] = (float)invstd;
  }
}

__device__ __forceinline__ void p2c_bn_fold_channel(const BnFoldDev& f, int c, bool writer, float& sc, float& sh) {
  p2c_bn_fold_sums(f, c, writer, f.stats ? f.stats[c] : 0.0, f.stats ? f.stats[f.C + c] : 0.0, sc, sh);
}

// BatchNorm sums of the xyz-only first SA conv y_c = w_c . d + b_c from the nine moments m of d over n rows
// (sa_first.cu: m[0..2] = sum d, m[3..8] = sum d d^T as xx, xy, xz, yy, yz, zz)
__device__ __forceinline__ void p2c_xyz_first_sums(const double* __restrict__ m, double n, double w0, double w1, double w2,
                                                   double b, double& sum, double& sumsq) {
  const double ws = w0 * m[0] + w1 * m[1] + w2 * m[2];
  const double q = w0 * w0 * m[3] + w1 * w1 * m[6] + w2 * w2 * m[8] + 2.0 * (w0 * w1 * m[4] + w0 * w2 * m[5] + w1 * w2 * m[7]);
  sum = n * b + ws;
  sumsq = q + 2.0 * b * ws + n * b * b;
}
