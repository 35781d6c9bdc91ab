/* Plain-C, loop-level restatement of the layer arithmetic of the reference's training step, written from the
 * TensorFlow / Keras *definitions* of the ops train.py calls -- not from any library's convolution:
 *
 *   Conv2D(filters, 4, 2, 'same') + bias + relu            train.py:158-169  (DownShuffle)
 *   Conv2D(filters, 3, 1, 'same') + bias + relu            train.py:131-139  (Block, block_depth > 0; same loop, k = 3, s = 1)
 *   Conv2DTranspose(filters, 4, 2, 'same') + bias + relu   train.py:145-156  (UpShuffle)
 *   Dense(3) on the last axis                              train.py:198-202
 *   mean squared difference                                train.py:272
 *   noising x*sqrt(abar) + eps*sqrt(1-abar), alpha_dash    train.py:85-93,231-234
 *   tf.keras.optimizers.Adam single step                   train.py:75
 *
 * TEST INFRASTRUCTURE ONLY (like oracle.py): it exists to cross-check oracle.py's PyTorch-CPU (oneDNN) restatement
 * with an implementation that shares no code with it.  PARITY UNPINNED against TensorFlow itself, for the reasons given
 * at the top of oracle.py.  Built by oracle/Makefile into oracle/_direct.so (git-ignored), loaded through ctypes by
 * tests/test_oracle_direct.py.
 *
 * SAME padding (TensorFlow's rule, computed here, not hard-coded): out = ceil(in / stride);
 * pad_total = max((out - 1) * stride + k - in, 0); pad_before = pad_total / 2 (integer division), the rest after.
 * Conv2DTranspose is defined by TensorFlow as the gradient of Conv2D with respect to its input: every input pixel
 * scatters  x[iy, ix, ci] * w[ky, kx, co, ci]  to  out[iy * stride - pad_before + ky, ...]  where pad_before is the
 * SAME padding of the forward convolution that maps the (2x larger) output back onto the input.  */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

static int same_out(int in, int stride) { return (in + stride - 1) / stride; }
static int same_pad_before(int in, int k, int stride) {
  int out = same_out(in, stride);
  int total = (out - 1) * stride + k - in;
  if (total < 0) total = 0;
  return total / 2;
}

/* x [B,H,W,Cin], w [k,k,Cin,Cout] (HWIO), bias [Cout] -> y [B,ceil(H/s),ceil(W/s),Cout]; relu != 0 applies max(.,0) */
void gct2_direct_conv2d_same(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Cin,
                             int Cout, int k, int s, int relu) {
  const int Ho = same_out(H, s), Wo = same_out(W, s);
  const int py = same_pad_before(H, k, s), px = same_pad_before(W, k, s);
  for (int b = 0; b < B; ++b)
    for (int oy = 0; oy < Ho; ++oy)
      for (int ox = 0; ox < Wo; ++ox)
        for (int co = 0; co < Cout; ++co) {
          double acc = bias ? bias[co] : 0.0;
          for (int ky = 0; ky < k; ++ky) {
            const int iy = oy * s - py + ky;
            if (iy < 0 || iy >= H) continue;
            for (int kx = 0; kx < k; ++kx) {
              const int ix = ox * s - px + kx;
              if (ix < 0 || ix >= W) continue;
              const float* xp = x + (((size_t)b * H + iy) * W + ix) * Cin;
              const float* wp = w + (((size_t)ky * k + kx) * Cin) * Cout + co;
              for (int ci = 0; ci < Cin; ++ci) acc += (double)xp[ci] * wp[(size_t)ci * Cout];
            }
          }
          float v = (float)acc;
          if (relu && v < 0.f) v = 0.f;
          y[(((size_t)b * Ho + oy) * Wo + ox) * Cout + co] = v;
        }
}

/* x [B,H,W,Cin], w [k,k,Cout,Cin] (Keras Conv2DTranspose layout), bias [Cout] -> y [B,H*s,W*s,Cout] */
void gct2_direct_conv2d_transpose_same(const float* x, const float* w, const float* bias, float* y, int B, int H, int W,
                                       int Cin, int Cout, int k, int s, int relu) {
  const int Ho = H * s, Wo = W * s;
  /* padding of the forward conv (Ho -> H) whose input-gradient this op is */
  const int py = same_pad_before(Ho, k, s), px = same_pad_before(Wo, k, s);
  const size_t n = (size_t)B * Ho * Wo * Cout;
  /* accumulate in double, then add bias / relu */
  static double* acc = 0;
  static size_t cap = 0;
  if (n > cap) {
    acc = (double*)realloc(acc, n * sizeof(double));
    cap = n;
  }
  memset(acc, 0, n * sizeof(double));
  for (int b = 0; b < B; ++b)
    for (int iy = 0; iy < H; ++iy)
      for (int ix = 0; ix < W; ++ix) {
        const float* xp = x + (((size_t)b * H + iy) * W + ix) * Cin;
        for (int ky = 0; ky < k; ++ky) {
          const int oy = iy * s - py + ky;
          if (oy < 0 || oy >= Ho) continue;
          for (int kx = 0; kx < k; ++kx) {
            const int ox = ix * s - px + kx;
            if (ox < 0 || ox >= Wo) continue;
            double* op = acc + (((size_t)b * Ho + oy) * Wo + ox) * Cout;
            const float* wp = w + (((size_t)ky * k + kx) * Cout) * Cin;
            for (int co = 0; co < Cout; ++co) {
              double t = 0.0;
              for (int ci = 0; ci < Cin; ++ci) t += (double)xp[ci] * wp[(size_t)co * Cin + ci];
              op[co] += t;
            }
          }
        }
      }
  for (size_t i = 0; i < n; ++i) {
    float v = (float)(acc[i] + (bias ? bias[i % Cout] : 0.0));
    if (relu && v < 0.f) v = 0.f;
    y[i] = v;
  }
}

/* Keras Dense on a rank-4 input: contracts the last axis only.  x [rows,Cin], w [Cin,Cout], bias [Cout] */
void gct2_direct_dense(const float* x, const float* w, const float* bias, float* y, long long rows, int Cin, int Cout) {
  for (long long r = 0; r < rows; ++r)
    for (int co = 0; co < Cout; ++co) {
      double acc = bias[co];
      for (int ci = 0; ci < Cin; ++ci) acc += (double)x[r * Cin + ci] * w[(size_t)ci * Cout + co];
      y[r * Cout + co] = (float)acc;
    }
}

/* tf.reduce_mean(tf.math.squared_difference(a, b)) */
double gct2_direct_mse(const float* a, const float* b, long long n) {
  double s = 0.0;
  for (long long i = 0; i < n; ++i) {
    const double d = (double)a[i] - (double)b[i];
    s += d * d;
  }
  return s / (double)n;
}

/* alpha_dash (train.py:85-93) and the noising of Trainer.call (train.py:231-234), per image */
double gct2_direct_alpha_dash(double t, int steps) {
  const double u = 1.0 - t / (double)(steps + 1);
  return u * u * 0.25;
}
void gct2_direct_noise(const float* x, const float* eps, const int* t_int, float* out, int B, long long per_image,
                       int steps) {
  for (int b = 0; b < B; ++b) {
    const double a = gct2_direct_alpha_dash((double)t_int[b], steps);
    const float sa = (float)sqrt(a), sb = (float)sqrt(1.0 - a);
    for (long long i = 0; i < per_image; ++i)
      out[b * per_image + i] = x[b * per_image + i] * sa + eps[b * per_image + i] * sb;
  }
}

/* One tf.keras.optimizers.Adam step (OptimizerV2 / ResourceApplyAdam, no amsgrad): t = 1-based iteration.
 *   alpha = lr * sqrt(1 - b2^t) / (1 - b1^t);  m += (g - m)(1 - b1);  v += (g*g - v)(1 - b2);  w -= alpha m / (sqrt(v) + eps) */
void gct2_direct_adam(float* w, float* m, float* v, const float* g, long long n, double lr, double b1, double b2,
                      double eps, long long t) {
  /* ResourceApplyAdam works in the variable's dtype: (1 - beta) is formed in fp32 (1 - 0.999f = 0.0010000467f) */
  const float c1 = 1.0f - (float)b1, c2 = 1.0f - (float)b2;
  const double alpha = lr * sqrt(1.0 - pow(b2, (double)t)) / (1.0 - pow(b1, (double)t));
  for (long long i = 0; i < n; ++i) {
    m[i] = m[i] + (g[i] - m[i]) * c1;
    v[i] = v[i] + (g[i] * g[i] - v[i]) * c2;
    w[i] = (float)(w[i] - alpha * m[i] / (sqrt((double)v[i]) + eps));
  }
}
