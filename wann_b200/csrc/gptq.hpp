// gptq.hpp -- data-aware rounding of a conv layer's weights to fp16 (host code, calibration time only).
//
// The tensor-core kernels multiply fp16 weights.  Rounding every weight to the NEAREST fp16 value is optimal
// weight by weight, but the layer's output error  sum_k (w_k - q_k) x_k  depends on how the inputs x co-vary,
// and the inputs of the 192->192 conv layers -- the 9 x 192 centred activations under a 3x3 window -- are
// strongly correlated (the second-moment matrix H = E[x x^T] of the policy net has an effective rank of
// ~15 of 1,728).  The rounding here is the sequential one of GPTQ / optimal brain quantisation: weights are
// fixed one input index at a time, and the rounding error made so far is pushed onto the not-yet-rounded
// weights of the same output channel along the regression of x_k on the remaining inputs, so that the
// errors cancel in the directions in which the inputs actually vary.  With
//     H^-1 = U^T U   (U upper triangular)
// the update after fixing column k is  w_j -= (w_k - q_k) / U_kk * U_kj  for j > k.
// Measured on 100,000 held-out positions: the weight rounding error of the three 192->192 layers all but
// disappears (layer-local output error energy 0.03-0.11 of round-to-nearest), without any extra MMA.
//
// Nothing here is on the evaluation path; the oracle restates it in numpy (oracle/oracle.py gptq_round).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

namespace wann_gptq {

inline int worker_threads() {
  const unsigned hc = std::thread::hardware_concurrency();
  return static_cast<int>(std::max(1u, std::min(16u, hc == 0 ? 4u : hc)));
}

// fp16 round-to-nearest-even of a double, saturating to +-65504 (what cvt.rn.satfinite.f16.f32 of the float does)
inline double round_f16(double v) {
  float f = static_cast<float>(v);
  if (!(f == f)) return 0.0;
  const float a = fabsf(f);
  if (a >= 65520.f) return f < 0 ? -65504.0 : 65504.0;
  if (a < 6.103515625e-05f) {                       // subnormal range: multiples of 2^-24
    const float q = nearbyintf(f * 16777216.f);     // default rounding mode = nearest even
    return static_cast<double>(q) / 16777216.0;
  }
  int e;
  frexpf(a, &e);                                    // a = m * 2^e, m in [0.5, 1): ulp = 2^(e - 11)
  const float ulp = ldexpf(1.f, e - 11);
  return static_cast<double>(nearbyintf(f / ulp) * ulp);
}

// In place: lower Cholesky factor of the symmetric positive definite A [n][n] (row major; the upper part is
// ignored and left untouched).  Returns false if a pivot is not positive.
inline bool cholesky_lower(double* A, int n, int threads) {
  for (int j = 0; j < n; ++j) {
    double* rj = A + static_cast<size_t>(j) * n;
    double d = rj[j];
    for (int k = 0; k < j; ++k) d -= rj[k] * rj[k];
    if (!(d > 0.0)) return false;
    const double ljj = sqrt(d);
    rj[j] = ljj;
    const double inv = 1.0 / ljj;
#pragma omp parallel for schedule(static) num_threads(threads) if (n - j > 256)
    for (int i = j + 1; i < n; ++i) {
      double* ri = A + static_cast<size_t>(i) * n;
      double s = ri[j];
      for (int k = 0; k < j; ++k) s -= ri[k] * rj[k];
      ri[j] = s * inv;
    }
  }
  return true;
}

// X = L^-1 for the lower triangular L [n][n]; X is lower triangular, row major, fully written (zeros above).
inline void invert_lower(const double* L, double* X, int n, int threads) {
  // column c of X solves L x = e_c; rows are produced in order, row i of x needs x[c..i-1].  Work on the
  // TRANSPOSE (Xt[c][i] = X[i][c]) so that both operands of the dot product are contiguous.
  std::vector<double> Xt(static_cast<size_t>(n) * n, 0.0);
#pragma omp parallel for schedule(dynamic, 8) num_threads(threads)
  for (int c = 0; c < n; ++c) {
    double* xc = Xt.data() + static_cast<size_t>(c) * n;
    xc[c] = 1.0 / L[static_cast<size_t>(c) * n + c];
    for (int i = c + 1; i < n; ++i) {
      const double* li = L + static_cast<size_t>(i) * n;
      double s = 0.0;
      for (int k = c; k < i; ++k) s -= li[k] * xc[k];
      xc[i] = s / li[i];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < n; ++c) X[static_cast<size_t>(i) * n + c] = c <= i ? Xt[static_cast<size_t>(c) * n + i] : 0.0;
}

// U [n][n] upper triangular with (H + damp * mean(diag H) * I)^-1 = U^T U.  H: symmetric second-moment matrix
// (float, row major).  Returns false if the damped matrix is not positive definite.
inline bool inverse_factor(const float* H, int n, double damp, std::vector<double>& U) {
  const int threads = worker_threads();
  double tr = 0;
  for (int i = 0; i < n; ++i) tr += H[static_cast<size_t>(i) * n + i];
  const double ridge = damp * tr / n;
  // reversed ordering: A = P (H + ridge I) P, A = L L^T, then U = P L^-1 P
  std::vector<double> A(static_cast<size_t>(n) * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      const int ri = n - 1 - i, rj = n - 1 - j;
      A[static_cast<size_t>(i) * n + j] = 0.5 * (static_cast<double>(H[static_cast<size_t>(ri) * n + rj]) +
                                                 static_cast<double>(H[static_cast<size_t>(rj) * n + ri])) + (i == j ? ridge : 0.0);
    }
  if (!cholesky_lower(A.data(), n, threads)) return false;
  std::vector<double> X(static_cast<size_t>(n) * n);
  invert_lower(A.data(), X.data(), n, threads);
  U.assign(static_cast<size_t>(n) * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) U[static_cast<size_t>(i) * n + j] = X[static_cast<size_t>(n - 1 - i) * n + (n - 1 - j)];
  return true;
}

// Sequential rounding of W [rows][n] (row major, in/out: on return every entry is the weight the kernel will
// multiply).  exact[r * n + k] != 0 marks entries the kernel sees to fp32-class precision (hi + lo operand
// pair): they absorb their share of the compensation but contribute no rounding error.
inline void round_rows(const double* U, int n, double* W, int rows, const uint8_t* exact) {
  const int threads = worker_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
  for (int r = 0; r < rows; ++r) {
    double* w = W + static_cast<size_t>(r) * n;
    const uint8_t* ex = exact ? exact + static_cast<size_t>(r) * n : nullptr;
    for (int k = 0; k < n; ++k) {
      const double* uk = U + static_cast<size_t>(k) * n;
      const double q = (ex && ex[k]) ? static_cast<double>(static_cast<float>(w[k])) : round_f16(w[k]);
      const double err = (w[k] - q) / uk[k];
      w[k] = q;
      if (err != 0.0)
        for (int j = k + 1; j < n; ++j) w[j] -= err * uk[j];
    }
  }
}

}  // namespace wann_gptq
