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
#include <fcntl.h>
#include <math.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <string.h>
#include <sys/file.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <memory>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace wann_gptq {

// Host threads for the calibration-time linear algebra: the CPUs this process may run on, at most 16.  The workers
// below never spin (blocking barriers, plain std::thread): torchrun exports OMP_NUM_THREADS=1 and the ranks of a
// job share the box's cores.
inline int worker_threads() {
  int cpus = 0;
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof set, &set) == 0) cpus = CPU_COUNT(&set);
  if (cpus <= 0) cpus = static_cast<int>(std::thread::hardware_concurrency());
  if (cpus <= 0) cpus = 4;
  if (const char* e = getenv("WANN_B200_HOST_THREADS")) return std::max(1, std::min(64, atoi(e)));
  return std::max(1, std::min(16, cpus));       // the heavy sections run one process at a time (MachineLock)
}

// The factorisations stream 25-100 MB of matrices many times: alone they live in the last-level cache (0.5 s per
// layer), but several processes doing it at once evict each other and crawl (measured: 2 concurrent processes
// 8x slower EACH, 8 processes 50x).  The ranks of a multi-GPU job calibrate at the same moment, so the heavy
// sections take a machine-wide advisory file lock and run one after the other, each with all the threads it may
// use (8 ranks: 8 x 0.5 s per layer instead of minutes).  No lock file, no lock: single processes are unaffected.
class MachineLock {
 public:
  MachineLock() {
    fd_ = open("/tmp/wann_b200.calibration.lock", O_CREAT | O_RDWR | O_CLOEXEC, 0666);
    if (fd_ >= 0 && flock(fd_, LOCK_EX) != 0) { close(fd_); fd_ = -1; }
  }
  ~MachineLock() {
    if (fd_ >= 0) { flock(fd_, LOCK_UN); close(fd_); }
  }
  MachineLock(const MachineLock&) = delete;
  MachineLock& operator=(const MachineLock&) = delete;
 private:
  int fd_ = -1;
};

// Blocking barrier (mutex + condition variable: a waiting thread sleeps).
class Barrier {
 public:
  explicit Barrier(int n) : n_(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m_);
    const unsigned gen = gen_;
    if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); }
    else cv_.wait(lk, [&] { return gen != gen_; });
  }
 private:
  std::mutex m_;
  std::condition_variable cv_;
  int n_, count_ = 0;
  unsigned gen_ = 0;
};

// run body(thread index) on `threads` threads (the caller is thread 0)
inline void run_threads(int threads, const std::function<void(int)>& body) {
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(body, t);
  body(0);
  for (auto& th : pool) th.join();
}

// dot product with four independent accumulators (no reassociation needed from the compiler)
inline double dot4(const double* a, const double* b, int n) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  int k = 0;
  for (; k + 4 <= n; k += 4) { s0 += a[k] * b[k]; s1 += a[k + 1] * b[k + 1]; s2 += a[k + 2] * b[k + 2]; s3 += a[k + 3] * b[k + 3]; }
  for (; k < n; ++k) s0 += a[k] * b[k];
  return (s0 + s1) + (s2 + s3);
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
// ignored and left untouched).  Blocked left-looking: per block column J (64 wide) the block rows I >= J are
// first updated with the block columns to the left (parallel over I), the diagonal block is factored by one
// thread, the panel below is solved (parallel over I): three barriers per block column.  Returns false if a pivot
// is not positive.
inline bool cholesky_lower(double* A, int n, int threads) {
  constexpr int nb = 64;
  const int nblk = (n + nb - 1) / nb;
  threads = std::max(1, std::min(threads, nblk));
  Barrier bar(threads);
  bool ok = true;
  run_threads(threads, [&](int t) {
    for (int J = 0; J < nblk; ++J) {
      const int j0 = J * nb, j1 = std::min(n, j0 + nb);
      // (1) A[I,J] -= sum_{K<J} L[I,K] L[J,K]^T : element (i, j) -= dot(L[i][0..j0), L[j][0..j0))
      for (int I = J + t; I < nblk; I += threads) {
        const int i0 = I * nb, i1 = std::min(n, i0 + nb);
        for (int i = i0; i < i1; ++i) {
          double* ri = A + static_cast<size_t>(i) * n;
          const int jend = std::min(j1, i + 1);
          for (int j = j0; j < jend; ++j) ri[j] -= dot4(ri, A + static_cast<size_t>(j) * n, j0);
        }
      }
      bar.wait();
      // (2) the diagonal block
      if (t == 0)
        for (int j = j0; j < j1 && ok; ++j) {
          double* rj = A + static_cast<size_t>(j) * n;
          const double d = rj[j] - dot4(rj + j0, rj + j0, j - j0);
          if (!(d > 0.0)) { ok = false; break; }
          rj[j] = sqrt(d);
          const double inv = 1.0 / rj[j];
          for (int i = j + 1; i < j1; ++i) {
            double* ri = A + static_cast<size_t>(i) * n;
            ri[j] = (ri[j] - dot4(ri + j0, rj + j0, j - j0)) * inv;
          }
        }
      bar.wait();
      if (!ok) return;
      // (3) the panel below: L[i][j] = (A[i][j] - dot(L[i][j0..j), L[j][j0..j))) / L[j][j]
      for (int I = J + 1 + t; I < nblk; I += threads) {
        const int i0 = I * nb, i1 = std::min(n, i0 + nb);
        for (int i = i0; i < i1; ++i) {
          double* ri = A + static_cast<size_t>(i) * n;
          for (int j = j0; j < j1; ++j) {
            const double* rj = A + static_cast<size_t>(j) * n;
            ri[j] = (ri[j] - dot4(ri + j0, rj + j0, j - j0)) / rj[j];
          }
        }
      }
      bar.wait();
    }
  });
  return ok;
}

// X = L^-1 for the lower triangular L [n][n]; X is lower triangular, row major, fully written (zeros above).
inline void invert_lower(const double* L, double* X, int n, int threads) {
  // column c of X solves L x = e_c; rows are produced in order, row i of x needs x[c..i-1].  Work on the
  // TRANSPOSE (Xt[c][i] = X[i][c]) so that both operands of the dot product are contiguous.  Columns are
  // independent; they are dealt out round-robin (column c costs (n - c)^2 / 2).
  std::vector<double> Xt(static_cast<size_t>(n) * n, 0.0);
  threads = std::max(1, std::min(threads, n));
  run_threads(threads, [&](int t) {
    for (int c = t; c < n; c += threads) {
      double* xc = Xt.data() + static_cast<size_t>(c) * n;
      xc[c] = 1.0 / L[static_cast<size_t>(c) * n + c];
      for (int i = c + 1; i < n; ++i) {
        const double* li = L + static_cast<size_t>(i) * n;
        xc[i] = -dot4(li + c, xc + c, i - c) / li[i];
      }
    }
  });
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < n; ++c) X[static_cast<size_t>(i) * n + c] = c <= i ? Xt[static_cast<size_t>(c) * n + i] : 0.0;
}

// U [n][n] upper triangular with (H + damp * mean(diag H) * I)^-1 = U^T U.  H: symmetric second-moment matrix
// (float, row major).  Returns false if the damped matrix is not positive definite.
inline bool inverse_factor(const float* H, int n, double damp, std::vector<double>& U) {
  MachineLock machine_lock;
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
  const bool trace = getenv("WANN_B200_TRACE_CALIBRATION") != nullptr;
  auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
  const double t0 = now();
  if (!cholesky_lower(A.data(), n, threads)) return false;
  const double t1 = now();
  std::vector<double> X(static_cast<size_t>(n) * n);
  invert_lower(A.data(), X.data(), n, threads);
  if (trace) fprintf(stderr, "[wann_b200] calibration: n = %d, %d threads, cholesky %.3f s, inverse %.3f s\n", n, threads, t1 - t0, now() - t1);
  U.assign(static_cast<size_t>(n) * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) U[static_cast<size_t>(i) * n + j] = X[static_cast<size_t>(n - 1 - i) * n + (n - 1 - j)];
  return true;
}

// inverse_factor behind a small process-wide cache keyed by the contents of H: the contexts of one process very
// often hold the same weights (a white and a black net from one checkpoint, a bench's evaluators), and the factor
// is a pure function of H.  Returns null if H is degenerate.
inline std::shared_ptr<const std::vector<double>> cached_inverse_factor(const float* H, int n, double damp) {
  static std::mutex mu;
  static std::deque<std::pair<uint64_t, std::shared_ptr<const std::vector<double>>>> cache;
  uint64_t key = 1469598103934665603ull ^ static_cast<uint64_t>(n);             // FNV-1a over the 32-bit words of H
  const uint32_t* w = reinterpret_cast<const uint32_t*>(H);
  for (size_t i = 0, m = static_cast<size_t>(n) * n; i < m; ++i) key = (key ^ w[i]) * 1099511628211ull;
  key ^= static_cast<uint64_t>(damp * 1e9);
  {
    std::lock_guard<std::mutex> g(mu);
    for (auto& e : cache)
      if (e.first == key) return e.second;
  }
  auto U = std::make_shared<std::vector<double>>();
  if (!inverse_factor(H, n, damp, *U)) return nullptr;
  std::lock_guard<std::mutex> g(mu);
  cache.emplace_back(key, U);
  while (cache.size() > 6) cache.pop_front();                                 // <= 6 x 24 MB
  return U;
}

// Sequential rounding of W [rows][n] (row major, in/out: on return every entry is the weight the kernel will
// multiply).  exact[r * n + k] != 0 marks entries the kernel sees to fp32-class precision (hi + lo operand
// pair): they absorb their share of the compensation but contribute no rounding error.
inline void round_rows(const double* U, int n, double* W, int rows, const uint8_t* exact) {
  MachineLock machine_lock;
  const int threads = std::max(1, std::min(worker_threads(), rows));
  run_threads(threads, [&](int t) {
    for (int r = t; r < rows; r += threads) {
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
  });
}

}  // namespace wann_gptq
