// C-ABI + host runtime of the B200-native WaNN policy-net leaf evaluator (include/wann_b200.h).
// Host logic only: weight upload / tensor-core weight image construction, lanes (streams +
// workspaces) for re-entrant concurrent callers, chunking, copy/compute overlap, error
// reporting.  All arithmetic is in the kernels (conv_stack_tc.cuh, aux_kernels.cuh).
// There is NO CPU fallback: every entry point either runs the CUDA kernels or fails.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <omp.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/wann_b200.h"
#include "aux_kernels.cuh"
#include "conv_stack_tc.cuh"
#include "conv_stack_split.cuh"
#include "array_tree.hpp"
#include "gptq.hpp"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(expr)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(WANN_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),    \
                  __FILE__, __LINE__);                                                    \
  } while (0)

// Switches to the context's device for the duration of an API call and restores the caller's
// current device afterwards (the host application -- e.g. torch -- keeps its own notion of it).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t status = cudaSuccess;
  explicit DeviceGuard(int dev) {
    status = cudaGetDevice(&prev);
    if (status == cudaSuccess && prev != dev) {
      status = cudaSetDevice(dev);
      switched = status == cudaSuccess;
    }
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

uint16_t f32_to_bf16(float f) {   // round-to-nearest-even, like cvt.rn.bf16.f32
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x7FFFFFu)) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

uint16_t f32_to_f16(float f) {    // round-to-nearest-even, saturating to the largest finite fp16
  if (f > 65504.f) f = 65504.f;
  if (f < -65504.f) f = -65504.f;
  const __half h = __float2half_rn(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  bool busy = false;
  int8_t* d_squares = nullptr;
  uint8_t* d_black = nullptr;
  void* d_planes = nullptr;
  float* d_h5 = nullptr;
  float* d_probs = nullptr;
  float* d_logits = nullptr;
  uint8_t* d_idx = nullptr;
  float* d_prior = nullptr;
  uint8_t* d_nlegal = nullptr;
  int8_t* d_planes_out = nullptr;
  uint8_t* d_mask = nullptr;
  int8_t* d_children = nullptr;   // [cap][48][64], allocated on first use
  uint8_t* d_cflags = nullptr;    // [cap][48]
  int32_t* d_cstats = nullptr;    // [cap][48][4] child seeds (lazy)
  int32_t* d_psum = nullptr;      // [cap][8] parent summaries (lazy)
  int* d_err = nullptr;
  int* h_err = nullptr;     // pinned
  float* d_act0 = nullptr;  // fp32 mode ping-pong activations
  float* d_act1 = nullptr;
  float* d_h5r = nullptr;   // margin refinement: compact h5 rows of the flagged boards (lazy)
  int* d_rlist = nullptr;   //                    their board indices (lazy)
  uint32_t refine_seen = 0; // value of the monotonic device counter d_err[5] already added to the stats
  uint8_t* z_in = nullptr;  // zero-copy small host calls: pinned, device-mapped staging (lazy)
  uint8_t* z_out = nullptr;
  long long cap = 0;        // positions
};

struct Lane {
  Slot slot[2];
  bool in_use = false;
  cudaEvent_t last = nullptr;   // completion of the last async (device/stream) call
  bool last_valid = false;
  cudaStream_t last_stream = nullptr;   // stream `last` was recorded on
};

// Readers-writer gate with WRITER preference (glibc's rwlock lets a steady stream of readers starve
// the writer): evaluations enter as readers, wann_set_weights as the writer.
struct WeightsGate {
  std::mutex m;
  std::condition_variable cv;
  int readers = 0, writers_waiting = 0;
  bool writer = false;
  void reader_enter() {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return !writer && writers_waiting == 0; });
    ++readers;
  }
  void reader_exit() {
    std::lock_guard<std::mutex> lk(m);
    if (--readers == 0) cv.notify_all();
  }
  void writer_enter() {
    std::unique_lock<std::mutex> lk(m);
    ++writers_waiting;
    cv.wait(lk, [&] { return !writer && readers == 0; });
    --writers_waiting;
    writer = true;
  }
  void writer_exit() {
    std::lock_guard<std::mutex> lk(m);
    writer = false;
    cv.notify_all();
  }
};
struct ReaderScope {
  WeightsGate* g;
  explicit ReaderScope(WeightsGate* gate) : g(gate) { if (g) g->reader_enter(); }
  ~ReaderScope() { if (g) g->reader_exit(); }
  ReaderScope(const ReaderScope&) = delete;
  ReaderScope& operator=(const ReaderScope&) = delete;
};

constexpr int kLanes = 4;
constexpr long long kDevChunkMul = 4;
constexpr long long kCombineMaxCall = 256;    // host calls up to this many positions may be merged
constexpr long long kCombineMaxBatch = 4096;  // positions per merged launch
constexpr long long kZeroCopyMax = 512;       // host calls up to this many positions skip the DMA copies

// One small host call waiting to be merged with its concurrent siblings (flat combining).
struct CombineReq {
  const int8_t* squares; const uint8_t* black; long long n;
  float* probs; float* logits; uint8_t* idx; float* prior; uint8_t* nlegal;
  int rc = 0; std::string err; bool done = false;
};

// Centred fp16 mode (WANN_MODE_FP16C): per conv layer 1..4 a permutation of the output channels and one offset
// per BUCKET of 8 channels.  Positions are sorted by activation variance (descending) into the 3 groups of 64
// that are the K-blocks of the next layer's weight stream -- the group whose weight rounding error matters
// most comes first, which is what the partial hi/lo split (options "split_layers", "lo_n", "lo_kgroups")
// relies on; inside a group, bucket = channel position % 8 (the halo cells of the shared-memory layout are
// shared by the 8 channel chunks of a group, so an offset can only depend on the group and position % 8),
// filled by ascending mean activation.  Identity / zero = the plain layout.
inline int cen_bucket(int q) { return (q / 64) * 8 + (q % 8); }
struct Centering {
  uint8_t pos_of[4][192];   // TF channel -> position in the kernel's layout
  uint8_t tf_of[4][192];    // position -> TF channel
  float c[4][wann::kCenBuckets];   // offsets (exact fp16 values) of the stored output of conv l+1, by cen_bucket(position)
  Centering() {
    for (int l = 0; l < 4; ++l) {
      for (int i = 0; i < 192; ++i) pos_of[l][i] = tf_of[l][i] = static_cast<uint8_t>(i);
      for (int j = 0; j < wann::kCenBuckets; ++j) c[l][j] = 0.f;
    }
  }
};

}  // namespace

struct wann_ctx {
  int device = 0;
  int mode = WANN_MODE_BF16;
  int sm_count = 0;
  int final_kernel = 3;
  long long chunk = 16384;
  // options may be changed by wann_set_option while other threads evaluate: atomics, read once per call
  std::atomic<int> variant{0};
  std::atomic<int> pdl{-1};                // programmatic dependent launch between consecutive evaluations of a stream:
                                           // -1 auto (on for <= 2 waves of supertiles: N=1 35 -> 31 us, 1,024 95 -> 92 us back to
                                           // back; off beyond: 4,096 positions measured 7 % slower with it), 0 off, 1 on
  std::atomic<bool> small_batch_halves{true};   // batches <= 4 x (SM pairs) run 4-board supertiles
  std::atomic<bool> fuse_tail{true};            // bf16/fp16 modes: tail warps inside the conv-stack kernel (one kernel per evaluation)
  std::atomic<bool> zero_copy{true};            // host calls <= kZeroCopyMax positions: kernels read/write mapped host memory
  std::atomic<int> skew{4};                // conv_stack_tc: stages by which half 1 trails the weight stream (0..6)
  // weights on device
  float* d_conv_w[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float* d_conv_b[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float* d_fc_w = nullptr;
  float* d_fc_b = nullptr;
  uint8_t* d_wimg = nullptr;   // tensor-core weight stage images, [2][kImgBytesPerRank]
  uint8_t* d_wimg_split = nullptr;   // bf16/fp16 modes: the fp16 hi/lo split image of the refinement pass
  // margin refinement: boards whose two best legal priors are closer than `margin` logits are
  // re-evaluated by conv_stack_split_kernel (fp32-class); refine_ratio = exp(-margin), 0 = off
  std::atomic<float> refine_ratio{0.f};
  float* d_bias_tc = nullptr;  // [4*192 + 4] bias table of the context's own tensor-core kernel (centred mode: b' - c)
  float* d_bias_plain = nullptr;   // [4*192 + 4] the TF biases (fp32_tc mode and the refinement pass)
  // centred fp16 mode: calibrated channel buckets / offsets, conv5 tap corrections, weight stream geometry
  Centering cen;
  float kappa[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  uint8_t* d_unperm = nullptr;     // [4][192] channel position -> TF channel
  std::atomic<int> split_mask{0};  // option "split_layers": conv layers (bits 1..3) whose weights stream as hi + lo
  std::atomic<int> lo_n{64};       // option "lo_n": the lo stages cover output positions [0, lo_n) ...
  std::atomic<int> lo_kgroups{1};  // option "lo_kgroups": ... and the first lo_kgroups 64-channel input groups of every tap
  int img_lo_n = 192, img_lo_kgroups = 3, img_split_mask = 0;   // what the uploaded stage image was built with
  // data-aware weight rounding (gptq.hpp; option "gptq", fp16c mode, default on): per 192->192 layer the
  // factor U of the inverse second-moment matrix of its calibration inputs (kept: re-rounding after a change
  // of the split options needs no new calibration), and the weights the kernel actually multiplies
  std::atomic<bool> gptq{true};
  std::shared_ptr<const std::vector<double>> gptq_u[3];   // (shared with the process-wide cache, wann_gptq::cached_inverse_factor)
  std::vector<float> w_eff[3];     // [9][192][192] HWIO: effective weights of hidden_layer/2..4 (wann_debug_effective_weights)
  int stages_per_st = wann::kKbPerSupertile;
  long long img_stride = wann::kImgBytesPerRank;
  size_t wimg_capacity = 0;
  std::vector<float> host_w[12];   // copy of the weights (images are rebuilt when split_layers changes)
  WeightsGate weights_gate;       // evaluations are readers, wann_set_weights the writer
  Lane lanes[kLanes];
  std::mutex mu;
  std::condition_variable cv;
  std::atomic<uint64_t> stats[8];
  // flat-combining coalescer for concurrent small HOST calls (the reference calls evaluate with
  // batch 1 from 11 threads, tree_builder.pyx:186-189): whoever finds the combiner idle becomes
  // the leader, merges every queued request into ONE launch and hands the results back.
  std::atomic<bool> coalesce{false};
  std::mutex cmu;
  std::condition_variable ccv;
  std::deque<CombineReq*> cq;
  bool cbusy = false;
  int8_t* c_sq = nullptr; uint8_t* c_bl = nullptr; float* c_probs = nullptr; float* c_logits = nullptr;
  uint8_t* c_idx = nullptr; float* c_prior = nullptr; uint8_t* c_nlegal = nullptr;   // pinned staging
  // optional CUDA-event timing of the conv-stack kernel (bench.py roofline)
  std::atomic<bool> time_kernels{false};
  std::mutex ev_mu;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pairs;
};

namespace {

using namespace wann;

float f16_value(float v) {
  const uint16_t h = f32_to_f16(v);
  __half hh;
  memcpy(&hh, &h, 2);
  return __half2float(hh);
}

struct ImageOpts {
  bool f16 = false;
  bool exact15 = false;     // conv1: lo parts (x 4096) in input channels 4..7; conv5: lo parts (x 4096) in rows 16..24
  int split_mask = 0;       // bit l (1..3): conv(l+1) K-blocks (tap, j < lo_kgroups) stream a lo stage (w - fp16(w),
                            // unscaled, output positions [0, lo_n)) after the hi stage
  int lo_n = 192;
  int lo_kgroups = 3;
  const Centering* cen = nullptr;
  // conv_stack_tc_kernel only: every layer conv1..4 starts with a BIAS stage [96 rows x 16 k]: k = 0..2 hold the
  // bias of the output position as three 16-bit parts (hi + mid + lo); the kernel multiplies it with a tile of
  // ones as the layer's first MMA (accumulate = 0), so the epilogue has no bias to load or add.
  // [4][192] by output position (centred mode: b' - c), or null = no bias stages
  const float* bias = nullptr;
};
constexpr int kBiasStageBytes = 96 * 16 * 2;

int split_layer_count(int split_mask) {
  int n = 0;
  for (int l = 1; l <= 3; ++l) n += (split_mask >> l) & 1;
  return n;
}
int stages_per_supertile(int split_mask, int lo_n, int lo_kgroups, bool bias_stages) {
  const int bps = static_cast<int>(lo_blocks_per_stage(lo_n));        // (tap, j) lo blocks per stage
  return kKbPerSupertile + split_layer_count(split_mask) * ((9 * lo_kgroups + bps - 1) / bps) + (bias_stages ? 4 : 0);
}
size_t image_bytes_per_rank(int split_mask, int lo_n, int lo_kgroups, bool bias_stages) {
  return static_cast<size_t>(kKbPerSupertile - kKbL5) * kStageBytes + kKbL5 * kStageBytesL5 +
         static_cast<size_t>(split_layer_count(split_mask)) * 9 * lo_kgroups * 64 * lo_n +
         (bias_stages ? 4 * kBiasStageBytes : 0);
}

// Build the shared-memory stage images of operand B (see conv_stack_tc.cuh).
// Element (n, k) of a stage lives at (k/8)*LBO + (n/8)*128 + (n%8)*16 + (k%8)*2 bytes.
void build_weight_image(const wann_weights& w, const ImageOpts& o, std::vector<uint8_t>& img) {
  static const Centering identity;
  const Centering& cen = o.cen ? *o.cen : identity;
  const size_t per_rank = image_bytes_per_rank(o.split_mask, o.lo_n, o.lo_kgroups, o.bias != nullptr);
  img.assign(2 * per_rank, 0);
  for (int rank = 0; rank < 2; ++rank) {
    uint8_t* base = img.data() + static_cast<size_t>(rank) * per_rank;
    size_t off = 0;
    auto put = [&](uint8_t* stage, int lbo, int n, int k, float v) {
      const uint16_t h = o.f16 ? f32_to_f16(v) : f32_to_bf16(v);
      memcpy(stage + (k / 8) * lbo + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2, &h, 2);
    };
    auto bias_stage = [&](int layer) {                // three 16-bit parts of the bias at k = 0, 1, 2
      if (o.bias == nullptr) return;
      for (int nl = 0; nl < 96; ++nl) {
        float rest = o.bias[layer * 192 + rank * 96 + nl];
        for (int k = 0; k < 3; ++k) {
          put(base + off, kLboB, nl, k, rest);
          uint16_t h = o.f16 ? f32_to_f16(rest) : f32_to_bf16(rest);
          float part;
          if (o.f16) part = f16_value(rest);
          else { const uint32_t u = static_cast<uint32_t>(h) << 16; memcpy(&part, &u, 4); }
          rest -= part;
        }
      }
      off += kBiasStageBytes;
    };
    // conv1: k-block kb holds taps 4kb..4kb+3, 16 k per tap (4 real input planes)
    bias_stage(0);
    for (int kb = 0; kb < kKbL1; ++kb, off += kStageBytes)
      for (int ks = 0; ks < 4; ++ks) {
        const int tap = kb * 4 + ks;
        if (tap >= 9) continue;
        for (int c = 0; c < 4; ++c)
          for (int nl = 0; nl < 96; ++nl) {
            const int co = cen.tf_of[0][rank * 96 + nl];
            const float v = w.conv_w[0][(tap * 4 + c) * 192 + co];
            put(base + off, kLboB, nl, ks * 16 + c, v);
            if (o.exact15) put(base + off, kLboB, nl, ks * 16 + 4 + c, (v - f16_value(v)) * 4096.f);
          }
      }
    // conv2..4: k-block kb = tap*3 + j holds input channel positions 64j..64j+63 of tap
    for (int layer = 1; layer <= 3; ++layer) {
      bias_stage(layer);
      const int lo_j = ((o.split_mask >> layer) & 1) ? o.lo_kgroups : 0;
      const int bps = static_cast<int>(lo_blocks_per_stage(o.lo_n));
      int pend_tap[16], pend_j[16], pend = 0;
      auto weight = [&](int tap, int j, int out_pos, int k) {
        const int co = cen.tf_of[layer][out_pos], ci = cen.tf_of[layer - 1][j * 64 + k];
        return w.conv_w[layer][(static_cast<size_t>(tap) * 192 + ci) * 192 + co];
      };
      for (int kb = 0; kb < kKbMid; ++kb) {
        const int tap = kb / 3, j = kb % 3;
        for (int k = 0; k < 64; ++k)
          for (int nl = 0; nl < 96; ++nl) put(base + off, kLboB, nl, k, weight(tap, j, rank * 96 + nl, k));
        off += kStageBytes;
        if (j < lo_j) {
          pend_tap[pend] = tap; pend_j[pend] = j; ++pend;
          if (pend == bps || (tap == 8 && j == lo_j - 1)) {
            // one lo stage = the pending (tap, j) blocks, [lo_n / 2 rows x 64 k] each: an N = lo_n MMA takes
            // columns [0, lo_n / 2) from CTA rank 0 and [lo_n / 2, lo_n) from rank 1
            const int half = o.lo_n / 2;
            for (int b = 0; b < pend; ++b, off += static_cast<size_t>(64) * o.lo_n)
              for (int k = 0; k < 64; ++k)
                for (int nl = 0; nl < half; ++nl) {
                  const float v = weight(pend_tap[b], pend_j[b], rank * half + nl, k);
                  put(base + off, 8 * o.lo_n, nl, k, v - f16_value(v));
                }
            pend = 0;
          }
        }
      }
    }
    // conv5 as a 1x1 GEMM whose N index is the TAP: rows 0..8 (CTA rank 0) = the taps, rows 16..24
    // (CTA rank 1) = their lo parts x 4096 (exact15 only)
    for (int j = 0; j < kKbL5; ++j, off += kStageBytesL5)
      for (int nl = 0; nl < 16; ++nl) {
        const int tap = nl;
        for (int k = 0; k < 64; ++k) {
          const int ci = cen.tf_of[3][j * 64 + k];
          float v = 0.f;
          if (w.final_kernel == 3) {
            if (tap < 9) v = w.conv_w[4][tap * 192 + ci];
          } else if (tap == 4) {
            v = w.conv_w[4][ci];
          }
          if (rank == 1) v = o.exact15 ? (v - f16_value(v)) * 4096.f : 0.f;
          put(base + off, kLboB5, nl, k, v);
        }
      }
  }
}

// WANN_MODE_FP32_TC: every weight as hi = fp16(w), lo = fp16((w - hi) * 4096); per K-block the hi
// stage is followed by the lo stage (conv_stack_split.cuh).
void build_weight_image_split(const wann_weights& w, std::vector<uint8_t>& img) {
  std::vector<uint8_t> hi_img, lo_img;
  ImageOpts fo;
  fo.f16 = true;
  build_weight_image(w, fo, hi_img);
  // residuals: same layout, value (w - float(hi)) * 4096 -- build by transforming the weights
  std::vector<float> res[5];
  const size_t k5 = static_cast<size_t>(w.final_kernel) * w.final_kernel;
  const size_t wsz[5] = {9 * 4 * 192, 9 * 192 * 192, 9 * 192 * 192, 9 * 192 * 192, k5 * 192};
  wann_weights r = w;
  for (int l = 0; l < 5; ++l) {
    res[l].resize(wsz[l]);
    for (size_t i = 0; i < wsz[l]; ++i) {
      const uint16_t h = f32_to_f16(w.conv_w[l][i]);
      __half hh;
      memcpy(&hh, &h, 2);
      res[l][i] = (w.conv_w[l][i] - __half2float(hh)) * 4096.f;
    }
    r.conv_w[l] = res[l].data();
  }
  build_weight_image(r, fo, lo_img);
  img.assign(2 * static_cast<size_t>(kSplitImgBytesPerRank), 0);
  for (int rank = 0; rank < 2; ++rank) {
    const uint8_t* h = hi_img.data() + static_cast<size_t>(rank) * kImgBytesPerRank;
    const uint8_t* l = lo_img.data() + static_cast<size_t>(rank) * kImgBytesPerRank;
    uint8_t* o = img.data() + static_cast<size_t>(rank) * kSplitImgBytesPerRank;
    size_t off = 0;
    for (int kb = 0; kb < kKbPerSupertile; ++kb) {
      const size_t bytes = (kb >= kKbPerSupertile - kKbL5) ? kStageBytesL5 : kStageBytes;
      memcpy(o + 2 * off, h + off, bytes);
      memcpy(o + 2 * off + bytes, l + off, bytes);
      off += bytes;
    }
  }
}

int alloc_slot(Slot& s, long long cap, int mode) {
  s.cap = cap;
  CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
  CU(cudaMalloc(&s.d_squares, cap * 64));
  CU(cudaMalloc(&s.d_black, cap));
  CU(cudaMalloc(&s.d_planes, cap * 256 * 4));
  // device-pointer calls need no staging, only h5: they run in chunks 4x as large (fewer launches,
  // smaller wave-quantisation tail) in the tensor-core modes
  CU(cudaMalloc(&s.d_h5, cap * 64 * 4 * (mode == WANN_MODE_FP32 ? 1 : kDevChunkMul)));
  CU(cudaMalloc(&s.d_probs, cap * kMoves * 4));
  CU(cudaMalloc(&s.d_logits, cap * kMoves * 4));
  CU(cudaMalloc(&s.d_idx, cap * kMaxLegal));
  CU(cudaMalloc(&s.d_prior, cap * kMaxLegal * 4));
  CU(cudaMalloc(&s.d_nlegal, cap));
  CU(cudaMalloc(&s.d_planes_out, cap * 256));
  CU(cudaMalloc(&s.d_mask, cap * kMoves));
  CU(cudaMalloc(&s.d_err, 8 * sizeof(int)));
  CU(cudaMemset(s.d_err, 0, 8 * sizeof(int)));
  CU(cudaMallocHost(&s.h_err, 8 * sizeof(int)));
  memset(s.h_err, 0, 8 * sizeof(int));
  if (mode == WANN_MODE_FP32) {
    CU(cudaMalloc(&s.d_act0, cap * 64 * 192 * 4));
    CU(cudaMalloc(&s.d_act1, cap * 64 * 192 * 4));
  }
  return WANN_OK;
}

void free_slot(Slot& s) {
  if (s.stream) cudaStreamDestroy(s.stream);
  if (s.done) cudaEventDestroy(s.done);
  cudaFree(s.d_squares); cudaFree(s.d_black); cudaFree(s.d_planes); cudaFree(s.d_h5);
  cudaFree(s.d_probs); cudaFree(s.d_logits); cudaFree(s.d_idx); cudaFree(s.d_prior);
  cudaFree(s.d_nlegal); cudaFree(s.d_planes_out); cudaFree(s.d_mask); cudaFree(s.d_err);
  if (s.h_err) cudaFreeHost(s.h_err);
  cudaFree(s.d_act0); cudaFree(s.d_act1); cudaFree(s.d_children); cudaFree(s.d_cflags);
  cudaFree(s.d_h5r); cudaFree(s.d_rlist); cudaFree(s.d_cstats); cudaFree(s.d_psum);
  if (s.z_in) cudaFreeHost(s.z_in);
  if (s.z_out) cudaFreeHost(s.z_out);
  s = Slot();
}

// lanes are allocated lazily (first use) so that idle lanes cost no HBM
int acquire_lane(wann_ctx* ctx, Lane** out) {
  std::unique_lock<std::mutex> lk(ctx->mu);
  while (true) {
    for (int i = 0; i < kLanes; ++i) {
      Lane& l = ctx->lanes[i];
      if (!l.in_use) {
        l.in_use = true;
        lk.unlock();
        if (l.slot[0].stream == nullptr) {
          const long long cap = (ctx->mode == WANN_MODE_FP32) ? std::min<long long>(ctx->chunk, 2048)
                                                               : ctx->chunk;
          for (int k = 0; k < 2; ++k) {
            int rc = alloc_slot(l.slot[k], cap, ctx->mode);
            if (rc != WANN_OK) {
              free_slot(l.slot[0]);   // leave the lane unallocated so that a later call retries cleanly
              free_slot(l.slot[1]);
              std::lock_guard<std::mutex> g(ctx->mu);
              l.in_use = false;
              ctx->cv.notify_one();
              return rc;
            }
          }
          cudaEventCreateWithFlags(&l.last, cudaEventDisableTiming);
        }
        *out = &l;
        return WANN_OK;
      }
    }
    ctx->cv.wait(lk);
  }
}

// A lane's workspaces may still be in use by an earlier asynchronous (device-pointer) call.
// Work enqueued on the SAME stream is ordered behind it already; any other stream is made to
// wait on the recorded event.  No host synchronisation: back-to-back small calls pipeline.
void lane_wait(Lane* lane, cudaStream_t stream) {
  if (lane->last_valid && lane->last_stream != stream) cudaStreamWaitEvent(stream, lane->last, 0);
}

void release_lane(wann_ctx* ctx, Lane* l) {
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    l->in_use = false;
  }
  ctx->cv.notify_one();
}

bool is_tc16(int mode) { return mode == WANN_MODE_BF16 || mode == WANN_MODE_FP16 || mode == WANN_MODE_FP16C; }

wann_weights host_weights_view(const wann_ctx* ctx) {
  wann_weights w = {};
  for (int i = 0; i < 5; ++i) { w.conv_w[i] = ctx->host_w[2 * i].data(); w.conv_b[i] = ctx->host_w[2 * i + 1].data(); }
  w.fc_w = ctx->host_w[10].data();
  w.fc_b = ctx->host_w[11].data();
  w.final_kernel = ctx->final_kernel;
  return w;
}

// Stage image + bias table (+ the centred mode's tap corrections and channel table) of the context's own
// tensor-core kernel, for the given centring.
int upload_tc_image(wann_ctx* ctx, const wann_weights& w, const Centering& cen) {
  const bool centred = ctx->mode == WANN_MODE_FP16C;
  const int split_mask = centred ? ctx->split_mask.load() : 0;
  const int lo_n = ctx->lo_n.load(), lo_kgroups = ctx->lo_kgroups.load();
  std::vector<uint8_t> img;
  // the weights the image is built from: the true ones, or -- 192->192 layers of the centred mode -- the ones
  // rounded with error compensation (entries covered by a lo stage stay unrounded: the image splits them hi + lo)
  wann_weights wi = w;
  for (int l = 0; l < 3; ++l) {
    std::vector<float>& we = ctx->w_eff[l];
    const float* src = w.conv_w[l + 1];
    we.assign(src, src + static_cast<size_t>(9) * 192 * 192);
    const bool lo_layer = centred && ((split_mask >> (l + 1)) & 1);
    auto covered = [&](int ci, int co) {
      return lo_layer && cen.pos_of[l][ci] < 64 * lo_kgroups && cen.pos_of[l + 1][co] < lo_n;
    };
    if (centred && ctx->gptq.load() && ctx->gptq_u[l] != nullptr) {
      constexpr int K = 9 * 192;
      std::vector<double> W(static_cast<size_t>(192) * K);
      std::vector<uint8_t> exact(static_cast<size_t>(192) * K, 0);
      for (int k = 0; k < K; ++k)
        for (int co = 0; co < 192; ++co) {
          W[static_cast<size_t>(co) * K + k] = src[static_cast<size_t>(k) * 192 + co];
          exact[static_cast<size_t>(co) * K + k] = covered(k % 192, co);
        }
      wann_gptq::round_rows(ctx->gptq_u[l]->data(), K, W.data(), 192, exact.data());
      for (int k = 0; k < K; ++k)
        for (int co = 0; co < 192; ++co) we[static_cast<size_t>(k) * 192 + co] = static_cast<float>(W[static_cast<size_t>(co) * K + k]);
      wi.conv_w[l + 1] = we.data();
    } else if (ctx->mode != WANN_MODE_FP32_TC) {
      // round to nearest: record what the kernel multiplies (16-bit value, or hi + lo where a lo stage covers it)
      for (int k = 0; k < 9 * 192; ++k)
        for (int co = 0; co < 192; ++co) {
          float& v = we[static_cast<size_t>(k) * 192 + co];
          if (ctx->mode == WANN_MODE_BF16) {
            const uint16_t h = f32_to_bf16(v);
            const uint32_t u = static_cast<uint32_t>(h) << 16;
            memcpy(&v, &u, 4);
          } else {
            const float hi = f16_value(v);
            v = covered(k % 192, co) ? hi + f16_value(v - hi) : hi;
          }
        }
    }
  }
  // bias table: conv1..4 by channel POSITION, then b5.  Centred: b'[n] - c, where
  // b'[n] = b[n] + sum_{tap,k} W[tap][k][n] * c_prev[k] (true fp32 weights, summed in double) gives back
  // what subtracting c_prev from the layer's input took away (halo cells hold -c_prev, so the sum runs
  // over all 9 taps for every pixel).
  std::vector<float> bias(4 * 192 + 4, 0.f);
  for (int l = 0; l < 4; ++l)
    for (int q = 0; q < 192; ++q) {
      const int n = cen.tf_of[l][q];
      double b = w.conv_b[l][n];
      if (centred && l >= 1)
        for (int tap = 0; tap < 9; ++tap)
          for (int k = 0; k < 192; ++k)
            b += static_cast<double>(w.conv_w[l][(static_cast<size_t>(tap) * 192 + k) * 192 + n]) *
                 cen.c[l - 1][cen_bucket(cen.pos_of[l - 1][k])];
      bias[l * 192 + q] = static_cast<float>(b - (centred ? cen.c[l][cen_bucket(q)] : 0.f));
    }
  bias[4 * 192] = w.conv_b[4][0];
  if (ctx->mode == WANN_MODE_FP32_TC) {
    build_weight_image_split(w, img);
  } else {
    ImageOpts o;
    o.bias = bias.data();
    o.f16 = ctx->mode != WANN_MODE_BF16;
    o.exact15 = centred;
    o.split_mask = split_mask;
    o.lo_n = lo_n;
    o.lo_kgroups = lo_kgroups;
    o.cen = &cen;
    build_weight_image(wi, o, img);
  }
  if (ctx->d_wimg != nullptr && ctx->wimg_capacity < img.size()) { cudaFree(ctx->d_wimg); ctx->d_wimg = nullptr; }
  if (ctx->d_wimg == nullptr) { CU(cudaMalloc(&ctx->d_wimg, img.size())); ctx->wimg_capacity = img.size(); }
  CU(cudaMemcpy(ctx->d_wimg, img.data(), img.size(), cudaMemcpyHostToDevice));
  ctx->stages_per_st = (ctx->mode == WANN_MODE_FP32_TC) ? kKbPerSupertile : stages_per_supertile(split_mask, lo_n, lo_kgroups, true);
  ctx->img_split_mask = split_mask;
  ctx->img_stride = static_cast<long long>(img.size() / 2);
  ctx->img_lo_n = lo_n;
  ctx->img_lo_kgroups = lo_kgroups;
  for (int t = 0; t < 9; ++t) {
    double kp = 0;
    if (centred)
      for (int k = 0; k < 192; ++k) {
        const double c = cen.c[3][cen_bucket(cen.pos_of[3][k])];
        if (w.final_kernel == 3) kp += static_cast<double>(w.conv_w[4][t * 192 + k]) * c;
        else if (t == 4) kp += static_cast<double>(w.conv_w[4][k]) * c;
      }
    ctx->kappa[t] = static_cast<float>(kp);
  }
  if (ctx->d_bias_tc == nullptr) CU(cudaMalloc(&ctx->d_bias_tc, bias.size() * 4));
  CU(cudaMemcpy(ctx->d_bias_tc, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
  if (centred) {
    if (ctx->d_unperm == nullptr) CU(cudaMalloc(&ctx->d_unperm, 4 * 192));
    CU(cudaMemcpy(ctx->d_unperm, cen.tf_of, 4 * 192, cudaMemcpyHostToDevice));
  }
  ctx->cen = cen;
  return WANN_OK;
}

void fill_tc_params(const wann_ctx* ctx, TcParams& p) {
  p.wimg = ctx->d_wimg;
  p.bias = ctx->d_bias_tc;
  p.stages_per_st = ctx->stages_per_st;
  p.img_stride = ctx->img_stride;
  if (ctx->mode == WANN_MODE_FP16C) {
    p.split_mask = ctx->img_split_mask;
    p.lo_n = ctx->img_lo_n;
    p.lo_kgroups = ctx->img_lo_kgroups;
    static_assert(sizeof p.cen == sizeof ctx->cen.c, "centring tables");
    memcpy(p.cen, ctx->cen.c, sizeof p.cen);
    for (int l = 0; l < 4; ++l)
      for (int g = 0; g < kGroups; ++g)
        for (int i = 0; i < 4; ++i) {
          const float c0 = ctx->cen.c[l][8 * g + 2 * i], c1 = ctx->cen.c[l][8 * g + 2 * i + 1];
          const uint32_t lo = c0 != 0.f ? f32_to_f16(-c0) : 0u, hi = c1 != 0.f ? f32_to_f16(-c1) : 0u;   // +0, not -0
          p.cen_neg16[l][g][i] = lo | (hi << 16);
        }
    memcpy(p.kappa, ctx->kappa, sizeof p.kappa);
    p.dbg_unperm = ctx->d_unperm;
  }
}

// Calibration positions of the centred mode: seeded random legal play-outs from the opening, cut at a
// random ply (the rules of board_utils.pyx:331-400; a game that reaches the far rank restarts).
void calibration_positions(int n, std::vector<int8_t>& squares, std::vector<uint8_t>& black) {
  squares.assign(static_cast<size_t>(n) * 64, 0);
  black.assign(n, 0);
  uint64_t state = 0x57414E4E42323030ull;
  auto rnd = [&]() {
    uint64_t z = (state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  };
  for (int g = 0; g < n; ++g) {
    int8_t sq[64];
    int blk = 0;
    auto reset = [&]() { for (int i = 0; i < 64; ++i) sq[i] = i < 16 ? 1 : (i >= 48 ? 2 : 0); blk = 0; };
    reset();
    const int plies = static_cast<int>(rnd() % 60);
    for (int ply = 0; ply < plies; ++ply) {
      int from[48], to[48], cnt = 0;
      const int own = blk ? 2 : 1, dir = blk ? -8 : 8;
      for (int i = 0; i < 64; ++i) {
        if (sq[i] != own) continue;
        const int r = i / 8, f = i % 8, t = i + dir;
        if ((blk && r == 0) || (!blk && r == 7)) continue;
        if (sq[t] == 0) { from[cnt] = i; to[cnt++] = t; }
        if (f > 0 && sq[t - 1] != own) { from[cnt] = i; to[cnt++] = t - 1; }
        if (f < 7 && sq[t + 1] != own) { from[cnt] = i; to[cnt++] = t + 1; }
      }
      if (cnt == 0) break;
      const int m = static_cast<int>(rnd() % static_cast<uint64_t>(cnt));
      const int tr = to[m] / 8;
      if ((blk && tr == 0) || (!blk && tr == 7)) break;      // keep the pre-terminal position
      sq[to[m]] = sq[from[m]];
      sq[from[m]] = 0;
      blk ^= 1;
    }
    memcpy(squares.data() + static_cast<size_t>(g) * 64, sq, 64);
    black[g] = static_cast<uint8_t>(blk);
  }
}

__global__ void channel_partial_sums_kernel(const float* __restrict__ act, long long rows, float* __restrict__ partial) {
  const int c = threadIdx.x;                 // 192 threads: one channel each, coalesced rows
  float acc = 0.f, acc2 = 0.f;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const float a = act[r * 192 + c];
    acc += a;
    acc2 += a * a;
  }
  partial[blockIdx.x * 384 + c] = acc;               // sum
  partial[blockIdx.x * 384 + 192 + c] = acc2;        // sum of squares
}

// Second-moment matrix H = E[x x^T] of the inputs of a 192->192 conv layer (gptq.hpp): x[tap * 192 + c] =
// the CENTRED stored activation a - off[c] of TF channel c at the pixel's 3x3 neighbour `tap`, -off[c] outside
// the board (what the shared-memory halo holds).  One block per 64 x 64 tile of the upper triangle (a tile =
// one tap x one 64-channel group), 256 threads x 4 x 4 outputs, fp32 accumulation over n_pos * 64 pixels.
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ act, const float* __restrict__ off,
                                                   int n_pos, float* __restrict__ H) {
  const int bi = blockIdx.x, bj = blockIdx.y;
  if (bj < bi) return;
  __shared__ __align__(16) float As[32][64];
  __shared__ __align__(16) float Bs[32][64];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int samples = n_pos * 64;
  for (int s0 = 0; s0 < samples; s0 += 32) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int idx = threadIdx.x + 256 * r, sl = idx >> 6, ch = idx & 63;
      const int smp = s0 + sl, pos = smp >> 6, y = (smp >> 3) & 7, x = smp & 7;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const int b = which ? bj : bi, tap = b / 3, c = (b % 3) * 64 + ch;
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        float v = -off[c];
        if (yy >= 0 && yy < 8 && xx >= 0 && xx < 8) v += act[(static_cast<size_t>(pos) * 64 + yy * 8 + xx) * 192 + c];
        (which ? Bs : As)[sl][ch] = v;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int sl = 0; sl < 32; ++sl) {
      const float4 a = *reinterpret_cast<const float4*>(&As[sl][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[sl][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float inv = 1.f / static_cast<float>(samples);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gi = bi * 64 + ty * 4 + i, gj = bj * 64 + tx * 4 + j;
      H[static_cast<size_t>(gi) * 1728 + gj] = acc[i][j] * inv;
      H[static_cast<size_t>(gj) * 1728 + gi] = acc[i][j] * inv;
    }
}

// Centred mode calibration: run the (not yet centred) kernel on the built-in positions, take mean and variance
// of the post-ReLU activation of every channel of conv1..4, sort the channels of a layer by variance into the
// 3 groups of 64 (highest first), the channels of a group by mean into 8 buckets of 8, and give every bucket
// the fp16-rounded mean of its channel means as offset.  Deterministic: fixed positions, fixed summation order.
int calibrate_centering(wann_ctx* ctx, Centering* out) {
  constexpr int kN = 1024, kBlocks = 128;
  std::vector<int8_t> sq;
  std::vector<uint8_t> bl;
  calibration_positions(kN, sq, bl);
  int8_t* d_sq = nullptr; uint8_t* d_bl = nullptr; float* d_act = nullptr; float* d_part = nullptr; int* d_err = nullptr;
  float* d_gram = nullptr; float* d_off = nullptr;
  constexpr int kGramPositions = 512, kGramDim = 9 * 192;
  const bool want_gptq = ctx->gptq.load();
  cudaStream_t stream = nullptr;
  int rc = WANN_OK;
  auto body = [&]() -> int {
    CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (want_gptq) {
      CU(cudaMalloc(&d_gram, static_cast<size_t>(kGramDim) * kGramDim * 4));
      CU(cudaMalloc(&d_off, 192 * 4));
    }
    CU(cudaMalloc(&d_sq, kN * 64));
    CU(cudaMalloc(&d_bl, kN));
    CU(cudaMalloc(&d_act, static_cast<size_t>(kN) * 64 * 192 * 4));
    CU(cudaMalloc(&d_part, kBlocks * 384 * 4));
    CU(cudaMalloc(&d_err, 8 * sizeof(int)));
    CU(cudaMemsetAsync(d_err, 0, 8 * sizeof(int), stream));
    CU(cudaMemcpyAsync(d_sq, sq.data(), kN * 64, cudaMemcpyHostToDevice, stream));
    CU(cudaMemcpyAsync(d_bl, bl.data(), kN, cudaMemcpyHostToDevice, stream));
    Centering cen;
    for (int l = 0; l < 4; ++l) {
      TcParams p = {};
      fill_tc_params(ctx, p);
      p.squares = d_sq; p.black = d_bl; p.n = kN; p.err = d_err; p.halves = 2; p.skew = ctx->skew;
      p.dbg = d_act; p.dbg_layer = l + 1;
      const int npairs = std::min(ctx->sm_count / 2, kN / kBoardsPerPair);
      conv_stack_tc_kernel<true, true><<<2 * npairs, kThreads, kSmemBytes, stream>>>(p);
      channel_partial_sums_kernel<<<kBlocks, 192, 0, stream>>>(d_act, static_cast<long long>(kN) * 64, d_part);
      std::vector<float> part(kBlocks * 384);
      int herr[8];
      CU(cudaMemcpyAsync(part.data(), d_part, part.size() * 4, cudaMemcpyDeviceToHost, stream));
      CU(cudaMemcpyAsync(herr, d_err, sizeof herr, cudaMemcpyDeviceToHost, stream));
      CU(cudaStreamSynchronize(stream));
      CU(cudaGetLastError());
      if (herr[0] != 0) return fail(WANN_E_KERNEL, "calibration: conv_stack_tc watchdog, wait site %d", herr[0]);
      double mean[192], var[192];
      const double rows = static_cast<double>(kN) * 64;
      for (int c = 0; c < 192; ++c) {
        double a = 0, a2 = 0;
        for (int b = 0; b < kBlocks; ++b) { a += part[b * 384 + c]; a2 += part[b * 384 + 192 + c]; }
        mean[c] = a / rows;
        var[c] = a2 / rows - mean[c] * mean[c];
      }
      int order[192];
      for (int c = 0; c < 192; ++c) order[c] = c;
      std::stable_sort(order, order + 192, [&](int a, int b) { return var[a] > var[b]; });
      struct Bucket { double mean; int l, j; };
      std::vector<Bucket> buckets;
      for (int g = 0; g < 3; ++g) {
        int* grp = order + 64 * g;
        std::stable_sort(grp, grp + 64, [&](int a, int b) { return mean[a] < mean[b]; });
        for (int j = 0; j < 8; ++j) {
          int* bk = grp + 8 * j;                               // bucket j of the group: 8 channels of similar mean,
          std::stable_sort(bk, bk + 8, [&](int a, int b) { return var[a] > var[b]; });   // the most variable first
          double m = 0;
          for (int i = 0; i < 8; ++i) {
            const int c = bk[i], q = 64 * g + 8 * i + j;
            m += mean[c];
            cen.pos_of[l][c] = static_cast<uint8_t>(q);
            cen.tf_of[l][q] = static_cast<uint8_t>(c);
          }
          float cj = f16_value(static_cast<float>(m / 8));
          // development knob: keep only the top WANN_B200_OFFSET_BITS mantissa bits of the offset (0 = a power of
          // two): about half of all stored activations are the clamped zeros, i.e. exactly -offset, and an operand
          // with few mantissa bits toggles less of the tensor core's multiplier array
          if (const char* ob = getenv("WANN_B200_OFFSET_BITS")) {
            const int bits = atoi(ob);
            if (bits >= 0 && bits < 10 && cj > 0.f) {
              int e;
              const float fr = frexpf(cj, &e);                       // cj = fr * 2^e, fr in [0.5, 1)
              const float q = static_cast<float>(1 << (bits + 1));
              cj = f16_value(ldexpf(roundf(fr * q) / q, e));
            }
          }
          cen.c[l][8 * g + j] = cj;
          buckets.push_back({m / 8, l, 8 * g + j});
        }
      }
      // The lowest-mean half of the buckets keeps offset 0: their clamped activations stay exact zeros (an
      // operand of 0 toggles nothing in the tensor core's multiplier array -- under the power cap 12 zero buckets
      // of 24 run 3.7 % faster than none, 1,515 -> 1,570 MHz) and, being small, they contribute little to the
      // rounding error (probability median 8.9e-4 -> 9.0e-4; 15: 9.8e-4, 18: 1.05e-3).
      // WANN_B200_ZERO_BUCKETS overrides the count (of 24) for experiments (tools/gpu_partial_split.py).
      int zero_buckets = 12;
      if (const char* zb = getenv("WANN_B200_ZERO_BUCKETS")) zero_buckets = atoi(zb);
      std::stable_sort(buckets.begin(), buckets.end(), [](const Bucket& a, const Bucket& b) { return a.mean < b.mean; });
      for (int i = 0; i < zero_buckets && i < static_cast<int>(buckets.size()); ++i) cen.c[l][buckets[i].j] = 0.f;
      // data-aware weight rounding of the NEXT layer (hidden_layer/(l+2)): second moments of its centred inputs
      // on the device, the factor of their inverse on the host
      if (l < 3) ctx->gptq_u[l].reset();
      if (want_gptq && l < 3) {
        float off_tf[192];
        for (int c = 0; c < 192; ++c) off_tf[c] = cen.c[l][cen_bucket(cen.pos_of[l][c])];
        CU(cudaMemcpyAsync(d_off, off_tf, sizeof off_tf, cudaMemcpyHostToDevice, stream));
        gram_kernel<<<dim3(27, 27), 256, 0, stream>>>(d_act, d_off, kGramPositions, d_gram);
        std::vector<float> H(static_cast<size_t>(kGramDim) * kGramDim);
        CU(cudaMemcpyAsync(H.data(), d_gram, H.size() * 4, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        CU(cudaGetLastError());
        // (a degenerate H -- e.g. a layer whose inputs are all zero -- leaves the factor empty: that layer's
        // weights are then rounded to nearest, which is always correct)
        ctx->gptq_u[l] = wann_gptq::cached_inverse_factor(H.data(), kGramDim, 0.01);   // null if H is degenerate
      }
    }
    *out = cen;
    return WANN_OK;
  };
  rc = body();
  if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
  cudaFree(d_sq); cudaFree(d_bl); cudaFree(d_act); cudaFree(d_part); cudaFree(d_err); cudaFree(d_gram); cudaFree(d_off);
  return rc;
}

// (Re)build every device-side form of the weights: the fp32 TF-layout tensors (fp32 mode, FC of the
// tail), the tensor-core stage images and the bias vector.  Buffers are allocated on first use and
// overwritten afterwards (wann_set_weights); the caller guarantees that no evaluation is in flight.
int upload_weights(wann_ctx* ctx, const wann_weights& w) {
  const size_t k5 = static_cast<size_t>(w.final_kernel) * w.final_kernel;
  const size_t wsz[5] = {9 * 4 * 192, 9 * 192 * 192, 9 * 192 * 192, 9 * 192 * 192, k5 * 192};
  const size_t bsz[5] = {192, 192, 192, 192, 1};
  auto up = [&](float** d, const float* h, size_t nfl) -> int {
    if (*d == nullptr) CU(cudaMalloc(d, nfl * 4));
    CU(cudaMemcpy(*d, h, nfl * 4, cudaMemcpyHostToDevice));
    return WANN_OK;
  };
  for (int i = 0; i < 5; ++i) {
    int rc = up(&ctx->d_conv_w[i], w.conv_w[i], wsz[i]);
    if (rc == WANN_OK) rc = up(&ctx->d_conv_b[i], w.conv_b[i], bsz[i]);
    if (rc != WANN_OK) return rc;
    if (w.conv_w[i] != ctx->host_w[2 * i].data()) {       // (not when re-uploading the context's own copy)
      ctx->host_w[2 * i].assign(w.conv_w[i], w.conv_w[i] + wsz[i]);
      ctx->host_w[2 * i + 1].assign(w.conv_b[i], w.conv_b[i] + bsz[i]);
    }
  }
  int rc = up(&ctx->d_fc_w, w.fc_w, 64 * kMoves);
  if (rc == WANN_OK) rc = up(&ctx->d_fc_b, w.fc_b, kMoves);
  if (rc != WANN_OK) return rc;
  if (w.fc_w != ctx->host_w[10].data()) {
    ctx->host_w[10].assign(w.fc_w, w.fc_w + 64 * kMoves);
    ctx->host_w[11].assign(w.fc_b, w.fc_b + kMoves);
  }
  std::vector<float> bias(4 * 192 + 4, 0.f);
  for (int l = 0; l < 4; ++l) memcpy(bias.data() + l * 192, w.conv_b[l], 192 * 4);
  bias[4 * 192] = w.conv_b[4][0];
  if (ctx->d_bias_plain == nullptr) CU(cudaMalloc(&ctx->d_bias_plain, bias.size() * 4));
  CU(cudaMemcpy(ctx->d_bias_plain, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
  if (ctx->mode != WANN_MODE_FP32) {
    Centering cen;                       // identity: the plain layout
    for (auto& u : ctx->gptq_u) u.reset();   // (factors of an earlier weight set)
    rc = upload_tc_image(ctx, w, cen);
    if (rc == WANN_OK && ctx->mode == WANN_MODE_FP16C) {
      rc = calibrate_centering(ctx, &cen);
      if (rc == WANN_OK) rc = upload_tc_image(ctx, w, cen);
    }
    if (rc != WANN_OK) return rc;
  }
  if (is_tc16(ctx->mode)) {   // the refinement pass's fp16 hi/lo split image
    std::vector<uint8_t> simg;
    build_weight_image_split(w, simg);
    if (ctx->d_wimg_split == nullptr) CU(cudaMalloc(&ctx->d_wimg_split, simg.size()));
    CU(cudaMemcpy(ctx->d_wimg_split, simg.data(), simg.size(), cudaMemcpyHostToDevice));
  }
  return WANN_OK;
}

struct ChunkIO {   // device pointers for one chunk
  const int8_t* squares = nullptr;
  const uint8_t* black = nullptr;
  const void* planes = nullptr;
  int planes_f32 = 0;
  float* probs = nullptr;
  float* logits = nullptr;
  uint8_t* idx = nullptr;
  float* prior = nullptr;
  uint8_t* nlegal = nullptr;
  float* dbg = nullptr;
  int dbg_layer = 0;
  int top_k = kMaxLegal;
  const int* n_dev = nullptr;   // device-side row count (device-resident forest): rows [0, *n_dev - n_dev_off) of this chunk
  long long n_dev_off = 0;
  long long* prof = nullptr;
  int8_t* children = nullptr;
  uint8_t* cflags = nullptr;
  int32_t* cstats = nullptr;    // child seeds [n][48][4] (with cflags and psum)
  int32_t* psum = nullptr;      // parent summaries [n][8]
  int* err_mirror = nullptr;    // host-mapped copy of the slot's d_err, written by the last kernel
  bool allow_defer = false;     // the caller synchronises anyway: it may read the flagged count itself and run
                                // enqueue_refine_pass only when needed (decided by enqueue_eval -> *deferred)
};

bool refine_active(const wann_ctx* ctx, const ChunkIO& io, float refine_ratio) {
  return (io.probs != nullptr || io.logits != nullptr || io.idx != nullptr) && refine_ratio > 0.f &&
         io.squares != nullptr && io.dbg == nullptr && io.prof == nullptr && is_tc16(ctx->mode);
}

// Refinement pass 2: the flagged boards (ws.d_rlist[0 .. d_err[4])) again, fp32-class (fp16 hi/lo
// split on the tensor cores), their rows of every output overwritten.  `count` < 0: the number of
// flagged boards is only known on the device -- grids are sized for "all flagged" and CTAs beyond
// the device-side count return at once; otherwise grids are sized for `count`.
int enqueue_refine_pass(wann_ctx* ctx, Slot& ws, cudaStream_t stream, const ChunkIO& io, long long n,
                        long long count) {
  const long long m = count < 0 ? n : count;
  int* d_rcount = ws.d_err + 4;
  TcParams p = {};
  p.squares = io.squares; p.black = io.black; p.n = n; p.wimg = ctx->d_wimg_split;
  p.bias = ctx->d_bias_plain; p.h5 = ws.d_h5r; p.err = ws.d_err; p.halves = 2;
  p.gather = ws.d_rlist; p.n_dev = d_rcount;
  const long long total_st = (m + kSplitBoardsPerPair - 1) / kSplitBoardsPerPair;
  const int npairs = static_cast<int>(std::min<long long>(ctx->sm_count / 2, total_st));
  conv_stack_split_kernel<<<2 * npairs, kThreads, kSmemBytes, stream>>>(p);
  TailParams t2;
  t2.h5 = ws.d_h5r; t2.squares = io.squares; t2.black = io.black; t2.n = n;
  t2.fc_w = ctx->d_fc_w; t2.fc_b = ctx->d_fc_b; t2.probs = io.probs; t2.logits = io.logits;
  t2.idx = io.idx; t2.prior = io.prior; t2.n_legal = io.nlegal; t2.top_k = io.top_k;
  t2.gather = ws.d_rlist; t2.n_dev = d_rcount;
  if (io.err_mirror != nullptr) { t2.err_src = ws.d_err; t2.err_dst = io.err_mirror; }
  const long long groups = (m + kTailBoards - 1) / kTailBoards;
  const int blocks = static_cast<int>(std::min<long long>((groups + kTailWarps - 1) / kTailWarps, 2ll * ctx->sm_count));
  tail_kernel<<<blocks, kTailWarps * 32, kTailSmemBytes, stream>>>(t2);
  CU(cudaGetLastError());
  ctx->stats[1] += 2;
  return WANN_OK;
}

// Enqueue encode -> conv stack -> tail for `n` positions on `stream`, using `ws` workspaces.
int enqueue_eval(wann_ctx* ctx, Slot& ws, cudaStream_t stream, const ChunkIO& io, long long n,
                 bool* deferred = nullptr) {
  if (deferred != nullptr) *deferred = false;
  if (n <= 0) return WANN_OK;
  uint64_t launches = 0;
  // ---- the tail (FC-155 -> softmax -> legal gather -> rank) and the refinement set-up ---------
  const bool want_tail = io.probs != nullptr || io.logits != nullptr || io.idx != nullptr;
  const bool tc16 = is_tc16(ctx->mode);
  TailParams t;
  t.h5 = ws.d_h5; t.squares = io.squares; t.black = io.black; t.n = n;
  t.fc_w = ctx->d_fc_w; t.fc_b = ctx->d_fc_b; t.probs = io.probs; t.logits = io.logits;
  t.idx = io.idx; t.prior = io.prior; t.n_legal = io.nlegal; t.top_k = io.top_k;
  t.n_dev = io.n_dev; t.n_dev_off = io.n_dev_off;
  const float refine_ratio = ctx->refine_ratio.load();
  const bool refine = refine_active(ctx, io, refine_ratio);
  int* d_rcount = ws.d_err + 4;     // [0] flagged boards of this chunk, [1] running total
  if (refine) {
    if (ws.d_h5r == nullptr) {
      CU(cudaMalloc(&ws.d_h5r, ws.cap * kDevChunkMul * 64 * 4));
      CU(cudaMalloc(&ws.d_rlist, ws.cap * kDevChunkMul * sizeof(int)));
    }
    CU(cudaMemsetAsync(d_rcount, 0, sizeof(int), stream));
    t.refine_list = ws.d_rlist; t.refine_count = d_rcount; t.refine_ratio = refine_ratio;
  }
  // bf16/fp16 modes: the tail runs INSIDE the conv-stack kernel (two tail warps per CTA); the
  // stand-alone tail_kernel serves the other modes, the refinement pass and the debug hooks
  const bool fused = want_tail && tc16 && ctx->fuse_tail && io.dbg == nullptr;
  // deferral needs the flagged count mirrored AFTER the flagging: only the fused kernel's last CTA
  // does that (the stand-alone tail kernel mirrors the record when it starts)
  const bool defer = refine && fused && io.allow_defer && io.err_mirror != nullptr;
  if (deferred != nullptr) *deferred = defer;
  const bool mirror_here = io.err_mirror != nullptr && (!refine || defer);   // else pass 2 mirrors
  if (ctx->mode != WANN_MODE_FP32) {
    TcParams p;
    p.squares = io.squares; p.black = io.black; p.planes = io.planes; p.planes_f32 = io.planes_f32;
    fill_tc_params(ctx, p);
    p.n = n; p.skew = ctx->skew; p.h5 = ws.d_h5;
    p.n_dev = io.n_dev; p.n_dev_off = io.n_dev_off;
    p.dbg = io.dbg; p.dbg_layer = io.dbg_layer; p.variant = ctx->variant; p.err = ws.d_err; p.prof = io.prof;
    if (fused) {
      p.fuse_tail = 1;
      p.tail = t;
      p.h5 = nullptr;                       // the 5th feature map never leaves the SM
      if (mirror_here) p.err_mirror = io.err_mirror;
    }
    // small batches: one half per CTA (4 boards per pair) while that still fits one wave of pairs
    p.halves = (ctx->mode != WANN_MODE_FP32_TC && ctx->small_batch_halves &&
                n <= 4ll * (ctx->sm_count / 2)) ? 1 : 2;
    const long long per_pair = (ctx->mode == WANN_MODE_FP32_TC) ? kSplitBoardsPerPair
                                                               : (p.halves == 1 ? 4 : kBoardsPerPair);
    const long long total_st = (n + per_pair - 1) / per_pair;
    const int npairs = static_cast<int>(std::min<long long>(ctx->sm_count / 2, total_st));
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (ctx->time_kernels) {
      CU(cudaEventCreate(&t0));
      CU(cudaEventCreate(&t1));
      CU(cudaEventRecord(t0, stream));
    }
    {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * npairs);
      cfg.blockDim = dim3(kThreads);
      cfg.dynamicSmemBytes = kSmemBytes;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      const int pdl = ctx->pdl.load();
      // (a kernel that reads its position count from device memory must not start before its predecessor ends)
      attr[0].val.programmaticStreamSerializationAllowed =
          (io.n_dev == nullptr && (pdl > 0 || (pdl < 0 && total_st <= 2ll * (ctx->sm_count / 2)))) ? 1 : 0;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (ctx->mode == WANN_MODE_FP32_TC) CU(cudaLaunchKernelEx(&cfg, conv_stack_split_kernel, p));
      else if (ctx->mode == WANN_MODE_FP16C) CU(cudaLaunchKernelEx(&cfg, conv_stack_tc_kernel<true, true>, p));
      else if (ctx->mode == WANN_MODE_FP16) CU(cudaLaunchKernelEx(&cfg, conv_stack_tc_kernel<true, false>, p));
      else CU(cudaLaunchKernelEx(&cfg, conv_stack_tc_kernel<false, false>, p));
    }
    if (t0 != nullptr) {
      CU(cudaEventRecord(t1, stream));
      std::lock_guard<std::mutex> g(ctx->ev_mu);
      ctx->ev_pairs.emplace_back(t0, t1);
    }
    ++launches;
  } else {
    const int blocks = static_cast<int>(std::min<long long>((n * 64 + 255) / 256, 65535));
    planes_to_f32_kernel<<<blocks, 256, 0, stream>>>(io.squares, io.black, io.planes, io.planes_f32, n,
                                                     ws.d_act1);
    conv3x3_f32_kernel<4><<<static_cast<unsigned>(n), 256, 100 * 4 * 4, stream>>>(
        ws.d_act1, ctx->d_conv_w[0], ctx->d_conv_b[0], ws.d_act0);
    float* cur = ws.d_act0;
    float* nxt = ws.d_act1;
    for (int l = 1; l <= 3; ++l) {
      if (io.dbg != nullptr && io.dbg_layer == l)
        CU(cudaMemcpyAsync(io.dbg, cur, n * 64 * 192 * 4, cudaMemcpyDeviceToDevice, stream));
      conv3x3_f32_kernel<192><<<static_cast<unsigned>(n), 256, 100 * 192 * 4, stream>>>(
          cur, ctx->d_conv_w[l], ctx->d_conv_b[l], nxt);
      std::swap(cur, nxt);
    }
    if (io.dbg != nullptr && io.dbg_layer == 4)
      CU(cudaMemcpyAsync(io.dbg, cur, n * 64 * 192 * 4, cudaMemcpyDeviceToDevice, stream));
    conv_small_f32_kernel<<<static_cast<unsigned>(n), 256, 0, stream>>>(
        cur, ctx->d_conv_w[4], ctx->d_conv_b[4], ctx->final_kernel, ws.d_h5);
    launches += 6;
  }
  if (io.dbg != nullptr && io.dbg_layer == 5)
    CU(cudaMemcpyAsync(io.dbg, ws.d_h5, n * 64 * 4, cudaMemcpyDeviceToDevice, stream));
  if (want_tail) {
    if (mirror_here && !fused) { t.err_src = ws.d_err; t.err_dst = io.err_mirror; }
    const long long groups = (n + kTailBoards - 1) / kTailBoards;
    const int blocks = static_cast<int>(
        std::min<long long>((groups + kTailWarps - 1) / kTailWarps, 2ll * ctx->sm_count));
    if (!fused) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(blocks);
      cfg.blockDim = dim3(kTailWarps * 32);
      cfg.dynamicSmemBytes = kTailSmemBytes;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = (io.n_dev == nullptr && ctx->pdl.load() > 0) ? 1 : 0;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      CU(cudaLaunchKernelEx(&cfg, tail_kernel, t));
      ++launches;
    }
    if (refine && !defer) {
      int rc = enqueue_refine_pass(ctx, ws, stream, io, n, -1);
      if (rc != WANN_OK) return rc;
    }
    if (io.children != nullptr) {
      const int eb = static_cast<int>(std::min<long long>((n + 7) / 8, 16ll * ctx->sm_count));
      expand_children_kernel<<<eb, 256, 0, stream>>>(io.squares, io.black, io.idx, io.nlegal, n, io.children,
                                                     io.cflags);
      ++launches;
    }
    if (io.cstats != nullptr) {
      const int eb = static_cast<int>(std::min<long long>((n + 7) / 8, 16ll * ctx->sm_count));
      seed_children_kernel<<<eb, 256, 0, stream>>>(io.squares, io.black, io.idx, io.prior, io.nlegal, n, io.cstats,
                                                   io.cflags, io.psum, io.n_dev, io.n_dev_off);
      ++launches;
    }
  }
  CU(cudaGetLastError());
  if (io.n_dev == nullptr) ctx->stats[0] += static_cast<uint64_t>(n);   // (device-side counts are added by their owner)
  ctx->stats[1] += launches;
  return WANN_OK;
}

// D2H of the slot's watchdog record and refinement counter (read after the stream has drained)
cudaError_t copy_device_err(wann_ctx*, Slot& s, cudaStream_t stream) {
  return cudaMemcpyAsync(s.h_err, s.d_err, 8 * sizeof(int), cudaMemcpyDeviceToHost, stream);
}

int check_device_err(wann_ctx* ctx, Slot& s) {
  // d_err[5]: monotonic count of positions re-evaluated by the refinement pass on this slot
  const uint32_t refined = static_cast<uint32_t>(s.h_err[5]);
  ctx->stats[7] += static_cast<uint64_t>(refined - s.refine_seen);
  s.refine_seen = refined;
  if (s.h_err[0] != 0) {
    const int code = s.h_err[0], blk = s.h_err[1], par = s.h_err[2];
    s.refine_seen = 0;
    memset(s.h_err, 0, 8 * sizeof(int));
    cudaMemset(s.d_err, 0, 8 * sizeof(int));
    return fail(WANN_E_KERNEL, "conv_stack_tc watchdog: wait site %d timed out (block %d, parity %d)",
                code, blk, par);
  }
  return WANN_OK;
}

struct EvalArgs {
  int8_t* children = nullptr;     // [n][48][64] optional (needs idx)
  uint8_t* cflags = nullptr;      // [n][48]
  int32_t* cstats = nullptr;      // [n][48][4] optional child seeds (needs idx, cflags, psum)
  int32_t* psum = nullptr;        // [n][8]
  const int8_t* squares = nullptr;
  const uint8_t* black = nullptr;
  const void* planes = nullptr;
  int planes_dtype = 0;
  float* probs = nullptr;
  float* logits = nullptr;
  uint8_t* idx = nullptr;
  float* prior = nullptr;
  uint8_t* nlegal = nullptr;
  float* dbg = nullptr;
  int dbg_layer = 0;
  int top_k = kMaxLegal;          // ranked children kept per position
  const int* n_dev = nullptr;     // optional device-side count of valid rows (<= n)
};

// the k argument of the C ABI: 0 (or anything outside 1..47) = every legal child
int clamp_top_k(int k) { return (k >= 1 && k < kMaxLegal) ? k : kMaxLegal; }

int eval_device(wann_ctx* ctx, Lane* lane, const EvalArgs& a, long long n, void* stream_v) {
  const long long cap = lane->slot[0].cap * (ctx->mode == WANN_MODE_FP32 ? 1 : kDevChunkMul);
  const size_t plane_elt = a.planes_dtype == WANN_PLANES_F32 ? 4 : 1;
  const long long dbg_c = (a.dbg_layer == 5) ? 1 : 192;
  Slot& ws = lane->slot[0];
  cudaStream_t stream = stream_v ? static_cast<cudaStream_t>(stream_v) : ws.stream;
  lane_wait(lane, stream);
  for (long long off = 0; off < n; off += cap) {
    const long long m = std::min(cap, n - off);
    ChunkIO io;
    io.squares = a.squares ? a.squares + off * 64 : nullptr;
    io.black = a.black ? a.black + off : nullptr;
    io.planes = a.planes ? static_cast<const uint8_t*>(a.planes) + off * 256 * plane_elt : nullptr;
    io.planes_f32 = a.planes_dtype == WANN_PLANES_F32;
    io.probs = a.probs ? a.probs + off * kMoves : nullptr;
    io.logits = a.logits ? a.logits + off * kMoves : nullptr;
    io.idx = a.idx ? a.idx + off * kMaxLegal : nullptr;
    io.prior = a.prior ? a.prior + off * kMaxLegal : nullptr;
    io.nlegal = a.nlegal ? a.nlegal + off : nullptr;
    io.dbg = a.dbg ? a.dbg + off * 64 * dbg_c : nullptr;
    io.dbg_layer = a.dbg_layer;
    io.top_k = a.top_k;
    io.n_dev = a.n_dev;
    io.n_dev_off = off;
    io.children = a.children ? a.children + off * kMaxLegal * 64 : nullptr;
    io.cflags = a.cflags ? a.cflags + off * kMaxLegal : nullptr;
    io.cstats = a.cstats ? a.cstats + off * kMaxLegal * 4 : nullptr;
    io.psum = a.psum ? a.psum + off * 8 : nullptr;
    int rc = enqueue_eval(ctx, ws, stream, io, m);
    if (rc != WANN_OK) return rc;
  }
  if (stream_v == nullptr) {
    CU(copy_device_err(ctx, ws, stream));
    CU(cudaStreamSynchronize(stream));
    lane->last_valid = false;
    return check_device_err(ctx, ws);
  }
  // caller's stream: later users of this lane on OTHER streams wait for this event
  CU(cudaEventRecord(lane->last, stream));
  lane->last_valid = true;
  lane->last_stream = stream;
  return WANN_OK;
}

// Small HOST calls (<= kZeroCopyMax positions -- what the reference's tree produces): no DMA copies
// at all.  Inputs are staged into pinned, device-mapped memory that the kernels read directly over
// PCIe (65 B per position), the tail kernel writes probabilities / ranked priors and the watchdog
// record straight into mapped host memory, and the call ends with one stream synchronisation.
// Replaces 2 H2D + up to 6 D2H cudaMemcpyAsync per call.
int eval_host_small(wann_ctx* ctx, Lane* lane, const EvalArgs& a, long long n) {
  Slot& s = lane->slot[0];
  const size_t plane_elt = a.planes_dtype == WANN_PLANES_F32 ? 4 : 1;
  constexpr size_t kIn = static_cast<size_t>(kZeroCopyMax) * 256 * 4;
  constexpr size_t kOutProbs = static_cast<size_t>(kZeroCopyMax) * kMoves * 4;
  constexpr size_t kOutIdx = static_cast<size_t>(kZeroCopyMax) * kMaxLegal;
  if (s.z_in == nullptr) {
    CU(cudaHostAlloc(&s.z_in, kIn, cudaHostAllocMapped));
    CU(cudaHostAlloc(&s.z_out, 2 * kOutProbs + 5 * kOutIdx + kZeroCopyMax + 64, cudaHostAllocMapped));
  }
  lane_wait(lane, s.stream);
  lane->last_valid = false;
  uint8_t* z_probs = s.z_out;
  uint8_t* z_logits = z_probs + kOutProbs;
  uint8_t* z_prior = z_logits + kOutProbs;
  uint8_t* z_idx = z_prior + 4 * kOutIdx;
  uint8_t* z_nlegal = z_idx + kOutIdx;
  int* z_err = reinterpret_cast<int*>(z_nlegal + kZeroCopyMax);   // 32 B, 4-byte aligned
  ChunkIO io;
  uint64_t h2d = 0, d2h = 0;
  if (a.squares) {
    int8_t* zs = reinterpret_cast<int8_t*>(s.z_in);
    uint8_t* zb = s.z_in + n * 64;
    memcpy(zs, a.squares, n * 64);
    memcpy(zb, a.black, n);
    io.squares = zs; io.black = zb;
    h2d = n * 65;
  } else {
    memcpy(s.z_in, a.planes, n * 256 * plane_elt);
    io.planes = s.z_in; io.planes_f32 = a.planes_dtype == WANN_PLANES_F32;
    h2d = n * 256 * plane_elt;
  }
  io.probs = a.probs ? reinterpret_cast<float*>(z_probs) : nullptr;
  io.logits = a.logits ? reinterpret_cast<float*>(z_logits) : nullptr;
  if (a.idx) { io.idx = z_idx; io.prior = reinterpret_cast<float*>(z_prior); io.nlegal = z_nlegal; }
  io.err_mirror = z_err;
  io.top_k = a.top_k;
  // margin refinement: the host is about to synchronise anyway, so it reads the number of flagged
  // boards itself and launches pass 2 only when there are any (98 % of batch-1 calls: none)
  io.allow_defer = true;
  bool deferred = false;
  int rc = enqueue_eval(ctx, s, s.stream, io, n, &deferred);
  if (rc != WANN_OK) return rc;
  CU(cudaStreamSynchronize(s.stream));
  if (deferred && z_err[0] == 0 && z_err[4] > 0) {
    rc = enqueue_refine_pass(ctx, s, s.stream, io, n, z_err[4]);
    if (rc != WANN_OK) return rc;
    CU(cudaStreamSynchronize(s.stream));
  }
  memcpy(s.h_err, z_err, 8 * sizeof(int));
  rc = check_device_err(ctx, s);
  if (rc != WANN_OK) return rc;
  if (a.probs) { memcpy(a.probs, z_probs, n * kMoves * 4); d2h += n * kMoves * 4; }
  if (a.logits) { memcpy(a.logits, z_logits, n * kMoves * 4); d2h += n * kMoves * 4; }
  if (a.idx) {
    memcpy(a.idx, z_idx, n * kMaxLegal);
    memcpy(a.prior, z_prior, n * kMaxLegal * 4);
    memcpy(a.nlegal, z_nlegal, n);
    d2h += n * (kMaxLegal * 5 + 1);
  }
  ctx->stats[3] += h2d;
  ctx->stats[4] += d2h;
  return WANN_OK;
}

// HOST memory: chunked, double-buffered over the lane's two slots so that the H2D copy of
// chunk i+1 and the D2H copy of chunk i-1 overlap the kernels of chunk i.
int eval_host(wann_ctx* ctx, Lane* lane, const EvalArgs& a, long long n) {
  // (the zero-copy path runs the whole call as ONE chunk on the slot's workspaces: only when it fits them
  // -- the "chunk" option may be as small as 8)
  const long long zc_cap = std::min<long long>(
      kZeroCopyMax, lane->slot[0].cap * (ctx->mode == WANN_MODE_FP32 ? 1 : kDevChunkMul));
  if (ctx->zero_copy && n <= zc_cap && a.dbg == nullptr && a.children == nullptr && a.cstats == nullptr &&
      (a.probs != nullptr || a.logits != nullptr || a.idx != nullptr))
    return eval_host_small(ctx, lane, a, n);
  const long long cap = lane->slot[0].cap;
  const size_t plane_elt = a.planes_dtype == WANN_PLANES_F32 ? 4 : 1;
  const long long dbg_c = (a.dbg_layer == 5) ? 1 : 192;
  uint64_t h2d = 0, d2h = 0;
  long long ci = 0;
  lane_wait(lane, lane->slot[0].stream);
  lane_wait(lane, lane->slot[1].stream);
  lane->last_valid = false;          // this call ends with host synchronisation of both slots
  for (long long off = 0; off < n; off += cap, ++ci) {
    const long long m = std::min(cap, n - off);
    Slot& s = lane->slot[ci & 1];
    if (s.busy) {
      s.busy = false;
      CU(cudaEventSynchronize(s.done));
      int rc = check_device_err(ctx, s);
      if (rc != WANN_OK) return rc;
    }
    ChunkIO io;
    io.top_k = a.top_k;
    if (a.squares) {
      CU(cudaMemcpyAsync(s.d_squares, a.squares + off * 64, m * 64, cudaMemcpyHostToDevice, s.stream));
      CU(cudaMemcpyAsync(s.d_black, a.black + off, m, cudaMemcpyHostToDevice, s.stream));
      h2d += m * 65;
      io.squares = s.d_squares; io.black = s.d_black;
    } else {
      CU(cudaMemcpyAsync(s.d_planes, static_cast<const uint8_t*>(a.planes) + off * 256 * plane_elt,
                         m * 256 * plane_elt, cudaMemcpyHostToDevice, s.stream));
      h2d += m * 256 * plane_elt;
      io.planes = s.d_planes; io.planes_f32 = a.planes_dtype == WANN_PLANES_F32;
    }
    io.probs = a.probs ? s.d_probs : nullptr;
    io.logits = a.logits ? s.d_logits : nullptr;
    if (a.idx) { io.idx = s.d_idx; io.prior = s.d_prior; io.nlegal = s.d_nlegal; }
    if (a.children) {
      if (s.d_children == nullptr) CU(cudaMalloc(&s.d_children, s.cap * kMaxLegal * 64));
      if (s.d_cflags == nullptr) CU(cudaMalloc(&s.d_cflags, s.cap * kMaxLegal));
      io.children = s.d_children; io.cflags = s.d_cflags;
    }
    if (a.cstats) {
      if (s.d_cflags == nullptr) CU(cudaMalloc(&s.d_cflags, s.cap * kMaxLegal));
      if (s.d_cstats == nullptr) {
        CU(cudaMalloc(&s.d_cstats, s.cap * kMaxLegal * 4 * sizeof(int32_t)));
        CU(cudaMalloc(&s.d_psum, s.cap * 8 * sizeof(int32_t)));
      }
      io.cflags = s.d_cflags; io.cstats = s.d_cstats; io.psum = s.d_psum;
    }
    float* d_dbg = nullptr;
    if (a.dbg) {
      CU(cudaMalloc(&d_dbg, m * 64 * dbg_c * 4));
      io.dbg = d_dbg; io.dbg_layer = a.dbg_layer;
    }
    int rc = enqueue_eval(ctx, s, s.stream, io, m);
    if (rc == WANN_OK) {
      cudaError_t e = cudaSuccess;
      auto D2H = [&](void* h, const void* d, size_t bytes) {
        d2h += bytes;
        if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s.stream);
      };
      if (a.probs) D2H(a.probs + off * kMoves, s.d_probs, m * kMoves * 4);
      if (a.logits) D2H(a.logits + off * kMoves, s.d_logits, m * kMoves * 4);
      if (a.idx) {
        D2H(a.idx + off * kMaxLegal, s.d_idx, m * kMaxLegal);
        D2H(a.prior + off * kMaxLegal, s.d_prior, m * kMaxLegal * 4);
        D2H(a.nlegal + off, s.d_nlegal, m);
      }
      if (a.children) D2H(a.children + off * kMaxLegal * 64, s.d_children, m * kMaxLegal * 64);
      if (a.cflags) D2H(a.cflags + off * kMaxLegal, s.d_cflags, m * kMaxLegal);
      if (a.cstats) {
        D2H(a.cstats + off * kMaxLegal * 4, s.d_cstats, m * kMaxLegal * 4 * sizeof(int32_t));
        D2H(a.psum + off * 8, s.d_psum, m * 8 * sizeof(int32_t));
      }
      if (a.dbg) D2H(a.dbg + off * 64 * dbg_c, d_dbg, m * 64 * dbg_c * 4);
      if (e == cudaSuccess)
        e = copy_device_err(ctx, s, s.stream);
      if (e == cudaSuccess) e = cudaEventRecord(s.done, s.stream);
      if (e == cudaSuccess) s.busy = true;
      else rc = fail(WANN_E_CUDA, "D2H enqueue: %s", cudaGetErrorString(e));
    }
    if (d_dbg) { cudaStreamSynchronize(s.stream); cudaFree(d_dbg); }
    if (rc != WANN_OK) return rc;
  }
  int rc = WANN_OK;
  for (int k = 0; k < 2; ++k) {
    Slot& s = lane->slot[k];
    if (s.busy) {
      s.busy = false;
      cudaError_t e = cudaEventSynchronize(s.done);
      if (e != cudaSuccess && rc == WANN_OK) rc = fail(WANN_E_CUDA, "event sync: %s", cudaGetErrorString(e));
      if (rc == WANN_OK) rc = check_device_err(ctx, s);
    }
  }
  ctx->stats[3] += h2d;
  ctx->stats[4] += d2h;
  return rc;
}

int run_eval(wann_ctx* ctx, const EvalArgs& a, long long n, int mem, void* stream_v);

int combine_alloc(wann_ctx* ctx) {
  if (ctx->c_sq != nullptr) return WANN_OK;
  const long long m = kCombineMaxBatch;
  CU(cudaMallocHost(&ctx->c_sq, m * 64));
  CU(cudaMallocHost(&ctx->c_bl, m));
  CU(cudaMallocHost(&ctx->c_probs, m * kMoves * 4));
  CU(cudaMallocHost(&ctx->c_logits, m * kMoves * 4));
  CU(cudaMallocHost(&ctx->c_idx, m * kMaxLegal));
  CU(cudaMallocHost(&ctx->c_prior, m * kMaxLegal * 4));
  CU(cudaMallocHost(&ctx->c_nlegal, m));
  return WANN_OK;
}

// Flat combining: enqueue; if nobody is processing, become the leader, take a batch of queued
// requests (FIFO, up to kCombineMaxBatch positions), evaluate them as ONE host-path call on the
// pinned staging buffers, scatter the results and wake the owners.  A lone caller therefore pays
// no hand-off latency; concurrent callers share launches.
int run_combined(wann_ctx* ctx, const EvalArgs& a, long long n) {
  CombineReq req;
  req.squares = a.squares; req.black = a.black; req.n = n; req.probs = a.probs; req.logits = a.logits;
  req.idx = a.idx; req.prior = a.prior; req.nlegal = a.nlegal;
  std::unique_lock<std::mutex> lk(ctx->cmu);
  ctx->cq.push_back(&req);
  while (!req.done) {
    if (ctx->cbusy) { ctx->ccv.wait(lk); continue; }
    ctx->cbusy = true;
    std::vector<CombineReq*> batch;
    long long total = 0;
    while (!ctx->cq.empty() && total + ctx->cq.front()->n <= kCombineMaxBatch) {
      batch.push_back(ctx->cq.front());
      total += ctx->cq.front()->n;
      ctx->cq.pop_front();
    }
    lk.unlock();
    int rc = combine_alloc(ctx);
    EvalArgs m;
    if (rc == WANN_OK) {
      long long off = 0;
      for (CombineReq* r : batch) {
        memcpy(ctx->c_sq + off * 64, r->squares, r->n * 64);
        memcpy(ctx->c_bl + off, r->black, r->n);
        if (r->probs) m.probs = ctx->c_probs;
        if (r->logits) m.logits = ctx->c_logits;
        if (r->idx) { m.idx = ctx->c_idx; m.prior = ctx->c_prior; m.nlegal = ctx->c_nlegal; }
        off += r->n;
      }
      m.squares = ctx->c_sq; m.black = ctx->c_bl;
      rc = run_eval(ctx, m, -total, WANN_MEM_HOST, nullptr);   // negative n: bypass the combiner
    }
    const std::string err = g_err;
    long long off = 0;
    for (CombineReq* r : batch) {
      if (rc == WANN_OK) {
        if (r->probs) memcpy(r->probs, ctx->c_probs + off * kMoves, r->n * kMoves * 4);
        if (r->logits) memcpy(r->logits, ctx->c_logits + off * kMoves, r->n * kMoves * 4);
        if (r->idx) {
          memcpy(r->idx, ctx->c_idx + off * kMaxLegal, r->n * kMaxLegal);
          memcpy(r->prior, ctx->c_prior + off * kMaxLegal, r->n * kMaxLegal * 4);
          memcpy(r->nlegal, ctx->c_nlegal + off, r->n);
        }
      }
      off += r->n;
    }
    ctx->stats[5] += 1;                                   // merged launches
    ctx->stats[6] += static_cast<uint64_t>(batch.size()); // requests served by them
    lk.lock();
    for (CombineReq* r : batch) { r->rc = rc; r->err = err; r->done = true; }
    ctx->cbusy = false;
    ctx->ccv.notify_all();
  }
  lk.unlock();
  if (req.rc != WANN_OK) g_err = req.err;
  return req.rc;
}

// The one driver behind wann_eval_probs / _planes / _topk / _debug_activations.
int run_eval(wann_ctx* ctx, const EvalArgs& a, long long n, int mem, void* stream_v) {
  bool bypass = false;
  if (n < 0) { n = -n; bypass = true; }
  if (ctx == nullptr) return fail(WANN_E_INVALID, "null context");
  if (n == 0) return WANN_OK;
  if (mem != WANN_MEM_HOST && mem != WANN_MEM_DEVICE) return fail(WANN_E_INVALID, "bad mem kind %d", mem);
  if (a.squares == nullptr && a.planes == nullptr) return fail(WANN_E_INVALID, "no input");
  if (a.squares != nullptr && a.black == nullptr) return fail(WANN_E_INVALID, "black_to_move is null");
  if (a.idx != nullptr && (a.squares == nullptr || a.prior == nullptr || a.nlegal == nullptr))
    return fail(WANN_E_INVALID, "top-k needs squares, prior and n_legal");
  if (mem == WANN_MEM_DEVICE && a.idx != nullptr &&
      ((reinterpret_cast<uintptr_t>(a.idx) | reinterpret_cast<uintptr_t>(a.prior)) & 15u))
    return fail(WANN_E_INVALID, "device idx / prior buffers must be 16-byte aligned");
  DeviceGuard dev_guard(ctx->device);
  CU(dev_guard.status);
  ReaderScope weights_scope(bypass ? nullptr : &ctx->weights_gate);   // (a combined call runs under its leader's)
  if (!bypass) ctx->stats[2] += 1;
  if (!bypass && ctx->coalesce && mem == WANN_MEM_HOST && a.squares != nullptr && a.dbg == nullptr &&
      a.children == nullptr && a.cstats == nullptr && a.top_k == kMaxLegal && n <= kCombineMaxCall)
    return run_combined(ctx, a, n);
  Lane* lane = nullptr;
  int rc = acquire_lane(ctx, &lane);
  if (rc != WANN_OK) return rc;
  rc = (mem == WANN_MEM_DEVICE) ? eval_device(ctx, lane, a, n, stream_v) : eval_host(ctx, lane, a, n);
  if (rc != WANN_OK) {   // drain whatever is still in flight on this lane before handing it back
    for (int k = 0; k < 2; ++k) { cudaStreamSynchronize(lane->slot[k].stream); lane->slot[k].busy = false; }
  }
  release_lane(ctx, lane);
  return rc;
}

}  // namespace

// =============================================================================================
extern "C" {

int wann_version(void) { return 200; }

const char* wann_last_error(void) { return g_err.c_str(); }

int wann_create(const wann_weights* w, int device, int mode, wann_ctx** out_ctx) {
  if (out_ctx == nullptr || w == nullptr) return fail(WANN_E_INVALID, "null argument");
  *out_ctx = nullptr;
  if (mode != WANN_MODE_BF16 && mode != WANN_MODE_FP32 && mode != WANN_MODE_FP16 && mode != WANN_MODE_FP32_TC &&
      mode != WANN_MODE_FP16C)
    return fail(WANN_E_INVALID, "bad mode %d", mode);
  if (w->final_kernel != 3 && w->final_kernel != 1)
    return fail(WANN_E_INVALID, "final_kernel must be 3 or 1");
  for (int i = 0; i < 5; ++i)
    if (!w->conv_w[i] || !w->conv_b[i]) return fail(WANN_E_INVALID, "null conv weight %d", i + 1);
  if (!w->fc_w || !w->fc_b) return fail(WANN_E_INVALID, "null output_layer weights");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(WANN_E_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(WANN_E_NO_DEVICE, "device %d out of range (%d)", device, count);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(WANN_E_NO_DEVICE, "cannot query device");
  if (prop.major != 10)
    return fail(WANN_E_NO_DEVICE, "device %d is sm_%d%d; this library is sm_100a (B200) only", device,
                prop.major, prop.minor);
  DeviceGuard dev_guard(device);
  CU(dev_guard.status);
  wann_ctx* ctx = new wann_ctx();
  for (auto& s : ctx->stats) s = 0;
  ctx->device = device;
  ctx->mode = mode;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->final_kernel = w->final_kernel;
  int rc = WANN_OK;
  {
    e = cudaFuncSetAttribute(conv_stack_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_stack_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_stack_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_stack_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv3x3_f32_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               100 * 192 * 4);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailSmemBytes);
    // keep the SM's L1/shared split at "max shared" for the small kernels too, so that alternating
    // conv-stack (227 KB) and tail launches never reconfigure the carve-out
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(tail_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(expand_children_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) rc = fail(WANN_E_CUDA, "kernel attributes: %s", cudaGetErrorString(e));
  }
  if (rc == WANN_OK) rc = upload_weights(ctx, *w);   // (the centred mode calibrates here: kernels run)
  if (rc != WANN_OK) {
    wann_destroy(ctx);
    return rc;
  }
  *out_ctx = ctx;
  return WANN_OK;
}

int wann_set_weights(wann_ctx* ctx, const wann_weights* w) {
  if (ctx == nullptr || w == nullptr) return fail(WANN_E_INVALID, "null argument");
  if (w->final_kernel != ctx->final_kernel)
    return fail(WANN_E_INVALID, "final_kernel %d differs from the context's %d", w->final_kernel, ctx->final_kernel);
  for (int i = 0; i < 5; ++i)
    if (!w->conv_w[i] || !w->conv_b[i]) return fail(WANN_E_INVALID, "null conv weight %d", i + 1);
  if (!w->fc_w || !w->fc_b) return fail(WANN_E_INVALID, "null output_layer weights");
  DeviceGuard dev_guard(ctx->device);
  CU(dev_guard.status);
  // new evaluations wait; asynchronous ones already enqueued finish with the old weights
  ctx->weights_gate.writer_enter();
  int rc = WANN_OK;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) rc = fail(WANN_E_CUDA, "device synchronise: %s", cudaGetErrorString(e));
  else rc = upload_weights(ctx, *w);
  ctx->weights_gate.writer_exit();
  return rc;
}

int wann_destroy(wann_ctx* ctx) {
  if (ctx == nullptr) return WANN_OK;
  DeviceGuard dev_guard(ctx->device);
  cudaDeviceSynchronize();
  for (auto& pr : ctx->ev_pairs) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  ctx->ev_pairs.clear();
  for (auto& l : ctx->lanes) {
    free_slot(l.slot[0]);
    free_slot(l.slot[1]);
    if (l.last) cudaEventDestroy(l.last);
  }
  if (ctx->c_sq) {
    cudaFreeHost(ctx->c_sq); cudaFreeHost(ctx->c_bl); cudaFreeHost(ctx->c_probs); cudaFreeHost(ctx->c_logits);
    cudaFreeHost(ctx->c_idx); cudaFreeHost(ctx->c_prior); cudaFreeHost(ctx->c_nlegal);
  }
  for (int i = 0; i < 5; ++i) { cudaFree(ctx->d_conv_w[i]); cudaFree(ctx->d_conv_b[i]); }
  cudaFree(ctx->d_fc_w); cudaFree(ctx->d_fc_b); cudaFree(ctx->d_wimg); cudaFree(ctx->d_wimg_split); cudaFree(ctx->d_bias_tc);
  cudaFree(ctx->d_bias_plain); cudaFree(ctx->d_unperm);
  delete ctx;
  return WANN_OK;
}

int wann_eval_probs(wann_ctx* ctx, const int8_t* squares, const uint8_t* black, int64_t n, float* probs,
                    float* logits, int mem, void* stream) {
  if (probs == nullptr && logits == nullptr) return fail(WANN_E_INVALID, "no output buffer");
  if (n < 0) return fail(WANN_E_INVALID, "negative n");
  EvalArgs a;
  a.squares = squares; a.black = black; a.probs = probs; a.logits = logits;
  if (squares == nullptr && n > 0) return fail(WANN_E_INVALID, "squares is null");
  return run_eval(ctx, a, n, mem, stream);
}

int wann_eval_planes(wann_ctx* ctx, const void* planes, int planes_dtype, int64_t n, float* probs,
                     float* logits, int mem, void* stream) {
  if (probs == nullptr && logits == nullptr) return fail(WANN_E_INVALID, "no output buffer");
  if (planes == nullptr && n > 0) return fail(WANN_E_INVALID, "planes is null");
  if (planes_dtype != WANN_PLANES_I8 && planes_dtype != WANN_PLANES_F32)
    return fail(WANN_E_INVALID, "bad planes dtype");
  if (n < 0) return fail(WANN_E_INVALID, "negative n");
  EvalArgs a;
  a.planes = planes; a.planes_dtype = planes_dtype; a.probs = probs; a.logits = logits;
  return run_eval(ctx, a, n, mem, stream);
}

int wann_eval_topk(wann_ctx* ctx, const int8_t* squares, const uint8_t* black, int64_t n, int k, uint8_t* idx,
                   float* prior, uint8_t* n_legal, int mem, void* stream) {
  if ((idx == nullptr || prior == nullptr || n_legal == nullptr) && n > 0)
    return fail(WANN_E_INVALID, "null output buffer");
  if (squares == nullptr && n > 0) return fail(WANN_E_INVALID, "squares is null");
  if (n < 0) return fail(WANN_E_INVALID, "negative n");
  EvalArgs a;
  a.squares = squares; a.black = black; a.idx = idx; a.prior = prior; a.nlegal = n_legal; a.top_k = clamp_top_k(k);
  return run_eval(ctx, a, n, mem, stream);
}

int wann_eval_expand(wann_ctx* ctx, const int8_t* squares, const uint8_t* black, int64_t n, int k, uint8_t* idx,
                     float* prior, uint8_t* n_legal, int8_t* child_squares, uint8_t* child_flags, int mem,
                     void* stream) {
  if ((idx == nullptr || prior == nullptr || n_legal == nullptr || child_squares == nullptr ||
       child_flags == nullptr) && n > 0)
    return fail(WANN_E_INVALID, "null output buffer");
  if (squares == nullptr && n > 0) return fail(WANN_E_INVALID, "squares is null");
  if (n < 0) return fail(WANN_E_INVALID, "negative n");
  EvalArgs a;
  a.squares = squares; a.black = black; a.idx = idx; a.prior = prior; a.nlegal = n_legal;
  a.children = child_squares; a.cflags = child_flags; a.top_k = clamp_top_k(k);
  return run_eval(ctx, a, n, mem, stream);
}

int wann_eval_expand_seeded(wann_ctx* ctx, const int8_t* squares, const uint8_t* black, int64_t n, int k, uint8_t* idx,
                            float* prior, uint8_t* n_legal, int8_t* child_squares, uint8_t* child_flags,
                            int32_t* child_stats, int32_t* parent_summary, int mem, void* stream) {
  if ((idx == nullptr || prior == nullptr || n_legal == nullptr || child_flags == nullptr ||
       child_stats == nullptr || parent_summary == nullptr) && n > 0)
    return fail(WANN_E_INVALID, "null output buffer");
  if (squares == nullptr && n > 0) return fail(WANN_E_INVALID, "squares is null");
  if (n < 0) return fail(WANN_E_INVALID, "negative n");
  if (mem == WANN_MEM_DEVICE &&
      ((reinterpret_cast<uintptr_t>(child_stats) | reinterpret_cast<uintptr_t>(parent_summary)) & 15u))
    return fail(WANN_E_INVALID, "device child_stats / parent_summary buffers must be 16-byte aligned");
  EvalArgs a;
  a.squares = squares; a.black = black; a.idx = idx; a.prior = prior; a.nlegal = n_legal;
  a.children = child_squares; a.cflags = child_flags; a.cstats = child_stats; a.psum = parent_summary;
  a.top_k = clamp_top_k(k);
  return run_eval(ctx, a, n, mem, stream);
}

int wann_expand_ranked(wann_ctx* ctx, const int8_t* squares, const uint8_t* black, int64_t n, const uint8_t* idx,
                       const float* prior, const uint8_t* n_legal, int8_t* child_squares, uint8_t* child_flags,
                       int32_t* child_stats, int32_t* parent_summary, void* stream_v) {
  if (ctx == nullptr) return fail(WANN_E_INVALID, "null context");
  if (n < 0) return fail(WANN_E_INVALID, "negative n");
  if (n == 0) return WANN_OK;
  if (!squares || !black || !idx || !prior || !n_legal || !child_flags)
    return fail(WANN_E_INVALID, "null buffer");
  if (child_squares == nullptr && child_stats == nullptr) return fail(WANN_E_INVALID, "nothing to produce");
  if (child_stats != nullptr && parent_summary == nullptr) return fail(WANN_E_INVALID, "child_stats needs parent_summary");
  if ((reinterpret_cast<uintptr_t>(child_squares) | reinterpret_cast<uintptr_t>(child_stats) |
       reinterpret_cast<uintptr_t>(parent_summary) | reinterpret_cast<uintptr_t>(idx) |
       reinterpret_cast<uintptr_t>(prior)) & 15u)
    return fail(WANN_E_INVALID, "device buffers must be 16-byte aligned");
  DeviceGuard dev_guard(ctx->device);
  CU(dev_guard.status);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int eb = static_cast<int>(std::min<long long>((n + 7) / 8, 16ll * ctx->sm_count));
  if (child_squares != nullptr) {
    expand_children_kernel<<<eb, 256, 0, stream>>>(squares, black, idx, n_legal, n, child_squares, child_flags);
    ctx->stats[1] += 1;
  }
  if (child_stats != nullptr) {
    seed_children_kernel<<<eb, 256, 0, stream>>>(squares, black, idx, prior, n_legal, n, child_stats, child_flags,
                                                 parent_summary);
    ctx->stats[1] += 1;
  }
  CU(cudaGetLastError());
  if (stream_v == nullptr) CU(cudaStreamSynchronize(stream));
  return WANN_OK;
}

int wann_tree_step(wann_ctx* ctx, int8_t* squares, uint8_t* black, int64_t n, int greedy, uint64_t seed,
                   uint64_t step, uint8_t* chosen_idx, uint8_t* finished, void* stream_v) {
  if (ctx == nullptr) return fail(WANN_E_INVALID, "null context");
  if (n < 0) return fail(WANN_E_INVALID, "negative n");
  if (n == 0) return WANN_OK;
  if (squares == nullptr || black == nullptr) return fail(WANN_E_INVALID, "null buffer");
  DeviceGuard dev_guard(ctx->device);
  CU(dev_guard.status);
  ReaderScope weights_scope(&ctx->weights_gate);
  ctx->stats[2] += 1;
  Lane* lane = nullptr;
  int rc = acquire_lane(ctx, &lane);
  if (rc != WANN_OK) return rc;
  Slot& ws = lane->slot[0];
  cudaStream_t stream = stream_v ? static_cast<cudaStream_t>(stream_v) : ws.stream;
  lane_wait(lane, stream);
  auto body = [&]() -> int {
    for (long long off = 0; off < n; off += ws.cap) {
      const long long m = std::min<long long>(ws.cap, n - off);
      ChunkIO io;
      io.squares = squares + off * 64; io.black = black + off;
      io.idx = ws.d_idx; io.prior = ws.d_prior; io.nlegal = ws.d_nlegal;
      int r = enqueue_eval(ctx, ws, stream, io, m);
      if (r != WANN_OK) return r;
      tree_step_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, stream>>>(
          squares + off * 64, black + off, ws.d_idx, ws.d_prior, ws.d_nlegal, m, off, greedy, seed, step,
          chosen_idx ? chosen_idx + off : nullptr, finished ? finished + off : nullptr);
      CU(cudaGetLastError());
      ctx->stats[1] += 1;
    }
    if (stream_v == nullptr) {
      CU(cudaStreamSynchronize(stream));
      lane->last_valid = false;
    } else {
      CU(cudaEventRecord(lane->last, stream));
      lane->last_valid = true;
      lane->last_stream = stream;
    }
    return WANN_OK;
  };
  rc = body();
  if (rc != WANN_OK) cudaStreamSynchronize(stream);
  release_lane(ctx, lane);
  return rc;
}

int wann_debug_activations(wann_ctx* ctx, const int8_t* squares, const uint8_t* black, int64_t n, int layer,
                           float* out, int mem) {
  if (layer < 1 || layer > 5) return fail(WANN_E_INVALID, "layer must be 1..5");
  if ((squares == nullptr || out == nullptr) && n > 0) return fail(WANN_E_INVALID, "null buffer");
  if (n < 0) return fail(WANN_E_INVALID, "negative n");
  EvalArgs a;
  a.squares = squares; a.black = black; a.dbg = out; a.dbg_layer = layer;
  return run_eval(ctx, a, n, mem, nullptr);
}

// encode / legal mask: standalone byte kernels (parity surface of tools.utils / board_utils)
static int run_bytes(wann_ctx* ctx, const int8_t* squares, const uint8_t* black, int64_t n, void* out,
                     int out_per_pos, bool is_mask, int mem, void* stream_v) {
  if (ctx == nullptr) return fail(WANN_E_INVALID, "null context");
  if (n < 0) return fail(WANN_E_INVALID, "negative n");
  if (n == 0) return WANN_OK;
  if (!squares || !black || !out) return fail(WANN_E_INVALID, "null buffer");
  if (mem == WANN_MEM_DEVICE && !is_mask &&
      ((reinterpret_cast<uintptr_t>(squares) & 3u) || (reinterpret_cast<uintptr_t>(out) & 15u)))
    return fail(WANN_E_INVALID, "device squares must be 4-byte and planes 16-byte aligned");
  if (mem == WANN_MEM_DEVICE && is_mask &&
      ((reinterpret_cast<uintptr_t>(squares) | reinterpret_cast<uintptr_t>(out)) & 15u))
    return fail(WANN_E_INVALID, "device squares and the mask buffer must be 16-byte aligned");
  DeviceGuard dev_guard(ctx->device);
  CU(dev_guard.status);
  Lane* lane = nullptr;
  int rc = acquire_lane(ctx, &lane);
  if (rc != WANN_OK) return rc;
  Slot& s = lane->slot[0];
  cudaStream_t stream = (mem == WANN_MEM_DEVICE && stream_v) ? static_cast<cudaStream_t>(stream_v) : s.stream;
  lane_wait(lane, stream);
  auto body = [&]() -> int {
    // HOST buffers are staged through the slot (chunks of its capacity); DEVICE buffers need no
    // staging: one launch over all n positions (the kernels are grid-stride loops)
    const long long step = (mem == WANN_MEM_HOST) ? s.cap : n;
    for (long long off = 0; off < n; off += step) {
      const long long m = std::min<long long>(step, n - off);
      const int8_t* dsq = squares + off * 64;
      const uint8_t* dbl = black + off;
      uint8_t* dout = static_cast<uint8_t*>(out) + off * out_per_pos;
      if (mem == WANN_MEM_HOST) {
        CU(cudaMemcpyAsync(s.d_squares, dsq, m * 64, cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(s.d_black, dbl, m, cudaMemcpyHostToDevice, stream));
        dsq = s.d_squares; dbl = s.d_black;
        dout = is_mask ? s.d_mask : reinterpret_cast<uint8_t*>(s.d_planes_out);
      }
      if (is_mask) {
        const long long per_block = kMaskWarps * 32;
        const int blocks = static_cast<int>(std::min<long long>((m + per_block - 1) / per_block, 11ll * ctx->sm_count));
        legal_mask_kernel<<<blocks, kMaskWarps * 32, 0, stream>>>(dsq, dbl, m, dout);
      } else {
        // (experiments: option "variant" bits 8-10 = unroll / store kind, bits 12-15 = blocks per SM)
        const int var = ctx->variant.load();
        const int per_sm = ((var >> 12) & 15) ? ((var >> 12) & 15) : 16;
        const int blocks = static_cast<int>(std::min<long long>((m * 16 + 255) / 256, static_cast<long long>(per_sm) * ctx->sm_count));
        int8_t* po = reinterpret_cast<int8_t*>(dout);
        switch ((var >> 8) & 7) {
          case 1: encode_kernel<2, true><<<blocks, 256, 0, stream>>>(dsq, dbl, m, po); break;
          case 2: encode_kernel<8, true><<<blocks, 256, 0, stream>>>(dsq, dbl, m, po); break;
          case 3: encode_kernel<4, false><<<blocks, 256, 0, stream>>>(dsq, dbl, m, po); break;
          case 4: encode_kernel<8, false><<<blocks, 256, 0, stream>>>(dsq, dbl, m, po); break;
          case 5: encode_kernel<1, false><<<blocks, 256, 0, stream>>>(dsq, dbl, m, po); break;
          default: encode_kernel<4, true><<<blocks, 256, 0, stream>>>(dsq, dbl, m, po); break;
        }
      }
      CU(cudaGetLastError());
      ctx->stats[1] += 1;
      if (mem == WANN_MEM_HOST) {
        CU(cudaMemcpyAsync(static_cast<uint8_t*>(out) + off * out_per_pos, dout,
                           static_cast<size_t>(m) * out_per_pos, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
      }
    }
    if (mem == WANN_MEM_DEVICE) {
      if (stream_v == nullptr) CU(cudaStreamSynchronize(stream));
      else { cudaEventRecord(lane->last, stream); lane->last_valid = true; lane->last_stream = stream; }
    }
    return WANN_OK;
  };
  rc = body();
  release_lane(ctx, lane);
  return rc;
}

int wann_encode(wann_ctx* ctx, const int8_t* squares, const uint8_t* black, int64_t n, int8_t* planes, int mem,
                void* stream) {
  return run_bytes(ctx, squares, black, n, planes, 256, false, mem, stream);
}

int wann_legal_mask(wann_ctx* ctx, const int8_t* squares, const uint8_t* black, int64_t n, uint8_t* mask,
                    int mem, void* stream) {
  return run_bytes(ctx, squares, black, n, mask, kMoves, true, mem, stream);
}

int wann_debug_profile(wann_ctx* ctx, const int8_t* squares, const uint8_t* black, int64_t n,
                       int64_t* stamps /* [4][64] */) {
  if (ctx == nullptr || !squares || !black || !stamps || n <= 0) return fail(WANN_E_INVALID, "bad argument");
  if (ctx->mode == WANN_MODE_FP32 || ctx->mode == WANN_MODE_FP32_TC)
    return fail(WANN_E_INVALID, "profile needs the bf16/fp16 tensor-core mode");
  DeviceGuard dev_guard(ctx->device);
  CU(dev_guard.status);
  Lane* lane = nullptr;
  int rc = acquire_lane(ctx, &lane);
  if (rc != WANN_OK) return rc;
  Slot& s = lane->slot[0];
  lane_wait(lane, s.stream);
  auto body = [&]() -> int {
    if (n > s.cap) return fail(WANN_E_INVALID, "n exceeds the chunk size");
    long long* d_prof = nullptr;
    CU(cudaMalloc(&d_prof, 8 * kProfSteps * 8));
    CU(cudaMemsetAsync(d_prof, 0, 8 * kProfSteps * 8, s.stream));
    CU(cudaMemcpyAsync(s.d_squares, squares, n * 64, cudaMemcpyHostToDevice, s.stream));
    CU(cudaMemcpyAsync(s.d_black, black, n, cudaMemcpyHostToDevice, s.stream));
    ChunkIO io;
    io.squares = s.d_squares; io.black = s.d_black; io.prof = d_prof;
    io.idx = s.d_idx; io.prior = s.d_prior; io.nlegal = s.d_nlegal;
    int r = enqueue_eval(ctx, s, s.stream, io, n);
    if (r == WANN_OK) {
      cudaMemcpyAsync(stamps, d_prof, 8 * kProfSteps * 8, cudaMemcpyDeviceToHost, s.stream);
      cudaError_t e = cudaStreamSynchronize(s.stream);
      if (e != cudaSuccess) r = fail(WANN_E_CUDA, "sync: %s", cudaGetErrorString(e));
    }
    cudaFree(d_prof);
    return r;
  };
  rc = body();
  release_lane(ctx, lane);
  return rc;
}

int wann_kernel_times(wann_ctx* ctx, double out[2]) {
  if (ctx == nullptr || out == nullptr) return fail(WANN_E_INVALID, "null argument");
  DeviceGuard dev_guard(ctx->device);
  CU(dev_guard.status);
  CU(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> g(ctx->ev_mu);
  double ms = 0;
  for (auto& pr : ctx->ev_pairs) {
    float t = 0;
    if (cudaEventElapsedTime(&t, pr.first, pr.second) == cudaSuccess) ms += t;
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  out[0] = ms;
  out[1] = static_cast<double>(ctx->ev_pairs.size());
  ctx->ev_pairs.clear();
  return WANN_OK;
}

int wann_debug_centering(wann_ctx* ctx, float offsets[96], uint8_t position_of[768], float kappa[9]) {
  if (ctx == nullptr || offsets == nullptr || position_of == nullptr) return fail(WANN_E_INVALID, "null argument");
  if (ctx->mode != WANN_MODE_FP16C) return fail(WANN_E_INVALID, "centring exists in the fp16c mode only");
  ReaderScope weights_scope(&ctx->weights_gate);
  memcpy(offsets, ctx->cen.c, 4 * kCenBuckets * sizeof(float));
  memcpy(position_of, ctx->cen.pos_of, 768);
  if (kappa != nullptr) memcpy(kappa, ctx->kappa, 9 * sizeof(float));
  return WANN_OK;
}

int wann_debug_effective_weights(wann_ctx* ctx, int layer, float* out) {
  if (ctx == nullptr || out == nullptr) return fail(WANN_E_INVALID, "null argument");
  if (layer < 2 || layer > 4) return fail(WANN_E_INVALID, "layer must be 2..4");
  if (ctx->mode == WANN_MODE_FP32 || ctx->mode == WANN_MODE_FP32_TC) return fail(WANN_E_INVALID, "16-bit tensor-core modes only");
  ReaderScope weights_scope(&ctx->weights_gate);
  const std::vector<float>& we = ctx->w_eff[layer - 2];
  if (we.empty()) return fail(WANN_E_INVALID, "no weights uploaded");
  memcpy(out, we.data(), we.size() * sizeof(float));
  return WANN_OK;
}

int wann_gptq_round(const float* H, int n, float* W, int rows, const uint8_t* exact, double damp) {
  if (H == nullptr || W == nullptr || n <= 0 || rows <= 0) return fail(WANN_E_INVALID, "null argument");
  std::vector<double> U;
  if (!wann_gptq::inverse_factor(H, n, damp, U)) return fail(WANN_E_INVALID, "matrix not positive definite");
  std::vector<double> Wd(W, W + static_cast<size_t>(rows) * n);
  wann_gptq::round_rows(U.data(), n, Wd.data(), rows, exact);
  for (size_t i = 0; i < Wd.size(); ++i) W[i] = static_cast<float>(Wd[i]);
  return WANN_OK;
}

// CRC-32C (Castagnoli), the checksum of TF-V2 checkpoint bundles (model.index entries).
uint32_t wann_crc32c(const void* data, uint64_t bytes, uint32_t seed) {
  static uint32_t table[8][256];
  static std::once_flag once;
  std::call_once(once, [] {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      table[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) table[t][i] = (table[t - 1][i] >> 8) ^ table[0][table[t - 1][i] & 0xFF];
  });
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = ~seed;
  while (bytes >= 8) {   // slicing-by-8
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = table[7][lo & 0xFF] ^ table[6][(lo >> 8) & 0xFF] ^ table[5][(lo >> 16) & 0xFF] ^ table[4][lo >> 24] ^
        table[3][hi & 0xFF] ^ table[2][(hi >> 8) & 0xFF] ^ table[1][(hi >> 16) & 0xFF] ^ table[0][hi >> 24];
    p += 8;
    bytes -= 8;
  }
  while (bytes--) c = table[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return ~c;
}

int wann_host_alloc(void** ptr, uint64_t bytes) {
  if (ptr == nullptr) return fail(WANN_E_INVALID, "null ptr");
  CU(cudaMallocHost(ptr, bytes));
  return WANN_OK;
}

int wann_host_free(void* ptr) {
  CU(cudaFreeHost(ptr));
  return WANN_OK;
}

int wann_get_stats(wann_ctx* ctx, uint64_t stats[8]) {
  if (ctx == nullptr || stats == nullptr) return fail(WANN_E_INVALID, "null argument");
  for (int i = 0; i < 8; ++i) stats[i] = ctx->stats[i].load();
  return WANN_OK;
}

int wann_set_option(wann_ctx* ctx, const char* key, int64_t value) {
  if (ctx == nullptr || key == nullptr) return fail(WANN_E_INVALID, "null argument");
  if (strcmp(key, "variant") == 0) { ctx->variant = static_cast<int>(value); return WANN_OK; }
  if (strcmp(key, "skew") == 0) {
    if (value < 0 || value > 6) return fail(WANN_E_INVALID, "skew out of range");
    ctx->skew = static_cast<int>(value);
    return WANN_OK;
  }
  if (strcmp(key, "small_batch_halves") == 0) { ctx->small_batch_halves = value != 0; return WANN_OK; }
  if (strcmp(key, "pdl") == 0) { ctx->pdl = value < 0 ? -1 : (value != 0); return WANN_OK; }
  if (strcmp(key, "zero_copy") == 0) { ctx->zero_copy = value != 0; return WANN_OK; }
  if (strcmp(key, "fuse_tail") == 0) { ctx->fuse_tail = value != 0; return WANN_OK; }
  if (strcmp(key, "coalesce") == 0) { ctx->coalesce = value != 0; return WANN_OK; }
  if (strcmp(key, "refine_margin_micro") == 0) {
    // margin in 1e-6 logit units; 0 switches the refinement pass off
    if (value < 0 || value > 20000000) return fail(WANN_E_INVALID, "refine margin out of range");
    if (value > 0 && !is_tc16(ctx->mode))
      return fail(WANN_E_INVALID, "margin refinement applies to the bf16/fp16 tensor-core modes");
    ctx->refine_ratio = value == 0 ? 0.f : expf(-static_cast<float>(value) * 1e-6f);
    return WANN_OK;
  }
  if (strcmp(key, "time_kernels") == 0) { ctx->time_kernels = value != 0; return WANN_OK; }
  if (strcmp(key, "gptq") == 0) {
    // centred mode: data-aware rounding of the 192->192 layers' weights (gptq.hpp) on / off; re-calibrates
    if (ctx->mode != WANN_MODE_FP16C) return fail(WANN_E_INVALID, "gptq applies to the fp16c mode");
    if ((value != 0) == ctx->gptq.load()) return WANN_OK;
    DeviceGuard dev_guard(ctx->device);
    CU(dev_guard.status);
    ctx->weights_gate.writer_enter();
    int rc = WANN_OK;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = fail(WANN_E_CUDA, "device synchronise: %s", cudaGetErrorString(e));
    else {
      ctx->gptq = value != 0;
      rc = upload_weights(ctx, host_weights_view(ctx));
    }
    ctx->weights_gate.writer_exit();
    return rc;
  }
  if (strcmp(key, "split_layers") == 0 || strcmp(key, "lo_n") == 0 || strcmp(key, "lo_kgroups") == 0) {
    // centred mode, partial hi/lo weight split: in the conv layers of "split_layers" (bit l = conv(l+1), l = 1..3)
    // every K-block (tap, j < "lo_kgroups") streams a lo stage w - fp16(w) for the output positions [0, "lo_n")
    // and runs a second, N = lo_n MMA per K-step against it: the weight rounding error of that block of the layer
    // disappears.  Positions are sorted by activation variance, so the leading block carries most of the error
    // (lo_n = 192, lo_kgroups = 3: the whole layer).  Rebuilds the stage image.
    if (ctx->mode != WANN_MODE_FP16C) return fail(WANN_E_INVALID, "%s applies to the fp16c mode", key);
    if (key[0] == 's' && (value < 0 || (value & ~14ll))) return fail(WANN_E_INVALID, "split_layers is a mask of bits 1..3");
    if (strcmp(key, "lo_n") == 0 && (value < 16 || value > 192 || value % 16))
      return fail(WANN_E_INVALID, "lo_n is a multiple of 16 in 16..192");
    if (strcmp(key, "lo_kgroups") == 0 && (value < 1 || value > 3)) return fail(WANN_E_INVALID, "lo_kgroups is 1..3");
    DeviceGuard dev_guard(ctx->device);
    CU(dev_guard.status);
    ctx->weights_gate.writer_enter();
    int rc = WANN_OK;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = fail(WANN_E_CUDA, "device synchronise: %s", cudaGetErrorString(e));
    else {
      if (key[0] == 's') ctx->split_mask = static_cast<int>(value);
      else if (strcmp(key, "lo_n") == 0) ctx->lo_n = static_cast<int>(value);
      else ctx->lo_kgroups = static_cast<int>(value);
      const Centering cen = ctx->cen;
      rc = upload_tc_image(ctx, host_weights_view(ctx), cen);
    }
    ctx->weights_gate.writer_exit();
    return rc;
  }
  if (strcmp(key, "chunk") == 0) {
    if (value < 8 || value > (1 << 20)) return fail(WANN_E_INVALID, "chunk out of range");
    std::lock_guard<std::mutex> g(ctx->mu);
    for (auto& l : ctx->lanes)
      if (l.in_use || l.slot[0].stream != nullptr)
        return fail(WANN_E_INVALID, "chunk must be set before the first evaluation");
    ctx->chunk = value;
    return WANN_OK;
  }
  return fail(WANN_E_INVALID, "unknown option %s", key);
}

// ---------------------------------------------------------------------------------------------
// Forest of array-based search trees (array_tree.hpp): host-side, no context needed.
}  // extern "C"

// Device-resident forest: the trees (headers + node / board arenas) live in HBM; a lock step is
// forest_next_kernel -> conv stack (+ fused tail) -> seed_children_kernel -> forest_apply_kernel on one
// stream, the number of requested rows staying on the device.  One warp per tree, lane 0 walks it.
struct DeviceForest {
  wann_ctx* ctx = nullptr;
  int device = 0;                               // (kept here: the forest may be destroyed after its context)
  int64_t n = 0;
  int32_t cap_nodes = 0, cap_boards = 0;
  wann_tree::Tree* d_trees = nullptr;
  wann_tree::Tree::Node* d_nodes = nullptr;     // [2][n][cap_nodes]: arena + the one advance_root compacts into
  int8_t* d_boards = nullptr;                   // [2][n][cap_boards][64]
  int8_t* d_sq = nullptr; uint8_t* d_bl = nullptr; int32_t* d_ids = nullptr;
  uint8_t* d_idx = nullptr; float* d_prior = nullptr; uint8_t* d_nlegal = nullptr; uint8_t* d_cflags = nullptr;
  int32_t* d_cstats = nullptr; int32_t* d_psum = nullptr;
  int* d_counts = nullptr;                      // [2][kForestHist] rows requested by each of the last steps; trees that yielded
  int* h_counts = nullptr;                      // pinned copy
  int8_t* d_restart = nullptr;                  // [64] + colour: wann_forest_play's new-game position
  uint8_t* d_moves = nullptr; uint8_t* d_fin = nullptr; int* d_flag = nullptr;
  unsigned long long* d_cycles = nullptr;       // [kForestHist][4]: per lock step, clock64 cycles of the trees in the tree
                                                // kernels: next sum, next max, apply sum, apply max
  cudaStream_t stream = nullptr;
  int64_t steps = 0;
};
constexpr int kForestHist = 64, kForestBurst = 8;

struct wann_forest {
  wann_tree::Forest f;
  std::vector<int64_t> waiting;      // trees whose rows were handed out by the last wann_forest_next
  DeviceForest* dev = nullptr;       // non-null: the trees live on the GPU (wann_forest_create_device)
};

namespace {

__global__ void forest_init_kernel(wann_tree::Tree* trees, long long n, const int8_t* __restrict__ squares,
                                   const uint8_t* __restrict__ black, const int32_t* __restrict__ heights) {
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (t >= n) return;
  wann_tree::Tree tr = trees[t];
  tr.init_in_place(squares + t * 64, black[t] != 0, heights != nullptr ? heights[t] : 0);
  trees[t] = tr;
}

// every tree runs to its next evaluation (or to the end of its simulation budget); the requested
// positions are appended to squares_out / black_out / ids (order = arrival: the evaluation of a row does
// not depend on its place in the batch)
__global__ void __launch_bounds__(128) forest_next_kernel(wann_tree::Tree* trees, long long n, long long max_sims,
                                                          int8_t* __restrict__ squares_out, uint8_t* __restrict__ black_out,
                                                          int32_t* __restrict__ ids, int* count, unsigned long long* cycles) {
  // one WARP per tree: all lanes run the tree code on identical state (array_tree.hpp, "DEVICE execution model")
  const long long t = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  if (t >= n) return;
  const int lane = threadIdx.x & 31;
  const long long c0 = clock64();
  wann_tree::Tree tr = trees[t];
  int32_t node = tr.advance(max_sims, 1);
  if (node == -2 && lane == 0) atomicAdd(count + kForestHist, 1);      // more to do, nothing to evaluate this step
  const int8_t* b = node >= 0 ? tr.board_of(node) : nullptr;
  if (node >= 0 && b == nullptr) {             // no room for the position: the tree stops (overflow is set)
    tr.phase = wann_tree::Tree::kFinished;
    tr.pending = -1;
    node = -1;
  }
  __syncwarp();
  if (node >= 0) {
    int slot = 0;
    if (lane == 0) slot = atomicAdd(count, 1);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (lane < 4) reinterpret_cast<int4*>(squares_out + static_cast<long long>(slot) * 64)[lane] = reinterpret_cast<const int4*>(b)[lane];
    if (lane == 0) {
      black_out[slot] = tr.nodes[node].black;
      ids[slot] = static_cast<int32_t>(t);
    }
  }
  if (lane == 0) {
    trees[t] = tr;
    const unsigned long long dc = static_cast<unsigned long long>(clock64() - c0);
    atomicAdd(cycles, dc);
    atomicMax(cycles + 1, dc);
  }
}

__global__ void __launch_bounds__(128) forest_apply_kernel(wann_tree::Tree* trees, const int32_t* __restrict__ ids,
                                                           const int* __restrict__ count, const uint8_t* __restrict__ idx,
                                                           const float* __restrict__ prior, const uint8_t* __restrict__ n_legal,
                                                           const uint8_t* __restrict__ cflags, const int32_t* __restrict__ cstats,
                                                           const int32_t* __restrict__ psum, unsigned long long* cycles, int* count_next,
                                                           unsigned long long* cycles_next) {
  const long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  if (blockIdx.x == 0 && threadIdx.x < 5) {      // the next step's row counter and statistics start from zero
    if (threadIdx.x == 4) { count_next[0] = 0; count_next[kForestHist] = 0; } else cycles_next[threadIdx.x] = 0ull;
  }
  if (i >= *count) return;
  const long long c0 = clock64();
  const int32_t t = ids[i];
  wann_tree::Tree tr = trees[t];
  tr.expand_resume(idx + i * kMaxLegal, prior + i * kMaxLegal, min(static_cast<int>(n_legal[i]), kMaxLegal),
                   cflags + i * kMaxLegal, cstats + i * kMaxLegal * 4, psum[i * 8 + 5] != 0);
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    trees[t] = tr;
    const unsigned long long dc = static_cast<unsigned long long>(clock64() - c0);
    atomicAdd(cycles + 2, dc);
    atomicMax(cycles + 3, dc);
  }
}

// a tree that is mid-expansion cannot move: the caller checks ALL trees before any of them plays
__global__ void forest_waiting_kernel(const wann_tree::Tree* trees, long long n, int* flag) {
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (t < n && trees[t].phase == wann_tree::Tree::kWaitEval) atomicExch(flag, static_cast<int>(t) + 1);
}

__global__ void __launch_bounds__(128) forest_play_kernel(wann_tree::Tree* trees, long long n, const int8_t* __restrict__ restart,
                                                          uint8_t* __restrict__ moves, uint8_t* __restrict__ fin, int* flag) {
  const long long t = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  if (t >= n) return;
  wann_tree::Tree tr = trees[t];
  if (tr.phase == wann_tree::Tree::kWaitEval) { atomicExch(flag, static_cast<int>(t) + 1); return; }
  uint8_t mv, f;
  tr.play(restart, restart[64] != 0, mv, f);
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    moves[t] = mv;
    fin[t] = f;
    trees[t] = tr;
  }
}

void free_device_forest(DeviceForest* d) {
  if (d == nullptr) return;
  if (d->stream) { cudaStreamSynchronize(d->stream); cudaStreamDestroy(d->stream); }
  cudaFree(d->d_trees); cudaFree(d->d_nodes); cudaFree(d->d_boards); cudaFree(d->d_sq); cudaFree(d->d_bl);
  cudaFree(d->d_ids); cudaFree(d->d_idx); cudaFree(d->d_prior); cudaFree(d->d_nlegal); cudaFree(d->d_cflags);
  cudaFree(d->d_cstats); cudaFree(d->d_psum); cudaFree(d->d_counts); cudaFree(d->d_restart); cudaFree(d->d_moves);
  cudaFree(d->d_fin); cudaFree(d->d_flag); cudaFree(d->d_cycles);
  if (d->h_counts) cudaFreeHost(d->h_counts);
  delete d;
}

// host copy of one device tree (header + the used part of its arenas), owning its buffers
int download_tree(DeviceForest* d, int64_t t, wann_tree::Tree* out) {
  CU(cudaStreamSynchronize(d->stream));
  wann_tree::Tree h;
  CU(cudaMemcpy(&h, d->d_trees + t, sizeof h, cudaMemcpyDeviceToHost));
  wann_tree::Tree r = h;
  r.owns = 1;
  r.alt_nodes = nullptr;
  r.alt_boards = nullptr;
  r.cap_nodes = std::max(h.n_nodes, 1);
  r.cap_boards = std::max(h.n_boards, 1);
  r.nodes = static_cast<wann_tree::Tree::Node*>(malloc(sizeof(wann_tree::Tree::Node) * static_cast<size_t>(r.cap_nodes)));
  r.boards = static_cast<int8_t*>(malloc(static_cast<size_t>(r.cap_boards) * 64));
  if (r.nodes == nullptr || r.boards == nullptr) { free(r.nodes); free(r.boards); return fail(WANN_E_INVALID, "out of host memory"); }
  cudaError_t e = cudaMemcpy(r.nodes, h.nodes, sizeof(wann_tree::Tree::Node) * static_cast<size_t>(h.n_nodes), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(r.boards, h.boards, static_cast<size_t>(h.n_boards) * 64, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { free(r.nodes); free(r.boards); return fail(WANN_E_CUDA, "tree download: %s", cudaGetErrorString(e)); }
  *out = r;
  return WANN_OK;
}

}  // namespace

extern "C" {

int wann_host_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
}

int wann_forest_create(int64_t n, const int8_t* squares, const uint8_t* black, const int32_t* heights,
                       wann_forest** out) {
  if (out == nullptr || n < 0 || (n > 0 && (squares == nullptr || black == nullptr)))
    return fail(WANN_E_INVALID, "bad argument");
  wann_forest* fo = new wann_forest();
  fo->f.trees.resize(static_cast<size_t>(n));
  for (int64_t t = 0; t < n; ++t) fo->f.trees[t].init(squares + t * 64, black[t] != 0, heights ? heights[t] : 0);
  *out = fo;
  return WANN_OK;
}

int wann_forest_destroy(wann_forest* fo) {
  if (fo != nullptr && fo->dev != nullptr) {
    DeviceGuard dev_guard(fo->dev->device);
    free_device_forest(fo->dev);
  }
  delete fo;
  return WANN_OK;
}

int wann_forest_create_device(wann_ctx* ctx, int64_t n, const int8_t* squares, const uint8_t* black,
                              const int32_t* heights, int32_t node_capacity, wann_forest** out) {
  if (ctx == nullptr || out == nullptr || n <= 0 || squares == nullptr || black == nullptr)
    return fail(WANN_E_INVALID, "bad argument");
  if (node_capacity <= 0) node_capacity = 16384;
  if (node_capacity < 256) return fail(WANN_E_INVALID, "node_capacity must be at least 256");
  DeviceGuard dev_guard(ctx->device);
  CU(dev_guard.status);
  wann_forest* fo = new wann_forest();
  DeviceForest* d = new DeviceForest();
  fo->dev = d;
  d->ctx = ctx;
  d->device = ctx->device;
  d->n = n;
  d->cap_nodes = node_capacity;
  d->cap_boards = std::max(64, node_capacity / 8);      // a node gets a board when it is evaluated: 1 in ~23
  auto body = [&]() -> int {
    using Node = wann_tree::Tree::Node;
    const size_t nn = static_cast<size_t>(n);
    CU(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    CU(cudaMalloc(&d->d_trees, nn * sizeof(wann_tree::Tree)));
    CU(cudaMalloc(&d->d_nodes, 2 * nn * d->cap_nodes * sizeof(Node)));
    CU(cudaMalloc(&d->d_boards, 2 * nn * d->cap_boards * 64));
    CU(cudaMalloc(&d->d_sq, nn * 64));
    CU(cudaMalloc(&d->d_bl, nn));
    CU(cudaMalloc(&d->d_ids, nn * sizeof(int32_t)));
    CU(cudaMalloc(&d->d_idx, nn * kMaxLegal));
    CU(cudaMalloc(&d->d_prior, nn * kMaxLegal * 4));
    CU(cudaMalloc(&d->d_nlegal, nn));
    CU(cudaMalloc(&d->d_cflags, nn * kMaxLegal));
    CU(cudaMalloc(&d->d_cstats, nn * kMaxLegal * 4 * sizeof(int32_t)));
    CU(cudaMalloc(&d->d_psum, nn * 8 * sizeof(int32_t)));
    CU(cudaMalloc(&d->d_counts, 2 * kForestHist * sizeof(int)));
    CU(cudaMallocHost(&d->h_counts, 2 * kForestHist * sizeof(int)));
    CU(cudaMalloc(&d->d_restart, 80));
    CU(cudaMalloc(&d->d_moves, nn));
    CU(cudaMalloc(&d->d_fin, nn));
    CU(cudaMalloc(&d->d_flag, sizeof(int)));
    CU(cudaMalloc(&d->d_cycles, kForestHist * 4 * sizeof(unsigned long long)));
    std::vector<wann_tree::Tree> hdr(nn);
    for (size_t t = 0; t < nn; ++t) {
      wann_tree::Tree& tr = hdr[t];
      tr.nodes = d->d_nodes + t * d->cap_nodes;
      tr.alt_nodes = d->d_nodes + (nn + t) * d->cap_nodes;
      tr.boards = d->d_boards + t * d->cap_boards * 64;
      tr.alt_boards = d->d_boards + (nn + t) * d->cap_boards * 64;
      tr.cap_nodes = d->cap_nodes;
      tr.cap_boards = d->cap_boards;
      tr.owns = 0;
    }
    CU(cudaMemcpyAsync(d->d_trees, hdr.data(), nn * sizeof(wann_tree::Tree), cudaMemcpyHostToDevice, d->stream));
    int32_t* d_h = nullptr;
    // the root positions go through the step buffers (d_sq / d_bl hold n rows)
    CU(cudaMemcpyAsync(d->d_sq, squares, nn * 64, cudaMemcpyHostToDevice, d->stream));
    CU(cudaMemcpyAsync(d->d_bl, black, nn, cudaMemcpyHostToDevice, d->stream));
    if (heights != nullptr) {
      CU(cudaMalloc(&d_h, nn * sizeof(int32_t)));
      CU(cudaMemcpyAsync(d_h, heights, nn * sizeof(int32_t), cudaMemcpyHostToDevice, d->stream));
    }
    forest_init_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, d->stream>>>(d->d_trees, n, d->d_sq, d->d_bl, d_h);
    cudaError_t e = cudaStreamSynchronize(d->stream);
    cudaFree(d_h);
    if (e != cudaSuccess) return fail(WANN_E_CUDA, "forest init: %s", cudaGetErrorString(e));
    return WANN_OK;
  };
  int rc = body();
  if (rc != WANN_OK) {
    free_device_forest(d);
    delete fo;
    return rc;
  }
  *out = fo;
  return WANN_OK;
}

// Search: every tree does `max_simulations` simulations (run_MCTS_with_expansions_simulation,
// expansion_MCTS_functions.pyx:300-350); out[0] = lock steps, out[1] = evaluations.
int wann_forest_run(wann_ctx* ctx, wann_forest* fo, int64_t max_simulations, int64_t out[2]) {
  return wann_forest_run_steps(ctx, fo, max_simulations, INT64_MAX, out);
}

// ... with a budget of lock steps as well (the reference's search is time-budgeted, `time_to_think`,
// expansion_MCTS_functions.pyx:113-119): stops after max_lock_steps steps even if trees are mid-simulation;
// a following wann_forest_run(ctx, forest, 0, ...) finishes the simulations in flight without starting new ones.
int wann_forest_run_steps(wann_ctx* ctx, wann_forest* fo, int64_t max_simulations, int64_t max_lock_steps, int64_t out[2]) {
  if (ctx == nullptr || fo == nullptr) return fail(WANN_E_INVALID, "null argument");
  if (max_lock_steps <= 0) { if (out) out[0] = out[1] = 0; return WANN_OK; }
  if (!fo->waiting.empty()) return fail(WANN_E_INVALID, "wann_forest_apply must follow wann_forest_next");
  int64_t steps = 0, evals = 0;
  if (fo->dev == nullptr) {
    // host forest: next -> ONE wann_eval_expand_seeded over the requested rows -> apply, until no tree asks
    const int64_t n = static_cast<int64_t>(fo->f.trees.size());
    if (n == 0) { if (out) out[0] = out[1] = 0; return WANN_OK; }
    int8_t* sq = nullptr; uint8_t* bl = nullptr; uint8_t* idx = nullptr; float* prior = nullptr; uint8_t* nl = nullptr;
    uint8_t* fl = nullptr; int32_t* st = nullptr; int32_t* ps = nullptr;
    DeviceGuard dev_guard(ctx->device);
    CU(dev_guard.status);
    auto body = [&]() -> int {
      CU(cudaMallocHost(&sq, n * 64)); CU(cudaMallocHost(&bl, n)); CU(cudaMallocHost(&idx, n * kMaxLegal));
      CU(cudaMallocHost(&prior, n * kMaxLegal * 4)); CU(cudaMallocHost(&nl, n)); CU(cudaMallocHost(&fl, n * kMaxLegal));
      CU(cudaMallocHost(&st, n * kMaxLegal * 16)); CU(cudaMallocHost(&ps, n * 32));
      while (steps < max_lock_steps) {
        int64_t m = 0;
        int rc = wann_forest_next(fo, max_simulations, sq, bl, nullptr, &m);
        if (rc != WANN_OK) return rc;
        if (m == 0) break;
        rc = wann_eval_expand_seeded(ctx, sq, bl, m, 0, idx, prior, nl, nullptr, fl, st, ps, WANN_MEM_HOST, nullptr);
        if (rc == WANN_OK) rc = wann_forest_apply(fo, idx, prior, nl, fl, st, ps);
        if (rc != WANN_OK) { fo->waiting.clear(); return rc; }
        ++steps;
        evals += m;
      }
      return WANN_OK;
    };
    int rc = body();
    cudaFreeHost(sq); cudaFreeHost(bl); cudaFreeHost(idx); cudaFreeHost(prior); cudaFreeHost(nl); cudaFreeHost(fl);
    cudaFreeHost(st); cudaFreeHost(ps);
    if (out) { out[0] = steps; out[1] = evals; }
    return rc;
  }
  DeviceForest* d = fo->dev;
  if (d->ctx != ctx) return fail(WANN_E_INVALID, "the forest was created on another context");
  DeviceGuard dev_guard(ctx->device);
  CU(dev_guard.status);
  ReaderScope weights_scope(&ctx->weights_gate);
  Lane* lane = nullptr;
  int rc = acquire_lane(ctx, &lane);
  if (rc != WANN_OK) return rc;
  lane_wait(lane, d->stream);
  auto body = [&]() -> int {
    const long long n = d->n;
    const unsigned blocks = static_cast<unsigned>((n * 32 + 127) / 128);
    const long long cap = lane->slot[0].cap * (ctx->mode == WANN_MODE_FP32 ? 1 : kDevChunkMul);
    Slot& ws = lane->slot[0];
    // development aid: WANN_FOREST_TIMING=1 brackets the phases of the first 48 lock steps with CUDA events
    const bool timing = getenv("WANN_FOREST_TIMING") != nullptr;
    std::vector<cudaEvent_t> evs;
    auto mark = [&]() {
      if (timing && evs.size() < 48 * 4 + 1) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, d->stream);
        evs.push_back(e);
      }
    };
    // the row counter and the cycle statistics of a step are zeroed by the apply kernel of the step before
    // (nothing touches them in between); the very first step of a run is zeroed here
    CU(cudaMemsetAsync(d->d_counts + (d->steps % kForestHist), 0, sizeof(int), d->stream));
    CU(cudaMemsetAsync(d->d_counts + kForestHist + (d->steps % kForestHist), 0, sizeof(int), d->stream));
    CU(cudaMemsetAsync(d->d_cycles + (d->steps % kForestHist) * 4, 0, 4 * sizeof(unsigned long long), d->stream));
    mark();
    double host_enqueue_us = 0, host_wait_us = 0;
    long long bursts = 0;
    int64_t launched = 0;
    while (true) {
      const int first = static_cast<int>(d->steps % kForestHist);
      const double tq0 = omp_get_wtime();
      const int burst = static_cast<int>(std::min<int64_t>(kForestBurst, max_lock_steps - launched));
      launched += burst;
      for (int b = 0; b < burst; ++b, ++d->steps) {
        int* cnt = d->d_counts + (d->steps % kForestHist);
        unsigned long long* cyc = d->d_cycles + (d->steps % kForestHist) * 4;
        int* cnt_next = d->d_counts + ((d->steps + 1) % kForestHist);
        unsigned long long* cyc_next = d->d_cycles + ((d->steps + 1) % kForestHist) * 4;
        forest_next_kernel<<<blocks, 128, 0, d->stream>>>(d->d_trees, n, max_simulations, d->d_sq, d->d_bl, d->d_ids, cnt, cyc);
        mark();
        for (long long off = 0; off < n; off += cap) {
          const long long m = std::min(cap, n - off);
          ChunkIO io;
          io.squares = d->d_sq + off * 64; io.black = d->d_bl + off;
          io.idx = d->d_idx + off * kMaxLegal; io.prior = d->d_prior + off * kMaxLegal; io.nlegal = d->d_nlegal + off;
          io.cflags = d->d_cflags + off * kMaxLegal; io.cstats = d->d_cstats + off * kMaxLegal * 4; io.psum = d->d_psum + off * 8;
          io.n_dev = cnt; io.n_dev_off = off;
          int r = enqueue_eval(ctx, ws, d->stream, io, m);
          if (r != WANN_OK) return r;
        }
        mark();
        forest_apply_kernel<<<blocks, 128, 0, d->stream>>>(d->d_trees, d->d_ids, cnt, d->d_idx, d->d_prior, d->d_nlegal,
                                                          d->d_cflags, d->d_cstats, d->d_psum, cyc, cnt_next, cyc_next);
        mark();
        mark();
        ctx->stats[1] += 2;
      }
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(d->h_counts, d->d_counts, 2 * kForestHist * sizeof(int), cudaMemcpyDeviceToHost, d->stream));
      CU(copy_device_err(ctx, ws, d->stream));
      const double tq1 = omp_get_wtime();
      CU(cudaStreamSynchronize(d->stream));
      host_enqueue_us += (tq1 - tq0) * 1e6;
      host_wait_us += (omp_get_wtime() - tq1) * 1e6;
      ++bursts;
      int r = check_device_err(ctx, ws);
      if (r != WANN_OK) return r;
      bool done = launched >= max_lock_steps;
      for (int b = 0; b < burst; ++b) {
        const int c = d->h_counts[(first + b) % kForestHist], busy = d->h_counts[kForestHist + (first + b) % kForestHist];
        if (c > 0) { ++steps; evals += c; }
        else if (busy == 0) done = true;                           // no tree asked or yielded: none will again
      }
      if (done) break;
    }
    if (timing && evs.size() > 8) {
      double ph[4] = {0, 0, 0, 0};
      const size_t nst = (evs.size() - 1) / 4;
      for (size_t i = 0; i < nst; ++i)
        for (int k = 0; k < 4; ++k) {
          float ms = 0;
          cudaEventElapsedTime(&ms, evs[4 * i + k], evs[4 * i + k + 1]);
          ph[k] += ms * 1e3 / nst;
        }
      fprintf(stderr, "[wann_forest_run] %zu steps, us per step: next %.1f  eval(+seed) %.1f  apply %.1f  (gap %.1f); host per "
              "burst of %d steps: enqueue %.1f us, wait %.1f us\n", nst, ph[0], ph[1], ph[2], ph[3], kForestBurst,
              host_enqueue_us / bursts, host_wait_us / bursts);
    }
    for (cudaEvent_t e : evs) cudaEventDestroy(e);
    return WANN_OK;
  };
  rc = body();
  if (rc != WANN_OK) cudaStreamSynchronize(d->stream);
  lane->last_valid = false;
  release_lane(ctx, lane);
  ctx->stats[0] += static_cast<uint64_t>(evals);
  if (out) { out[0] = steps; out[1] = evals; }
  return rc;
}

int wann_forest_next(wann_forest* fo, int64_t max_simulations, int8_t* squares_out, uint8_t* black_out,
                     int64_t* tree_ids, int64_t* m_out) {
  if (fo == nullptr || squares_out == nullptr || black_out == nullptr || m_out == nullptr)
    return fail(WANN_E_INVALID, "null argument");
  if (!fo->waiting.empty()) return fail(WANN_E_INVALID, "wann_forest_apply must follow wann_forest_next");
  if (fo->dev != nullptr) return fail(WANN_E_INVALID, "a device-resident forest is driven by wann_forest_run");
  const int64_t n = static_cast<int64_t>(fo->f.trees.size());
  std::vector<int32_t> node(static_cast<size_t>(n), -1);
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t t = 0; t < n; ++t) {
    wann_tree::Tree& tr = fo->f.trees[t];
    node[t] = tr.advance(max_simulations);
    if (node[t] >= 0) tr.board_of(node[t]);          // materialise the position while we are parallel
  }
  int64_t m = 0;
  for (int64_t t = 0; t < n; ++t)
    if (node[t] >= 0) fo->waiting.push_back(t);
  m = static_cast<int64_t>(fo->waiting.size());
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < m; ++i) {
    const int64_t t = fo->waiting[i];
    wann_tree::Tree& tr = fo->f.trees[t];
    memcpy(squares_out + i * 64, tr.board_of(node[t]), 64);
    black_out[i] = tr.nodes[node[t]].black;
    if (tree_ids != nullptr) tree_ids[i] = t;
  }
  *m_out = m;
  return WANN_OK;
}

int wann_forest_apply(wann_forest* fo, const uint8_t* idx, const float* prior, const uint8_t* n_legal,
                      const uint8_t* child_flags, const int32_t* child_stats, const int32_t* parent_summary) {
  if (fo == nullptr) return fail(WANN_E_INVALID, "null forest");
  if (fo->dev != nullptr) return fail(WANN_E_INVALID, "a device-resident forest is driven by wann_forest_run");
  const int64_t m = static_cast<int64_t>(fo->waiting.size());
  if (m > 0 && (!idx || !prior || !n_legal || !child_flags || !child_stats || !parent_summary))
    return fail(WANN_E_INVALID, "null argument");
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t i = 0; i < m; ++i) {
    wann_tree::Tree& tr = fo->f.trees[fo->waiting[i]];
    tr.expand_resume(idx + i * kMaxLegal, prior + i * kMaxLegal, std::min<int>(n_legal[i], kMaxLegal),
                     child_flags + i * kMaxLegal, child_stats + i * kMaxLegal * 4, parent_summary[i * 8 + 5] != 0);
  }
  fo->waiting.clear();
  return WANN_OK;
}

int wann_forest_info(wann_forest* fo, int64_t tree, int64_t out[8]) {
  if (fo == nullptr || out == nullptr || tree < 0 ||
      tree >= (fo->dev ? fo->dev->n : static_cast<int64_t>(fo->f.trees.size())))
    return fail(WANN_E_INVALID, "bad argument");
  wann_tree::Tree copy;
  if (fo->dev != nullptr) {                      // a host copy of the device tree answers the query
    DeviceGuard dev_guard(fo->dev->device);
    int rc = download_tree(fo->dev, tree, &copy);
    if (rc != WANN_OK) return rc;
  }
  struct Release { wann_tree::Tree* t; ~Release() { if (t->owns) t->release(); } } release_copy{&copy};
  const wann_tree::Tree& tr = fo->dev ? copy : fo->f.trees[tree];
  out[0] = static_cast<int64_t>(tr.n_nodes);
  out[1] = tr.simulations;
  out[2] = tr.evaluations;
  out[3] = tr.nodes[tr.root].visits;
  out[4] = tr.nodes[tr.root].wins;
  out[5] = tr.nodes[tr.root].ws;
  out[6] = tr.phase == wann_tree::Tree::kFinished ? (tr.overflow ? 2 : 1) : 0;   // 2: stopped because its arena is full
  const int32_t mv = tr.best_move();       // randomly_choose_a_winning_move (tree_search_utils.pyx:469-486)
  out[7] = mv >= 0 ? tr.nodes[mv].index : -1;
  return WANN_OK;
}

int wann_forest_advance(wann_forest* fo, int64_t tree, int move_index) {
  if (fo == nullptr || tree < 0 || tree >= static_cast<int64_t>(fo->f.trees.size()))
    return fail(WANN_E_INVALID, fo != nullptr && fo->dev != nullptr ? "device-resident forests move with wann_forest_play" : "bad argument");
  if (!fo->waiting.empty()) return fail(WANN_E_INVALID, "wann_forest_apply must follow wann_forest_next");
  if (!fo->f.trees[tree].advance_root(move_index))
    return fail(WANN_E_INVALID, "tree %lld: the root has no child for move index %d (or is mid-simulation)",
                static_cast<long long>(tree), move_index);
  return WANN_OK;
}

int wann_forest_play(wann_forest* fo, const int8_t* restart_squares, int restart_black, uint8_t* moves_out,
                     uint8_t* finished_out) {
  if (fo == nullptr || restart_squares == nullptr) return fail(WANN_E_INVALID, "null argument");
  if (!fo->waiting.empty()) return fail(WANN_E_INVALID, "wann_forest_apply must follow wann_forest_next");
  if (fo->dev != nullptr) {
    DeviceForest* d = fo->dev;
    DeviceGuard dev_guard(d->device);
    CU(dev_guard.status);
    int8_t rs[80] = {0};
    memcpy(rs, restart_squares, 64);
    rs[64] = restart_black != 0;
    int flag = 0;
    CU(cudaMemcpyAsync(d->d_restart, rs, 80, cudaMemcpyHostToDevice, d->stream));
    CU(cudaMemsetAsync(d->d_flag, 0, sizeof(int), d->stream));
    forest_waiting_kernel<<<static_cast<unsigned>((d->n + 255) / 256), 256, 0, d->stream>>>(d->d_trees, d->n, d->d_flag);
    CU(cudaMemcpyAsync(&flag, d->d_flag, sizeof(int), cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    if (flag != 0)
      return fail(WANN_E_INVALID, "tree %d is waiting for an evaluation: finish wann_forest_run before wann_forest_play", flag - 1);
    forest_play_kernel<<<static_cast<unsigned>((d->n * 32 + 127) / 128), 128, 0, d->stream>>>(
        d->d_trees, d->n, d->d_restart, d->d_moves, d->d_fin, d->d_flag);
    CU(cudaGetLastError());
    if (moves_out != nullptr) CU(cudaMemcpyAsync(moves_out, d->d_moves, d->n, cudaMemcpyDeviceToHost, d->stream));
    if (finished_out != nullptr) CU(cudaMemcpyAsync(finished_out, d->d_fin, d->n, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaMemcpyAsync(&flag, d->d_flag, sizeof(int), cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    if (flag != 0)
      return fail(WANN_E_INVALID, "tree %d is waiting for an evaluation: finish wann_forest_run before wann_forest_play", flag - 1);
    return WANN_OK;
  }
  const int64_t n = static_cast<int64_t>(fo->f.trees.size());
  for (int64_t t = 0; t < n; ++t)      // a tree that is mid-expansion cannot move: finish the next / apply cycle first
    if (fo->f.trees[t].phase == wann_tree::Tree::kWaitEval)
      return fail(WANN_E_INVALID, "tree %lld is waiting for an evaluation: run wann_forest_next / _apply until "
                  "wann_forest_next returns no rows before wann_forest_play", static_cast<long long>(t));
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t t = 0; t < n; ++t) {
    uint8_t mv = 255, fin = 0;
    fo->f.trees[t].play(restart_squares, restart_black != 0, mv, fin);
    if (moves_out != nullptr) moves_out[t] = mv;
    if (finished_out != nullptr) finished_out[t] = fin;
  }
  return WANN_OK;
}

int wann_forest_debug_cycles(wann_forest* fo, int64_t* out) {
  if (fo == nullptr || fo->dev == nullptr || out == nullptr) return fail(WANN_E_INVALID, "needs a device-resident forest");
  DeviceGuard dev_guard(fo->dev->device);
  CU(cudaStreamSynchronize(fo->dev->stream));
  // rows: the last kForestHist lock steps, oldest first: rows requested, next sum, next max, apply sum, apply max
  DeviceForest* d = fo->dev;
  std::vector<unsigned long long> cyc(kForestHist * 4);
  std::vector<int> cnt(kForestHist);
  CU(cudaMemcpy(cyc.data(), d->d_cycles, cyc.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(cnt.data(), d->d_counts, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost));
  for (int i = 0; i < kForestHist; ++i) {
    const int slot = static_cast<int>((d->steps + i) % kForestHist);
    out[i * 5] = cnt[slot];
    for (int k = 0; k < 4; ++k) out[i * 5 + 1 + k] = static_cast<int64_t>(cyc[slot * 4 + k]);
  }
  return WANN_OK;
}

int64_t wann_forest_dump(wann_forest* fo, int64_t tree, int64_t* rows, int64_t capacity) {
  if (fo == nullptr || tree < 0 || tree >= (fo->dev ? fo->dev->n : static_cast<int64_t>(fo->f.trees.size()))) return -1;
  if (fo->dev != nullptr) {
    DeviceGuard dev_guard(fo->dev->device);
    wann_tree::Tree copy;
    if (download_tree(fo->dev, tree, &copy) != WANN_OK) return -1;
    const int64_t r = copy.dump(rows, rows ? capacity : 0);
    copy.release();
    return r;
  }
  return fo->f.trees[tree].dump(rows, rows ? capacity : 0);
}

}  // extern "C"
