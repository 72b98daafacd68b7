// conv_stack_tc: the WaNN policy-net conv stack (encode -> conv1..conv5) as ONE persistent,
// warp-specialised sm_100a kernel.  Replaces the TF graph ops
//   hidden_layer/{1..5}/{Conv2D,BiasAdd,Relu}   (Breakthrough_Player/policy_net_utils.pyx:149-178)
// and the Python board->plane encoder (tools/utils.pyx:546-602,819-866) for batched leaf
// positions.  Output: h5[n][64] fp32 = the post-ReLU 1-channel 5th feature map, which the tail
// kernel (aux_kernels.cuh) turns into logits / softmax / ranked legal priors.
//
// Design (see DESIGN.md "K2"):
//   * CTA pair (cluster of 2, tcgen05 cta_group::2).  A pair owns a "supertile" of 8 boards:
//     each CTA holds 4 boards = two M=128 row tiles (2 boards x 64 squares each); one
//     tcgen05.mma covers M=256 (the same tile index in both CTAs = a "half") x N=192 x K=16.
//   * Activations NEVER leave the SM: a layer's 16-bit output is written by the epilogue warps
//     straight into the shared-memory layout the next layer's MMAs read as operand A.
//     Layout per row tile ("A tile"):  3 groups x [8 channel-chunks of 8][9 rows y=-1..7][2 boards]
//     [9 cells x=-1..7][16 B] (+ a trailing halo row per group), i.e. a haloed board image whose
//     16-byte cells are the 8x16B "core matrix" rows of a K-major SWIZZLE_NONE UMMA descriptor.
//     (Inside a group the y=8 halo row of a chunk IS the y=-1 row of the next chunk, the x=8 halo
//     cell of a row IS the x=-1 cell of the next row; a group = the 64 channels of one K-block of
//     the weight stream, and -- centred mode -- the unit that shares halo cells, hence offsets.)
//     MMA row m of a tile is square (y = m/16, board = (m/8)%2, x = m%8), so
//     an 8-row core-matrix group is one board row, the group stride (SBO) is one padded row =
//     144 B and the K-chunk stride (LBO) is 2592 B.  The 3x3 taps are then nothing but nine
//     descriptor START ADDRESSES ((dy+1)*288 + (dx+1)*16 bytes): implicit GEMM with zero im2col
//     traffic, SAME padding supplied by the halo (zero and never written in the bf16 / fp16 modes;
//     minus the layer's offsets, rewritten per layer, in the centred mode).
//   * Weights (operand B) stream from L2 through a ring of shared-memory stages via
//     cp.async.bulk (TMA engine, 1-D).  They are pre-arranged on the host into the exact
//     shared-memory image of each stage (K-major SWIZZLE_NONE core matrices; each CTA of the
//     pair holds its half of the 192 output channels), so no tensor map is needed.  Every layer
//     conv1..4 opens with a BIAS stage (the bias in three 16-bit parts at k = 0..2) that is
//     multiplied with a tile of ones as the layer's first MMA: the accumulator starts at the bias
//     and the epilogue neither loads nor adds one.  Optional lo stages (w - fp16(w) of the leading
//     input group x leading output positions, several taps per stage) follow their hi stages.
//   * The two halves (row tile 0 / row tile 1) are issued by TWO MMA warps and consume every
//     weight stage one after the other: a stage is freed when BOTH have used it.  Because the
//     eight epilogue warps serve the halves alternately, half 1 settles a few K-blocks behind
//     half 0, and each half's epilogue (TMEM -> ReLU (centred: max with -offset), 16-bit -> shared
//     memory) runs while the tensor core works on the other half.
//   * conv1 (4->192) runs as 9 taps x K=16 (4 real channels), conv2-4 as 9 taps x 3 x K=64,
//     conv5 (192->1, 3x3) as a 1x1 GEMM with N=32 whose columns are the 9 TAPS (and their lo parts)
//     (P[pixel][tap] = <act4[pixel], w5[tap]>) followed by a 9-point stencil sum in the epilogue.
//   * Warp roles per CTA (416 threads): warp 0 = weight producer, warps 1/2 = MMA issuers of
//     half 0/1 (leader CTA; in the peer CTA warp 1 relays "my half of the stage has landed"),
//     warps 3..10 = epilogue, warps 11/12 = FUSED TAIL of half 0/1: the conv5 epilogue hands the
//     5th feature map of the half's two boards over in shared memory and the tail warp runs
//     FC-155 + bias -> softmax -> legal gather -> rank (aux_kernels.cuh tail_group) while the tensor
//     pipe is already on the next supertile: the whole evaluation is ONE kernel and h5 never
//     exists in global memory.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "aux_kernels.cuh"
#include "ptx.cuh"

namespace wann {

// ---- geometry ------------------------------------------------------------------------------
constexpr int kC = 192;                       // hidden channels
constexpr int kKChunks = kC / 8;              // 24 chunks of 8 channels (16 B of bf16/fp16)
constexpr int kCell = 16;                     // bytes per (pixel, channel-chunk) cell
constexpr int kRowStride = 9 * kCell;         // 144 B: x = -1..7 (x = 8 halo is next row's x = -1)
constexpr int kYStride = 2 * kRowStride;      // 288 B: two boards interleaved per y
constexpr int kChunkStride = 9 * kYStride;    // 2592 B: y = -1..7 (y = 8 halo is next chunk's y = -1)
constexpr int kTilesPerCta = 2;
// The 24 chunks of a tile form 3 GROUPS of 8 (= the 64-channel K-blocks of the weight stream).  Every group
// carries its OWN trailing halo row (the y = 8 row of its last chunk + the x = 8 cell of that row), so that
// no halo cell is shared between groups (centred mode: a halo cell holds -offset, and offsets differ per
// group) or between the two tiles of a CTA (which are in different layers at the same time).
constexpr int kGroupChunks = 8;
constexpr int kGroups = kKChunks / kGroupChunks;                         // 3
constexpr int kGroupGap = kYStride + kCell;                              // 304 B
constexpr int kGroupStride = kGroupChunks * kChunkStride + kGroupGap;    // 21040 B
constexpr int kATileStride = kGroups * kGroupStride;                     // 63120 B per 128-row tile
constexpr int kABytes = kTilesPerCta * kATileStride;
constexpr int kHaloCells = kKChunks * 34 + kGroups * 19;   // per tile: 18 (y = -1 row) + 16 (x = -1 cells) per chunk, + a trailing row per group
constexpr int kCenBuckets = kGroups * 8;      // centred mode: one offset per (group, channel position % 8)
// lo weight stages (partial hi/lo split): a (tap, j) block is [lo_n / 2 rows x 64 k] = 64 * lo_n bytes per CTA;
// a stage carries as many consecutive blocks as fit a ring slot
__host__ __device__ constexpr uint32_t lo_blocks_per_stage(int lo_n) {
  return (12288 / (64 * lo_n)) > 0 ? static_cast<uint32_t>(12288 / (64 * lo_n)) : 1u;
}
// byte offset of channel chunk q inside a tile
__host__ __device__ constexpr int chunk_off(int q) { return q * kChunkStride + (q / kGroupChunks) * kGroupGap; }
constexpr int kBoardsPerCta = 4;
constexpr int kBoardsPerPair = 8;

constexpr int kStageBytes = 96 * 64 * 2;      // 12288 B: half of [192 x 64] 16-bit
constexpr int kStageBytesL5 = 16 * 64 * 2;    // 2048 B: half of [32 x 64] 16-bit (rows 0-15 = taps, 16-31 = their lo parts)
constexpr int kStageBytesBias = 96 * 16 * 2;  // 3072 B: the bias stage of a layer, half of [192 x 16] (k = 0..2: the bias in three 16-bit parts)
constexpr int kStages = 8;
constexpr int kLboB = 96 / 8 * 128;           // 1536 B between K-chunks of a stage
constexpr int kLboB5 = 16 / 8 * 128;          // 256 B
constexpr int kSboB = 128;

constexpr int kKbL1 = 3, kKbMid = 27, kKbL5 = 3;
constexpr int kKbPerSupertile = kKbL1 + 3 * kKbMid + kKbL5;
constexpr int kImgBytesPerRank = (kKbL1 + 3 * kKbMid) * kStageBytes + kKbL5 * kStageBytesL5;

constexpr int kNumEpiWarps = 8;
constexpr int kFirstEpiWarp = 3;
constexpr int kNumTailWarps = 2;
constexpr int kFirstTailWarp = kFirstEpiWarp + kNumEpiWarps;                    // 11
constexpr int kThreads = 32 * (kFirstEpiWarp + kNumEpiWarps + kNumTailWarps);   // 416
constexpr int kBiasFloats = 4 * kC + 4;

// shared memory carve-up (bytes from the 128-aligned base)
// (the weight ring first: its 128-byte core matrices then sit on 128-byte lines; the activation cells are read at
// tap offsets anyway and need 16-byte alignment only)
constexpr int kOffB = 0;
constexpr int kOffA = kOffB + kStages * kStageBytes;
constexpr int kOffBias = kOffA + kABytes;
static_assert(kOffA % 128 == 0 && kOffBias % 16 == 0, "shared memory carve-up alignment");
constexpr int kOffBar = kOffBias + kBiasFloats * 4;
constexpr int kNumBars = 3 * kStages + 8;
constexpr int kOffTmemSlot = kOffBar + kNumBars * 8;
// fused tail: h5 hand-over [2 halves][2 boards][64] f32, FC bias [160] f32, move table [160] u16,
// then per tail warp 240 floats (h copy [2][64], compact priors [64], ranked priors [48]) and
// 240 bytes (pov [2][64], compact indices [64], ranked indices [48])
constexpr int kOffTail = kOffTmemSlot + 16;
constexpr int kTailH5Bytes = 2 * 2 * 64 * 4;
constexpr int kTailWarpF = 240, kTailWarpB = 240;
constexpr int kOffTailBias = kOffTail + kTailH5Bytes;
constexpr int kOffTailMoves = kOffTailBias + 160 * 4;
constexpr int kOffTailWarpF = kOffTailMoves + 160 * 2;
constexpr int kOffTailWarpB = kOffTailWarpF + kNumTailWarps * kTailWarpF * 4;
constexpr int kSmemBytes = kOffTailWarpB + kNumTailWarps * kTailWarpB + 128;   // +128 for base alignment slack
static_assert(kOffTail % 16 == 0 && kOffTailWarpF % 16 == 0 && kOffTailWarpB % 16 == 0, "tail scratch alignment");
static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB) exceeded");

constexpr long long kWatchdogCycles = 1ll << 31;   // ~1.2 s: no legitimate wait comes close
constexpr int kProfSteps = 64;

struct TcParams {
  const int8_t* squares;     // [n][64] or null
  const uint8_t* black;      // [n]
  const void* planes;        // [n][8][8][4] int8/f32 or null
  int planes_f32;
  long long n;
  const uint8_t* wimg;       // [2][kImgBytesPerRank]: the stage images of CTA rank 0 / 1
  const float* bias;         // [4][192] conv1..4 bias, then b5
  float* h5;                 // [n][64]
  float* dbg;                // optional activation dump [n][64][192] (layer 1..4)
  int dbg_layer;
  int variant;               // debug: bit2 = swap the B halves between the CTAs of a pair, bit3 = skip the MMAs
  int skew;                  // weight stages by which half 1 trails the stream (0 = free-running)
  int halves;                // 2 = 8 boards per CTA pair (throughput), 1 = 4 boards (small batches)
  int* err;                  // [8] watchdog record
  long long* prof;           // optional [8][kProfSteps] clock64 stamps of pair 0 (debug):
                             // half h: [4h+0] MMA start, [4h+1] MMA issue end,
                             //         [4h+2] epilogue start, [4h+3] epilogue end
  // margin refinement pass (conv_stack_split_kernel only): evaluate the boards gather[0..*n_dev)
  // of `squares` and write their h5 rows COMPACTLY (row i <- board gather[i]); both null otherwise
  const int* gather = nullptr;
  // device-side position count (both kernels): evaluate min(n, max(0, *n_dev - n_dev_off)) positions -- the
  // refinement pass's flagged boards, or the rows a device-resident search forest asked for this step.
  // Read at kernel start: such a launch is never made under programmatic dependent launch.
  const int* n_dev = nullptr;
  long long n_dev_off = 0;
  // fused tail (conv_stack_tc_kernel only): when fuse_tail != 0 the tail warps produce tail.probs /
  // logits / idx / prior / n_legal (and the refinement flags) themselves and h5 may be null
  int fuse_tail = 0;
  TailParams tail;
  int* err_mirror = nullptr;   // host-mapped copy of err[0..5], written by the last CTA to finish
  // weight stream geometry: stages per supertile (kKbPerSupertile + 4 bias stages + the lo stages of the split layers) and the
  // byte distance between the stage images of CTA rank 0 and 1
  int stages_per_st = kKbPerSupertile;
  long long img_stride = kImgBytesPerRank;
  // hi/lo weight split: in conv(l+1), bit l (1..3) of split_mask, every K-block (tap, j) with j < lo_kgroups
  // streams a second stage -- the lo parts w - fp16(w) of output positions [0, lo_n) x its 64 input
  // positions -- and the issuer runs the same A descriptors against it with N = lo_n: the weight rounding
  // error of that block of the layer disappears.  (Channel positions are sorted by activation variance,
  // so the leading positions are the ones whose weight error matters.)
  int split_mask = 0;
  int lo_n = 192;              // multiple of 16, 16..192
  int lo_kgroups = 3;          // 1..3
  // centred mode (kCen kernels): the stored activation of channel position q of conv l+1 is
  // a' = a - cen[l][(q / 64) * 8 + q % 8]; the halo holds -cen (so that a' + cen == 0 there) and the next
  // layer's bias absorbs sum(w * cen).  kappa[t] = sum_k w5[t][k] * cen[3][..] corrects conv5's tap products.
  float cen[4][kCenBuckets] = {};
  uint32_t cen_neg16[4][kGroups][4] = {};   // the same offsets, negated and packed as fp16 pairs (what the epilogue needs)
  float kappa[9] = {};
  const uint8_t* dbg_unperm = nullptr;   // [4][192] channel position -> TF channel (debug dump of a permuted net)
};

// bounded mbarrier wait: returns false (and records who/where) instead of hanging the GPU
__device__ __forceinline__ bool wait_bar(uint32_t bar, uint32_t parity, bool cluster_scope,
                                         int* err, int code) {
  if (cluster_scope ? ptx::mbar_try_wait_cluster(bar, parity) : ptx::mbar_try_wait(bar, parity))
    return true;
  long long t0 = clock64();
  while (true) {
    if (ptx::mbar_try_wait_parked(bar, parity)) return true;
    if (clock64() - t0 > kWatchdogCycles) {
      if (atomicCAS(err, 0, code) == 0) {
        err[1] = static_cast<int>(blockIdx.x);
        err[2] = static_cast<int>(parity);
      }
      return false;
    }
  }
}

// two floats -> packed 16-bit pair in the operand format (bf16 or fp16), lo = a
template <bool kF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (kF16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// The conv1 input cell of one square = planes P,O,E,1 in the operand format + 4 zero channels
// (tools/utils.pyx:563-600: Black sees the board rotated by 180 degrees; :842-865: P/O/E/bias
// planes).  Split into a global-memory load (issued early, during the conv4 epilogue, so that its
// latency is off the conv5 -> conv1 critical chain) and the shared-memory store.
__device__ __forceinline__ long long effective_n(const TcParams& p) {
  if (p.n_dev == nullptr) return p.n;
  return max(0ll, min(p.n, static_cast<long long>(*p.n_dev) - p.n_dev_off));
}
template <bool kF16>
__device__ __forceinline__ uint2 load_input_cell(const TcParams& p, long long n_tot, long long board, int y, int x) {
  uint2 v = make_uint2(0u, 0u);
  if (board < n_tot) {
    const int sq = y * 8 + x;
    if (p.squares != nullptr) {
      const int8_t* b64 = p.squares + board * 64;
      const int s_w = b64[sq], s_b = b64[63 - sq];       // both candidates: no dependent load
      const int blk = p.black[board];
      int s = blk ? s_b : s_w;
      if (blk && s) s = 3 - s;
      const uint32_t one = kF16 ? 0x3C00u : 0x3F80u;      // 1.0 in fp16 / bf16
      v.x = (s == 1 ? one : 0u) | ((s == 2 ? one : 0u) << 16);
      v.y = (s == 0 ? one : 0u) | (one << 16);
    } else if (p.planes_f32) {
      const float4 f = reinterpret_cast<const float4*>(p.planes)[board * 64 + sq];
      v.x = pack2<kF16>(f.x, f.y);
      v.y = pack2<kF16>(f.z, f.w);
    } else {
      const char4 c = reinterpret_cast<const char4*>(p.planes)[board * 64 + sq];
      v.x = pack2<kF16>(float(c.x), float(c.y));
      v.y = pack2<kF16>(float(c.z), float(c.w));
    }
  }
  return v;
}
// Centred mode: channels 4..7 carry the same planes scaled by 2^-12; the weight image holds
// (w1 - fp16(w1)) * 4096 there, so conv1 sees its weights to ~22 bits at no extra MMA.
template <bool kCen>
__device__ __forceinline__ void store_input_cell(uint32_t cell0, uint2 v) {
  if (kCen) {
    const __half2 s = __half2half2(__ushort_as_half(static_cast<unsigned short>(0x0C00)));   // 2^-12
    const __half2 dx = __hmul2(*reinterpret_cast<const __half2*>(&v.x), s);
    const __half2 dy = __hmul2(*reinterpret_cast<const __half2*>(&v.y), s);
    ptx::st_shared_v4(cell0, make_uint4(v.x, v.y, *reinterpret_cast<const uint32_t*>(&dx),
                                        *reinterpret_cast<const uint32_t*>(&dy)));
  } else {
    ptx::st_shared_v4(cell0, make_uint4(v.x, v.y, 0u, 0u));
  }
}
// byte offset (inside a tile) of halo cell i in [0, kHaloCells), with the cell's group in bits 0..1
__device__ __forceinline__ uint32_t halo_cell_offset(int i) {
  if (i >= kKChunks * 34) {                                   // the trailing rows of the groups
    const int t = i - kKChunks * 34, g = t / 19, r = t - 19 * g;
    return static_cast<uint32_t>(g * kGroupStride + kGroupChunks * kChunkStride + r * kCell) | g;
  }
  const int q = i / 34, r = i - 34 * q;
  return static_cast<uint32_t>(chunk_off(q) + (r < 18 ? r * kCell : kYStride + (r - 18) * kRowStride)) | (q / kGroupChunks);
}

template <bool kF16, bool kCen>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_stack_tc_kernel(const TcParams p) {
  static_assert(kF16 || !kCen, "the centred mode is fp16 only");
  constexpr uint32_t kIdescN192 = ptx::make_idesc_16bit(256, 192, kF16);
  constexpr uint32_t kIdescN32 = ptx::make_idesc_16bit(256, 32, kF16);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t a_base = smem_base + kOffA;
  const uint32_t b_base = smem_base + kOffB;
  const uint32_t bar_base = smem_base + kOffBar;
  const uint32_t slot_addr = smem_base + kOffTmemSlot;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const float* bias_s = reinterpret_cast<const float*>(smem_gen + kOffBias);

  auto bar_full = [&](int s) { return bar_base + 8u * s; };
  auto bar_peer_full = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto bar_empty = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto bar_acc_full = [&](int h) { return bar_base + 8u * (3 * kStages + h); };
  auto bar_a_ready = [&](int h) { return bar_base + 8u * (3 * kStages + 2 + h); };
  auto bar_h5_full = [&](int h) { return bar_base + 8u * (3 * kStages + 4 + h); };
  auto bar_h5_empty = [&](int h) { return bar_base + 8u * (3 * kStages + 6 + h); };
  float* h5s = reinterpret_cast<float*>(smem_gen + kOffTail);      // [half][board][64]

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  // small batches run with ONE half per CTA (4 boards per pair): twice the supertiles to spread
  // over the SM pairs, and each 5-layer chain has the tensor pipe to itself (lower latency)
  const int nh = (p.halves == 1) ? 1 : 2;
  const int boards_per_pair = 4 * nh;
  const long long n_tot = effective_n(p);
  const long long total_st = (n_tot + boards_per_pair - 1) / boards_per_pair;
  constexpr uint32_t nst = kStages;

  // ---- one-time setup ---------------------------------------------------------------------
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(bar_full(s), 1);
      ptx::mbar_init(bar_peer_full(s), 1);
      ptx::mbar_init(bar_empty(s), nh);     // freed when ALL active halves' MMAs on the stage are done
    }
    for (int h = 0; h < 2; ++h) {
      ptx::mbar_init(bar_acc_full(h), 1);
      ptx::mbar_init(bar_a_ready(h), 2);    // one arrival per CTA of the pair
      ptx::mbar_init(bar_h5_full(h), 128);  // the 128 conv5 epilogue threads of the half
      ptx::mbar_init(bar_h5_empty(h), 1);   // the half's tail warp
    }
    ptx::fence_mbar_init();
  }
  // zero the activation tiles (the halo cells stay zero for the whole kernel)
  for (int i = tid; i < kABytes / 16; i += kThreads)
    ptx::st_shared_v4(a_base + 16u * i, make_uint4(0u, 0u, 0u, 0u));
  // bias region: the first 128 bytes are the ONES tile -- one 8 x 16-byte core matrix whose rows are
  // {1, 1, 1, 0, 0, 0, 0, 0}: operand A of the bias MMA that opens every layer conv1..4 (descriptor strides 0: all
  // 128 rows and both K chunks read this one matrix; the bias stage holds the bias in three parts at k = 0..2 and
  // zeros at k = 8..10).  The conv1..4 biases themselves are not needed here any more; b5 is (bias_s[4 * kC]).
  for (int i = tid; i < kBiasFloats; i += kThreads) {
    constexpr uint32_t one = kF16 ? 0x3C00u : 0x3F80u;
    uint32_t bits = (i < 4 * kC + 1) ? __float_as_uint(p.bias[i]) : 0u;
    if (i < 32) bits = (i % 4 == 0) ? (one | (one << 16)) : ((i % 4 == 1) ? one : 0u);
    reinterpret_cast<uint32_t*>(smem_gen + kOffBias)[i] = bits;
  }
  ptx::fence_proxy_async_smem();
  if (warp == 0) {
    ptx::tmem_alloc<2>(slot_addr, 512);
    ptx::tmem_relinquish<2>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + kOffTmemSlot);
  // Programmatic dependent launch: everything above (and the weight prefetch below) only touches
  // constants, so it may overlap the previous kernel of the stream; let the next kernel start
  // its own prologue too.  The epilogue warps -- the only ones that read the positions and write
  // h5 -- wait for the previous kernels before their first access.
  ptx::pdl_launch_dependents();

  // ---- roles -------------------------------------------------------------------------------
  if (warp == 0) {
    // ===== weight producer: stream this CTA's half of every weight stage ====================
    if (lane == 0) {
      const uint32_t img_rank = (p.variant & 4) ? (rank ^ 1u) : rank;
      const uint8_t* img = p.wimg + static_cast<size_t>(img_rank) * p.img_stride;
      uint32_t s = 0, ph = 0;
      bool ok = true;
      size_t off = 0;
      auto stage = [&](uint32_t bytes) {
        if (!ok) return;
        if (!wait_bar(bar_empty(s), ph ^ 1u, false, p.err, 100 + s)) { ok = false; return; }
        ptx::mbar_arrive_expect_tx(bar_full(s), bytes);
        ptx::bulk_g2s(b_base + s * kStageBytes, img + off, bytes, bar_full(s));
        off += bytes;
        if (++s == nst) { s = 0; ph ^= 1u; }
      };
      const uint32_t lo_block = 64u * static_cast<uint32_t>(p.lo_n);      // [lo_n / 2 rows x 64 k] 16-bit
      const uint32_t bps = lo_blocks_per_stage(p.lo_n);
      for (long long st = pair; st < total_st && ok; st += npairs) {
        off = 0;
        stage(kStageBytesBias);
        for (int kb = 0; kb < kKbL1; ++kb) stage(kStageBytes);
        for (int layer = 1; layer <= 3; ++layer) {
          stage(kStageBytesBias);
          const int lo_j = ((p.split_mask >> layer) & 1) ? p.lo_kgroups : 0;
          uint32_t pend = 0;
          for (int tap = 0; tap < 9; ++tap)
            for (int j = 0; j < 3; ++j) {
              stage(kStageBytes);
              if (j < lo_j && (++pend == bps || (tap == 8 && j == lo_j - 1))) {
                stage(pend * lo_block);
                pend = 0;
              }
            }
        }
        for (int kb = 0; kb < kKbL5; ++kb) stage(kStageBytesL5);
      }
    }
  } else if (warp == 1 && rank != 0) {
    // ===== relay (peer CTA): forward "my half of stage s has landed" to the leader ===========
    if (lane == 0) {
      const int nstages = p.stages_per_st;
      uint32_t s = 0, ph = 0;
      bool ok = true;
      for (long long st = pair; st < total_st && ok; st += npairs) {
        for (int kb = 0; kb < nstages; ++kb) {
          if (!wait_bar(bar_full(s), ph, false, p.err, 500 + s)) { ok = false; break; }
          ptx::mbar_arrive_cluster(ptx::mapa(bar_peer_full(s), 0));
          if (++s == nst) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp <= 2) {
    if (rank == 0 && warp - 1 < nh) {
      // ===== MMA issuer of half h = warp - 1 (leader CTA).  The WHOLE warp runs the loop with
      // uniform control flow (descriptors live in uniform registers); one elected lane issues
      // the tcgen05 ops.  The 64-bit descriptors are {lo = start>>4 | LBO<<16, hi = SBO |
      // version}; only `lo` changes from MMA to MMA (one add).
      const int h = warp - 1;
      const uint32_t issuer = ptx::elect_one();
      const bool skip_mma = p.variant & 8;        // timing experiment: weight stream only
      const uint32_t a_hi = (kRowStride >> 4) | (1u << 14);
      const uint32_t b_hi = (kSboB >> 4) | (1u << 14);
      const uint32_t a_lo = ((a_base + h * kATileStride) >> 4) | ((kChunkStride >> 4) << 16);
      const uint32_t b_lo_mid = (b_base >> 4) | ((kLboB >> 4) << 16);
      const uint32_t b_lo_l5 = (b_base >> 4) | ((kLboB5 >> 4) << 16);
      const uint32_t ones_lo = (smem_base + kOffBias) >> 4;       // LBO = 0
      const uint32_t ones_hi = 1u << 14;                          // SBO = 0
      // lo stages: [lo_n / 2 rows x 64 k] per CTA, LBO = 8 * lo_n bytes, K-step = 2 LBO = lo_n sixteen-byte units
      const uint32_t lo_n = static_cast<uint32_t>(p.lo_n);
      const uint32_t b_lo_lo = (b_base >> 4) | (((8u * lo_n) >> 4) << 16);
      const uint32_t idesc_lo = ptx::make_idesc_16bit(256, p.lo_n, kF16);
      const uint32_t lo_bps = lo_blocks_per_stage(p.lo_n);
      const uint32_t tmem_d = tmem_base + h * kC;
      const uint32_t acc_bar = bar_acc_full(h), ready_bar = bar_a_ready(h);
      constexpr uint32_t kChunk16 = kChunkStride >> 4;          // 162
      constexpr uint32_t kGroup16 = kGroupStride >> 4;          // 1315
      constexpr uint32_t kStage16 = kStageBytes >> 4;           // 768
      uint32_t s = 0, ph = 0, lstep = 0;
      bool ok = true;
      // half 1 may be held `lookahead` weight stages behind the stream so that it still has MMAs
      // to run while half 0 sits in its epilogue (and vice versa at the start of the next layer)
      const uint32_t lookahead = (h == 1) ? min(static_cast<uint32_t>(p.skew), nst - 2u) : 0u;
      long long kb_left = 0;
      for (long long st = pair; st < total_st; st += npairs) kb_left += p.stages_per_st;
      long long* prof = (p.prof != nullptr && pair == 0) ? p.prof + 4 * h * kProfSteps : nullptr;
      // one K-block: wait for the stage, 4 (or fewer) K=16 steps, release my claim on the stage
      // one weight stage: wait for it (WANN_STAGE_BEGIN), issue the MMAs, release my claim on it (WANN_STAGE_END)
#define WANN_STAGE_BEGIN                                                                          \
      {                                                                                           \
        if (lookahead != 0 && kb_left > lookahead) {   /* half 1 trails the stream by d stages */ \
          const uint32_t s2 = (s + lookahead >= nst) ? s + lookahead - nst : s + lookahead;       \
          const uint32_t ph2 = (s + lookahead >= nst) ? ph ^ 1u : ph;                             \
          if (!wait_bar(bar_full(s2), ph2, false, p.err, 700 + s2)) { ok = false; break; }        \
        }                                                                                         \
        --kb_left;                                                                                \
        if (!wait_bar(bar_full(s), ph, false, p.err, 300 + s) ||                                 \
            !wait_bar(bar_peer_full(s), ph, true, p.err, 400 + s)) { ok = false; break; }       \
        ptx::tc_fence_after();
#define WANN_STAGE_END                                                                            \
        if (issuer) ptx::umma_commit<2>(bar_empty(s), 3);   /* my claim, in BOTH CTAs */          \
        if (++s == nst) { s = 0; ph ^= 1u; }                                                  \
      }
#define WANN_KBLOCK(NKS, A_OFF16_OF_KS, B_LO, B_KS16, IDESC, ACC_FIRST)                          \
      WANN_STAGE_BEGIN                                                                            \
        const uint32_t blo = (B_LO) + s * kStage16;                                               \
        _Pragma("unroll") for (int ks = 0; ks < (NKS); ++ks) {                                   \
          const uint32_t ao = (A_OFF16_OF_KS);                                                    \
          const uint32_t acc = ((ACC_FIRST) && ks == 0) ? 0u : 1u;                                \
          if (issuer && !skip_mma)                                                                \
            ptx::umma_bf16_2cta(tmem_d, a_lo + ao, a_hi, blo + ks * (B_KS16), b_hi, (IDESC), acc); \
        }                                                                                         \
      WANN_STAGE_END
      // the layer's first MMA: accumulator = ones x bias stage (no accumulate), i.e. the bias in every row
#define WANN_BIAS_BLOCK                                                                           \
      WANN_STAGE_BEGIN                                                                            \
        if (issuer && !skip_mma)                                                                  \
          ptx::umma_bf16_2cta(tmem_d, ones_lo, ones_hi, b_lo_mid + s * kStage16, b_hi, kIdescN192, 0u); \
      WANN_STAGE_END
      for (long long st = pair; st < total_st && ok; st += npairs) {
        // ---- conv1: 9 taps x K=16 (4 real planes); K-block kb holds taps 4kb..4kb+3
        do {
          if (!wait_bar(ready_bar, lstep & 1u, true, p.err, 200 + 10 * h)) { ok = false; break; }
          ptx::tc_fence_after();
          if (prof && lstep < kProfSteps && issuer) prof[lstep] = clock64();
          constexpr uint32_t kTap16[12] = {0, 1, 2, 18, 19, 20, 36, 37, 38, 0, 0, 0};
          WANN_BIAS_BLOCK
          WANN_KBLOCK(4, kTap16[ks], b_lo_mid, (2 * kLboB) >> 4, kIdescN192, false)
          WANN_KBLOCK(4, kTap16[4 + ks], b_lo_mid, (2 * kLboB) >> 4, kIdescN192, false)
          WANN_KBLOCK(1, kTap16[8 + ks], b_lo_mid, (2 * kLboB) >> 4, kIdescN192, false)
          if (issuer) ptx::umma_commit<2>(acc_bar, 3);
          if (prof && lstep < kProfSteps && issuer) prof[kProfSteps + lstep] = clock64();
          ++lstep;
        } while (false);
        // ---- conv2..4: K-block (tap, j) = input channels 64j..64j+63 of tap
        for (int layer = 1; layer <= 3 && ok; ++layer) {
          if (!wait_bar(ready_bar, lstep & 1u, true, p.err, 200 + 10 * h + layer)) { ok = false; break; }
          ptx::tc_fence_after();
          if (prof && lstep < kProfSteps && issuer) prof[lstep] = clock64();
          WANN_BIAS_BLOCK
          const uint32_t first = 0;
          // hi/lo split: the blocks (tap, j < lo_j) also have lo weights; they arrive lo_bps blocks per stage,
          // right after the hi stage of the last block, and run with N = lo_n against the same A descriptors
          const uint32_t lo_j = ((p.split_mask >> layer) & 1) ? static_cast<uint32_t>(p.lo_kgroups) : 0u;
          uint32_t pend = 0, first_blk = 0;
          for (uint32_t ty = 0; ty < 3 && ok; ++ty)
            for (uint32_t tx = 0; tx < 3 && ok; ++tx) {
              const uint32_t tap16 = ty * (kYStride >> 4) + tx;
              for (uint32_t j = 0; j < 3; ++j) {
                const uint32_t ao0 = j * kGroup16 + tap16;
                WANN_KBLOCK(4, ao0 + ks * 2 * kChunk16, b_lo_mid, (2 * kLboB) >> 4, kIdescN192, first)
                if (j < lo_j && (++pend == lo_bps || (ty == 2 && tx == 2 && j == lo_j - 1))) {
                  // one lo stage = the blocks first_blk .. first_blk + pend - 1 (block index = tap * lo_j + j)
                  WANN_STAGE_BEGIN
                  const uint32_t blo = b_lo_lo + s * kStage16;
                  for (uint32_t b = 0; b < pend; ++b) {
                    const uint32_t blk = first_blk + b, btap = blk / lo_j, bj = blk - btap * lo_j;
                    const uint32_t bao = bj * kGroup16 + (btap / 3u) * (kYStride >> 4) + (btap % 3u);
                    _Pragma("unroll") for (uint32_t ks = 0; ks < 4; ++ks)
                      if (issuer && !skip_mma)
                        ptx::umma_bf16_2cta(tmem_d, a_lo + bao + ks * 2 * kChunk16, a_hi,
                                            blo + b * 4u * lo_n + ks * lo_n, b_hi, idesc_lo, 1u);
                  }
                  WANN_STAGE_END
                  first_blk += pend;
                  pend = 0;
                }
              }
            }
          if (!ok) break;
          if (issuer) ptx::umma_commit<2>(acc_bar, 3);   // this half's accumulators ready, both CTAs
          if (prof && lstep < kProfSteps && issuer) prof[kProfSteps + lstep] = clock64();
          ++lstep;
        }
        if (!ok) break;
        // ---- conv5 as a 1x1 GEMM over the centre tap with N = 16 (columns = the 9 taps)
        do {
          if (!wait_bar(ready_bar, lstep & 1u, true, p.err, 204 + 10 * h)) { ok = false; break; }
          ptx::tc_fence_after();
          if (prof && lstep < kProfSteps && issuer) prof[lstep] = clock64();
          constexpr uint32_t kCentre16 = (kYStride >> 4) + 1;
          for (uint32_t j = 0; j < 3; ++j) {
            const uint32_t ao0 = j * kGroup16 + kCentre16;
            WANN_KBLOCK(4, ao0 + ks * 2 * kChunk16, b_lo_l5, (2 * kLboB5) >> 4, kIdescN32, j == 0)
          }
          if (!ok) break;
          if (issuer) ptx::umma_commit<2>(acc_bar, 3);
          if (prof && lstep < kProfSteps && issuer) prof[kProfSteps + lstep] = clock64();
          ++lstep;
        } while (false);
      }
#undef WANN_KBLOCK
#undef WANN_BIAS_BLOCK
#undef WANN_STAGE_BEGIN
#undef WANN_STAGE_END
    }
  } else if (warp < kFirstTailWarp) {
    // ===== epilogue warps: all 8 serve half 0, then half 1, alternately =======================
    const int ew = warp - kFirstEpiWarp;       // 0..7
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int chalf = ew >> 2;                 // which 96 of the 192 accumulator columns
    const int row = quad * 32 + lane;          // MMA row inside the tile
    const int y = row >> 4, b = (row >> 3) & 1, x = row & 7;
    const uint32_t cell_in_tile = ((y + 1) * 2 + b) * kRowStride + (x + 1) * kCell;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const bool signaller = (tid == kFirstEpiWarp * 32);
    const uint32_t a_ready_leader0 = ptx::mapa(bar_a_ready(0), 0);
    const uint32_t a_ready_leader1 = ptx::mapa(bar_a_ready(1), 0);
    const bool profiler = (p.prof != nullptr && blockIdx.x == 0 && signaller);
    uint32_t halo_off[4];      // centred mode: the (up to) 4 halo cells of a tile this thread rewrites per layer
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = ew * 32 + lane + i * kNumEpiWarps * 32;
      halo_off[i] = c < kHaloCells ? halo_cell_offset(c) : 0xFFFFFFFFu;
    }

    auto signal_a_ready = [&](int h) {
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(kNumEpiWarps * 32) : "memory");
      if (signaller) ptx::mbar_arrive_cluster(h ? a_ready_leader1 : a_ready_leader0);
    };
    auto board_of = [&](long long st, int h) {
      return st * boards_per_pair + static_cast<int>(rank) * (2 * nh) + h * 2 + b;
    };

    ptx::pdl_wait();
    long long st = pair;
    if (st < total_st) {
      for (int h = 0; h < nh; ++h) {
        if (chalf == 0)
          store_input_cell<kCen>(a_base + h * kATileStride + cell_in_tile, load_input_cell<kF16>(p, n_tot, board_of(st, h), y, x));
        signal_a_ready(h);
      }
    }
    uint32_t lstep = 0;
    bool ok = true;
    uint2 next_in[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};
    for (; st < total_st && ok; st += npairs) {
      for (int layer = 0; layer < 5 && ok; ++layer, ++lstep) {
        for (int h = 0; h < nh; ++h) {
          const long long board = board_of(st, h);
          const uint32_t cell0 = a_base + h * kATileStride + cell_in_tile;
          const uint32_t taddr = lane_taddr + h * kC;
          if (!wait_bar(bar_acc_full(h), lstep & 1u, true, p.err, 600 + 10 * h + layer)) { ok = false; break; }
          ptx::tc_fence_after();
          if (profiler && lstep < kProfSteps) p.prof[(4 * h + 2) * kProfSteps + lstep] = clock64();
          if (layer < 4) {
            if (layer == 3 && chalf == 0 && st + npairs < total_st)   // prefetch next supertile's input
              next_in[h] = load_input_cell<kF16>(p, n_tot, board_of(st + npairs, h), y, x);
            float* dbg = (p.dbg != nullptr && p.dbg_layer == layer + 1 && board < n_tot)
                             ? p.dbg + (board * 64 + y * 8 + x) * kC : nullptr;
            // centred mode: store a' = max(acc + (b' - c), -c) = relu(acc + b') - c   (c by chunk group and
            // channel % 8); bias_s already holds b' - c.
            // -c as packed 16-bit pairs: rounding is monotonic and -c is exact in the operand format, so
            // fp16(max(t, -c)) == max(fp16(t), -c): one packed max per PAIR after the conversion
            uint4 hneg[kGroups];            // -c of the three chunk groups, packed (host-made: TcParams::cen_neg16)
#pragma unroll
            for (int g = 0; g < kGroups; ++g)
              hneg[g] = kCen ? make_uint4(p.cen_neg16[layer][g][0], p.cen_neg16[layer][g][1], p.cen_neg16[layer][g][2],
                                          p.cen_neg16[layer][g][3])
                             : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
            for (int ch = 0; ch < 3; ++ch) {
              const int qc0 = chalf * 12 + ch * 4;              // this pass writes chunks qc0..qc0+3: one group
              const int grp = qc0 / kGroupChunks;
              const uint32_t cell_q0 = cell0 + qc0 * kChunkStride + grp * kGroupGap;
              const uint4 hv = grp == 0 ? hneg[0] : (grp == 1 ? hneg[1] : hneg[2]);
              uint32_t v[32];
              ptx::tmem_ld_x32(taddr + chalf * 96 + ch * 32, v);
              if (kCen && ch == 0) {
                // (under the latency of the first accumulator load) the tile's halo must read as "activation 0" for
                // the NEXT layer: -c of this layer's output (conv4's output: zeros -- conv5 uses the centre tap
                // only, its stencil scratch and the next supertile's conv1 need a zero halo)
                const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
                const uint32_t tile0 = a_base + h * kATileStride;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (i < 3 || halo_off[3] != 0xFFFFFFFFu) {
                    const uint32_t g = halo_off[i] & 3u;
                    ptx::st_shared_v4(tile0 + (halo_off[i] & ~15u),
                                      layer < 3 ? (g == 0 ? hneg[0] : (g == 1 ? hneg[1] : hneg[2])) : zero4);
                  }
              }
              ptx::tmem_ld_wait();
#pragma unroll
              for (int kc = 0; kc < 4; ++kc) {
                uint4 o;
                if (kCen) {
                  o.x = ptx::max16x2<kF16>(ptx::pack_sat<kF16>(__uint_as_float(v[kc * 8 + 0]),
                                                               __uint_as_float(v[kc * 8 + 1])), hv.x);
                  o.y = ptx::max16x2<kF16>(ptx::pack_sat<kF16>(__uint_as_float(v[kc * 8 + 2]),
                                                               __uint_as_float(v[kc * 8 + 3])), hv.y);
                  o.z = ptx::max16x2<kF16>(ptx::pack_sat<kF16>(__uint_as_float(v[kc * 8 + 4]),
                                                               __uint_as_float(v[kc * 8 + 5])), hv.z);
                  o.w = ptx::max16x2<kF16>(ptx::pack_sat<kF16>(__uint_as_float(v[kc * 8 + 6]),
                                                               __uint_as_float(v[kc * 8 + 7])), hv.w);
                } else {
                  o.x = ptx::pack_relu<kF16>(__uint_as_float(v[kc * 8 + 0]),
                                             __uint_as_float(v[kc * 8 + 1]));
                  o.y = ptx::pack_relu<kF16>(__uint_as_float(v[kc * 8 + 2]),
                                             __uint_as_float(v[kc * 8 + 3]));
                  o.z = ptx::pack_relu<kF16>(__uint_as_float(v[kc * 8 + 4]),
                                             __uint_as_float(v[kc * 8 + 5]));
                  o.w = ptx::pack_relu<kF16>(__uint_as_float(v[kc * 8 + 6]),
                                             __uint_as_float(v[kc * 8 + 7]));
                }
                ptx::st_shared_v4(cell_q0 + kc * kChunkStride, o);
              }
            }
            if (dbg != nullptr) {
              // debug / calibration dump (kept OUT of the loop above: a branch per store would split it into
              // basic blocks and serialise the bias loads): read this thread's 12 cells back, TF channel order
              for (int qc = chalf * 12; qc < chalf * 12 + 12; ++qc) {
                const uint4 o = ptx::ld_shared_v4(cell0 + chunk_off(qc));
                const uint32_t w[4] = {o.x, o.y, o.z, o.w};
                for (int i = 0; i < 8; ++i) {
                  const int q = qc * 8 + i;                        // channel position
                  const int c_tf = p.dbg_unperm != nullptr ? p.dbg_unperm[layer * kC + q] : q;
                  const float c = kCen ? p.cen[layer][8 * (qc / kGroupChunks) + i] : 0.f;
                  dbg[c_tf] = ptx::unpack16<kF16>((w[i >> 1] >> (16 * (i & 1))) & 0xFFFFu) + c;
                }
              }
            }
          } else if (chalf == 0) {
            // conv5: column t of the accumulator is P[pixel][tap t]; out = 9-point stencil sum.
            uint32_t v[16];
            ptx::tmem_ld_x16(taddr, v);
            if (kCen) {
              // columns 16..24: the taps' lo weights (x 4096); kappa: what the centring took out of a4
              uint32_t vl[16];
              ptx::tmem_ld_x16(taddr + 16, vl);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int t = 0; t < 9; ++t)
                v[t] = __float_as_uint(__uint_as_float(v[t]) + __uint_as_float(vl[t]) * (1.f / 4096.f) + p.kappa[t]);
            }
            ptx::tmem_ld_wait();
            // scratch = cells of channel-chunks 2..4 at this pixel (fp32 x 4 per cell); halo = 0
            ptx::st_shared_v4(cell0 + 2 * kChunkStride, make_uint4(v[0], v[1], v[2], v[3]));
            ptx::st_shared_v4(cell0 + 3 * kChunkStride, make_uint4(v[4], v[5], v[6], v[7]));
            ptx::st_shared_v4(cell0 + 4 * kChunkStride, make_uint4(v[8], 0u, 0u, 0u));
            asm volatile("bar.sync 2, 128;" ::: "memory");        // the 4 chalf==0 warps
            float acc = 0.f;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int dy = tap / 3 - 1, dx = tap % 3 - 1;
              acc += ptx::ld_shared_f32(cell0 + (2 + tap / 4) * kChunkStride + dy * kYStride +
                                        dx * kCell + (tap % 4) * 4);
            }
            const float hval = fmaxf(acc + bias_s[4 * kC], 0.f);
            if (p.fuse_tail) {
              // hand the half's 5th feature map to its tail warp (which took the previous
              // supertile's values long ago: a supertile is ~70 k cycles)
              const uint32_t tcount = lstep / 5u;
              if (!wait_bar(bar_h5_empty(h), (tcount & 1u) ^ 1u, false, p.err, 820 + h)) { ok = false; break; }
              h5s[(h * 2 + b) * 64 + y * 8 + x] = hval;
              ptx::mbar_arrive(bar_h5_full(h));
            }
            if (p.h5 != nullptr && board < n_tot) p.h5[board * 64 + y * 8 + x] = hval;
            if (st + npairs < total_st) store_input_cell<kCen>(cell0, next_in[h]);
          }
          signal_a_ready(h);
          if (profiler && lstep < kProfSteps) p.prof[(4 * h + 3) * kProfSteps + lstep] = clock64();
        }
      }
    }
  }

  else {
    // ===== fused tail warp of half h: FC-155 + bias -> softmax -> legal gather -> rank ==========
    const int h = warp - kFirstTailWarp;
    if (p.fuse_tail) {
      float* tb_s = reinterpret_cast<float*>(smem_gen + kOffTailBias);
      uint16_t* tmv_s = reinterpret_cast<uint16_t*>(smem_gen + kOffTailMoves);
      for (int i = tid - kFirstTailWarp * 32; i < 160; i += kNumTailWarps * 32) {
        tb_s[i] = i < kMoves ? p.tail.fc_b[i] : 0.f;
        int r = 0, c = 0, d = 0;
        if (i < 154) decode_move(i, r, c, d);
        tmv_s[i] = static_cast<uint16_t>((r * 8 + c) | (((r + 1) * 8 + c + d) << 6) | ((d == 0) << 12));
      }
      asm volatile("bar.sync 3, %0;" ::"n"(kNumTailWarps * 32) : "memory");
      if (h < nh) {
        float* wf = reinterpret_cast<float*>(smem_gen + kOffTailWarpF) + h * kTailWarpF;
        uint8_t* wb = smem_gen + kOffTailWarpB + h * kTailWarpB;
        float* hc = wf;                                   // [2][64]
        int8_t* pov = reinterpret_cast<int8_t*>(wb);      // [2][64]
        TailScratch sc = {hc, wf + 128, wf + 192, pov, wb + 128, wb + 192};
        const bool flagging = p.tail.refine_list != nullptr && p.tail.squares != nullptr;
        ptx::pdl_wait();
        uint32_t tcount = 0;
        for (long long st = pair; st < total_st; st += npairs, ++tcount) {
          const long long board0 = st * boards_per_pair + static_cast<int>(rank) * (2 * nh) + h * 2;
          const int nb = static_cast<int>(max(0ll, min(2ll, n_tot - board0)));
          const long long dst[2] = {board0, board0 + 1};
          if (p.tail.squares != nullptr) {
#pragma unroll
            for (int t = 0; t < 2; ++t)
              if (t < nb) {
                const int blk = p.tail.black[dst[t]];
                const int8_t* sq = p.tail.squares + dst[t] * 64;
                pov[t * 64 + lane] = static_cast<int8_t>(pov_square(sq, blk, lane));
                pov[t * 64 + lane + 32] = static_cast<int8_t>(pov_square(sq, blk, lane + 32));
              }
          }
          if (!wait_bar(bar_h5_full(h), tcount & 1u, false, p.err, 800 + h)) break;
          reinterpret_cast<float4*>(hc)[lane] = reinterpret_cast<const float4*>(h5s + h * 128)[lane];
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_h5_empty(h));
          if (nb > 0) tail_group<2, true>(p.tail, p.tail.fc_w, tb_s, tmv_s, sc, nb, dst, lane, flagging);
          __syncwarp();
        }
      }
    }
  }

  // ---- teardown ------------------------------------------------------------------------------
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 0) ptx::tmem_dealloc<2>(tmem_base, 512);
  if (p.err_mirror != nullptr && tid == 0) {
    // zero-copy calls: the LAST CTA to finish mirrors the watchdog record into mapped host memory
    __threadfence();
    if (atomicAdd(p.err + 6, 1) == static_cast<int>(gridDim.x) - 1) {
      __threadfence();
      for (int i = 0; i < 6; ++i) p.err_mirror[i] = reinterpret_cast<volatile int*>(p.err)[i];
      p.err[6] = 0;
    }
  }
}

}  // namespace wann
