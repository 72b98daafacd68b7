// Auxiliary sm_100a kernels of the WaNN leaf-evaluation path:
//   encode_kernel      board -> int8 POEB planes        (tools/utils.pyx:546-602,819-866)
//   legal_mask_kernel  board -> uint8[155] legality mask (board_utils.pyx:293-400, utils.pyx:605-641)
//   tail_kernel        h5[64] -> FC-155 + bias -> softmax -> (legal gather + descending sort)
//                      (policy_net_utils.pyx:215-239; tree_search_utils.pyx:979-986;
//                       tree_builder.pyx:557; board_utils.pyx:262-291)
//   conv3x3_f32_kernel / conv_small_f32_kernel / planes_to_f32_kernel
//                      CUDA-core fp32 conv stack for WANN_MODE_FP32 (accuracy mode)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wann {

constexpr int kMoves = 155;
constexpr int kMaxLegal = 48;

// side-to-move view of one square: 0 empty, 1 player, 2 opponent
__device__ __forceinline__ int pov_square(const int8_t* sq64, int blk, int i) {
  int s = sq64[blk ? 63 - i : i];
  if (blk && s) s = 3 - s;
  return s;
}

// One thread per FOUR squares: one 4-byte load (for Black the mirrored word, bytes reversed: the
// 180-degree rotation of tools/utils.pyx:563-600) and one 16-byte store of the four P/O/E/1 cells.
// HBM-bound: 65 B in, 256 B out per position.  Four grid-stride items per thread are in flight at once
// (loads first, then the stores): with one item per iteration the ~10 KB of reads a full SM keeps in
// flight did not cover the memory latency (4.5 TB/s; profiles/r02_byte_kernels.json has the new figure).
__device__ __forceinline__ uint4 encode_four_squares(uint32_t w, int blk) {
  if (blk) w = __byte_perm(w, 0u, 0x0123);            // reverse the four squares
  uint32_t out[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int sq = (w >> (8 * k)) & 0xFF;
    if (blk && sq) sq = 3 - sq;                       // Black sees its own pieces as "player"
    out[k] = (sq == 1 ? 1u : 0u) | (sq == 2 ? 0x100u : 0u) | (sq == 0 ? 0x10000u : 0u) | 0x1000000u;
  }
  return make_uint4(out[0], out[1], out[2], out[3]);
}
template <int kUnroll, bool kStream>
__global__ void __launch_bounds__(256) encode_kernel(const int8_t* __restrict__ squares,
                                                     const uint8_t* __restrict__ black, long long n,
                                                     int8_t* __restrict__ planes) {
  const long long total = n * 16;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i0 < total; i0 += kUnroll * stride) {
    uint32_t w[kUnroll];
    int blk[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i < total) {
        const long long board = i >> 4;
        blk[u] = __ldg(black + board);
        w[u] = __ldg(reinterpret_cast<const uint32_t*>(squares) + i);     // independent of the colour: for Black the
      }                                                                    // OUTPUT cell is the mirrored one
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i < total) {
        const long long o = blk[u] ? (i | 15) - (i & 15) : i;             // cell 15 - q of the same board
        if (kStream) __stcs(reinterpret_cast<uint4*>(planes) + o, encode_four_squares(w[u], blk[u]));
        else reinterpret_cast<uint4*>(planes)[o] = encode_four_squares(w[u], blk[u]);
      }
    }
  }
}

// decode a move index into its POV (row, col, dcol); idx in [0,154)
__device__ __forceinline__ void decode_move(int idx, int& r, int& c, int& d) {
  r = idx / 22;
  const int rem = idx - 22 * r;
  c = (rem + 1) / 3;
  d = rem - 3 * c;
}

__device__ __forceinline__ bool move_is_legal(const int8_t* pov, int idx) {
  if (idx >= 154) return false;   // 'no-move' is never a legal child
  int r, c, d;
  decode_move(idx, r, c, d);
  if (pov[r * 8 + c] != 1) return false;
  const int t = pov[(r + 1) * 8 + c + d];
  return d == 0 ? (t == 0) : (t != 1);
}

// One THREAD per board, bitboards in registers; all global traffic as coalesced 16-byte vectors:
//   * a warp loads its 32 boards (2 KB) with four 16-byte loads per lane into shared memory (80-byte pitch:
//     conflict-free both ways) and every lane reads its own board back as four 16-byte words;
//   * SWAR byte-equality tests (exact for any byte value) and a multiply-compress turn them into 64-bit
//     "white / black / empty" boards; Black's point of view is the bit-reversed board with the colours
//     swapped (the 180-degree rotation of tools/utils.pyx:563-600; a stray value 3 reads as empty for
//     Black, as in pov_square);
//   * legality for all 154 moves is three shifts: forward = player & (empty >> 8), the diagonals
//     = player & ~(player >> 7 | 9) off the edge files (board_utils.pyx:331-400);
//   * the 155 mask bytes of a board are built IN REGISTERS as 39 packed words: per rank the three 8-bit
//     rows (left / forward / right) are spread to bytes with a multiply (bit i -> byte i), interleaved to
//     the move alphabet's order (22 indices per rank: F0 R0 | L1 F1 R1 | ... | L7 F7,
//     tools/utils.pyx:605-641) with byte permutes, and shifted to the row's byte offset; the row is then
//     shifted once more by its own alignment (155 * lane mod 4) and written with WORD stores into the warp's
//     packed 4,960-byte block (a word shared by two rows is written by the lower row, which fetches the
//     upper row's bytes with a shuffle), and the block leaves as 16-byte vectors.
//   The first version wrote the 155 bytes of a board with 155 byte stores per lane: shared-memory store
//   bound (2.8 TB/s = 0.43 of the HBM copy peak).  65 B in, 155 B out per position.
constexpr int kMaskWarps = 4;
constexpr int kMaskInPitch = 80;
__device__ __forceinline__ uint64_t swar_eq_flags(uint64_t w, uint64_t pattern) {   // 0x80 in every byte == pattern's
  const uint64_t x = w ^ pattern;
  const uint64_t t = (x & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full;
  return ~(t | x | 0x7F7F7F7F7F7F7F7Full);
}
__device__ __forceinline__ uint64_t swar_compress(uint64_t flags) {   // byte c's 0x80 flag -> bit c
  return ((flags >> 7) * 0x0102040810204080ull) >> 56;
}
// bit i (i < 4) of x -> byte i (0 / 1)
__device__ __forceinline__ uint32_t spread4(uint32_t x) { return ((x & 0xFu) * 0x00204081u) & 0x01010101u; }

__global__ void __launch_bounds__(kMaskWarps * 32) legal_mask_kernel(const int8_t* __restrict__ squares,
                                                                    const uint8_t* __restrict__ black,
                                                                    long long n, uint8_t* __restrict__ mask) {
  __shared__ __align__(16) uint8_t in_s[kMaskWarps][32 * kMaskInPitch];
  __shared__ __align__(16) uint32_t out_s[kMaskWarps][32 * kMoves / 4];      // 1,240 words
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long base = (blockIdx.x * static_cast<long long>(kMaskWarps) + w) * 32; base < n;
       base += static_cast<long long>(gridDim.x) * kMaskWarps * 32) {
    const int nvalid = static_cast<int>(min(32ll, n - base));
    const long long board = base + lane;
    // ---- the warp's boards: global -> shared, coalesced --------------------------------------------
    const uint4* src = reinterpret_cast<const uint4*>(squares + base * 64);
    uint4 chunk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = j * 32 + lane;
      chunk[j] = q < nvalid * 4 ? __ldg(src + q) : make_uint4(0u, 0u, 0u, 0u);
    }
    const bool blk = board < n && __ldg(black + board) != 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = j * 32 + lane;
      *reinterpret_cast<uint4*>(&in_s[w][kMaskInPitch * (q >> 2) + 16 * (q & 3)]) = chunk[j];
    }
    __syncwarp();
    // 4 squares per 32-bit word: a byte of y1 is zero where the mover's piece stands, a byte of y2 where the
    // square is empty (for Black: 0, or a stray 3 -- pov_square reads 3 - 3 = 0); zero bytes -> 0x80 flags
    // (exact for any byte value), 4 flags -> 4 bits with one multiply
    const uint32_t pcode = blk ? 0x02020202u : 0x01010101u;
    uint32_t pl[2] = {0u, 0u}, em[2] = {0u, 0u};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint4 v = *reinterpret_cast<const uint4*>(&in_s[w][kMaskInPitch * lane + 16 * k]);
      const uint32_t xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int wi = 4 * k + t;
        const uint32_t x = xs[t];
        const uint32_t y1 = x ^ pcode;
        const uint32_t y2 = blk ? (((x ^ (x >> 1)) & 0x01010101u) | (x & 0xFCFCFCFCu)) : x;
        const uint32_t z1 = ~(((y1 & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y1) & 0x80808080u;
        const uint32_t z2 = ~(((y2 & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y2) & 0x80808080u;
        pl[wi >> 3] |= (((z1 >> 7) * 0x01020408u) >> 24) << (4 * (wi & 7));
        em[wi >> 3] |= (((z2 >> 7) * 0x01020408u) >> 24) << (4 * (wi & 7));
      }
    }
    const uint64_t pl64 = static_cast<uint64_t>(pl[0]) | (static_cast<uint64_t>(pl[1]) << 32);
    const uint64_t em64 = static_cast<uint64_t>(em[0]) | (static_cast<uint64_t>(em[1]) << 32);
    const uint64_t player = blk ? __brevll(pl64) : pl64;
    const uint64_t empty = blk ? __brevll(em64) : em64;
    constexpr uint64_t kRows06 = 0x00FFFFFFFFFFFFFFull, kFileA = 0x0101010101010101ull;
    uint64_t fwd = player & (empty >> 8) & kRows06;
    uint64_t lft = player & ~(player >> 7) & ~kFileA & kRows06;          // d = -1: to = from + 7
    uint64_t rgt = player & ~(player >> 9) & ~(kFileA << 7) & kRows06;   // d = +1: to = from + 9
    if (board >= n) fwd = lft = rgt = 0;
    // ---- the row of 155 bytes as 39 packed words (byte 155 = pad) -----------------------------------
    // per rank: T[r] = the 24 bytes L0 F0 G0 L1 F1 G1 ... L7 F7 G7; L0 and G7 are always 0, bytes 1..22 are
    // the rank's 22 row bytes
    uint32_t T[7][6];
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      const uint32_t f = static_cast<uint32_t>(fwd >> (8 * r)) & 0xFFu;
      const uint32_t l = static_cast<uint32_t>(lft >> (8 * r)) & 0xFFu;
      const uint32_t g = static_cast<uint32_t>(rgt >> (8 * r)) & 0xFFu;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const uint32_t L = spread4(l >> (4 * hh)), F = spread4(f >> (4 * hh)), G = spread4(g >> (4 * hh));
        T[r][3 * hh + 0] = __byte_perm(__byte_perm(L, F, 0x0140), G, 0x2410);      // L0 F0 G0 L1
        T[r][3 * hh + 1] = __byte_perm(__byte_perm(F, G, 0x0251), L, 0x2610);      // F1 G1 L2 F2
        T[r][3 * hh + 2] = __byte_perm(__byte_perm(G, L, 0x0372), F, 0x2710);      // G2 L3 F3 G3
      }
    }
    // row word j = row bytes 4j .. 4j+3: a funnel shift inside a rank's 24 bytes, OR (where the word crosses
    // into the next rank) that rank's first word shifted -- the always-zero L0 / G7 bytes make the ORs safe
    uint32_t U[40];
#pragma unroll
    for (int j = 0; j < 39; ++j) {
      const int b0 = 4 * j, r = b0 / 22, m0 = b0 % 22 + 1, a = m0 / 4, sh = m0 % 4;
      uint32_t v = 0u;
      if (r < 7) {
        const uint32_t lo = T[r][a], hi = a + 1 < 6 ? T[r][a + 1] : 0u;
        v = sh == 0 ? lo : __funnelshift_r(lo, hi, 8 * sh);
        const int cnt = 23 - m0 < 4 ? 23 - m0 : 4;            // bytes of this word that belong to rank r
        if (cnt < 4 && r + 1 < 7) v |= T[r + 1][0] << (8 * (cnt - 1));
      }
      U[j] = v;
    }
    U[39] = 0u;
    U[38] &= 0x00FFFFFFu;                             // (byte 155 is the pad; byte 154 = 'no-move' = 0 already)
    // ---- into the warp's packed block, word stores ---------------------------------------------------
    const int s_me = (kMoves * lane) & 3, w0 = (kMoves * lane) >> 2;
    const int s_next = (kMoves * (lane + 1)) & 3;
    const int j_last = (s_me + kMoves - 1) >> 2;      // 38 or 39
    const uint32_t head = U[0] << (8 * s_me);         // my first bytes, in the position they have in the shared word
    uint32_t next_head = __shfl_down_sync(0xffffffffu, head, 1);
    if (lane == 31 || s_next == 0) next_head = 0u;
#pragma unroll
    for (int j = 0; j < 40; ++j) {
      const uint32_t lo = j > 0 ? U[j - 1] : 0u, hi = j < 39 ? U[j] : 0u;
      uint32_t v = __funnelshift_l(lo, hi, 8 * s_me);
      if (j == j_last) v |= next_head;
      if (j <= j_last && (j > 0 || s_me == 0)) out_s[w][w0 + j] = v;
    }
    __syncwarp();
    // ---- shared -> global, 16-byte vectors ------------------------------------------------------------
    const int bytes = nvalid * kMoves;
    uint8_t* dst = mask + base * kMoves;              // base is a multiple of 32 -> 16-byte aligned if mask is
    const uint8_t* srcb = reinterpret_cast<const uint8_t*>(out_s[w]);
    for (int i = lane; i < bytes / 16; i += 32)
      __stcs(reinterpret_cast<uint4*>(dst) + i, reinterpret_cast<const uint4*>(srcb)[i]);
    for (int i = (bytes & ~15) + lane; i < bytes; i += 32) dst[i] = srcb[i];
    __syncwarp();
  }
}

struct TailParams {
  const float* h5;          // [n][64]
  const int8_t* squares;    // [n][64] or null (no legality -> no top-k)
  const uint8_t* black;     // [n]
  long long n;
  const float* fc_w;        // [64][155]
  const float* fc_b;        // [155]
  float* probs;             // [n][155] or null
  float* logits;            // [n][155] or null
  uint8_t* idx;             // [n][48] or null
  float* prior;             // [n][48]
  uint8_t* n_legal;         // [n]
  int top_k = kMaxLegal;    // keep the best top_k legal children per position (num_top / num_to_keep of the reference)
  // margin refinement (wann_set_option "refine_margin_milli"):
  //   flagging pass: boards whose two best LEGAL priors are closer than the margin
  //                  (p2 >= p1 * refine_ratio, refine_ratio = exp(-margin)) are appended to
  //                  refine_list, counted in refine_count[0] (this chunk) and [1] (running total);
  //   refinement pass: h5 holds COMPACT rows for the boards gather[0 .. *n_dev); board row
  //                  gather[i] of squares / black / every output is (re)written from h5 row i.
  int* refine_list = nullptr;
  int* refine_count = nullptr;
  float refine_ratio = 0.f;
  const int* gather = nullptr;
  const int* n_dev = nullptr;      // device-side position count: min(n, max(0, *n_dev - n_dev_off)) positions
  long long n_dev_off = 0;
  // zero-copy small calls: the slot's device-side watchdog / refinement record is mirrored into
  // host-mapped memory by the last kernel of the call (saves a D2H copy on the latency path)
  const int* err_src = nullptr;
  int* err_dst = nullptr;
};

// One warp handles 4 boards per pass (each FC weight read from shared memory feeds 4 FMAs);
// softmax by warp shuffles; legality from a per-CTA move table; ranking = compact the <= 48
// legal candidates (ascending move index), then each lane counts how many candidates beat its
// own (prior greater, or equal with a smaller index) -> its output slot.
constexpr int kTailWarps = 8;
constexpr int kTailBoards = 4;                 // boards per warp per pass
constexpr int kTailWarpFloats = kTailBoards * 64 + 64 + kMaxLegal;   // h[4][64] + compact priors[48] (+pad) + ranked priors[48]
constexpr int kTailWarpBytes = kTailBoards * 64 + 64 + kMaxLegal;    // pov[4][64] + compact indices[48] (+pad) + ranked indices[48]
constexpr int kTailSmemBytes = (64 * kMoves + 160) * 4 + 160 * 2 +
                               kTailWarps * (kTailWarpFloats * 4 + kTailWarpBytes);

// The per-group body shared by tail_kernel (FC weights in shared memory, 4 boards per warp) and the
// fused tail warps of conv_stack_tc_kernel (FC weights read through L1/L2, 2 boards per warp):
// logits -> softmax -> outputs -> legal gather -> rank -> (margin flag).  h / pov hold the group's
// 5th feature maps and side-to-move boards; dst[t] is the output row of board t.
struct TailScratch {
  const float* h;      // [NB][64]
  float* cp;           // [64]  compact priors (+ the two best at [48], [49])
  float* rp;           // [48]  ranked priors, 16-byte aligned
  const int8_t* pov;   // [NB][64]
  uint8_t* cj;         // [64]  compact indices
  uint8_t* ri;         // [48]  ranked indices, 16-byte aligned
};

template <int NB, bool kGlobalW>
__device__ __forceinline__ void tail_group(const TailParams& p, const float* __restrict__ w,
                                           const float* __restrict__ b_s, const uint16_t* __restrict__ mv_s,
                                           const TailScratch& sc, int nb, const long long (&dst)[NB], int lane,
                                           bool flagging) {
  const float* h = sc.h;
  float* cp = sc.cp;
  float* rp = sc.rp;
  const int8_t* pov = sc.pov;
  uint8_t* cj = sc.cj;
  uint8_t* ri = sc.ri;
  // logits[t][j] = sum_q h[t][q] * W[q][j] + b[j]   (reshape [-1,64] row-major, matmul, bias_add)
  float lg[NB][5];
#pragma unroll
  for (int t = 0; t < NB; ++t)
#pragma unroll
    for (int k = 0; k < 5; ++k) lg[t][k] = 0.f;
#pragma unroll 8
  for (int q = 0; q < 64; ++q) {
    const float* wr = w + q * kMoves + lane;
    float wv[5];
#pragma unroll
    for (int k = 0; k < 4; ++k) wv[k] = kGlobalW ? __ldg(wr + 32 * k) : wr[32 * k];
    wv[4] = (lane < kMoves - 128) ? (kGlobalW ? __ldg(wr + 128) : wr[128]) : 0.f;
#pragma unroll
    for (int t = 0; t < NB; ++t) {
      const float hq = h[t * 64 + q];
#pragma unroll
      for (int k = 0; k < 5; ++k) lg[t][k] = fmaf(hq, wv[k], lg[t][k]);
    }
  }
#pragma unroll
  for (int t = 0; t < NB; ++t) {
    if (t >= nb) break;
    const long long board = dst[t];
    float e[5];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int j = lane + 32 * k;
      if (j < kMoves) {
        e[k] = lg[t][k] + b_s[j];
        mx = fmaxf(mx, e[k]);
      } else {
        e[k] = -INFINITY;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f, pr[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      pr[k] = (lane + 32 * k < kMoves) ? expf(e[k] - mx) : 0.f;
      sum += pr[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int j = lane + 32 * k;
      pr[k] = pr[k] / sum;
      if (j < kMoves) {
        if (p.probs != nullptr) p.probs[board * kMoves + j] = pr[k];
        if (p.logits != nullptr) p.logits[board * kMoves + j] = e[k];
      }
    }
    if (p.idx != nullptr || flagging) {
      // legal gather: prior_i = softmax155[idx_i] (no renormalisation); compact in ascending
      // index order, then rank: descending by prior, ties by ascending index.
      const int8_t* pv = pov + t * 64;
      int cnt = 0;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int j = lane + 32 * k;
        bool legal = false;
        if (j < 154) {
          const uint32_t m = mv_s[j];
          const int tgt = pv[(m >> 6) & 63];
          legal = pv[m & 63] == 1 && ((m >> 12) ? tgt == 0 : tgt != 1);
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, legal);
        if (legal) {
          const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
          if (pos < kMaxLegal) {   // a real position has <= 48 legal moves (16 pieces x 3);
            cp[pos] = pr[k];       // malformed boards are truncated to the 48 lowest indices
            cj[pos] = static_cast<uint8_t>(j);
          }
        }
        cnt += __popc(mask);
      }
      cnt = min(cnt, kMaxLegal);
      __syncwarp();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c = lane + 32 * half;
        if (c < kMaxLegal) {
          if (c < cnt) {
            const float me = cp[c];
            int rank = 0;
            for (int o = 0; o < cnt; ++o) {
              const float other = cp[o];
              rank += (other > me) || (other == me && o < c);
            }
            const bool keep = rank < p.top_k;       // pruned children read as padding
            ri[rank] = keep ? cj[c] : static_cast<uint8_t>(255);
            rp[rank] = keep ? me : 0.f;
            if (flagging && rank < 2) cp[kMaxLegal + rank] = me;   // the two best legal priors
          } else {
            ri[c] = 255;
            rp[c] = 0.f;
          }
        }
      }
      __syncwarp();
      if (p.idx != nullptr) {   // one ranked row = 12 + 3 sixteen-byte stores (+ the count)
        if (lane < 12)
          reinterpret_cast<float4*>(p.prior + board * kMaxLegal)[lane] = reinterpret_cast<const float4*>(rp)[lane];
        else if (lane < 15)
          reinterpret_cast<uint4*>(p.idx + board * kMaxLegal)[lane - 12] = reinterpret_cast<const uint4*>(ri)[lane - 12];
        else if (lane == 15)
          p.n_legal[board] = static_cast<uint8_t>(min(cnt, p.top_k));
      }
      if (flagging && lane == 0 && cnt >= 2 && cp[kMaxLegal + 1] >= cp[kMaxLegal] * p.refine_ratio) {
        p.refine_list[atomicAdd(p.refine_count, 1)] = static_cast<int>(board);
        atomicAdd(p.refine_count + 1, 1);
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(kTailWarps * 32) tail_kernel(const TailParams p) {
  extern __shared__ float tsm[];
  // refinement pass: the board count lives on the device (this launch is never made under PDL, so
  // the flagging pass has completed); CTAs without work leave before touching the FC weights
  const long long n_eff = (p.n_dev != nullptr) ? max(0ll, min(p.n, static_cast<long long>(*p.n_dev) - p.n_dev_off)) : p.n;
  if (p.err_dst != nullptr && blockIdx.x == 0 && threadIdx.x < 8) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    p.err_dst[threadIdx.x] = p.err_src[threadIdx.x];
  }
  if (blockIdx.x * static_cast<long long>(kTailWarps * kTailBoards) >= n_eff) return;
  float* w_s = tsm;                                   // [64][155]
  float* b_s = w_s + 64 * kMoves;                     // [160]
  uint16_t* mv_s = reinterpret_cast<uint16_t*>(b_s + 160);   // [160] from | to<<6 | fwd<<12
  float* warp_f = reinterpret_cast<float*>(mv_s + 160);
  uint8_t* warp_b = reinterpret_cast<uint8_t*>(warp_f + kTailWarps * kTailWarpFloats);
  for (int i = threadIdx.x; i < 64 * kMoves; i += blockDim.x) w_s[i] = p.fc_w[i];
  for (int i = threadIdx.x; i < 160; i += blockDim.x) {
    b_s[i] = i < kMoves ? p.fc_b[i] : 0.f;
    int r = 0, c = 0, d = 0;
    if (i < 154) decode_move(i, r, c, d);
    mv_s[i] = static_cast<uint16_t>((r * 8 + c) | (((r + 1) * 8 + c + d) << 6) | ((d == 0) << 12));
  }
  // PDL: the FC weights above are constants; h5 (and the boards) are only valid once the
  // conv-stack kernel before us has completed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* h = warp_f + w * kTailWarpFloats;            // [4][64]
  float* cp = h + kTailBoards * 64;                   // [48] compact priors
  float* rp = cp + 64;                                // [48] ranked priors (16-byte aligned)
  int8_t* pov = reinterpret_cast<int8_t*>(warp_b + w * kTailWarpBytes);            // [4][64]
  uint8_t* cj = reinterpret_cast<uint8_t*>(pov + kTailBoards * 64);                // [48] compact indices
  uint8_t* ri = cj + 64;                                                           // [48] ranked indices (16-byte aligned)
  const bool flagging = p.refine_list != nullptr && p.squares != nullptr;
  const long long ngroups = (n_eff + kTailBoards - 1) / kTailBoards;
  for (long long g = blockIdx.x * static_cast<long long>(kTailWarps) + w; g < ngroups;
       g += static_cast<long long>(gridDim.x) * kTailWarps) {
    const long long board0 = g * kTailBoards;
    const int nb = static_cast<int>(min(static_cast<long long>(kTailBoards), n_eff - board0));
    long long dst[kTailBoards];     // row of squares / black / outputs that h5 row board0 + t belongs to
#pragma unroll
    for (int t = 0; t < kTailBoards; ++t)
      dst[t] = (p.gather != nullptr && t < nb) ? static_cast<long long>(p.gather[board0 + t]) : board0 + t;
#pragma unroll
    for (int t = 0; t < kTailBoards * 2; ++t) {        // 4 boards x 64 floats, coalesced
      const int i = t * 32 + lane;
      h[i] = (i >> 6) < nb ? p.h5[board0 * 64 + i] : 0.f;
    }
    if (p.squares != nullptr) {
#pragma unroll
      for (int t = 0; t < kTailBoards; ++t)
        if (t < nb) {
          const int blk = p.black[dst[t]];
          const int8_t* sq = p.squares + dst[t] * 64;
          pov[t * 64 + lane] = static_cast<int8_t>(pov_square(sq, blk, lane));
          pov[t * 64 + lane + 32] = static_cast<int8_t>(pov_square(sq, blk, lane + 32));
        }
    }
    __syncwarp();
    TailScratch sc = {h, cp, rp, pov, cj, ri};
    tail_group<kTailBoards, false>(p, w_s, b_s, mv_s, sc, nb, dst, lane, flagging);
    __syncwarp();
  }
}

// Child construction for every ranked legal move of every position:
// get_child_attributes (tree_builder.pyx:641-658) + move_piece_update_piece_arrays
// (board_utils.pyx:90-131) + check_for_winning_move (tree_builder.pyx:902-905).
// One warp per board; 4 lanes build one 64-byte child (16 B each), 8 children per pass, so every
// store is a coalesced 16-byte vector.  flags: bit0 = move reaches the far rank (index > 131),
// bit1 = capture.  Pad slots (>= n_legal) get zero boards / zero flags.
__global__ void __launch_bounds__(256) expand_children_kernel(const int8_t* __restrict__ squares,
                                                              const uint8_t* __restrict__ black,
                                                              const uint8_t* __restrict__ idx,
                                                              const uint8_t* __restrict__ n_legal,
                                                              long long n, int8_t* __restrict__ children,
                                                              uint8_t* __restrict__ flags) {
  __shared__ __align__(16) int8_t brd[8][64];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  for (long long board = blockIdx.x * 8ll + w; board < n; board += gridDim.x * 8ll) {
    reinterpret_cast<uint16_t*>(brd[w])[lane] = reinterpret_cast<const uint16_t*>(squares + board * 64)[lane];
    const int blk = black[board];
    const int cnt = n_legal[board];
    __syncwarp();
    const uint4 base = reinterpret_cast<const uint4*>(brd[w])[q];
#pragma unroll 1
    for (int k0 = 0; k0 < kMaxLegal; k0 += 8) {
      const int k = k0 + g;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      uint32_t f = 0;
      if (k < cnt) {
        const int j = idx[board * kMaxLegal + k];
        int r, c, d;
        decode_move(j, r, c, d);
        int frm = r * 8 + c, to = (r + 1) * 8 + c + d;
        if (blk) { frm = 63 - frm; to = 63 - to; }
        const uint32_t mover = static_cast<uint32_t>(brd[w][frm]) & 0xFFu;
        f = (j > 131 ? 1u : 0u) | (brd[w][to] != 0 ? 2u : 0u);
        uint32_t wd[4] = {base.x, base.y, base.z, base.w};
        if ((frm >> 4) == q) wd[(frm >> 2) & 3] &= ~(0xFFu << ((frm & 3) * 8));
        if ((to >> 4) == q) {
          const int sh = (to & 3) * 8;
          wd[(to >> 2) & 3] = (wd[(to >> 2) & 3] & ~(0xFFu << sh)) | (mover << sh);
        }
        v = make_uint4(wd[0], wd[1], wd[2], wd[3]);
      }
      reinterpret_cast<uint4*>(children + (board * kMaxLegal + k) * 64)[q] = v;
      if (q == 0) flags[board * kMaxLegal + k] = static_cast<uint8_t>(f);
    }
    __syncwarp();
  }
}

// Child seeding (SURVEY 8(a) a10 + 8(f)-1): what the reference does with the priors right after the
// evaluation, for every ranked legal child of every position --
//   check_for_winning_move / set_game_over_values           tree_builder.pyx:902-914
//   update_child (capture close to home, prior -> seeds)     tree_search_utils.pyx:963-1014
//   assign_children: "maybe gameover next move" branch       tree_builder.pyx:563-576 (game-saving
//     captures, set_game_over_values_1_step_lookahead :916-921), else eval_children with
//     evaluation_function                                     tree_search_utils.pyx:812-858,1056-1071
//   one level of update_tree_wins / _losses                  tree_search_utils.pyx:43-66
//   update_win_status_from_children                          tree_search_utils.pyx:86-132
// One warp per board; lane c seeds children c and c + 32; the parent summary is built from warp
// ballots.  stats: int32 [n][48][4] = visits, wins, gameover_visits, gameover_wins (one 16-byte
// store per child); flags: uint8 [n][48] (bit0 win, bit1 capture, bit2 capture close to home,
// bit3 game-saving move, bit4 doomed = loses to a 1-step lookahead, bit5 evaluation_function == 1);
// parent: int32 [n][8] = visits, wins, gameover_visits, gameover_wins gained by the parent, its
// win_status (-1 None / 0 False / 1 True), threat, number of winning children, of doomed children.
__constant__ int8_t kEvalTable[64] = {5, 28, 28, 12, 12, 28, 28, 5,   2, 3, 3, 3, 3, 3, 3, 2,
                                      3, 5, 10, 10, 10, 10, 5, 3,     6, 9, 16, 16, 16, 16, 9, 6,
                                      9, 15, 21, 21, 21, 21, 15, 9,   14, 22, 22, 22, 22, 22, 22, 14,
                                      21, 23, 23, 23, 23, 23, 23, 21, 24, 24, 24, 24, 24, 24, 24, 24};

__global__ void __launch_bounds__(256) seed_children_kernel(const int8_t* __restrict__ squares,
                                                            const uint8_t* __restrict__ black,
                                                            const uint8_t* __restrict__ idx,
                                                            const float* __restrict__ prior,
                                                            const uint8_t* __restrict__ n_legal, long long n,
                                                            int32_t* __restrict__ stats, uint8_t* __restrict__ flags,
                                                            int32_t* __restrict__ parent,
                                                            const int* __restrict__ n_dev = nullptr,
                                                            long long n_dev_off = 0) {
  if (n_dev != nullptr) n = max(0ll, min(n, static_cast<long long>(*n_dev) - n_dev_off));
  __shared__ int8_t brd[8][64];
  __shared__ int8_t tab[64];        // lanes index the table divergently: shared memory, not the constant cache
  if (threadIdx.x < 64) tab[threadIdx.x] = kEvalTable[threadIdx.x];
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long board = blockIdx.x * 8ll + w; board < n; board += gridDim.x * 8ll) {
    const int blk = black[board];
    const int cnt = min(static_cast<int>(n_legal[board]), kMaxLegal);
    const int s0 = squares[board * 64 + lane], s1 = squares[board * 64 + lane + 32];
    brd[w][lane] = static_cast<int8_t>(s0);
    brd[w][lane + 32] = static_cast<int8_t>(s1);
    // evaluation_function of the PARENT board: sum over white - sum over black, absolute squares
    int res = (s0 == 1 ? tab[lane] : (s0 == 2 ? -tab[lane] : 0)) +
              (s1 == 1 ? tab[lane + 32] : (s1 == 2 ? -tab[lane + 32] : 0));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) res += __shfl_xor_sync(0xffffffffu, res, o);
    // threat: an opponent piece one step from the mover's home rank (rank 2 for White, rank 7 for Black)
    const int opp = blk ? 1 : 2;
    const bool on_row = blk ? (lane + 32 >= 48 && lane + 32 < 56 && s1 == opp) : (lane >= 8 && lane < 16 && s0 == opp);
    const bool threat = __any_sync(0xffffffffu, on_row);
    __syncwarp();
    int n_win = 0, n_doomed = 0, n_l = 0, n_w = 0;
    bool any_false = false, any_true = false, any_none = false;        // all children
    bool c_any = false, c_true = false, c_none = false;                // children over the 1.001 threshold
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int k = lane + 32 * half;
      const bool live = k < cnt;
      int v = 0, wn = 0, gv = 0, gw = 0, f = 0, ws = -1;
      bool cons = false;
      if (live) {
        const int j = idx[board * kMaxLegal + k];
        const float p = prior[board * kMaxLegal + k];
        int r, c, d;
        decode_move(j, r, c, d);
        int frm = r * 8 + c, to = (r + 1) * 8 + c + d;
        if (blk) { frm = 63 - frm; to = 63 - to; }
        const bool capture = brd[w][to] == opp;
        const bool win = j > 131;
        f = (win ? 1 : 0) | (capture ? 2 : 0);
        double uct = 1.0 + static_cast<double>(p);
        if (win) { wn = 0; gv = 65537; v = 65537; ws = 0; }
        if (j >= 23 && j <= 42 && capture) {
          gv = 100; gw = 0; v = 100; wn = 0;
          uct = static_cast<double>(__fadd_rn(1.0f, p));
          f |= 4;
        } else if (!win) {
          const int nv = static_cast<int>(__fmul_rn(p, 100.0f));
          v = 100;
          const int ww = v - nv;
          wn = min(ww, 75);
          if (ww < 70) { gv = v; gw = ww; }
        }
        if (threat) {
          if (capture) { gv = 1000; gw = 0; v = 65536; wn = 0; f |= 8; }
          else if (!win) { v = wn = gw = gv = 65537; ws = 1; f |= 16; }
        } else {
          const int t_to = tab[to], t_frm = tab[frm];
          int rc = blk ? res - (t_to - t_frm) : res + (t_to - t_frm);
          if (capture) rc += blk ? -t_to : t_to;
          const bool child_is_white = blk != 0;
          if ((rc > 0) == child_is_white) f |= 32;
        }
        cons = uct > static_cast<double>(1.001f);
        reinterpret_cast<int4*>(stats)[board * kMaxLegal + k] = make_int4(v, wn, gv, gw);
      } else if (k < kMaxLegal) {
        reinterpret_cast<int4*>(stats)[board * kMaxLegal + k] = make_int4(0, 0, 0, 0);
      }
      if (k < kMaxLegal) flags[board * kMaxLegal + k] = static_cast<uint8_t>(f);
      n_win += __popc(__ballot_sync(0xffffffffu, live && (f & 1)));
      n_doomed += __popc(__ballot_sync(0xffffffffu, live && (f & 16)));
      n_l += __popc(__ballot_sync(0xffffffffu, live && !threat && (f & 32)));
      n_w += __popc(__ballot_sync(0xffffffffu, live && !threat && !(f & 32)));
      any_false |= __any_sync(0xffffffffu, live && ws == 0);
      any_true |= __any_sync(0xffffffffu, live && ws == 1);
      any_none |= __any_sync(0xffffffffu, live && ws == -1);
      c_any |= __any_sync(0xffffffffu, live && cons);
      c_true |= __any_sync(0xffffffffu, live && cons && ws == 1);
      c_none |= __any_sync(0xffffffffu, live && cons && ws == -1);
    }
    if (lane == 0) {
      int pv = n_win + n_doomed + n_l + n_w, pw = n_win + n_w, pgv = n_win + n_doomed, pgw = n_win, pws = -1;
      if (cnt > 0) {
        const bool use_c = !any_false && c_any;          // tree_search_utils.pyx:113-114
        const bool has_true = use_c ? c_true : any_true, has_none = use_c ? c_none : any_none;
        if (any_false) pws = 1;
        else if (has_true && !has_none) { pws = 0; pv += 1; pgv += 1; }
      }
      int4* out = reinterpret_cast<int4*>(parent + board * 8);
      out[0] = make_int4(pv, pw, pgv, pgw);
      out[1] = make_int4(pws, threat ? 1 : 0, n_win, n_doomed);
    }
    __syncwarp();
  }
}

// Self-play style step (BASELINE config 5: independent game trees, one leaf per tree per step):
// pick ONE child per game from its ranked legal priors -- greedy (slot 0 = the policy move,
// board_utils.pyx:262-277) or sampled in proportion to the priors -- apply the move in place
// (board_utils.pyx:119-120), flip the side to move, and restart a finished game (move index > 131,
// tree_builder.pyx:902-905, or no legal move) from the opening position (utils.pyx:518-541).
// Counter-based RNG: u = splitmix64(seed, step, game) -> 24-bit uniform; reproducible on the host.
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256) tree_step_kernel(int8_t* __restrict__ squares, uint8_t* __restrict__ black,
                                                        const uint8_t* __restrict__ idx,
                                                        const float* __restrict__ prior,
                                                        const uint8_t* __restrict__ n_legal, long long n,
                                                        long long game0, int greedy, uint64_t seed, uint64_t step,
                                                        uint8_t* __restrict__ chosen, uint8_t* __restrict__ finished) {
  // (no griddepcontrol.launch_dependents here: letting the next evaluation's 148 CTAs start early and
  // spin next to this small kernel was measured slower, 106 -> 114 us per 1,024-game step)
  const long long g = blockIdx.x * 256ll + threadIdx.x;
  if (g >= n) return;
  const int cnt = n_legal[g];
  int k = 0;
  if (!greedy && cnt > 1) {
    const float* pr = prior + g * kMaxLegal;
    float total = 0.f;
    for (int i = 0; i < cnt; ++i) total += pr[i];
    const uint64_t r = splitmix64(splitmix64(seed ^ (step * 0xD1342543DE82EF95ull)) ^ static_cast<uint64_t>(game0 + g));
    const float u = static_cast<float>(r >> 40) * (1.0f / 16777216.0f) * total;
    float cum = 0.f;
    k = cnt - 1;
    for (int i = 0; i < cnt; ++i) {
      cum += pr[i];
      if (u < cum) { k = i; break; }
    }
  }
  uint8_t mv = 255, fin = 1;
  int8_t* b = squares + g * 64;
  if (cnt > 0) {
    mv = idx[g * kMaxLegal + k];
    fin = mv > 131;
  }
  if (fin) {   // game over (or stuck): restart from the opening, White to move
    uint32_t* w = reinterpret_cast<uint32_t*>(b);
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = i < 4 ? 0x01010101u : (i >= 12 ? 0x02020202u : 0u);
    black[g] = 0;
  } else {
    int r, c, d;
    decode_move(mv, r, c, d);
    int frm = r * 8 + c, to = (r + 1) * 8 + c + d;
    if (black[g]) { frm = 63 - frm; to = 63 - to; }
    b[to] = b[frm];
    b[frm] = 0;
    black[g] ^= 1;
  }
  if (chosen != nullptr) chosen[g] = mv;
  if (finished != nullptr) finished[g] = fin;
}

// --------------------------------------------------------------------------------------------
// fp32 accuracy mode (CUDA cores).  One CTA per board, 256 threads: warp = output row y,
// lane -> output channels lane + 32*j.  Input board staged zero-padded in shared memory.
// tf.nn.conv2d NHWC/HWIO SAME stride 1 + bias_add + relu (policy_net_utils.pyx:169-176).
__global__ void planes_to_f32_kernel(const int8_t* __restrict__ squares,
                                     const uint8_t* __restrict__ black,
                                     const void* __restrict__ planes_in, int planes_f32,
                                     long long n, float* __restrict__ out) {
  const long long total = n * 64;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 v;
    if (squares != nullptr) {
      const long long board = i >> 6;
      const int s = pov_square(squares + board * 64, black[board], static_cast<int>(i & 63));
      v = make_float4(s == 1, s == 2, s == 0, 1.f);
    } else if (planes_f32) {
      v = reinterpret_cast<const float4*>(planes_in)[i];
    } else {
      const char4 c = reinterpret_cast<const char4*>(planes_in)[i];
      v = make_float4(c.x, c.y, c.z, c.w);
    }
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

template <int CIN>
__global__ void __launch_bounds__(256) conv3x3_f32_kernel(const float* __restrict__ in,   // [n][64][CIN]
                                                          const float* __restrict__ w,    // [3][3][CIN][192]
                                                          const float* __restrict__ bias, // [192]
                                                          float* __restrict__ out) {      // [n][64][192]
  extern __shared__ float a_s[];   // [10][10][CIN]
  const long long board = blockIdx.x;
  for (int i = threadIdx.x; i < 100 * CIN; i += 256) a_s[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * CIN; i += 256) {
    const int px = i / CIN, c = i - px * CIN;
    a_s[((px / 8 + 1) * 10 + (px % 8) + 1) * CIN + c] = in[board * 64 * CIN + i];
  }
  __syncthreads();
  const int y = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[8][6];
#pragma unroll
  for (int x = 0; x < 8; ++x)
#pragma unroll
    for (int j = 0; j < 6; ++j) acc[x][j] = 0.f;
  for (int dy = 0; dy < 3; ++dy) {
    for (int ci = 0; ci < CIN; ++ci) {
      float a[10];
#pragma unroll
      for (int x = 0; x < 10; ++x) a[x] = a_s[((y + dy) * 10 + x) * CIN + ci];
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float* wr = w + (static_cast<size_t>(dy * 3 + dx) * CIN + ci) * 192 + lane;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const float wv = __ldg(wr + 32 * j);
#pragma unroll
          for (int x = 0; x < 8; ++x) acc[x][j] = fmaf(a[x + dx], wv, acc[x][j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const float bv = bias[lane + 32 * j];
#pragma unroll
    for (int x = 0; x < 8; ++x)
      out[(board * 64 + y * 8 + x) * 192 + lane + 32 * j] = fmaxf(acc[x][j] + bv, 0.f);
  }
}

// 192 -> 1 channel conv (k = 3 or 1): one warp per output pixel, lanes split the channels.
__global__ void __launch_bounds__(256) conv_small_f32_kernel(const float* __restrict__ in,  // [n][64][192]
                                                             const float* __restrict__ w,   // [k][k][192][1]
                                                             const float* __restrict__ bias, int k,
                                                             float* __restrict__ h5) {      // [n][64]
  const long long board = blockIdx.x;
  const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = (k - 1) / 2;
  for (int px = wp; px < 64; px += 8) {
    const int y = px >> 3, x = px & 7;
    float acc = 0.f;
    for (int i = 0; i < k; ++i) {
      const int yy = y + i - p;
      if (yy < 0 || yy > 7) continue;
      for (int j = 0; j < k; ++j) {
        const int xx = x + j - p;
        if (xx < 0 || xx > 7) continue;
        const float* a = in + (board * 64 + yy * 8 + xx) * 192;
        const float* wr = w + (i * k + j) * 192;
        for (int c = lane; c < 192; c += 32) acc = fmaf(a[c], wr[c], acc);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) h5[board * 64 + px] = fmaxf(acc + bias[0], 0.f);
  }
}

}  // namespace wann
