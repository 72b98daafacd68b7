// conv_stack_split: fp32-CLASS accuracy on the tcgen05 tensor cores (WANN_MODE_FP32_TC).
//
// Same implicit-GEMM conv stack, same shared-memory activation layout and weight ring as
// conv_stack_tc.cuh, but every operand is carried as TWO fp16 numbers
//        a = a_hi + a_lo / 4096,   a_hi = fp16(a),   a_lo = fp16((a - a_hi) * 4096)
// (22 mantissa bits; fp16 x fp16 products are exact in the fp32 accumulator), and
//        a . w  ~=  a_hi.w_hi  +  (a_lo.w_hi + a_hi.w_lo) / 4096          (lo.lo term dropped: 2^-24)
// is formed by THREE MMAs per K-step into TWO TMEM accumulators:
//        acc0 += A_hi * W_hi          acc1 += A_lo * W_hi + A_hi * W_lo
// and combined in the epilogue (acc0 + acc1 * 2^-12 + bias, ReLU, split again for the next layer).
// The two row tiles of a CTA hold the hi and the lo parts of the SAME two boards, so a CTA pair
// evaluates 4 boards per supertile; each K-block streams two weight stages (hi, lo).
// Cost: 3 MMAs per K-step and half the boards per tile -> about a third of the bf16/fp16 modes'
// throughput, ~20x the CUDA-core fp32 mode; logits agree with the fp32 oracle to ~1e-6.
//
// Structure is deliberately the simple one: one MMA issuer warp, epilogue serial with the MMAs.
#pragma once
#include "conv_stack_tc.cuh"

namespace wann {

constexpr int kSplitBoardsPerPair = 4;
constexpr float kSplitScale = 4096.f;
constexpr int kSplitStagesPerSupertile = 2 * kKbPerSupertile;
constexpr int kSplitImgBytesPerRank = 2 * kImgBytesPerRank;

// x >= 0 -> (hi, lo) fp16 pair with x ~= hi + lo / 4096
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(fminf(x, 65504.f));
  lo = __float2half_rn((x - __half2float(hi)) * kSplitScale);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_stack_split_kernel(const TcParams p) {
  constexpr uint32_t kIdescN192 = ptx::make_idesc_16bit(256, 192, true);
  constexpr uint32_t kIdescN32 = ptx::make_idesc_16bit(256, 32, true);
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
  const uint32_t bar_acc_full = bar_base + 8u * (3 * kStages);
  const uint32_t bar_a_ready = bar_base + 8u * (3 * kStages + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  // refinement pass: the number of boards is only known on the device (written by the tail kernel
  // launched before us on the same stream, never under programmatic dependent launch)
  const long long n_eff = effective_n(p);
  const long long total_st = (n_eff + kSplitBoardsPerPair - 1) / kSplitBoardsPerPair;
  if (pair >= total_st) return;   // both CTAs of the pair leave together, before any cluster traffic

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(bar_full(s), 1);
      ptx::mbar_init(bar_peer_full(s), 1);
      ptx::mbar_init(bar_empty(s), 1);
    }
    ptx::mbar_init(bar_acc_full, 1);
    ptx::mbar_init(bar_a_ready, 2);
    ptx::fence_mbar_init();
  }
  for (int i = tid; i < kABytes / 16; i += kThreads)
    ptx::st_shared_v4(a_base + 16u * i, make_uint4(0u, 0u, 0u, 0u));
  for (int i = tid; i < kBiasFloats; i += kThreads)
    reinterpret_cast<float*>(smem_gen + kOffBias)[i] = (i < 4 * kC + 1) ? p.bias[i] : 0.f;
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

  if (warp == 0) {
    // ===== weight producer: per K-block the hi stage, then the lo stage ======================
    if (lane == 0) {
      const uint8_t* img = p.wimg + static_cast<size_t>(rank) * kSplitImgBytesPerRank;
      uint32_t s = 0, ph = 0;
      bool ok = true;
      for (long long st = pair; st < total_st && ok; st += npairs) {
        size_t off = 0;
        for (int q = 0; q < kSplitStagesPerSupertile; ++q) {
          const uint32_t bytes = (q >= kSplitStagesPerSupertile - 2 * kKbL5) ? kStageBytesL5 : kStageBytes;
          if (!wait_bar(bar_empty(s), ph ^ 1u, false, p.err, 100 + s)) { ok = false; break; }
          ptx::mbar_arrive_expect_tx(bar_full(s), bytes);
          ptx::bulk_g2s(b_base + s * kStageBytes, img + off, bytes, bar_full(s));
          off += bytes;
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1 && rank != 0) {
    if (lane == 0) {   // relay: my half of stage s has landed
      uint32_t s = 0, ph = 0;
      bool ok = true;
      for (long long st = pair; st < total_st && ok; st += npairs) {
        for (int q = 0; q < kSplitStagesPerSupertile; ++q) {
          if (!wait_bar(bar_full(s), ph, false, p.err, 500 + s)) { ok = false; break; }
          ptx::mbar_arrive_cluster(ptx::mapa(bar_peer_full(s), 0));
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA) ============================================================
    const uint32_t issuer = ptx::elect_one();
    const uint32_t a_hi_desc = (kRowStride >> 4) | (1u << 14);
    const uint32_t b_hi_desc = (kSboB >> 4) | (1u << 14);
    const uint32_t a_lo_hi = (a_base >> 4) | ((kChunkStride >> 4) << 16);                  // tile 0: hi parts
    const uint32_t a_lo_lo = ((a_base + kATileStride) >> 4) | ((kChunkStride >> 4) << 16);  // tile 1: lo parts
    const uint32_t b_lo_mid = (b_base >> 4) | ((kLboB >> 4) << 16);
    const uint32_t b_lo_l5 = (b_base >> 4) | ((kLboB5 >> 4) << 16);
    const uint32_t acc0 = tmem_base, acc1 = tmem_base + kC;
    constexpr uint32_t kChunk16 = kChunkStride >> 4;
    constexpr uint32_t kStage16 = kStageBytes >> 4;
    uint32_t s = 0, ph = 0, lstep = 0;
    bool ok = true;
    // one K-block = stage s (W_hi) and stage s+1 (W_lo): up to 4 K=16 steps x 3 MMAs
#define WANN_SPLIT_KBLOCK(NKS, A_OFF16_OF_KS, B_LO, B_KS16, IDESC, FIRST, WITH_A_LO)               \
    {                                                                                              \
      const uint32_t s1 = (s + 1 == kStages) ? 0u : s + 1;                                         \
      const uint32_t ph1 = (s + 1 == kStages) ? ph ^ 1u : ph;                                      \
      if (!wait_bar(bar_full(s), ph, false, p.err, 300 + s) ||                                     \
          !wait_bar(bar_peer_full(s), ph, true, p.err, 400 + s) ||                                 \
          !wait_bar(bar_full(s1), ph1, false, p.err, 300 + s1) ||                                  \
          !wait_bar(bar_peer_full(s1), ph1, true, p.err, 400 + s1)) { ok = false; break; }         \
      ptx::tc_fence_after();                                                                       \
      const uint32_t bh = (B_LO) + s * kStage16, bl = (B_LO) + s1 * kStage16;                      \
      _Pragma("unroll") for (int ks = 0; ks < (NKS); ++ks) {                                       \
        const uint32_t ao = (A_OFF16_OF_KS);                                                       \
        const uint32_t is_first = ((FIRST) && ks == 0) ? 1u : 0u;                                  \
        if (issuer) {                                                                              \
          ptx::umma_bf16_2cta(acc0, a_lo_hi + ao, a_hi_desc, bh + ks * (B_KS16), b_hi_desc, (IDESC), is_first ^ 1u); \
          if (WITH_A_LO) {                                                                         \
            ptx::umma_bf16_2cta(acc1, a_lo_lo + ao, a_hi_desc, bh + ks * (B_KS16), b_hi_desc, (IDESC), is_first ^ 1u); \
            ptx::umma_bf16_2cta(acc1, a_lo_hi + ao, a_hi_desc, bl + ks * (B_KS16), b_hi_desc, (IDESC), 1u); \
          } else {                                                                                 \
            ptx::umma_bf16_2cta(acc1, a_lo_hi + ao, a_hi_desc, bl + ks * (B_KS16), b_hi_desc, (IDESC), is_first ^ 1u); \
          }                                                                                        \
        }                                                                                          \
      }                                                                                            \
      if (issuer) { ptx::umma_commit<2>(bar_empty(s), 3); ptx::umma_commit<2>(bar_empty(s1), 3); } \
      s = (s1 + 1 == kStages) ? 0u : s1 + 1;                                                       \
      ph = (s1 + 1 == kStages) ? ph1 ^ 1u : ph1;                                                   \
    }
    for (long long st = pair; st < total_st && ok; st += npairs) {
      do {   // conv1: the 0/1 planes are exact in fp16 (A_lo = 0): acc1 = A_hi * W_lo only
        if (!wait_bar(bar_a_ready, lstep & 1u, true, p.err, 200)) { ok = false; break; }
        ptx::tc_fence_after();
        constexpr uint32_t kTap16[12] = {0, 1, 2, 18, 19, 20, 36, 37, 38, 0, 0, 0};
        WANN_SPLIT_KBLOCK(4, kTap16[ks], b_lo_mid, (2 * kLboB) >> 4, kIdescN192, true, false)
        WANN_SPLIT_KBLOCK(4, kTap16[4 + ks], b_lo_mid, (2 * kLboB) >> 4, kIdescN192, false, false)
        WANN_SPLIT_KBLOCK(1, kTap16[8 + ks], b_lo_mid, (2 * kLboB) >> 4, kIdescN192, false, false)
        if (issuer) ptx::umma_commit<2>(bar_acc_full, 3);
        ++lstep;
      } while (false);
      for (int layer = 1; layer <= 3 && ok; ++layer) {
        if (!wait_bar(bar_a_ready, lstep & 1u, true, p.err, 200 + layer)) { ok = false; break; }
        ptx::tc_fence_after();
        uint32_t first = 1;
        for (uint32_t ty = 0; ty < 3 && ok; ++ty)
          for (uint32_t tx = 0; tx < 3 && ok; ++tx) {
            const uint32_t tap16 = ty * (kYStride >> 4) + tx;
            for (uint32_t j = 0; j < 3; ++j) {
              const uint32_t ao0 = j * 8 * kChunk16 + tap16;
              WANN_SPLIT_KBLOCK(4, ao0 + ks * 2 * kChunk16, b_lo_mid, (2 * kLboB) >> 4, kIdescN192, first, true)
              first = 0;
            }
          }
        if (!ok) break;
        if (issuer) ptx::umma_commit<2>(bar_acc_full, 3);
        ++lstep;
      }
      if (!ok) break;
      do {   // conv5 as a 1x1 GEMM with N = 32 (columns 0..8 = the 9 taps; 16..31 unused here)
        if (!wait_bar(bar_a_ready, lstep & 1u, true, p.err, 204)) { ok = false; break; }
        ptx::tc_fence_after();
        constexpr uint32_t kCentre16 = (kYStride >> 4) + 1;
        for (uint32_t j = 0; j < 3; ++j) {
          const uint32_t ao0 = j * 8 * kChunk16 + kCentre16;
          WANN_SPLIT_KBLOCK(4, ao0 + ks * 2 * kChunk16, b_lo_l5, (2 * kLboB5) >> 4, kIdescN32, j == 0, true)
        }
        if (!ok) break;
        if (issuer) ptx::umma_commit<2>(bar_acc_full, 3);
        ++lstep;
      } while (false);
    }
#undef WANN_SPLIT_KBLOCK
  } else if (warp >= kFirstEpiWarp && warp < kFirstTailWarp) {
    // ===== epilogue warps =======================================================================
    const int ew = warp - kFirstEpiWarp;
    const int quad = warp & 3;
    const int chalf = ew >> 2;
    const int row = quad * 32 + lane;
    const int y = row >> 4, b = (row >> 3) & 1, x = row & 7;
    const uint32_t cell_hi = a_base + ((y + 1) * 2 + b) * kRowStride + (x + 1) * kCell;   // tile 0
    const uint32_t cell_lo = cell_hi + kATileStride;                                       // tile 1
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const bool signaller = (tid == kFirstEpiWarp * 32);
    const uint32_t a_ready_leader = ptx::mapa(bar_a_ready, 0);

    auto signal_a_ready = [&]() {
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(kNumEpiWarps * 32) : "memory");
      if (signaller) ptx::mbar_arrive_cluster(a_ready_leader);
    };
    auto board_of = [&](long long st) {
      return st * kSplitBoardsPerPair + static_cast<int>(rank) * 2 + b;
    };
    auto load_in = [&](long long board) {
      if (board >= n_eff) return make_uint2(0u, 0u);
      return load_input_cell<true>(p, p.n, p.gather != nullptr ? static_cast<long long>(p.gather[board]) : board, y, x);
    };

    ptx::pdl_wait();
    long long st = pair;
    if (st < total_st) {
      if (chalf == 0) store_input_cell<false>(cell_hi, load_in(board_of(st)));
      signal_a_ready();
    }
    uint32_t lstep = 0;
    bool ok = true;
    for (; st < total_st && ok; st += npairs) {
      const long long board = board_of(st);
      for (int layer = 0; layer < 5; ++layer, ++lstep) {
        if (!wait_bar(bar_acc_full, lstep & 1u, true, p.err, 600 + layer)) { ok = false; break; }
        ptx::tc_fence_after();
        if (layer < 4) {
          const float* bs = bias_s + layer * kC + chalf * 96;
          float* dbg = (p.dbg != nullptr && p.dbg_layer == layer + 1 && board < n_eff)
                           ? p.dbg + (board * 64 + y * 8 + x) * kC + chalf * 96 : nullptr;
#pragma unroll 1
          for (int ch = 0; ch < 3; ++ch) {
            uint32_t v0[32], v1[32];
            ptx::tmem_ld_x32(taddr + chalf * 96 + ch * 32, v0);
            ptx::tmem_ld_x32(taddr + kC + chalf * 96 + ch * 32, v1);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int c = kc * 8 + 2 * i;
                const float x0 = fmaxf(__uint_as_float(v0[c]) + __uint_as_float(v1[c]) * (1.f / kSplitScale) +
                                       bs[ch * 32 + c], 0.f);
                const float x1 = fmaxf(__uint_as_float(v0[c + 1]) + __uint_as_float(v1[c + 1]) * (1.f / kSplitScale) +
                                       bs[ch * 32 + c + 1], 0.f);
                __half h0, l0, h1, l1;
                split_f16(x0, h0, l0);
                split_f16(x1, h1, l1);
                __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
                hi[i] = *reinterpret_cast<uint32_t*>(&hh);
                lo[i] = *reinterpret_cast<uint32_t*>(&ll);
                if (dbg != nullptr) {
                  dbg[ch * 32 + c] = __half2float(h0) + __half2float(l0) * (1.f / kSplitScale);
                  dbg[ch * 32 + c + 1] = __half2float(h1) + __half2float(l1) * (1.f / kSplitScale);
                }
              }
              const uint32_t coff = (chalf * 12 + ch * 4 + kc) * kChunkStride;
              ptx::st_shared_v4(cell_hi + coff, make_uint4(hi[0], hi[1], hi[2], hi[3]));
              ptx::st_shared_v4(cell_lo + coff, make_uint4(lo[0], lo[1], lo[2], lo[3]));
            }
          }
        } else if (chalf == 0) {
          uint32_t v0[16], v1[16];
          ptx::tmem_ld_x16(taddr, v0);
          ptx::tmem_ld_x16(taddr + kC, v1);
          ptx::tmem_ld_wait();
          float pt[9];
#pragma unroll
          for (int t = 0; t < 9; ++t)
            pt[t] = __uint_as_float(v0[t]) + __uint_as_float(v1[t]) * (1.f / kSplitScale);
          ptx::st_shared_v4(cell_hi + 2 * kChunkStride, make_uint4(__float_as_uint(pt[0]), __float_as_uint(pt[1]),
                                                                   __float_as_uint(pt[2]), __float_as_uint(pt[3])));
          ptx::st_shared_v4(cell_hi + 3 * kChunkStride, make_uint4(__float_as_uint(pt[4]), __float_as_uint(pt[5]),
                                                                   __float_as_uint(pt[6]), __float_as_uint(pt[7])));
          ptx::st_shared_v4(cell_hi + 4 * kChunkStride, make_uint4(__float_as_uint(pt[8]), 0u, 0u, 0u));
          asm volatile("bar.sync 2, 128;" ::: "memory");
          float acc = 0.f;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            acc += ptx::ld_shared_f32(cell_hi + (2 + tap / 4) * kChunkStride + dy * kYStride + dx * kCell +
                                      (tap % 4) * 4);
          }
          const float hval = fmaxf(acc + bias_s[4 * kC], 0.f);
          if (board < n_eff) p.h5[board * 64 + y * 8 + x] = hval;
          if (st + npairs < total_st)
            store_input_cell<false>(cell_hi, load_in(board_of(st + npairs)));
        }
        signal_a_ready();
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 0) ptx::tmem_dealloc<2>(tmem_base, 512);
}

}  // namespace wann
