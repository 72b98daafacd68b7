// array_tree.hpp -- the reference's search tree over flat arrays, many independent trees per forest.
//
// What ONE thread of the reference's search does per tree, restated over flat node records instead of
// Python dicts.  The same code runs on the HOST (wann_forest_next / _apply: OpenMP over
// trees, buffers grow on demand) and on the DEVICE (forest kernels in wann_api.cu: one warp per tree, lane
// 0 walks the tree, fixed-capacity arenas in HBM), hence no std:: containers.  It is RESUMABLE at the one point
// where the reference blocks -- policy_net.evaluate (tree_builder.pyx:189) -- so that a forest of
// trees advances in lock step: every tree hands out the position it needs evaluated, ONE batched
// wann_eval_expand_seeded call serves them all, every tree applies its row and runs on to its next
// evaluation.  BASELINE config 5 with the reference's own search semantics per tree.
//
//   run_simulation           expansion_MCTS_functions.pyx:300-350
//   choose / find_best_UCT   tree_search_utils.pyx:251-403, get_UCT :405-467
//   expand_descendants       tree_builder.pyx:122-215, do_updates_in_the_same_process :294-340,
//                            update_parent :344-355, get_child :530-536, assign_children :538-620
//   statistics / solver      update_tree_wins/_losses :43-66, update_win_statuses & co :70-132,
//                            backpropagate_num_checked_children :134-158, thread counters :161-242
// The child seeds (update_child, check_for_winning_move, the "maybe gameover next move" branch,
// evaluation_function) arrive from seed_children_kernel as flags + statistics per ranked child.
// Semantics are pinned by tests/golden/ref_tree_golden.npz (whole trees grown by the compiled
// reference) through oracle/tree_oracle.py and directly.
#pragma once
#include <stdint.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#ifdef __CUDACC__
#define WANN_HD __host__ __device__
#else
#define WANN_HD
#endif
// Read-modify-write of tree memory on the device: ONE lane writes, the warp synchronises before anyone
// reads the result (plain stores of a computed value are made by all lanes: same address, same value).
#ifdef __CUDA_ARCH__
#define WANN_LEADER ((threadIdx.x & 31u) == 0u)
#define WANN_WARP_SYNC() __syncwarp()
#else
#define WANN_LEADER true
#define WANN_WARP_SYNC() ((void)0)
#endif

namespace wann_tree {

constexpr int kMaxLegal = 48;
enum : uint8_t { kSeedWin = 1, kSeedCapture = 2, kSeedCloseToHome = 4, kSeedGameSaving = 8, kSeedDoomed = 16, kSeedEvalOne = 32 };

struct Tree {
  // ---- nodes (TreeNode.pyx:7-50): one 72-byte record per node (children of a node are contiguous,
  // so selection walks one cache-friendly run instead of seventeen arrays).  Boards are kept only for
  // nodes that get evaluated (1 in ~23): a child's board is its parent's board plus its move. ----------
  struct Node {
    int64_t visits, wins, gv, gw;
    double uct;                                 // UCT_multiplier (a float32-rounded number when uct_f32)
    int32_t parent, first, count, height, threads;
    int32_t board;                              // slot in `boards` (x 64 bytes), -1 until the node is evaluated
    int8_t ws;                                  // -1 None / 0 False / 1 True
    uint8_t index, black, gameover, checked, being_checked, uct_f32;
  };
  // Storage: flat arenas.  Host trees own malloc'ed buffers that grow (owns = 1); device trees are given
  // fixed slices of HBM (and a second pair, alt_*, that advance_root compacts into) and set `overflow`
  // instead of growing: the tree then stops (phase = kFinished) with its statistics intact.
  Node* nodes = nullptr;
  int8_t* boards = nullptr;                     // slots of 64 bytes
  Node* alt_nodes = nullptr;
  int8_t* alt_boards = nullptr;
  int32_t n_nodes = 0, cap_nodes = 0, n_boards = 0, cap_boards = 0;
  uint8_t owns = 0, overflow = 0;
  int32_t root = 0;
  // ---- resumable state ---------------------------------------------------------------------------
  enum Phase : uint8_t { kIdle, kWaitEval, kFinished };
  Phase phase = kIdle;
  int32_t ex_leaf = -1, ex_depth = 0, ex_limit = 0, pending = -1;
  bool ex_final = false;                        // the expansion run_simulation makes for a solved, childless root
  bool ex_no_root = false;                      // init_new_root: sim_info['root'] is None while it expands
  int pending_move = -1;                        // ... and the move to advance by once that expansion is done
  int32_t frontier[4] = {0, 0, 0, 0};           // (the reference's frontier list never holds more than two nodes)
  int32_t n_frontier = 0;
  int64_t simulations = 0, evaluations = 0;

  // DEVICE execution model: all 32 lanes of a warp run this code on the SAME tree with identical state
  // and uniform control flow (loads are broadcasts, the redundant stores write identical values); the
  // loops over the children of a node -- where the time goes -- are spread over the lanes and combined
  // with warp collectives, which also re-converge the warp.  On the host the loops are plain loops.
#ifdef __CUDA_ARCH__
  static __device__ __forceinline__ int lane_id() { return static_cast<int>(threadIdx.x & 31u); }
  static __device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  }
#endif
  WANN_HD void frontier_push(int32_t v) { if (n_frontier < 4) frontier[n_frontier++] = v; }
  // room for `extra` more nodes / boards?  Host: grow.  Device: no.
  WANN_HD bool reserve_nodes(int32_t extra) {
    if (n_nodes + extra <= cap_nodes) return true;
#ifndef __CUDA_ARCH__
    if (owns) {
      int32_t cap = cap_nodes > 0 ? cap_nodes : 16384;
      while (cap < n_nodes + extra) cap *= 2;
      Node* nn = static_cast<Node*>(malloc(sizeof(Node) * static_cast<size_t>(cap)));
      if (nn != nullptr) {
        if (n_nodes > 0) memcpy(nn, nodes, sizeof(Node) * static_cast<size_t>(n_nodes));
        free(nodes);
        nodes = nn;
        cap_nodes = cap;
        return true;
      }
    }
#endif
    overflow = 1;
    return false;
  }
  WANN_HD bool reserve_boards(int32_t extra) {
    if (n_boards + extra <= cap_boards) return true;
#ifndef __CUDA_ARCH__
    if (owns) {
      int32_t cap = cap_boards > 0 ? cap_boards : 64;
      while (cap < n_boards + extra) cap *= 2;
      int8_t* nb = static_cast<int8_t*>(malloc(static_cast<size_t>(cap) * 64));
      if (nb != nullptr) {
        if (n_boards > 0) memcpy(nb, boards, static_cast<size_t>(n_boards) * 64);
        free(boards);
        boards = nb;
        cap_boards = cap;
        return true;
      }
    }
#endif
    overflow = 1;
    return false;
  }
  void release() {                              // host trees only
    if (owns) { free(nodes); free(boards); }
    *this = Tree();
  }

  WANN_HD int32_t new_node(int32_t par, int idx, bool blk, int32_t h, double u, bool u_f32) {   // caller reserved the room
    Node nd;
    nd.visits = nd.wins = nd.gv = nd.gw = 0;
    nd.uct = u;
    nd.parent = par; nd.first = -1; nd.count = -1; nd.height = h; nd.threads = 0;
    nd.ws = -1;
    nd.index = static_cast<uint8_t>(idx); nd.black = blk; nd.gameover = 0; nd.checked = 0; nd.being_checked = 0;
    nd.uct_f32 = u_f32;
    nd.board = -1;
    nodes[n_nodes] = nd;
    return n_nodes++;
  }
  // host tree: its own buffers (16,384 nodes of address space up front -- pages are touched as nodes
  // arrive --: the first ~700 evaluations of a search never re-allocate / copy)
  void init(const int8_t* squares, bool blk, int32_t h) {
    owns = 1;
    init_in_place(squares, blk, h);
  }
  // (re)start on the buffers the tree already has (device trees; host trees after reserve)
  WANN_HD bool init_in_place(const int8_t* squares, bool blk, int32_t h) {
    n_nodes = n_boards = 0;
    root = 0;
    phase = kIdle;
    ex_leaf = -1; ex_depth = ex_limit = 0; pending = -1;
    ex_final = ex_no_root = false;
    pending_move = -1;
    n_frontier = 0;
    simulations = evaluations = 0;
    overflow = 0;
    if (!reserve_nodes(1) || !reserve_boards(1)) { phase = kFinished; return false; }
    root = new_node(-1, 0, blk, h, 1.0, false);
    nodes[root].board = 0;
    memcpy(boards, squares, 64);
    n_boards = 1;
    return true;
  }
  // the position of node n (materialised from its parent's on first use; parents are always expanded,
  // hence evaluated, hence have theirs)
  // (nullptr: no room for another board on a fixed-capacity tree -- `overflow` is set)
  WANN_HD const int8_t* board_of(int32_t n) {
    if (nodes[n].board < 0) {
      const int32_t par = nodes[n].parent;
      if (!reserve_boards(1)) return nullptr;
      const size_t slot = static_cast<size_t>(n_boards++);
      apply_move(boards + static_cast<size_t>(nodes[par].board) * 64, nodes[par].black, nodes[n].index, boards + slot * 64);
      nodes[n].board = static_cast<int32_t>(slot);
    }
    return boards + static_cast<size_t>(nodes[n].board) * 64;
  }

  // ---- statistics (tree_search_utils.pyx:43-66) -------------------------------------------------
  WANN_HD void update_tree_wins(int32_t n, int64_t amount, bool go) {
    bool win = true;                            // node win => parent loss => grandparent win ...
    while (n >= 0) {
      if (WANN_LEADER) {
        if (win) nodes[n].wins += amount;
        nodes[n].visits += amount;
        if (go) { if (win) nodes[n].gw += amount; nodes[n].gv += amount; }
      }
      n = nodes[n].parent;
      win = !win;
    }
    WANN_WARP_SYNC();
  }
  // Several updates that start at the same node, in ONE walk to the root (they are all additions, and
  // nothing reads the statistics in between): update_tree_wins(n, w), update_tree_losses(n, l),
  // update_tree_wins(n, wg, gameover), update_tree_losses(n, lg, gameover).  A walk up a deep tree is a
  // chain of cache misses; an expansion needs this one instead of 2 + (winning + doomed children).
  WANN_HD void propagate(int32_t n, int64_t w, int64_t l, int64_t wg, int64_t lg) {
    while (n >= 0 && (w | l | wg | lg) != 0) {
      Node& nd = nodes[n];
      if (WANN_LEADER) {
        nd.wins += w + wg;
        nd.visits += w + l + wg + lg;
        nd.gw += wg;
        nd.gv += wg + lg;
      }
      n = nd.parent;
      int64_t t = w; w = l; l = t;               // a win here is a loss for the parent
      t = wg; wg = lg; lg = t;
    }
    WANN_WARP_SYNC();
  }
  WANN_HD void update_tree_losses(int32_t n, int64_t amount, bool go) {
    if (WANN_LEADER) {
      nodes[n].visits += amount;
      if (go) nodes[n].gv += amount;
    }
    WANN_WARP_SYNC();
    if (nodes[n].parent >= 0) update_tree_wins(nodes[n].parent, amount, go);
  }

  // mixed float32 / double comparison of the reference (numpy scalar vs Python float: in float32)
  WANN_HD bool uct_gt(int32_t c, float thr) const { return uct_gt_node(nodes[c], thr); }
  WANN_HD static bool uct_gt_node(const Node& nd, float thr) {
    return nd.uct_f32 ? static_cast<float>(nd.uct) > thr : nd.uct > static_cast<double>(thr);
  }
  // the cascading child filters of find_best_UCT_child, stage by stage
  WANN_HD static bool admit_node(int stage, const Node& nd, float first_thr, float second_thr) {
    const bool ok_threads = nd.threads <= 0 || nd.count >= 0;
    const bool open_ = nd.ws == -1 && !nd.being_checked;
    switch (stage) {
      case 0: return open_ && ok_threads;                                                   // solved node
      case 1: return !(nd.checked || nd.being_checked) && ok_threads;
      case 2: return !nd.checked && ok_threads;
      case 10: return open_ && uct_gt_node(nd, first_thr) && ok_threads;                    // open node
      case 11: return open_ && uct_gt_node(nd, second_thr) && (nd.gw <= 1000 || nd.gw * 2 < nd.gv) && ok_threads;
      case 12: return open_ && uct_gt_node(nd, second_thr) && ok_threads;
      default: return open_ && ok_threads;                                                  // 13
    }
  }

  // ---- solver (tree_search_utils.pyx:70-132) ----------------------------------------------------
  // what update_win_status_from_children decides for node n: 1 / 0, or -2 = nothing to do
  WANN_HD int status_from_children(int32_t n) const {
    if (nodes[n].count < 0) return -2;
    bool any_f = false, any_t = false, any_n = false, c_any = false, c_t = false, c_n = false;
#ifdef __CUDA_ARCH__
    {
      unsigned bits = 0;
      const int32_t hi = nodes[n].first + nodes[n].count;
      for (int32_t c = nodes[n].first + lane_id(); c < hi; c += 32) {
        const int8_t s = nodes[c].ws;
        bits |= (s == 0 ? 1u : 0u) | (s == 1 ? 2u : 0u) | (s == -1 ? 4u : 0u);
        if (uct_gt(c, 1.001f)) bits |= 8u | (s == 1 ? 16u : 0u) | (s == -1 ? 32u : 0u);
      }
      bits = __reduce_or_sync(0xffffffffu, bits);
      any_f = bits & 1u; any_t = bits & 2u; any_n = bits & 4u; c_any = bits & 8u; c_t = bits & 16u; c_n = bits & 32u;
    }
#else
    for (int32_t c = nodes[n].first; c < nodes[n].first + nodes[n].count; ++c) {
      const int8_t s = nodes[c].ws;
      any_f |= s == 0; any_t |= s == 1; any_n |= s == -1;
      if (uct_gt(c, 1.001f)) { c_any = true; c_t |= s == 1; c_n |= s == -1; }
    }
#endif
    const bool use_c = !any_f && c_any;
    const bool has_t = use_c ? c_t : any_t, has_n = use_c ? c_n : any_n;
    if (any_f) return 1;
    if (has_t && !has_n) return 0;
    return -2;
  }
  // update_win_statuses <-> update_win_status_from_children call each other in tail position in the
  // reference; here as one loop up the tree (no recursion: the same code runs in a CUDA thread)
  WANN_HD void update_win_statuses(int32_t n, int8_t status) {
    while (true) {
      if (status == 0) update_tree_losses(n, 1, true);
      nodes[n].ws = status;
      n = nodes[n].parent;
      if (n < 0) return;
      const int st = status_from_children(n);
      if (st == -2) return;
      status = static_cast<int8_t>(st);
    }
  }
  WANN_HD void update_win_status_from_children(int32_t n) {
    const int st = status_from_children(n);
    if (st != -2) update_win_statuses(n, static_cast<int8_t>(st));
  }
  WANN_HD void backpropagate_num_checked_children(int32_t n) {      // :134-158
    while (n >= 0) {
      int k = 0;
#ifdef __CUDA_ARCH__
      for (int32_t c = nodes[n].first + lane_id(); c < nodes[n].first + nodes[n].count; c += 32) k += nodes[c].checked;
      k = __reduce_add_sync(0xffffffffu, k);
#else
      for (int32_t c = nodes[n].first; c < nodes[n].first + nodes[n].count; ++c) k += nodes[c].checked;
#endif
      if (k == nodes[n].count) { nodes[n].checked = 1; n = nodes[n].parent; }
      else { nodes[n].checked = 0; n = -1; }
    }
  }

  // ---- thread counters (:161-242) ----------------------------------------------------------------
  WANN_HD void crement_threads(int32_t n, int amount) {
    if (amount == 0) nodes[n].threads = 0;
    else {
      if (WANN_LEADER) nodes[n].threads += amount;
      WANN_WARP_SYNC();
    }
    if (nodes[n].parent >= 0) update_num_children_being_checked(nodes[n].parent);
  }
  WANN_HD void update_num_children_being_checked(int32_t n) {
    while (n >= 0 && nodes[n].count >= 0) {
      int k = 0;
#ifdef __CUDA_ARCH__
      for (int32_t c = nodes[n].first + lane_id(); c < nodes[n].first + nodes[n].count; c += 32)
        k += (nodes[c].threads > 0 && nodes[c].count < 0) || nodes[c].being_checked || nodes[c].ws != -1;
      k = __reduce_add_sync(0xffffffffu, k);
#else
      for (int32_t c = nodes[n].first; c < nodes[n].first + nodes[n].count; ++c)
        k += (nodes[c].threads > 0 && nodes[c].count < 0) || nodes[c].being_checked || nodes[c].ws != -1;
#endif
      const uint8_t previous = nodes[n].being_checked;
      nodes[n].being_checked = k == nodes[n].count;
      if (previous != nodes[n].being_checked && nodes[n].parent >= 0) n = nodes[n].parent; else break;
    }
  }

  // ---- selection (:251-467) -----------------------------------------------------------------------
  WANN_HD double get_uct(int32_t c, int64_t parent_visits) const { return get_uct_node(nodes[c], parent_visits, nodes[root].black); }
  WANN_HD static double get_uct_node(const Node& nd, int64_t parent_visits, bool root_black) {
    const int64_t nw = nd.wins, nv = nd.visits, tw = nd.gw, tg = nd.gv;
    // max(0.10, UCT_multiplier - 1): a float32 multiplier is subtracted and compared in float32
    double nn_prob;
    if (nd.uct_f32) {
      const float d = static_cast<float>(nd.uct) - 1.0f;
      nn_prob = d > 0.10f ? static_cast<double>(d) : 0.10;
    } else {
      const double d = nd.uct - 1.0;
      nn_prob = d > 0.10 ? d : 0.10;
    }
    double rate;
    if (tg > 10 && (nd.black != 0) != root_black) rate = static_cast<double>(tg - tw) / static_cast<double>(tg) + 1.0;
    else rate = static_cast<double>(nv - nw) / static_cast<double>(nv);
    const int64_t sib = parent_visits - nv > 1 ? parent_visits - nv : 1;
#ifdef __CUDA_ARCH__
    // every operation rounded on its own, like the host (nvcc would contract the product and the sum into an FMA)
    return __dadd_rn(rate, __dmul_rn(__dmul_rn(1.0, nn_prob),
                                     __ddiv_rn(sqrt(static_cast<double>(sib)), static_cast<double>(1 + nv))));
#else
    return rate + 1.0 * nn_prob * (std::sqrt(static_cast<double>(sib)) / static_cast<double>(1 + nv));
#endif
  }
  WANN_HD int32_t find_best_uct_child(int32_t n) const {
    const int32_t lo = nodes[n].first, hi = nodes[n].first + nodes[n].count;
    // the cascading filters, as predicates tried in order until one admits a child
    float first_thr = 0.f, second_thr = 0.f;
    const bool solved = nodes[n].ws != -1;
    if (!solved) {
      if (nodes[n].black == nodes[root].black) {
        int band;                                  // best_prob < .1 / < .2 / < .3 / else
        if (nodes[n].height >= 60) band = 0;             // best_prob = .01
        else if (nodes[lo].uct_f32) {                    // float32 multiplier: float32 difference, float32 comparisons
          const float bp = static_cast<float>(nodes[lo].uct) - 1.0f;
          band = bp < .1f ? 0 : bp < .2f ? 1 : bp < .3f ? 2 : 3;
        } else {
          const double bp = nodes[lo].uct - 1.0;
          band = bp < .1 ? 0 : bp < .2 ? 1 : bp < .3 ? 2 : 3;
        }
        if (band == 0) { first_thr = second_thr = 1.04f; }
        else if (band == 1) { first_thr = 1.10f; second_thr = 1.05f; }
        else if (band == 2) { first_thr = 1.15f; second_thr = 1.10f; }
        else { first_thr = 1.20f; second_thr = 1.15f; }
      } else { first_thr = 1.1f; second_thr = 1.01f; }
    }
    int stages[4];
    int ns = 0;
    if (solved) { stages[0] = 0; stages[1] = 1; stages[2] = 2; ns = 3; }
    else if (n == root) { stages[0] = 10; stages[1] = 11; stages[2] = 12; stages[3] = 13; ns = 4; }
    else { stages[0] = 10; stages[1] = 13; ns = 2; }
#ifdef __CUDA_ARCH__
    // lane-parallel: the node has <= 48 children = two rounds of 32 lanes.  The sequential scan's result is
    // reproduced exactly: the first admitted child if its value is NaN (nothing compares greater), else
    // the lowest-indexed child among those with the greatest non-NaN value.
    // Every lane loads its (at most two) children ONCE, decides for every stage whether the child is
    // admitted, and computes its value once.
    const int64_t pv = nodes[n].visits;
    const bool root_black = nodes[root].black != 0;
    unsigned adm_bits[2] = {0u, 0u};
    double v[2] = {0.0, 0.0};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int32_t c = lo + lane_id() + 32 * r;
      if (c < hi) {
        const Node nd = nodes[c];
        for (int s = 0; s < ns; ++s)
          if (admit_node(stages[s], nd, first_thr, second_thr)) adm_bits[r] |= 1u << s;
        if (adm_bits[r]) v[r] = get_uct_node(nd, pv, root_black);
      }
    }
    for (int s = 0; s < ns; ++s) {
      const bool adm[2] = {((adm_bits[0] >> s) & 1u) != 0, ((adm_bits[1] >> s) & 1u) != 0};
      const unsigned b0 = __ballot_sync(0xffffffffu, adm[0]), b1 = __ballot_sync(0xffffffffu, adm[1]);
      if ((b0 | b1) == 0) continue;
      const int first_r = b0 ? 0 : 1;
      const int first_lane = __ffs(b0 ? b0 : b1) - 1;
      const double v_first = __shfl_sync(0xffffffffu, first_r == 0 ? v[0] : v[1], first_lane);
      if (v_first != v_first) return lo + first_lane + 32 * first_r;
      double m = -INFINITY;
#pragma unroll
      for (int r = 0; r < 2; ++r) if (adm[r] && v[r] == v[r]) m = fmax(m, v[r]);
      m = warp_max(m);
      const unsigned e0 = __ballot_sync(0xffffffffu, adm[0] && v[0] == m), e1 = __ballot_sync(0xffffffffu, adm[1] && v[1] == m);
      return e0 ? lo + __ffs(e0) - 1 : lo + 32 + __ffs(e1) - 1;
    }
    return -1;
#else
    for (int s = 0; s < ns; ++s) {
      int32_t best = -1;
      double best_v = 0;
      const int64_t pv = nodes[n].visits;
      for (int32_t c = lo; c < hi; ++c) {
        if (!admit_node(stages[s], nodes[c], first_thr, second_thr)) continue;
        const double v = get_uct(c, pv);
        if (best < 0 || v > best_v) { best = c; best_v = v; }      // the FIRST maximum wins
      }
      if (best >= 0) return best;
    }
    return -1;
#endif
  }
  WANN_HD int32_t choose_uct_or_best_child(int32_t n) const { return nodes[n].count == 1 ? nodes[n].first : find_best_uct_child(n); }

  // ---- expansion ----------------------------------------------------------------------------------
  WANN_HD static void apply_move(const int8_t* board, bool blk, int idx, int8_t* out) {
#ifdef __CUDA_ARCH__
    reinterpret_cast<int16_t*>(out)[threadIdx.x & 31u] = reinterpret_cast<const int16_t*>(board)[threadIdx.x & 31u];   // 64-byte slots
    __syncwarp();
#else
    memcpy(out, board, 64);
#endif
    const int r = idx / 22, rem = idx - 22 * r, c = (rem + 1) / 3, d = rem - 3 * c;
    int frm = r * 8 + c, to = (r + 1) * 8 + c + d;
    if (blk) { frm = 63 - frm; to = 63 - to; }
    const int8_t piece = out[frm];
    out[to] = piece;
    out[frm] = 0;
    WANN_WARP_SYNC();
  }
  // update_parent -> get_child x k -> assign_children with the seeds of one evaluated position.
  // `next` receives what the reference's assign_children returns (next-frontier candidates).
  WANN_HD void expand_one(int32_t n, const uint8_t* idx, const float* prior, int k, const uint8_t* flags,
                          const int32_t* stats, bool threat, int32_t* next, int& n_next) {
    n_next = 0;
    ++evaluations;
    const bool blk = nodes[n].black;
    const int32_t base = n_nodes;
    int64_t up_w = 0, up_l = 0, up_wg = 0, up_lg = 0;   // what this expansion adds to n and its ancestors
#ifdef __CUDA_ARCH__
    {
      // lane j builds child j completely (both loops of the host version below: neither crement_threads nor
      // anything else in between looks at the new children) and stores its record once
      const bool attach = k > 0 && nodes[n].count < 0;
      const int32_t h = nodes[n].height + 1;
      int wg = 0, lg = 0, evl = 0, evw = 0;
      unsigned saving[2] = {0u, 0u};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int j = lane_id() + 32 * r;
        bool is_saving = false;
        if (j < k) {
          const uint8_t f = flags[j];
          const float p = prior[j];
          const bool f32 = f & kSeedCloseToHome;
          Node nd;
          nd.visits = nd.wins = nd.gv = nd.gw = 0;
          nd.uct = f32 ? static_cast<double>(1.0f + p) : 1.0 + static_cast<double>(p);
          nd.parent = n; nd.first = -1; nd.count = -1; nd.height = h; nd.threads = 0;
          nd.board = -1;
          nd.ws = -1;
          nd.index = idx[j]; nd.black = !blk; nd.gameover = 0; nd.checked = 0; nd.being_checked = 0;
          nd.uct_f32 = f32;
          if (f & kSeedWin) {
            nd.gameover = nd.checked = 1;
            nd.ws = 0;
            nd.gv = nd.visits = 65536 + 1;
            ++wg;
          }
          if (attach) {
            const int32_t* st = stats + 4 * j;
            if (threat) {
              if (f & kSeedGameSaving) { nd.visits = st[0]; nd.wins = st[1]; nd.gv = st[2]; nd.gw = st[3]; is_saving = true; }
              else if (f & kSeedDoomed) { nd.visits = nd.wins = nd.gw = nd.gv = 65536 + 1; nd.ws = 1; ++lg; }
            } else {
              if (!(f & kSeedWin)) { nd.visits = st[0]; nd.wins = st[1]; nd.gv = st[2]; nd.gw = st[3]; }
              if (f & kSeedEvalOne) ++evl; else ++evw;
            }
          }
          nodes[base + j] = nd;
        }
        saving[r] = __ballot_sync(0xffffffffu, is_saving);
      }
      n_nodes += k;
      up_wg = __reduce_add_sync(0xffffffffu, wg);
      up_lg = __reduce_add_sync(0xffffffffu, lg);
      const int ev_l = __reduce_add_sync(0xffffffffu, evl), ev_w = __reduce_add_sync(0xffffffffu, evw);
      __syncwarp();                                // the children are in memory for every lane
      crement_threads(n, -1);
      if (attach) {
        nodes[n].first = base;
        nodes[n].count = k;
        if (threat) {                              // `next`: the game-saving children, in order (its head is what is used)
          n_next = __popc(saving[0]) + __popc(saving[1]);
          if (n_next > 0) next[0] = saving[0] ? base + __ffs(saving[0]) - 1 : base + 32 + __ffs(saving[1]) - 1;
        } else {
          n_next = k;
          next[0] = base;
          up_l += ev_l;
          up_w += ev_w;
        }
        propagate(n, up_w, up_l, up_wg, up_lg);
        update_win_status_from_children(n);
        backpropagate_num_checked_children(n);
        update_num_children_being_checked(n);
      } else {
        propagate(n, up_w, up_l, up_wg, up_lg);
      }
    }
#else
    for (int j = 0; j < k; ++j) {                  // get_child: created (a winner's loss propagated) before the sort
      const uint8_t f = flags[j];
      const float p = prior[j];
      const bool f32 = f & kSeedCloseToHome;
      const double u = f32 ? static_cast<double>(1.0f + p) : 1.0 + static_cast<double>(p);
      const int32_t c = new_node(n, idx[j], !blk, nodes[n].height + 1, u, f32);
      if (f & kSeedWin) {                          // set_game_over_values + update_tree_losses(child)
        nodes[c].gameover = nodes[c].checked = 1;
        nodes[c].ws = 0;
        nodes[c].gv = nodes[c].visits = 65536 + 1;   // ... and its own share of update_tree_losses(child)
        ++up_wg;                                     // the rest of it: update_tree_wins(parent, 1, gameover)
      }
    }
    crement_threads(n, -1);                        // assign_children
    if (k > 0 && nodes[n].count < 0) {
      int64_t ev_l = 0, ev_w = 0;
      for (int j = 0; j < k; ++j) {
        const int32_t c = base + j;
        const uint8_t f = flags[j];
        const int32_t* st = stats + 4 * j;
        if (threat) {
          if (f & kSeedGameSaving) { nodes[c].visits = st[0]; nodes[c].wins = st[1]; nodes[c].gv = st[2]; nodes[c].gw = st[3]; next[n_next++] = c; }
          else if (f & kSeedDoomed) {              // set_game_over_values_1_step_lookahead + update_tree_wins(child)
            nodes[c].visits = nodes[c].wins = nodes[c].gw = nodes[c].gv = 65536 + 1;   // incl. update_tree_wins(child)
            nodes[c].ws = 1;
            ++up_lg;                                 // -> update_tree_losses(parent, 1, gameover)
          }
        } else {
          if (!(f & kSeedWin)) { nodes[c].visits = st[0]; nodes[c].wins = st[1]; nodes[c].gv = st[2]; nodes[c].gw = st[3]; }
          if (f & kSeedEvalOne) ++ev_l; else ++ev_w;
        }
      }
      nodes[n].first = base;
      nodes[n].count = k;
      if (!threat) {
        for (int j = 0; j < k; ++j) next[n_next++] = base + j;
        up_l += ev_l;                              // eval_children: update_tree_losses(n, L); update_tree_wins(n, W)
        up_w += ev_w;
      }
      propagate(n, up_w, up_l, up_wg, up_lg);
      update_win_status_from_children(n);
      backpropagate_num_checked_children(n);
      update_num_children_being_checked(n);
    } else {
      propagate(n, up_w, up_l, up_wg, up_lg);      // (children that were not attached still reported their results)
    }
#endif
  }

  // expand_descendants_to_depth_wrt_NN, resumable: runs until an evaluation is needed (returns the
  // node) or the expansion is over (returns -1).
  WANN_HD int32_t expand_run() {
    while (ex_depth < ex_limit) {
      if (!ex_no_root && nodes[root].ws == -1) { if (nodes[ex_leaf].count < 0 && nodes[ex_leaf].ws == -1) frontier_push(ex_leaf); }
      else if (nodes[ex_leaf].count < 0 && !nodes[ex_leaf].gameover) frontier_push(ex_leaf);
      if (n_frontier > 0) {                        // (the reference's 0 < len < 256; a frontier has one node)
        ex_limit = 800;
        const int32_t p = frontier[0];
        if (nodes[p].threads <= 1 && nodes[p].count < 0) {     // do_updates_in_the_same_process
          if (nodes[p].threads <= 0) crement_threads(p, 1);
          pending = p;
          return p;                                // -> policy_net.evaluate
        }
        crement_threads(p, -1);
        n_frontier = 0;
        ++ex_depth;
      } else {
        ex_depth = ex_limit;
        crement_threads(ex_leaf, 0);
      }
    }
    return -1;
  }
  WANN_HD void expand_begin(int32_t leaf, int depth, int limit) {
    ex_leaf = leaf; ex_depth = depth; ex_limit = limit;
    n_frontier = 0;
  }
  // continuation after the evaluation of `pending`
  WANN_HD void expand_resume(const uint8_t* idx, const float* prior, int k, const uint8_t* flags, const int32_t* stats,
                             bool threat) {
    if (k > kMaxLegal) k = kMaxLegal;
    if (!reserve_nodes(k)) {                       // a fixed-capacity (device) tree is full: it stops here
      phase = kFinished;
      pending = -1;
      return;
    }
    int32_t kids[kMaxLegal];
    int nk = 0;
    expand_one(pending, idx, prior, k, flags, stats, threat, kids, nk);
    n_frontier = 0;
    if (ex_depth + 1 < ex_limit && nk > 0) {
      const int32_t c0 = kids[0];
      if (nodes[c0].ws == -1 && nodes[c0].threads <= 0 && !nodes[c0].being_checked) frontier_push(c0);
    }
    ++ex_depth;
    if (ex_depth < ex_limit)
      for (int i = 0; i < n_frontier; ++i) crement_threads(frontier[i], 1);
    pending = -1;
  }

  // ---- the per-tree driver: advance until an evaluation is needed or `max_sims` are done ---------
  // Returns the node to evaluate, or -1 (finished / budget reached), or -2: `yield_after` simulations ended
  // inside this call without needing an evaluation (they stopped at a finished game or a solved node) and
  // the tree has more to do -- a lock-step driver calls again next step instead of letting one tree with a
  // run of such simulations hold up the step of all the others (device forests; hosts pass no limit).
  WANN_HD int32_t advance(int64_t max_sims, int yield_after = 0x7fffffff) {
    int ended_here = 0;
    while (true) {
      if (phase == kFinished) return -1;
      if (phase == kWaitEval) {                    // continue the expansion that was interrupted
        const int32_t need = expand_run();
        if (need >= 0) return need;
        if (pending_move >= 0) {                   // init_new_root's expansion is done: now reuse the child
          const int mv = pending_move;
          pending_move = -1;
          ex_no_root = ex_final = false;
          phase = kIdle;
          if (!advance_root(mv)) phase = kFinished;   // (a position without that move: nothing to search)
          continue;
        }
        if (!ex_final) {                           // tail of run_MCTS_with_expansions_simulation
          const bool done = nodes[root].checked || nodes[root].ws != -1;
          if (done && nodes[root].count < 0) {
            ex_final = true;
            expand_begin(root, 0, 1);
            continue;
          }
        }
        ex_final = false;
        phase = kIdle;
        ++simulations;
        if (++ended_here >= yield_after && simulations < max_sims && !nodes[root].checked) return -2;
        continue;
      }
      // kIdle: start a simulation
      if (simulations >= max_sims) return -1;
      if (nodes[root].checked) { phase = kFinished; return -1; }
      int32_t n = root;
      bool expanding = false;
      while (true) {                               // play_MCTS_game_with_expansions / select_UCT_child
        if (nodes[n].gameover) break;
        if (nodes[n].count < 0) {
          if (nodes[n].threads <= 0) {
            crement_threads(n, 1);
            expand_begin(n, 0, 5);
            expanding = true;
          }
          break;
        }
        while (n >= 0 && nodes[n].count >= 0) {
          nodes[n].threads = 0;
          n = choose_uct_or_best_child(n);
        }
        if (n < 0) break;
      }
      phase = kWaitEval;                           // (an empty expansion falls straight through above)
      if (!expanding) { ex_depth = ex_limit = 0; n_frontier = 0; }
    }
  }

  // ---- root reuse between moves: assign_root's "Reused old tree" branch (expansion_MCTS_functions.pyx:
  // 236-247,253-257) + reset_thread_flag (:197-207).  Returns false if the root has no such child.
  WANN_HD bool advance_root(int move_index) {
    if (phase == kWaitEval) return false;
    if (nodes[root].count < 0) {
      // init_new_root (tree_builder.pyx:809-897): the previous node was never expanded.  Expand it first
      // (threads_checking_node = 1; expand_descendants([parent], 0, 1) with sim_info['root'] still None);
      // the evaluations go through the usual next / apply cycle and the advance completes when they are
      // done.  (Its closing backpropagate_num_checked_children / update_win_status_from_children
      // (new_subtree=True) only touch the parent and above.)
      if (nodes[root].gameover) return false;
      nodes[root].threads = 1;
      expand_begin(root, 0, 1);
      ex_no_root = true;
      ex_final = true;
      pending_move = move_index;
      phase = kWaitEval;
      return true;
    }
    int32_t kid = -1;
    for (int32_t c = nodes[root].first; c < nodes[root].first + nodes[root].count; ++c)
      if (nodes[c].index == move_index) { kid = c; break; }
    if (kid < 0) return false;
    // Keep only the new root's subtree (the reference drops everything above the new root's
    // grandparent, :259-266; nothing above the root influences the search below it): breadth-first
    // copy into the second arena, so that the children of a node stay contiguous.  Thread bookkeeping
    // is cleared on the way.  A copied record keeps its OLD first/count until it is processed itself.
    if (board_of(kid) == nullptr) return false;  // its position, before its parent disappears
    Node* dst = alt_nodes;
    int8_t* dstb = alt_boards;
#ifndef __CUDA_ARCH__
    if (owns) {
      dst = static_cast<Node*>(malloc(sizeof(Node) * static_cast<size_t>(cap_nodes)));
      dstb = static_cast<int8_t*>(malloc(static_cast<size_t>(cap_boards) * 64));
      if (dst == nullptr || dstb == nullptr) { free(dst); free(dstb); overflow = 1; return false; }
    }
#endif
    if (dst == nullptr || dstb == nullptr) { overflow = 1; return false; }
    int32_t m = 0, mb = 0;
    dst[m++] = nodes[kid];
    for (int32_t i = 0; i < m; ++i) {
      Node nd = dst[i];
      nd.threads = 0;
      nd.being_checked = 0;
      if (nd.board >= 0) {
        memcpy(dstb + static_cast<size_t>(mb) * 64, boards + static_cast<size_t>(nd.board) * 64, 64);
        nd.board = mb++;
      }
      if (nd.count >= 0) {
        const int32_t old_first = nd.first;
        nd.first = m;
#ifdef __CUDA_ARCH__
        for (int32_t c = lane_id(); c < nd.count; c += 32) {
          Node ch = nodes[old_first + c];
          ch.parent = i;
          dst[m + c] = ch;
        }
        m += nd.count;
#else
        for (int32_t c = 0; c < nd.count; ++c) {
          dst[m] = nodes[old_first + c];
          dst[m].parent = i;
          ++m;
        }
#endif
      }
      dst[i] = nd;
      WANN_WARP_SYNC();                            // (a later iteration reads records another lane copied)
    }
    dst[0].parent = -1;
#ifndef __CUDA_ARCH__
    if (owns) {
      free(nodes);
      free(boards);
      nodes = dst;
      boards = dstb;
    } else
#endif
    {
      alt_nodes = nodes;
      alt_boards = boards;
      nodes = dst;
      boards = dstb;
    }
    n_nodes = m;
    n_boards = mb;
    root = 0;
    phase = kIdle;
    simulations = 0;                           // the budget of wann_forest_next counts per search
    return true;
  }

  // ---- the move the search plays: randomly_choose_a_winning_move (tree_search_utils.pyx:469-486;
  // deterministic despite its name), check_for_forced_win :559-582, choose_best_true_loser :493-557.
  // Num mirrors the reference's scalars: a Python float (weak) or a numpy float32 (the UCT_multiplier
  // of a capture close to home); numpy's promotion rules decide the precision of each operation.
  struct Num { double v; bool f32; };
  WANN_HD static Num num_op(Num a, Num b, char op) {
    if (a.f32 || b.f32) {
      const float x = static_cast<float>(a.v), y = static_cast<float>(b.v);
      const float r = op == '+' ? x + y : op == '-' ? x - y : x / y;
      return {static_cast<double>(r), true};
    }
    return {op == '+' ? a.v + b.v : op == '-' ? a.v - b.v : a.v / b.v, false};
  }
  WANN_HD static bool num_gt(Num a, Num b) {
    return (a.f32 || b.f32) ? static_cast<float>(a.v) > static_cast<float>(b.v) : a.v > b.v;
  }
  WANN_HD Num num_uct(int32_t c) const { return {nodes[c].uct, nodes[c].uct_f32 != 0}; }
  WANN_HD Num loss_rate_of(int32_t c) const {   // true_losses / max(1, gameover_visits) + (UCT_multiplier - 1) / 3
    const int64_t losses = nodes[c].gv - nodes[c].gw, total = nodes[c].gv > 1 ? nodes[c].gv : 1;
    const Num rate = {static_cast<double>(losses) / static_cast<double>(total), false};
    return num_op(rate, num_op(num_op(num_uct(c), Num{1.0, false}, '-'), Num{3.0, false}, '/'), '+');
  }
  WANN_HD int32_t best_move() const {
    const int32_t lo = nodes[root].first, k = nodes[root].count;
    if (k <= 0) return -1;
    for (int32_t c = lo; c < lo + k; ++c) if (nodes[c].gameover) return c;      // check_for_forced_win
    for (int32_t c = lo; c < lo + k; ++c) if (nodes[c].ws == 0) return c;
    // choose_best_true_loser
    const double action_value_threshold = static_cast<double>(nodes[lo].gw) / 2.0;
    const Num top = num_uct(lo);
    double first_thr, second_thr;
    if (num_gt(Num{1.1, false}, top)) first_thr = second_thr = 1.04;
    else if (num_gt(Num{1.2, false}, top)) { first_thr = 1.10; second_thr = 1.05; }
    else if (num_gt(Num{1.3, false}, top)) { first_thr = 1.15; second_thr = 1.10; }
    else { first_thr = 1.20; second_thr = 1.15; }
    auto losses = [&](int32_t c) { return nodes[c].gv - nodes[c].gw; };
    int32_t filtered[kMaxLegal];
    int nf = 0;
    for (int pass = 0; pass < 3 && nf == 0; ++pass)
      for (int32_t c = lo; c < lo + k; ++c) {
        if (nodes[c].ws == 1) continue;
        const bool keep = pass == 2 || num_gt(num_uct(c), Num{pass == 0 ? first_thr : second_thr, false}) ||
                          static_cast<double>(losses(c)) > action_value_threshold;
        if (keep && nf < kMaxLegal) filtered[nf++] = c;
      }
    if (nf == 0) return lo;
    int32_t best = filtered[0];
    for (int i = 0; i < nf; ++i) if (losses(filtered[i]) > losses(best)) best = filtered[i];   // max(): the first maximum
    const int64_t highest = losses(best);
    Num best_rate = loss_rate_of(best);
    for (int i = 0; i < nf; ++i) {
      const int32_t c = filtered[i];
      const Num rate = loss_rate_of(c);
      if (num_gt(rate, best_rate) && num_gt(num_op(rate, best_rate, '-'), Num{0.05, false}) && highest != 0 &&
          static_cast<double>(losses(c)) / static_cast<double>(highest) > .30 && nodes[c].ws != 1) {
        best = c;
        best_rate = rate;
      }
    }
    return best;
  }

  // Self-play step of one tree (wann_forest_play): play the move the search chose -- the game continues from
  // that child with its subtree reused -- or, when the move wins the game (or the root has no children: a
  // finished game, a tree that was never searched), start a new game from `restart_squares`.
  WANN_HD void play(const int8_t* restart_squares, bool restart_black, uint8_t& move_out, uint8_t& finished_out) {
    const int32_t c = best_move();
    uint8_t mv = 255, fin = 0;
    bool restart = false;
    if (c >= 0) {
      mv = nodes[c].index;
      restart = nodes[c].gameover || !advance_root(mv);
    } else {
      restart = true;
    }
    if (restart) {
      fin = 1;
      init_in_place(restart_squares, restart_black, 0);
    }
    move_out = mv;
    finished_out = fin;
  }

  // BFS dump in the layout of tests/golden/make_tree_golden.py; returns the number of rows (nodes)
  int64_t dump(int64_t* rows, int64_t capacity) const {
    std::vector<std::pair<int32_t, int64_t>> order;
    order.emplace_back(root, -1);
    for (size_t i = 0; i < order.size(); ++i) {
      const int32_t n = order[i].first;
      if (static_cast<int64_t>(i) < capacity) {
        int64_t* r = rows + i * 14;
        r[0] = order[i].second; r[1] = nodes[n].index; r[2] = nodes[n].black; r[3] = nodes[n].visits; r[4] = nodes[n].wins; r[5] = nodes[n].gv;
        r[6] = nodes[n].gw; r[7] = nodes[n].ws; r[8] = nodes[n].gameover; r[9] = nodes[n].checked; r[10] = nodes[n].being_checked;
        r[11] = nodes[n].threads; r[12] = nodes[n].count; r[13] = nodes[n].height;
      }
      if (nodes[n].count >= 0)
        for (int32_t c = nodes[n].first; c < nodes[n].first + nodes[n].count; ++c) order.emplace_back(c, static_cast<int64_t>(i));
    }
    return static_cast<int64_t>(order.size());
  }
};

struct Forest {
  std::vector<Tree> trees;
  ~Forest() { for (Tree& t : trees) t.release(); }
};

}  // namespace wann_tree
