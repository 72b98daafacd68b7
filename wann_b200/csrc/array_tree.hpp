// array_tree.hpp -- the reference's search tree over flat arrays, many independent trees per forest.
//
// Host-side C++ (no CUDA): what ONE thread of the reference's search does per tree, restated over
// struct-of-arrays nodes instead of Python dicts, and made RESUMABLE at the one point
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
#include <cstring>
#include <vector>

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
  std::vector<Node> nodes;
  std::vector<int8_t> boards;
  int32_t root = 0;
  // ---- resumable state ---------------------------------------------------------------------------
  enum Phase : uint8_t { kIdle, kWaitEval, kFinished };
  Phase phase = kIdle;
  int32_t ex_leaf = -1, ex_depth = 0, ex_limit = 0, pending = -1;
  bool ex_final = false;                        // the expansion run_simulation makes for a solved, childless root
  bool ex_no_root = false;                      // init_new_root: sim_info['root'] is None while it expands
  int pending_move = -1;                        // ... and the move to advance by once that expansion is done
  std::vector<int32_t> frontier;
  int64_t simulations = 0, evaluations = 0;

  int32_t new_node(int32_t par, int idx, bool blk, int32_t h, double u, bool u_f32) {
    Node nd;
    nd.visits = nd.wins = nd.gv = nd.gw = 0;
    nd.uct = u;
    nd.parent = par; nd.first = -1; nd.count = -1; nd.height = h; nd.threads = 0;
    nd.ws = -1;
    nd.index = static_cast<uint8_t>(idx); nd.black = blk; nd.gameover = 0; nd.checked = 0; nd.being_checked = 0;
    nd.uct_f32 = u_f32;
    nd.board = -1;
    nodes.push_back(nd);
    return static_cast<int32_t>(nodes.size()) - 1;
  }
  void init(const int8_t* squares, bool blk, int32_t h) {
    nodes.reserve(16384);                        // address space only (pages are touched as nodes arrive): the
                                                 // first ~700 evaluations of a search never re-allocate / copy
    root = new_node(-1, 0, blk, h, 1.0, false);
    nodes[root].board = 0;
    boards.assign(squares, squares + 64);
  }
  // the position of node n (materialised from its parent's on first use; parents are always expanded,
  // hence evaluated, hence have theirs)
  const int8_t* board_of(int32_t n) {
    if (nodes[n].board < 0) {
      const int32_t par = nodes[n].parent;
      const size_t slot = boards.size() / 64;
      boards.resize(boards.size() + 64);
      apply_move(&boards[static_cast<size_t>(nodes[par].board) * 64], nodes[par].black, nodes[n].index, &boards[slot * 64]);
      nodes[n].board = static_cast<int32_t>(slot);
    }
    return &boards[static_cast<size_t>(nodes[n].board) * 64];
  }

  // ---- statistics (tree_search_utils.pyx:43-66) -------------------------------------------------
  void update_tree_wins(int32_t n, int64_t amount, bool go) {
    bool win = true;                            // node win => parent loss => grandparent win ...
    while (n >= 0) {
      if (win) nodes[n].wins += amount;
      nodes[n].visits += amount;
      if (go) { if (win) nodes[n].gw += amount; nodes[n].gv += amount; }
      n = nodes[n].parent;
      win = !win;
    }
  }
  // Several updates that start at the same node, in ONE walk to the root (they are all additions, and
  // nothing reads the statistics in between): update_tree_wins(n, w), update_tree_losses(n, l),
  // update_tree_wins(n, wg, gameover), update_tree_losses(n, lg, gameover).  A walk up a deep tree is a
  // chain of cache misses; an expansion needs this one instead of 2 + (winning + doomed children).
  void propagate(int32_t n, int64_t w, int64_t l, int64_t wg, int64_t lg) {
    while (n >= 0 && (w | l | wg | lg) != 0) {
      Node& nd = nodes[n];
      nd.wins += w + wg;
      nd.visits += w + l + wg + lg;
      nd.gw += wg;
      nd.gv += wg + lg;
      n = nd.parent;
      int64_t t = w; w = l; l = t;               // a win here is a loss for the parent
      t = wg; wg = lg; lg = t;
    }
  }
  void update_tree_losses(int32_t n, int64_t amount, bool go) {
    nodes[n].visits += amount;
    if (go) nodes[n].gv += amount;
    if (nodes[n].parent >= 0) update_tree_wins(nodes[n].parent, amount, go);
  }

  // mixed float32 / double comparison of the reference (numpy scalar vs Python float: in float32)
  bool uct_gt(int32_t c, float thr) const {
    return nodes[c].uct_f32 ? static_cast<float>(nodes[c].uct) > thr : nodes[c].uct > static_cast<double>(thr);
  }

  // ---- solver (tree_search_utils.pyx:70-132) ----------------------------------------------------
  void update_win_statuses(int32_t n, int8_t status) {
    if (status == 0) update_tree_losses(n, 1, true);
    nodes[n].ws = status;
    if (nodes[n].parent >= 0) update_win_status_from_children(nodes[n].parent);
  }
  void update_win_status_from_children(int32_t n) {
    if (nodes[n].count < 0) return;
    bool any_f = false, any_t = false, any_n = false, c_any = false, c_t = false, c_n = false;
    for (int32_t c = nodes[n].first; c < nodes[n].first + nodes[n].count; ++c) {
      const int8_t s = nodes[c].ws;
      any_f |= s == 0; any_t |= s == 1; any_n |= s == -1;
      if (uct_gt(c, 1.001f)) { c_any = true; c_t |= s == 1; c_n |= s == -1; }
    }
    const bool use_c = !any_f && c_any;
    const bool has_t = use_c ? c_t : any_t, has_n = use_c ? c_n : any_n;
    if (any_f) update_win_statuses(n, 1);
    else if (has_t && !has_n) update_win_statuses(n, 0);
  }
  void backpropagate_num_checked_children(int32_t n) {      // :134-158
    while (n >= 0) {
      int k = 0;
      for (int32_t c = nodes[n].first; c < nodes[n].first + nodes[n].count; ++c) k += nodes[c].checked;
      if (k == nodes[n].count) { nodes[n].checked = 1; n = nodes[n].parent; }
      else { nodes[n].checked = 0; n = -1; }
    }
  }

  // ---- thread counters (:161-242) ----------------------------------------------------------------
  void crement_threads(int32_t n, int amount) {
    if (amount == 0) nodes[n].threads = 0; else nodes[n].threads += amount;
    if (nodes[n].parent >= 0) update_num_children_being_checked(nodes[n].parent);
  }
  void update_num_children_being_checked(int32_t n) {
    while (n >= 0 && nodes[n].count >= 0) {
      int k = 0;
      for (int32_t c = nodes[n].first; c < nodes[n].first + nodes[n].count; ++c)
        k += (nodes[c].threads > 0 && nodes[c].count < 0) || nodes[c].being_checked || nodes[c].ws != -1;
      const uint8_t previous = nodes[n].being_checked;
      nodes[n].being_checked = k == nodes[n].count;
      if (previous != nodes[n].being_checked && nodes[n].parent >= 0) n = nodes[n].parent; else break;
    }
  }

  // ---- selection (:251-467) -----------------------------------------------------------------------
  double get_uct(int32_t c, int64_t parent_visits) const {
    const int64_t nw = nodes[c].wins, nv = nodes[c].visits, tw = nodes[c].gw, tg = nodes[c].gv;
    // max(0.10, UCT_multiplier - 1): a float32 multiplier is subtracted and compared in float32
    double nn_prob;
    if (nodes[c].uct_f32) {
      const float d = static_cast<float>(nodes[c].uct) - 1.0f;
      nn_prob = d > 0.10f ? static_cast<double>(d) : 0.10;
    } else {
      const double d = nodes[c].uct - 1.0;
      nn_prob = d > 0.10 ? d : 0.10;
    }
    double rate;
    if (tg > 10 && nodes[c].black != nodes[root].black) rate = static_cast<double>(tg - tw) / static_cast<double>(tg) + 1.0;
    else rate = static_cast<double>(nv - nw) / static_cast<double>(nv);
    const int64_t sib = parent_visits - nv > 1 ? parent_visits - nv : 1;
    return rate + 1.0 * nn_prob * (std::sqrt(static_cast<double>(sib)) / static_cast<double>(1 + nv));
  }
  int32_t find_best_uct_child(int32_t n) const {
    const int32_t lo = nodes[n].first, hi = nodes[n].first + nodes[n].count;
    auto ok_threads = [&](int32_t c) { return nodes[c].threads <= 0 || nodes[c].count >= 0; };
    auto open_ = [&](int32_t c) { return nodes[c].ws == -1 && !nodes[c].being_checked; };
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
    auto admit = [&](int stage, int32_t c) -> bool {
      switch (stage) {
        case 0: return open_(c) && ok_threads(c);                                            // solved node
        case 1: return !(nodes[c].checked || nodes[c].being_checked) && ok_threads(c);
        case 2: return !nodes[c].checked && ok_threads(c);
        case 10: return open_(c) && uct_gt(c, first_thr) && ok_threads(c);                   // open node
        case 11: return open_(c) && uct_gt(c, second_thr) && (nodes[c].gw <= 1000 || nodes[c].gw * 2 < nodes[c].gv) && ok_threads(c);
        case 12: return open_(c) && uct_gt(c, second_thr) && ok_threads(c);
        default: return open_(c) && ok_threads(c);                                           // 13
      }
    };
    int stages[4];
    int ns = 0;
    if (solved) { stages[0] = 0; stages[1] = 1; stages[2] = 2; ns = 3; }
    else if (n == root) { stages[0] = 10; stages[1] = 11; stages[2] = 12; stages[3] = 13; ns = 4; }
    else { stages[0] = 10; stages[1] = 13; ns = 2; }
    for (int s = 0; s < ns; ++s) {
      int32_t best = -1;
      double best_v = 0;
      const int64_t pv = nodes[n].visits;
      for (int32_t c = lo; c < hi; ++c) {
        if (!admit(stages[s], c)) continue;
        const double v = get_uct(c, pv);
        if (best < 0 || v > best_v) { best = c; best_v = v; }      // the FIRST maximum wins
      }
      if (best >= 0) return best;
    }
    return -1;
  }
  int32_t choose_uct_or_best_child(int32_t n) const { return nodes[n].count == 1 ? nodes[n].first : find_best_uct_child(n); }

  // ---- expansion ----------------------------------------------------------------------------------
  static void apply_move(const int8_t* board, bool blk, int idx, int8_t* out) {
    memcpy(out, board, 64);
    const int r = idx / 22, rem = idx - 22 * r, c = (rem + 1) / 3, d = rem - 3 * c;
    int frm = r * 8 + c, to = (r + 1) * 8 + c + d;
    if (blk) { frm = 63 - frm; to = 63 - to; }
    out[to] = out[frm];
    out[frm] = 0;
  }
  // update_parent -> get_child x k -> assign_children with the seeds of one evaluated position.
  // `next` receives what the reference's assign_children returns (next-frontier candidates).
  void expand_one(int32_t n, const uint8_t* idx, const float* prior, int k, const uint8_t* flags,
                  const int32_t* stats, bool threat, std::vector<int32_t>& next) {
    next.clear();
    ++evaluations;
    const bool blk = nodes[n].black;
    const int32_t base = static_cast<int32_t>(nodes.size());
    int64_t up_w = 0, up_l = 0, up_wg = 0, up_lg = 0;   // what this expansion adds to n and its ancestors
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
          if (f & kSeedGameSaving) { nodes[c].visits = st[0]; nodes[c].wins = st[1]; nodes[c].gv = st[2]; nodes[c].gw = st[3]; next.push_back(c); }
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
        for (int j = 0; j < k; ++j) next.push_back(base + j);
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
  }

  // expand_descendants_to_depth_wrt_NN, resumable: runs until an evaluation is needed (returns the
  // node) or the expansion is over (returns -1).
  int32_t expand_run() {
    while (ex_depth < ex_limit) {
      if (!ex_no_root && nodes[root].ws == -1) { if (nodes[ex_leaf].count < 0 && nodes[ex_leaf].ws == -1) frontier.push_back(ex_leaf); }
      else if (nodes[ex_leaf].count < 0 && !nodes[ex_leaf].gameover) frontier.push_back(ex_leaf);
      if (!frontier.empty()) {                     // (the reference's 0 < len < 256; a frontier has one node)
        ex_limit = 800;
        const int32_t p = frontier[0];
        if (nodes[p].threads <= 1 && nodes[p].count < 0) {     // do_updates_in_the_same_process
          if (nodes[p].threads <= 0) crement_threads(p, 1);
          pending = p;
          return p;                                // -> policy_net.evaluate
        }
        crement_threads(p, -1);
        frontier.clear();
        ++ex_depth;
      } else {
        ex_depth = ex_limit;
        crement_threads(ex_leaf, 0);
      }
    }
    return -1;
  }
  void expand_begin(int32_t leaf, int depth, int limit) {
    ex_leaf = leaf; ex_depth = depth; ex_limit = limit;
    frontier.clear();
  }
  // continuation after the evaluation of `pending`
  void expand_resume(const uint8_t* idx, const float* prior, int k, const uint8_t* flags, const int32_t* stats,
                     bool threat) {
    std::vector<int32_t> kids;
    expand_one(pending, idx, prior, k, flags, stats, threat, kids);
    frontier.clear();
    if (ex_depth + 1 < ex_limit && !kids.empty()) {
      const int32_t c0 = kids[0];
      if (nodes[c0].ws == -1 && nodes[c0].threads <= 0 && !nodes[c0].being_checked) frontier.push_back(c0);
    }
    ++ex_depth;
    if (ex_depth < ex_limit)
      for (int32_t c : frontier) crement_threads(c, 1);
    pending = -1;
  }

  // ---- the per-tree driver: advance until an evaluation is needed or `max_sims` are done ---------
  // Returns the node to evaluate, or -1 (finished / budget reached).
  int32_t advance(int64_t max_sims) {
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
      if (!expanding) { ex_depth = ex_limit = 0; frontier.clear(); }
    }
  }

  // ---- root reuse between moves: assign_root's "Reused old tree" branch (expansion_MCTS_functions.pyx:
  // 236-247,253-257) + reset_thread_flag (:197-207).  Returns false if the root has no such child.
  bool advance_root(int move_index) {
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
    // copy, so that the children of a node stay contiguous.  Thread bookkeeping is cleared on the way.
    board_of(kid);                               // its position, before its parent disappears
    std::vector<Node> kept;
    std::vector<int8_t> kept_boards;
    std::vector<int32_t> order(1, kid);
    kept.reserve(16384);
    for (size_t i = 0; i < order.size(); ++i) {
      Node nd = nodes[order[i]];
      nd.threads = 0;
      nd.being_checked = 0;
      if (nd.board >= 0) {
        const int8_t* b = &boards[static_cast<size_t>(nd.board) * 64];
        nd.board = static_cast<int32_t>(kept_boards.size() / 64);
        kept_boards.insert(kept_boards.end(), b, b + 64);
      }
      if (nd.count >= 0) {
        const int32_t new_first = static_cast<int32_t>(order.size());
        for (int32_t c = nd.first; c < nd.first + nd.count; ++c) order.push_back(c);
        nd.first = new_first;
      }
      kept.push_back(nd);
    }
    for (size_t i = 0; i < kept.size(); ++i)     // parents: children of kept[i] are kept[first .. first + count)
      if (kept[i].count >= 0)
        for (int32_t c = kept[i].first; c < kept[i].first + kept[i].count; ++c) kept[c].parent = static_cast<int32_t>(i);
    kept[0].parent = -1;
    nodes.swap(kept);
    boards.swap(kept_boards);
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
  static Num num_op(Num a, Num b, char op) {
    if (a.f32 || b.f32) {
      const float x = static_cast<float>(a.v), y = static_cast<float>(b.v);
      const float r = op == '+' ? x + y : op == '-' ? x - y : x / y;
      return {static_cast<double>(r), true};
    }
    return {op == '+' ? a.v + b.v : op == '-' ? a.v - b.v : a.v / b.v, false};
  }
  static bool num_gt(Num a, Num b) {
    return (a.f32 || b.f32) ? static_cast<float>(a.v) > static_cast<float>(b.v) : a.v > b.v;
  }
  Num num_uct(int32_t c) const { return {nodes[c].uct, nodes[c].uct_f32 != 0}; }
  Num loss_rate_of(int32_t c) const {   // true_losses / max(1, gameover_visits) + (UCT_multiplier - 1) / 3
    const int64_t losses = nodes[c].gv - nodes[c].gw, total = nodes[c].gv > 1 ? nodes[c].gv : 1;
    const Num rate = {static_cast<double>(losses) / static_cast<double>(total), false};
    return num_op(rate, num_op(num_op(num_uct(c), Num{1.0, false}, '-'), Num{3.0, false}, '/'), '+');
  }
  int32_t best_move() const {
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
    std::vector<int32_t> filtered;
    for (int pass = 0; pass < 3 && filtered.empty(); ++pass)
      for (int32_t c = lo; c < lo + k; ++c) {
        if (nodes[c].ws == 1) continue;
        const bool keep = pass == 2 || num_gt(num_uct(c), Num{pass == 0 ? first_thr : second_thr, false}) ||
                          static_cast<double>(losses(c)) > action_value_threshold;
        if (keep) filtered.push_back(c);
      }
    if (filtered.empty()) return lo;
    int32_t best = filtered[0];
    for (int32_t c : filtered) if (losses(c) > losses(best)) best = c;            // max(): the first maximum
    const int64_t highest = losses(best);
    Num best_rate = loss_rate_of(best);
    for (int32_t c : filtered) {
      const Num rate = loss_rate_of(c);
      if (num_gt(rate, best_rate) && num_gt(num_op(rate, best_rate, '-'), Num{0.05, false}) && highest != 0 &&
          static_cast<double>(losses(c)) / static_cast<double>(highest) > .30 && nodes[c].ws != 1) {
        best = c;
        best_rate = rate;
      }
    }
    return best;
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
  std::vector<int32_t> need;                       // node each tree is waiting on (-1 = none)
};

}  // namespace wann_tree
