/* wann_b200.h -- C ABI of the B200-native WaNN policy-net leaf evaluator.
 *
 * This library replaces, for ONE path, what the reference (TeoZosa/WaNN) does through a
 * TensorFlow 1.0 CPU session:   board(s) -> feature planes -> 5-layer 3x3 conv stack ->
 * FC-155 -> softmax -> (legal-move gather + descending sort).
 * Citations are into /root/reference.  The reference has no C ABI of its own (it is
 * Python/Cython calling `sess.run`); each entry point names the reference interface it
 * stands in for.  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - plain pointers and sizes; no torch / numpy types.
 *   - every function returns 0 on success, a negative WANN_E_* code otherwise;
 *     wann_last_error() returns a human-readable message for the calling thread.
 *   - `mem` says where ALL data pointers of the call live: WANN_MEM_HOST (pageable or
 *     pinned host memory; the call stages through the context's pinned buffers, runs the
 *     kernels and returns when the outputs are in place) or WANN_MEM_DEVICE (device
 *     pointers on the context's GPU; work is enqueued on `stream` -- a cudaStream_t passed
 *     as void*, NULL = the context's own stream followed by a synchronize).
 *   - board encoding ("squares"): int8[64] per position, index = (rank-1)*8 + file
 *     (a1=0, b1=1, ..., h8=63), value 0 = empty 'e', 1 = white 'w', 2 = black 'b'
 *     (the compact form of the dict-of-dicts board, tools/utils.pyx:518-541).
 *     `black_to_move`: uint8 per position, 0 = White to move, 1 = Black to move
 *     (the `player_color` / node['color'] argument, MCTS.pyx:98-102).
 *   - move indices are the reference's 155-way alphabet (tools/utils.pyx:69-114,605-641),
 *     in the side-to-move frame; 154 = 'no-move'.
 *   - DEVICE buffers are read and written as vectors: `squares` must be 16-byte aligned and every
 *     output buffer 16-byte aligned (anything from cudaMalloc / a torch tensor, and any row slice
 *     of such a buffer, is); misaligned device pointers are refused with WANN_E_INVALID.
 *   - there is NO CPU fallback: without an sm_100 device wann_create fails.
 *   - all entry points are re-entrant; concurrent callers (the reference calls evaluate
 *     from 11 threads without a lock, tree_builder.pyx:186-189) run on separate lanes.
 */
#ifndef WANN_B200_H
#define WANN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WANN_N_MOVES 155   /* tools/utils.pyx:605-641 */
#define WANN_MAX_LEGAL 48  /* 16 pieces x 3 directions */

enum wann_status {
  WANN_OK = 0,
  WANN_E_INVALID = -1,   /* bad argument */
  WANN_E_NO_DEVICE = -2, /* no sm_100 GPU / CUDA failure at init */
  WANN_E_CUDA = -3,      /* CUDA runtime error during a call */
  WANN_E_KERNEL = -4     /* device-side watchdog tripped (see wann_last_error) */
};

enum wann_mode {
  WANN_MODE_BF16 = 0, /* tcgen05 tensor-core path: bf16 operands, fp32 accumulate */
  WANN_MODE_FP32 = 1, /* CUDA-core fp32 path (accuracy mode, validation) */
  WANN_MODE_FP16 = 2, /* tcgen05 path with fp16 operands (same speed as BF16, 3 more mantissa
                         bits; activations saturate at 65504) */
  WANN_MODE_FP32_TC = 3, /* fp32-class accuracy ON the tensor cores: operands split into fp16
                         hi + lo/4096 pairs, 3 MMAs per K-step, fp32 accumulate (about a third
                         of the BF16 throughput, ~20x the CUDA-core FP32 mode) */
  WANN_MODE_FP16C = 4   /* "centred" fp16: the FP16 kernel at the same MMA count with a third of its error
                         (logits and probabilities within 1e-3 of the fp32 graph, north_star's gate) --
                         activations are stored as a - c (c = calibrated per-bucket mean: the sign bit of fp16
                         is no longer wasted on non-negative ReLU outputs; the halo holds -c and the next bias
                         absorbs sum(w*c)), conv1 and conv5 see their weights as fp16 hi + lo pairs in spare
                         K / N slots, and the weights of the 192->192 layers are rounded to fp16 with error
                         compensation against their inputs' second moments (option "gptq", default on).
                         Calibrated at wann_create / wann_set_weights on built-in positions (+0.1 s).  Options
                         "split_layers", "lo_n" and "lo_kgroups" add lo weight stages for the block of chosen
                         192->192 layers whose rounding error matters most (wann_set_option). */
};

enum wann_mem { WANN_MEM_HOST = 0, WANN_MEM_DEVICE = 1 };

enum wann_planes_dtype { WANN_PLANES_I8 = 0, WANN_PLANES_F32 = 1 };

/* Weights in TensorFlow layout, fp32, HOST pointers (copied at wann_create).
 * Names/shapes are those of policy_net_model/model.index:
 *   hidden_layer/{1..5}/weights  HWIO [3,3,4,192] [3,3,192,192]x3 [k,k,192,1]
 *   hidden_layer/{1..5}/bias     [192]x4, [1]
 *   output_layer/weights [64,155], output_layer/bias [155]
 * (Breakthrough_Player/policy_net_utils.pyx:106-124,149-178,215-239). */
typedef struct wann_weights {
  const float* conv_w[5];
  const float* conv_b[5];
  const float* fc_w;
  const float* fc_b;
  int final_kernel; /* 3 = shipped checkpoint; 1 = the `_128` variant (policy_net_utils.pyx:135) */
} wann_weights;

typedef struct wann_ctx wann_ctx;

/* Replaces instantiate_session() / Saver.restore (policy_net_utils.pyx:64-76): uploads the
 * weights to `device`, builds the tensor-core weight images, allocates lanes/workspaces. */
int wann_create(const wann_weights* weights, int device, int mode, wann_ctx** out_ctx);

/* Replace the weights of a live context (a newer checkpoint of the same architecture: the
 * reference's training scripts overwrite policy_net_model/model and a self-play loop reloads it,
 * policy_net_utils.pyx:68-73 Saver.restore).  Blocks until the evaluations in flight have finished;
 * evaluations issued afterwards use the new weights.  final_kernel must equal the context's. */
int wann_set_weights(wann_ctx* ctx, const wann_weights* weights);

/* Replaces sess.close() (MCTS.pyx:132-133). */
int wann_destroy(wann_ctx* ctx);

/* Message for the last failing call made by the calling thread ("" if none). */
const char* wann_last_error(void);

/* Replaces convert_board_to_2d_matrix_POEB + generate_binary_plane_POEB
 * (tools/utils.pyx:546-602, 819-866): planes int8 [n][8][8][4] = {player, opponent, empty, 1}
 * in the side-to-move frame.  Standalone (the eval entry points fuse this step). */
int wann_encode(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move, int64_t n,
                int8_t* planes, int mem, void* stream);

/* Replaces enumerate_legal_moves + index_lookup_by_move
 * (Breakthrough_Player/board_utils.pyx:293-400, tools/utils.pyx:69-114):
 * mask uint8 [n][155], 1 where the move index is legal for the side to move. */
int wann_legal_mask(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move,
                    int64_t n, uint8_t* mask, int mem, void* stream);

/* Replaces NeuralNetsCombined_128.evaluate on board inputs (MCTS.pyx:85-127):
 * encode + sess.run(y_pred, {X: batch}).  probs float [n][155] (softmax over all 155,
 * not renormalised over legal moves); logits float [n][155] optional (may be NULL). */
int wann_eval_probs(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move,
                    int64_t n, float* probs, float* logits, int mem, void* stream);

/* Replaces sess.run(y_pred, feed_dict={X: planes}) for pre-encoded input
 * (evaluate(..., already_converted=True), MCTS.pyx:103-104; call_policy_net,
 * policy_net_utils.pyx:9-16).  planes [n][8][8][4], int8 or float32 per planes_dtype. */
int wann_eval_planes(wann_ctx* ctx, const void* planes, int planes_dtype, int64_t n,
                     float* probs, float* logits, int mem, void* stream);

/* Replaces evaluate + the per-parent consumer: legal enumeration
 * (board_utils.pyx:311-319), prior gather NN_output[child_index]
 * (tree_search_utils.pyx:979-986) and the descending sort (tree_builder.pyx:557).
 * For each position: idx uint8 [n][48] legal move indices, prior float [n][48] their
 * softmax-155 probabilities, sorted by prior descending (ties: ascending index), padded
 * with 255 / 0; n_legal uint8 [n].  idx[.][0] is the policy move of get_best_move
 * (board_utils.pyx:262-277).  DEVICE idx / prior buffers must be 16-byte aligned (rows are
 * written as 16-byte vectors).
 * k: keep only the k best legal children per position -- num_top of get_top_children
 * (board_utils.pyx:279-291), num_to_keep of the pruning branch (tree_builder.pyx:585-609,
 * TreeNode.pyx:18).  k = 0 (like num_top = 0), or any k outside 1..47, keeps every legal child.
 * With 1 <= k <= 47 the entries from k on are padding (255 / 0) and n_legal = min(legal moves, k),
 * so wann_eval_expand / _expand_seeded build and seed the kept children only (the parent summary
 * then covers the kept children, as update_win_status_from_children does after pruning). */
int wann_eval_topk(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move,
                   int64_t n, int k, uint8_t* idx, float* prior, uint8_t* n_legal, int mem,
                   void* stream);

/* wann_eval_topk + child construction in one call (no host round trip between them): for every
 * ranked legal move k < n_legal the child board after the move, child_squares int8 [n][48][64]
 * (absolute frame, same encoding as `squares`; the child is to be moved by the OTHER colour) and
 * child_flags uint8 [n][48]: bit0 = the move reaches the far rank (index > 131: game over, won
 * by the mover), bit1 = capture.  Pad slots are zero.  Replaces get_child /
 * init_unique_child_node_and_board / get_child_attributes (tree_builder.pyx:530-537,623-658),
 * move_piece_update_piece_arrays (board_utils.pyx:90-131) and check_for_winning_move
 * (tree_builder.pyx:902-905). */
int wann_eval_expand(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move,
                     int64_t n, int k, uint8_t* idx, float* prior, uint8_t* n_legal,
                     int8_t* child_squares, uint8_t* child_flags, int mem, void* stream);

/* wann_eval_expand + CHILD SEEDING: everything the reference derives from the priors right after
 * the evaluation, for every ranked legal child -- get_child / check_for_winning_move /
 * set_game_over_values (tree_builder.pyx:530-536,902-914), update_child (capture close to home and
 * the prior -> visits/wins seeds, tree_search_utils.pyx:963-1014), assign_children's "maybe gameover
 * next move" branch (game-saving captures, set_game_over_values_1_step_lookahead,
 * tree_builder.pyx:563-576,916-921) or else eval_children / evaluation_function
 * (tree_search_utils.pyx:812-858,1056-1071), one level of update_tree_wins/_losses (:43-66) and
 * update_win_status_from_children (:86-132).
 *   child_stats    int32 [n][48][4]  visits, wins, gameover_visits, gameover_wins (zeros for pads)
 *   child_flags    uint8 [n][48]     bit0 win, bit1 capture, bit2 capture close to home, bit3 game-
 *                                    saving move, bit4 doomed (loses to the 1-step lookahead; the
 *                                    child's win_status is True), bit5 evaluation_function(child) == 1
 *                                    (win_status of a child: bit0 -> False, bit4 -> True, else None;
 *                                    UCT_multiplier = 1 + prior)
 *   parent_summary int32 [n][8]      visits, wins, gameover_visits, gameover_wins the parent gains,
 *                                    its win_status (-1 None / 0 False / 1 True), threat (opponent one
 *                                    step from the far rank), winning children, doomed children.
 * The propagation ABOVE the parent is the tree's business.  child_squares may be NULL (seeds only).
 * DEVICE child_stats / parent_summary must be 16-byte aligned. */
int wann_eval_expand_seeded(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move,
                            int64_t n, int k, uint8_t* idx, float* prior, uint8_t* n_legal,
                            int8_t* child_squares, uint8_t* child_flags, int32_t* child_stats,
                            int32_t* parent_summary, int mem, void* stream);

/* The child half of wann_eval_expand_seeded on its own, for positions whose ranked children are
 * already known (idx / prior / n_legal as written by wann_eval_topk): child boards and flags
 * (get_child_attributes + move_piece_update_piece_arrays + check_for_winning_move,
 * tree_builder.pyx:623-658,902-905, board_utils.pyx:90-131) when child_squares != NULL, and the
 * child seeds / parent summary (update_child, assign_children, eval_children,
 * tree_search_utils.pyx:963-1071, tree_builder.pyx:538-620) when child_stats != NULL.  DEVICE
 * pointers only, 16-byte aligned; enqueued on `stream` (NULL = the legacy default stream, then a
 * synchronize).  Also what bench.py times to report these two kernels' HBM bandwidth. */
int wann_expand_ranked(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move, int64_t n,
                       const uint8_t* idx, const float* prior, const uint8_t* n_legal,
                       int8_t* child_squares, uint8_t* child_flags, int32_t* child_stats,
                       int32_t* parent_summary, void* stream);

/* One self-play style step for n independent games, entirely on the device (BASELINE config 5:
 * independent game trees, each contributing one leaf per step): evaluate the positions, pick ONE
 * child per game -- greedy != 0: the policy move (get_best_move, board_utils.pyx:262-277);
 * greedy == 0: sampled in proportion to the ranked legal priors with the counter-based RNG
 * u = splitmix64(splitmix64(seed ^ step * 0xD1342543DE82EF95) ^ game) >> 40 -- apply it IN PLACE
 * to `squares` / `black_to_move` (DEVICE pointers), and restart finished games (move index > 131
 * or no legal move) from the opening position.  chosen_idx uint8 [n] (255 = none) and
 * finished uint8 [n] are optional outputs.  Asynchronous on `stream` (NULL = synchronise). */
int wann_tree_step(wann_ctx* ctx, int8_t* squares, uint8_t* black_to_move, int64_t n, int greedy,
                   uint64_t seed, uint64_t step, uint8_t* chosen_idx, uint8_t* finished,
                   void* stream);

/* ---- Forest of array-based search trees (host side; no context, no GPU work of its own) --------
 * The reference's search tree -- TreeNode dicts, select_UCT_child / find_best_UCT_child / get_UCT,
 * expand_descendants_to_depth_wrt_NN, update_tree_wins/_losses, the win-status solver, the
 * checked / being-checked bookkeeping (expansion_MCTS_functions.pyx:300-350, tree_search_utils.pyx:43-467,
 * tree_builder.pyx:122-215,294-355,530-620) -- over flat arrays, many independent trees advancing in
 * lock step (BASELINE config 5).  Each tree does exactly what ONE thread of the reference's search
 * does, suspended where the reference calls policy_net.evaluate (tree_builder.pyx:189):
 *   wann_forest_next   runs every tree until it needs a position evaluated (or has done
 *                      max_simulations calls of run_MCTS_with_expansions_simulation, or is solved) and
 *                      writes the m <= n requested positions, compacted, to squares_out / black_out
 *                      (tree_ids[i] = the tree row i belongs to, optional);
 *   (caller)           ONE wann_eval_expand_seeded call over those m rows (child_squares may be NULL);
 *   wann_forest_apply  attaches the children with the reference's seeds, propagates statistics and
 *                      win statuses, and leaves every tree ready for the next wann_forest_next.
 * A forest is driven by ONE host thread (its entry points parallelise over the trees internally, see
 * wann_host_threads); different forests may be driven by different threads on the same context.
 * Trees grown this way are node-for-node the trees the compiled reference grows single-threaded
 * (tests/golden/ref_tree_golden.npz).  Tie-break: children arrive ranked by prior, equal priors by
 * ascending move index (the reference's order for exact ties depends on its piece-list history). */
/* Host threads used by wann_forest_next / wann_forest_apply (OpenMP over the trees): n > 0 sets the
 * number (launchers such as torchrun export OMP_NUM_THREADS=1), any n returns the current maximum. */
int wann_host_threads(int n);

typedef struct wann_forest wann_forest;
int wann_forest_create(int64_t n_trees, const int8_t* squares, const uint8_t* black_to_move,
                       const int32_t* heights /* ply of each root, may be NULL */, wann_forest** out);
int wann_forest_destroy(wann_forest* forest);
/* The same forest, DEVICE-RESIDENT: tree headers, node records (72 B) and evaluated positions live in
 * the context's GPU memory, `node_capacity` nodes per tree (0 = 16,384; fixed -- a tree that fills its
 * arena stops and reports finished = 2 in wann_forest_info).  The tree code is the host forest's,
 * compiled for the device (array_tree.hpp): one warp per tree, so the trees are node for node the
 * same.  Such a forest is driven by wann_forest_run and wann_forest_play only; _info / _dump copy the
 * tree back.  Destroy it before its context. */
int wann_forest_create_device(wann_ctx* ctx, int64_t n_trees, const int8_t* squares,
                              const uint8_t* black_to_move, const int32_t* heights, int32_t node_capacity,
                              wann_forest** out);
/* The search loop inside the library (expansion_MCTS_functions.pyx:105-119 for many trees at once):
 * every tree does max_simulations simulations.  Host forest: wann_forest_next -> one
 * wann_eval_expand_seeded -> wann_forest_apply per lock step, pinned rows.  Device forest: per lock step
 * forest_next_kernel -> conv stack + fused tail -> seed_children_kernel -> forest_apply_kernel are
 * enqueued on one stream, the number of requested rows stays on the device, and the host only looks
 * at the per-step row counts every 8 steps.  out[0] = lock steps, out[1] = evaluations (optional). */
int wann_forest_run(wann_ctx* ctx, wann_forest* forest, int64_t max_simulations, int64_t out[2]);
/* The same with a budget of lock steps (the reference's search thinks for a fixed TIME, `time_to_think`,
 * expansion_MCTS_functions.pyx:113-119, not for a fixed number of simulations): returns after
 * max_lock_steps steps even if trees are mid-simulation.  wann_forest_run(ctx, forest, 0, ...) afterwards
 * finishes the simulations in flight without starting new ones, so that moves can be played. */
int wann_forest_run_steps(wann_ctx* ctx, wann_forest* forest, int64_t max_simulations, int64_t max_lock_steps,
                          int64_t out[2]);
/* Inspection hook (device-resident forests): the last 64 lock steps, oldest first, out int64 [64][5]:
 * rows requested, then the clock64 cycles the trees spent in forest_next_kernel (sum over trees, max)
 * and in forest_apply_kernel (sum, max). */
int wann_forest_debug_cycles(wann_forest* forest, int64_t* out);
int wann_forest_next(wann_forest* forest, int64_t max_simulations, int8_t* squares_out,
                     uint8_t* black_out, int64_t* tree_ids, int64_t* m_out);
int wann_forest_apply(wann_forest* forest, const uint8_t* idx, const float* prior, const uint8_t* n_legal,
                      const uint8_t* child_flags, const int32_t* child_stats, const int32_t* parent_summary);
/* out[8] = nodes, simulations, evaluations, root visits, root wins, root win_status (-1/0/1),
 * finished (1 = root subtree checked, 2 = a device tree whose arena is full), and the move the search would play now: the move index of the
 * child chosen by randomly_choose_a_winning_move (tree_search_utils.pyx:469-486 with
 * check_for_forced_win :559-582 and choose_best_true_loser :493-557; expansion_MCTS_functions.pyx:129),
 * -1 while the root is unexpanded. */
int wann_forest_info(wann_forest* forest, int64_t tree, int64_t out[8]);
/* Root reuse between moves (assign_root's "Reused old tree" branch + reset_thread_flag,
 * expansion_MCTS_functions.pyx:197-207,236-257): the child of the root reached by `move_index`
 * becomes the root of the tree's next search (its subtree, statistics and win statuses are kept, the
 * thread bookkeeping is cleared, the simulation count restarts).  Fails if the root has no such child.
 * If the root was never expanded (a move that is not in the tree: init_new_root, tree_builder.pyx:
 * 809-897) it is expanded first -- the evaluations this needs are handed out by the following
 * wann_forest_next calls like any others -- and the advance completes when they are done. */
int wann_forest_advance(wann_forest* forest, int64_t tree, int move_index);

/* Self-play step for the whole forest: every tree plays the move its search chose (wann_forest_info's
 * move) -- wann_forest_advance, i.e. the game continues from the chosen child with its subtree reused;
 * a winning move ends the game and the tree restarts from `restart_squares` (the opening position,
 * tools/utils.pyx:518-541).  moves_out uint8 [n] (255 = root not expanded yet), finished_out uint8 [n];
 * both optional.  Between two calls: wann_forest_next / wann_eval_expand_seeded / wann_forest_apply
 * until the simulation budget is used up. */
int wann_forest_play(wann_forest* forest, const int8_t* restart_squares, int restart_black,
                     uint8_t* moves_out, uint8_t* finished_out);

/* Breadth-first dump of one tree, 14 int64 per node: parent row, move index, black to move, visits,
 * wins, gameover_visits, gameover_wins, win_status, gameover, subtree_checked, subtree_being_checked,
 * threads_checking_node, number of children (-1 = not expanded), height.  Returns the number of
 * nodes (call with rows = NULL to size the buffer), -1 on a bad argument. */
int64_t wann_forest_dump(wann_forest* forest, int64_t tree, int64_t* rows, int64_t capacity);

/* Test/inspection hook: activations after hidden layer `layer` (1..5) as float
 * [n][8][8][C] (C = 192, or 1 for layer 5), host or device per `mem`.  In BF16 mode the
 * values are the bf16-rounded activations the next layer consumes (layer 5: fp32). */
int wann_debug_activations(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move,
                           int64_t n, int layer, float* out, int mem);

/* Test/inspection hook (tensor-core mode): clock64 stamps of CTA pair 0 for its first 64
 * layer-steps: for half h (0/1): [4h][k] MMA issue start, [4h+1][k] MMA issue end, [4h+2][k] epilogue start,
 * [4h+3][k] epilogue end (int64 [8][64]).  Host pointers; n <= chunk. */
int wann_debug_profile(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move,
                       int64_t n, int64_t* stamps);

/* Test/inspection hook (WANN_MODE_FP16C): the calibrated centring of the context's weights --
 * offsets[l*24 + (q/64)*8 + q%8] = offset of the stored output of hidden layer l+1 at channel position q
 * (exact fp16 values), position_of[l*192 + c] = position of TF channel c in the kernel's
 * layout, kappa[9] (optional) = what the centring of layer 4 took out of each tap product of
 * hidden_layer/5.  The emulation in tests/ uses it to reproduce the kernel's roundings. */
int wann_debug_centering(wann_ctx* ctx, float offsets[96], uint8_t position_of[768], float kappa[9]);

/* Test/inspection hook (bf16 / fp16 / fp16c modes): the weights of hidden_layer/`layer` (2..4) exactly as the
 * tensor-core kernel multiplies them, fp32 HWIO [3,3,192,192]: the 16-bit operand value, hi + lo where a lo
 * stage covers the entry, and -- WANN_MODE_FP16C with option "gptq" (default) -- the result of the data-aware
 * rounding (wann_b200/csrc/gptq.hpp) instead of round-to-nearest.  The emulation in tests/ is fed with it. */
int wann_debug_effective_weights(wann_ctx* ctx, int layer, float* out);

/* Host utility (no device, no context): the sequential error-compensated rounding of gptq.hpp on its own.
 * H [n][n] symmetric second-moment matrix of the inputs, W [rows][n] weights (in/out: rounded to fp16 values,
 * except where exact[r*n + k] != 0: those entries keep fp32 precision), damp = ridge as a fraction of
 * mean(diag H) (the calibration uses 0.01).  Exists so that the rounding can be tested against its numpy
 * restatement (oracle/oracle.py gptq_round) without a GPU. */
int wann_gptq_round(const float* H, int n, float* W, int rows, const uint8_t* exact, double damp);

/* Measurement hook: with option "time_kernels" = 1 every conv-stack kernel launch is bracketed
 * by CUDA events on its own stream; this call synchronises the device and returns
 * out[0] = summed kernel milliseconds, out[1] = number of launches since the last call. */
int wann_kernel_times(wann_ctx* ctx, double out[2]);

/* CRC-32C (Castagnoli) of a host buffer, continuing from `seed` (0 to start): the checksum
 * TF-V2 checkpoint bundles store per tensor (policy_net_model/model.index).  Host utility. */
uint32_t wann_crc32c(const void* data, uint64_t bytes, uint32_t seed);

/* Pinned host buffers for callers that want copy/compute overlap on the HOST path. */
int wann_host_alloc(void** ptr, uint64_t bytes);
int wann_host_free(void* ptr);

/* Counters since create: [0] positions evaluated, [1] kernel launches, [2] calls,
 * [3] H2D bytes, [4] D2H bytes, [5] merged launches made by the coalescer, [6] calls they
 * served, [7] positions re-evaluated by the margin-refinement pass (counted when a call ends with a
 * host synchronisation). */
int wann_get_stats(wann_ctx* ctx, uint64_t stats[8]);

/* Tunables (optional).  key: "chunk" (positions per internal chunk, default 16384),
 * "coalesce" (0/1: merge concurrent HOST calls of <= 256 positions into one launch -- flat
 * combining, no added latency for a lone caller; the reference calls evaluate with batch 1 from
 * 11 threads), "variant" (debug bits),
 * "time_kernels" (0/1, see wann_kernel_times),
 * "fuse_tail" (bf16/fp16 modes, default 1: FC-155 -> softmax -> legal gather -> rank run on two extra
 * warps inside the conv-stack kernel, one kernel per evaluation; 0 = the stand-alone tail kernel,
 * bit-identical), "zero_copy" (default 1: HOST calls of <= 512 positions read their inputs from and
 * write their outputs to pinned mapped host memory, no DMA copies), "pdl" (-1 auto / 0 / 1:
 * programmatic dependent launch between consecutive evaluations on a stream),
 * "small_batch_halves" (default 1: <= 296 positions run 4-board supertiles), "skew" (0..6),
 * "refine_margin_micro" (bf16/fp16 modes; margin in 1e-6 logit units, 0 = off): positions whose two
 * best LEGAL moves are closer than the margin -- log(prior_1 / prior_2) < margin, the only
 * positions where 16-bit operand rounding can flip the policy move of get_best_move
 * (board_utils.pyx:262-277) -- are evaluated a second time in the same call by the fp32-class
 * split kernel (WANN_MODE_FP32_TC arithmetic) and every output row of theirs is overwritten.
 * Needs board inputs (squares); plane inputs are not refined.
 * "gptq" (WANN_MODE_FP16C, default 1): the weights of the three 192->192 layers are rounded to fp16 with
 * error compensation against the second moments of the layer's inputs on 512 built-in calibration positions
 * (GPTQ-style, wann_b200/csrc/gptq.hpp) instead of to nearest: their rounding error all but disappears from
 * the output at no run-time cost (probability median error 1.0e-3 -> 7.3e-4); 0 = round to nearest.
 * "split_layers" / "lo_n" / "lo_kgroups" (WANN_MODE_FP16C; defaults 0 / 64 / 1): partial hi/lo weight split.
 * The calibration orders the channels of every hidden layer by activation variance, so the weight
 * rounding error of a 192->192 layer is concentrated in its leading input positions (the first of the three
 * 64-channel K-groups carries ~70 % of the input variance) and leading output positions.  In the layers of
 * the mask split_layers (bit l = hidden_layer/(l+1), l = 1..3) every K-block of the first lo_kgroups groups
 * streams a second stage w - fp16(w) for the output positions [0, lo_n) and a second MMA with N = lo_n runs
 * against it: split_layers = 14, lo_n = 64, lo_kgroups = 1 costs +11 % MMAs and meets the 1e-3 gates;
 * lo_n = 192, lo_kgroups = 3 removes a layer's weight error completely (+33 % MMAs per layer). */
int wann_set_option(wann_ctx* ctx, const char* key, int64_t value);

/* Library/ABI version (major*100+minor). */
int wann_version(void);

#ifdef __cplusplus
}
#endif
#endif /* WANN_B200_H */
