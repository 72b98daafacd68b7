# wann_b200.pxd -- Cython declarations of the C ABI in wann_b200.h, for the reference's own
# language (TeoZosa/WaNN is Python/Cython: MCTS.pyx, tree_builder.pyx, policy_net_utils.pyx).
# `cimport wann_b200` from a .pyx built with include_dirs=[<repo>/include] and linked against
# libwann_b200.so; see examples/cython_binding/ and INTEGRATION.md section 2.
from libc.stdint cimport int8_t, uint8_t, int32_t, int64_t, uint32_t, uint64_t

cdef extern from "wann_b200.h" nogil:
    enum: WANN_N_MOVES
    enum: WANN_MAX_LEGAL
    enum: WANN_OK
    enum: WANN_E_INVALID
    enum: WANN_E_NO_DEVICE
    enum: WANN_E_CUDA
    enum: WANN_E_KERNEL
    enum: WANN_MODE_BF16
    enum: WANN_MODE_FP32
    enum: WANN_MODE_FP16
    enum: WANN_MODE_FP32_TC
    enum: WANN_MODE_FP16C
    enum: WANN_MEM_HOST
    enum: WANN_MEM_DEVICE
    enum: WANN_PLANES_I8
    enum: WANN_PLANES_F32

    ctypedef struct wann_weights:
        const float* conv_w[5]
        const float* conv_b[5]
        const float* fc_w
        const float* fc_b
        int final_kernel

    ctypedef struct wann_ctx:
        pass

    int wann_create(const wann_weights* weights, int device, int mode, wann_ctx** out_ctx)
    int wann_set_weights(wann_ctx* ctx, const wann_weights* weights)
    int wann_destroy(wann_ctx* ctx)
    const char* wann_last_error()
    int wann_encode(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move, int64_t n,
                    int8_t* planes, int mem, void* stream)
    int wann_legal_mask(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move, int64_t n,
                        uint8_t* mask, int mem, void* stream)
    int wann_eval_probs(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move, int64_t n,
                        float* probs, float* logits, int mem, void* stream)
    int wann_eval_planes(wann_ctx* ctx, const void* planes, int planes_dtype, int64_t n,
                         float* probs, float* logits, int mem, void* stream)
    int wann_eval_topk(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move, int64_t n, int k,
                       uint8_t* idx, float* prior, uint8_t* n_legal, int mem, void* stream)
    int wann_eval_expand(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move, int64_t n, int k,
                         uint8_t* idx, float* prior, uint8_t* n_legal, int8_t* child_squares,
                         uint8_t* child_flags, int mem, void* stream)
    int wann_eval_expand_seeded(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move, int64_t n, int k,
                                uint8_t* idx, float* prior, uint8_t* n_legal, int8_t* child_squares,
                                uint8_t* child_flags, int32_t* child_stats, int32_t* parent_summary,
                                int mem, void* stream)
    int wann_expand_ranked(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move, int64_t n,
                           const uint8_t* idx, const float* prior, const uint8_t* n_legal,
                           int8_t* child_squares, uint8_t* child_flags, int32_t* child_stats,
                           int32_t* parent_summary, void* stream)
    int wann_tree_step(wann_ctx* ctx, int8_t* squares, uint8_t* black_to_move, int64_t n, int greedy,
                       uint64_t seed, uint64_t step, uint8_t* chosen_idx, uint8_t* finished, void* stream)
    int wann_debug_activations(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move,
                               int64_t n, int layer, float* out, int mem)
    int wann_debug_profile(wann_ctx* ctx, const int8_t* squares, const uint8_t* black_to_move,
                           int64_t n, int64_t* stamps)
    int wann_debug_centering(wann_ctx* ctx, float* offsets, uint8_t* position_of, float* kappa)
    int wann_debug_effective_weights(wann_ctx* ctx, int layer, float* out)
    int wann_gptq_round(const float* H, int n, float* W, int rows, const uint8_t* exact, double damp)
    int wann_kernel_times(wann_ctx* ctx, double* out)
    uint32_t wann_crc32c(const void* data, uint64_t nbytes, uint32_t seed)
    int wann_host_alloc(void** ptr, uint64_t nbytes)
    int wann_host_free(void* ptr)
    int wann_get_stats(wann_ctx* ctx, uint64_t* stats)
    int wann_set_option(wann_ctx* ctx, const char* key, int64_t value)
    int wann_version()

    int wann_host_threads(int n)

    ctypedef struct wann_forest:
        pass

    int wann_forest_create(int64_t n_trees, const int8_t* squares, const uint8_t* black_to_move,
                           const int32_t* heights, wann_forest** out)
    int wann_forest_destroy(wann_forest* forest)
    int wann_forest_create_device(wann_ctx* ctx, int64_t n_trees, const int8_t* squares, const uint8_t* black_to_move,
                                  const int32_t* heights, int32_t node_capacity, wann_forest** out)
    int wann_forest_run(wann_ctx* ctx, wann_forest* forest, int64_t max_simulations, int64_t* out)
    int wann_forest_run_steps(wann_ctx* ctx, wann_forest* forest, int64_t max_simulations, int64_t max_lock_steps,
                              int64_t* out)
    int wann_forest_debug_cycles(wann_forest* forest, int64_t* out)
    int wann_forest_next(wann_forest* forest, int64_t max_simulations, int8_t* squares_out, uint8_t* black_out,
                         int64_t* tree_ids, int64_t* m_out)
    int wann_forest_apply(wann_forest* forest, const uint8_t* idx, const float* prior, const uint8_t* n_legal,
                          const uint8_t* child_flags, const int32_t* child_stats, const int32_t* parent_summary)
    int wann_forest_info(wann_forest* forest, int64_t tree, int64_t* out)
    int wann_forest_advance(wann_forest* forest, int64_t tree, int move_index)
    int wann_forest_play(wann_forest* forest, const int8_t* restart_squares, int restart_black,
                         uint8_t* moves_out, uint8_t* finished_out)
    int64_t wann_forest_dump(wann_forest* forest, int64_t tree, int64_t* rows, int64_t capacity)
