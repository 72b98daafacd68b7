/* CPU oracle for the WaNN policy-net leaf-evaluation path -- plain C restatement.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  Never by wann_b200/.
 *
 * PARITY: encode / move index / legal moves are pinned against the compiled reference
 * (oracle/_ref) and tests/golden fixtures.  Logits/probabilities: PARITY UNPINNED --
 * TensorFlow 1.0.0 (the reference's arithmetic) and the weight shard are absent; this
 * file restates the published semantics of the ops at the reference's call sites.
 * All paths below are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NMOVES 155
#define MAXLEGAL 48

/* tools/utils.pyx:563-600 (180-degree rotation for Black) + :842-848 (player/opponent
 * letters).  out: 0 empty, 1 player, 2 opponent in the side-to-move frame. */
static void pov64(const int8_t *sq, int black, int8_t *pov) {
    for (int i = 0; i < 64; ++i) {
        int8_t v = black ? sq[63 - i] : sq[i];
        if (black && v) v = (int8_t)(3 - v);
        pov[i] = v;
    }
}

/* convert_board_to_2d_matrix_POEB tools/utils.pyx:546-602 +
 * generate_binary_plane_POEB :819-866.  planes: int8 [n][8][8][4]. */
void oracle_encode(const int8_t *squares, const uint8_t *black, int64_t n, int8_t *planes) {
    for (int64_t b = 0; b < n; ++b) {
        int8_t pov[64];
        pov64(squares + 64 * b, black[b], pov);
        int8_t *p = planes + 256 * b;
        for (int i = 0; i < 64; ++i) {
            p[4 * i + 0] = pov[i] == 1;
            p[4 * i + 1] = pov[i] == 2;
            p[4 * i + 2] = pov[i] == 0;
            p[4 * i + 3] = 1; /* utils.pyx:828 bias plane */
        }
    }
}

/* Breakthrough_Player/board_utils.pyx:331-400 legality + tools/utils.pyx:605-641 index.
 * mask: uint8 [n][155]. */
void oracle_legal_mask(const int8_t *squares, const uint8_t *black, int64_t n, uint8_t *mask) {
    for (int64_t b = 0; b < n; ++b) {
        int8_t pov[64];
        pov64(squares + 64 * b, black[b], pov);
        uint8_t *m = mask + NMOVES * b;
        memset(m, 0, NMOVES);
        for (int r = 0; r < 7; ++r)
            for (int c = 0; c < 8; ++c) {
                if (pov[r * 8 + c] != 1) continue;
                for (int d = -1; d <= 1; ++d) {
                    int cc = c + d;
                    if (cc < 0 || cc > 7) continue;
                    int8_t t = pov[(r + 1) * 8 + cc];
                    int ok = d == 0 ? (t == 0) : (t != 1);
                    if (ok) m[22 * r + 3 * c + d] = 1;
                }
            }
    }
}

/* tf.nn.conv2d NHWC/HWIO SAME stride 1 + bias_add + relu
 * (Breakthrough_Player/policy_net_utils.pyx:169-176).  in [8][8][cin] -> out [8][8][cout]. */
static void conv_bias_relu(const float *in, int cin, const float *w, const float *bias, int k,
                           int cout, float *out) {
    int p = (k - 1) / 2;
    for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x) {
            float *o = out + (y * 8 + x) * cout;
            for (int co = 0; co < cout; ++co) o[co] = 0.f;
            for (int i = 0; i < k; ++i) {
                int yy = y + i - p;
                if (yy < 0 || yy > 7) continue;
                for (int j = 0; j < k; ++j) {
                    int xx = x + j - p;
                    if (xx < 0 || xx > 7) continue;
                    const float *a = in + (yy * 8 + xx) * cin;
                    const float *wk = w + (size_t)((i * k + j) * cin) * cout;
                    for (int ci = 0; ci < cin; ++ci) {
                        float av = a[ci];
                        const float *wr = wk + (size_t)ci * cout;
                        for (int co = 0; co < cout; ++co) o[co] += av * wr[co];
                    }
                }
            }
            for (int co = 0; co < cout; ++co) {
                float v = o[co] + bias[co];
                o[co] = v > 0.f ? v : 0.f;
            }
        }
}

/* build_policy_net policy_net_utils.pyx:106-124 + output_layer_init :215-239.
 * planes float [n][8][8][4]; weights in TF layout (HWIO, [64][155]); final_k = 3 (shipped
 * checkpoint) or 1 (_128 variant :127-146).  logits/probs [n][155] (either may be NULL);
 * h5 [n][64] optional (post-ReLU 5th map). */
void oracle_forward(const float *planes, int64_t n, const float *w1, const float *b1,
                    const float *w2, const float *b2, const float *w3, const float *b3,
                    const float *w4, const float *b4, const float *w5, const float *b5,
                    int final_k, const float *wfc, const float *bfc, float *logits, float *probs,
                    float *h5) {
#pragma omp parallel
    {
        float *a = (float *)malloc(64 * 192 * sizeof(float));
        float *b = (float *)malloc(64 * 192 * sizeof(float));
#pragma omp for schedule(dynamic, 4)
        for (int64_t i = 0; i < n; ++i) {
            conv_bias_relu(planes + 256 * i, 4, w1, b1, 3, 192, a);
            conv_bias_relu(a, 192, w2, b2, 3, 192, b);
            conv_bias_relu(b, 192, w3, b3, 3, 192, a);
            conv_bias_relu(a, 192, w4, b4, 3, 192, b);
            float h[64];
            conv_bias_relu(b, 192, w5, b5, final_k, 1, h);
            if (h5) memcpy(h5 + 64 * i, h, sizeof h);
            float lg[NMOVES];
            for (int j = 0; j < NMOVES; ++j) lg[j] = 0.f;
            for (int q = 0; q < 64; ++q) /* reshape [-1,64] row-major r*8+c, :216 */
                for (int j = 0; j < NMOVES; ++j) lg[j] += h[q] * wfc[q * NMOVES + j];
            float mx = -INFINITY;
            for (int j = 0; j < NMOVES; ++j) {
                lg[j] += bfc[j];
                if (lg[j] > mx) mx = lg[j];
            }
            if (logits) memcpy(logits + NMOVES * i, lg, sizeof lg);
            if (probs) {
                float s = 0.f, e[NMOVES];
                for (int j = 0; j < NMOVES; ++j) {
                    e[j] = expf(lg[j] - mx);
                    s += e[j];
                }
                for (int j = 0; j < NMOVES; ++j) probs[NMOVES * i + j] = e[j] / s;
            }
        }
        free(a);
        free(b);
    }
}

/* prior gather (tree_search_utils.pyx:979-986) + descending sort (tree_builder.pyx:557);
 * ties -> ascending index.  idx uint8 [n][48] pad 255, prior float [n][48] pad 0. */
void oracle_rank(const float *probs, const uint8_t *mask, int64_t n, uint8_t *idx, float *prior,
                 uint8_t *count) {
    for (int64_t b = 0; b < n; ++b) {
        int k = 0;
        uint8_t *ib = idx + MAXLEGAL * b;
        float *pb = prior + MAXLEGAL * b;
        for (int j = 0; j < MAXLEGAL; ++j) { ib[j] = 255; pb[j] = 0.f; }
        for (int j = 0; j < NMOVES; ++j) {
            if (!mask[NMOVES * b + j]) continue;
            float p = probs[NMOVES * b + j];
            int pos = k;
            while (pos > 0 && pb[pos - 1] < p) { /* stable insertion: strict < keeps lower idx first */
                pb[pos] = pb[pos - 1];
                ib[pos] = ib[pos - 1];
                --pos;
            }
            pb[pos] = p;
            ib[pos] = (uint8_t)j;
            ++k;
        }
        count[b] = (uint8_t)k;
    }
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
