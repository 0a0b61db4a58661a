/* leela_oracle.c — CPU restatement of the reference's policy/value evaluation path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker. The product path (leela_b200/csrc) has no CPU fallback.
 *
 * Parity pinning: the reference ships no golden vectors for the 128-wide policy net or the
 * value net (its only known-answer test, GTP.cpp:105-125, pins the 192-wide OpenCL net whose
 * weights are missing). This restatement is pinned instead against outputs of the reference
 * ITSELF, compiled from /root/reference by oracle/ref/Makefile and run on identical synthetic
 * weights and positions: tests/golden/ref_golden.npz (made by tests/golden/make_golden.py),
 * checked by tests/test_oracle_vs_reference.py.
 *
 * Each function cites the reference lines it follows. Arithmetic is fp32 like the
 * reference's; the GEMM is a plain blocked C loop instead of OpenBLAS' cblas_sgemm
 * (third-party, OpenBLAS 0.3.15 in this image; summation order differs, so agreement with
 * the reference is to ~1e-6 relative, not bit-exact).
 *
 * `emulate` flags reproduce the roundings the B200 path applies so that kernel exactness
 * (tier A) can be separated from fp16 model fidelity (tier B):
 *   LB2O_ROUND_W    trunk conv weights rounded to fp16 (tensor-core B operand)
 *   LB2O_ROUND_ACT  trunk activations rounded to fp16 when stored between layers
 *   LB2O_ROUND_LAST also round the last trunk layer's output (it feeds the CUDA-core heads)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LB2O_P 361
#define LB2O_MAXC 256   /* widest layer any caller uses (the OpenCL-build policy net is 192 wide, Network.cpp:55-80) */
#define LB2O_ROUND_W 1
#define LB2O_ROUND_ACT 2
#define LB2O_ROUND_LAST 4

static inline float round_f16(float x) { return (float)(_Float16)x; }

/* Network::rotate_nn_idx, Network.cpp:1348-1379: symmetry >= 4 swaps x,y first, then
 * s&3: 0 identity, 1 flip y, 2 flip x, 3 both. */
int lb2o_rotate_nn_idx(int vertex, int symmetry) {
    int x = vertex % 19, y = vertex / 19;
    if (symmetry >= 4) { int t = x; x = y; y = t; symmetry -= 4; }
    int nx = x, ny = y;
    if (symmetry == 1) { ny = 18 - y; }
    else if (symmetry == 2) { nx = 18 - x; }
    else if (symmetry == 3) { nx = 18 - x; ny = 18 - y; }
    return ny * 19 + nx;
}

/* Network::rev_rotate_nn_idx, Network.cpp:1341-1346: inverse table {0,1,2,3,4,6,5,7}. */
int lb2o_rev_rotate_nn_idx(int vertex, int symmetry) {
    static const int invert[8] = {0, 1, 2, 3, 4, 6, 5, 7};
    return lb2o_rotate_nn_idx(vertex, invert[symmetry]);
}

/* Plane expansion, Network.cpp:765-773 / 691-699:
 * in[c][h][w] = planes[c][rotate_nn_idx(h*19+w, rotation)], planes packed one uint32 per
 * board point (bit c = plane c). */
void lb2o_expand_planes(const uint32_t* packed, int rotation, float* out /*[32][361]*/) {
    for (int i = 0; i < LB2O_P; i++) {
        uint32_t w = packed[lb2o_rotate_nn_idx(i, rotation)];
        for (int c = 0; c < 32; c++) out[c * LB2O_P + i] = (float)((w >> c) & 1u);
    }
}

/* im2col<channels, filter_size>, Im2Col.h:8-50: zero-padded unfold of [C][19][19] into
 * [C*k*k][361], row order (c, kernel_row, kernel_col). */
void lb2o_im2col(int k, int channels, const float* in, float* col) {
    const int pad = k / 2;
    for (int c = 0; c < channels; c++)
        for (int kr = 0; kr < k; kr++)
            for (int kc = 0; kc < k; kc++) {
                float* dst = col + (size_t)((c * k + kr) * k + kc) * LB2O_P;
                for (int y = 0; y < 19; y++) {
                    int iy = y - pad + kr;
                    for (int x = 0; x < 19; x++) {
                        int ix = x - pad + kc;
                        dst[y * 19 + x] = (iy >= 0 && iy < 19 && ix >= 0 && ix < 19)
                                              ? in[c * LB2O_P + iy * 19 + ix] : 0.0f;
                    }
                }
            }
}

/* C[M][N] = A[M][K] * B[K][N], row-major, alpha 1 beta 0 — the cblas_sgemm call of
 * Network.cpp:375-380 (M = outputs, N = 361, K = C*k*k). */
void lb2o_sgemm(int M, int N, int K, const float* A, const float* B, float* C) {
    int i = 0;
    for (; i + 4 <= M; i += 4) {
        float* restrict c0 = C + (size_t)i * N; float* restrict c1 = c0 + N;
        float* restrict c2 = c1 + N; float* restrict c3 = c2 + N;
        memset(c0, 0, sizeof(float) * 4 * (size_t)N);
        for (int k = 0; k < K; k++) {
            const float a0 = A[(size_t)i * K + k], a1 = A[(size_t)(i + 1) * K + k];
            const float a2 = A[(size_t)(i + 2) * K + k], a3 = A[(size_t)(i + 3) * K + k];
            const float* restrict b = B + (size_t)k * N;
            for (int j = 0; j < N; j++) {
                float bv = b[j];
                c0[j] += a0 * bv; c1[j] += a1 * bv; c2[j] += a2 * bv; c3[j] += a3 * bv;
            }
        }
    }
    for (; i < M; i++) {
        float* c0 = C + (size_t)i * N;
        memset(c0, 0, sizeof(float) * (size_t)N);
        for (int k = 0; k < K; k++) {
            const float a0 = A[(size_t)i * K + k];
            const float* b = B + (size_t)k * N;
            for (int j = 0; j < N; j++) c0[j] += a0 * b[j];
        }
    }
}

static inline float elu(float v) { return v > 0.0f ? v : 1.0f * (expf(v) - 1.0f); }

/* convolve<k, C, O>, Network.cpp:345-393: im2col, sgemm with OIHW weights viewed as
 * [O][C*k*k], then out = ELU(out + bias[o]) on every layer. round_w / round_out apply the
 * fp16 roundings of the B200 path (0 = pure fp32 like the reference). */
void lb2o_convolve(int k, int channels, int outputs, const float* in, const float* w,
                   const float* b, float* out, int round_w, int round_out) {
    const int kdim = k * k * channels;
    float* col = (float*)malloc(sizeof(float) * (size_t)kdim * LB2O_P);
    float* wr = NULL;
    lb2o_im2col(k, channels, in, col);
    if (round_w) {
        wr = (float*)malloc(sizeof(float) * (size_t)kdim * outputs);
        for (size_t i = 0; i < (size_t)kdim * outputs; i++) wr[i] = round_f16(w[i]);
        w = wr;
    }
    lb2o_sgemm(outputs, LB2O_P, kdim, w, col, out);
    for (int o = 0; o < outputs; o++)
        for (int p = 0; p < LB2O_P; p++) {
            float v = elu(b[o] + out[o * LB2O_P + p]);
            out[o * LB2O_P + p] = round_out ? round_f16(v) : v;
        }
    free(col);
    free(wr);
}

/* innerproduct<inputs, outputs>, Network.cpp:395-423: y = W x (sgemv, W row-major
 * [outputs][inputs]) + bias, ELU only when outputs > 1. */
void lb2o_innerproduct(int inputs, int outputs, const float* in, const float* w,
                       const float* b, float* out) {
    for (int o = 0; o < outputs; o++) {
        float acc = 0.0f;
        for (int i = 0; i < inputs; i++) acc += w[(size_t)o * inputs + i] * in[i];
        float v = b[o] + acc;
        out[o] = outputs > 1 ? elu(v) : v;
    }
}

/* Network::softmax, Network.cpp:450-469: p_i = exp(x_i/T - max/T) / sum. */
void lb2o_softmax(const float* in, float* out, int n, float temperature) {
    float alpha = in[0];
    for (int i = 1; i < n; i++) if (in[i] > alpha) alpha = in[i];
    alpha /= temperature;
    float denom = 0.0f;
    for (int i = 0; i < n; i++) { out[i] = expf(in[i] / temperature - alpha); denom += out[i]; }
    for (int i = 0; i < n; i++) out[i] /= denom;
}

/* A net = its conv layers in order (k, c_in, c_out, weights OIHW, bias) followed, for the
 * value net, by two inner products. Mirrors the push_convolve / push_innerproduct order of
 * Network::initialize (Network.cpp:206-233). */
typedef struct {
    int n_conv;
    int k[16], c_in[16], c_out[16];
    const float* w[16];
    const float* b[16];
    int n_ip;
    int ip_in[4], ip_out[4];
    const float* ip_w[4];
    const float* ip_b[4];
} lb2o_net;

static void run_trunk(const lb2o_net* net, const uint32_t* packed, int rotation, int emulate,
                      float* buf_a, float* buf_b, float** last) {
    lb2o_expand_planes(packed, rotation, buf_a);
    float* in = buf_a; float* out = buf_b;
    for (int l = 0; l < net->n_conv; l++) {
        int is_head = (net->c_out[l] == 1);           /* 1-output conv runs on CUDA cores, fp32 weights */
        int is_last_trunk = (l == net->n_conv - 2);
        int rw = !is_head && (emulate & LB2O_ROUND_W);
        int ro = !is_head && (emulate & LB2O_ROUND_ACT) && (!is_last_trunk || (emulate & LB2O_ROUND_LAST));
        lb2o_convolve(net->k[l], net->c_in[l], net->c_out[l], in, net->w[l], net->b[l], out, rw, ro);
        float* t = in; in = out; out = t;
    }
    *last = in;
}

/* Network::get_scored_moves_internal, Network.cpp:742-832, for one position, WITHOUT the
 * EMPTY-point filter (that needs the board and stays on the host side): expands planes under
 * `rotation`, runs the 13 convs, softmax with temperature, and un-rotates:
 * probs[idx] = softmax[rev_rotate_nn_idx(idx, rotation)] (Network.cpp:820-823).
 * logits_out (optional) gets the 361 pre-softmax outputs in network orientation. */
void lb2o_policy_forward(const lb2o_net* net, const uint32_t* packed, int rotation,
                         float temperature, int emulate, float* probs, float* logits_out) {
    float* a = (float*)malloc(sizeof(float) * LB2O_MAXC * LB2O_P);
    float* b = (float*)malloc(sizeof(float) * LB2O_MAXC * LB2O_P);
    float sm[LB2O_P];
    float* last;
    run_trunk(net, packed, rotation, emulate, a, b, &last);
    if (logits_out) memcpy(logits_out, last, sizeof(float) * LB2O_P);
    lb2o_softmax(last, sm, LB2O_P, temperature);
    for (int idx = 0; idx < LB2O_P; idx++) probs[idx] = sm[lb2o_rev_rotate_nn_idx(idx, rotation)];
    free(a); free(b);
}

/* Network::get_value_internal, Network.cpp:676-740: 12 convs, innerproduct<361,256> (ELU),
 * innerproduct<256,1> (linear), winrate = (1 + tanh(x)) / 2 for the side to move. */
float lb2o_value_forward(const lb2o_net* net, const uint32_t* packed, int rotation, int emulate) {
    float* a = (float*)malloc(sizeof(float) * LB2O_MAXC * LB2O_P);
    float* b = (float*)malloc(sizeof(float) * LB2O_MAXC * LB2O_P);
    float h[256], o[1];
    float* last;
    run_trunk(net, packed, rotation, emulate, a, b, &last);
    lb2o_innerproduct(net->ip_in[0], net->ip_out[0], last, net->ip_w[0], net->ip_b[0], h);
    lb2o_innerproduct(net->ip_in[1], net->ip_out[1], h, net->ip_w[1], net->ip_b[1], o);
    free(a); free(b);
    return (1.0f + tanhf(o[0])) / 2.0f;
}

/* Batch drivers (each position is an independent reference-style
 * batch-1 evaluation, as in the reference where batch is always 1). */
void lb2o_policy_forward_batch(const lb2o_net* net, const uint32_t* packed, const uint8_t* rotation,
                               int n, float temperature, int emulate, float* probs, float* logits) {
    for (int i = 0; i < n; i++)
        lb2o_policy_forward(net, packed + (size_t)i * LB2O_P, rotation[i], temperature, emulate,
                            probs + (size_t)i * LB2O_P, logits ? logits + (size_t)i * LB2O_P : NULL);
}

void lb2o_value_forward_batch(const lb2o_net* net, const uint32_t* packed, const uint8_t* rotation,
                              int n, int emulate, float* winrate) {
    for (int i = 0; i < n; i++)
        winrate[i] = lb2o_value_forward(net, packed + (size_t)i * LB2O_P, rotation[i], emulate);
}

/* All trunk activations of one position, for per-layer parity: acts[l] = output of conv l
 * ([c_out][361], network orientation, after bias+ELU and any emulated rounding). */
void lb2o_trunk_activations(const lb2o_net* net, const uint32_t* packed, int rotation,
                            int emulate, float* const* acts) {
    float* in = (float*)malloc(sizeof(float) * LB2O_MAXC * LB2O_P);
    lb2o_expand_planes(packed, rotation, in);
    const float* cur = in;
    for (int l = 0; l < net->n_conv; l++) {
        int is_head = (net->c_out[l] == 1);
        int is_last_trunk = (l == net->n_conv - 2);
        int rw = !is_head && (emulate & LB2O_ROUND_W);
        int ro = !is_head && (emulate & LB2O_ROUND_ACT) && (!is_last_trunk || (emulate & LB2O_ROUND_LAST));
        lb2o_convolve(net->k[l], net->c_in[l], net->c_out[l], cur, net->w[l], net->b[l], acts[l], rw, ro);
        cur = acts[l];
    }
    free(in);
}
