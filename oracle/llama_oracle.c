/*
 * oracle/llama_oracle.c -- TEST INFRASTRUCTURE (parity oracle). See ggml_oracle.h.
 *
 * Whole-model restatement of one LLaMA forward pass as the reference computes it on its ggml CPU path:
 * the node sequence of crates/models/llama/src/lib.rs:166-362 evaluated with the row kernels restated in
 * ggml_oracle.c, the f16 KV cache laid out as crates/models/llama/src/lib.rs:227-239 (K: [n_ctx, n_embd_gqa] per
 * layer by position; V: stored transposed, [n_embd_gqa, n_ctx] per layer), and session state (n_past) as
 * crates/llm-base/src/inference_session.rs:220-295.  Travels to the GPU box (no /root/reference needed).
 */
#include "ggml_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float *attention_norm, *ffn_norm; void *wq, *wk, *wv, *wo, *w1, *w2, *w3; } or_layer;

struct or_llama {
    or_hparams hp;
    void *wte, *output; float *norm;
    or_layer *layers;
    uint16_t *memory_k, *memory_v;
    int n_past;
    float *tap; int tap_layer, tap_stage;
    float rope_base, rope_scale;
};

static void *xmalloc(size_t n) { void *p = malloc(n ? n : 1); if (!p) { fprintf(stderr, "oracle: out of memory (%zu)\n", n); abort(); } return p; }

or_llama *or_llama_new(const or_hparams *hp) {
    or_llama *m = calloc(1, sizeof(*m));
    m->hp = *hp; m->tap_layer = -2; m->tap_stage = 11; m->rope_base = 10000.0f; m->rope_scale = 1.0f;
    const int e = hp->n_embd, f = hp->n_ff, v = hp->n_vocab, t = hp->wtype;
    const int gqa = e / (hp->n_head / hp->n_head_kv);
    m->wte = xmalloc(or_row_bytes(t, e) * v);
    m->output = xmalloc(or_row_bytes(t, e) * v);
    m->norm = xmalloc((size_t)e * 4);
    m->layers = calloc(hp->n_layer, sizeof(or_layer));
    for (int i = 0; i < hp->n_layer; i++) {
        or_layer *L = &m->layers[i];
        L->attention_norm = xmalloc((size_t)e * 4); L->ffn_norm = xmalloc((size_t)e * 4);
        L->wq = xmalloc(or_row_bytes(t, e) * e);   L->wo = xmalloc(or_row_bytes(t, e) * e);
        L->wk = xmalloc(or_row_bytes(t, e) * gqa); L->wv = xmalloc(or_row_bytes(t, e) * gqa);
        L->w1 = xmalloc(or_row_bytes(t, e) * f);   L->w3 = xmalloc(or_row_bytes(t, e) * f);
        L->w2 = xmalloc(or_row_bytes(t, f) * e);
    }
    const size_t n_el = (size_t)e * hp->n_layer * hp->n_ctx;   /* inference_session.rs:156-158 (n_embd, not n_embd_gqa) */
    m->memory_k = calloc(n_el, 2); m->memory_v = calloc(n_el, 2);
    return m;
}

void *or_llama_tensor(or_llama *m, const char *name, size_t *nbytes) {
    const or_hparams *hp = &m->hp;
    const int e = hp->n_embd, f = hp->n_ff, v = hp->n_vocab, t = hp->wtype;
    const int gqa = e / (hp->n_head / hp->n_head_kv);
    void *p = NULL; size_t nb = 0; int il = -1; char sub[64];
    if (!strcmp(name, "tok_embeddings.weight")) { p = m->wte; nb = or_row_bytes(t, e) * v; }
    else if (!strcmp(name, "norm.weight")) { p = m->norm; nb = (size_t)e * 4; }
    else if (!strcmp(name, "output.weight")) { p = m->output; nb = or_row_bytes(t, e) * v; }
    else if (sscanf(name, "layers.%d.%63s", &il, sub) == 2 && il >= 0 && il < hp->n_layer) {
        or_layer *L = &m->layers[il];
        if      (!strcmp(sub, "attention_norm.weight")) { p = L->attention_norm; nb = (size_t)e * 4; }
        else if (!strcmp(sub, "ffn_norm.weight"))       { p = L->ffn_norm; nb = (size_t)e * 4; }
        else if (!strcmp(sub, "attention.wq.weight"))   { p = L->wq; nb = or_row_bytes(t, e) * e; }
        else if (!strcmp(sub, "attention.wk.weight"))   { p = L->wk; nb = or_row_bytes(t, e) * gqa; }
        else if (!strcmp(sub, "attention.wv.weight"))   { p = L->wv; nb = or_row_bytes(t, e) * gqa; }
        else if (!strcmp(sub, "attention.wo.weight"))   { p = L->wo; nb = or_row_bytes(t, e) * e; }
        else if (!strcmp(sub, "feed_forward.w1.weight")) { p = L->w1; nb = or_row_bytes(t, e) * f; }
        else if (!strcmp(sub, "feed_forward.w2.weight")) { p = L->w2; nb = or_row_bytes(t, f) * e; }
        else if (!strcmp(sub, "feed_forward.w3.weight")) { p = L->w3; nb = or_row_bytes(t, e) * f; }
    }
    if (nbytes) *nbytes = nb;
    return p;
}

void or_llama_reset(or_llama *m) { m->n_past = 0; }
void or_llama_set_n_past(or_llama *m, int n) { m->n_past = n; }
void or_llama_set_rope(or_llama *m, float freq_base, float freq_scale) { m->rope_base = freq_base; m->rope_scale = freq_scale; }
void or_llama_set_tap(or_llama *m, float *buf, int il) { m->tap = buf; m->tap_layer = il; m->tap_stage = 11; }
/* stage taps inside layer il (debugging GPU parity): 1 cur after attn rms_norm*gain, 2 q|k|v before rope (q then k then v, each [N][.]),
 * 3 q|k after rope, 4 KQ raw, 5 KQ after scale+mask+softmax, 6 merged KQV, 7 inpFF, 8 cur after ffn norm, 9 w1x|w3x, 10 silu*mul, 11 layer out */
void or_llama_set_tap_stage(or_llama *m, float *buf, int il, int stage) { m->tap = buf; m->tap_layer = il; m->tap_stage = stage; }
#define TAP(stage, ptr, count) do { if (m->tap && m->tap_layer == il && m->tap_stage == (stage)) memcpy(m->tap, (ptr), (size_t)(count) * 4); } while (0)
#define TAP2(stage, p1, c1, p2, c2) do { if (m->tap && m->tap_layer == il && m->tap_stage == (stage)) { memcpy(m->tap, (p1), (size_t)(c1) * 4); memcpy(m->tap + (c1), (p2), (size_t)(c2) * 4); } } while (0)
void *or_llama_kv(or_llama *m, int which, size_t *nbytes) {
    if (nbytes) *nbytes = (size_t)m->hp.n_embd * m->hp.n_layer * m->hp.n_ctx * 2;
    return which ? m->memory_v : m->memory_k;
}

static void mul_rows(float *x, const float *g, int64_t n, int64_t rows) { /* ggml_mul broadcast, LC/ggml.c:8852-8886 */
    for (int64_t r = 0; r < rows; r++) for (int64_t i = 0; i < n; i++) x[r*n + i] *= g[i];
}

int or_llama_eval(or_llama *m, const int32_t *tokens, int N, float *logits_all) {
    const or_hparams *hp = &m->hp;
    const int e = hp->n_embd, f = hp->n_ff, V = hp->n_vocab, t = hp->wtype;
    const int n_head = hp->n_head, n_head_kv = hp->n_head_kv, hd = e / n_head;
    const int gqa = e / (n_head / n_head_kv), n_ctx = hp->n_ctx, n_past = m->n_past, n_kv = n_past + N;
    if (N < 1 || n_kv > n_ctx) return -1;

    float *x   = xmalloc((size_t)N * e * 4), *cur = xmalloc((size_t)N * e * 4), *ff = xmalloc((size_t)N * e * 4);
    float *q   = xmalloc((size_t)N * e * 4), *k = xmalloc((size_t)N * gqa * 4), *v = xmalloc((size_t)N * gqa * 4);
    float *kq  = xmalloc((size_t)n_head * N * n_kv * 4), *kqv = xmalloc((size_t)n_head * N * hd * 4);
    float *h1  = xmalloc((size_t)N * f * 4), *h3 = xmalloc((size_t)N * f * 4);
    uint16_t *q16 = xmalloc((size_t)N * e * 2), *p16 = xmalloc((size_t)n_head * N * n_kv * 2);

    for (int i = 0; i < N; i++)                                                             /* get_rows, llama lib.rs:170 */
        or_dequantize_row(t, (const char *)m->wte + (size_t)tokens[i] * or_row_bytes(t, e), x + (size_t)i * e, e);

    for (int il = 0; il < hp->n_layer; il++) {
        const or_layer *L = &m->layers[il];
        uint16_t *Kl = m->memory_k + (size_t)il * n_ctx * gqa;      /* [n_ctx][gqa] */
        uint16_t *Vl = m->memory_v + (size_t)il * n_ctx * gqa;      /* [gqa][n_ctx] */
        or_rms_norm(x, cur, e, N, 5e-6f);                                                   /* :183 */
        mul_rows(cur, L->attention_norm, e, N);                                             /* :186 */
        TAP(1, cur, (size_t)N * e);
        or_mul_mat(t, L->wq, cur, q, e, e, N);                                              /* :194 */
        or_mul_mat(t, L->wk, cur, k, e, gqa, N);                                            /* :208 */
        or_mul_mat(t, L->wv, cur, v, e, gqa, N);                                            /* :223 */
        if (m->tap && m->tap_layer == il && m->tap_stage == 2) { memcpy(m->tap, q, (size_t)N*e*4); memcpy(m->tap + (size_t)N*e, k, (size_t)N*gqa*4); memcpy(m->tap + (size_t)N*(e+gqa), v, (size_t)N*gqa*4); }
        or_rope(q, hd, n_head, N, n_past, hp->n_rot, 0, m->rope_base, m->rope_scale);                    /* :190-203 */
        or_rope(k, hd, n_head_kv, N, n_past, hp->n_rot, 0, m->rope_base, m->rope_scale);                 /* :204-217 */
        TAP2(3, q, (size_t)N * e, k, (size_t)N * gqa);
        for (int i = 0; i < N; i++)                                                         /* cpy f32->f16, :243-244 */
            for (int c = 0; c < gqa; c++) {
                Kl[(size_t)(n_past + i) * gqa + c] = or_fp32_to_fp16(k[(size_t)i * gqa + c]);
                Vl[(size_t)c * n_ctx + n_past + i] = or_fp32_to_fp16(v[(size_t)i * gqa + c]);
            }
        /* KQ = mul_mat(K, Q): src1 rows go f32 -> f16 in INIT (LC/ggml.c:10504-10520, F16 traits :1650-1656) */
        for (size_t i = 0; i < (size_t)N * e; i++) q16[i] = or_fp32_to_fp16(q[i]);
        #pragma omp parallel for collapse(2) schedule(static)
        for (int h = 0; h < n_head; h++)
            for (int i = 0; i < N; i++) {
                const int hk = h / (n_head / n_head_kv);                                    /* broadcast, LC/ggml.c:10549 */
                float *row = kq + ((size_t)h * N + i) * n_kv;
                for (int j = 0; j < n_kv; j++)
                    row[j] = or_vec_dot_f16(hd, Kl + (size_t)j * gqa + (size_t)hk * hd, q16 + (size_t)i * e + (size_t)h * hd);
            }
        TAP(4, kq, (size_t)n_head * N * n_kv);
        or_scale_mask_soft_max(kq, n_kv, N, n_head, 1.0f / sqrtf((float)e / (float)n_head), n_past);  /* :268-281 */
        TAP(5, kq, (size_t)n_head * N * n_kv);
        for (size_t i = 0; i < (size_t)n_head * N * n_kv; i++) p16[i] = or_fp32_to_fp16(kq[i]);
        #pragma omp parallel for collapse(2) schedule(static)
        for (int h = 0; h < n_head; h++)
            for (int i = 0; i < N; i++) {
                const int hk = h / (n_head / n_head_kv);
                for (int c = 0; c < hd; c++)                                                /* KQV = mul_mat(V, softmax), :284-296 */
                    kqv[((size_t)h * N + i) * hd + c] =
                        or_vec_dot_f16(n_kv, Vl + ((size_t)hk * hd + c) * n_ctx, p16 + ((size_t)h * N + i) * n_kv);
            }
        for (int i = 0; i < N; i++)                                                         /* permute + cpy, :299-307 */
            for (int h = 0; h < n_head; h++)
                memcpy(cur + (size_t)i * e + (size_t)h * hd, kqv + ((size_t)h * N + i) * hd, (size_t)hd * 4);
        TAP(6, cur, (size_t)N * e);
        or_mul_mat(t, L->wo, cur, ff, e, e, N);                                             /* :310 */
        for (size_t i = 0; i < (size_t)N * e; i++) ff[i] = ff[i] + x[i];                    /* inpFF, :314 */
        TAP(7, ff, (size_t)N * e);
        or_rms_norm(ff, cur, e, N, 5e-6f);                                                  /* :318 */
        mul_rows(cur, L->ffn_norm, e, N);                                                   /* :321 */
        TAP(8, cur, (size_t)N * e);
        or_mul_mat(t, L->w3, cur, h3, e, f, N);                                             /* :323 */
        or_mul_mat(t, L->w1, cur, h1, e, f, N);                                             /* :325 */
        TAP2(9, h1, (size_t)N * f, h3, (size_t)N * f);
        or_silu(h1, h1, (int64_t)N * f);                                                    /* :328 */
        for (size_t i = 0; i < (size_t)N * f; i++) h1[i] = h1[i] * h3[i];                   /* :330 */
        TAP(10, h1, (size_t)N * f);
        or_mul_mat(t, L->w2, h1, cur, f, e, N);                                             /* :332 */
        for (size_t i = 0; i < (size_t)N * e; i++) x[i] = cur[i] + ff[i];                   /* :334 */
        TAP(11, x, (size_t)N * e);
    }
    or_rms_norm(x, cur, e, N, 5e-6f);                                                       /* :343 */
    mul_rows(cur, m->norm, e, N);                                                           /* :346 */
    if (m->tap && m->tap_layer == -1) memcpy(m->tap, cur, (size_t)N * e * 4);
    or_mul_mat(t, m->output, cur, logits_all, e, V, N);                                     /* :352 */
    m->n_past += N;                                                                         /* inference_session.rs:288 */

    free(x); free(cur); free(ff); free(q); free(k); free(v); free(kq); free(kqv); free(h1); free(h3); free(q16); free(p16);
    return 0;
}

void or_llama_free(or_llama *m) {
    if (!m) return;
    for (int i = 0; i < m->hp.n_layer; i++) {
        or_layer *L = &m->layers[i];
        free(L->attention_norm); free(L->ffn_norm); free(L->wq); free(L->wk); free(L->wv); free(L->wo); free(L->w1); free(L->w2); free(L->w3);
    }
    free(m->layers); free(m->wte); free(m->output); free(m->norm); free(m->memory_k); free(m->memory_v); free(m);
}
