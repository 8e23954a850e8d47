/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE, never shipped, never on the product path.
 *
 * Drives the UNMODIFIED reference ggml (LC/ggml.c + LC/k_quants.c, compiled from /root/reference
 * by oracle/Makefile into oracle/_ref/) through its public C API, doing by hand what the
 * reference's Rust layers do:
 *   - crates/llm-base/src/inference_session.rs:114-295  (InferenceSession::new / ::compute)
 *   - crates/models/llama/src/lib.rs:43-140,144-368      (Llama::new / Llama::evaluate)
 *   - crates/ggml/src/context.rs:200-262,636-646         (set_offloading / use_scratch / new_tensor_raw)
 *   - crates/ggml/src/tensor.rs:56-112                   (transfer_to / offload / offload_no_scratch)
 * There is no Rust toolchain in this environment, so the graph construction is restated here in C
 * node for node; every arithmetic kernel that runs is the reference's own.
 *
 * Built twice by oracle/Makefile:
 *   oracle/_ref/libggml_ref.so   -- plain CPU reference (the parity oracle + "reference" CPU baseline)
 *   oracle/_ref/libggml_seam.so  -- the same ggml.c compiled with -DGGML_USE_CUBLAS and linked against
 *                                   libllm_b200.so: the reference graph executor calling OUR
 *                                   ggml_cuda_* seam (the drop-in test).
 */
#include "ggml.h"
#ifdef GGML_USE_CUBLAS
#include "ggml-cuda.h"
#endif

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define RH_MAX_LAYERS 128

typedef struct {
    int32_t n_vocab, n_embd, n_head, n_head_kv, n_layer, n_ff, n_rot, n_ctx;
    int32_t wtype;      /* enum ggml_type of every 2-D weight (quantize_tensors ".*weight", llama lib.rs:390) */
    int32_t use_gpu;    /* ModelParameters::use_gpu (model/mod.rs:197-229); needs the seam build */
    int32_t n_threads;  /* InferenceSessionConfig::n_threads */
    int32_t n_batch;    /* InferenceSessionConfig::n_batch -> CUDA scratch of n_batch MiB (inference_session.rs:146-149) */
} rh_params;

typedef struct {
    struct ggml_tensor *attention_norm, *wq, *wk, *wv, *wo, *ffn_norm, *w1, *w2, *w3;
} rh_layer;

typedef struct {
    rh_params hp;
    struct ggml_context *model_ctx, *session_ctx, *ctx0;
    struct ggml_tensor *wte, *norm, *output;
    rh_layer layers[RH_MAX_LAYERS];
    struct ggml_tensor *memory_k, *memory_v;
    void *eval_buf;  size_t eval_size;
    void *scratch[2]; size_t scratch_size;
    int n_past;
    int can_offload;
    int finalized;
    int rope_custom; float rope_base, rope_scale;   /* ModelParameters::rope_overrides (model/mod.rs:197-229) */
} rh_model;

static size_t rh_tensor_bytes(enum ggml_type t, int64_t ne0, int64_t ne1) {
    return (size_t)(ne0 / ggml_blck_size(t)) * ggml_type_size(t) * (size_t)ne1;
}

/* crates/ggml/src/context.rs:636-646 new_tensor_raw: every tensor created while offloading is on
 * is handed to ggml_cuda_assign_buffers (tensor.rs:87-94). */
static struct ggml_tensor *W(rh_model *m, struct ggml_tensor *t) {
#ifdef GGML_USE_CUBLAS
    if (m->can_offload) ggml_cuda_assign_buffers(t);
#else
    (void)m;
#endif
    return t;
}

static void rh_use_scratch(rh_model *m, int idx) { /* context.rs:212-229: offs always 0 */
    struct ggml_scratch s = {0, 0, NULL};
    if (idx >= 0) { s.size = m->scratch_size; s.data = m->scratch[idx]; }
    ggml_set_scratch(m->ctx0, s);
}

rh_model *rh_llama_new(const rh_params *p) {
    if (p->n_layer > RH_MAX_LAYERS) return NULL;
#ifndef GGML_USE_CUBLAS
    if (p->use_gpu) { fprintf(stderr, "rh_llama_new: use_gpu needs the seam build\n"); return NULL; }
#endif
    rh_model *m = calloc(1, sizeof(*m));
    m->hp = *p;
    const enum ggml_type wt = (enum ggml_type)p->wtype;
    const int n_embd_gqa = p->n_embd / (p->n_head / p->n_head_kv);

    size_t wbytes = 0;
    wbytes += rh_tensor_bytes(wt, p->n_embd, p->n_vocab) * 2;               /* tok_embeddings, output */
    wbytes += (size_t)p->n_embd * 4;                                       /* norm */
    wbytes += (size_t)p->n_layer * (rh_tensor_bytes(wt, p->n_embd, p->n_embd) * 2          /* wq wo */
                                  + rh_tensor_bytes(wt, p->n_embd, n_embd_gqa) * 2         /* wk wv */
                                  + rh_tensor_bytes(wt, p->n_embd, p->n_ff) * 2            /* w1 w3 */
                                  + rh_tensor_bytes(wt, p->n_ff, p->n_embd)                /* w2 */
                                  + (size_t)p->n_embd * 8);                                /* norms */
    wbytes += (size_t)(3 + 9 * p->n_layer) * 512 + (1u << 20);

    struct ggml_init_params ip = { wbytes, NULL, false };
    m->model_ctx = ggml_init(ip);
    struct ggml_context *c = m->model_ctx;
    m->wte    = ggml_new_tensor_2d(c, wt, p->n_embd, p->n_vocab);
    m->norm   = ggml_new_tensor_1d(c, GGML_TYPE_F32, p->n_embd);
    m->output = ggml_new_tensor_2d(c, wt, p->n_embd, p->n_vocab);
    for (int i = 0; i < p->n_layer; i++) {
        rh_layer *L = &m->layers[i];
        L->attention_norm = ggml_new_tensor_1d(c, GGML_TYPE_F32, p->n_embd);
        L->wq = ggml_new_tensor_2d(c, wt, p->n_embd, p->n_embd);
        L->wk = ggml_new_tensor_2d(c, wt, p->n_embd, n_embd_gqa);
        L->wv = ggml_new_tensor_2d(c, wt, p->n_embd, n_embd_gqa);
        L->wo = ggml_new_tensor_2d(c, wt, p->n_embd, p->n_embd);
        L->ffn_norm = ggml_new_tensor_1d(c, GGML_TYPE_F32, p->n_embd);
        L->w1 = ggml_new_tensor_2d(c, wt, p->n_embd, p->n_ff);
        L->w2 = ggml_new_tensor_2d(c, wt, p->n_ff, p->n_embd);
        L->w3 = ggml_new_tensor_2d(c, wt, p->n_embd, p->n_ff);
    }
    return m;
}

/* Name lookup follows the loader's tensor names (llama lib.rs:52-91). Returns the host buffer to fill. */
void *rh_llama_tensor(rh_model *m, const char *name, size_t *nbytes) {
    struct ggml_tensor *t = NULL;
    int il = -1; char sub[64];
    if (!strcmp(name, "tok_embeddings.weight")) t = m->wte;
    else if (!strcmp(name, "norm.weight")) t = m->norm;
    else if (!strcmp(name, "output.weight")) t = m->output;
    else if (sscanf(name, "layers.%d.%63s", &il, sub) == 2 && il >= 0 && il < m->hp.n_layer) {
        rh_layer *L = &m->layers[il];
        if      (!strcmp(sub, "attention_norm.weight")) t = L->attention_norm;
        else if (!strcmp(sub, "attention.wq.weight")) t = L->wq;
        else if (!strcmp(sub, "attention.wk.weight")) t = L->wk;
        else if (!strcmp(sub, "attention.wv.weight")) t = L->wv;
        else if (!strcmp(sub, "attention.wo.weight")) t = L->wo;
        else if (!strcmp(sub, "ffn_norm.weight")) t = L->ffn_norm;
        else if (!strcmp(sub, "feed_forward.w1.weight")) t = L->w1;
        else if (!strcmp(sub, "feed_forward.w2.weight")) t = L->w2;
        else if (!strcmp(sub, "feed_forward.w3.weight")) t = L->w3;
    }
    if (!t) return NULL;
    if (nbytes) *nbytes = ggml_nbytes(t);
    return t->data;
}

#ifdef GGML_USE_CUBLAS
static void rh_transfer_to_gpu(struct ggml_tensor *t) { /* tensor.rs:56-80 */
    t->backend = GGML_BACKEND_GPU;
    ggml_cuda_transform_tensor(t->data, t);
}
#endif

/* Everything Llama::new does after the tensors are loaded + InferenceSession::new. */
int rh_llama_finalize(rh_model *m) {
    const rh_params *p = &m->hp;
#ifdef GGML_USE_CUBLAS
    if (p->use_gpu) {
        /* llama lib.rs:52-57: tok_embeddings stays on the CPU; norm / output / all layer tensors go to params.backend() */
        rh_transfer_to_gpu(m->norm);
        rh_transfer_to_gpu(m->output);
        for (int i = 0; i < p->n_layer; i++) {
            rh_layer *L = &m->layers[i];
            struct ggml_tensor *ts[9] = { L->attention_norm, L->wq, L->wk, L->wv, L->wo, L->ffn_norm, L->w1, L->w2, L->w3 };
            for (int k = 0; k < 9; k++) rh_transfer_to_gpu(ts[k]);
        }
        /* inference_session.rs:146-149 -> accelerator/mod.rs:68-94 */
        ggml_init_cublas();
        ggml_cuda_set_main_device(0);
        float split = 1.0f;
        ggml_cuda_set_tensor_split(&split);
        ggml_cuda_set_scratch_size((size_t)p->n_batch * 1024 * 1024);
    }
#endif
    /* inference_session.rs:127-160, 996-1021: f16 KV of n_layer*n_ctx*n_embd elements each */
    const size_t n_elements = (size_t)p->n_embd * p->n_layer * p->n_ctx;
    struct ggml_init_params ip = { n_elements * 2 * 2 + (size_t)(5 + 10 * p->n_layer) * 256 + 4096, NULL, false };
    m->session_ctx = ggml_init(ip);
    m->memory_k = ggml_new_tensor_1d(m->session_ctx, GGML_TYPE_F16, n_elements);
    m->memory_v = ggml_new_tensor_1d(m->session_ctx, GGML_TYPE_F16, n_elements);
    memset(m->memory_k->data, 0, ggml_nbytes(m->memory_k));
    memset(m->memory_v->data, 0, ggml_nbytes(m->memory_v));
#ifdef GGML_USE_CUBLAS
    if (p->use_gpu) {
        ggml_cuda_assign_buffers_no_scratch(m->memory_k);
        ggml_cuda_assign_buffers_no_scratch(m->memory_v);
    }
#endif
    /* ctx0 backing store + the two host scratch buffers (inference_session.rs:19-28,169-180). Sized for the
     * largest batch this harness will see rather than the fixed 1 GiB / 512 MiB of the reference. */
    const size_t B = (size_t)p->n_batch;
    size_t per = 0;
    per += B * p->n_ff * 4 * 4;                                   /* w1,w3,silu,mul */
    per += B * p->n_embd * 4 * 12;
    per += (size_t)p->n_head * B * p->n_ctx * 4 * 2;              /* KQ (+ one copy for non-inplace variants) */
    per += 64u << 20;
    m->scratch_size = per;
    m->scratch[0] = malloc(per);
    m->scratch[1] = malloc(per);
    m->eval_size = B * p->n_vocab * 4 + B * p->n_embd * 16 + (size_t)p->n_layer * 64 * 512
                 + B * p->n_ff * 40 /* mul_mat work buffer: q8 rows */ + ggml_graph_overhead() + (64u << 20);
    m->eval_buf = malloc(m->eval_size);
    m->finalized = 1;
    return 0;
}

void rh_llama_reset(rh_model *m) { m->n_past = 0; }
/* session rewind / restore: continue from position n with whatever the KV cache holds (the bench's CPU arm installs the cache contents) */
void rh_llama_set_n_past(rh_model *m, int n) { m->n_past = n; }
/* InferenceSessionConfig::n_threads of the next evaluate (bench.py tries several counts on ONE loaded model) */
void rh_llama_set_threads(rh_model *m, int n) { m->hp.n_threads = n; }
/* RoPEOverrides -> op_rope_inplace takes the ggml_rope_custom_inplace branch (crates/ggml/src/context.rs:558-590) */
void rh_llama_set_rope(rh_model *m, float freq_base, float freq_scale) { m->rope_custom = 1; m->rope_base = freq_base; m->rope_scale = freq_scale; }
int  rh_llama_n_past(rh_model *m) { return m->n_past; }

/* One forward pass = InferenceSession::compute(Llama::evaluate builder). Writes all n rows of logits
 * (OutputRequest::all_logits, model/common.rs:22-39) and, if embd_out, the final-norm embeddings. */
int rh_llama_eval(rh_model *m, const int32_t *tokens, int n, float *logits_out, float *embd_out) {
    const rh_params *p = &m->hp;
    if (!m->finalized || n < 1 || n > p->n_batch || m->n_past + n > p->n_ctx) return -1;
    const int n_embd = p->n_embd, n_head = p->n_head, n_head_kv = p->n_head_kv, n_rot = p->n_rot;
    const int ctx_size = p->n_ctx, session_len = m->n_past, input_len = n;
    const int n_embd_gqa = n_embd / (n_head / n_head_kv);
    const int use_gpu = p->use_gpu;

    if (m->ctx0) ggml_free(m->ctx0);                                  /* ctx0.recreate() */
    struct ggml_init_params ip = { m->eval_size, m->eval_buf, false };
    m->ctx0 = ggml_init(ip);
    struct ggml_context *ctx0 = m->ctx0;
    m->can_offload = 0;

    struct ggml_tensor *embd = W(m, ggml_new_tensor_1d(ctx0, GGML_TYPE_I32, input_len));
    struct ggml_tensor *inpL = W(m, ggml_get_rows(ctx0, m->wte, embd));       /* llama lib.rs:170 */
    struct ggml_cgraph *gf = ggml_new_graph(ctx0);
    const size_t ksz = 2, vsz = 2;                                             /* f16 element size */

    for (int il = 0; il < p->n_layer; il++) {
        const rh_layer *L = &m->layers[il];
        m->can_offload = use_gpu;                                              /* :175 should_offload(il) */
        struct ggml_tensor *inpSA = inpL, *cur;
        rh_use_scratch(m, 0);
        cur = W(m, ggml_rms_norm(ctx0, inpL, 5e-6f));                          /* :183, eps crates/ggml/src/lib.rs:132 */
        cur = W(m, ggml_mul(ctx0, cur, L->attention_norm));                    /* :186 */
        struct ggml_tensor *q3 = W(m, ggml_reshape_3d(ctx0, W(m, ggml_mul_mat(ctx0, L->wq, cur)), n_embd / n_head, n_head, input_len));
        struct ggml_tensor *k3 = W(m, ggml_reshape_3d(ctx0, W(m, ggml_mul_mat(ctx0, L->wk, cur)), n_embd / n_head, n_head_kv, input_len));
        struct ggml_tensor *Qcur = W(m, m->rope_custom ? ggml_rope_custom_inplace(ctx0, q3, session_len, n_rot, 0, 1, m->rope_base, m->rope_scale)
                                                       : ggml_rope_inplace(ctx0, q3, session_len, n_rot, 0, 0));       /* :190-203 */
        struct ggml_tensor *Kcur = W(m, m->rope_custom ? ggml_rope_custom_inplace(ctx0, k3, session_len, n_rot, 0, 1, m->rope_base, m->rope_scale)
                                                       : ggml_rope_inplace(ctx0, k3, session_len, n_rot, 0, 0));       /* :204-217 */
        struct ggml_tensor *Vcur = W(m, ggml_transpose(ctx0,
            W(m, ggml_reshape_2d(ctx0, W(m, ggml_mul_mat(ctx0, L->wv, cur)), n_embd_gqa, input_len))));   /* :221-225 */
        struct ggml_tensor *k = W(m, ggml_view_1d(ctx0, m->memory_k, (int64_t)input_len * n_embd_gqa,
            (ksz * n_embd_gqa) * ((size_t)il * ctx_size + session_len)));      /* :227-231 */
        struct ggml_tensor *v = W(m, ggml_view_2d(ctx0, m->memory_v, input_len, n_embd_gqa, (size_t)ctx_size * vsz,
            ((size_t)il * ctx_size) * vsz * n_embd_gqa + (size_t)session_len * vsz));                     /* :233-239 */
        ggml_build_forward_expand(gf, W(m, ggml_cpy(ctx0, Kcur, k)));          /* :243 */
        ggml_build_forward_expand(gf, W(m, ggml_cpy(ctx0, Vcur, v)));          /* :244 */
        struct ggml_tensor *Q = W(m, ggml_permute(ctx0, Qcur, 0, 2, 1, 3));    /* :246 */
        struct ggml_tensor *K = W(m, ggml_permute(ctx0,
            W(m, ggml_reshape_3d(ctx0,
                W(m, ggml_view_1d(ctx0, m->memory_k, (int64_t)(session_len + input_len) * n_embd_gqa,
                                  (size_t)il * ctx_size * ksz * n_embd_gqa)),
                n_embd / n_head, n_head_kv, session_len + input_len)),
            0, 2, 1, 3));                                                      /* :248-262 */
        struct ggml_tensor *KQ = W(m, ggml_mul_mat(ctx0, K, Q));               /* :265 */
        struct ggml_tensor *KQ_scale = W(m, ggml_new_f32(ctx0, 1.0f / sqrtf((float)n_embd / (float)n_head)));  /* :268-270 */
        struct ggml_tensor *KQ_scaled = W(m, ggml_scale_inplace(ctx0, KQ, KQ_scale));                     /* :271 */
        struct ggml_tensor *KQ_masked = W(m, ggml_diag_mask_inf_inplace(ctx0, KQ_scaled, session_len));  /* :274-276 */
        struct ggml_tensor *KQ_soft_max = W(m, ggml_soft_max_inplace(ctx0, KQ_masked));                  /* :279-281 */
        struct ggml_tensor *V = W(m, ggml_view_3d(ctx0, m->memory_v, session_len + input_len, n_embd / n_head, n_head_kv,
            (size_t)ctx_size * vsz, (size_t)ctx_size * vsz * n_embd / n_head,
            (size_t)il * ctx_size * vsz * n_embd_gqa));                        /* :284-294 */
        struct ggml_tensor *KQV = W(m, ggml_mul_mat(ctx0, V, KQ_soft_max));    /* :296 */
        struct ggml_tensor *KQV_merged = W(m, ggml_permute(ctx0, KQV, 0, 2, 1, 3));                      /* :299 */
        cur = W(m, ggml_cpy(ctx0, KQV_merged, W(m, ggml_new_tensor_2d(ctx0, GGML_TYPE_F32, n_embd, input_len))));  /* :302-307 */
        cur = W(m, ggml_mul_mat(ctx0, L->wo, cur));                            /* :310 */
        rh_use_scratch(m, 1);                                                  /* :312 */
        struct ggml_tensor *inpFF = W(m, ggml_add(ctx0, cur, inpSA));          /* :314 */
        cur = W(m, ggml_rms_norm(ctx0, inpFF, 5e-6f));                         /* :318 */
        cur = W(m, ggml_mul(ctx0, cur, L->ffn_norm));                          /* :321 */
        struct ggml_tensor *tmp = W(m, ggml_mul_mat(ctx0, L->w3, cur));        /* :323 */
        cur = W(m, ggml_mul_mat(ctx0, L->w1, cur));                            /* :325 */
        cur = W(m, ggml_silu(ctx0, cur));                                      /* :328 */
        cur = W(m, ggml_mul(ctx0, cur, tmp));                                  /* :330 */
        cur = W(m, ggml_mul_mat(ctx0, L->w2, cur));                            /* :332 */
        cur = W(m, ggml_add(ctx0, cur, inpFF));                                /* :334 */
        inpL = cur;
    }
    rh_use_scratch(m, 0);                                                      /* :340 */
    inpL = W(m, ggml_rms_norm(ctx0, inpL, 5e-6f));                             /* :343 */
    inpL = W(m, ggml_mul(ctx0, inpL, m->norm));                                /* :346 */
    struct ggml_tensor *embedding_result = inpL;
    m->can_offload = 0;                                                        /* :350 */
    inpL = W(m, ggml_mul_mat(ctx0, m->output, inpL));                          /* :352 */
    rh_use_scratch(m, -1);                                                     /* :354 */

    memcpy(embd->data, tokens, (size_t)input_len * 4);                         /* inference_session.rs:254 */
    ggml_build_forward_expand(gf, inpL);                                       /* :257 */
    struct ggml_cplan plan = ggml_graph_plan(gf, p->n_threads);                /* crates/ggml/src/lib.rs:346-377 */
    struct ggml_tensor *work = ggml_new_tensor_1d(ctx0, GGML_TYPE_I8, plan.work_size ? plan.work_size : 1);
    plan.work_data = work->data;
    ggml_graph_compute(gf, &plan);
    m->n_past += input_len;                                                    /* :288 */

    if (logits_out) memcpy(logits_out, inpL->data, (size_t)input_len * p->n_vocab * 4);
    if (embd_out) {
        if (embedding_result->backend == GGML_BACKEND_CPU)
            memcpy(embd_out, embedding_result->data, (size_t)input_len * n_embd * 4);
        else
            memset(embd_out, 0, (size_t)input_len * n_embd * 4); /* device-resident; same limitation as the reference */
    }
    return 0;
}

/* Raw KV cache bytes (host copy only; the get_snapshot path, inference_session.rs:599-646). */
void *rh_llama_kv(rh_model *m, int which, size_t *nbytes) {
    struct ggml_tensor *t = which ? m->memory_v : m->memory_k;
    if (nbytes) *nbytes = ggml_nbytes(t);
    return t->data;
}

void rh_llama_free(rh_model *m) {
    if (!m) return;
#ifdef GGML_USE_CUBLAS
    if (m->hp.use_gpu) {
        /* Context::drop (context.rs:649-661) + InferenceSession::drop (inference_session.rs:659-665) */
        ggml_cuda_free_data(m->norm); ggml_cuda_free_data(m->output);
        for (int i = 0; i < m->hp.n_layer; i++) {
            rh_layer *L = &m->layers[i];
            struct ggml_tensor *ts[9] = { L->attention_norm, L->wq, L->wk, L->wv, L->wo, L->ffn_norm, L->w1, L->w2, L->w3 };
            for (int k = 0; k < 9; k++) ggml_cuda_free_data(ts[k]);
        }
        if (m->memory_k) { ggml_cuda_free_data(m->memory_k); ggml_cuda_free_data(m->memory_v); }
        ggml_cuda_free_scratch();
    }
#endif
    if (m->ctx0) ggml_free(m->ctx0);
    if (m->session_ctx) ggml_free(m->session_ctx);
    if (m->model_ctx) ggml_free(m->model_ctx);
    free(m->eval_buf); free(m->scratch[0]); free(m->scratch[1]);
    free(m);
}

/* ------------------------------------------------------------------------------------------------
 * Unit-level entry points: the reference's own row kernels through ggml_internal_get_type_traits
 * (LC/ggml.c:1645-1743), its weight quantizers (LC/ggml.c:18083-18230) and single-node graphs.
 * ---------------------------------------------------------------------------------------------- */
static void rh_ensure_init(void) { /* first ggml_init fills the fp16 tables (LC/ggml.c:4313-4326) */
    static int done = 0;
    if (!done) { struct ggml_init_params ip = { 1024, NULL, false }; struct ggml_context *c = ggml_init(ip); ggml_free(c); done = 1; }
}

size_t rh_quantize(int type, const float *src, void *dst, int n, int k) {
    int64_t hist[16] = {0};
    rh_ensure_init();
    switch (type) {
        case GGML_TYPE_Q4_0: return ggml_quantize_q4_0(src, dst, n, k, hist);
        case GGML_TYPE_Q4_1: return ggml_quantize_q4_1(src, dst, n, k, hist);
        case GGML_TYPE_Q5_0: return ggml_quantize_q5_0(src, dst, n, k, hist);
        case GGML_TYPE_Q5_1: return ggml_quantize_q5_1(src, dst, n, k, hist);
        case GGML_TYPE_Q8_0: return ggml_quantize_q8_0(src, dst, n, k, hist);
        default: return 0;
    }
}
void rh_from_float(int type, const float *x, void *y, int k) { rh_ensure_init(); ggml_internal_get_type_traits((enum ggml_type)type).from_float(x, y, k); }
void rh_to_float(int type, const void *x, float *y, int k)   { rh_ensure_init(); ggml_internal_get_type_traits((enum ggml_type)type).to_float(x, y, k); }
int  rh_vec_dot_type(int type) { return (int)ggml_internal_get_type_traits((enum ggml_type)type).vec_dot_type; }
void rh_vec_dot(int type, int n, float *s, const void *x, const void *y) { rh_ensure_init(); ggml_internal_get_type_traits((enum ggml_type)type).vec_dot(n, s, x, y); }
size_t rh_type_size(int type) { return ggml_type_size((enum ggml_type)type); }
uint16_t rh_fp32_to_fp16(float x) { return ggml_fp32_to_fp16(x); }
float rh_fp16_to_fp32(uint16_t h) { rh_ensure_init(); return ggml_fp16_to_fp32(h); }

/* Generic single-op graph on host tensors. op: 0 mul_mat(W[type;K,N], X[f32;K,B]) 1 rms_norm(eps) 2 norm 3 soft_max
 * 4 silu 5 gelu 6 rope(n_past,n_dims,mode; x is [ne0,ne1,ne2]) 7 scale+diag_mask_inf+soft_max (the attention chain)
 * 8 get_rows(W[type;K,N], ids[i32;B]) */
int rh_op(int op, int type, const void *a, const float *b, float *out,
          int64_t ne0, int64_t ne1, int64_t ne2, const int32_t *iparams, const float *fparams, int n_threads) {
    rh_ensure_init();
    size_t an = 0, bn = 0, on = 0;
    switch (op) {
        case 0: an = rh_tensor_bytes((enum ggml_type)type, ne0, ne1); bn = (size_t)ne0 * ne2 * 4; on = (size_t)ne1 * ne2 * 4; break;
        case 8: an = rh_tensor_bytes((enum ggml_type)type, ne0, ne1); bn = (size_t)ne2 * 4; on = (size_t)ne0 * ne2 * 4; break;
        default: bn = (size_t)ne0 * ne1 * ne2 * 4; on = bn; break;
    }
    size_t mem = an + bn + on * 3 + (size_t)ne0 * ne2 * 40 + ggml_graph_overhead() + (16u << 20);
    struct ggml_init_params ip = { mem, NULL, false };
    struct ggml_context *c = ggml_init(ip);
    if (!c) return -1;
    struct ggml_tensor *r = NULL;
    if (op == 0) {
        struct ggml_tensor *w = ggml_new_tensor_2d(c, (enum ggml_type)type, ne0, ne1);
        struct ggml_tensor *x = ggml_new_tensor_2d(c, GGML_TYPE_F32, ne0, ne2);
        memcpy(w->data, a, an); memcpy(x->data, b, bn);
        r = ggml_mul_mat(c, w, x);
    } else if (op == 8) {
        struct ggml_tensor *w = ggml_new_tensor_2d(c, (enum ggml_type)type, ne0, ne1);
        struct ggml_tensor *ids = ggml_new_tensor_1d(c, GGML_TYPE_I32, ne2);
        memcpy(w->data, a, an); memcpy(ids->data, b, bn);
        r = ggml_get_rows(c, w, ids);
    } else {
        struct ggml_tensor *x = ggml_new_tensor_3d(c, GGML_TYPE_F32, ne0, ne1, ne2);
        memcpy(x->data, b, bn);
        switch (op) {
            case 1: r = ggml_rms_norm(c, x, fparams[0]); break;
            case 2: r = ggml_norm(c, x); break;
            case 3: r = ggml_soft_max(c, x); break;
            case 4: r = ggml_silu(c, x); break;
            case 5: r = ggml_gelu(c, x); break;
            case 6:
                if (iparams[3]) r = ggml_rope_custom_inplace(c, x, iparams[0], iparams[1], iparams[2], 1, fparams[0], fparams[1]);
                else            r = ggml_rope_inplace(c, x, iparams[0], iparams[1], iparams[2], 0);
                break;
            case 7:
                r = ggml_soft_max_inplace(c, ggml_diag_mask_inf_inplace(c, ggml_scale_inplace(c, x, ggml_new_f32(c, fparams[0])), iparams[0]));
                break;
            default: ggml_free(c); return -2;
        }
    }
    struct ggml_cgraph *gf = ggml_new_graph(c);
    ggml_build_forward_expand(gf, r);
    struct ggml_cplan plan = ggml_graph_plan(gf, n_threads);
    struct ggml_tensor *work = ggml_new_tensor_1d(c, GGML_TYPE_I8, plan.work_size ? plan.work_size : 1);
    plan.work_data = work->data;
    ggml_graph_compute(gf, &plan);
    memcpy(out, r->data, on);
    ggml_free(c);
    return 0;
}
