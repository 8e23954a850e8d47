/*
 * oracle/ref_gpt2.c -- TEST INFRASTRUCTURE (see ref_harness.c): the reference's GPT-2 model driven through the UNMODIFIED ggml C API.
 *
 * Restates node for node:
 *   Gpt2::new        crates/models/gpt2/src/lib.rs:48-123   tensor names "model/wte", "model/h{i}/attn/c_attn/w", ...; every tensor transfer_to(backend)
 *   Gpt2::evaluate   crates/models/gpt2/src/lib.rs:138-329  LayerNorm (norm, mul g, add b), fused qkv + bias, f16 KV cache by position (V is transposed by a
 *                                                            cpy at read time), gelu MLP, lm_head = model/lm_head or wte
 *   InferenceSession::new / ::compute   crates/llm-base/src/inference_session.rs:114-295
   GptNeoX::new / ::evaluate   crates/models/gptneox/src/lib.rs:44-135,156-352,487-515   (arch = 1) same tensor set under the NeoX names, no wpe,
 *                                                            embed_in stays on the CPU, embed_out is mandatory; cont(view_3d) q/k/v per head, rope mode 2 on
 *                                                            n_rot dims, V stored transposed (as LLaMA), parallel residual (BASELINE.json configs[4] geometry)
 * BASELINE.json configs[0] ("GPT-2 117M Q4_0, 32-token prompt on the reference ggml CPU path") runs through this file on the CPU build; the seam
 * build (-DGGML_USE_CUBLAS, linked against libllm_b200.so) sends the same graph through OUR ggml_cuda_* entry points.
 */
#include "ggml.h"
#ifdef GGML_USE_CUBLAS
#include "ggml-cuda.h"
#endif

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define RG_MAX_LAYERS 64

typedef struct {
    int32_t n_vocab, n_ctx, n_embd, n_head, n_layer;
    int32_t wtype;      /* enum ggml_type of the 2-D weights (quantize.rs quantizes every 2-D tensor whose name ends in "w"/"weight"; wpe stays f32 here) */
    int32_t use_gpu, n_threads, n_batch;
    int32_t has_lm_head;
    int32_t arch;       /* 0 GPT-2, 1 GPT-NeoX */
    int32_t n_rot, use_parallel_residual;   /* NeoX: gptneox lib.rs:425 (default true) */
} rg_params;

typedef struct {
    struct ggml_tensor *ln_1_g, *ln_1_b, *ln_2_g, *ln_2_b, *c_attn_attn_w, *c_attn_attn_b, *c_attn_proj_w, *c_attn_proj_b, *c_mlp_fc_w, *c_mlp_fc_b, *c_mlp_proj_w, *c_mlp_proj_b;
} rg_layer;

typedef struct {
    rg_params hp;
    struct ggml_context *model_ctx, *session_ctx, *ctx0;
    struct ggml_tensor *wte, *wpe, *ln_f_g, *ln_f_b, *lm_head;
    rg_layer layers[RG_MAX_LAYERS];
    struct ggml_tensor *memory_k, *memory_v;
    void *eval_buf; size_t eval_size;
    void *scratch[2]; size_t scratch_size;
    int n_past, can_offload, finalized;
} rg_model;

static size_t rg_bytes(enum ggml_type t, int64_t ne0, int64_t ne1) { return (size_t)(ne0 / ggml_blck_size(t)) * ggml_type_size(t) * (size_t)ne1; }

static struct ggml_tensor *G(rg_model *m, struct ggml_tensor *t) {        /* context.rs:636-646: tensors created while offloading is on */
#ifdef GGML_USE_CUBLAS
    if (m->can_offload) ggml_cuda_assign_buffers(t);
#else
    (void)m;
#endif
    return t;
}
static void rg_use_scratch(rg_model *m, int idx) {
    struct ggml_scratch s = {0, 0, NULL};
    if (idx >= 0) { s.size = m->scratch_size; s.data = m->scratch[idx]; }
    ggml_set_scratch(m->ctx0, s);
}

rg_model *rh_gpt2_new(const rg_params *p) {
    if (p->n_layer > RG_MAX_LAYERS) return NULL;
#ifndef GGML_USE_CUBLAS
    if (p->use_gpu) { fprintf(stderr, "rh_gpt2_new: use_gpu needs the seam build\n"); return NULL; }
#endif
    rg_model *m = calloc(1, sizeof(*m));
    m->hp = *p;
    const enum ggml_type wt = (enum ggml_type)p->wtype;
    const int64_t e = p->n_embd;
    size_t wbytes = rg_bytes(wt, e, p->n_vocab) * 2 + (size_t)e * p->n_ctx * 4 + (size_t)e * 8;
    wbytes += (size_t)p->n_layer * (rg_bytes(wt, e, 3 * e) + rg_bytes(wt, e, e) + rg_bytes(wt, e, 4 * e) + rg_bytes(wt, 4 * e, e) + (size_t)e * 4 * 13);
    wbytes += (size_t)(5 + 12 * p->n_layer) * 512 + (1u << 20);
    struct ggml_init_params ip = { wbytes, NULL, false };
    m->model_ctx = ggml_init(ip);
    struct ggml_context *c = m->model_ctx;
    m->wte = ggml_new_tensor_2d(c, wt, e, p->n_vocab);
    m->wpe = p->arch == 0 ? ggml_new_tensor_2d(c, GGML_TYPE_F32, e, p->n_ctx) : NULL;
    m->ln_f_g = ggml_new_tensor_1d(c, GGML_TYPE_F32, e);
    m->ln_f_b = ggml_new_tensor_1d(c, GGML_TYPE_F32, e);
    m->lm_head = (p->has_lm_head || p->arch == 1) ? ggml_new_tensor_2d(c, wt, e, p->n_vocab) : NULL;
    for (int i = 0; i < p->n_layer; i++) {
        rg_layer *L = &m->layers[i];
        L->ln_1_g = ggml_new_tensor_1d(c, GGML_TYPE_F32, e); L->ln_1_b = ggml_new_tensor_1d(c, GGML_TYPE_F32, e);
        L->ln_2_g = ggml_new_tensor_1d(c, GGML_TYPE_F32, e); L->ln_2_b = ggml_new_tensor_1d(c, GGML_TYPE_F32, e);
        L->c_attn_attn_w = ggml_new_tensor_2d(c, wt, e, 3 * e); L->c_attn_attn_b = ggml_new_tensor_1d(c, GGML_TYPE_F32, 3 * e);
        L->c_attn_proj_w = ggml_new_tensor_2d(c, wt, e, e);     L->c_attn_proj_b = ggml_new_tensor_1d(c, GGML_TYPE_F32, e);
        L->c_mlp_fc_w = ggml_new_tensor_2d(c, wt, e, 4 * e);    L->c_mlp_fc_b = ggml_new_tensor_1d(c, GGML_TYPE_F32, 4 * e);
        L->c_mlp_proj_w = ggml_new_tensor_2d(c, wt, 4 * e, e);  L->c_mlp_proj_b = ggml_new_tensor_1d(c, GGML_TYPE_F32, e);
    }
    return m;
}

void *rh_gpt2_tensor(rg_model *m, const char *name, size_t *nbytes) {       /* loader names, gpt2 lib.rs:59-107 */
    struct ggml_tensor *t = NULL;
    int il = -1; char sub[64];
    if (m->hp.arch == 1) {                                                 /* gptneox lib.rs:58-123 */
        if (!strcmp(name, "gpt_neox.embed_in.weight")) t = m->wte;
        else if (!strcmp(name, "gpt_neox.final_layer_norm.weight")) t = m->ln_f_g;
        else if (!strcmp(name, "gpt_neox.final_layer_norm.bias")) t = m->ln_f_b;
        else if (!strcmp(name, "embed_out.weight")) t = m->lm_head;
        else if (sscanf(name, "gpt_neox.layers.%d.%63s", &il, sub) == 2 && il >= 0 && il < m->hp.n_layer) {
            rg_layer *L = &m->layers[il];
            if      (!strcmp(sub, "input_layernorm.weight")) t = L->ln_1_g;
            else if (!strcmp(sub, "input_layernorm.bias")) t = L->ln_1_b;
            else if (!strcmp(sub, "post_attention_layernorm.weight")) t = L->ln_2_g;
            else if (!strcmp(sub, "post_attention_layernorm.bias")) t = L->ln_2_b;
            else if (!strcmp(sub, "attention.query_key_value.weight")) t = L->c_attn_attn_w;
            else if (!strcmp(sub, "attention.query_key_value.bias")) t = L->c_attn_attn_b;
            else if (!strcmp(sub, "attention.dense.weight")) t = L->c_attn_proj_w;
            else if (!strcmp(sub, "attention.dense.bias")) t = L->c_attn_proj_b;
            else if (!strcmp(sub, "mlp.dense_h_to_4h.weight")) t = L->c_mlp_fc_w;
            else if (!strcmp(sub, "mlp.dense_h_to_4h.bias")) t = L->c_mlp_fc_b;
            else if (!strcmp(sub, "mlp.dense_4h_to_h.weight")) t = L->c_mlp_proj_w;
            else if (!strcmp(sub, "mlp.dense_4h_to_h.bias")) t = L->c_mlp_proj_b;
        }
    } else if (!strcmp(name, "model/wte")) t = m->wte;
    else if (!strcmp(name, "model/wpe")) t = m->wpe;
    else if (!strcmp(name, "model/ln_f/g")) t = m->ln_f_g;
    else if (!strcmp(name, "model/ln_f/b")) t = m->ln_f_b;
    else if (!strcmp(name, "model/lm_head")) t = m->lm_head;
    else if (sscanf(name, "model/h%d/%63s", &il, sub) == 2 && il >= 0 && il < m->hp.n_layer) {
        rg_layer *L = &m->layers[il];
        if      (!strcmp(sub, "ln_1/g")) t = L->ln_1_g;
        else if (!strcmp(sub, "ln_1/b")) t = L->ln_1_b;
        else if (!strcmp(sub, "ln_2/g")) t = L->ln_2_g;
        else if (!strcmp(sub, "ln_2/b")) t = L->ln_2_b;
        else if (!strcmp(sub, "attn/c_attn/w")) t = L->c_attn_attn_w;
        else if (!strcmp(sub, "attn/c_attn/b")) t = L->c_attn_attn_b;
        else if (!strcmp(sub, "attn/c_proj/w")) t = L->c_attn_proj_w;
        else if (!strcmp(sub, "attn/c_proj/b")) t = L->c_attn_proj_b;
        else if (!strcmp(sub, "mlp/c_fc/w")) t = L->c_mlp_fc_w;
        else if (!strcmp(sub, "mlp/c_fc/b")) t = L->c_mlp_fc_b;
        else if (!strcmp(sub, "mlp/c_proj/w")) t = L->c_mlp_proj_w;
        else if (!strcmp(sub, "mlp/c_proj/b")) t = L->c_mlp_proj_b;
    }
    if (!t) return NULL;
    if (nbytes) *nbytes = ggml_nbytes(t);
    return t->data;
}

#ifdef GGML_USE_CUBLAS
static void rg_to_gpu(struct ggml_tensor *t) { if (t) { t->backend = GGML_BACKEND_GPU; ggml_cuda_transform_tensor(t->data, t); } }   /* tensor.rs:56-80 */
static void rg_free_gpu(struct ggml_tensor *t) { if (t) ggml_cuda_free_data(t); }
#endif

int rh_gpt2_finalize(rg_model *m) {
    const rg_params *p = &m->hp;
#ifdef GGML_USE_CUBLAS
    if (p->use_gpu) {
        /* Gpt2::new transfers wte and wpe too (gpt2 lib.rs:59-60), but the reference's CUDA backend has no get_rows: ggml.c:14589 then aborts on the
         * first node of the graph (GGML_ASSERT src0->backend == CPU), with their ggml-cuda.cu exactly as with our seam.  The two embedding tables
         * therefore stay on the host here, as Llama::new / GptNeoX::new keep theirs (llama lib.rs:52-57, gptneox lib.rs:58). */
        rg_to_gpu(m->ln_f_g); rg_to_gpu(m->ln_f_b); rg_to_gpu(m->lm_head);
        for (int i = 0; i < p->n_layer; i++) {
            struct ggml_tensor **ts = (struct ggml_tensor **)&m->layers[i];
            for (int k = 0; k < 12; k++) rg_to_gpu(ts[k]);
        }
        ggml_init_cublas();
        ggml_cuda_set_main_device(0);
        float split = 1.0f;
        ggml_cuda_set_tensor_split(&split);
        ggml_cuda_set_scratch_size((size_t)p->n_batch * 1024 * 1024);
    }
#endif
    const size_t n_elements = (size_t)p->n_embd * p->n_layer * p->n_ctx;     /* inference_session.rs:127-160 */
    struct ggml_init_params ip = { n_elements * 2 * 2 + 8192, NULL, false };
    m->session_ctx = ggml_init(ip);
    m->memory_k = ggml_new_tensor_1d(m->session_ctx, GGML_TYPE_F16, n_elements);
    m->memory_v = ggml_new_tensor_1d(m->session_ctx, GGML_TYPE_F16, n_elements);
    memset(m->memory_k->data, 0, ggml_nbytes(m->memory_k));
    memset(m->memory_v->data, 0, ggml_nbytes(m->memory_v));
#ifdef GGML_USE_CUBLAS
    if (p->use_gpu) { ggml_cuda_assign_buffers_no_scratch(m->memory_k); ggml_cuda_assign_buffers_no_scratch(m->memory_v); }
#endif
    const size_t B = (size_t)p->n_batch;
    size_t per = B * p->n_embd * 4 * (4 * 3 + 16) + (size_t)p->n_head * B * p->n_ctx * 4 * 2 + (size_t)p->n_ctx * p->n_embd * 2 * 2 + (64u << 20);
    m->scratch_size = per;
    m->scratch[0] = malloc(per); m->scratch[1] = malloc(per);
    m->eval_size = B * p->n_vocab * 4 + B * p->n_embd * 64 + (size_t)p->n_layer * 96 * 512 + B * p->n_embd * 4 * 40 + ggml_graph_overhead() + (64u << 20);
    m->eval_buf = malloc(m->eval_size);
    m->finalized = 1;
    return 0;
}

void rh_gpt2_reset(rg_model *m) { m->n_past = 0; }
int  rh_gpt2_n_past(rg_model *m) { return m->n_past; }

int rh_gpt2_eval(rg_model *m, const int32_t *tokens, int n, float *logits_out) {
    const rg_params *p = &m->hp;
    if (!m->finalized || p->arch != 0 || n < 1 || n > p->n_batch || m->n_past + n > p->n_ctx) return -1;
    const int n_embd = p->n_embd, n_head = p->n_head, ctx_size = p->n_ctx, session_len = m->n_past, input_len = n;
    const size_t ksz = 2, vsz = 2, f32sz = 4;

    if (m->ctx0) ggml_free(m->ctx0);
    struct ggml_init_params ip = { m->eval_size, m->eval_buf, false };
    m->ctx0 = ggml_init(ip);
    struct ggml_context *ctx0 = m->ctx0;
    m->can_offload = 0;

    struct ggml_tensor *embd = G(m, ggml_new_tensor_1d(ctx0, GGML_TYPE_I32, input_len));
    struct ggml_tensor *position = G(m, ggml_new_tensor_1d(ctx0, GGML_TYPE_I32, input_len));            /* :164-167 */
    for (int i = 0; i < input_len; i++) ((int32_t *)position->data)[i] = session_len + i;
    struct ggml_tensor *inpL = G(m, ggml_add(ctx0, G(m, ggml_get_rows(ctx0, m->wte, embd)), G(m, ggml_get_rows(ctx0, m->wpe, position))));   /* :169-172 */
    struct ggml_cgraph *gf = ggml_new_graph(ctx0);

    for (int il = 0; il < p->n_layer; il++) {
        const rg_layer *L = &m->layers[il];
        m->can_offload = p->use_gpu;                                                                     /* :176 */
        rg_use_scratch(m, 0);
        struct ggml_tensor *cur = G(m, ggml_norm(ctx0, inpL));                                           /* :179 */
        cur = G(m, ggml_add(ctx0, G(m, ggml_mul(ctx0, cur, L->ln_1_g)), L->ln_1_b));                     /* :180-183 */
        cur = G(m, ggml_mul_mat(ctx0, L->c_attn_attn_w, cur));                                           /* :186 */
        cur = G(m, ggml_add(ctx0, cur, L->c_attn_attn_b));                                               /* :187 */
        const size_t nb = cur->nb[1];
        struct ggml_tensor *qcur = G(m, ggml_view_2d(ctx0, cur, n_embd, input_len, nb, 0));              /* :190-195 */
        struct ggml_tensor *kcur = G(m, ggml_view_2d(ctx0, cur, n_embd, input_len, nb, f32sz * n_embd));
        struct ggml_tensor *vcur = G(m, ggml_view_2d(ctx0, cur, n_embd, input_len, nb, f32sz * n_embd * 2));
        struct ggml_tensor *k = G(m, ggml_view_1d(ctx0, m->memory_k, (int64_t)input_len * n_embd, (ksz * n_embd) * ((size_t)il * ctx_size + session_len)));   /* :198-202 */
        struct ggml_tensor *v = G(m, ggml_view_1d(ctx0, m->memory_v, (int64_t)input_len * n_embd, (vsz * n_embd) * ((size_t)il * ctx_size + session_len)));   /* :203-207 */
        ggml_build_forward_expand(gf, G(m, ggml_cpy(ctx0, kcur, k)));                                    /* :209 */
        ggml_build_forward_expand(gf, G(m, ggml_cpy(ctx0, vcur, v)));                                    /* :210 */
        struct ggml_tensor *q = G(m, ggml_permute(ctx0,
            G(m, ggml_cpy(ctx0, qcur, G(m, ggml_new_tensor_3d(ctx0, GGML_TYPE_F32, n_embd / n_head, n_head, input_len)))), 0, 2, 1, 3));      /* :213-219 */
        struct ggml_tensor *kk = G(m, ggml_permute(ctx0,
            G(m, ggml_reshape_3d(ctx0,
                G(m, ggml_view_1d(ctx0, m->memory_k, (int64_t)(session_len + input_len) * n_embd, (size_t)il * ctx_size * ksz * n_embd)),
                n_embd / n_head, n_head, session_len + input_len)), 0, 2, 1, 3));                        /* :221-233 */
        struct ggml_tensor *kq = G(m, ggml_mul_mat(ctx0, kk, q));                                        /* :235 */
        struct ggml_tensor *kq_scaled = G(m, ggml_scale_inplace(ctx0, kq, G(m, ggml_new_f32(ctx0, 1.0f / sqrtf((float)n_embd / (float)n_head)))));   /* :236-239 */
        struct ggml_tensor *kq_masked = G(m, ggml_diag_mask_inf_inplace(ctx0, kq_scaled, session_len));  /* :241 */
        struct ggml_tensor *kq_softmax = G(m, ggml_soft_max_inplace(ctx0, kq_masked));                   /* :242 */
        struct ggml_tensor *v_trans = G(m, ggml_cpy(ctx0,
            G(m, ggml_permute(ctx0,
                G(m, ggml_reshape_3d(ctx0,
                    G(m, ggml_view_1d(ctx0, m->memory_v, (int64_t)(session_len + input_len) * n_embd, (size_t)il * ctx_size * vsz * n_embd)),
                    n_embd / n_head, n_head, session_len + input_len)), 1, 2, 0, 3)),
            G(m, ggml_new_tensor_3d(ctx0, m->memory_v->type, session_len + input_len, n_embd / n_head, n_head))));                              /* :244-264 */
        struct ggml_tensor *kqv = G(m, ggml_mul_mat(ctx0, v_trans, kq_softmax));                         /* :266 */
        struct ggml_tensor *kqv_merged = G(m, ggml_permute(ctx0, kqv, 0, 2, 1, 3));                      /* :267 */
        cur = G(m, ggml_cpy(ctx0, kqv_merged, G(m, ggml_new_tensor_2d(ctx0, GGML_TYPE_F32, n_embd, input_len))));                              /* :269-272 */
        cur = G(m, ggml_mul_mat(ctx0, L->c_attn_proj_w, cur));                                           /* :275 */
        cur = G(m, ggml_add(ctx0, cur, L->c_attn_proj_b));                                               /* :276 */
        cur = G(m, ggml_add(ctx0, cur, inpL));                                                           /* :279 */
        struct ggml_tensor *ff_in = cur;                                                                 /* :282 */
        rg_use_scratch(m, 1);                                                                            /* :284 */
        cur = G(m, ggml_norm(ctx0, ff_in));                                                              /* :287 */
        cur = G(m, ggml_add(ctx0, G(m, ggml_mul(ctx0, cur, L->ln_2_g)), L->ln_2_b));                     /* :288-291 */
        cur = G(m, ggml_mul_mat(ctx0, L->c_mlp_fc_w, cur));                                              /* :294 */
        cur = G(m, ggml_add(ctx0, cur, L->c_mlp_fc_b));                                                  /* :295 */
        cur = G(m, ggml_gelu(ctx0, cur));                                                                /* :298 */
        cur = G(m, ggml_mul_mat(ctx0, L->c_mlp_proj_w, cur));                                            /* :301 */
        cur = G(m, ggml_add(ctx0, cur, L->c_mlp_proj_b));                                                /* :302 */
        inpL = G(m, ggml_add(ctx0, cur, ff_in));                                                         /* :305 */
    }
    rg_use_scratch(m, 0);                                                                                /* :308 */
    inpL = G(m, ggml_norm(ctx0, inpL));                                                                  /* :311 */
    inpL = G(m, ggml_add(ctx0, G(m, ggml_mul(ctx0, inpL, m->ln_f_g)), m->ln_f_b));                       /* :312 */
    rg_use_scratch(m, -1);                                                                               /* :314 */
    m->can_offload = 0;                                                                                  /* :315 */
    inpL = G(m, ggml_mul_mat(ctx0, m->lm_head ? m->lm_head : m->wte, inpL));                             /* :319-320 */

    memcpy(embd->data, tokens, (size_t)input_len * 4);
    ggml_build_forward_expand(gf, inpL);
    struct ggml_cplan plan = ggml_graph_plan(gf, p->n_threads);
    struct ggml_tensor *work = ggml_new_tensor_1d(ctx0, GGML_TYPE_I8, plan.work_size ? plan.work_size : 1);
    plan.work_data = work->data;
    ggml_graph_compute(gf, &plan);
    m->n_past += input_len;
    if (logits_out) memcpy(logits_out, inpL->data, (size_t)input_len * p->n_vocab * 4);
    return 0;
}

/* GptNeoX::evaluate (gptneox lib.rs:156-352) */
static struct ggml_tensor *rg_ffn(rg_model *m, struct ggml_context *ctx0, const rg_layer *L, struct ggml_tensor *input) {      /* :487-515 */
    struct ggml_tensor *cur = G(m, ggml_norm(ctx0, input));
    cur = G(m, ggml_add(ctx0, G(m, ggml_mul(ctx0, cur, L->ln_2_g)), L->ln_2_b));
    cur = G(m, ggml_mul_mat(ctx0, L->c_mlp_fc_w, cur));
    cur = G(m, ggml_add(ctx0, cur, L->c_mlp_fc_b));
    cur = G(m, ggml_gelu(ctx0, cur));
    cur = G(m, ggml_mul_mat(ctx0, L->c_mlp_proj_w, cur));
    return G(m, ggml_add(ctx0, cur, L->c_mlp_proj_b));
}

int rh_neox_eval(rg_model *m, const int32_t *tokens, int n, float *logits_out) {
    const rg_params *p = &m->hp;
    if (!m->finalized || p->arch != 1 || n < 1 || n > p->n_batch || m->n_past + n > p->n_ctx) return -1;
    const int n_embd = p->n_embd, n_head = p->n_head, n_ctx = p->n_ctx, n_past = m->n_past, n_rot = p->n_rot;
    const size_t ksz = 2, vsz = 2, f32sz = 4;

    if (m->ctx0) ggml_free(m->ctx0);
    struct ggml_init_params ip = { m->eval_size, m->eval_buf, false };
    m->ctx0 = ggml_init(ip);
    struct ggml_context *ctx0 = m->ctx0;
    m->can_offload = 0;
    struct ggml_tensor *embd = G(m, ggml_new_tensor_1d(ctx0, GGML_TYPE_I32, n));
    struct ggml_tensor *inpL = G(m, ggml_get_rows(ctx0, m->wte, embd));                                  /* :178 */
    struct ggml_cgraph *gf = ggml_new_graph(ctx0);
    for (int il = 0; il < p->n_layer; il++) {
        const rg_layer *L = &m->layers[il];
        m->can_offload = p->use_gpu;                                                                     /* :187 */
        rg_use_scratch(m, 0);
        struct ggml_tensor *cur = G(m, ggml_norm(ctx0, inpL));                                           /* :192 */
        cur = G(m, ggml_add(ctx0, G(m, ggml_mul(ctx0, cur, L->ln_1_g)), L->ln_1_b));                     /* :193-196 */
        cur = G(m, ggml_mul_mat(ctx0, L->c_attn_attn_w, cur));                                           /* :199 */
        cur = G(m, ggml_add(ctx0, cur, L->c_attn_attn_b));                                               /* :200 */
        const size_t nb = cur->nb[1];
        struct ggml_tensor *qcur = G(m, ggml_cont(ctx0, G(m, ggml_view_3d(ctx0, cur, n_embd / n_head, n_head, n, nb / n_head, nb, 0))));                        /* :205-210 */
        struct ggml_tensor *kcur = G(m, ggml_cont(ctx0, G(m, ggml_view_3d(ctx0, cur, n_embd / n_head, n_head, n, nb / n_head, nb, f32sz * n_embd / n_head))));  /* :211-216 */
        struct ggml_tensor *vcur = G(m, ggml_cont(ctx0, G(m, ggml_view_3d(ctx0, cur, n_embd / n_head, n_head, n, nb / n_head, nb, 2 * f32sz * n_embd / n_head))));
        qcur = G(m, ggml_rope_inplace(ctx0, qcur, n_past, n_rot, 2, 0));                                 /* :226-228 */
        kcur = G(m, ggml_rope_inplace(ctx0, kcur, n_past, n_rot, 2, 0));
        vcur = G(m, ggml_transpose(ctx0, G(m, ggml_reshape_2d(ctx0, vcur, n_embd, n))));                 /* :231 */
        struct ggml_tensor *k = G(m, ggml_view_1d(ctx0, m->memory_k, (int64_t)n * n_embd, (ksz * n_embd) * ((size_t)il * n_ctx + n_past)));      /* :233-237 */
        struct ggml_tensor *v = G(m, ggml_view_2d(ctx0, m->memory_v, n, n_embd, (size_t)n_ctx * vsz,
                                                  ((size_t)il * n_ctx) * vsz * n_embd + (size_t)n_past * vsz));                                  /* :239-244 */
        ggml_build_forward_expand(gf, G(m, ggml_cpy(ctx0, kcur, k)));                                    /* :246 */
        ggml_build_forward_expand(gf, G(m, ggml_cpy(ctx0, vcur, v)));                                    /* :247 */
        struct ggml_tensor *Q = G(m, ggml_permute(ctx0, qcur, 0, 2, 1, 3));                              /* :250 */
        struct ggml_tensor *K = G(m, ggml_permute(ctx0,
            G(m, ggml_reshape_3d(ctx0, G(m, ggml_view_1d(ctx0, m->memory_k, (int64_t)(n_past + n) * n_embd, (size_t)il * n_ctx * ksz * n_embd)),
                                 n_embd / n_head, n_head, n_past + n)), 0, 2, 1, 3));                    /* :252-264 */
        struct ggml_tensor *KQ = G(m, ggml_mul_mat(ctx0, K, Q));                                         /* :267 */
        struct ggml_tensor *KQ_scaled = G(m, ggml_scale_inplace(ctx0, KQ, G(m, ggml_new_f32(ctx0, 1.0f / sqrtf((float)n_embd / (float)n_head)))));   /* :270-273 */
        struct ggml_tensor *KQ_masked = G(m, ggml_diag_mask_inf_inplace(ctx0, KQ_scaled, n_past));       /* :276 */
        struct ggml_tensor *KQ_softmax = G(m, ggml_soft_max_inplace(ctx0, KQ_masked));                   /* :279 */
        struct ggml_tensor *V = G(m, ggml_view_3d(ctx0, m->memory_v, n_past + n, n_embd / n_head, n_head,
            (size_t)n_ctx * vsz, (size_t)n_ctx * vsz * n_embd / n_head, (size_t)il * n_ctx * vsz * n_embd));                                     /* :282-290 */
        struct ggml_tensor *KQV = G(m, ggml_mul_mat(ctx0, V, KQ_softmax));                               /* :293 */
        struct ggml_tensor *KQV_merged = G(m, ggml_permute(ctx0, KQV, 0, 2, 1, 3));                      /* :295 */
        cur = G(m, ggml_cpy(ctx0, KQV_merged, G(m, ggml_new_tensor_2d(ctx0, GGML_TYPE_F32, n_embd, n))));                                       /* :298 */
        cur = G(m, ggml_mul_mat(ctx0, L->c_attn_proj_w, cur));                                           /* :301 */
        cur = G(m, ggml_add(ctx0, cur, L->c_attn_proj_b));                                               /* :302 */
        rg_use_scratch(m, 1);                                                                            /* :305 */
        if (!p->use_parallel_residual) {                                                                 /* :308-312 */
            struct ggml_tensor *ff_in = G(m, ggml_add(ctx0, cur, inpL));
            cur = rg_ffn(m, ctx0, L, ff_in);
            inpL = G(m, ggml_add(ctx0, cur, ff_in));
        } else {                                                                                         /* :313-325 */
            struct ggml_tensor *ff_in = cur;
            cur = rg_ffn(m, ctx0, L, inpL);
            cur = G(m, ggml_add(ctx0, cur, ff_in));
            inpL = G(m, ggml_add(ctx0, cur, inpL));
        }
    }
    rg_use_scratch(m, 0);                                                                                /* :329 */
    inpL = G(m, ggml_norm(ctx0, inpL));                                                                  /* :332 */
    inpL = G(m, ggml_add(ctx0, G(m, ggml_mul(ctx0, inpL, m->ln_f_g)), m->ln_f_b));                       /* :334 */
    rg_use_scratch(m, -1);                                                                               /* :339 */
    m->can_offload = 0;                                                                                  /* :340 */
    inpL = G(m, ggml_mul_mat(ctx0, m->lm_head, inpL));                                                   /* :342 */

    memcpy(embd->data, tokens, (size_t)n * 4);
    ggml_build_forward_expand(gf, inpL);
    struct ggml_cplan plan = ggml_graph_plan(gf, p->n_threads);
    struct ggml_tensor *work = ggml_new_tensor_1d(ctx0, GGML_TYPE_I8, plan.work_size ? plan.work_size : 1);
    plan.work_data = work->data;
    ggml_graph_compute(gf, &plan);
    m->n_past += n;
    if (logits_out) memcpy(logits_out, inpL->data, (size_t)n * p->n_vocab * 4);
    return 0;
}

void rh_gpt2_free(rg_model *m) {
    if (!m) return;
#ifdef GGML_USE_CUBLAS
    if (m->hp.use_gpu) {
        rg_free_gpu(m->ln_f_g); rg_free_gpu(m->ln_f_b); rg_free_gpu(m->lm_head);
        for (int i = 0; i < m->hp.n_layer; i++) { struct ggml_tensor **ts = (struct ggml_tensor **)&m->layers[i]; for (int k = 0; k < 12; k++) rg_free_gpu(ts[k]); }
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
