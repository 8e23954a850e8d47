/*
 * oracle/ggml_oracle.h -- TEST INFRASTRUCTURE (parity oracle). Not shipped, not on the product path.
 *
 * Plain-C restatement of the arithmetic on the reference's quantized mul_mat hot path
 * (rustformers/llm @ 9376078c, vendored llama.cpp @ 1a941869; "LC/" = crates/ggml/sys/llama-cpp/).
 * Every function cites the reference lines it follows.  PARITY PINNED: tests/test_oracle_pin.py checks
 * this restatement bit-for-bit against the reference's own compiled ggml.c (oracle/_ref/libggml_ref.so)
 * and against golden vectors generated from it (tests/golden/, oracle/gen_golden.py).
 *
 * The restated arithmetic is the one rustformers' build actually runs on x86-64: the AVX2 code paths
 * (crates/ggml/sys/build.rs:46-63 passes -mavx -mavx2 -mfma -mf16c -msse3, never AVX-512/VNNI), i.e.
 * 8 f32 lanes per accumulator, fused multiply-add, round-half-even activation quantization.
 */
#ifndef GGML_ORACLE_H
#define GGML_ORACLE_H
#include <stddef.h>
#include <stdint.h>

enum { OR_F32 = 0, OR_F16 = 1, OR_Q4_0 = 2, OR_Q4_1 = 3, OR_Q5_0 = 6, OR_Q5_1 = 7, OR_Q8_0 = 8, OR_Q8_1 = 9 };

#pragma pack(push, 1)
typedef struct { uint16_t d; uint8_t qs[16]; } or_block_q4_0;                          /* LC/ggml.c:895-900 */
typedef struct { uint16_t d; uint16_t m; uint8_t qs[16]; } or_block_q4_1;              /* LC/ggml.c:902-908 */
typedef struct { uint16_t d; uint8_t qh[4]; uint8_t qs[16]; } or_block_q5_0;           /* LC/ggml.c:910-916 */
typedef struct { uint16_t d; uint16_t m; uint8_t qh[4]; uint8_t qs[16]; } or_block_q5_1; /* LC/ggml.c:918-925 */
typedef struct { uint16_t d; int8_t qs[32]; } or_block_q8_0;                           /* LC/ggml.c:927-932 */
typedef struct { float d; float s; int8_t qs[32]; } or_block_q8_1;                     /* LC/ggml.c:934-940 */
#pragma pack(pop)

uint16_t or_fp32_to_fp16(float f);
float    or_fp16_to_fp32(uint16_t h);

size_t or_row_bytes(int type, int64_t k);
int    or_vec_dot_type(int type);

/* weight quantizers = quantize_row_*_reference, LC/ggml.c:943-1111 (what ggml_quantize_* calls, :18083-18230) */
void or_quantize_weights(int type, const float *src, void *dst, int64_t nrows, int64_t k);
/* to_float, LC/ggml.c:1525-1635 */
void or_dequantize_row(int type, const void *x, float *y, int64_t k);
/* activation quantizers = the AVX2 bodies of quantize_row_q8_0 / q8_1, LC/ggml.c:1217-1300, 1427-1518 */
void or_quantize_row_q8_0(const float *x, void *y, int64_t k);
void or_quantize_row_q8_1(const float *x, void *y, int64_t k);
void or_quantize_row_act(int vec_dot_type, const float *x, void *y, int64_t k);
/* ggml_vec_dot_q*_q8_* AVX2 bodies, LC/ggml.c:2434-2457, 2702-2735, 2916-2938, 3166-3191, 3315-3336 */
float or_vec_dot(int type, int64_t n, const void *x, const void *y);
/* ggml_vec_dot_f16 AVX body, LC/ggml.c:2325-2359 with the F16 SIMD macros :1937-1975 */
float or_vec_dot_f16(int64_t n, const uint16_t *x, const uint16_t *y);

/* ggml_compute_forward_mul_mat, LC/ggml.c:10397-10586: W[type; K, N] (row stride = or_row_bytes) x X[f32; K, B] -> dst[B][N] */
void or_mul_mat(int type, const void *w, const float *x, float *dst, int64_t K, int64_t N, int64_t B);

void or_rms_norm(const float *x, float *y, int64_t n, int64_t rows, float eps);      /* LC/ggml.c:10129-10175 */
void or_norm(const float *x, float *y, int64_t n, int64_t rows);                     /* LC/ggml.c:10063-10111 */
void or_soft_max(const float *x, float *y, int64_t n, int64_t rows);                 /* LC/ggml.c:11352-11421 */
void or_scale_mask_soft_max(float *x, int64_t nc, int64_t nr, int64_t nz, float scale, int n_past); /* :10733, :11268-11316, :11352 */
void or_silu(const float *x, float *y, int64_t n);                                   /* LC/ggml.c:3556-3564 (GGML_SILU_FP16) */
void or_gelu(const float *x, float *y, int64_t n);                                   /* LC/ggml.c:3499-3507 (GGML_GELU_FP16) */
/* ggml_compute_forward_rope_f32, LC/ggml.c:11774-11901: x is [ne0, ne1, ne2] contiguous, in place; modes 0 and 2 */
void or_rope(float *x, int64_t ne0, int64_t ne1, int64_t ne2, int n_past, int n_dims, int mode, float freq_base, float freq_scale);

const uint16_t *or_table_silu(void);
const uint16_t *or_table_gelu(void);
const uint16_t *or_table_exp(void);

/* ---- whole-model restatement (llama_oracle.c) ---- */
typedef struct {
    int32_t n_vocab, n_embd, n_head, n_head_kv, n_layer, n_ff, n_rot, n_ctx, wtype;
} or_hparams;
typedef struct or_llama or_llama;
or_llama *or_llama_new(const or_hparams *hp);
/* returns the host buffer the named tensor lives in (loader names, crates/models/llama/src/lib.rs:52-91) */
void *or_llama_tensor(or_llama *m, const char *name, size_t *nbytes);
void  or_llama_reset(or_llama *m);
void  or_llama_set_n_past(or_llama *m, int n);
void  or_llama_set_rope(or_llama *m, float freq_base, float freq_scale);
int   or_llama_eval(or_llama *m, const int32_t *tokens, int n, float *logits_all);
void *or_llama_kv(or_llama *m, int which, size_t *nbytes);
void  or_llama_free(or_llama *m);
/* optional per-layer taps for debugging parity: copies the residual stream after layer il (or -1: final norm) */
void  or_llama_set_tap(or_llama *m, float *buf, int il);
void  or_llama_set_tap_stage(or_llama *m, float *buf, int il, int stage);

#endif
