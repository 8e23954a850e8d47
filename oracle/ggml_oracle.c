/*
 * oracle/ggml_oracle.c -- TEST INFRASTRUCTURE (parity oracle). See ggml_oracle.h for scope and pinning.
 *
 * Compiled with -ffp-contract=off: every fused multiply-add the reference build performs is written
 * explicitly as fmaf() (the reference is built with gcc's default -ffp-contract=fast plus -mfma, so its
 * scalar `a*b + c` expressions fuse), everything else rounds after each operation.
 */
#include "ggml_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- fp16 <-> fp32, round-to-nearest-even (LC/ggml.c:309-317: _cvtss_sh(x, 0) / _cvtsh_ss) ------------- */
uint16_t or_fp32_to_fp16(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? (0x0200u | ((ax >> 13) & 0x3ffu)) : 0)); /* inf / quiet NaN */
    if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);            /* rounds to >= 65520 -> inf */
    if (ax < 0x33000001u) return (uint16_t)sign;                         /* <= 2^-25 -> +-0 (ties to even) */
    int e = (int)(ax >> 23) - 127;
    uint32_t m = (ax & 0x7fffffu) | 0x800000u;
    int shift;
    uint32_t base;
    if (e < -14) { shift = 13 + (-14 - e); base = 0; }                   /* subnormal half */
    else         { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7fffffu; }
    uint32_t q = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u);
    const uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    return (uint16_t)(sign | (base + q));                                /* carry into the exponent is correct by construction */
}

float or_fp16_to_fp32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    const uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) x = sign;
        else { int s = 0; uint32_t mm = m; while (!(mm & 0x400u)) { mm <<= 1; s++; } x = sign | ((uint32_t)(113 - s) << 23) | ((mm & 0x3ffu) << 13); }
    } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &x, 4); return f;
}

size_t or_row_bytes(int type, int64_t k) {
    switch (type) {
        case OR_F32: return (size_t)k * 4;
        case OR_F16: return (size_t)k * 2;
        case OR_Q4_0: return (size_t)(k / 32) * sizeof(or_block_q4_0);
        case OR_Q4_1: return (size_t)(k / 32) * sizeof(or_block_q4_1);
        case OR_Q5_0: return (size_t)(k / 32) * sizeof(or_block_q5_0);
        case OR_Q5_1: return (size_t)(k / 32) * sizeof(or_block_q5_1);
        case OR_Q8_0: return (size_t)(k / 32) * sizeof(or_block_q8_0);
        case OR_Q8_1: return (size_t)(k / 32) * sizeof(or_block_q8_1);
    }
    return 0;
}

int or_vec_dot_type(int type) { /* type_traits[].vec_dot_type, LC/ggml.c:1645-1737 */
    switch (type) {
        case OR_Q4_0: case OR_Q5_0: case OR_Q8_0: return OR_Q8_0;
        case OR_Q4_1: case OR_Q5_1: return OR_Q8_1;
        case OR_F16: return OR_F16;
    }
    return OR_F32;
}

#define OR_MIN(a, b) ((a) < (b) ? (a) : (b))

/* ---- weight quantizers: quantize_row_q*_reference ------------------------------------------------------- */
static void q_w_q4_0(const float *x, or_block_q4_0 *y, int64_t k) { /* LC/ggml.c:943-977 */
    for (int64_t i = 0; i < k / 32; i++) {
        float amax = 0.0f, max = 0.0f;
        for (int j = 0; j < 32; j++) { const float v = x[i*32 + j]; if (amax < fabsf(v)) { amax = fabsf(v); max = v; } }
        const float d = max / -8;
        const float id = d ? 1.0f/d : 0.0f;
        y[i].d = or_fp32_to_fp16(d);
        for (int j = 0; j < 16; ++j) {
            const float x0 = x[i*32 + j]*id, x1 = x[i*32 + 16 + j]*id;
            const uint8_t xi0 = OR_MIN(15, (int8_t)(x0 + 8.5f));
            const uint8_t xi1 = OR_MIN(15, (int8_t)(x1 + 8.5f));
            y[i].qs[j] = xi0 | (xi1 << 4);
        }
    }
}
static void q_w_q4_1(const float *x, or_block_q4_1 *y, int64_t k) { /* LC/ggml.c:983-1017 */
    for (int64_t i = 0; i < k / 32; i++) {
        float min = FLT_MAX, max = -FLT_MAX;
        for (int j = 0; j < 32; j++) { const float v = x[i*32 + j]; if (v < min) min = v; if (v > max) max = v; }
        const float d = (max - min) / 15;
        const float id = d ? 1.0f/d : 0.0f;
        y[i].d = or_fp32_to_fp16(d); y[i].m = or_fp32_to_fp16(min);
        for (int j = 0; j < 16; ++j) {
            const float x0 = (x[i*32 + j] - min)*id, x1 = (x[i*32 + 16 + j] - min)*id;
            const uint8_t xi0 = OR_MIN(15, (int8_t)(x0 + 0.5f));
            const uint8_t xi1 = OR_MIN(15, (int8_t)(x1 + 0.5f));
            y[i].qs[j] = xi0 | (xi1 << 4);
        }
    }
}
static void q_w_q5_0(const float *x, or_block_q5_0 *y, int64_t k) { /* LC/ggml.c:1023-1064 */
    for (int64_t i = 0; i < k / 32; i++) {
        float amax = 0.0f, max = 0.0f;
        for (int j = 0; j < 32; j++) { const float v = x[i*32 + j]; if (amax < fabsf(v)) { amax = fabsf(v); max = v; } }
        const float d = max / -16;
        const float id = d ? 1.0f/d : 0.0f;
        y[i].d = or_fp32_to_fp16(d);
        uint32_t qh = 0;
        for (int j = 0; j < 16; ++j) {
            const float x0 = x[i*32 + j]*id, x1 = x[i*32 + 16 + j]*id;
            const uint8_t xi0 = OR_MIN(31, (int8_t)(x0 + 16.5f));
            const uint8_t xi1 = OR_MIN(31, (int8_t)(x1 + 16.5f));
            y[i].qs[j] = (xi0 & 0x0F) | ((xi1 & 0x0F) << 4);
            qh |= ((xi0 & 0x10u) >> 4) << (j + 0);
            qh |= ((xi1 & 0x10u) >> 4) << (j + 16);
        }
        memcpy(y[i].qh, &qh, 4);
    }
}
static void q_w_q5_1(const float *x, or_block_q5_1 *y, int64_t k) { /* LC/ggml.c:1070-1111 */
    for (int64_t i = 0; i < k / 32; i++) {
        float min = FLT_MAX, max = -FLT_MAX;
        for (int j = 0; j < 32; j++) { const float v = x[i*32 + j]; if (v < min) min = v; if (v > max) max = v; }
        const float d = (max - min) / 31;
        const float id = d ? 1.0f/d : 0.0f;
        y[i].d = or_fp32_to_fp16(d); y[i].m = or_fp32_to_fp16(min);
        uint32_t qh = 0;
        for (int j = 0; j < 16; ++j) {
            const float x0 = (x[i*32 + j] - min)*id, x1 = (x[i*32 + 16 + j] - min)*id;
            const uint8_t xi0 = (uint8_t)(x0 + 0.5f);
            const uint8_t xi1 = (uint8_t)(x1 + 0.5f);
            y[i].qs[j] = (xi0 & 0x0F) | ((xi1 & 0x0F) << 4);
            qh |= ((xi0 & 0x10u) >> 4) << (j + 0);
            qh |= ((xi1 & 0x10u) >> 4) << (j + 16);
        }
        memcpy(y[i].qh, &qh, 4);
    }
}
static void q_w_q8_0(const float *x, or_block_q8_0 *y, int64_t k) { /* LC/ggml.c:1121-1145 (roundf, id = 1/d) */
    for (int64_t i = 0; i < k / 32; i++) {
        float amax = 0.0f;
        for (int j = 0; j < 32; j++) { const float v = fabsf(x[i*32 + j]); amax = amax > v ? amax : v; }
        const float d = amax / 127;
        const float id = d ? 1.0f/d : 0.0f;
        y[i].d = or_fp32_to_fp16(d);
        for (int j = 0; j < 32; ++j) y[i].qs[j] = (int8_t)roundf(x[i*32 + j]*id);
    }
}

void or_quantize_weights(int type, const float *src, void *dst, int64_t nrows, int64_t k) {
    const size_t rb = or_row_bytes(type, k);
    #pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nrows; r++) {
        const float *x = src + r * k; char *y = (char *)dst + r * rb;
        switch (type) {
            case OR_Q4_0: q_w_q4_0(x, (or_block_q4_0 *)y, k); break;
            case OR_Q4_1: q_w_q4_1(x, (or_block_q4_1 *)y, k); break;
            case OR_Q5_0: q_w_q5_0(x, (or_block_q5_0 *)y, k); break;
            case OR_Q5_1: q_w_q5_1(x, (or_block_q5_1 *)y, k); break;
            case OR_Q8_0: q_w_q8_0(x, (or_block_q8_0 *)y, k); break;
        }
    }
}

/* ---- to_float: dequantize_row_*, LC/ggml.c:1525-1635 (q4_1/q5_1 `x*d + m` fuses in the reference build) - */
static inline int q5_hi(uint32_t qh, int j) { return (int)((qh >> j) & 1u) << 4; }

void or_dequantize_row(int type, const void *vx, float *y, int64_t k) {
    const int64_t nb = k / 32;
    for (int64_t i = 0; i < nb; i++) {
        switch (type) {
            case OR_Q4_0: { const or_block_q4_0 *x = vx; const float d = or_fp16_to_fp32(x[i].d);
                for (int j = 0; j < 16; j++) { y[i*32+j] = ((x[i].qs[j] & 0xF) - 8)*d; y[i*32+j+16] = ((x[i].qs[j] >> 4) - 8)*d; } } break;
            case OR_Q4_1: { const or_block_q4_1 *x = vx; const float d = or_fp16_to_fp32(x[i].d), m = or_fp16_to_fp32(x[i].m);
                for (int j = 0; j < 16; j++) { y[i*32+j] = fmaf((float)(x[i].qs[j] & 0xF), d, m); y[i*32+j+16] = fmaf((float)(x[i].qs[j] >> 4), d, m); } } break;
            case OR_Q5_0: { const or_block_q5_0 *x = vx; const float d = or_fp16_to_fp32(x[i].d); uint32_t qh; memcpy(&qh, x[i].qh, 4);
                for (int j = 0; j < 16; j++) { y[i*32+j] = (((x[i].qs[j] & 0xF) | q5_hi(qh, j)) - 16)*d; y[i*32+j+16] = (((x[i].qs[j] >> 4) | q5_hi(qh, j+16)) - 16)*d; } } break;
            case OR_Q5_1: { const or_block_q5_1 *x = vx; const float d = or_fp16_to_fp32(x[i].d), m = or_fp16_to_fp32(x[i].m); uint32_t qh; memcpy(&qh, x[i].qh, 4);
                for (int j = 0; j < 16; j++) { y[i*32+j] = fmaf((float)((x[i].qs[j] & 0xF) | q5_hi(qh, j)), d, m); y[i*32+j+16] = fmaf((float)((x[i].qs[j] >> 4) | q5_hi(qh, j+16)), d, m); } } break;
            case OR_Q8_0: { const or_block_q8_0 *x = vx; const float d = or_fp16_to_fp32(x[i].d);
                for (int j = 0; j < 32; j++) y[i*32+j] = x[i].qs[j]*d; } break;
        }
    }
}

/* ---- activation quantizers: AVX2 bodies ---------------------------------------------------------------- */
/* LC/ggml.c:1217-1300. d = amax/127 (stored fp16); multiplier 127/amax (NOT 1/d); _mm256_round_ps(NEAREST) =
 * round-half-to-even, reproduced with rintf() under the default rounding mode. */
void or_quantize_row_q8_0(const float *x, void *vy, int64_t k) {
    or_block_q8_0 *y = vy;
    for (int64_t i = 0; i < k / 32; i++) {
        float amax = 0.0f;
        for (int j = 0; j < 32; j++) { const float v = fabsf(x[i*32 + j]); if (v > amax) amax = v; }
        const float d = amax / 127.f;
        y[i].d = or_fp32_to_fp16(d);
        const float id = (amax != 0.0f) ? 127.f / amax : 0.0f;
        for (int j = 0; j < 32; j++) y[i].qs[j] = (int8_t)rintf(x[i*32 + j] * id);
    }
}
/* LC/ggml.c:1427-1518. d kept f32; s = d * (float)sum(q) (:1474). */
void or_quantize_row_q8_1(const float *x, void *vy, int64_t k) {
    or_block_q8_1 *y = vy;
    for (int64_t i = 0; i < k / 32; i++) {
        float amax = 0.0f;
        for (int j = 0; j < 32; j++) { const float v = fabsf(x[i*32 + j]); if (v > amax) amax = v; }
        const float d = amax / 127.f;
        y[i].d = d;
        const float id = (amax != 0.0f) ? 127.f / amax : 0.0f;
        int sum = 0;
        for (int j = 0; j < 32; j++) { const int q = (int)rintf(x[i*32 + j] * id); y[i].qs[j] = (int8_t)q; sum += q; }
        y[i].s = d * (float)sum;
    }
}
void or_quantize_row_act(int t, const float *x, void *y, int64_t k) {
    if (t == OR_Q8_0) or_quantize_row_q8_0(x, y, k); else or_quantize_row_q8_1(x, y, k);
}

/* ---- vec_dot: AVX2 bodies. One __m256 accumulator = 8 f32 lanes; lane L holds the four int8 products of
 * elements 4L..4L+3 (mul_sum_i8_pairs_float, LC/ggml.c:685-700: maddubs -> madd -> cvtepi32_ps, all exact),
 * accumulated with _mm256_fmadd_ps, reduced by hsum_float_8 (LC/ggml.c:608-616). ----------------------- */
static inline float hsum8(const float a[8]) {
#ifdef OR_PERTURB_HSUM
    /* liboracle_perturbed.so only (tests/test_chaos.py): the SAME eight f32 values, added in a different association.  This is
     * the smallest possible deviation from the reference (~1e-7 per mat-mul output) and is used to show what it does to logits. */
    return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
#endif
    const float r0 = a[4] + a[0], r1 = a[5] + a[1], r2 = a[6] + a[2], r3 = a[7] + a[3];
    const float s0 = r0 + r2, s1 = r1 + r3;
    return s0 + s1;
}
static inline void unpack_w(int type, const void *blk, int8_t w[32], float *d, float *m) {
    *m = 0.0f;
    switch (type) {
        case OR_Q4_0: { const or_block_q4_0 *b = blk; *d = or_fp16_to_fp32(b->d);            /* bytes_from_nibbles_32 - 8, :2443-2447 */
            for (int j = 0; j < 16; j++) { w[j] = (int8_t)((b->qs[j] & 0xF) - 8); w[j+16] = (int8_t)((b->qs[j] >> 4) - 8); } } break;
        case OR_Q4_1: { const or_block_q4_1 *b = blk; *d = or_fp16_to_fp32(b->d); *m = or_fp16_to_fp32(b->m);
            for (int j = 0; j < 16; j++) { w[j] = (int8_t)(b->qs[j] & 0xF); w[j+16] = (int8_t)(b->qs[j] >> 4); } } break;
        case OR_Q5_0: { const or_block_q5_0 *b = blk; *d = or_fp16_to_fp32(b->d); uint32_t qh; memcpy(&qh, b->qh, 4); /* :2924-2927: (nibble | ~bit<<4..) == q5 - 16 */
            for (int j = 0; j < 16; j++) { w[j] = (int8_t)(((b->qs[j] & 0xF) | q5_hi(qh, j)) - 16); w[j+16] = (int8_t)(((b->qs[j] >> 4) | q5_hi(qh, j+16)) - 16); } } break;
        case OR_Q5_1: { const or_block_q5_1 *b = blk; *d = or_fp16_to_fp32(b->d); *m = or_fp16_to_fp32(b->m); uint32_t qh; memcpy(&qh, b->qh, 4);
            for (int j = 0; j < 16; j++) { w[j] = (int8_t)((b->qs[j] & 0xF) | q5_hi(qh, j)); w[j+16] = (int8_t)((b->qs[j] >> 4) | q5_hi(qh, j+16)); } } break;
        case OR_Q8_0: { const or_block_q8_0 *b = blk; *d = or_fp16_to_fp32(b->d); memcpy(w, b->qs, 32); } break;
    }
}

float or_vec_dot(int type, int64_t n, const void *vx, const void *vy) {
    const int64_t nb = n / 32;
    const size_t wb = or_row_bytes(type, 32);
    const int q81 = (or_vec_dot_type(type) == OR_Q8_1);
    float acc[8] = {0}, summs = 0.0f;
    int8_t w[32];
    for (int64_t i = 0; i < nb; i++) {
        float dw, mw, dx, sx = 0.0f; const int8_t *q;
        unpack_w(type, (const char *)vx + i * wb, w, &dw, &mw);
        if (q81) { const or_block_q8_1 *y = (const or_block_q8_1 *)vy + i; dx = y->d; sx = y->s; q = y->qs; }
        else     { const or_block_q8_0 *y = (const or_block_q8_0 *)vy + i; dx = or_fp16_to_fp32(y->d); q = y->qs; }
        if (q81) summs = fmaf(mw, sx, summs);                 /* `summs += m*s` (:2713, :3178) fuses under -ffp-contract=fast */
        const float d = dw * dx;                              /* _mm256_mul_ps / scalar f32 product, one rounding */
        for (int L = 0; L < 8; L++) {
            int s = 0;
            for (int t = 0; t < 4; t++) s += (int)w[4*L + t] * (int)q[4*L + t];
            acc[L] = fmaf(d, (float)s, acc[L]);               /* _mm256_fmadd_ps */
        }
    }
    return q81 ? hsum8(acc) + summs : hsum8(acc);
}

/* ggml_vec_dot_f16, LC/ggml.c:2325-2359: GGML_F16_STEP 32, 4 accumulators x 8 lanes, f16->f32 loads (exact),
 * _mm256_fmadd_ps; GGML_F32x8_REDUCE (:1899-1916); leftovers accumulate in double on top of the reduced f32. */
float or_vec_dot_f16(int64_t n, const uint16_t *x, const uint16_t *y) {
    const int64_t np = n & ~31LL;
    float sum[4][8]; memset(sum, 0, sizeof(sum));
    for (int64_t i = 0; i < np; i += 32)
        for (int j = 0; j < 4; j++)
            for (int l = 0; l < 8; l++)
                sum[j][l] = fmaf(or_fp16_to_fp32(x[i + j*8 + l]), or_fp16_to_fp32(y[i + j*8 + l]), sum[j][l]);
    float t0[4];
    for (int l = 0; l < 8; l++) { sum[0][l] += sum[2][l]; sum[1][l] += sum[3][l]; }
    for (int l = 0; l < 8; l++) sum[0][l] += sum[1][l];
    for (int l = 0; l < 4; l++) t0[l] = sum[0][l] + sum[0][l + 4];
    const float h0 = t0[0] + t0[1], h1 = t0[2] + t0[3];
    double sumf = (double)(h0 + h1);
    for (int64_t i = np; i < n; ++i) sumf += (double)(or_fp16_to_fp32(x[i]) * or_fp16_to_fp32(y[i]));
    return (float)sumf;
}

/* ---- mul_mat, LC/ggml.c:10397-10586: INIT quantizes every src1 row (:10504-10520), COMPUTE is one vec_dot per
 * dst element (:10570-10572), so the result is independent of the thread split. ---------------------------- */
void or_mul_mat(int type, const void *w, const float *x, float *dst, int64_t K, int64_t N, int64_t B) {
    const int vt = or_vec_dot_type(type);
    const size_t wrb = or_row_bytes(type, K), xrb = or_row_bytes(vt, K);
    char *xq = malloc(xrb * (size_t)B);
    for (int64_t b = 0; b < B; b++) or_quantize_row_act(vt, x + b * K, xq + b * xrb, K);
    #pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; r++)
        for (int64_t b = 0; b < B; b++)
            dst[b * N + r] = or_vec_dot(type, K, (const char *)w + r * wrb, xq + b * xrb);
    free(xq);
}

/* ---- fp16 look-up tables, LC/ggml.c:4313-4326 (built with this host's libm, exactly as ggml_init does) ---- */
static uint16_t T_SILU[1 << 16], T_GELU[1 << 16], T_EXP[1 << 16];
static int tables_ready = 0;
static inline float gelu_f32(float x) { /* LC/ggml.c:3484-3490; `1.0f + A*x*x` fuses to fma(A*x, x, 1) in the reference build */
    return 0.5f*x*(1.0f + tanhf(0.79788456080286535587989211986876f*x*fmaf(0.044715f*x, x, 1.0f)));
}
static void ensure_tables(void) {
    if (tables_ready) return;
    #pragma omp critical
    if (!tables_ready) {
        for (int i = 0; i < (1 << 16); ++i) {
            const float f = or_fp16_to_fp32((uint16_t)i);
            T_GELU[i] = or_fp32_to_fp16(gelu_f32(f));
            T_SILU[i] = or_fp32_to_fp16(f/(1.0f + expf(-f)));             /* ggml_silu_f32 :3545-3547 */
            T_EXP[i]  = or_fp32_to_fp16(expf(f));
        }
        tables_ready = 1;
    }
}
const uint16_t *or_table_silu(void) { ensure_tables(); return T_SILU; }
const uint16_t *or_table_gelu(void) { ensure_tables(); return T_GELU; }
const uint16_t *or_table_exp(void)  { ensure_tables(); return T_EXP; }

void or_silu(const float *x, float *y, int64_t n) { ensure_tables(); for (int64_t i = 0; i < n; i++) y[i] = or_fp16_to_fp32(T_SILU[or_fp32_to_fp16(x[i])]); }
void or_gelu(const float *x, float *y, int64_t n) { ensure_tables(); for (int64_t i = 0; i < n; i++) y[i] = or_fp16_to_fp32(T_GELU[or_fp32_to_fp16(x[i])]); }

/* ---- norms: double accumulators (ggml_float, LC/ggml.c:270) ------------------------------------------------- */
void or_rms_norm(const float *x, float *y, int64_t n, int64_t rows, float eps) { /* LC/ggml.c:10129-10175 */
    for (int64_t r = 0; r < rows; r++) {
        const float *xr = x + r * n; float *yr = y + r * n;
        double sum = 0.0;
        for (int64_t i = 0; i < n; i++) sum += (double)(xr[i] * xr[i]);
        const float mean = (float)(sum / n);
        const float scale = 1.0f / sqrtf(mean + eps);
        for (int64_t i = 0; i < n; i++) yr[i] = xr[i] * scale;
    }
}
void or_norm(const float *x, float *y, int64_t n, int64_t rows) { /* LC/ggml.c:10063-10111, eps 1e-5 */
    for (int64_t r = 0; r < rows; r++) {
        const float *xr = x + r * n; float *yr = y + r * n;
        double sum = 0.0;
        for (int64_t i = 0; i < n; i++) sum += (double)xr[i];
        const float mean = (float)(sum / n);
        double sum2 = 0.0;
        for (int64_t i = 0; i < n; i++) { const float v = xr[i] - mean; yr[i] = v; sum2 += (double)(v * v); }
        const float variance = (float)(sum2 / n);
        const float scale = 1.0f / sqrtf(variance + 1e-5f);
        for (int64_t i = 0; i < n; i++) yr[i] *= scale;
    }
}

/* LC/ggml.c:11352-11421: max-subtract, exp through the fp16 table, -inf -> 0, double sum, scale by (float)(1/sum) */
void or_soft_max(const float *x, float *y, int64_t n, int64_t rows) {
    ensure_tables();
    for (int64_t r = 0; r < rows; r++) {
        const float *sp = x + r * n; float *dp = y + r * n;
        float max = -INFINITY;
        for (int64_t i = 0; i < n; i++) if (sp[i] > max) max = sp[i];
        double sum = 0.0;
        for (int64_t i = 0; i < n; i++) {
            if (sp[i] == -INFINITY) dp[i] = 0.0f;
            else { const float val = or_fp16_to_fp32(T_EXP[or_fp32_to_fp16(sp[i] - max)]); sum += (double)val; dp[i] = val; }
        }
        const float inv = (float)(1.0 / sum);
        for (int64_t i = 0; i < n; i++) dp[i] *= inv;
    }
}

/* the attention chain scale_inplace (:10733) -> diag_mask_inf_inplace (:11268-11316) -> soft_max_inplace on x[nz][nr][nc] */
void or_scale_mask_soft_max(float *x, int64_t nc, int64_t nr, int64_t nz, float scale, int n_past) {
    for (int64_t k = 0; k < nz; k++)
        for (int64_t j = 0; j < nr; j++) {
            float *row = x + (k * nr + j) * nc;
            for (int64_t i = 0; i < nc; i++) row[i] *= scale;
            for (int64_t i = n_past; i < nc; i++) if (i > n_past + j) row[i] = -INFINITY;
        }
    or_soft_max(x, x, nc, nr * nz);
}

/* LC/ggml.c:11774-11901, modes 0 (adjacent pairs over ALL of ne0) and 2 (NeoX halves, theta continues across chunks).
 * theta is advanced by repeated f32 multiplication (:11864, :11883); `x0*c - x1*s` fuses to fma(x0, c, -(x1*s)) in the
 * reference build (gcc contracts the second product's consumer), reproduced below. */
void or_rope(float *x, int64_t ne0, int64_t ne1, int64_t ne2, int n_past, int n_dims, int mode, float freq_base, float freq_scale) {
    const float theta_scale = powf(freq_base, -2.0f/n_dims);
    const int is_neox = mode & 2;
    for (int64_t i2 = ((mode & 1) == 0 ? 0 : n_past); i2 < ne2; i2++) {
        const int64_t p = ((mode & 1) == 0 ? n_past + i2 : i2);
        for (int64_t i1 = 0; i1 < ne1; i1++) {
            float theta = freq_scale * (float)p;
            float *row = x + (i2 * ne1 + i1) * ne0;
            if (!is_neox) {
                for (int64_t i0 = 0; i0 < ne0; i0 += 2) {
                    const float c = cosf(theta), s = sinf(theta);
                    theta *= theta_scale;
                    const float x0 = row[i0], x1 = row[i0 + 1];
                    row[i0]     = fmaf(x0, c, -(x1*s));
                    row[i0 + 1] = fmaf(x0, s, x1*c);
                }
            } else {
                for (int64_t ib = 0; ib < ne0/n_dims; ++ib)
                    for (int64_t ic = 0; ic < n_dims; ic += 2) {
                        const float c = cosf(theta), s = sinf(theta);
                        theta *= theta_scale;
                        const int64_t i0 = ib*n_dims + ic/2;
                        const float x0 = row[i0], x1 = row[i0 + n_dims/2];
                        row[i0]            = fmaf(x0, c, -(x1*s));
                        row[i0 + n_dims/2] = fmaf(x0, s, x1*c);
                    }
            }
        }
    }
}
