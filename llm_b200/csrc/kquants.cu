// llm_b200/csrc/kquants.cu -- K-quant weights Q2_K .. Q6_K (256-element super-blocks, LC/k_quants.h:28-120): bit-exact mat-mul against Q8_K activations.
//
// Serves ggml_compute_forward_mul_mat (LC/ggml.c:10397-10586) when src0 is Q2_K / Q3_K / Q4_K / Q5_K / Q6_K: type_traits[] pairs them with
// vec_dot_type = Q8_K (LC/ggml.c:1700-1737), i.e. the INIT phase runs quantize_row_q8_K on every src1 row and the COMPUTE phase calls
// ggml_vec_dot_q{2,3,4,5,6}_K_q8_K (LC/k_quants.c:1240, :1763, :2492, :3023, :3592) once per (weight row, src1 row).
//
// What "the reference's result" is (the x86 build of crates/ggml/sys/build.rs: -mavx2 -mfma -mf16c, QK_K = 256):
//   * quantize_row_q8_K_reference (LC/k_quants.c:1133-1168): the FIRST element of largest magnitude gives iscale = -128/max, q = min(127, nearest_int(iscale*x)),
//     d = 1/iscale, bsums = sums of 16.  gcc contracts nearest_int's `iscale*x + 12582912.f` into ONE fused multiply-add (default -ffp-contract=fast; verified in the
//     disassembly of oracle/_ref: vfmadd132ps) -- so does this file (__fmaf_rn); a separately rounded product differs in the last bit for ~1e-3 of the inputs.
//   * the AVX2 dot products keep 8 int32 lanes per super-block (lane L = bytes 4L..4L+3 of every 32-byte group, each group weighted by its 6-bit / 8-bit sub-block
//     scale: all integer, exact) and ONE 8-lane f32 accumulator: acc_L = fma(d, (float)sumi_L, acc_L) per super-block in order, d = y.d * fp16(x.d); the result is
//     hsum_float_8(acc) (LC/k_quants.c:1193-1199) plus the "mins" term: Q4_K keeps 4 f32 lanes acc_m[t] = fma(dmin, (float)prod[t], acc_m[t]) reduced as
//     (m0+m2)+(m1+m3) (:2618-2620, :2652-2655); Q5_K a scalar summs = fma(dmin, (float)(prod0+..+prod3), summs) (:3157-3160; contracted, vfmadd231ss);
//     Q2_K folds its mins into the SAME 8 lanes first: acc_L = fma(dmin, (float)(m[2L] bsums[2L] + m[2L+1] bsums[2L+1]), acc_L) (:1346-1350).
// One warp per (weight row, src1 row): lane = 8*part + L owns byte column L of two of the eight 32-byte groups; the integer partials meet by shuffles,
// every lane then carries the f32 chain of its L.  This is a correctness-first kernel (weights re-read per src1 row through L2): the fused decode schedule
// and the tensor-core prefill GEMMs serve the five classic formats only.
#include "kernels.cuh"

namespace b200 {

namespace {

constexpr int QKK = 256;
struct __align__(4) BlockQ8K { float d; int8_t qs[QKK]; int16_t bsums[QKK / 16]; };   // LC/k_quants.h:112-117
static_assert(sizeof(BlockQ8K) == 292, "block_q8_K");

__device__ __forceinline__ int nearest_int_fma(float a, float b) {        // nearest_int(a*b) of LC/k_quants.c:36-42 as compiled (see the header)
    const float val = __fmaf_rn(a, b, 12582912.f);
    return (__float_as_int(val) & 0x007fffff) - 0x00400000;
}

// one warp per super-block; lane holds elements 8*lane .. 8*lane+7
__global__ void __launch_bounds__(128) quantize_q8k_kernel(const float *__restrict__ x, int64_t ldx, BlockQ8K *__restrict__ y, int64_t nsb, int64_t total) {
    const int lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (gw >= total) return;
    const int64_t row = gw / nsb, sb = gw % nsb;
    const float *src = x + row * ldx + sb * QKK + lane * 8;
    float v[8];
    *(float4 *)&v[0] = *(const float4 *)src; *(float4 *)&v[4] = *(const float4 *)(src + 4);
    float amax = 0.f, mx = 0.f;
    int idx = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { const float ax = fabsf(v[j]); if (ax > amax) { amax = ax; mx = v[j]; idx = lane * 8 + j; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                                   // largest magnitude, ties to the lower index (the scalar loop's strict `>`)
        const float oa = __shfl_xor_sync(0xffffffffu, amax, o), om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (oa > amax || (oa == amax && oi < idx)) { amax = oa; mx = om; idx = oi; }
    }
    BlockQ8K *yb = y + gw;
    int q[8];
    float d = 0.f;
    if (amax == 0.f) {
#pragma unroll
        for (int j = 0; j < 8; j++) q[j] = 0;
    } else {
        const float iscale = __fdiv_rn(-128.f, mx);
#pragma unroll
        for (int j = 0; j < 8; j++) { const int t = nearest_int_fma(iscale, v[j]); q[j] = t < 127 ? t : 127; }
        d = __fdiv_rn(1.f, iscale);
    }
    int s8 = 0;
    uint32_t w0 = 0, w1 = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) { s8 += q[j] + q[j + 4]; w0 |= (uint32_t)(q[j] & 0xff) << (8 * j); w1 |= (uint32_t)(q[j + 4] & 0xff) << (8 * j); }
    *(uint32_t *)(yb->qs + lane * 8) = w0; *(uint32_t *)(yb->qs + lane * 8 + 4) = w1;
    const int s16 = s8 + __shfl_xor_sync(0xffffffffu, s8, 1);
    if ((lane & 1) == 0) yb->bsums[lane >> 1] = (int16_t)s16;
    if (lane == 0) yb->d = d;
}

__device__ __forceinline__ uint32_t ld_u32_a2(const uint8_t *p) {        // 4 bytes from a 2-byte aligned address (block_q6_K is 210 bytes)
    return (uint32_t)*(const uint16_t *)p | ((uint32_t)*(const uint16_t *)(p + 2) << 16);
}
__device__ __forceinline__ float h2f(const uint8_t *p) { return __half2float(*(const __half *)p); }

// 6-bit scale / min j of block_q4_K / block_q5_K (the utmp shuffles of LC/k_quants.c:2604-2609 = get_scale_min_k4, :316-324)
__device__ __forceinline__ int k4_scale(const uint8_t *q, int j) { return j < 4 ? (q[j] & 63) : ((q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4)); }
__device__ __forceinline__ int k4_min(const uint8_t *q, int j) { return j < 4 ? (q[j + 4] & 63) : ((q[j + 4] >> 4) | ((q[j] >> 6) << 4)); }

template <int TYPE> struct KQ;
template <> struct KQ<T_Q2_K> { static constexpr int BYTES = 84, QS = 16, QH = 0; };      // scales[16] qs[64] d dmin   (LC/k_quants.h:34-39)
template <> struct KQ<T_Q3_K> { static constexpr int BYTES = 110, QS = 32, QH = 0; };     // hmask[32] qs[64] scales[12] d (:52-57)
template <> struct KQ<T_Q4_K> { static constexpr int BYTES = 144, QS = 16, QH = 0; };
template <> struct KQ<T_Q5_K> { static constexpr int BYTES = 176, QS = 48, QH = 16; };
template <> struct KQ<T_Q6_K> { static constexpr int BYTES = 210, QS = 0, QH = 128; };

template <int TYPE>
__global__ void __launch_bounds__(128) mul_mat_kq_exact_kernel(const uint8_t *__restrict__ W, const BlockQ8K *__restrict__ X, float *__restrict__ dst, int64_t ldd,
                                                               int64_t N, int64_t nsb, const float *__restrict__ addend, int64_t lda) {
    const int lane = threadIdx.x & 31, L = lane & 7, part = lane >> 3;
    const int64_t n = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5), b = blockIdx.y;
    if (n >= N) return;
    const uint8_t *wrow = W + n * nsb * KQ<TYPE>::BYTES;
    const BlockQ8K *xrow = X + b * nsb;
    float acc = 0.f, accm = 0.f;
    for (int64_t i = 0; i < nsb; i++) {
        const uint8_t *wb = wrow + i * KQ<TYPE>::BYTES;
        const BlockQ8K *xb = xrow + i;
        const float yd = xb->d;
        int p;
        if (TYPE == T_Q2_K || TYPE == T_Q3_K) {
            // 2-bit planes: group (j, k) = bits 2k of qs[32j ..] against q8[128j + 32k ..]; sub-block scale index 8j + 2k + (L >= 4) (get_scale_shuffle_q3k)
            const int j = part >> 1;
            float d;
            uint32_t w, hm = 0;
            if (TYPE == T_Q2_K) {
                d = __fmul_rn(yd, h2f(wb + 80));
                const float dmin = __fmul_rn(-yd, h2f(wb + 82));
                const int prod = (int)(wb[2 * L] >> 4) * (int)xb->bsums[2 * L] + (int)(wb[2 * L + 1] >> 4) * (int)xb->bsums[2 * L + 1];
                acc = __fmaf_rn(dmin, (float)prod, acc);              // the mins go into the SAME accumulator, before the quants (:1350)
                w = *(const uint32_t *)(wb + 16 + 32 * j + 4 * L);
            } else {
                d = __fmul_rn(yd, h2f(wb + 108));
                hm = ld_u32_a2(wb + 4 * L);
                w = ld_u32_a2(wb + 32 + 32 * j + 4 * L);
            }
            p = 0;
#pragma unroll
            for (int kk = 0; kk < 2; kk++) {
                const int k = 2 * (part & 1) + kk, si = 8 * j + 2 * k + (L >= 4);
                const uint32_t lo = (w >> (2 * k)) & 0x03030303u;
                const int xw = *(const int *)(xb->qs + 128 * j + 32 * k + 4 * L);
                if (TYPE == T_Q2_K) p += (int)(wb[si] & 0xF) * __dp4a((int)lo, xw, 0);
                else {
                    const uint32_t q3h = ((~(hm >> (4 * j + k))) & 0x01010101u) << 2;       // 4 where the high bit is NOT set (:1936-1948)
                    const uint8_t *s = wb + 96;                                             // 6-bit scales, the aux[] shuffles of :1907-1913
                    const int wd = si >> 2, c = si & 3;
                    const int sc = (int)(((s[(wd & 1) * 4 + c] >> (4 * (wd >> 1))) & 0xF) | (((s[8 + c] >> (2 * wd)) & 3) << 4)) - 32;
                    p += sc * (__dp4a((int)lo, xw, 0) - __dp4a((int)q3h, xw, 0));
                }
            }
            p += __shfl_xor_sync(0xffffffffu, p, 8);
            p += __shfl_xor_sync(0xffffffffu, p, 16);
            acc = __fmaf_rn(d, (float)p, acc);
        } else if (TYPE == T_Q6_K) {
            const float d = __fmul_rn(yd, h2f(wb + 208));
            const int j = part >> 1, hs = part & 1;
            const uint32_t wl = ld_u32_a2(wb + 64 * j + 32 * hs + 4 * L), wh = ld_u32_a2(wb + 128 + 32 * j + 4 * L);
            const int8_t *sc = (const int8_t *)(wb + 192);
            p = 0;
#pragma unroll
            for (int kk = 0; kk < 2; kk++) {                              // groups k = hs and hs + 2 of this 128-element half share the ql bytes
                const int k = hs + 2 * kk;
                const uint32_t nib = (kk ? (wl >> 4) : wl) & 0x0F0F0F0Fu;
                const uint32_t q = nib | (((wh >> (2 * k)) & 0x03030303u) << 4);
                const int xw = *(const int *)(xb->qs + 128 * j + 32 * k + 4 * L);
                const int dot = __dp4a((int)q, xw, 0) - 32 * __dp4a(0x01010101, xw, 0);
                p += (int)sc[2 * (4 * j + k) + (L >= 4)] * dot;
            }
            p += __shfl_xor_sync(0xffffffffu, p, 8);
            p += __shfl_xor_sync(0xffffffffu, p, 16);
            acc = __fmaf_rn(d, (float)p, acc);
        } else {
            const float d = __fmul_rn(yd, h2f(wb)), dmin = __fmul_rn(-yd, h2f(wb + 2));
            const uint8_t *scq = wb + 4;
            const uint32_t w = *(const uint32_t *)(wb + KQ<TYPE>::QS + 32 * part + 4 * L);
            uint32_t lo = w & 0x0F0F0F0Fu, hi = (w >> 4) & 0x0F0F0F0Fu;
            if (TYPE == T_Q5_K) {
                const uint32_t hb = *(const uint32_t *)(wb + KQ<TYPE>::QH + 4 * L);
                lo |= ((hb >> (2 * part)) & 0x01010101u) << 4;
                hi |= ((hb >> (2 * part + 1)) & 0x01010101u) << 4;
            }
            const int x0 = *(const int *)(xb->qs + 64 * part + 4 * L), x1 = *(const int *)(xb->qs + 64 * part + 32 + 4 * L);
            p = k4_scale(scq, 2 * part) * __dp4a((int)lo, x0, 0) + k4_scale(scq, 2 * part + 1) * __dp4a((int)hi, x1, 0);
            p += __shfl_xor_sync(0xffffffffu, p, 8);
            p += __shfl_xor_sync(0xffffffffu, p, 16);
            acc = __fmaf_rn(d, (float)p, acc);
            const int t = lane & 3;                                       // prod[t] of :2614-2617 (every lane carries the copy of its t)
            const int q8a = (int)(int16_t)(xb->bsums[4 * t] + xb->bsums[4 * t + 1]), q8b = (int)(int16_t)(xb->bsums[4 * t + 2] + xb->bsums[4 * t + 3]);
            int prod = k4_min(scq, 2 * t) * q8a + k4_min(scq, 2 * t + 1) * q8b;
            if (TYPE == T_Q5_K) {
                prod += __shfl_xor_sync(0xffffffffu, prod, 1);
                prod += __shfl_xor_sync(0xffffffffu, prod, 2);
            }
            accm = __fmaf_rn(dmin, (float)prod, accm);
        }
    }
    float v = acc;                                                        // hsum_float_8: ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7))
    v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 4));
    v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 2));
    v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 1));
    if (TYPE == T_Q4_K) {                                                 // (m0+m2) + (m1+m3)
        accm = __fadd_rn(accm, __shfl_xor_sync(0xffffffffu, accm, 2));
        accm = __fadd_rn(accm, __shfl_xor_sync(0xffffffffu, accm, 1));
    }
    if (TYPE == T_Q4_K || TYPE == T_Q5_K) v = __fadd_rn(v, accm);
    if (lane == 0) dst[b * ldd + n] = addend ? __fadd_rn(v, addend[b * lda + n]) : v;
}

}  // namespace

size_t q8k_bytes(int64_t K, int64_t B) { return (size_t)B * (size_t)(K / QKK) * sizeof(BlockQ8K); }

void quantize_act_q8k(const float *x, int64_t ldx, void *y, int64_t K, int64_t B, cudaStream_t st) {
    B200_ASSERT(K % QKK == 0 && ldx % 4 == 0 && ((uintptr_t)x & 15) == 0);
    const int64_t nsb = K / QKK, total = nsb * B;
    if (!total) return;
    quantize_q8k_kernel<<<(unsigned)((total + 3) / 4), 128, 0, st>>>(x, ldx, (BlockQ8K *)y, nsb, total);
    B200_CHECK(cudaGetLastError());
}

void mul_mat_kq_exact(int type, const void *w_raw, const void *xq8k, float *dst, int64_t ldd, int64_t K, int64_t N, int64_t B, const float *addend, int64_t lda,
                      cudaStream_t st) {
    B200_ASSERT(is_kquant(type) && K % QKK == 0);
    if (!N || !B) return;
    B200_ASSERT(B <= 65535);
    const dim3 grid((unsigned)((N + 3) / 4), (unsigned)B);
    const int64_t nsb = K / QKK;
    switch (type) {
        case T_Q2_K: mul_mat_kq_exact_kernel<T_Q2_K><<<grid, 128, 0, st>>>((const uint8_t *)w_raw, (const BlockQ8K *)xq8k, dst, ldd, N, nsb, addend, lda); break;
        case T_Q3_K: mul_mat_kq_exact_kernel<T_Q3_K><<<grid, 128, 0, st>>>((const uint8_t *)w_raw, (const BlockQ8K *)xq8k, dst, ldd, N, nsb, addend, lda); break;
        case T_Q4_K: mul_mat_kq_exact_kernel<T_Q4_K><<<grid, 128, 0, st>>>((const uint8_t *)w_raw, (const BlockQ8K *)xq8k, dst, ldd, N, nsb, addend, lda); break;
        case T_Q5_K: mul_mat_kq_exact_kernel<T_Q5_K><<<grid, 128, 0, st>>>((const uint8_t *)w_raw, (const BlockQ8K *)xq8k, dst, ldd, N, nsb, addend, lda); break;
        case T_Q6_K: mul_mat_kq_exact_kernel<T_Q6_K><<<grid, 128, 0, st>>>((const uint8_t *)w_raw, (const BlockQ8K *)xq8k, dst, ldd, N, nsb, addend, lda); break;
        default: B200_ASSERT(!"mul_mat_kq_exact: type");
    }
    B200_CHECK(cudaGetLastError());
}

}  // namespace b200
