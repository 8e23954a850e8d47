// llm_b200/csrc/exact_mma.cu -- bit-exact ggml_mul_mat for a BATCH of activation rows (prefill), with the integer work on tensor cores.
//
// The reference needs, per output and per 32-element block, the EIGHT four-element partial dots S_L = sum_{t<4} w[4L+t] x[4L+t] separately:
// each one feeds its own f32 lane accumulator acc_L = fma(d_w d_x, (float)S_L, acc_L) (AVX2 ggml_vec_dot_q*_q8_*, LC/ggml.c:2434-2457 ...).
// A 32-deep integer MMA sums all 32 products and is useless here; a BLOCK-DIAGONAL B operand is not:
//   mma.m16n8k16 (f16 x f16 -> f32):  A = 16 tokens x 16 elements of the Q8 activations (int8 values, exact in f16)
//                                     B = 16 elements x 8 columns, column c = lane (c & 3) of weight row (c < 4 ? n : n+1): only the 4
//                                         elements of that lane are non-zero (the dequantised integer weight, exact in f16)
//   => C[token][c] = S_lane exactly (integers < 2^17, every product and partial sum exact in f32, accumulator input zero),
//      two MMAs per block (elements 0-15 -> lanes 0-3, elements 16-31 -> lanes 4-7) give all 8 partials of 16 tokens x 2 weight rows.
// The tensor core replaces 8 dp4a + 8 int->float conversions per (token, row, block); what remains on the CUDA cores is the part that
// defines the reference's rounding: one f32 product d_w*d_x and eight ordered fmas.  Results are bit-identical to exact.cu (tests).
//
// CTA = 8 warps = 128 tokens x 16 weight rows; warp w owns rows (2w, 2w+1) for all 8 token tiles (64 accumulator registers).
// K loop: double-buffered cp.async pipeline over 4-block stages (activations as f16 [token][k], weights in the planes layout).
#include <string.h>

#include "kernels.cuh"

namespace b200 {

namespace {

constexpr int XM = 128, XN = 16, XKB = 4, XST = 2, XTH = 256;   // 2 stages of 41 KB: two CTAs (16 warps) per SM
constexpr int XA_STRIDE = XKB * 64 + 16;          // bytes per token row of f16 activations in smem (+16: conflict-free ldmatrix)
constexpr int XS_STRIDE = XKB * 8 + 8;            // float2 {d, s} per block

template <int TYPE> struct Xm {
    static constexpr int QS = (TYPE == T_Q8_0) ? 32 : 16;
    static constexpr int DM = (TYPE == T_Q4_1 || TYPE == T_Q5_1) ? 4 : 2;
    static constexpr bool QH = (TYPE == T_Q5_0 || TYPE == T_Q5_1);
    static constexpr bool MIN = (TYPE == T_Q4_1 || TYPE == T_Q5_1);
    static constexpr int A_BYTES = XM * XA_STRIDE, S_BYTES = XM * XS_STRIDE;
    static constexpr int Q_BYTES = XN * XKB * QS, D_BYTES = XN * XKB * DM, H_BYTES = XN * XKB * 4;
    static constexpr int STAGE = A_BYTES + S_BYTES + Q_BYTES + ((D_BYTES + 15) & ~15) + H_BYTES;
};

__device__ __forceinline__ void cpa16(void *smem, const void *g, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cpa8(void *smem, const void *g, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cpa4(void *smem, const void *g, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void ldm_x4(uint32_t (&r)[4], const void *smem) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"((uint32_t)__cvta_generic_to_shared(smem)));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};"
                 : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.0f));
}
// two small integers (each in one byte of `v`: byte 0 and byte 1) -> half2, minus `off`: exact (0x6400 | n == 1024 + n in fp16)
__device__ __forceinline__ uint32_t ints_to_half2(uint32_t v, uint32_t mask, uint32_t off_h2) {
    uint32_t p = (__byte_perm(v, 0, 0x4140) & mask) | 0x64006400u;        // [n0, 0x64, n1, 0x64] = half2(1024 + n0, 1024 + n1)
    __half2 a, o;
    memcpy(&a, &p, 4); memcpy(&o, &off_h2, 4);
    const __half2 h = __hsub2(a, o);
    uint32_t r; memcpy(&r, &h, 4);
    return r;
}

template <int TYPE>
__global__ void __launch_bounds__(XTH, 2) mm_exact_mma_kernel(const QWeight w, const __half *__restrict__ xh, const float2 *__restrict__ xds,
                                                              float *__restrict__ dst, int64_t ldd, int64_t B,
                                                              const float *__restrict__ addend, int64_t lda) {
    using T = Xm<TYPE>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int64_t m_base = (int64_t)blockIdx.y * XM, n_base = (int64_t)blockIdx.x * XN;
    const int nb = (int)w.nb, ktiles = (nb + XKB - 1) / XKB;
    const int64_t K = (int64_t)nb * QK;

    auto stage = [&](int s) { return smem + (size_t)s * T::STAGE; };
    auto load_stage = [&](int s, int kt) {
        uint8_t *sA = stage(s), *sS = sA + T::A_BYTES, *sQ = sS + T::S_BYTES, *sD = sQ + T::Q_BYTES, *sH = sD + ((T::D_BYTES + 15) & ~15);
        const int b0 = kt * XKB;
        for (int c = tid; c < XM * XKB * 4; c += XTH) {                  // activations: 4 x 16 B per (token, block)
            const int r = c / (XKB * 4), cc = c % (XKB * 4), b = b0 + cc / 4;
            const int64_t m = m_base + r < B ? m_base + r : B - 1;
            const bool ok = b < nb;
            cpa16(sA + r * XA_STRIDE + cc * 16, xh + m * K + (int64_t)(ok ? b : 0) * QK + (cc & 3) * 8, ok ? 16 : 0);
        }
        for (int c = tid; c < XM * XKB; c += XTH) {                      // {d, s} per (token, block)
            const int r = c / XKB, b = b0 + c % XKB;
            const int64_t m = m_base + r < B ? m_base + r : B - 1;
            const bool ok = b < nb;
            cpa8(sS + r * XS_STRIDE + (c % XKB) * 8, xds + m * nb + (ok ? b : 0), ok ? 8 : 0);
        }
        for (int c = tid; c < XN * XKB * (T::QS / 16); c += XTH) {       // weight quants
            const int r = c / (XKB * (T::QS / 16)), cc = c % (XKB * (T::QS / 16)), b = b0 + cc / (T::QS / 16);
            const int64_t n = n_base + r < w.N ? n_base + r : w.N - 1;
            const bool ok = b < nb;
            cpa16(sQ + (r * XKB * T::QS) + cc * 16, w.qs + (n * nb + (ok ? b : 0)) * T::QS + (cc % (T::QS / 16)) * 16, ok ? 16 : 0);
        }
        for (int c = tid; c < XN * XKB * T::DM / 4; c += XTH) {          // weight scales (4-byte copies: 2 blocks of fp16 d, or one {d, m})
            const int per_row = XKB * T::DM / 4, r = c / per_row, cc = c % per_row, b = b0 + cc * (4 / T::DM);
            const int64_t n = n_base + r < w.N ? n_base + r : w.N - 1;
            const bool ok = b < nb;                                       // nb is even
            cpa4(sD + r * XKB * T::DM + cc * 4, (const uint8_t *)w.dm + (n * nb + (ok ? b : 0)) * T::DM, ok ? 4 : 0);
        }
        if (T::QH)
            for (int c = tid; c < XN * XKB; c += XTH) {
                const int r = c / XKB, b = b0 + c % XKB;
                const int64_t n = n_base + r < w.N ? n_base + r : w.N - 1;
                const bool ok = b < nb;
                cpa4(sH + (r * XKB + c % XKB) * 4, w.qh + n * nb + (ok ? b : 0), ok ? 4 : 0);
            }
    };

    // this thread's role in the B operand: column g = lane (g & 3) of row (g < 4 ? pair row 0 : pair row 1)
    const int cg = g & 3;
    const bool b_nonzero = (t >> 1) == (cg & 1);
    const bool b_second = (cg >> 1) != 0;                  // value sits in b1 (k 8..15 of the chunk) instead of b0
    const int kbyte = 2 * t + 8 * (cg >> 1);               // first of the two element indices inside the 16-element chunk
    const int brow = warp * 2 + (g >> 2);                  // weight row (in the CTA tile) feeding this thread's B column
    const int crow = warp * 2 + (t >> 1);                  // weight row of this thread's two C columns

    float acc[8][8];                                        // [token tile][chunk*4 + e]
    float summs[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { summs[i][0] = summs[i][1] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f; }

#pragma unroll
    for (int s = 0; s < XST - 1; s++) { if (s < ktiles) load_stage(s, s); asm volatile("cp.async.commit_group;"); }

    for (int kt = 0; kt < ktiles; kt++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(XST - 2));
        __syncthreads();
        { const int nk = kt + XST - 1; if (nk < ktiles) load_stage(nk % XST, nk); asm volatile("cp.async.commit_group;"); }
        const uint8_t *sA = stage(kt % XST), *sS = sA + T::A_BYTES, *sQ = sS + T::S_BYTES, *sD = sQ + T::Q_BYTES, *sH = sD + ((T::D_BYTES + 15) & ~15);
#pragma unroll
        for (int b = 0; b < XKB; b++) {
            // ---- B fragments of this block: chunk 0 = elements 0-15 (low nibbles), chunk 1 = elements 16-31 (high nibbles) ----
            uint32_t bf[2] = {0u, 0u};
            if (b_nonzero) {
                if (TYPE == T_Q8_0) {
                    const uint8_t *q = sQ + (brow * XKB + b) * 32;
                    const int a0 = (int8_t)q[kbyte], a1 = (int8_t)q[kbyte + 1], c0 = (int8_t)q[16 + kbyte], c1 = (int8_t)q[16 + kbyte + 1];
                    const __half2 h0 = __halves2half2(__int2half_rn(a0), __int2half_rn(a1)), h1 = __halves2half2(__int2half_rn(c0), __int2half_rn(c1));
                    bf[0] = *(const uint32_t *)&h0; bf[1] = *(const uint32_t *)&h1;
                } else {
                    const uint32_t v = *(const uint16_t *)(sQ + (brow * XKB + b) * 16 + kbyte);      // bytes kbyte, kbyte+1
                    uint32_t lo = v & 0x0F0Fu, hi = (v >> 4) & 0x0F0Fu;
                    if (T::QH) {
                        const uint32_t qh = *(const uint32_t *)(sH + (brow * XKB + b) * 4);
                        lo |= (((qh >> kbyte) & 1u) << 4) | (((qh >> (kbyte + 1)) & 1u) << 12);
                        hi |= (((qh >> (16 + kbyte)) & 1u) << 4) | (((qh >> (17 + kbyte)) & 1u) << 12);
                    }
                    const uint32_t off = TYPE == T_Q4_0 ? 0x64086408u : (TYPE == T_Q5_0 ? 0x64106410u : 0x64006400u);   // 1024 + {8, 16, 0}
                    bf[0] = ints_to_half2(lo, 0x00FF00FFu, off);
                    bf[1] = ints_to_half2(hi, 0x00FF00FFu, off);
                }
            }
            float dw, mw = 0.f;
            if (T::MIN) { const __half2 dm = *(const __half2 *)(sD + (crow * XKB + b) * 4); dw = __low2float(dm); mw = __high2float(dm); }
            else dw = __half2float(*(const __half *)(sD + (crow * XKB + b) * 2));
#pragma unroll
            for (int mt = 0; mt < 8; mt++) {
                const int rm = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                uint32_t a0[4], a1[4];
                ldm_x4(a0, sA + rm * XA_STRIDE + b * 64 + (lane >> 4) * 16);            // elements 0-15
                ldm_x4(a1, sA + rm * XA_STRIDE + b * 64 + 32 + (lane >> 4) * 16);       // elements 16-31
                float c0[4], c1[4];
                mma_f16(c0, a0, b_second ? 0u : bf[0], b_second ? bf[0] : 0u);
                mma_f16(c1, a1, b_second ? 0u : bf[1], b_second ? bf[1] : 0u);
                const float2 x0 = *(const float2 *)(sS + (mt * 16 + g) * XS_STRIDE + b * 8);
                const float2 x1 = *(const float2 *)(sS + (mt * 16 + g + 8) * XS_STRIDE + b * 8);
                const float d0 = __fmul_rn(dw, x0.x), d1 = __fmul_rn(dw, x1.x);
                acc[mt][0] = __fmaf_rn(d0, c0[0], acc[mt][0]); acc[mt][1] = __fmaf_rn(d0, c0[1], acc[mt][1]);
                acc[mt][2] = __fmaf_rn(d1, c0[2], acc[mt][2]); acc[mt][3] = __fmaf_rn(d1, c0[3], acc[mt][3]);
                acc[mt][4] = __fmaf_rn(d0, c1[0], acc[mt][4]); acc[mt][5] = __fmaf_rn(d0, c1[1], acc[mt][5]);
                acc[mt][6] = __fmaf_rn(d1, c1[2], acc[mt][6]); acc[mt][7] = __fmaf_rn(d1, c1[3], acc[mt][7]);
                if (T::MIN) { summs[mt][0] = __fmaf_rn(mw, x0.y, summs[mt][0]); summs[mt][1] = __fmaf_rn(mw, x1.y, summs[mt][1]); }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;");

    // ---- hsum_float_8 (LC/ggml.c:608-616): ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7)); lanes 2tau, 2tau+1 (+4) live in thread tau = t & 1 ----
    const int64_t n = n_base + crow;
#pragma unroll
    for (int mt = 0; mt < 8; mt++)
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {                              // token g (hh = 0) / g + 8 (hh = 1)
            const float r_lo = __fadd_rn(acc[mt][4 + 2 * hh], acc[mt][2 * hh]);          // a_{4+2tau} + a_{2tau}
            const float r_hi = __fadd_rn(acc[mt][5 + 2 * hh], acc[mt][1 + 2 * hh]);      // a_{5+2tau} + a_{1+2tau}
            const float s0 = __fadd_rn(r_lo, __shfl_xor_sync(0xffffffffu, r_lo, 1));      // (a0+a4) + (a2+a6)
            const float s1 = __fadd_rn(r_hi, __shfl_xor_sync(0xffffffffu, r_hi, 1));      // (a1+a5) + (a3+a7)
            float v = __fadd_rn(s0, s1);
            if (T::MIN) v = __fadd_rn(v, summs[mt][hh]);
            const int64_t m = m_base + mt * 16 + g + hh * 8;
            if ((t & 1) == 0 && m < B && n < w.N) dst[m * ldd + n] = addend ? __fadd_rn(v, addend[m * lda + n]) : v;
        }
}

// quantize_act with the quants written as fp16 (exact: |q| <= 127): same arithmetic as quantize_act_kernel (quant.cu)
template <bool Q81>
__global__ void __launch_bounds__(256) quantize_act_f16_kernel(const float *__restrict__ x, int64_t ldx, __half *__restrict__ xh, float2 *__restrict__ ds,
                                                               int64_t nbk, int64_t total_blocks) {
    const int64_t blk = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (blk >= total_blocks) return;
    const int lane = threadIdx.x & 31;
    const int64_t row = blk / nbk, b = blk - row * nbk;
    const float v = x[row * ldx + b * QK + lane];
    const float amax = warp_max(fabsf(v));
    const float d = __fdiv_rn(amax, 127.f);
    const float id = (amax != 0.0f) ? __fdiv_rn(127.f, amax) : 0.0f;
    const int q = __float2int_rn(__fmul_rn(v, id));
    const int isum = warp_sum(q);
    xh[blk * QK + lane] = __int2half_rn(q);
    if (lane == 0) ds[blk] = Q81 ? make_float2(d, __fmul_rn(d, (float)isum)) : make_float2(__half2float(__float2half_rn(d)), (float)isum);
}

template <int TYPE>
void launch_xmma(const QWeight &w, const __half *xh, const float2 *xds, float *dst, int64_t ldd, int64_t B, const float *addend, int64_t lda, cudaStream_t st) {
    using T = Xm<TYPE>;
    constexpr int smem = XST * T::STAGE;
    static bool set = false;
    if (!set) { B200_CHECK(cudaFuncSetAttribute(mm_exact_mma_kernel<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
    dim3 grid((unsigned)((w.N + XN - 1) / XN), (unsigned)((B + XM - 1) / XM));
    mm_exact_mma_kernel<TYPE><<<grid, XTH, smem, st>>>(w, xh, xds, dst, ldd, B, addend, lda);
    B200_CHECK(cudaGetLastError());
}

}  // namespace

void quantize_act_f16(int vdt, const float *x, int64_t ldx, __half *xh, float2 *ds, int64_t K, int64_t B, cudaStream_t st) {
    const int64_t nbk = K / QK, total = nbk * B;
    if (total == 0) return;
    if (vdt == T_Q8_1) quantize_act_f16_kernel<true><<<(unsigned)((total + 7) / 8), 256, 0, st>>>(x, ldx, xh, ds, nbk, total);
    else               quantize_act_f16_kernel<false><<<(unsigned)((total + 7) / 8), 256, 0, st>>>(x, ldx, xh, ds, nbk, total);
    B200_CHECK(cudaGetLastError());
}

// bit-exact batched mat-mul on tensor cores; xh = quantized activations as fp16 [B][K], xds = {d, aux} per block (quantize_act_f16)
void mul_mat_q_exact_mma(const QWeight &w, const __half *xh, const float2 *xds, float *dst, int64_t ldd, int64_t B, const float *addend, int64_t lda, cudaStream_t st) {
    if (w.N == 0 || B == 0) return;
    B200_ASSERT(w.nb % 2 == 0);
    switch (w.type) {
        case T_Q4_0: launch_xmma<T_Q4_0>(w, xh, xds, dst, ldd, B, addend, lda, st); break;
        case T_Q4_1: launch_xmma<T_Q4_1>(w, xh, xds, dst, ldd, B, addend, lda, st); break;
        case T_Q5_0: launch_xmma<T_Q5_0>(w, xh, xds, dst, ldd, B, addend, lda, st); break;
        case T_Q5_1: launch_xmma<T_Q5_1>(w, xh, xds, dst, ldd, B, addend, lda, st); break;
        case T_Q8_0: launch_xmma<T_Q8_0>(w, xh, xds, dst, ldd, B, addend, lda, st); break;
        default: B200_ASSERT(!"mul_mat_q_exact_mma: unsupported weight type");
    }
}

}  // namespace b200
