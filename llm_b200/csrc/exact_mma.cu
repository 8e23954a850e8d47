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
// CTA = 8 warps = 64 tokens x 32 weight rows (L2 bytes per unit of work ~ 72/rows + 18/tokens); warp tile = 32 tokens x 8 rows
// (2 token tiles x 4 row pairs, 64 accumulator registers).  K loop: 3-stage cp.async pipeline over 4-block stages.
//
// k ORDER inside a 16-element chunk.  The k index is a dummy, so the chunk's elements are assigned to MMA k slots such that lane L' (elements
// 4L'..4L'+3) occupies k = {2L', 2L'+1, 2L'+8, 2L'+9}: exactly the four k slots ONE thread (t = L') holds in both the A and the B fragment.
//   A: thread (g, t) needs elements 4t..4t+3 of a chunk for tokens g and g+8: quantize_act_f16 writes xh directly in that fragment order
//      ([16-token tile][block][chunk][thread] x 16 B), so one conflict-free 16-byte shared load IS the {a0..a3} register quad (no ldmatrix, no moves);
//   B: column g = (row g >> 2 of the pair, lane g & 3) is non-zero only in thread t == (g & 3): one predicated 16-byte load per (row pair, block)
//      of fragments that the CTA expands ONCE per stage (nibbles -> f16) into `sB`; the other 24 threads keep registers that stay zero.
// Per warp and block: 16 HMMA + 64 FFMA + 16 FMUL (the reference's arithmetic) + 16 shared loads.
#include <string.h>

#include "kernels.cuh"

namespace b200 {

namespace {

constexpr int XM = 64, XN = 32, XKB = 4, XST = 3, XTH = 256;    // 3 stages of ~25 KB + 9 KB of expanded operands, two CTAs (16 warps) per SM
constexpr int XTG = XM / 32;                                     // warps along the token axis (the other XTH/32/XTG split the rows, 8 each)
static_assert((XTH / 32 / XTG) * 8 == XN, "warp grid must cover the CTA tile");
constexpr int XS_STRIDE = XKB * 8 + 8;            // float2 {d, s} per block
constexpr int XB_BYTES = (XN / 2) * XKB * 8 * 16; // expanded B fragments of one stage: [row pair][block][column g] x 16 B
constexpr int XW_BYTES = XN * XKB * 8;            // {d, m} as f32 per (row, block)

template <int TYPE> struct Xm {
    static constexpr int QS = (TYPE == T_Q8_0) ? 32 : 16;
    static constexpr int DM = (TYPE == T_Q4_1 || TYPE == T_Q5_1) ? 4 : 2;
    static constexpr bool QH = (TYPE == T_Q5_0 || TYPE == T_Q5_1);
    static constexpr bool MIN = (TYPE == T_Q4_1 || TYPE == T_Q5_1);
    static constexpr int A_BYTES = (XM / 16) * XKB * 1024, S_BYTES = XM * XS_STRIDE;
    static constexpr int Q_BYTES = XN * XKB * QS, D_BYTES = (XN * XKB * DM + 15) & ~15, H_BYTES = QH ? XN * XKB * 4 : 0;
    static constexpr int STAGE = A_BYTES + S_BYTES + Q_BYTES + D_BYTES + H_BYTES;
    static constexpr int SMEM = XST * STAGE + XB_BYTES + XW_BYTES;
    // fp16 magic: 0x6400 | n == 1024 + n; subtracting 1024 + bias leaves the signed integer weight exactly
    static constexpr uint32_t OFF = TYPE == T_Q4_0 ? 0x64086408u : TYPE == T_Q5_0 ? 0x64106410u : TYPE == T_Q8_0 ? 0x64806480u : 0x64006400u;
};

__device__ __forceinline__ void cpa16(uint32_t smem, const void *g, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cpa8(uint32_t smem, const void *g, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cpa4(uint32_t smem, const void *g, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};"
                 : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.0f));
}
// bytes (lo, hi) of `v` -> half2(1024 + lo, 1024 + hi) - off   (exact)
__device__ __forceinline__ uint32_t bytes_to_half2(uint32_t v, uint32_t sel, uint32_t off_h2) {
    const uint32_t p = __byte_perm(v, 0x64646464u, sel);
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
    uint8_t *const sB = smem + XST * T::STAGE, *const sW = sB + XB_BYTES;
    const uint32_t smem_u = (uint32_t)__cvta_generic_to_shared(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int64_t m_base = (int64_t)blockIdx.y * XM, n_base = (int64_t)blockIdx.x * XN;
    const int nb = (int)w.nb, ktiles = (nb + XKB - 1) / XKB;

    // ---- per-thread copy plan: byte offsets of this thread's rows inside each plane (32-bit: planes are < 4 GB), fixed for the whole K loop ----
    auto clamp_m = [&](int r) { const int64_t m = m_base + r; return m < B ? m : B - 1; };
    auto clamp_n = [&](int r) { const int64_t n = n_base + r; return n < w.N ? n : w.N - 1; };
    uint32_t a_off[XM / 16];                                            // activations: token tile i of the CTA = 256 x 16 B per block-stage, contiguous in xh
    const int64_t n_tiles = (B + 15) / 16;
#pragma unroll
    for (int i = 0; i < XM / 16; i++) {
        const int64_t tile = m_base / 16 + i < n_tiles ? m_base / 16 + i : n_tiles - 1;
        a_off[i] = (uint32_t)(tile * nb * 1024 + tid * 16);
    }
    const uint32_t s_off = (uint32_t)((clamp_m(tid >> 2) * nb + (tid & 3)) * 8);          // {d, s}: thread -> (token tid/4, block tid%4)
    constexpr int QC = T::QS / 16;                                      // 16-byte copies per weight block
    const uint32_t q_off = (uint32_t)(clamp_n(tid / (XKB * QC)) * nb * T::QS + (tid % (XKB * QC)) * 16);
    constexpr int DC = XKB * T::DM / 4;                                 // 4-byte copies per weight row per stage
    const uint32_t d_off = (uint32_t)(clamp_n(tid / DC) * nb * T::DM + (tid % DC) * 4);
    const uint32_t h_off = (uint32_t)((clamp_n(tid >> 2) * nb + (tid & 3)) * 4);

    auto load_stage = [&](int s, int kt) {
        const uint32_t sA = smem_u + s * T::STAGE, sS = sA + T::A_BYTES, sQ = sS + T::S_BYTES, sD = sQ + T::Q_BYTES, sH = sD + T::D_BYTES;
        const int b0 = kt * XKB;
        if (b0 + XKB <= nb) {                                          // every k-tile but possibly the last: no per-copy predicates
            const uint8_t *src = (const uint8_t *)xh + (size_t)b0 * 1024;
#pragma unroll
            for (int i = 0; i < XM / 16; i++) cpa16(sA + i * (XKB * 1024) + tid * 16, src + a_off[i], 16);
            if (tid < XM * XKB) cpa8(sS + (tid >> 2) * XS_STRIDE + (tid & 3) * 8, (const uint8_t *)xds + (size_t)b0 * 8 + s_off, 8);
            if (tid < XN * XKB * QC) cpa16(sQ + tid * 16, w.qs + (size_t)b0 * T::QS + q_off, 16);
            if (tid < XN * DC) cpa4(sD + tid * 4, (const uint8_t *)w.dm + (size_t)b0 * T::DM + d_off, 4);
            if (T::QH && tid < XN * XKB) cpa4(sH + tid * 4, (const uint8_t *)w.qh + (size_t)b0 * 4 + h_off, 4);
            return;
        }
        {
            const int ok = b0 + (tid >> 6) < nb ? 16 : 0;                                  // 64 copies per block
            const uint8_t *src = (const uint8_t *)xh + (size_t)b0 * 1024;
#pragma unroll
            for (int i = 0; i < XM / 16; i++) cpa16(sA + i * (XKB * 1024) + tid * 16, src + (ok ? a_off[i] : 0u), ok);
        }
        if (tid < XM * XKB) cpa8(sS + (tid >> 2) * XS_STRIDE + (tid & 3) * 8, (const uint8_t *)xds + (size_t)b0 * 8 + (b0 + (tid & 3) < nb ? s_off : 0u), b0 + (tid & 3) < nb ? 8 : 0);
        if (tid < XN * XKB * QC) {
            const int ok = b0 + (tid % (XKB * QC)) / QC < nb ? 16 : 0;
            cpa16(sQ + tid * 16, w.qs + (size_t)b0 * T::QS + (ok ? q_off : 0u), ok);
        }
        if (tid < XN * DC) {
            const int ok = b0 + (tid % DC) * (4 / T::DM) < nb ? 4 : 0;                    // nb is even: a 4-byte copy of two fp16 scales is all-or-nothing
            cpa4(sD + tid * 4, (const uint8_t *)w.dm + (size_t)b0 * T::DM + (ok ? d_off : 0u), ok);
        }
        if (T::QH && tid < XN * XKB) {
            const int ok = b0 + (tid & 3) < nb ? 4 : 0;
            cpa4(sH + tid * 4, (const uint8_t *)w.qh + (size_t)b0 * 4 + (ok ? h_off : 0u), ok);
        }
    };

    // ---- once per stage: nibbles -> f16 B fragments (sB) and fp16 scales -> f32 (sW) ----
    auto expand_stage = [&](int s) {
        const uint8_t *sQ = smem + s * T::STAGE + T::A_BYTES + T::S_BYTES, *sD = sQ + T::Q_BYTES, *sH = sD + T::D_BYTES;
#pragma unroll
        for (int it = 0; it < XN * XKB * 4 / XTH; it++) {
            const int i = tid + it * XTH, cg = i & 3, b = (i >> 2) & (XKB - 1), r = i >> 4;      // (row r, block b, lane pair cg / cg + 4)
            uint32_t lo, hi;                                                                       // 4 weights of chunk 0 / chunk 1, one per byte
            if (TYPE == T_Q8_0) {
                lo = *(const uint32_t *)(sQ + (r * XKB + b) * 32 + 4 * cg) ^ 0x80808080u;         // int8 -> biased 0..255
                hi = *(const uint32_t *)(sQ + (r * XKB + b) * 32 + 16 + 4 * cg) ^ 0x80808080u;
            } else {
                const uint32_t v = *(const uint32_t *)(sQ + (r * XKB + b) * 16 + 4 * cg);
                lo = v & 0x0F0F0F0Fu; hi = (v >> 4) & 0x0F0F0F0Fu;
                if (T::QH) {
                    const uint32_t qh = *(const uint32_t *)(sH + (r * XKB + b) * 4);
                    lo |= ((((qh >> (4 * cg)) & 0xFu) * 0x00204081u) & 0x01010101u) << 4;          // bit j -> bit 4 of byte j
                    hi |= ((((qh >> (16 + 4 * cg)) & 0xFu) * 0x00204081u) & 0x01010101u) << 4;
                }
            }
            uint4 f;
            f.x = bytes_to_half2(lo, 0x4140u, T::OFF); f.y = bytes_to_half2(lo, 0x4342u, T::OFF);
            f.z = bytes_to_half2(hi, 0x4140u, T::OFF); f.w = bytes_to_half2(hi, 0x4342u, T::OFF);
            *(uint4 *)(sB + ((((r >> 1) * XKB + b) * 8) + (r & 1) * 4 + cg) * 16) = f;
        }
        if (tid < XN * XKB) {
            float2 dm;
            if (T::MIN) { const __half2 h = *(const __half2 *)(sD + tid * 4); dm = make_float2(__low2float(h), __high2float(h)); }
            else dm = make_float2(__half2float(*(const __half *)(sD + tid * 2)), 0.f);
            *(float2 *)(sW + tid * 8) = dm;
        }
    };

    const int tg = warp % XTG, rg = warp / XTG;
    const bool b_active = t == (g & 3);

    float acc[2][4][8];                                     // [token tile][row pair][chunk*4 + e]
    float summs[2][4][2];
    uint4 bf[4];
#pragma unroll
    for (int p = 0; p < 4; p++) bf[p] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int p = 0; p < 4; p++) { summs[i][p][0] = summs[i][p][1] = 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) acc[i][p][j] = 0.f; }

#pragma unroll
    for (int s = 0; s < XST - 1; s++) { if (s < ktiles) load_stage(s, s); asm volatile("cp.async.commit_group;"); }

    for (int kt = 0; kt < ktiles; kt++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(XST - 2));
        __syncthreads();                                    // stage kt landed for everyone; everyone is done with sB / sW / stage kt-1
        { const int nk = kt + XST - 1; if (nk < ktiles) load_stage(nk % XST, nk); asm volatile("cp.async.commit_group;"); }
        expand_stage(kt % XST);
        __syncthreads();
        const uint8_t *sA = smem + (kt % XST) * T::STAGE, *sS = sA + T::A_BYTES;
#pragma unroll 1
        for (int b = 0; b < XKB; b++) {
            float dw[4], mw[4];
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const float2 dm = *(const float2 *)(sW + ((rg * 8 + p * 2 + (t >> 1)) * XKB + b) * 8);   // weight row of this thread's two C columns
                dw[p] = dm.x; mw[p] = dm.y;
                if (b_active) bf[p] = *(const uint4 *)(sB + (((rg * 4 + p) * XKB + b) * 8 + g) * 16);
            }
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                const int tok0 = tg * 32 + mt * 16;
                const uint8_t *at = sA + ((tg * 2 + mt) * XKB + b) * 1024 + lane * 16;
                const uint4 xa = *(const uint4 *)at;                // chunk 0: the A fragment {a0, a1, a2, a3} as stored by quantize_act_f16
                const uint4 xb = *(const uint4 *)(at + 512);        // chunk 1
                const float2 x0 = *(const float2 *)(sS + (tok0 + g) * XS_STRIDE + b * 8);
                const float2 x1 = *(const float2 *)(sS + (tok0 + g + 8) * XS_STRIDE + b * 8);
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    float c0[4], c1[4];
                    mma_f16(c0, xa.x, xa.y, xa.z, xa.w, bf[p].x, bf[p].y);
                    mma_f16(c1, xb.x, xb.y, xb.z, xb.w, bf[p].z, bf[p].w);
                    const float d0 = __fmul_rn(dw[p], x0.x), d1 = __fmul_rn(dw[p], x1.x);
                    float *a = acc[mt][p];
                    a[0] = __fmaf_rn(d0, c0[0], a[0]); a[1] = __fmaf_rn(d0, c0[1], a[1]);
                    a[2] = __fmaf_rn(d1, c0[2], a[2]); a[3] = __fmaf_rn(d1, c0[3], a[3]);
                    a[4] = __fmaf_rn(d0, c1[0], a[4]); a[5] = __fmaf_rn(d0, c1[1], a[5]);
                    a[6] = __fmaf_rn(d1, c1[2], a[6]); a[7] = __fmaf_rn(d1, c1[3], a[7]);
                    if (T::MIN) { summs[mt][p][0] = __fmaf_rn(mw[p], x0.y, summs[mt][p][0]); summs[mt][p][1] = __fmaf_rn(mw[p], x1.y, summs[mt][p][1]); }
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;");

    // ---- hsum_float_8 (LC/ggml.c:608-616): ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7)); lanes 2tau, 2tau+1 (+4) live in thread tau = t & 1 ----
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int p = 0; p < 4; p++)
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {                              // token g (hh = 0) / g + 8 (hh = 1)
                const float *a = acc[mt][p];
                const float r_lo = __fadd_rn(a[4 + 2 * hh], a[2 * hh]);              // a_{4+2tau} + a_{2tau}
                const float r_hi = __fadd_rn(a[5 + 2 * hh], a[1 + 2 * hh]);          // a_{5+2tau} + a_{1+2tau}
                const float s0 = __fadd_rn(r_lo, __shfl_xor_sync(0xffffffffu, r_lo, 1));      // (a0+a4) + (a2+a6)
                const float s1 = __fadd_rn(r_hi, __shfl_xor_sync(0xffffffffu, r_hi, 1));      // (a1+a5) + (a3+a7)
                float v = __fadd_rn(s0, s1);
                if (T::MIN) v = __fadd_rn(v, summs[mt][p][hh]);
                const int64_t m = m_base + tg * 32 + mt * 16 + g + hh * 8;
                const int64_t n = n_base + rg * 8 + p * 2 + (t >> 1);
                if ((t & 1) == 0 && m < B && n < w.N) dst[m * ldd + n] = addend ? __fadd_rn(v, addend[m * lda + n]) : v;
            }
}

// quantize_act with the quants written as fp16 (exact: |q| <= 127): same arithmetic as quantize_act_kernel (quant.cu)
template <bool Q81>
__global__ void __launch_bounds__(256) quantize_act_f16_kernel(const float *__restrict__ x, int64_t ldx, __half *__restrict__ xh, float2 *__restrict__ ds,
                                                               int64_t nbk, int64_t total_blocks, int64_t rows) {
    const int64_t blk = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (blk >= total_blocks) return;
    const int lane = threadIdx.x & 31;
    const int64_t row = blk / nbk, b = blk - row * nbk;
    // xh is stored in MMA A-fragment order (see the kernel header): [16-token tile][block][chunk c][g][t]{tok g: j0 j1 | tok g+8: j0 j1 | tok g: j2 j3 | tok g+8: j2 j3}
    // for element e = 16c + 4t + j of token 16*tile + 8*h + g
    const int c = lane >> 4, tt = (lane >> 2) & 3, j = lane & 3, gg = (int)(row & 7), h = (int)((row >> 3) & 1);
    __half *out = xh + ((row >> 4) * nbk + b) * 512 + c * 256 + (gg * 4 + tt) * 8 + (j >> 1) * 4 + h * 2 + (j & 1);
    if (row >= rows) { *out = __float2half_rn(0.f); return; }           // padding tokens of the last 16-token tile: defined (zero) operands
    const float v = x[row * ldx + b * QK + lane];
    const float amax = warp_max(fabsf(v));
    const float d = __fdiv_rn(amax, 127.f);
    const float id = (amax != 0.0f) ? __fdiv_rn(127.f, amax) : 0.0f;
    const int q = __float2int_rn(__fmul_rn(v, id));
    const int isum = warp_sum(q);
    *out = __int2half_rn(q);
    if (lane == 0) ds[blk] = Q81 ? make_float2(d, __fmul_rn(d, (float)isum)) : make_float2(__half2float(__float2half_rn(d)), (float)isum);
}

template <int TYPE>
void launch_xmma(const QWeight &w, const __half *xh, const float2 *xds, float *dst, int64_t ldd, int64_t B, const float *addend, int64_t lda, cudaStream_t st) {
    using T = Xm<TYPE>;
    constexpr int smem = T::SMEM;
    static bool set = false;
    if (!set) { B200_CHECK(cudaFuncSetAttribute(mm_exact_mma_kernel<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
    dim3 grid((unsigned)((w.N + XN - 1) / XN), (unsigned)((B + XM - 1) / XM));
    mm_exact_mma_kernel<TYPE><<<grid, XTH, smem, st>>>(w, xh, xds, dst, ldd, B, addend, lda);
    B200_CHECK(cudaGetLastError());
}

}  // namespace

void quantize_act_f16(int vdt, const float *x, int64_t ldx, __half *xh, float2 *ds, int64_t K, int64_t B, cudaStream_t st) {
    const int64_t nbk = K / QK, total = nbk * ((B + 15) / 16 * 16);      // every token slot of the last tile is written
    if (total == 0) return;
    if (vdt == T_Q8_1) quantize_act_f16_kernel<true><<<(unsigned)((total + 7) / 8), 256, 0, st>>>(x, ldx, xh, ds, nbk, total, B);
    else               quantize_act_f16_kernel<false><<<(unsigned)((total + 7) / 8), 256, 0, st>>>(x, ldx, xh, ds, nbk, total, B);
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
