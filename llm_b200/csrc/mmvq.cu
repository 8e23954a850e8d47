// llm_b200/csrc/mmvq.cu -- decode-path mat-vec on quantized weights: the HBM-bound half of ggml_mul_mat.
//
// Arithmetic contract (what the reference's CPU path computes, LC/ggml.c:10397-10586 with the AVX2 vec_dots
// :2434-2457, 2702-2735, 2916-2938, 3166-3191, 3315-3336): per 32-element block an EXACT integer dot of the weight
// quants with the Q8-quantized activations, scaled by the f32 product d_w*d_x and accumulated in f32 (fma); Q4_1/Q5_1 add
// m_w * s_x.  Only the f32 summation ORDER differs from the CPU (one partial per lane + shuffle tree instead of 8 AVX lanes).
//
// Data movement: each weight byte is read exactly once with 128-bit streaming loads (planes layout, common.cuh);
// lane L of a warp owns blocks L, L+32, ... of its rows so a warp load covers 512 contiguous bytes; the 4-11 KB quantized
// activation row is re-read through L1 (it is shared by every CTA on the SM).  Algorithmic bytes per row of K weights:
// K/32 * {18,20,22,24,34}.
#include "kernels.cuh"

namespace b200 {

template <int TYPE> struct WBlock;   // one weight block in registers

template <> struct WBlock<T_Q4_0> { int4 q; float d;
    __device__ __forceinline__ void load(const QWeight &w, int64_t i) { q = ld_stream_int4(w.qs + i * 16); d = f16_bits_to_f32(ld_stream_u16((const uint16_t *)w.dm + i)); } };
template <> struct WBlock<T_Q4_1> { int4 q; float d, m;
    __device__ __forceinline__ void load(const QWeight &w, int64_t i) { q = ld_stream_int4(w.qs + i * 16); const uint32_t dm = ld_stream_u32((const uint32_t *)w.dm + i);
        d = f16_bits_to_f32((uint16_t)(dm & 0xffffu)); m = f16_bits_to_f32((uint16_t)(dm >> 16)); } };
template <> struct WBlock<T_Q5_0> { int4 q; uint32_t qh; float d;
    __device__ __forceinline__ void load(const QWeight &w, int64_t i) { q = ld_stream_int4(w.qs + i * 16); qh = ld_stream_u32(w.qh + i); d = f16_bits_to_f32(ld_stream_u16((const uint16_t *)w.dm + i)); } };
template <> struct WBlock<T_Q5_1> { int4 q; uint32_t qh; float d, m;
    __device__ __forceinline__ void load(const QWeight &w, int64_t i) { q = ld_stream_int4(w.qs + i * 16); qh = ld_stream_u32(w.qh + i); const uint32_t dm = ld_stream_u32((const uint32_t *)w.dm + i);
        d = f16_bits_to_f32((uint16_t)(dm & 0xffffu)); m = f16_bits_to_f32((uint16_t)(dm >> 16)); } };
template <> struct WBlock<T_Q8_0> { int4 q0, q1; float d;
    __device__ __forceinline__ void load(const QWeight &w, int64_t i) { q0 = ld_stream_int4(w.qs + i * 32); q1 = ld_stream_int4(w.qs + i * 32 + 16); d = f16_bits_to_f32(ld_stream_u16((const uint16_t *)w.dm + i)); } };

// nibble word v: byte j low nibble = element 4w+j, high nibble = element 16+4w+j  (LC/ggml.c:1535-1540)
__device__ __forceinline__ int dot_nibbles(const int4 &q, const int4 &xa, const int4 &xb) {
    int s = 0;
    s = __dp4a((int)(q.x & 0x0F0F0F0F), xa.x, s); s = __dp4a((int)((q.x >> 4) & 0x0F0F0F0F), xb.x, s);
    s = __dp4a((int)(q.y & 0x0F0F0F0F), xa.y, s); s = __dp4a((int)((q.y >> 4) & 0x0F0F0F0F), xb.y, s);
    s = __dp4a((int)(q.z & 0x0F0F0F0F), xa.z, s); s = __dp4a((int)((q.z >> 4) & 0x0F0F0F0F), xb.z, s);
    s = __dp4a((int)(q.w & 0x0F0F0F0F), xa.w, s); s = __dp4a((int)((q.w >> 4) & 0x0F0F0F0F), xb.w, s);
    return s;
}
// with the fifth bits: qh bit j <-> element j, bit j+16 <-> element j+16 (LC/ggml.c:1576-1587)
__device__ __forceinline__ int dot_nibbles5(const int4 &q, uint32_t qh, const int4 &xa, const int4 &xb) {
    int s = 0;
    const uint32_t h = qh >> 16;
    s = __dp4a((int)((q.x & 0x0F0F0F0F) | spread4_to_bit4(qh)),       xa.x, s); s = __dp4a((int)(((q.x >> 4) & 0x0F0F0F0F) | spread4_to_bit4(h)),       xb.x, s);
    s = __dp4a((int)((q.y & 0x0F0F0F0F) | spread4_to_bit4(qh >> 4)),  xa.y, s); s = __dp4a((int)(((q.y >> 4) & 0x0F0F0F0F) | spread4_to_bit4(h >> 4)),  xb.y, s);
    s = __dp4a((int)((q.z & 0x0F0F0F0F) | spread4_to_bit4(qh >> 8)),  xa.z, s); s = __dp4a((int)(((q.z >> 4) & 0x0F0F0F0F) | spread4_to_bit4(h >> 8)),  xb.z, s);
    s = __dp4a((int)((q.w & 0x0F0F0F0F) | spread4_to_bit4(qh >> 12)), xa.w, s); s = __dp4a((int)(((q.w >> 4) & 0x0F0F0F0F) | spread4_to_bit4(h >> 12)), xb.w, s);
    return s;
}
__device__ __forceinline__ int dot_bytes(const int4 &q0, const int4 &q1, const int4 &xa, const int4 &xb) {
    int s = 0;
    s = __dp4a(q0.x, xa.x, s); s = __dp4a(q0.y, xa.y, s); s = __dp4a(q0.z, xa.z, s); s = __dp4a(q0.w, xa.w, s);
    s = __dp4a(q1.x, xb.x, s); s = __dp4a(q1.y, xb.y, s); s = __dp4a(q1.z, xb.z, s); s = __dp4a(q1.w, xb.w, s);
    return s;
}

// acc += (d_w * d_x) * S ; accm += m_w * s_x.   xs = {d_x, aux} as written by quantize_act.
template <int TYPE>
__device__ __forceinline__ void block_fma(const WBlock<TYPE> &wb, const int4 &xa, const int4 &xb, const float2 xs, float &acc, float &accm);

template <> __device__ __forceinline__ void block_fma<T_Q4_0>(const WBlock<T_Q4_0> &wb, const int4 &xa, const int4 &xb, const float2 xs, float &acc, float &accm) {
    const float S = __fmaf_rn(-8.0f, xs.y, (float)dot_nibbles(wb.q, xa, xb));      // sum (q-8) x = sum q x - 8 sum x : exact in f32
    acc = __fmaf_rn(__fmul_rn(wb.d, xs.x), S, acc);
}
template <> __device__ __forceinline__ void block_fma<T_Q5_0>(const WBlock<T_Q5_0> &wb, const int4 &xa, const int4 &xb, const float2 xs, float &acc, float &accm) {
    const float S = __fmaf_rn(-16.0f, xs.y, (float)dot_nibbles5(wb.q, wb.qh, xa, xb));
    acc = __fmaf_rn(__fmul_rn(wb.d, xs.x), S, acc);
}
template <> __device__ __forceinline__ void block_fma<T_Q8_0>(const WBlock<T_Q8_0> &wb, const int4 &xa, const int4 &xb, const float2 xs, float &acc, float &accm) {
    acc = __fmaf_rn(__fmul_rn(wb.d, xs.x), (float)dot_bytes(wb.q0, wb.q1, xa, xb), acc);
}
template <> __device__ __forceinline__ void block_fma<T_Q4_1>(const WBlock<T_Q4_1> &wb, const int4 &xa, const int4 &xb, const float2 xs, float &acc, float &accm) {
    acc = __fmaf_rn(__fmul_rn(wb.d, xs.x), (float)dot_nibbles(wb.q, xa, xb), acc);
    accm = __fmaf_rn(wb.m, xs.y, accm);
}
template <> __device__ __forceinline__ void block_fma<T_Q5_1>(const WBlock<T_Q5_1> &wb, const int4 &xa, const int4 &xb, const float2 xs, float &acc, float &accm) {
    acc = __fmaf_rn(__fmul_rn(wb.d, xs.x), (float)dot_nibbles5(wb.q, wb.qh, xa, xb), acc);
    accm = __fmaf_rn(wb.m, xs.y, accm);
}

constexpr int MMVQ_WARPS = 8;

template <int TYPE, int RPW>
__global__ void __launch_bounds__(MMVQ_WARPS * 32) mmvq_kernel(const QWeight w, const int8_t *__restrict__ xq, const float2 *__restrict__ xds,
                                                               float *__restrict__ dst, const float *__restrict__ addend) {
    const int lane = threadIdx.x & 31;
    const int64_t row0 = ((int64_t)blockIdx.x * MMVQ_WARPS + (threadIdx.x >> 5)) * RPW;
    if (row0 >= w.N) return;
    const int nb = (int)w.nb;
    float acc[RPW], accm[RPW];
    int64_t rbase[RPW];
#pragma unroll
    for (int r = 0; r < RPW; r++) { acc[r] = 0.f; accm[r] = 0.f; const int64_t row = row0 + r < w.N ? row0 + r : w.N - 1; rbase[r] = row * nb; }

#pragma unroll 2
    for (int b = lane; b < nb; b += 32) {
        WBlock<TYPE> wb[RPW];
#pragma unroll
        for (int r = 0; r < RPW; r++) wb[r].load(w, rbase[r] + b);
        const int4 xa = __ldg((const int4 *)(xq + (int64_t)b * QK));
        const int4 xb = __ldg((const int4 *)(xq + (int64_t)b * QK) + 1);
        const float2 xs = __ldg(xds + b);
#pragma unroll
        for (int r = 0; r < RPW; r++) block_fma<TYPE>(wb[r], xa, xb, xs, acc[r], accm[r]);
    }
#pragma unroll
    for (int r = 0; r < RPW; r++) {
        float v = warp_sum(acc[r]);
        if (has_min(TYPE)) v += warp_sum(accm[r]);           // hsum(acc) + summs, LC/ggml.c:2735 / 3191
        if (lane == 0 && row0 + r < w.N) dst[row0 + r] = addend ? v + addend[row0 + r] : v;
    }
}

template <int TYPE>
static void launch_mmvq(const QWeight &w, const int8_t *xq, const float2 *xds, float *dst, const float *addend, cudaStream_t st) {
    // rows per warp: enough independent 128-bit loads in flight per lane without starving the grid
    if (w.N >= 8192) {
        const int64_t rows_per_cta = MMVQ_WARPS * 4;
        mmvq_kernel<TYPE, 4><<<(unsigned)((w.N + rows_per_cta - 1) / rows_per_cta), MMVQ_WARPS * 32, 0, st>>>(w, xq, xds, dst, addend);
    } else {
        const int64_t rows_per_cta = MMVQ_WARPS * 2;
        mmvq_kernel<TYPE, 2><<<(unsigned)((w.N + rows_per_cta - 1) / rows_per_cta), MMVQ_WARPS * 32, 0, st>>>(w, xq, xds, dst, addend);
    }
    B200_CHECK(cudaGetLastError());
}

void mul_mat_vec_q(const QWeight &w, const int8_t *xq, const float2 *xds, float *dst, const float *addend, cudaStream_t st) {
    if (w.N == 0) return;
    switch (w.type) {
        case T_Q4_0: launch_mmvq<T_Q4_0>(w, xq, xds, dst, addend, st); break;
        case T_Q4_1: launch_mmvq<T_Q4_1>(w, xq, xds, dst, addend, st); break;
        case T_Q5_0: launch_mmvq<T_Q5_0>(w, xq, xds, dst, addend, st); break;
        case T_Q5_1: launch_mmvq<T_Q5_1>(w, xq, xds, dst, addend, st); break;
        case T_Q8_0: launch_mmvq<T_Q8_0>(w, xq, xds, dst, addend, st); break;
        default: B200_ASSERT(!"mul_mat_vec_q: unsupported weight type");
    }
}

// ---- cross-check GEMM on CUDA cores: every (row, token) pair through the same block_fma; used by tests and as the
//      reference point the tensor-core kernel in mmq.cu is compared with on the device. ---------------------------------
template <int TYPE, int TOK>
__global__ void __launch_bounds__(MMVQ_WARPS * 32) mmq_simple_kernel(const QWeight w, const int8_t *__restrict__ xq, const float2 *__restrict__ xds,
                                                                     float *__restrict__ dst, int64_t ldd, int64_t B,
                                                                     const float *__restrict__ addend, int64_t lda) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * MMVQ_WARPS + (threadIdx.x >> 5);
    const int64_t t0 = (int64_t)blockIdx.y * TOK;
    if (row >= w.N) return;
    const int nb = (int)w.nb;
    float acc[TOK], accm[TOK];
#pragma unroll
    for (int t = 0; t < TOK; t++) { acc[t] = 0.f; accm[t] = 0.f; }
    for (int b = lane; b < nb; b += 32) {
        WBlock<TYPE> wb;
        wb.load(w, row * nb + b);
#pragma unroll
        for (int t = 0; t < TOK; t++) {
            const int64_t tok = t0 + t < B ? t0 + t : B - 1;
            const int4 xa = __ldg((const int4 *)(xq + (tok * nb + b) * QK));
            const int4 xb = __ldg((const int4 *)(xq + (tok * nb + b) * QK) + 1);
            const float2 xs = __ldg(xds + tok * nb + b);
            block_fma<TYPE>(wb, xa, xb, xs, acc[t], accm[t]);
        }
    }
#pragma unroll
    for (int t = 0; t < TOK; t++) {
        float v = warp_sum(acc[t]);
        if (has_min(TYPE)) v += warp_sum(accm[t]);
        if (lane == 0 && t0 + t < B) dst[(t0 + t) * ldd + row] = addend ? v + addend[(t0 + t) * lda + row] : v;
    }
}

template <int TYPE>
static void launch_simple(const QWeight &w, const int8_t *xq, const float2 *xds, float *dst, int64_t ldd, int64_t B, const float *addend, int64_t lda, cudaStream_t st) {
    constexpr int TOK = 8;
    dim3 grid((unsigned)((w.N + MMVQ_WARPS - 1) / MMVQ_WARPS), (unsigned)((B + TOK - 1) / TOK));
    mmq_simple_kernel<TYPE, TOK><<<grid, MMVQ_WARPS * 32, 0, st>>>(w, xq, xds, dst, ldd, B, addend, lda);
    B200_CHECK(cudaGetLastError());
}

void mul_mat_q_simple(const QWeight &w, const int8_t *xq, const float2 *xds, float *dst, int64_t ldd, int64_t B, const float *addend, int64_t lda, cudaStream_t st) {
    if (w.N == 0 || B == 0) return;
    switch (w.type) {
        case T_Q4_0: launch_simple<T_Q4_0>(w, xq, xds, dst, ldd, B, addend, lda, st); break;
        case T_Q4_1: launch_simple<T_Q4_1>(w, xq, xds, dst, ldd, B, addend, lda, st); break;
        case T_Q5_0: launch_simple<T_Q5_0>(w, xq, xds, dst, ldd, B, addend, lda, st); break;
        case T_Q5_1: launch_simple<T_Q5_1>(w, xq, xds, dst, ldd, B, addend, lda, st); break;
        case T_Q8_0: launch_simple<T_Q8_0>(w, xq, xds, dst, ldd, B, addend, lda, st); break;
        default: B200_ASSERT(!"mul_mat_q_simple: unsupported weight type");
    }
}

}  // namespace b200
