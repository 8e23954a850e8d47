// llm_b200/csrc/mmq.cu -- prefill path: ggml_mul_mat for a batch of activation rows, integer-exact.
//
// Contract (LC/ggml.c:10397-10586 + vec_dots): dst[m][n] = sum_b (d_w[n,b] * d_x[m,b]) * S[m,n,b]  (+ m_w[n,b] * s_x[m,b]),
// S = exact int dot of one 32-element quant block.  Because the per-block scale is rank-1 in (m, n), the K loop cannot stay
// inside the tensor core: each k=32 MMA (exactly one quant block, s8 x s8 -> s32) is followed by a CUDA-core
// convert + scale + f32-fma into the register accumulators.  That epilogue, not the tensor pipe, bounds this kernel
// (3 fp32-pipe ops per (m, n, block); DESIGN.md "prefill roofline").
//
// v1 (this file): warp-level mma.sync.m16n8k32.s8 with a 3-stage cp.async pipeline.  GGML's nibble order (byte j = elements
// j and j+16, LC/ggml.c:1535-1540) is exactly the B-fragment order of m16n8k32 (k = 4t..4t+3 and 16+4t..16+4t+3 per thread), so
// a packed weight block is unpacked with one 32-bit shared load + two mask/shift per fragment; activations arrive as int8
// rows and are read with ldmatrix.
#include "kernels.cuh"

namespace b200 {

namespace {

constexpr int BM = 128, BN = 128, KB = 4, STAGES = 3, NTHREADS = 256;
constexpr int WM = 64, WN = 32;                 // warp tile: 2 warps along M x 4 along N
constexpr int A_STRIDE = KB * 32 + 16;          // bytes per activation row in smem (pad -> conflict-free ldmatrix)
constexpr int AS_STRIDE = KB * 8 + 8;           // float2 {d, aux} per block

template <int TYPE> struct Tr {
    static constexpr int QS = (TYPE == T_Q8_0) ? 32 : 16;
    static constexpr int DM = (TYPE == T_Q4_1 || TYPE == T_Q5_1) ? 4 : 2;
    static constexpr bool QH = (TYPE == T_Q5_0 || TYPE == T_Q5_1);
    static constexpr bool MIN = (TYPE == T_Q4_1 || TYPE == T_Q5_1);
    static constexpr int BQ_STRIDE = KB * QS + 16;
    static constexpr int BD_STRIDE = KB * DM + 4;
    static constexpr int BH_STRIDE = KB * 4 + 4;
    static constexpr int A_BYTES = BM * A_STRIDE, AS_BYTES = BM * AS_STRIDE;
    static constexpr int BQ_BYTES = BN * BQ_STRIDE, BD_BYTES = BN * BD_STRIDE, BH_BYTES = QH ? BN * BH_STRIDE : 0;
    static constexpr int STAGE_BYTES = A_BYTES + AS_BYTES + BQ_BYTES + BD_BYTES + BH_BYTES;
};

__device__ __forceinline__ void cp_async(void *smem, const void *gmem, int bytes_total, int src_bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    if (bytes_total == 16)     asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
    else if (bytes_total == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
    else                       asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void *smem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void mma_s8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};"
                 : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(0));
}

// per-byte: 4-bit n in [0,15] -> int8 (n - 8);  5-bit q in [0,31] -> int8 (q - 16).  No cross-byte carries.
__device__ __forceinline__ uint32_t sub8_nibbles(uint32_t w) { const uint32_t u = w ^ 0x08080808u; return u | ((u & 0x08080808u) * 30u); }
__device__ __forceinline__ uint32_t sub16_q5(uint32_t w)     { const uint32_t u = w ^ 0x10101010u; return u | ((u & 0x10101010u) * 14u); }

template <int TYPE>
__global__ void __launch_bounds__(NTHREADS, 2) mmq_kernel(const QWeight w, const int8_t *__restrict__ xq, const float2 *__restrict__ xds,
                                                          float *__restrict__ dst, int64_t ldd, int64_t B,
                                                          const float *__restrict__ addend, int64_t lda) {
    using T = Tr<TYPE>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp >> 2) * WM, wn0 = (warp & 3) * WN;
    const int64_t m_base = (int64_t)blockIdx.y * BM, n_base = (int64_t)blockIdx.x * BN;
    const int nb = (int)w.nb;
    const int ktiles = (nb + KB - 1) / KB;

    auto stage_ptr = [&](int s) { return smem + (size_t)s * T::STAGE_BYTES; };

    // ---- producer: one k-tile (KB blocks) of activations + weights into stage s.  Out-of-range rows are clamped (their
    //      results are never stored); blocks past nb are zero-filled (zero quants, zero scales contribute nothing). ----
    auto load_stage = [&](int s, int kt) {
        uint8_t *sA = stage_ptr(s), *sAS = sA + T::A_BYTES, *sBQ = sAS + T::AS_BYTES, *sBD = sBQ + T::BQ_BYTES, *sBH = sBD + T::BD_BYTES;
        const int b0 = kt * KB;
        // activations: BM rows x (KB*32) bytes as 16 B chunks (2 chunks per block)
        for (int c = tid; c < BM * KB * 2; c += NTHREADS) {
            const int r = c / (KB * 2), cc = c % (KB * 2), b = b0 + cc / 2;
            const int64_t m = m_base + r < B ? m_base + r : B - 1;
            const bool ok = b < nb;
            cp_async(sA + r * A_STRIDE + cc * 16, xq + (m * nb + (ok ? b : 0)) * QK + (cc & 1) * 16, 16, ok ? 16 : 0);
        }
        // activation scales: 8 B per block
        for (int c = tid; c < BM * KB; c += NTHREADS) {
            const int r = c / KB, b = b0 + c % KB;
            const int64_t m = m_base + r < B ? m_base + r : B - 1;
            const bool ok = b < nb;
            cp_async(sAS + r * AS_STRIDE + (c % KB) * 8, xds + m * nb + (ok ? b : 0), 8, ok ? 8 : 0);
        }
        // weight quants: 16 B chunks
        constexpr int QCH = T::QS / 16;
        for (int c = tid; c < BN * KB * QCH; c += NTHREADS) {
            const int r = c / (KB * QCH), cc = c % (KB * QCH), b = b0 + cc / QCH;
            const int64_t n = n_base + r < w.N ? n_base + r : w.N - 1;
            const bool ok = b < nb;
            cp_async(sBQ + r * T::BQ_STRIDE + cc * 16, w.qs + (n * nb + (ok ? b : 0)) * T::QS + (cc % QCH) * 16, 16, ok ? 16 : 0);
        }
        // weight scales (and mins): 2 or 4 B per block -> 4 B copies
        if (T::DM == 2) {
            for (int c = tid; c < BN * KB / 2; c += NTHREADS) {
                const int r = c / (KB / 2), b = b0 + (c % (KB / 2)) * 2;
                const int64_t n = n_base + r < w.N ? n_base + r : w.N - 1;
                const bool ok = b < nb;   // nb is even (Q4 rows need K % 64 == 0, crates/ggml/src/lib.rs:112-118)
                cp_async(sBD + r * T::BD_STRIDE + (c % (KB / 2)) * 4, (const uint8_t *)w.dm + (n * nb + (ok ? b : 0)) * 2, 4, ok ? 4 : 0);
            }
        } else {
            for (int c = tid; c < BN * KB; c += NTHREADS) {
                const int r = c / KB, b = b0 + c % KB;
                const int64_t n = n_base + r < w.N ? n_base + r : w.N - 1;
                const bool ok = b < nb;
                cp_async(sBD + r * T::BD_STRIDE + (c % KB) * 4, (const uint8_t *)w.dm + (n * nb + (ok ? b : 0)) * 4, 4, ok ? 4 : 0);
            }
        }
        if (T::QH) {
            for (int c = tid; c < BN * KB; c += NTHREADS) {
                const int r = c / KB, b = b0 + c % KB;
                const int64_t n = n_base + r < w.N ? n_base + r : w.N - 1;
                const bool ok = b < nb;
                cp_async(sBH + r * T::BH_STRIDE + (c % KB) * 4, w.qh + n * nb + (ok ? b : 0), 4, ok ? 4 : 0);
            }
        }
    };

    float acc[WM / 16][WN / 8][4];
#pragma unroll
    for (int i = 0; i < WM / 16; i++)
#pragma unroll
        for (int j = 0; j < WN / 8; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[i][j][e] = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) { if (s < ktiles) load_stage(s, s); cp_async_commit(); }

    for (int kt = 0; kt < ktiles; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        { const int nk = kt + STAGES - 1; if (nk < ktiles) load_stage(nk % STAGES, nk); cp_async_commit(); }

        const uint8_t *sA = stage_ptr(kt % STAGES), *sAS = sA + T::A_BYTES, *sBQ = sAS + T::AS_BYTES, *sBD = sBQ + T::BQ_BYTES, *sBH = sBD + T::BD_BYTES;
#pragma unroll
        for (int b = 0; b < KB; b++) {
            // B fragments + weight scales for the 4 n-tiles of this warp
            uint32_t bf[WN / 8][2];
            float dw[WN / 8][2], mw[WN / 8][2];
#pragma unroll
            for (int j = 0; j < WN / 8; j++) {
                const int rn = wn0 + j * 8 + g;
                if (TYPE == T_Q8_0) {
                    bf[j][0] = *(const uint32_t *)(sBQ + rn * T::BQ_STRIDE + b * 32 + 4 * t);
                    bf[j][1] = *(const uint32_t *)(sBQ + rn * T::BQ_STRIDE + b * 32 + 16 + 4 * t);
                } else {
                    const uint32_t wq = *(const uint32_t *)(sBQ + rn * T::BQ_STRIDE + b * 16 + 4 * t);
                    uint32_t lo = wq & 0x0F0F0F0Fu, hi = (wq >> 4) & 0x0F0F0F0Fu;
                    if (T::QH) {
                        const uint32_t qh = *(const uint32_t *)(sBH + rn * T::BH_STRIDE + b * 4);
                        lo |= spread4_to_bit4(qh >> (4 * t));
                        hi |= spread4_to_bit4(qh >> (16 + 4 * t));
                    }
                    if (TYPE == T_Q4_0) { lo = sub8_nibbles(lo); hi = sub8_nibbles(hi); }
                    if (TYPE == T_Q5_0) { lo = sub16_q5(lo); hi = sub16_q5(hi); }
                    bf[j][0] = lo; bf[j][1] = hi;
                }
                // C-fragment columns of this thread: n = wn0 + j*8 + 2t, +1
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const int cn = wn0 + j * 8 + 2 * t + c;
                    if (T::MIN) {
                        const __half2 dm = *(const __half2 *)(sBD + cn * T::BD_STRIDE + b * 4);
                        dw[j][c] = __low2float(dm); mw[j][c] = __high2float(dm);
                    } else {
                        dw[j][c] = __half2float(*(const __half *)(sBD + cn * T::BD_STRIDE + b * 2)); mw[j][c] = 0.f;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < WM / 16; i++) {
                uint32_t af[4];
                const int rm = wm0 + i * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldmatrix_x4(af, sA + rm * A_STRIDE + b * 32 + (lane >> 4) * 16);
                const float2 x0 = *(const float2 *)(sAS + (wm0 + i * 16 + g) * AS_STRIDE + b * 8);        // row g
                const float2 x1 = *(const float2 *)(sAS + (wm0 + i * 16 + g + 8) * AS_STRIDE + b * 8);    // row g + 8
#pragma unroll
                for (int j = 0; j < WN / 8; j++) {
                    int c[4];
                    mma_s8(c, af, bf[j][0], bf[j][1]);
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float2 xs = (e < 2) ? x0 : x1;
                        const float d = __fmul_rn(dw[j][e & 1], xs.x);
                        float a = __fmaf_rn(d, (float)c[e], acc[i][j][e]);
                        if (T::MIN) a = __fmaf_rn(mw[j][e & 1], xs.y, a);
                        acc[i][j][e] = a;
                    }
                }
            }
        }
    }
    cp_async_wait<0>();

    // ---- store ----
#pragma unroll
    for (int i = 0; i < WM / 16; i++)
#pragma unroll
        for (int j = 0; j < WN / 8; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int64_t m = m_base + wm0 + i * 16 + g + (e >> 1) * 8;
                const int64_t n = n_base + wn0 + j * 8 + 2 * t + (e & 1);
                if (m < B && n < w.N) {
                    float v = acc[i][j][e];
                    if (addend) v = __fadd_rn(v, addend[m * lda + n]);
                    dst[m * ldd + n] = v;
                }
            }
}

template <int TYPE>
void launch_mmq(const QWeight &w, const int8_t *xq, const float2 *xds, float *dst, int64_t ldd, int64_t B, const float *addend, int64_t lda, cudaStream_t st) {
    using T = Tr<TYPE>;
    constexpr int smem = STAGES * T::STAGE_BYTES;
    static bool attr_set = false;
    if (!attr_set) {
        B200_CHECK(cudaFuncSetAttribute(mmq_kernel<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    dim3 grid((unsigned)((w.N + BN - 1) / BN), (unsigned)((B + BM - 1) / BM));
    mmq_kernel<TYPE><<<grid, NTHREADS, smem, st>>>(w, xq, xds, dst, ldd, B, addend, lda);
    B200_CHECK(cudaGetLastError());
}

}  // namespace

void mul_mat_q(const QWeight &w, const int8_t *xq, const float2 *xds, float *dst, int64_t ldd, int64_t B, const float *addend, int64_t lda, cudaStream_t st) {
    if (w.N == 0 || B == 0) return;
    B200_ASSERT(w.nb % 2 == 0);
    switch (w.type) {
        case T_Q4_0: launch_mmq<T_Q4_0>(w, xq, xds, dst, ldd, B, addend, lda, st); break;
        case T_Q4_1: launch_mmq<T_Q4_1>(w, xq, xds, dst, ldd, B, addend, lda, st); break;
        case T_Q5_0: launch_mmq<T_Q5_0>(w, xq, xds, dst, ldd, B, addend, lda, st); break;
        case T_Q5_1: launch_mmq<T_Q5_1>(w, xq, xds, dst, ldd, B, addend, lda, st); break;
        case T_Q8_0: launch_mmq<T_Q8_0>(w, xq, xds, dst, ldd, B, addend, lda, st); break;
        default: B200_ASSERT(!"mul_mat_q: unsupported weight type");
    }
}

}  // namespace b200
