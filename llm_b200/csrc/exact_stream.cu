// llm_b200/csrc/exact_stream.cu -- the decode mat-vec at HBM speed WITH the reference's bit-exact operation order.
//
// Arithmetic = exact.cu (AVX2 lane chains of ggml_vec_dot_q*_q8_*).  Data movement is what changes:
//   * a dedicated producer warp streams the weight rows of a 32-row tile chunk by chunk (32 quant blocks = 512 B of nibbles per
//     row) into a 4-stage shared-memory ring with TMA bulk copies (cp.async.bulk.shared::cluster.global + mbarrier complete_tx):
//     ~55 KB in flight per CTA without spending a single register or LSU slot of the compute warps on global loads;
//   * 4 compute warps (4 threads per row: thread w owns AVX lanes w and w+4 == packed word w of every block) walk the blocks
//     IN ORDER out of shared memory: 1 LDS.32 (nibbles) + 1 LDS.U16 (d) + 1 LDS.128 (activation pack) + 2 dp4a + 2 fma per block;
//   * the quantized activation row is re-packed once per mat-vec (quantize_act_pack) into 16-byte records per (block, word):
//     {x word w, x word w+4, -offset * (byte sums of both words) as 2 x int16 | s, d_x}, so the -8 / -16 nibble offsets ride in
//     the dp4a accumulator and need no per-block unpack arithmetic.
// Algorithmic bytes per row of K weights: K/32 * {18, 20, 22, 24, 34}; each is read exactly once.
#include "kernels.cuh"

namespace b200 {

namespace {

constexpr int SR = 32;        // rows per CTA tile
constexpr int SCB = 32;       // quant blocks per chunk
constexpr int SST = 4;        // ring stages
constexpr int SCOMPUTE = 128; // 4 compute warps
constexpr int STHREADS = SCOMPUTE + 32;

template <int TYPE> struct St {
    static constexpr int QS = (TYPE == T_Q8_0) ? 32 : 16;
    static constexpr int DM = (TYPE == T_Q4_1 || TYPE == T_Q5_1) ? 4 : 2;
    static constexpr bool QH = (TYPE == T_Q5_0 || TYPE == T_Q5_1);
    static constexpr bool MIN = (TYPE == T_Q4_1 || TYPE == T_Q5_1);
    static constexpr int QS_STRIDE = SCB * QS + 16;     // +16 B: rows land in different banks (stride % 128 B == 16)
    static constexpr int DM_STRIDE = SCB * DM + 16;
    static constexpr int QH_STRIDE = SCB * 4 + 16;
    static constexpr int QS_BYTES = SR * QS_STRIDE, DM_BYTES = SR * DM_STRIDE, QH_BYTES = QH ? SR * QH_STRIDE : 0;
    static constexpr int STAGE_BYTES = QS_BYTES + DM_BYTES + QH_BYTES;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar)) : "memory");
}

template <int TYPE>
__global__ void __launch_bounds__(STHREADS) mmv_exact_stream_kernel(const QWeight w, const int4 *__restrict__ xpack, float *__restrict__ dst,
                                                                    const float *__restrict__ addend) {
    using T = St<TYPE>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = (uint64_t *)smem, *empty = full + SST;
    uint8_t *ring = smem + 128;
    const int nb = (int)w.nb;
    int4 *sx = (int4 *)(ring + SST * T::STAGE_BYTES);     // [nb][4] activation records
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row_base = (int64_t)blockIdx.x * SR;
    const int nchunks = (nb + SCB - 1) / SCB;

    if (tid == 0) {
        for (int s = 0; s < SST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], SCOMPUTE / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == SCOMPUTE / 32) {
        // ===== producer warp: lane r streams row r of the tile =====
        const int64_t row = row_base + lane < w.N ? row_base + lane : w.N - 1;      // tail tile: re-read a valid row, byte counts stay fixed
        for (int c = 0; c < nchunks; c++) {
            const int s = c % SST;
            const int b0 = c * SCB, cb = nb - b0 < SCB ? nb - b0 : SCB;
            mbar_wait(&empty[s], ((c / SST) & 1) ^ 1);
            if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)(SR * cb * (T::QS + T::DM + (T::QH ? 4 : 0))));
            __syncwarp();
            uint8_t *st = ring + s * T::STAGE_BYTES;
            bulk_g2s(st + lane * T::QS_STRIDE, w.qs + (row * nb + b0) * T::QS, cb * T::QS, &full[s]);
            bulk_g2s(st + T::QS_BYTES + lane * T::DM_STRIDE, (const uint8_t *)w.dm + (row * nb + b0) * T::DM, cb * T::DM, &full[s]);
            if (T::QH) bulk_g2s(st + T::QS_BYTES + T::DM_BYTES + lane * T::QH_STRIDE, w.qh + row * nb + b0, cb * 4, &full[s]);
        }
        return;
    }

    // ===== compute warps =====
    for (int i = tid; i < nb * 4; i += SCOMPUTE) sx[i] = __ldg(xpack + i);           // activation records -> shared memory
    asm volatile("bar.sync 1, %0;" ::"n"(SCOMPUTE));                                 // compute warps only
    const int r = tid >> 2, wd = tid & 3;
    float a_lo = 0.f, a_hi = 0.f, summs = 0.f;
    for (int c = 0; c < nchunks; c++) {
        const int s = c % SST;
        const int b0 = c * SCB, cb = nb - b0 < SCB ? nb - b0 : SCB;
        mbar_wait(&full[s], (c / SST) & 1);
        const uint8_t *st = ring + s * T::STAGE_BYTES;
        const uint8_t *qrow = st + r * T::QS_STRIDE, *drow = st + T::QS_BYTES + r * T::DM_STRIDE, *hrow = st + T::QS_BYTES + T::DM_BYTES + r * T::QH_STRIDE;
#pragma unroll 4
        for (int b = 0; b < cb; b++) {
            const int4 xp = sx[(b0 + b) * 4 + wd];
            int lo, hi;
            if (TYPE == T_Q8_0) {
                lo = *(const int *)(qrow + b * 32 + 4 * wd);
                hi = *(const int *)(qrow + b * 32 + 16 + 4 * wd);
            } else {
                const uint32_t q = *(const uint32_t *)(qrow + b * 16 + 4 * wd);
                uint32_t l = q & 0x0F0F0F0Fu, h = (q >> 4) & 0x0F0F0F0Fu;
                if (T::QH) {
                    const uint32_t qh = *(const uint32_t *)(hrow + b * 4);
                    l |= spread4_to_bit4(qh >> (4 * wd));
                    h |= spread4_to_bit4(qh >> (16 + 4 * wd));
                }
                lo = (int)l; hi = (int)h;
            }
            float dw, mw = 0.f;
            if (T::MIN) { const __half2 dm = *(const __half2 *)(drow + b * 4); dw = __low2float(dm); mw = __high2float(dm); }
            else dw = __half2float(*(const __half *)(drow + b * 2));
            // Q4_0 / Q5_0: the -8 / -16 offset of every nibble is pre-multiplied into the accumulator seed (exact integers)
            const int seed_lo = T::MIN || TYPE == T_Q8_0 ? 0 : (int)(short)(xp.z & 0xffff);
            const int seed_hi = T::MIN || TYPE == T_Q8_0 ? 0 : (xp.z >> 16);
            const float d = __fmul_rn(dw, __int_as_float(xp.w));
            a_lo = __fmaf_rn(d, (float)__dp4a(lo, xp.x, seed_lo), a_lo);
            a_hi = __fmaf_rn(d, (float)__dp4a(hi, xp.y, seed_hi), a_hi);
            if (T::MIN) summs = __fmaf_rn(mw, __int_as_float(xp.z), summs);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    float v = __fadd_rn(a_hi, a_lo);                                   // hsum_float_8 (see exact.cu)
    v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 2));
    v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 1));
    if (T::MIN) v = __fadd_rn(v, summs);
    const int64_t row = row_base + r;
    if (wd == 0 && row < w.N) dst[row] = addend ? __fadd_rn(v, addend[row]) : v;
}

// ---- activation quantizer that emits the packed records (one warp per block; arithmetic identical to quantize_act) ----------------
// record (block b, word w) = { x bytes 4w..4w+3, x bytes 16+4w..16+4w+3, z, d_x }:
//   z = 2 x int16 { -off * sum(bytes of word w), -off * sum(bytes of word w+4) }   for Q8_0 activations (off = 8: Q4_0, 16: Q5_0, 0: Q8_0)
//   z = bits of s = d * sum(q)                                                      for Q8_1 activations (Q4_1 / Q5_1)
__global__ void __launch_bounds__(256) quantize_act_pack_kernel(const float *__restrict__ x, int4 *__restrict__ pack, int nbk, int q81, int off) {
    const int blk = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (blk >= nbk) return;
    const int lane = threadIdx.x & 31;
    const float v = x[blk * QK + lane];
    const float amax = warp_max(fabsf(v));
    const float d = __fdiv_rn(amax, 127.f);
    const float id = (amax != 0.0f) ? __fdiv_rn(127.f, amax) : 0.0f;
    const int q = __float2int_rn(__fmul_rn(v, id));
    const int isum = warp_sum(q);
    uint32_t word = (uint32_t)(q & 0xff) << (8 * (lane & 3));
    word |= __shfl_xor_sync(0xffffffffu, word, 1);
    word |= __shfl_xor_sync(0xffffffffu, word, 2);
    int s4 = q + __shfl_xor_sync(0xffffffffu, q, 1);
    s4 += __shfl_xor_sync(0xffffffffu, s4, 2);
    const uint32_t word_hi = __shfl_down_sync(0xffffffffu, word, 16);
    const int s4_hi = __shfl_down_sync(0xffffffffu, s4, 16);
    if (lane < 16 && (lane & 3) == 0) {
        const int wd = lane >> 2;
        int4 rec;
        rec.x = (int)word; rec.y = (int)word_hi;
        if (q81) { rec.z = __float_as_int(__fmul_rn(d, (float)isum)); rec.w = __float_as_int(d); }
        else { rec.z = (int)(((uint32_t)(-off * s4) & 0xffffu) | ((uint32_t)(-off * s4_hi) << 16)); rec.w = __float_as_int(__half2float(__float2half_rn(d))); }
        pack[blk * 4 + wd] = rec;
    }
}

template <int TYPE>
void launch_stream(const QWeight &w, const int4 *xpack, float *dst, const float *addend, cudaStream_t st) {
    using T = St<TYPE>;
    const int smem = 128 + SST * T::STAGE_BYTES + (int)w.nb * 64;
    static int smem_set = 0;
    if (smem > smem_set) {
        B200_ASSERT(smem <= 227 * 1024);
        B200_CHECK(cudaFuncSetAttribute(mmv_exact_stream_kernel<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        smem_set = smem;
    }
    mmv_exact_stream_kernel<TYPE><<<(unsigned)((w.N + SR - 1) / SR), STHREADS, smem, st>>>(w, xpack, dst, addend);
    B200_CHECK(cudaGetLastError());
}

}  // namespace

bool mmv_exact_stream_supported(const QWeight &w) { return w.nb % 8 == 0 && w.nb * 64 + 128 + SST * St<T_Q8_0>::STAGE_BYTES <= 227 * 1024; }

void quantize_act_pack(int wtype, const float *x, int4 *pack, int64_t K, cudaStream_t st) {
    const int nbk = (int)(K / QK);
    const int off = wtype == T_Q4_0 ? 8 : (wtype == T_Q5_0 ? 16 : 0);
    quantize_act_pack_kernel<<<(nbk + 7) / 8, 256, 0, st>>>(x, pack, nbk, has_min(wtype) ? 1 : 0, off);
    B200_CHECK(cudaGetLastError());
}

void mul_mat_vec_q_exact_stream(const QWeight &w, const int4 *xpack, float *dst, const float *addend, cudaStream_t st) {
    if (w.N == 0) return;
    B200_ASSERT(mmv_exact_stream_supported(w));
    switch (w.type) {
        case T_Q4_0: launch_stream<T_Q4_0>(w, xpack, dst, addend, st); break;
        case T_Q4_1: launch_stream<T_Q4_1>(w, xpack, dst, addend, st); break;
        case T_Q5_0: launch_stream<T_Q5_0>(w, xpack, dst, addend, st); break;
        case T_Q5_1: launch_stream<T_Q5_1>(w, xpack, dst, addend, st); break;
        case T_Q8_0: launch_stream<T_Q8_0>(w, xpack, dst, addend, st); break;
        default: B200_ASSERT(!"mul_mat_vec_q_exact_stream: unsupported weight type");
    }
}

}  // namespace b200
