// llm_b200/csrc/decode.cu -- one decoded token = ONE persistent cooperative kernel (Llama::evaluate with a single token,
// crates/models/llama/src/lib.rs:144-368), bit-exact with the reference's CPU path.
//
// Why: at batch 1 the 7B graph is ~129 weight mat-vecs of 1.5-8 us each plus ~400 tiny row ops; launched one by one the GPU
// spends more time filling and draining pipelines than streaming weights (profiles/r01_notes.md: 2.5 TB/s inside the mat-vecs,
// 0.9 TB/s per token).  Here every CTA stays resident for the whole token:
//   * its producer warp walks the model's weight matrices in graph order and keeps the shared-memory ring full with cp.async;
//     weights do not depend on activations, so streaming continues while the compute warps wait at a phase boundary or run attention;
//   * its 4 compute warps execute the graph phase by phase, separated by grid-wide barriers (6 per layer):
//       A  pack(rms_norm(x) * attn_norm) -> [wq|wk|wv] mat-vec -> RoPE, K/V rows stored to the f16 cache     (llama lib.rs:183-244)
//       B  KQ = K . f16(Q) for every cached position                                                         (:246-265)
//       C  scale + soft_max (fp16 exp table) + KQV = V^T . f16(P), written in merged [n_embd] order          (:268-307)
//       D  pack(attn) -> wo mat-vec + residual                                                               (:310-314)
//       E  pack(rms_norm(inpFF) * ffn_norm) -> [w1|w3] mat-vec                                               (:318-325)
//       F  pack(silu(w1 x) * (w3 x)) -> w2 mat-vec + residual                                                (:328-334)
//     then final norm -> lm_head.  Every CTA re-derives the (tiny) quantized activation vector it needs on its own; only the
//     mat-vec outputs travel through global memory (L2).
// Arithmetic is exactly that of exact.cu / rowops.cu (AVX2 lane chains, ggml_vec_dot_f16 order, fp16 tables, host-built RoPE table).
#include "decode.h"
#include "stream_core.cuh"

namespace b200 {

using namespace stream;

namespace {

__device__ __forceinline__ float lutf(const uint16_t *t, float x) { return f16_bits_to_f32(__ldg(t + f32_to_f16_bits(x))); }

// all CTAs resident (cooperative launch); only the 128 compute threads of each CTA take part
__device__ __forceinline__ void grid_sync(unsigned int *bar, int tid) {
    compute_sync();
    if (tid == 0) {
        volatile unsigned int *genp = bar + 1;
        const unsigned int gen = *genp;
        __threadfence();
        if (atomicAdd(bar, 1u) == gridDim.x - 1) {
            bar[0] = 0;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*genp == gen) __nanosleep(32);
        }
        __threadfence();
    }
    compute_sync();
}

__device__ __forceinline__ float cta_rms_scale(const float *x, int n, float eps, double *shd, int tid) {
    double s = 0.0;
    for (int i = tid; i < n; i += SCOMPUTE) { const float v = __ldcg(x + i); s += (double)__fmul_rn(v, v); }
    s = warp_sum(s);
    if ((tid & 31) == 0) shd[tid >> 5] = s;
    compute_sync();
    const double tot = (shd[0] + shd[1]) + (shd[2] + shd[3]);
    compute_sync();
    const float mean = (float)(tot / (double)n);
    return __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, eps)));
}

template <int TYPE, class F>
__device__ __forceinline__ void build_pack(int4 *sx, int nbk, F val, int tid) {
    const int lane = tid & 31;
    for (int b = tid >> 5; b < nbk; b += SCOMPUTE / 32)
        pack_block(val(b * QK + lane), sx + b * 4, lane, has_min(TYPE) ? 1 : 0, TYPE == T_Q5_0 ? 16 : 0, TYPE == T_Q4_0 ? 1 : 0);
    compute_sync();
}

// element i (i < 16: low half, element i and i+16 of the block) of one quant block of a planes matrix (dequantize_row_*, LC/ggml.c:1525-1635)
template <int TYPE>
__device__ __forceinline__ void dequant_pair(const QWeight &w, int64_t blk, int j, float &lo, float &hi) {
    float d, m = 0.f;
    if (has_min(TYPE)) { const __half2 dm = ((const __half2 *)w.dm)[blk]; d = __low2float(dm); m = __high2float(dm); }
    else d = __half2float(((const __half *)w.dm)[blk]);
    if (TYPE == T_Q8_0) { const int8_t *q = (const int8_t *)w.qs + blk * 32; lo = (float)q[j] * d; hi = (float)q[j + 16] * d; return; }
    const int q = w.qs[blk * 16 + j];
    int q0 = q & 0xF, q1 = q >> 4;
    if (has_qh(TYPE)) { const uint32_t qh = w.qh[blk]; q0 |= ((qh >> j) & 1) << 4; q1 |= ((qh >> (j + 16)) & 1) << 4; }
    if (TYPE == T_Q4_0) { lo = (float)(q0 - 8) * d; hi = (float)(q1 - 8) * d; }
    else if (TYPE == T_Q5_0) { lo = (float)(q0 - 16) * d; hi = (float)(q1 - 16) * d; }
    else { lo = __fmaf_rn((float)q0, d, m); hi = __fmaf_rn((float)q1, d, m); }
}

// ggml_vec_dot_f16 reduction tree over a warp (see exact.cu)
__device__ __forceinline__ float f16dot_tree(float s) {
    s = __fadd_rn(s, __shfl_down_sync(0xffffffffu, s, 16));
    s = __fadd_rn(s, __shfl_down_sync(0xffffffffu, s, 8));
    s = __fadd_rn(s, __shfl_down_sync(0xffffffffu, s, 4));
    s = __fadd_rn(s, __shfl_down_sync(0xffffffffu, s, 1));
    s = __fadd_rn(s, __shfl_down_sync(0xffffffffu, s, 2));
    return s;
}

template <int TYPE>
__global__ void __launch_bounds__(STHREADS) llama_decode_kernel(const DecodeParams P) {
    using T = St<TYPE>;
    extern __shared__ __align__(128) uint8_t smem[];
    Ring R{(uint64_t *)smem, (uint64_t *)smem + SST, smem + 128, 0u};
    double *shd = (double *)(smem + 64);                       // 8 doubles of reduction scratch
    uint8_t *scratch = R.base + T::RING_BYTES;                 // activation records (mat-vec phases) | attention scratch (phases B, C)
    int4 *sx = (int4 *)scratch;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, ncta = gridDim.x;
    if (tid == 0) ring_init(R.full, R.empty);
    __syncthreads();

    if (tid >= SCOMPUTE) {
        // ===== producer warp: the token's weights, in graph order =====
        for (int il = 0; il < P.n_layer; il++) {
            const DecodeLayer *L = P.layers + il;
            { const QWeight w = L->wqkv; produce_matvec<TYPE>(w, R, cta, ncta, lane); }
            { const QWeight w = L->wo;   produce_matvec<TYPE>(w, R, cta, ncta, lane); }
            { const QWeight w = L->w13;  produce_matvec<TYPE>(w, R, cta, ncta, lane); }
            { const QWeight w = L->w2;   produce_matvec<TYPE>(w, R, cta, ncta, lane); }
        }
        produce_matvec<TYPE>(P.output, R, cta, ncta, lane);
        return;
    }

    // ===== compute warps =====
    const int e = P.e, f = P.f, hd = P.hd, gqa = P.gqa, n_ctx = P.n_ctx;
    const int p = __ldcg(P.n_past), n_kv = p + 1;
    const int tok = __ldcg(P.token);

    // embedding row: get_rows(tok_embeddings, token)                                             llama lib.rs:170
    for (int b = cta; b < e / QK; b += ncta)
        if (tid < 16) {
            float lo, hi;
            dequant_pair<TYPE>(P.wte, (int64_t)tok * (e / QK) + b, tid, lo, hi);
            P.x[b * QK + tid] = lo; P.x[b * QK + tid + 16] = hi;
        }
    grid_sync(P.bar, tid);

    for (int il = 0; il < P.n_layer; il++) {
        const DecodeLayer *L = P.layers + il;
        __half *Kl = L->K, *Vl = L->V;
        // ---- A: attention norm -> QKV mat-vec -> RoPE + KV store ----
        {
            const float scale = cta_rms_scale(P.x, e, P.eps, shd, tid);
            const float *gain = L->attn_norm;
            build_pack<TYPE>(sx, e / QK, [&](int i) { return __fmul_rn(__fmul_rn(__ldcg(P.x + i), scale), __ldg(gain + i)); }, tid);
            const QWeight w = L->wqkv;
            consume_matvec<TYPE>(w, sx, R, cta, ncta, tid, [&](int64_t row, float v) {
                const float other = __shfl_xor_sync(0xffffffffu, v, 4);          // the rotation partner: rows 2i and 2i+1 sit in adjacent quads
                if ((tid & 3) != 0 || row >= w.N) return;
                if (row < e + gqa) {                                               // Q or K row: ggml_rope mode 0 (LC/ggml.c:11859-11874)
                    const int within = (int)(row < e ? row : row - e);
                    const float2 cs = __ldg(P.rope_cs + (int64_t)p * P.rope_half + (within % hd) / 2);
                    const bool even = (row & 1) == 0;
                    const float x0 = even ? v : other, x1 = even ? other : v;
                    const float out = even ? __fmaf_rn(x0, cs.x, -__fmul_rn(x1, cs.y)) : __fmaf_rn(x0, cs.y, __fmul_rn(x1, cs.x));
                    if (row < e) P.q[row] = out;
                    else Kl[(int64_t)p * gqa + within] = __float2half_rn(out);    // cpy f32 -> f16 into memory_k          (:243)
                } else {
                    Vl[(int64_t)(row - e - gqa) * n_ctx + p] = __float2half_rn(v); // transposed V store                    (:244)
                }
            });
        }
        grid_sync(P.bar, tid);
        // ---- B: KQ[h][j] = K[j][h] . f16(Q[h])  (ggml_vec_dot_f16 order), 64 cached positions per work item ----
        {
            __half *q16 = (__half *)scratch;
            const int chunks = (n_kv + 63) / 64, items = P.n_head * chunks;
            for (int it = cta; it < items; it += ncta) {
                const int h = it / chunks, j0 = (it - h * chunks) * 64;
                const int hk = h / (P.n_head / P.n_head_kv);
                for (int i = tid; i < hd; i += SCOMPUTE) q16[i] = __float2half_rn(__ldcg(P.q + h * hd + i));
                compute_sync();
                const int np = hd & ~31;
                for (int jj = 0; jj < 16; jj++) {
                    const int j = j0 + warp * 16 + jj;
                    if (j >= n_kv) break;
                    const __half *krow = Kl + (int64_t)j * gqa + hk * hd;
                    float s = 0.f;
                    for (int k = lane; k < np; k += 32) s = __fmaf_rn(__half2float(__ldcg(krow + k)), __half2float(q16[k]), s);
                    s = f16dot_tree(s);
                    if (lane == 0) {
                        double sumf = (double)s;
                        for (int k = np; k < hd; k++) sumf += (double)__fmul_rn(__half2float(__ldcg(krow + k)), __half2float(q16[k]));
                        P.kq[(int64_t)h * n_ctx + j] = (float)sumf;
                    }
                }
                compute_sync();
            }
        }
        grid_sync(P.bar, tid);
        // ---- C: scale, soft_max, KQV for 32 channels of one head per work item ----
        {
            float *sc = (float *)scratch;                       // [n_kv] exp values
            __half *p16 = (__half *)(scratch + (size_t)n_ctx * 4);
            float *shf = (float *)shd;
            const int cpi = 32, per_head = hd / cpi, items = P.n_head * per_head;
            for (int it = cta; it < items; it += ncta) {
                const int h = it / per_head, c0 = (it - h * per_head) * cpi;
                const int hk = h / (P.n_head / P.n_head_kv);
                float mx = -INFINITY;
                for (int j = tid; j < n_kv; j += SCOMPUTE) { const float v = __fmul_rn(__ldcg(P.kq + (int64_t)h * n_ctx + j), P.kq_scale); sc[j] = v; mx = fmaxf(mx, v); }
                mx = warp_max(mx);
                if (lane == 0) shf[warp] = mx;
                compute_sync();
                mx = fmaxf(fmaxf(shf[0], shf[1]), fmaxf(shf[2], shf[3]));
                compute_sync();
                double s = 0.0;
                for (int j = tid; j < n_kv; j += SCOMPUTE) { const float ev = lutf(P.lut_exp, __fsub_rn(sc[j], mx)); sc[j] = ev; s += (double)ev; }
                s = warp_sum(s);
                if (lane == 0) shd[4 + warp] = s;
                compute_sync();
                const float inv = (float)(1.0 / ((shd[4] + shd[5]) + (shd[6] + shd[7])));
                for (int j = tid; j < n_kv; j += SCOMPUTE) p16[j] = __float2half_rn(__fmul_rn(sc[j], inv));
                compute_sync();
                const int np = n_kv & ~31;
                for (int cc = 0; cc < cpi / 4; cc++) {
                    const int c = c0 + warp * (cpi / 4) + cc;
                    const __half *vrow = Vl + (int64_t)(hk * hd + c) * n_ctx;
                    float a = 0.f;
                    for (int k = lane; k < np; k += 32) a = __fmaf_rn(__half2float(__ldcg(vrow + k)), __half2float(p16[k]), a);
                    a = f16dot_tree(a);
                    if (lane == 0) {
                        double sumf = (double)a;
                        for (int k = np; k < n_kv; k++) sumf += (double)__fmul_rn(__half2float(__ldcg(vrow + k)), __half2float(p16[k]));
                        P.attn[h * hd + c] = (float)sumf;
                    }
                }
                compute_sync();
            }
        }
        grid_sync(P.bar, tid);
        // ---- D: wo + residual ----
        {
            build_pack<TYPE>(sx, e / QK, [&](int i) { return __ldcg(P.attn + i); }, tid);
            const QWeight w = L->wo;
            consume_matvec<TYPE>(w, sx, R, cta, ncta, tid, [&](int64_t row, float v) {
                if ((tid & 3) == 0 && row < w.N) P.ff[row] = __fadd_rn(v, __ldcg(P.x + row));
            });
        }
        grid_sync(P.bar, tid);
        // ---- E: ffn norm -> [w1|w3] ----
        {
            const float scale = cta_rms_scale(P.ff, e, P.eps, shd, tid);
            const float *gain = L->ffn_norm;
            build_pack<TYPE>(sx, e / QK, [&](int i) { return __fmul_rn(__fmul_rn(__ldcg(P.ff + i), scale), __ldg(gain + i)); }, tid);
            const QWeight w = L->w13;
            consume_matvec<TYPE>(w, sx, R, cta, ncta, tid, [&](int64_t row, float v) {
                if ((tid & 3) == 0 && row < w.N) P.h13[row] = v;
            });
        }
        grid_sync(P.bar, tid);
        // ---- F: silu(w1 x) * (w3 x) -> w2 + residual ----
        {
            build_pack<TYPE>(sx, f / QK, [&](int i) { return __fmul_rn(lutf(P.lut_silu, __ldcg(P.h13 + i)), __ldcg(P.h13 + f + i)); }, tid);
            const QWeight w = L->w2;
            consume_matvec<TYPE>(w, sx, R, cta, ncta, tid, [&](int64_t row, float v) {
                if ((tid & 3) == 0 && row < w.N) P.x[row] = __fadd_rn(v, __ldcg(P.ff + row));
            });
        }
        grid_sync(P.bar, tid);
    }
    // ---- final norm -> lm_head ----
    {
        const float scale = cta_rms_scale(P.x, e, P.eps, shd, tid);
        build_pack<TYPE>(sx, e / QK, [&](int i) { return __fmul_rn(__fmul_rn(__ldcg(P.x + i), scale), __ldg(P.norm + i)); }, tid);
        consume_matvec<TYPE>(P.output, sx, R, cta, ncta, tid, [&](int64_t row, float v) {
            if ((tid & 3) == 0 && row < P.output.N) P.logits[row] = v;
        });
    }
    if (cta == 0 && tid == 0) *P.n_past = p + 1;               // InferenceSession::n_past += 1 (inference_session.rs:288)
}

template <int TYPE>
int decode_smem_bytes(const DecodeParams &P) {
    const size_t sxb = (size_t)((P.f > P.e ? P.f : P.e) / QK) * 64;
    const size_t att = (size_t)P.n_ctx * 6 + 512;
    return (int)(128 + St<TYPE>::RING_BYTES + (sxb > att ? sxb : att));
}

template <int TYPE>
bool launch_decode_t(const DecodeParams &P, cudaStream_t st, int *grid_out) {
    const int smem = decode_smem_bytes<TYPE>(P);
    static int smem_set = 0, grid = 0;
    if (smem > 227 * 1024) return false;
    if (smem != smem_set) {
        B200_CHECK(cudaFuncSetAttribute(llama_decode_kernel<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int dev, sms, per_sm;
        B200_CHECK(cudaGetDevice(&dev));
        B200_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        B200_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, llama_decode_kernel<TYPE>, STHREADS, smem));
        if (per_sm < 1) return false;
        grid = sms * per_sm;                                   // every CTA resident: the grid barrier relies on it (cooperative launch checks)
        smem_set = smem;
    }
    void *args[] = {(void *)&P};
    B200_CHECK(cudaLaunchCooperativeKernel((const void *)llama_decode_kernel<TYPE>, dim3(grid), dim3(STHREADS), args, smem, st));
    if (grid_out) *grid_out = grid;
    return true;
}

}  // namespace

bool decode_supported(const DecodeParams &P, int wtype) {
    QWeight probe; probe.nb = P.e / QK;
    QWeight probe2; probe2.nb = P.f / QK;
    return is_quant(wtype) && mmv_exact_stream_supported(probe) && mmv_exact_stream_supported(probe2) && P.hd % 2 == 0 && P.hd <= 256 && P.e % 64 == 0 &&
           P.gqa % SR == 0 && P.e % SR == 0;
}

bool launch_decode(const DecodeParams &P, int wtype, cudaStream_t st, int *grid_out) {
    switch (wtype) {
        case T_Q4_0: return launch_decode_t<T_Q4_0>(P, st, grid_out);
        case T_Q4_1: return launch_decode_t<T_Q4_1>(P, st, grid_out);
        case T_Q5_0: return launch_decode_t<T_Q5_0>(P, st, grid_out);
        case T_Q5_1: return launch_decode_t<T_Q5_1>(P, st, grid_out);
        case T_Q8_0: return launch_decode_t<T_Q8_0>(P, st, grid_out);
    }
    return false;
}

}  // namespace b200
