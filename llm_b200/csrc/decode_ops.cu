// llm_b200/csrc/decode_ops.cu -- the default decode schedule: 8 fused kernels per layer, replayed from ONE CUDA graph per token.
//
//   1 norm_pack        rms_norm(x) * attn_norm -> Q8 activation records                       (llama lib.rs:183-186 + mul_mat INIT)
//   2 mmv<QKV>         [wq|wk|wv] x, epilogue: RoPE on Q/K rows, K/V rows -> f16 cache at n_past (:190-244)
//   3 attn_kq          KQ = K . f16(Q)                                                        (:246-265)
//   4 attn_sv          scale + soft_max + V^T . f16(P), epilogue: quantize the merged row     (:268-307)
//   5 mmv<RES>         wo x + inpSA                                                           (:310-314)
//   6 norm_pack        rms_norm(inpFF) * ffn_norm                                             (:318-321)
//   7 mmv<SILU>        [w1|w3] x (rows interleaved in 32-row pieces), epilogue: silu(w1 x) * (w3 x) quantized (:323-330)
//   8 mmv<RES>         w2 h + inpFF                                                           (:332-334)
// Everything that depends on the position reads n_past from DEVICE memory, so the captured graph is valid for every token; the
// last node increments it.  All arithmetic is the bit-exact arithmetic of exact.cu / rowops.cu (same device functions as decode.cu).
#include <vector>

#include "decode.h"
#include "stream_core.cuh"

namespace b200 {

using namespace stream;

namespace {

__device__ __forceinline__ float lutf(const uint16_t *t, float x) { return f16_bits_to_f32(__ldg(t + f32_to_f16_bits(x))); }
__device__ __forceinline__ float f16dot_tree(float s) {          // ggml_vec_dot_f16 reduction order, see exact.cu
    s = __fadd_rn(s, __shfl_down_sync(0xffffffffu, s, 16));
    s = __fadd_rn(s, __shfl_down_sync(0xffffffffu, s, 8));
    s = __fadd_rn(s, __shfl_down_sync(0xffffffffu, s, 4));
    s = __fadd_rn(s, __shfl_down_sync(0xffffffffu, s, 1));
    s = __fadd_rn(s, __shfl_down_sync(0xffffffffu, s, 2));
    return s;
}

// ---- 1 / 6: rms_norm * gain -> records.  grid = e/128 CTAs of 256 threads; every CTA reduces the whole row (16 KB from L2) and
//      quantizes its own 4 blocks per warp pass. ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) norm_pack_kernel(const float *__restrict__ x, const float *__restrict__ gain, int4 *__restrict__ pack,
                                                        int e, float eps, int q81, int off, int scale16) {
    __shared__ double shd[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double s = 0.0;
    for (int i = tid; i < e / 4; i += 256) {
        const float4 v = __ldg((const float4 *)x + i);
        s += (double)__fmul_rn(v.x, v.x); s += (double)__fmul_rn(v.y, v.y); s += (double)__fmul_rn(v.z, v.z); s += (double)__fmul_rn(v.w, v.w);
    }
    s = warp_sum(s);
    if (lane == 0) shd[warp] = s;
    __syncthreads();
    const double tot = ((shd[0] + shd[1]) + (shd[2] + shd[3])) + ((shd[4] + shd[5]) + (shd[6] + shd[7]));
    const float mean = (float)(tot / (double)e);
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, eps)));
    // this CTA's blocks: 8 warps x 4 blocks = 32 blocks per CTA
    const int b = blockIdx.x * 32 + warp * 4 + (lane >> 3);
    const bool active = b < e / QK;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
        const float4 xv = __ldg((const float4 *)x + b * 8 + (lane & 7)), gv = __ldg((const float4 *)gain + b * 8 + (lane & 7));
        v.x = __fmul_rn(__fmul_rn(xv.x, scale), gv.x); v.y = __fmul_rn(__fmul_rn(xv.y, scale), gv.y);
        v.z = __fmul_rn(__fmul_rn(xv.z, scale), gv.z); v.w = __fmul_rn(__fmul_rn(xv.w, scale), gv.w);
    }
    pack_quad(v, pack + (active ? b : 0) * 4, lane, active, q81, off, scale16);
}

// ---- mat-vec with fused epilogues --------------------------------------------------------------------------------------------------
enum { EPI_RES = 0, EPI_QKV = 1, EPI_SILU = 2, EPI_LOGITS = 3 };

struct MmvArgs {
    const int4 *xpack;            // input records
    float *dst; const float *addend;                       // EPI_RES / EPI_LOGITS
    // EPI_QKV
    float *q; __half *K, *V; const float2 *rope_cs; int rope_half, hd, e, gqa, n_ctx; const int *n_past;
    // EPI_SILU
    int4 *xpack_out; const uint16_t *lut_silu;
    int q81, off, scale16;
    int *n_past_inc;              // EPI_LOGITS: the last node of the token increments InferenceSession::n_past on the device
};

template <int TYPE, int EPI>
__global__ void __launch_bounds__(STHREADS) mmv_fused_kernel(const QWeight w, const MmvArgs A) {
    using T = St<TYPE>;
    extern __shared__ __align__(128) uint8_t smem[];
    Ring R{(uint64_t *)smem, (uint64_t *)smem + SST_MAX, smem + 256, 0u, SST};
    int4 *sx = (int4 *)(R.base + T::RING_BYTES);
    const int tid = threadIdx.x;
    constexpr int G = EPI == EPI_SILU ? 2 : 1;
    if (tid == 0) ring_init(R.full, R.empty, SST);
    __syncthreads();
    if (tid >= SCOMPUTE) { produce_matvec<TYPE>(w, R, blockIdx.x, gridDim.x, tid & 31, G); return; }
    for (int i = tid; i < (int)w.nb * 4; i += SCOMPUTE) cp16(smem_u32(sx + i), A.xpack + i);   // all 16-byte copies in flight at once
    asm volatile("cp.async.wait_all;" ::: "memory");
    float *stash = (float *)(sx + (size_t)w.nb * 4);      // 64 floats behind the records (EPI_SILU)
    compute_sync();
    const int lane = tid & 31, warp = tid >> 5;
    if (EPI == EPI_RES || EPI == EPI_LOGITS) {
        consume_matvec<TYPE>(w, sx, R, blockIdx.x, gridDim.x, tid, [&](int64_t row, float v) {
            if ((tid & 3) == 0 && row < w.N) A.dst[row] = A.addend ? __fadd_rn(v, A.addend[row]) : v;
        });
        if (EPI == EPI_LOGITS && blockIdx.x == 0 && tid == 0) *A.n_past_inc = *A.n_past_inc + 1;
    } else if (EPI == EPI_QKV) {
        const int p = __ldg(A.n_past);
        consume_matvec<TYPE>(w, sx, R, blockIdx.x, gridDim.x, tid, [&](int64_t row, float v) {
            const float other = __shfl_xor_sync(0xffffffffu, v, 4);              // rotation partner: rows 2i, 2i+1 sit in adjacent quads
            if ((tid & 3) != 0 || row >= w.N) return;
            if (row < A.e + A.gqa) {                                               // ggml_rope mode 0 (LC/ggml.c:11859-11874)
                const int within = (int)(row < A.e ? row : row - A.e);
                const float2 cs = __ldg(A.rope_cs + (int64_t)p * A.rope_half + (within % A.hd) / 2);
                const bool even = (row & 1) == 0;
                const float x0 = even ? v : other, x1 = even ? other : v;
                const float out = even ? __fmaf_rn(x0, cs.x, -__fmul_rn(x1, cs.y)) : __fmaf_rn(x0, cs.y, __fmul_rn(x1, cs.x));
                if (row < A.e) A.q[row] = out;
                else A.K[(int64_t)p * A.gqa + within] = __float2half_rn(out);
            } else {
                A.V[(int64_t)(row - A.e - A.gqa) * A.n_ctx + p] = __float2half_rn(v);
            }
        });
    } else {   // EPI_SILU
        consume_matvec<TYPE>(w, sx, R, blockIdx.x, gridDim.x, tid, [&](int64_t row, float v) {
            if ((tid & 3) == 0) stash[row & 63] = v;
            if (((row >> 5) & 1) == 0) return;
            compute_sync();
            if (warp == 0) {
                const int w8 = lane & 7;
                const float4 a = ((const float4 *)stash)[w8], b = ((const float4 *)stash)[8 + w8];
                float4 hm;
                hm.x = __fmul_rn(lutf(A.lut_silu, a.x), b.x); hm.y = __fmul_rn(lutf(A.lut_silu, a.y), b.y);
                hm.z = __fmul_rn(lutf(A.lut_silu, a.z), b.z); hm.w = __fmul_rn(lutf(A.lut_silu, a.w), b.w);
                pack_quad(hm, A.xpack_out + (row >> 6) * 4, lane, lane < 8, A.q81, A.off, A.scale16);
            }
            compute_sync();
        }, G);
    }
}

// ---- 3: KQ.  CTA = (64 cached positions, head); CTAs past n_kv exit at once (the grid is sized for the context bucket).  All
//      16-byte loads of the K tile are issued before the first one is consumed. ---------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(128) attn_kq_kernel(const float *__restrict__ q, const __half *__restrict__ Kl, float *__restrict__ kq,
                                                      const int *__restrict__ n_past, int gqa, int n_head, int n_head_kv, int n_ctx) {
    __shared__ __align__(16) __half q16[HD];
    __shared__ __align__(16) __half kt[64 * HD];
    const int n_kv = __ldg(n_past) + 1;
    const int j0 = blockIdx.x * 64, h = blockIdx.y;
    if (j0 >= n_kv) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hk = h / (n_head / n_head_kv);
    const int rows = n_kv - j0 < 64 ? n_kv - j0 : 64;
    constexpr int VPR = HD / 8;                       // 16-byte vectors per K row
    constexpr int NV = 64 * VPR / 128;                // vectors per thread
    int4 v[NV];
#pragma unroll
    for (int u = 0; u < NV; u++) {
        const int i = tid + u * 128, rr = i / VPR, cc = i % VPR;
        v[u] = rr < rows ? __ldg((const int4 *)(Kl + (int64_t)(j0 + rr) * gqa + hk * HD) + cc) : make_int4(0, 0, 0, 0);
    }
    for (int i = tid; i < HD; i += 128) q16[i] = __float2half_rn(q[h * HD + i]);
#pragma unroll
    for (int u = 0; u < NV; u++) ((int4 *)kt)[tid + u * 128] = v[u];
    __syncthreads();
    for (int jj = warp; jj < rows; jj += 4) {
        const __half *krow = kt + jj * HD;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < HD; k += 32) s = __fmaf_rn(__half2float(krow[k + lane]), __half2float(q16[k + lane]), s);
        s = f16dot_tree(s);
        if (lane == 0) kq[(int64_t)h * n_ctx + j0 + jj] = (float)(double)s;
    }
}

// ---- 4: scale + soft_max + KQV for 32 channels of one head, then quantize those 32 outputs (one block of wo's input) ---------------
constexpr int KC = 128;
__global__ void __launch_bounds__(128) attn_sv_kernel(const float *__restrict__ kq, const __half *__restrict__ Vl, int4 *__restrict__ xpack_out,
                                                      const int *__restrict__ n_past, const uint16_t *__restrict__ lut_exp, float kq_scale,
                                                      int hd, int n_head, int n_head_kv, int n_ctx, int q81, int off, int scale16) {
    extern __shared__ __align__(16) uint8_t sm[];
    __shared__ double shd[8];
    __shared__ float shf[4], stash[32];
    const int n_kv = __ldg(n_past) + 1;
    const int per_head = hd / 32, h = blockIdx.x / per_head, c0 = (blockIdx.x - h * per_head) * 32;
    const int hk = h / (n_head / n_head_kv);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *sc = (float *)sm;
    __half *p16 = (__half *)(sm + (size_t)n_ctx * 4);
    __half *vt = (__half *)(sm + (size_t)n_ctx * 6);
    __half *vleft = vt + 32 * KC;
    // leftover V columns and the first V tile are requested before anything else so that their latency hides behind the soft_max
    const int np = n_kv & ~31;
    if (np < n_kv) {
        const int rr = tid >> 2, part = tid & 3;
        ((int4 *)vleft)[tid] = __ldg((const int4 *)(Vl + (int64_t)(hk * hd + c0 + rr) * n_ctx + np) + part);
    }
    int4 pre[KC / 32];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int u = 0; u < KC / 32; u++) {
            const int i = tid + u * 128, rr = i / (KC / 8), cc = i - rr * (KC / 8);
            pre[u] = __ldg((const int4 *)(Vl + (int64_t)(hk * hd + c0 + rr) * n_ctx + k0) + cc);
        }
    };
    if (np > 0) fetch(0);
    float mx = -INFINITY;
    constexpr int SB = 8;                              // scores per thread per pass: 8 loads in flight
    for (int jb = 0; jb < n_kv; jb += SB * 128) {
        float sv[SB];
#pragma unroll
        for (int u = 0; u < SB; u++) { const int j = jb + tid + u * 128; sv[u] = j < n_kv ? kq[(int64_t)h * n_ctx + j] : 0.f; }
#pragma unroll
        for (int u = 0; u < SB; u++) { const int j = jb + tid + u * 128; if (j < n_kv) { const float v = __fmul_rn(sv[u], kq_scale); sc[j] = v; mx = fmaxf(mx, v); } }
    }
    mx = warp_max(mx);
    if (lane == 0) shf[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(shf[0], shf[1]), fmaxf(shf[2], shf[3]));
    double s = 0.0;
    for (int jb = 0; jb < n_kv; jb += SB * 128) {
        uint16_t ev16[SB];
#pragma unroll
        for (int u = 0; u < SB; u++) { const int j = jb + tid + u * 128; ev16[u] = j < n_kv ? __ldg(lut_exp + f32_to_f16_bits(__fsub_rn(sc[j], mx))) : (uint16_t)0; }
#pragma unroll
        for (int u = 0; u < SB; u++) { const int j = jb + tid + u * 128; if (j < n_kv) { const float ev = f16_bits_to_f32(ev16[u]); sc[j] = ev; s += (double)ev; } }
    }
    s = warp_sum(s);
    if (lane == 0) shd[warp] = s;
    __syncthreads();
    const float inv = (float)(1.0 / ((shd[0] + shd[1]) + (shd[2] + shd[3])));
    for (int j = tid; j < n_kv; j += 128) p16[j] = __float2half_rn(__fmul_rn(sc[j], inv));
    float acc[8];
#pragma unroll
    for (int cc = 0; cc < 8; cc++) acc[cc] = 0.f;
    for (int k0 = 0; k0 < np; k0 += KC) {
        __syncthreads();
#pragma unroll
        for (int u = 0; u < KC / 32; u++) ((int4 *)vt)[tid + u * 128] = pre[u];
        __syncthreads();
        if (k0 + KC < np) fetch(k0 + KC);
        const int kend = np - k0 < KC ? np - k0 : KC;
#pragma unroll
        for (int cc = 0; cc < 8; cc++) {
            const __half *vrow = vt + (warp * 8 + cc) * KC;
            for (int k = lane; k < kend; k += 32) acc[cc] = __fmaf_rn(__half2float(vrow[k]), __half2float(p16[k0 + k]), acc[cc]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int cc = 0; cc < 8; cc++) {
        const float a = f16dot_tree(acc[cc]);
        if (lane == 0) {
            const __half *vrow = vleft + (warp * 8 + cc) * 32;
            double sumf = (double)a;
            for (int k = np; k < n_kv; k++) sumf += (double)__fmul_rn(__half2float(vrow[k - np]), __half2float(p16[k]));
            stash[warp * 8 + cc] = (float)sumf;
        }
    }
    __syncthreads();
    if (warp == 0) pack_quad(((const float4 *)stash)[lane & 7], xpack_out + (int64_t)((h * hd + c0) / QK) * 4, lane, lane < 8, q81, off, scale16);
}

template <int TYPE, int EPI>
void launch_mmv(const QWeight &w, const MmvArgs &A, cudaStream_t st) {
    using T = St<TYPE>;
    const int smem = 256 + T::RING_BYTES + (int)w.nb * 64 + 256;
    static int smem_set = 0, ctas_per_sm = 0, sms = 0, occ_smem = -1;
    if (smem > smem_set) {
        B200_ASSERT(smem <= 227 * 1024);
        B200_CHECK(cudaFuncSetAttribute(mmv_fused_kernel<TYPE, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        smem_set = smem;
    }
    if (!sms) { int dev; B200_CHECK(cudaGetDevice(&dev)); B200_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)); }
    if (occ_smem != smem) { B200_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, mmv_fused_kernel<TYPE, EPI>, STHREADS, smem)); occ_smem = smem; }
    constexpr int G = EPI == EPI_SILU ? 2 : 1;
    const int64_t groups = ((w.N + SR - 1) / SR + G - 1) / G;
    const int64_t slots = (int64_t)sms * (ctas_per_sm > 0 ? ctas_per_sm : 1);
    mmv_fused_kernel<TYPE, EPI><<<(unsigned)(groups < slots ? groups : slots), STHREADS, smem, st>>>(w, A);
    B200_CHECK(cudaGetLastError());
}

template <int TYPE>
void decode_ops_t(const DecodeParams &P, const std::vector<DecodeLayer> &layers, int n_kv_bucket, int4 *xpack_a, cudaStream_t st, int *launches) {
    const int q81 = has_min(TYPE) ? 1 : 0, off = TYPE == T_Q5_0 ? 16 : 0, s16 = TYPE == T_Q4_0 ? 1 : 0;
    const int e = P.e, f = P.f;
    int n = 0;
    get_rows_q(P.wte, P.token, P.x, 1, st); n++;
    const size_t sv_smem = (size_t)P.n_ctx * 6 + 32 * KC * 2 + 32 * 32 * 2;
    static size_t sv_set = 48 * 1024;
    if (sv_smem > sv_set) { B200_CHECK(cudaFuncSetAttribute(attn_sv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sv_smem)); sv_set = sv_smem; }
    for (int il = 0; il < P.n_layer; il++) {
        const DecodeLayer &L = layers[il];
        norm_pack_kernel<<<(e / QK + 31) / 32, 256, 0, st>>>(P.x, L.attn_norm, xpack_a, e, P.eps, q81, off, s16); n++;
        MmvArgs A{}; A.xpack = xpack_a; A.q = P.q; A.K = L.K; A.V = L.V; A.rope_cs = P.rope_cs; A.rope_half = P.rope_half; A.hd = P.hd; A.e = e; A.gqa = P.gqa;
        A.n_ctx = P.n_ctx; A.n_past = P.n_past;
        launch_mmv<TYPE, EPI_QKV>(L.wqkv, A, st); n++;
        if (P.hd == 128) attn_kq_kernel<128><<<dim3((n_kv_bucket + 63) / 64, P.n_head), 128, 0, st>>>(P.q, L.K, P.kq, P.n_past, P.gqa, P.n_head, P.n_head_kv, P.n_ctx);
        else             attn_kq_kernel<64><<<dim3((n_kv_bucket + 63) / 64, P.n_head), 128, 0, st>>>(P.q, L.K, P.kq, P.n_past, P.gqa, P.n_head, P.n_head_kv, P.n_ctx);
        n++;
        attn_sv_kernel<<<P.n_head * (P.hd / 32), 128, sv_smem, st>>>(P.kq, L.V, P.xpack_d, P.n_past, P.lut_exp, P.kq_scale, P.hd, P.n_head, P.n_head_kv, P.n_ctx,
                                                                      q81, off, s16); n++;
        MmvArgs Bo{}; Bo.xpack = P.xpack_d; Bo.dst = P.ff; Bo.addend = P.x;
        launch_mmv<TYPE, EPI_RES>(L.wo, Bo, st); n++;
        norm_pack_kernel<<<(e / QK + 31) / 32, 256, 0, st>>>(P.ff, L.ffn_norm, xpack_a, e, P.eps, q81, off, s16); n++;
        MmvArgs C{}; C.xpack = xpack_a; C.xpack_out = P.xpack_f; C.lut_silu = P.lut_silu; C.q81 = q81; C.off = off; C.scale16 = s16;
        launch_mmv<TYPE, EPI_SILU>(L.w13, C, st); n++;
        MmvArgs D{}; D.xpack = P.xpack_f; D.dst = P.x; D.addend = P.ff;
        launch_mmv<TYPE, EPI_RES>(L.w2, D, st); n++;
    }
    norm_pack_kernel<<<(e / QK + 31) / 32, 256, 0, st>>>(P.x, P.norm, xpack_a, e, P.eps, q81, off, s16); n++;
    MmvArgs Z{}; Z.xpack = xpack_a; Z.dst = P.logits; Z.addend = nullptr; Z.n_past_inc = P.n_past;
    launch_mmv<TYPE, EPI_LOGITS>(P.output, Z, st); n++;
    B200_CHECK(cudaGetLastError());
    (void)f;
    if (launches) *launches = n;
}

}  // namespace

// Enqueue one decode step (position read from *P.n_past on the device) on `st`.  n_kv_bucket >= n_past + 1 sizes the KQ grid.
void decode_ops_enqueue(const DecodeParams &P, const std::vector<DecodeLayer> &layers, int wtype, int n_kv_bucket, int4 *xpack_a, cudaStream_t st, int *launches) {
    switch (wtype) {
        case T_Q4_0: decode_ops_t<T_Q4_0>(P, layers, n_kv_bucket, xpack_a, st, launches); break;
        case T_Q4_1: decode_ops_t<T_Q4_1>(P, layers, n_kv_bucket, xpack_a, st, launches); break;
        case T_Q5_0: decode_ops_t<T_Q5_0>(P, layers, n_kv_bucket, xpack_a, st, launches); break;
        case T_Q5_1: decode_ops_t<T_Q5_1>(P, layers, n_kv_bucket, xpack_a, st, launches); break;
        case T_Q8_0: decode_ops_t<T_Q8_0>(P, layers, n_kv_bucket, xpack_a, st, launches); break;
        default: B200_ASSERT(!"decode_ops_enqueue: unsupported weight type");
    }
}

}  // namespace b200
