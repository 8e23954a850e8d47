// llm_b200/csrc/attn.cu -- ggml_mul_mat with an F16 src0: the two attention mat-muls on the f16 KV cache
// (KQ = K x Q and KQV = V^T x softmax, crates/models/llama/src/lib.rs:265,296).
//
// Reference arithmetic (F16 type traits, LC/ggml.c:1650-1656; INIT conversion :10504-10520; ggml_vec_dot_f16 :2325-2359):
// each src1 row is first ROUNDED TO FP16, products f16 x f16 are exact in f32 and are accumulated with f32 fma.  This kernel
// keeps all of that and only changes the order of the f32 additions (one running sum per output instead of 32 AVX lanes).
#include "kernels.cuh"

namespace b200 {

constexpr int F16MM_TJ = 4;      // src1 rows per CTA
constexpr int F16MM_KC = 256;    // k chunk staged in shared memory
constexpr int F16MM_WARPS = 4;   // 128 src0 rows per CTA, lane <-> src0 row

template <bool VEC>
__global__ void __launch_bounds__(F16MM_WARPS * 32) mul_mat_f16_kernel(const char *__restrict__ src0, int64_t ne00, int64_t ne01, int64_t ne02, int64_t nb01, int64_t nb02,
                                                                       const char *__restrict__ src1, int64_t ne11, int64_t ne12, int64_t nb11, int64_t nb12,
                                                                       char *__restrict__ dst, int64_t nbd1, int64_t nbd2) {
    __shared__ float xs[F16MM_TJ][F16MM_KC];
    const int64_t i2 = blockIdx.z;
    const int64_t j0 = (int64_t)blockIdx.y * F16MM_TJ;
    const int64_t i0 = (int64_t)blockIdx.x * (F16MM_WARPS * 32) + threadIdx.x;
    const int64_t i02 = i2 / (ne12 / ne02);                                   // broadcast rule, LC/ggml.c:10549
    const bool active = i0 < ne01;
    const char *row0 = src0 + i02 * nb02 + (active ? i0 : 0) * nb01;
    float acc[F16MM_TJ];
#pragma unroll
    for (int j = 0; j < F16MM_TJ; j++) acc[j] = 0.f;

    for (int64_t k0 = 0; k0 < ne00; k0 += F16MM_KC) {
        const int kc = (int)((ne00 - k0) < F16MM_KC ? (ne00 - k0) : F16MM_KC);
        __syncthreads();
        for (int idx = threadIdx.x; idx < F16MM_TJ * F16MM_KC; idx += blockDim.x) {
            const int j = idx / F16MM_KC, k = idx % F16MM_KC;
            float v = 0.f;
            if (k < kc && j0 + j < ne11) v = __half2float(__float2half_rn(*(const float *)(src1 + i2 * nb12 + (j0 + j) * nb11 + (k0 + k) * 4)));
            xs[j][k] = v;
        }
        __syncthreads();
        if (!active) continue;
        const __half *a = (const __half *)row0 + k0;
        int k = 0;
        if (VEC) {
            for (; k + 8 <= kc; k += 8) {
                const int4 v = *(const int4 *)(a + k);
                const __half2 *h = (const __half2 *)&v;
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const float2 f = __half22float2(h[p]);
#pragma unroll
                    for (int j = 0; j < F16MM_TJ; j++) { acc[j] = __fmaf_rn(f.x, xs[j][k + 2 * p], acc[j]); acc[j] = __fmaf_rn(f.y, xs[j][k + 2 * p + 1], acc[j]); }
                }
            }
        }
        for (; k < kc; k++) {
            const float f = __half2float(a[k]);
#pragma unroll
            for (int j = 0; j < F16MM_TJ; j++) acc[j] = __fmaf_rn(f, xs[j][k], acc[j]);
        }
    }
    if (!active) return;
#pragma unroll
    for (int j = 0; j < F16MM_TJ; j++)
        if (j0 + j < ne11) *(float *)(dst + i2 * nbd2 + (j0 + j) * nbd1 + i0 * 4) = acc[j];
}

void mul_mat_f16(const __half *src0, int64_t ne00, int64_t ne01, int64_t ne02, int64_t nb01, int64_t nb02,
                 const float *src1, int64_t ne11, int64_t ne12, int64_t nb11, int64_t nb12,
                 float *dst, int64_t nbd1, int64_t nbd2, cudaStream_t st) {
    if (ne01 == 0 || ne11 == 0 || ne12 == 0) return;
    B200_ASSERT(ne12 % ne02 == 0);
    dim3 grid((unsigned)((ne01 + F16MM_WARPS * 32 - 1) / (F16MM_WARPS * 32)), (unsigned)((ne11 + F16MM_TJ - 1) / F16MM_TJ), (unsigned)ne12);
    const bool vec = ((uintptr_t)src0 % 16 == 0) && (nb01 % 16 == 0) && (nb02 % 16 == 0);
    if (vec) mul_mat_f16_kernel<true><<<grid, F16MM_WARPS * 32, 0, st>>>((const char *)src0, ne00, ne01, ne02, nb01, nb02, (const char *)src1, ne11, ne12, nb11, nb12, (char *)dst, nbd1, nbd2);
    else     mul_mat_f16_kernel<false><<<grid, F16MM_WARPS * 32, 0, st>>>((const char *)src0, ne00, ne01, ne02, nb01, nb02, (const char *)src1, ne11, ne12, nb11, nb12, (char *)dst, nbd1, nbd2);
    B200_CHECK(cudaGetLastError());
}

}  // namespace b200
