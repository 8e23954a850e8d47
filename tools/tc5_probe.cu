// tools/tc5_probe.cu -- stand-alone B200 probe for the tcgen05 plumbing of llm_b200/csrc/tc5.cuh (not part of the product):
//   1. one tcgen05.mma (M128 N128 K16, f16 -> f32) on operands laid out by the HOST in the candidate shared-memory layouts / descriptors
//      (128B-swizzled and un-swizzled K-major), checked against a CPU product: tells which (layout, LBO, SBO) reading is right;
//   2. a TMA tile load with CU_TENSOR_MAP_SWIZZLE_128B dumped back: checks the chunk ^ (row & 7) placement the MMA descriptor assumes;
//   3. micro-benchmarks the exact prefill GEMM's budget depends on: tcgen05.ld bandwidth per SM, packed fma.rn.f32x2 vs scalar fma issue rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I llm_b200/csrc tools/tc5_probe.cu -o tools/tc5_probe
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <cuda_fp16.h>

#include "tc5.cuh"

using namespace b200::tc5;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d: %s\n", cudaGetErrorName(e_), __LINE__, cudaGetErrorString(e_)); exit(2); } } while (0)

struct MmaCfg { uint32_t a_off, a_lbo, a_sbo, a_layout, b_off, b_lbo, b_sbo, b_layout; };

__global__ void __launch_bounds__(128) probe_mma(const uint8_t *a_img, const uint8_t *b_img, MmaCfg c, float *out) {
    extern __shared__ uint8_t raw[];
    const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t *p = raw + (sbase - smem_u32(raw));
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    for (int i = threadIdx.x; i < 32768 / 16; i += 128) { ((uint4 *)p)[i] = ((const uint4 *)a_img)[i]; ((uint4 *)(p + 32768))[i] = ((const uint4 *)b_img)[i]; }
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_slot), 128); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    bool dead = false;
    if (threadIdx.x == 0) {
        const uint64_t ad = make_smem_desc(sbase + c.a_off, c.a_lbo, c.a_sbo, c.a_layout);
        const uint64_t bd = make_smem_desc(sbase + 32768 + c.b_off, c.b_lbo, c.b_sbo, c.b_layout);
        mma_f16_ss(tmem, ad, bd, make_idesc_f16(128, 128), 0u);
        tc_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0, dead);
    tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int cgrp = 0; cgrp < 4; cgrp++) {
        uint32_t r[32];
        tmem_ld_x32(tmem + ((uint32_t)(warp * 32) << 16) + cgrp * 32, r);
        tc_wait_ld();
        for (int i = 0; i < 32; i++) out[(warp * 32 + lane) * 128 + cgrp * 32 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 128);
}

__global__ void __launch_bounds__(128) probe_tma(const __grid_constant__ CUtensorMap tm, int c0, int c1, uint8_t *dump) {
    extern __shared__ uint8_t raw[];
    const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t *p = raw + (sbase - smem_u32(raw));
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    __syncthreads();
    bool dead = false;
    if (threadIdx.x == 0) { mbar_expect_tx(smem_u32(&bar), 16384); tma_load_2d(sbase, &tm, c0, c1, smem_u32(&bar)); }
    mbar_wait(smem_u32(&bar), 0, dead);
    for (int i = threadIdx.x; i < 16384 / 16; i += 128) ((uint4 *)dump)[i] = ((const uint4 *)p)[i];
}

// tcgen05.ld bandwidth: NW warps each issue `iters` x32 loads (4 KB per warp-instruction) from an allocated (uninitialised) TMEM region
template <int NW>
__global__ void __launch_bounds__(NW * 32) bench_ldtm(int iters, long long *cycles, uint32_t *sink) {
    __shared__ uint32_t tmem_slot;
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int warp = threadIdx.x >> 5;
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        uint32_t r[32], r2[32];
        tmem_ld_x32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + ((i * 64) & 511), r);
        tmem_ld_x32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + ((i * 64 + 32) & 511), r2);
        tc_wait_ld();
#pragma unroll
        for (int k = 0; k < 32; k++) acc ^= r[k] + r2[k];
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// fp32 issue rate: 32 independent chains per thread, scalar fma vs packed f32x2
template <bool PACKED>
__global__ void __launch_bounds__(256) bench_fma(int iters, long long *cycles, float *sink, float s0) {
    float2 a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float2 s = make_float2(s0, s0), d = make_float2(1e-7f, 2e-7f);
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (PACKED) a[i] = ffma2(s, d, a[i]);
            else { a[i].x = __fmaf_rn(s.x, d.x, a[i].x); a[i].y = __fmaf_rn(s.y, d.y, a[i].y); }
        }
    }
    const long long t1 = clock64();
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) r += a[i].x + a[i].y;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (r == 123.456f) sink[0] = r;
}

static void put_f16(std::vector<uint8_t> &img, size_t off, float v) { __half h = __float2half(v); memcpy(&img[off], &h, 2); }

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("device: %s sm_%d%d, %d SMs\n", pr.name, pr.major, pr.minor, pr.multiProcessorCount);
    // ---- logical operands: A[128][64] (we use K columns k0..k0+15), B[128][16] ----
    std::vector<float> A(128 * 64), Bm(128 * 16);
    srand(1);
    for (auto &v : A) v = (float)(rand() % 255 - 127);
    for (auto &v : Bm) v = (float)(rand() % 31 - 15);
    auto expect = [&](int k0, std::vector<float> &D) { D.assign(128 * 128, 0.f); for (int m = 0; m < 128; m++) for (int n = 0; n < 128; n++) { float s = 0; for (int k = 0; k < 16; k++) s += A[m * 64 + k0 + k] * Bm[n * 16 + k]; D[m * 128 + n] = s; } };
    // images
    auto img_sw128 = [&](const std::vector<float> &M, int rows, int cols, bool swz) {      // rows x 64 f16, 128 B per row
        std::vector<uint8_t> img(32768, 0);
        for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) { const int chunk = c / 8, cc = swz ? (chunk ^ (r & 7)) : chunk; put_f16(img, (size_t)r * 128 + cc * 16 + (c % 8) * 2, M[r * cols + c]); }
        return img;
    };
    auto img_none = [&](const std::vector<float> &M, int ld, int k0, uint32_t kstride, uint32_t gstride) {   // 128 x 16 canonical un-swizzled: 8x16B core matrices
        std::vector<uint8_t> img(32768, 0);
        for (int r = 0; r < 128; r++) for (int k = 0; k < 16; k++) put_f16(img, (size_t)(r / 8) * gstride + (k / 8) * kstride + (r % 8) * 16 + (k % 8) * 2, M[r * ld + k0 + k]);
        return img;
    };
    uint8_t *da, *db; float *dout; CK(cudaMalloc(&da, 32768)); CK(cudaMalloc(&db, 32768)); CK(cudaMalloc(&dout, 128 * 128 * 4));
    CK(cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560 + 1024));
    struct Var { const char *name; std::vector<uint8_t> a, b; MmaCfg c; int k0; };
    std::vector<Var> vars;
    const std::vector<uint8_t> bn = img_none(Bm, 16, 0, 128, 256);
    vars.push_back({"A sw128(k0=0)  LBO16 SBO1024 | B none kstride=LBO=128 gstride=SBO=256", img_sw128(A, 128, 64, true), bn, {0, 16, 1024, 2, 0, 128, 256, 0}, 0});
    vars.push_back({"A sw128(k0=16, +32B) same                                             ", img_sw128(A, 128, 64, true), bn, {32, 16, 1024, 2, 0, 128, 256, 0}, 16});
    vars.push_back({"A sw128(k0=48, +96B) same                                             ", img_sw128(A, 128, 64, true), bn, {96, 16, 1024, 2, 0, 128, 256, 0}, 48});
    vars.push_back({"A sw128(k0=0) | B none but desc LBO=256 SBO=128 (swapped reading)     ", img_sw128(A, 128, 64, true), bn, {0, 16, 1024, 2, 0, 256, 128, 0}, 0});
    vars.push_back({"A none(k0=0) LBO=128 SBO=256 | B none LBO=128 SBO=256                  ", img_none(A, 64, 0, 128, 256), bn, {0, 128, 256, 0, 0, 128, 256, 0}, 0});
    vars.push_back({"A none LBO=2048(kstride) SBO=128 (TMA-3D-style image) | B none          ", img_none(A, 64, 0, 2048, 128), bn, {0, 2048, 128, 0, 0, 128, 256, 0}, 0});
    vars.push_back({"A sw128 image WITHOUT swizzle, layout=2 (expect wrong)                 ", img_sw128(A, 128, 64, false), bn, {0, 16, 1024, 2, 0, 128, 256, 0}, 0});
    vars.push_back({"A sw128(k0=0) LBO=0                                                    ", img_sw128(A, 128, 64, true), bn, {0, 0, 1024, 2, 0, 128, 256, 0}, 0});
    vars.push_back({"A sw128 | B none with staggered groups SBO=272                         ", img_sw128(A, 128, 64, true), img_none(Bm, 16, 0, 128, 272), {0, 16, 1024, 2, 0, 128, 272, 0}, 0});
    for (auto &v : vars) {
        std::vector<float> D, got(128 * 128);
        expect(v.k0, D);
        CK(cudaMemcpy(da, v.a.data(), 32768, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, v.b.data(), 32768, cudaMemcpyHostToDevice));
        CK(cudaMemset(dout, 0xff, 128 * 128 * 4));
        probe_mma<<<1, 128, 66560 + 1024>>>(da, db, v.c, dout);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(got.data(), dout, 128 * 128 * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0; int bad = 0;
        for (int i = 0; i < 128 * 128; i++) { const double e = fabs((double)got[i] - D[i]); if (!(e == 0)) bad++; if (e > maxerr || e != e) maxerr = e; }
        printf("MMA  %-76s : %s (mismatches %d / 16384, max err %g)\n", v.name, bad == 0 ? "EXACT" : "wrong", bad, maxerr);
    }
    printf("timeouts: %d\n", check_timeout("probe"));
    // ---- TMA swizzle placement ----
    {
        const int R = 256, Ccols = 256;
        std::vector<__half> G(R * Ccols);
        for (int r = 0; r < R; r++) for (int c = 0; c < Ccols; c++) G[r * Ccols + c] = __float2half((float)((r * 7 + c * 3) % 2048));
        __half *dg; uint8_t *dd; CK(cudaMalloc(&dg, G.size() * 2)); CK(cudaMalloc(&dd, 16384));
        CK(cudaMemcpy(dg, G.data(), G.size() * 2, cudaMemcpyHostToDevice));
        CUtensorMap tm = make_tmap_2d_f16_sw128(dg, Ccols, 200 /* rows visible: 200 < 256 tests the zero fill */, Ccols * 2, 128);
        CK(cudaFuncSetAttribute(probe_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1024));
        probe_tma<<<1, 128, 16384 + 1024>>>(tm, 64, 128, dd);
        CK(cudaDeviceSynchronize());
        std::vector<uint8_t> h(16384); CK(cudaMemcpy(h.data(), dd, 16384, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int r = 0; r < 128; r++) for (int c = 0; c < 64; c++) {
            const int chunk = c / 8, cc = chunk ^ (r & 7);
            __half v; memcpy(&v, &h[(size_t)r * 128 + cc * 16 + (c % 8) * 2], 2);
            const float want = (128 + r) < 200 ? (float)(((128 + r) * 7 + (64 + c) * 3) % 2048) : 0.f;
            if (__half2float(v) != want) bad++;
        }
        printf("TMA  128B-swizzled [128 x 64] box, rows >= 200 zero-filled: %s (%d mismatches)\n", bad == 0 ? "as assumed" : "DIFFERENT", bad);
        printf("timeouts: %d\n", check_timeout("probe"));
    }
    // ---- micro-benchmarks ----
    long long *dc; uint32_t *ds; CK(cudaMalloc(&dc, 1024 * 8)); CK(cudaMalloc(&ds, 64));
    const int nsm = pr.multiProcessorCount;
    std::vector<long long> hc(1024);
    auto avg = [&](int n) { CK(cudaMemcpy(hc.data(), dc, n * 8, cudaMemcpyDeviceToHost)); double s = 0; for (int i = 0; i < n; i++) s += hc[i]; return s / n; };
    for (int rep = 0; rep < 2; rep++) {
        bench_ldtm<4><<<nsm, 128>>>(2000, dc, ds); CK(cudaDeviceSynchronize());
        const double c4 = avg(nsm);
        bench_ldtm<8><<<nsm, 256>>>(2000, dc, ds); CK(cudaDeviceSynchronize());
        const double c8 = avg(nsm);
        bench_ldtm<16><<<nsm, 512>>>(2000, dc, ds); CK(cudaDeviceSynchronize());
        const double c16 = avg(nsm);
        printf("LDTM x32 pairs: 4 warps %.1f B/clk/SM, 8 warps %.1f B/clk/SM, 16 warps %.1f B/clk/SM\n", 4.0 * 2000 * 8192 / c4, 8.0 * 2000 * 8192 / c8, 16.0 * 2000 * 8192 / c16);
        bench_fma<false><<<nsm * 2, 256>>>(4000, dc, (float *)ds, 1.0001f); CK(cudaDeviceSynchronize());
        const double f1 = avg(nsm * 2);
        bench_fma<true><<<nsm * 2, 256>>>(4000, dc, (float *)ds, 1.0001f); CK(cudaDeviceSynchronize());
        const double f2 = avg(nsm * 2);
        // per SM: 2 CTAs x 256 threads x 32 fma per iteration
        printf("fp32: scalar fma %.1f fma/clk/SM, packed f32x2 %.1f fma/clk/SM\n", 2.0 * 256 * 32 * 4000 / f1, 2.0 * 256 * 32 * 4000 / f2);
    }
    return 0;
}
