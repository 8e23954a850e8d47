// llm_b200/csrc/tp.cuh -- tensor-parallel decode: the exchange of activation slices between the GPUs of one NVSwitch box.
//
// Scheme: every weight matrix is split by OUTPUT ROWS (heads for wq|wk|wv, n_ff/G rows of w1|w3, n_embd/G rows of wo / w2,
// n_vocab/G rows of the lm_head), so every dst element is still one complete ggml_vec_dot over the full K (LC/ggml.c:10570-10572) and the result is
// bit-identical to the single-GPU / CPU result by construction.  What travels between GPUs are the OUTPUT slices a following operator needs whole:
//   attention rows (as Q8 records)  -> wo's input        buffer XD
//   wo x + inpSA  (f32)             -> ffn rms_norm      buffer FF
//   silu(w1 x) * (w3 x) (records)   -> w2's input        buffer XF
//   w2 h + inpFF  (f32)             -> next rms_norm     buffer X
//   logits (f32)                    -> the caller        buffer LOGITS
// There is no collective call on the data path: the epilogue that produces a slice stores it straight into EVERY rank's buffer through peer-mapped
// pointers (NVLink / NVSwitch P2P stores, CUDA IPC mappings of one "exchange slab" per rank), then the last CTA of the kernel releases a per-source flag
// on every rank; the consuming kernel acquires the G flags before it reads the buffer.  Flags carry a monotonically increasing value
// (epoch * (n_layer + 1) + layer + 1, epoch = tokens decoded so far) so nothing is ever reset.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int TP_MAX = 8;
enum { TPB_X = 0, TPB_FF = 1, TPB_XD = 2, TPB_XF = 3, TPB_LOGITS = 4, TPB_COUNT = 5 };

struct TpCtx {                 // kernel argument (POD); world == 1: single GPU, every helper below degenerates to the local store
    int world = 1, rank = 0;
    char *peer[TP_MAX] = {};   // base of rank p's exchange slab as mapped in THIS process (peer[rank] is the local slab)
    unsigned *epoch = nullptr;     // local: [0] tokens decoded so far, [1] number of flag waits that timed out (a peer died)
    unsigned *arrivals = nullptr;  // local: one CTA-arrival counter per kernel site (5 per layer + 1)
    uint32_t off[TPB_COUNT] = {};  // byte offsets of the buffers inside a slab
    uint32_t off_flags = 0;        // [TPB_COUNT][TP_MAX] flags, 32 bytes apart
    unsigned vmul = 1;             // n_layer + 1
    int nowait = 0;                // measurement aid (b200_session_tp_set_nowait): skip the flag waits -- results are garbage, the time is compute + stores only
};

// what one kernel instance waits for / signals (baked into the CUDA graph; the epoch is read from device memory)
struct TpSync {
    int wait_buf = -1; unsigned wait_v = 0;     // acquire flags[wait_buf][0..G) >= epoch * vmul + wait_v before reading the buffer
    int sig_buf = -1; unsigned sig_v = 0;       // after the last CTA: flags[sig_buf][rank] = epoch * vmul + sig_v on every rank
    int site = 0;                               // index into TpCtx::arrivals
    int bump_epoch = 0;                         // the token's last kernel: epoch += 1 once every CTA has arrived
};

__device__ __forceinline__ unsigned *tp_flag(const TpCtx &T, int p, int buf, int src) { return (unsigned *)(T.peer[p] + T.off_flags + (size_t)(buf * TP_MAX + src) * 32); }

__device__ __forceinline__ void tp_store_f32(const TpCtx &T, int buf, int64_t idx, float v) {
#pragma unroll 1
    for (int p = 0; p < T.world; p++) ((float *)(T.peer[p] + T.off[buf]))[idx] = v;
}
__device__ __forceinline__ void tp_store_rec(const TpCtx &T, int buf, int64_t idx, int4 v) {
#pragma unroll 1
    for (int p = 0; p < T.world; p++) ((int4 *)(T.peer[p] + T.off[buf]))[idx] = v;
}

// one thread: spin until every rank's slice of `buf` for this (token, layer) has landed here
__device__ __forceinline__ void tp_wait_thread(const TpCtx &T, const TpSync &S) {
    if (T.world <= 1 || S.wait_buf < 0 || T.nowait) return;
    const unsigned want = *(volatile unsigned *)T.epoch * T.vmul + S.wait_v;
    for (int src = 0; src < T.world; src++) {
        const unsigned *f = tp_flag(T, T.rank, S.wait_buf, src);
        unsigned v;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (clock64() - t0 > 6000000000LL) { atomicAdd(T.epoch + 1, 1u); break; }     // a peer is gone: do not hang the GPU, count it (b200_tp_timeouts)
        } while ((int)(v - want) < 0);
    }
}

// Called by ONE thread per CTA after every thread of the CTA has issued its remote stores and a CTA-level barrier (so the fence below orders them):
// the last CTA to arrive publishes the flag on every rank.
__device__ __forceinline__ void tp_signal_thread(const TpCtx &T, const TpSync &S, unsigned n_ctas) {
    if (T.world <= 1 || S.sig_buf < 0) return;
    const unsigned value = *(volatile unsigned *)T.epoch * T.vmul + S.sig_v;
    __threadfence_system();
    if (atomicAdd(T.arrivals + S.site, 1u) == n_ctas - 1) {
        T.arrivals[S.site] = 0;
        __threadfence_system();
        for (int p = 0; p < T.world; p++) {
            unsigned *f = tp_flag(T, p, S.sig_buf, T.rank);
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(value) : "memory");
        }
        if (S.bump_epoch) *T.epoch = *(volatile unsigned *)T.epoch + 1;
    }
}

}  // namespace b200
