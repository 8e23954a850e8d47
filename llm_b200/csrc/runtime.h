// llm_b200/csrc/runtime.h -- process-wide device state shared by the seam (seam.cu) and the native session (session.cu).
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <vector>

namespace b200 {

// Bump allocator over one device block; grows by replacement. Pointers are valid until the next reset().
struct Arena {
    char *base = nullptr;
    size_t cap = 0, off = 0;
    std::vector<void *> retired;
    void *get(size_t bytes, cudaStream_t st);
    void reset();
    void release();
};

struct Runtime {
    bool inited = false;
    int device = 0, device_count = 0, sm_count = 0;
    cudaStream_t stream = nullptr;     // one non-blocking stream: every kernel and copy of this backend is ordered on it
    Arena op_arena;                    // per-node temporaries of the seam front end
    bool fast = false;                 // B200_FAST=1: order-free integer-exact kernels (mmvq/mmq/attn.cu) instead of the bit-exact ones
    std::mutex mu;
    void ensure_init();                // exits(1) if there is no CUDA device: this backend has no CPU fallback
};

Runtime &rt();

}  // namespace b200
