/*
 * include/ggml_b200.h -- the drop-in boundary: the reference's accelerator seam, served by libllm_b200.so.
 *
 * These are exactly the entry points the reference binds for its CUDA backend and that its graph executor
 * calls for every node ("LC/" = crates/ggml/sys/llama-cpp/ under /root/reference):
 *   declarations  LC/ggml-cuda.h:11-32
 *   Rust FFI      crates/ggml/sys/src/cuda.rs:7-77        (#[cfg(feature = "cublas")], sys/src/lib.rs:9-10)
 *   C callers     LC/ggml.c:4354-4355 (ggml_init -> ggml_init_cublas), :14584-14591 (per node ->
 *                 ggml_cuda_compute_forward), :16233-16237 (plan -> ggml_cuda_can_mul_mat)
 *   Rust callers  crates/ggml/src/accelerator/mod.rs:68-94, crates/ggml/src/tensor.rs:56-112,213-222
 * Link ggml.c (built with -DGGML_USE_CUBLAS) against libllm_b200.so instead of compiling LC/ggml-cuda.cu and
 * nothing above the seam changes; INTEGRATION.md shows the build.rs edit.
 *
 * The data contract is the reference's `struct ggml_tensor` / `struct ggml_compute_params`
 * (LC/ggml.h:395-431, :514-530; bindgen layout tests crates/ggml/sys/src/lib.rs:446 = 272 bytes).  A consumer
 * that already includes the reference's ggml.h keeps using those definitions (define GGML_B200_USE_GGML_H);
 * otherwise the mirror below is layout-identical -- tests/test_abi.py checks sizes and offsets.
 *
 * Error behaviour follows the reference backend: no error returns; a CUDA failure prints and exit(1)s
 * (LC/ggml-cuda.cu:24-53), an unsupported request aborts.  There is NO CPU fallback inside this library.
 */
#ifndef GGML_B200_H
#define GGML_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef GGML_B200_USE_GGML_H

enum b200_ggml_type {            /* enum ggml_type, LC/ggml.h:262-285 (only the values this backend touches) */
    B200_TYPE_F32 = 0, B200_TYPE_F16 = 1, B200_TYPE_Q4_0 = 2, B200_TYPE_Q4_1 = 3,
    B200_TYPE_Q5_0 = 6, B200_TYPE_Q5_1 = 7, B200_TYPE_Q8_0 = 8, B200_TYPE_Q8_1 = 9,
    B200_TYPE_I8 = 16, B200_TYPE_I16 = 17, B200_TYPE_I32 = 18,
};
enum b200_ggml_backend { B200_BACKEND_CPU = 0, B200_BACKEND_GPU = 10, B200_BACKEND_GPU_SPLIT = 20 };  /* LC/ggml.h:290-294 */
enum b200_ggml_op {              /* enum ggml_op, LC/ggml.h:312-386: ordinal positions of the ops the seam serves */
    B200_OP_NONE = 0, B200_OP_DUP = 1, B200_OP_ADD = 2, B200_OP_MUL = 6, B200_OP_NORM = 18, B200_OP_RMS_NORM = 19,
    B200_OP_MUL_MAT = 21, B200_OP_SCALE = 23, B200_OP_CPY = 25, B200_OP_CONT = 26, B200_OP_RESHAPE = 27,
    B200_OP_VIEW = 28, B200_OP_PERMUTE = 29, B200_OP_TRANSPOSE = 30, B200_OP_GET_ROWS = 31,
    B200_OP_DIAG_MASK_INF = 34, B200_OP_SOFT_MAX = 36, B200_OP_ROPE = 38, B200_OP_UNARY = 51,
};
enum b200_ggml_unary_op { B200_UNARY_GELU = 7, B200_UNARY_SILU = 9 };     /* LC/ggml.h:388-399 */
enum b200_ggml_task_type { B200_TASK_INIT = 0, B200_TASK_COMPUTE = 1, B200_TASK_FINALIZE = 2 };   /* LC/ggml.h:514-519 */

struct ggml_tensor {             /* LC/ggml.h:395-431; 272 bytes */
    int32_t  type;               /* enum ggml_type */
    int32_t  backend;            /* enum ggml_backend */
    int32_t  n_dims;
    int64_t  ne[4];              /* elements per dimension, ne[0] fastest */
    size_t   nb[4];              /* strides in bytes */
    int32_t  op;                 /* enum ggml_op */
    int32_t  op_params[8];
    bool     is_param;
    struct ggml_tensor *grad;
    struct ggml_tensor *src[6];
    int32_t  perf_runs;
    int64_t  perf_cycles;
    int64_t  perf_time_us;
    void    *data;               /* host pointer */
    char     name[48];
    void    *extra;              /* backend-private: this library hangs its device buffers here */
    char     padding[4];
};

struct ggml_compute_params {     /* LC/ggml.h:521-530 */
    int32_t type;                /* enum ggml_task_type */
    int32_t ith, nth;            /* every worker thread calls the seam; only ith == 0 && COMPUTE does work */
    size_t  wsize;
    void   *wdata;
};

#endif /* GGML_B200_USE_GGML_H */

#define GGML_CUDA_MAX_DEVICES 16                                            /* LC/ggml-cuda.h:9 */

/* ---- called by ggml.c itself --------------------------------------------------------------------------- */
void   ggml_init_cublas(void);                                              /* LC/ggml-cuda.h:11; cuda.rs:7-9 */
bool   ggml_cuda_can_mul_mat(const struct ggml_tensor *src0, const struct ggml_tensor *src1, struct ggml_tensor *dst); /* :15; cuda.rs:16-22 */
bool   ggml_cuda_compute_forward(struct ggml_compute_params *params, struct ggml_tensor *tensor);                    /* :31; cuda.rs:71-76 */

/* ---- called by crates/ggml (accelerator/mod.rs, tensor.rs) ------------------------------------------------ */
void   ggml_cuda_set_tensor_split(const float *tensor_split);              /* :12; cuda.rs:10-12 (reads ONE float: the Rust side passes &1.0f32) */
void   ggml_cuda_set_main_device(int main_device);                         /* :28; cuda.rs:56-58 */
void   ggml_cuda_set_scratch_size(size_t scratch_size);                    /* :29; cuda.rs:62-64 */
void   ggml_cuda_free_scratch(void);                                       /* :30; cuda.rs:65-67 */
void   ggml_cuda_transform_tensor(void *data, struct ggml_tensor *tensor); /* :22; cuda.rs:43-45 : upload a weight, set tensor->extra */
void   ggml_cuda_free_data(struct ggml_tensor *tensor);                    /* :24; cuda.rs:46-48 */
void   ggml_cuda_assign_buffers(struct ggml_tensor *tensor);               /* :25; cuda.rs:49-51 : activations, scratch bump allocator */
void   ggml_cuda_assign_buffers_no_scratch(struct ggml_tensor *tensor);    /* :26; cuda.rs:52-54 : KV cache, own zeroed allocation */
void   ggml_cuda_assign_buffers_force_inplace(struct ggml_tensor *tensor); /* :27; cuda.rs:55 (bound, unused by Rust) */

/* ---- bound by cuda.rs but unused by the Rust side (kept so the binding links) ---------- */
void   ggml_cuda_mul(const struct ggml_tensor *src0, const struct ggml_tensor *src1, struct ggml_tensor *dst);       /* :14 */
size_t ggml_cuda_mul_mat_get_wsize(const struct ggml_tensor *src0, const struct ggml_tensor *src1, struct ggml_tensor *dst); /* :16 (stale) */
void   ggml_cuda_mul_mat(const struct ggml_tensor *src0, const struct ggml_tensor *src1, struct ggml_tensor *dst, void *wdata, size_t wsize); /* :17 */
void  *ggml_cuda_host_malloc(size_t size);                                 /* :20 : pinned host memory */
void   ggml_cuda_host_free(void *ptr);                                     /* :21 */
void   ggml_cuda_set_mul_mat_q(bool mul_mat_q);                            /* cuda.rs:59-61 (stale upstream knob; integer mat-mul is always on here) */

#ifdef __cplusplus
}
#endif
#endif
