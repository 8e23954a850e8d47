// llm_b200/csrc/decode.h -- parameters of the one-launch-per-token decode kernel (decode.cu)
#pragma once
#include "kernels.cuh"

namespace b200 {

struct DecodeLayer {
    QWeight wqkv, wo, w13, w2;
    const float *attn_norm, *ffn_norm;
    __half *K, *V;                      // this layer's slice of memory_k ([n_ctx][gqa]) / memory_v ([gqa][n_ctx], transposed)
};

struct DecodeParams {
    const DecodeLayer *layers;
    int n_layer;
    QWeight wte, output;
    const float *norm;
    int e, f, hd, gqa, n_head, n_head_kv, n_ctx, n_vocab;
    float kq_scale, eps;
    const float2 *rope_cs; int rope_half;
    const uint16_t *lut_silu, *lut_exp;
    const int32_t *token;
    int *n_past;                        // device copy of InferenceSession::n_past; incremented at the end of the kernel
    float *x, *q, *kq, *attn, *ff, *h13, *logits;
    unsigned int *bar;                  // [0] arrival count, [1] generation
};

bool decode_supported(const DecodeParams &P, int wtype);
// cooperative launch on `st`; returns false if the kernel cannot be made resident (caller falls back to the per-op schedule)
bool launch_decode(const DecodeParams &P, int wtype, cudaStream_t st, int *grid_out);

}  // namespace b200
