"""Loader of the C-ABI product library (llm_b200/libllm_b200.so).

There is no Python/NumPy/CPU implementation of anything in this package: if the CUDA library is missing or no CUDA device
is present, calls fail loudly (ImportError here, exit(1) inside the library -- the reference backend's error behaviour,
LC/ggml-cuda.cu:24-53)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libllm_b200.so")

_lib = None


class LlamaHparams(C.Structure):
    """b200_llama_hparams (include/llm_b200.h) <- Hyperparameters, crates/models/llama/src/lib.rs:403-447."""
    _fields_ = [("n_vocab", C.c_int32), ("n_embd", C.c_int32), ("n_head", C.c_int32), ("n_head_kv", C.c_int32),
                ("n_layer", C.c_int32), ("n_rot", C.c_int32), ("n_ff", C.c_int32), ("wtype", C.c_int32),
                ("context_size", C.c_int32), ("rope_freq_base", C.c_float), ("rope_freq_scale", C.c_float)]


class GgmlTensorInfo(C.Structure):
    """b200_ggml_tensor_info <- TensorLoadInfo, crates/ggml/src/format/loader.rs:72-86."""
    _fields_ = [("name", C.c_char * 96), ("type", C.c_int32), ("n_dims", C.c_int32), ("ne", C.c_int64 * 2), ("offset", C.c_uint64), ("nbytes", C.c_uint64)]


class SessionConfig(C.Structure):
    """b200_session_config <- InferenceSessionConfig, crates/llm-base/src/inference_session.rs:799-841."""
    _fields_ = [("n_batch", C.c_int32), ("flags", C.c_int32)]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). llm_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    L.b200_init.argtypes = [C.c_int]
    L.b200_device_info.argtypes = [C.POINTER(i32), C.POINTER(sz), C.POINTER(sz)]
    L.b200_stream.restype = vp
    L.b200_llama_new.restype = vp
    L.b200_llama_new.argtypes = [C.POINTER(LlamaHparams)]
    L.b200_llama_new_tp.restype = vp
    L.b200_llama_new_tp.argtypes = [C.POINTER(LlamaHparams), i32, i32]
    L.b200_session_tp_handle.argtypes = [vp, vp]
    L.b200_session_tp_connect.argtypes = [vp, C.c_char_p]
    L.b200_session_tp_set_nowait.argtypes = [vp, i32]
    L.b200_session_tp_timeouts.restype = i32
    L.b200_session_tp_timeouts.argtypes = [vp]
    L.b200_model_load_tensor.argtypes = [vp, C.c_char_p, i32, vp, sz]
    L.b200_model_synthesize.argtypes = [vp, C.c_uint64]
    L.b200_model_read_tensor.argtypes = [vp, C.c_char_p, vp, sz]
    L.b200_model_tensor_nbytes.restype = sz
    L.b200_model_tensor_nbytes.argtypes = [vp, C.c_char_p]
    L.b200_model_weight_bytes.restype = sz
    L.b200_model_weight_bytes.argtypes = [vp]
    L.b200_model_free.argtypes = [vp]
    L.b200_model_start_session.restype = vp
    L.b200_model_start_session.argtypes = [vp, C.POINTER(SessionConfig)]
    L.b200_session_evaluate.argtypes = [vp, vp, i32, vp, i32]
    L.b200_session_feed_prompt.argtypes = [vp, vp, i32, vp]
    L.b200_session_top_k.argtypes = [vp, i32, vp, vp]
    L.b200_session_evaluate_device.argtypes = [vp, vp, i32]
    L.b200_session_device_logits.restype = vp
    L.b200_session_device_logits.argtypes = [vp]
    L.b200_session_n_past.argtypes = [vp]
    L.b200_session_set_n_past.argtypes = [vp, i32]
    L.b200_session_read_kv.argtypes = [vp, i32, vp, sz]
    L.b200_session_sync.argtypes = [vp]
    L.b200_session_decode_profile.argtypes = [vp, vp]
    L.b200_session_decode_timeline.argtypes = [vp, vp, C.c_int, C.c_int]
    L.b200_session_set_tap.argtypes = [vp, i32, i32]
    L.b200_session_read_tap.restype = i64
    L.b200_session_read_tap.argtypes = [vp, vp, i64]
    L.b200_session_last_launches.argtypes = [vp]
    L.b200_session_free.argtypes = [vp]
    L.b200_timing_begin.argtypes = []
    L.b200_timing_end_ms.restype = C.c_float
    L.b200_session_probe_matvec.restype = C.c_float
    L.b200_session_probe_matvec.argtypes = [vp, i32, C.POINTER(i64), C.POINTER(C.c_double)]
    L.b200_model_is_loaded.argtypes = [vp]
    L.b200_ggml_open.restype = vp
    L.b200_ggml_open.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    L.b200_ggml_open_arch.restype = vp
    L.b200_ggml_open_arch.argtypes = [C.c_char_p, i32, C.POINTER(C.c_int)]
    L.b200_ggml_hparams.argtypes = [vp, C.POINTER(i32), C.POINTER(i32 * 8), C.POINTER(i32)]
    L.b200_ggml_write.argtypes = [C.c_char_p, C.POINTER(i32), i32, i32, vp, vp, vp, C.POINTER(GgmlTensorInfo), vp, i64]
    L.b200_ggml_close.argtypes = [vp]
    L.b200_ggml_container.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.b200_ggml_n_tensors.restype = i64
    L.b200_ggml_n_tensors.argtypes = [vp]
    L.b200_ggml_tensor.argtypes = [vp, i64, C.POINTER(GgmlTensorInfo)]
    L.b200_ggml_tensor_data.restype = vp
    L.b200_ggml_tensor_data.argtypes = [vp, i64]
    L.b200_ggml_n_vocab.restype = i64
    L.b200_ggml_n_vocab.argtypes = [vp]
    L.b200_ggml_token.argtypes = [vp, i64, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
    L.b200_ggml_llama_hparams.argtypes = [vp, C.POINTER(LlamaHparams), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.b200_ggml_write_llama.argtypes = [C.c_char_p, C.POINTER(LlamaHparams), i32, i32, vp, vp, vp, C.POINTER(GgmlTensorInfo), vp, i64]
    L.b200_llama_load_file.restype = vp
    L.b200_llama_load_file.argtypes = [C.c_char_p, i32, C.c_float, C.c_float, C.POINTER(C.c_int)]
    L.b200_op_quantize_act.argtypes = [i32, vp, i64, i64, vp, vp, vp]
    L.b200_op_quantize_weights.argtypes = [i32, vp, i64, i64, vp]
    L.b200_op_mul_mat.argtypes = [i32, vp, i64, i64, vp, i64, vp, i32]
    L.b200_op_quantize_q8_K.argtypes = [vp, i64, i64, vp]
    _lib = L
    return L
