"""ctypes binding of the ggml_cuda_* seam (include/ggml_b200.h) -- the stub a ctypes host would write, and the way the
per-op parity tests call the product through its plugin ABI.

Mirrors the slice of crates/ggml the hot path uses (crates/ggml/src/context.rs:276-626 op_* constructors,
crates/ggml/src/tensor.rs:56-112 transfer_to / offload): tensors are `struct ggml_tensor` (272 bytes, LC/ggml.h:395-431)
built here exactly as the ggml constructors fill them (shapes, strides, op, op_params, src), then handed node by node to
ggml_cuda_compute_forward the way ggml_graph_compute_thread does (LC/ggml.c:14584-14591).  No arithmetic happens in Python.
"""
import ctypes as C
import struct

import numpy as np

from . import _lib

F32, F16, Q4_0, Q4_1, Q5_0, Q5_1, Q8_0, Q8_1, I8, I16, I32 = 0, 1, 2, 3, 6, 7, 8, 9, 16, 17, 18
BACKEND_CPU, BACKEND_GPU = 0, 10
OP_NONE, OP_DUP, OP_ADD, OP_MUL, OP_NORM, OP_RMS_NORM, OP_MUL_MAT, OP_SCALE, OP_CPY, OP_CONT = 0, 1, 2, 6, 18, 19, 21, 23, 25, 26
OP_RESHAPE, OP_VIEW, OP_PERMUTE, OP_TRANSPOSE, OP_GET_ROWS, OP_DIAG_MASK_INF, OP_SOFT_MAX, OP_ROPE, OP_UNARY = 27, 28, 29, 30, 31, 34, 36, 38, 51
UNARY_GELU, UNARY_SILU = 7, 9
TASK_INIT, TASK_COMPUTE, TASK_FINALIZE = 0, 1, 2
BLOCK_BYTES = {Q4_0: 18, Q4_1: 20, Q5_0: 22, Q5_1: 24, Q8_0: 34, Q8_1: 40}
Q2_K, Q3_K, Q4_K, Q5_K, Q6_K = 10, 11, 12, 13, 14                  # K-quants served by the seam (256-element super-blocks, LC/k_quants.h:28-110)
SUPER_BLOCK_BYTES = {Q2_K: 84, Q3_K: 110, Q4_K: 144, Q5_K: 176, Q6_K: 210}
TYPE_SIZE = {F32: 4, F16: 2, I8: 1, I16: 2, I32: 4, **BLOCK_BYTES, **SUPER_BLOCK_BYTES}
BLCK = {**{t: 32 for t in BLOCK_BYTES}, **{t: 256 for t in SUPER_BLOCK_BYTES}}


class Tensor(C.Structure):
    pass


Tensor._fields_ = [
    ("type", C.c_int32), ("backend", C.c_int32), ("n_dims", C.c_int32),
    ("ne", C.c_int64 * 4), ("nb", C.c_size_t * 4),
    ("op", C.c_int32), ("op_params", C.c_int32 * 8), ("is_param", C.c_bool),
    ("grad", C.POINTER(Tensor)), ("src", C.POINTER(Tensor) * 6),
    ("perf_runs", C.c_int32), ("perf_cycles", C.c_int64), ("perf_time_us", C.c_int64),
    ("data", C.c_void_p), ("name", C.c_char * 48), ("extra", C.c_void_p), ("padding", C.c_char * 4),
]


class ComputeParams(C.Structure):
    _fields_ = [("type", C.c_int32), ("ith", C.c_int32), ("nth", C.c_int32), ("wsize", C.c_size_t), ("wdata", C.c_void_p)]


def _seam():
    L = _lib.lib()
    if not getattr(L, "_seam_ready", False):
        tp = C.POINTER(Tensor)
        L.ggml_init_cublas.argtypes = []
        L.ggml_cuda_set_main_device.argtypes = [C.c_int]
        L.ggml_cuda_set_tensor_split.argtypes = [C.POINTER(C.c_float)]
        L.ggml_cuda_set_scratch_size.argtypes = [C.c_size_t]
        L.ggml_cuda_free_scratch.argtypes = []
        L.ggml_cuda_transform_tensor.argtypes = [C.c_void_p, tp]
        L.ggml_cuda_free_data.argtypes = [tp]
        L.ggml_cuda_assign_buffers.argtypes = [tp]
        L.ggml_cuda_assign_buffers_no_scratch.argtypes = [tp]
        L.ggml_cuda_assign_buffers_force_inplace.argtypes = [tp]
        L.ggml_cuda_can_mul_mat.restype = C.c_bool
        L.ggml_cuda_can_mul_mat.argtypes = [tp, tp, tp]
        L.ggml_cuda_compute_forward.restype = C.c_bool
        L.ggml_cuda_compute_forward.argtypes = [C.POINTER(ComputeParams), tp]
        L.ggml_cuda_host_malloc.restype = C.c_void_p
        L.ggml_cuda_host_malloc.argtypes = [C.c_size_t]
        L.ggml_cuda_host_free.argtypes = [C.c_void_p]
        L._seam_ready = True
    return L


class Context:
    """Owns tensors (and their host buffers) for one test graph; ops are executed eagerly node by node through the seam."""

    def __init__(self, device=0, scratch_mb=64):
        self.L = _seam()
        self.L.ggml_cuda_set_main_device(device)
        self.L.ggml_init_cublas()                       # accelerator::initialize, crates/ggml/src/accelerator/mod.rs:68-77
        one = C.c_float(1.0)
        self.L.ggml_cuda_set_tensor_split(C.byref(one))
        self.L.ggml_cuda_set_scratch_size(scratch_mb << 20)
        self._keep = []
        self._owned = []

    # ---- tensors -------------------------------------------------------------------------------------------------
    def new_tensor(self, typ, ne, array=None):
        ne = list(ne) + [1] * (4 - len(ne))
        t = Tensor()
        t.type, t.backend, t.n_dims = typ, BACKEND_CPU, max(1, sum(1 for i, v in enumerate(ne) if v != 1 or i == 0))
        ts, bs = TYPE_SIZE[typ], BLCK.get(typ, 1)
        t.ne[:] = ne
        t.nb[0] = ts
        t.nb[1] = ts * (ne[0] // bs)                    # LC/ggml.c:4623-4627
        t.nb[2] = t.nb[1] * ne[1]
        t.nb[3] = t.nb[2] * ne[2]
        nbytes = t.nb[3] * ne[3]
        if array is None:
            array = np.zeros(nbytes, np.uint8)
        else:
            array = np.ascontiguousarray(array)
            assert array.nbytes == nbytes, (array.nbytes, nbytes)
        t.data = array.ctypes.data
        self._keep.append((t, array))
        return t

    def from_numpy(self, a):
        """f32 array of shape [..., ne1, ne0] (C order) -> tensor with ne = reversed shape."""
        a = np.ascontiguousarray(a, np.float32)
        return self.new_tensor(F32, list(reversed(a.shape)), a)

    def quantized(self, typ, rows_bytes, K):
        """uint8 [N, K/32*blk] GGML block rows -> 2-D weight tensor [K, N]."""
        rows_bytes = np.ascontiguousarray(rows_bytes, np.uint8)
        return self.new_tensor(typ, [K, rows_bytes.shape[0]], rows_bytes)

    def host_array(self, t, dtype=np.float32):
        for tt, arr in self._keep:
            if tt is t:
                shape = [int(t.ne[i]) for i in (3, 2, 1, 0)]
                while len(shape) > 1 and shape[0] == 1:
                    shape.pop(0)
                return arr.view(dtype).reshape(shape)
        raise KeyError

    def transfer_to_gpu(self, t):                       # Tensor::transfer_to(Backend::Gpu), crates/ggml/src/tensor.rs:56-80
        t.backend = BACKEND_GPU
        self.L.ggml_cuda_transform_tensor(t.data, C.byref(t))
        self._owned.append(t)
        return t

    def offload(self, t):                               # Tensor::offload, tensor.rs:87-94
        self.L.ggml_cuda_assign_buffers(C.byref(t))
        return t

    def offload_no_scratch(self, t):                    # tensor.rs:101-112
        self.L.ggml_cuda_assign_buffers_no_scratch(C.byref(t))
        self._owned.append(t)
        return t

    # ---- op constructors (shape rules of the ggml_* builders) ----------------------------------------------------------
    def _result(self, typ, ne, op, src0, src1=None, params=()):
        t = self.new_tensor(typ, ne)
        t.op = op
        t.src[0] = C.pointer(src0)
        if src1 is not None:
            t.src[1] = C.pointer(src1)
        for i, v in enumerate(params):
            t.op_params[i] = v
        return t

    @staticmethod
    def _f2i(x):
        return struct.unpack("<i", struct.pack("<f", x))[0]

    def op_mul_mat(self, a, b):                         # LC/ggml.c:5846-5868
        return self._result(F32, [a.ne[1], b.ne[1], b.ne[2], b.ne[3]], OP_MUL_MAT, a, b)

    def op_rms_norm(self, a, eps=5e-6):                 # crates/ggml/src/context.rs:296-300, LLAMA_DEFAULT_RMS_EPS
        return self._result(F32, list(a.ne), OP_RMS_NORM, a, params=(self._f2i(eps),))

    def op_norm(self, a):
        return self._result(F32, list(a.ne), OP_NORM, a)

    def op_soft_max(self, a):
        return self._result(F32, list(a.ne), OP_SOFT_MAX, a)

    def op_diag_mask_inf(self, a, n_past):
        return self._result(F32, list(a.ne), OP_DIAG_MASK_INF, a, params=(n_past, 0))

    def op_scale(self, a, s):
        sv = self.from_numpy(np.array([s], np.float32))  # new_f32: 1-element host tensor
        return self._result(F32, list(a.ne), OP_SCALE, a, sv)

    def op_silu(self, a):
        return self._result(F32, list(a.ne), OP_UNARY, a, params=(UNARY_SILU,))

    def op_gelu(self, a):
        return self._result(F32, list(a.ne), OP_UNARY, a, params=(UNARY_GELU,))

    def op_add(self, a, b):
        return self._result(F32, list(a.ne), OP_ADD, a, b)

    def op_mul(self, a, b):
        return self._result(F32, list(a.ne), OP_MUL, a, b)

    def op_rope(self, a, n_past, n_dims, mode, freq_base=10000.0, freq_scale=1.0, n_ctx=0):   # context.rs:558-590
        return self._result(F32, list(a.ne), OP_ROPE, a,
                            params=(n_past, n_dims, mode, n_ctx, self._f2i(freq_base), self._f2i(freq_scale)))

    def op_cpy_to(self, a, typ):
        """ggml_cpy(a, new_tensor(typ, a.ne)) with a contiguous destination."""
        dst = self.new_tensor(typ, list(a.ne))
        r = self._result(typ, list(a.ne), OP_CPY, a, dst)
        return r

    # ---- execution -----------------------------------------------------------------------------------------------------
    def compute(self, t, nth=1):
        """What ggml_graph_compute_thread does for one node: every thread calls the seam; only ith==0/COMPUTE works."""
        handled = None
        for typ in (TASK_INIT, TASK_COMPUTE, TASK_FINALIZE):
            for ith in range(nth):
                p = ComputeParams(type=typ, ith=ith, nth=nth, wsize=0, wdata=None)
                h = self.L.ggml_cuda_compute_forward(C.byref(p), C.byref(t))
                handled = h if handled is None else (handled and h)
        return bool(handled)

    def close(self):
        for t in self._owned:
            self.L.ggml_cuda_free_data(C.byref(t))
        self._owned = []
        self.L.ggml_cuda_free_scratch()
