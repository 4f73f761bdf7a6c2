"""ctypes binding of include/lstmp_b200.h.  Thin: every method is one C-ABI call."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OK, EINVAL, ENODEV, ENOMEM, ESTATE, EUNSUPPORTED = 0, -1, -2, -3, -4, -5

# every symbol include/lstmp_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "lstmp_b200_abi_version", "lstmp_b200_last_error", "lstmp_b200_create", "lstmp_b200_destroy",
    "lstmp_b200_clone", "lstmp_b200_num_params", "lstmp_b200_set_params", "lstmp_b200_get_params",
    "lstmp_b200_get_flat", "lstmp_b200_set_flat", "lstmp_b200_arena", "lstmp_b200_get_state",
    "lstmp_b200_set_state", "lstmp_b200_reset", "lstmp_b200_propagate", "lstmp_b200_backpropagate",
    "lstmp_b200_update", "lstmp_b200_allreduce_grads_nccl", "lstmp_b200_get_info", "lstmp_b200_get_record",
    "lstmp_b200_timing_enable", "lstmp_b200_timing_read", "lstmp_b200_debug_gemm", "lstmp_b200_debug_gemm_group",
    "lstmp_b200_xent_create", "lstmp_b200_xent_destroy", "lstmp_b200_xent_eval_masked",
    "lstmp_b200_xent_get_stats", "lstmp_b200_xent_reset_stats",
    "lstmp_b200_update_clipped", "lstmp_b200_time_shift", "lstmp_b200_set_nccl",
    "lstmp_b200_dispatch_create", "lstmp_b200_dispatch_destroy", "lstmp_b200_dispatch_set_transform",
    "lstmp_b200_dispatch_load_utt", "lstmp_b200_dispatch_assemble", "lstmp_b200_dispatch_get_stats",
    "lstmp_b200_xent_eval_masked_logits",
    "lstmp_b200_tail_create", "lstmp_b200_tail_destroy", "lstmp_b200_tail_arena", "lstmp_b200_tail_set_flat",
    "lstmp_b200_tail_get_flat", "lstmp_b200_tail_propagate_eval", "lstmp_b200_tail_backpropagate",
    "lstmp_b200_tail_update", "lstmp_b200_tail_allreduce_grads_nccl", "lstmp_b200_tail_get_diff",
    "lstmp_b200_tail_get_stats", "lstmp_b200_tail_reset_stats",
]

TIMING_KINDS = ["input_gemm", "fwd_recurrent", "bwd_recurrent", "in_diff_gemm", "wgrad_gemms", "small_grads",
                "update", "reset"]


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("lstmp_b200 error %d: %s" % (code, msg))
        self.code = code


class Info(ctypes.Structure):
    _fields_ = [
        ("input_dim", ctypes.c_int), ("cell_dim", ctypes.c_int), ("recur_dim", ctypes.c_int),
        ("num_stream", ctypes.c_int), ("max_frames", ctypes.c_int), ("sm_count", ctypes.c_int),
        ("ngroups", ctypes.c_int), ("ctas_per_group", ctypes.c_int), ("streams_per_group", ctypes.c_int),
        ("cells_per_cta", ctypes.c_int), ("rcols_per_cta", ctypes.c_int),
        ("fwd_smem_bytes", ctypes.c_size_t), ("bwd_smem_bytes", ctypes.c_size_t),
        ("workspace_bytes", ctypes.c_size_t), ("kernel_launches", ctypes.c_ulonglong),
        ("gemm_backend", ctypes.c_int), ("weights_streamed", ctypes.c_int), ("fwd_tensor_core", ctypes.c_int),
        ("bwd_tensor_core", ctypes.c_int), ("bwd_ctas", ctypes.c_int), ("bwd_cluster", ctypes.c_int),
    ]


class XentStats(ctypes.Structure):
    _fields_ = [("loss", ctypes.c_double), ("entropy", ctypes.c_double), ("correct", ctypes.c_longlong),
                ("frames", ctypes.c_longlong), ("kernel_launches", ctypes.c_ulonglong)]


class Timing(ctypes.Structure):
    _fields_ = [("ms", ctypes.c_double * 8), ("count", ctypes.c_ulonglong * 8)]


def lib_path():
    # LSTMP_B200_LIB selects an alternative build of the same library (e.g. the clock-stamp build)
    return os.environ.get("LSTMP_B200_LIB") or os.path.join(_HERE, "_lib", "liblstmp_b200.so")


def load_library():
    """Load the CUDA engine.  Fails loudly when it has not been built -- there is no other path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise EngineError(ENODEV, "%s not built; run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
    L = ctypes.CDLL(path)
    vp, sz, fp, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float, ctypes.c_int
    L.lstmp_b200_abi_version.restype = ci
    L.lstmp_b200_last_error.restype = ctypes.c_char_p
    L.lstmp_b200_create.argtypes = [ci, ci, ci, ci, ci, ci, ctypes.POINTER(vp)]
    L.lstmp_b200_destroy.argtypes = [vp]
    L.lstmp_b200_clone.argtypes = [vp, ctypes.POINTER(vp)]
    L.lstmp_b200_num_params.argtypes = [vp, ctypes.POINTER(sz)]
    L.lstmp_b200_set_params.argtypes = [vp, vp, sz, vp, sz, vp, vp, vp, vp, vp, sz, vp]
    L.lstmp_b200_get_params.argtypes = [vp, vp, sz, vp, sz, vp, vp, vp, vp, vp, sz, vp]
    L.lstmp_b200_get_flat.argtypes = [vp, ci, vp, vp]
    L.lstmp_b200_set_flat.argtypes = [vp, ci, vp, vp]
    L.lstmp_b200_arena.argtypes = [vp, ci, ctypes.POINTER(vp), ctypes.POINTER(sz)]
    L.lstmp_b200_get_state.argtypes = [vp, vp, sz, vp, sz, vp]
    L.lstmp_b200_set_state.argtypes = [vp, vp, sz, vp, sz, vp]
    L.lstmp_b200_reset.argtypes = [vp, vp, ci, vp]
    L.lstmp_b200_propagate.argtypes = [vp, vp, sz, vp, sz, ci, vp]
    L.lstmp_b200_backpropagate.argtypes = [vp, vp, sz, vp, sz, vp, sz, ci, vp]
    L.lstmp_b200_update.argtypes = [vp, fp, fp, vp]
    L.lstmp_b200_allreduce_grads_nccl.argtypes = [vp, vp, vp]
    L.lstmp_b200_update_clipped.argtypes = [vp, fp, fp, fp, vp]
    L.lstmp_b200_set_nccl.argtypes = [vp, vp, vp]
    L.lstmp_b200_time_shift.argtypes = [vp, sz, vp, sz, ci, ci, ci, vp]
    L.lstmp_b200_get_info.argtypes = [vp, ctypes.POINTER(Info)]
    L.lstmp_b200_get_record.argtypes = [vp, ci, vp, sz, vp]
    L.lstmp_b200_debug_gemm.argtypes = [ci, vp, sz, ci, ci, ci, fp, vp, sz, ci, vp, sz, ci, fp, vp, vp]
    L.lstmp_b200_timing_enable.argtypes = [vp, ci]
    L.lstmp_b200_timing_read.argtypes = [vp, ctypes.POINTER(Timing)]
    L.lstmp_b200_xent_create.argtypes = [ci, ci, ctypes.POINTER(vp)]
    L.lstmp_b200_xent_destroy.argtypes = [vp]
    L.lstmp_b200_xent_eval_masked.argtypes = [vp, vp, vp, sz, ci, ci, vp, vp, vp, vp, sz, vp]
    L.lstmp_b200_xent_get_stats.argtypes = [vp, ctypes.POINTER(XentStats), vp]
    L.lstmp_b200_xent_reset_stats.argtypes = [vp, vp]
    for name in ABI_SYMBOLS:
        f = getattr(L, name)
        if name not in ("lstmp_b200_last_error",):
            f.restype = ci
    _LIB = L
    return L


def _chk(rc):
    if rc != 0:
        raise EngineError(rc, load_library().lstmp_b200_last_error().decode("utf-8", "replace"))


class _CudaArray:
    """__cuda_array_interface__ view of an engine-owned device arena (zero copy into torch)."""

    def __init__(self, ptr, n, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {
            "shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None,
        }


def _cur_stream(device):
    """The current torch stream of `device` (an engine's own device, not whatever device happens to be current)."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    """One LstmProjectedStreams layer resident on one B200."""

    def __init__(self, input_dim, cell_dim, recur_dim, num_stream, max_frames, device=0):
        L = load_library()
        h = ctypes.c_void_p()
        _chk(L.lstmp_b200_create(input_dim, cell_dim, recur_dim, num_stream, max_frames, device, ctypes.byref(h)))
        self._h = h
        self.I, self.C, self.R, self.S, self.Tmax = input_dim, cell_dim, recur_dim, num_stream, max_frames
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            load_library().lstmp_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return _cur_stream(self.device)

    @property
    def num_params(self):
        n = ctypes.c_size_t()
        _chk(load_library().lstmp_b200_num_params(self._h, ctypes.byref(n)))
        return n.value

    def info(self):
        i = Info()
        _chk(load_library().lstmp_b200_get_info(self._h, ctypes.byref(i)))
        return {k: getattr(i, k) for k, _ in Info._fields_}

    # ---- flat arenas (GetParams order) -------------------------------------------------
    def set_flat(self, which, tensor_or_array):
        import numpy as np
        import torch
        t = tensor_or_array
        if isinstance(t, np.ndarray):
            t = np.ascontiguousarray(t, dtype=np.float32)
            assert t.size == self.num_params
            ptr = t.ctypes.data
        else:
            t = t.contiguous().float()
            assert t.numel() == self.num_params
            ptr = t.data_ptr()
        _chk(load_library().lstmp_b200_set_flat(self._h, which, ctypes.c_void_p(ptr), self._stream()))

    def get_flat(self, which):
        import numpy as np
        out = np.empty(self.num_params, np.float32)
        _chk(load_library().lstmp_b200_get_flat(self._h, which, ctypes.c_void_p(out.ctypes.data), self._stream()))
        return out

    def arena_tensor(self, which):
        """Zero-copy torch view of arena `which` (0 params, 1 momentum-accumulated, 2 fresh grads)."""
        import torch
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _chk(load_library().lstmp_b200_arena(self._h, which, ctypes.byref(p), ctypes.byref(n)))
        return torch.as_tensor(_CudaArray(p.value, n.value, self), device="cuda:%d" % self.device)

    # ---- carried state -------------------------------------------------------------------
    def get_state(self):
        import numpy as np
        c = np.empty((self.S, self.C), np.float32)
        r = np.empty((self.S, self.R), np.float32)
        _chk(load_library().lstmp_b200_get_state(self._h, ctypes.c_void_p(c.ctypes.data), self.C,
                                                 ctypes.c_void_p(r.ctypes.data), self.R, self._stream()))
        return c, r

    def set_state(self, c, r):
        import numpy as np
        c = np.ascontiguousarray(c, np.float32)
        r = np.ascontiguousarray(r, np.float32)
        assert c.shape == (self.S, self.C) and r.shape == (self.S, self.R)
        _chk(load_library().lstmp_b200_set_state(self._h, ctypes.c_void_p(c.ctypes.data), self.C,
                                                 ctypes.c_void_p(r.ctypes.data), self.R, self._stream()))

    def reset(self, flags):
        import numpy as np
        f = np.ascontiguousarray(flags, np.int32)
        _chk(load_library().lstmp_b200_reset(self._h, ctypes.c_void_p(f.ctypes.data), int(f.size), self._stream()))

    # ---- hot path: device tensors (torch) -------------------------------------------------
    @staticmethod
    def _mat(t, cols, name):
        if not t.is_cuda or t.dtype.__str__() != "torch.float32" or t.dim() != 2 or t.shape[1] != cols or (
                t.shape[0] > 1 and t.stride(1) != 1):
            raise EngineError(EINVAL, "%s must be a CUDA float32 [rows x %d] matrix with unit column stride" % (name, cols))
        return ctypes.c_void_p(t.data_ptr()), (t.stride(0) if t.shape[0] > 1 else max(cols, t.stride(0)))

    def propagate(self, x, out):
        px, ldx = self._mat(x, self.I, "in")
        po, ldo = self._mat(out, self.R, "out")
        if out.shape[0] != x.shape[0]:
            raise EngineError(EINVAL, "out rows != in rows")
        _chk(load_library().lstmp_b200_propagate(self._h, px, ldx, po, ldo, x.shape[0], self._stream()))
        self._rows = x.shape[0]

    def backpropagate(self, x, out_diff, in_diff=None):
        px, ldx = self._mat(x, self.I, "in")
        pod, ldod = self._mat(out_diff, self.R, "out_diff")
        if in_diff is not None:
            pid, ldid = self._mat(in_diff, self.I, "in_diff")
        else:
            pid, ldid = None, 0
        _chk(load_library().lstmp_b200_backpropagate(self._h, px, ldx, pod, ldod, pid, ldid, x.shape[0], self._stream()))

    def update(self, learn_rate, momentum):
        _chk(load_library().lstmp_b200_update(self._h, float(learn_rate), float(momentum), self._stream()))

    def update_clipped(self, learn_rate, momentum, max_grad):
        """Update with the element-wise clip of the standard single-stream component (nnet-lstm-projected.h:480-493)."""
        _chk(load_library().lstmp_b200_update_clipped(self._h, float(learn_rate), float(momentum), float(max_grad),
                                                      self._stream()))

    def set_nccl(self, comm_ptr, stream_ptr):
        """Overlapped exchange inside backpropagate (lstmp_b200_set_nccl); comm_ptr None switches it off."""
        _chk(load_library().lstmp_b200_set_nccl(self._h, ctypes.c_void_p(comm_ptr) if comm_ptr else None,
                                                ctypes.c_void_p(stream_ptr) if stream_ptr else None))

    def allreduce_grads_nccl(self, comm_ptr, stream_ptr=None):
        """Sum all-reduce of the fresh-gradient arena over the raw ncclComm_t `comm_ptr` on `stream_ptr` (a
        cudaStream_t as int; default: the current stream)."""
        st = ctypes.c_void_p(stream_ptr) if stream_ptr is not None else self._stream()
        _chk(load_library().lstmp_b200_allreduce_grads_nccl(self._h, ctypes.c_void_p(comm_ptr), st))

    def timing_enable(self, on=True):
        _chk(load_library().lstmp_b200_timing_enable(self._h, 1 if on else 0))

    def timing_read(self):
        """{kind: (total_ms, launches)} since the last read (device-side CUDA events)."""
        t = Timing()
        _chk(load_library().lstmp_b200_timing_read(self._h, ctypes.byref(t)))
        return {k: (t.ms[i], int(t.count[i])) for i, k in enumerate(TIMING_KINDS)}

    def get_record(self, backward=False):
        import numpy as np
        T = self._last_rows() // self.S
        W = 7 * self.C + self.R
        out = np.empty((T * self.S, W), np.float32)
        _chk(load_library().lstmp_b200_get_record(self._h, 1 if backward else 0, ctypes.c_void_p(out.ctypes.data), W,
                                                  self._stream()))
        return out

    def _last_rows(self):
        return getattr(self, "_rows", self.Tmax * self.S)


def debug_gemm(backend, C, M, N, K, alpha, A, tA, B, tB, beta=0.0, bias=None):
    """C = alpha*op(A)*op(B) + beta*C (+bias) with one of the engine's GEMM kernels (torch CUDA tensors)."""
    import torch
    L = load_library()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _chk(L.lstmp_b200_debug_gemm(int(backend), ctypes.c_void_p(C.data_ptr()), C.stride(0), M, N, K, float(alpha),
                                 ctypes.c_void_p(A.data_ptr()), A.stride(0), int(tA), ctypes.c_void_p(B.data_ptr()),
                                 B.stride(0), int(tB), float(beta),
                                 ctypes.c_void_p(bias.data_ptr()) if bias is not None else None, st))


class GemmDesc(ctypes.Structure):
    """lstmp_b200_gemm_desc (include/lstmp_b200.h)."""
    _fields_ = [("C", ctypes.c_void_p), ("ldc", ctypes.c_size_t), ("M", ctypes.c_int), ("N", ctypes.c_int),
                ("K", ctypes.c_int), ("alpha", ctypes.c_float), ("A", ctypes.c_void_p), ("lda", ctypes.c_size_t),
                ("tA", ctypes.c_int), ("B", ctypes.c_void_p), ("ldb", ctypes.c_size_t), ("tB", ctypes.c_int),
                ("beta", ctypes.c_float), ("bias", ctypes.c_void_p)]


def debug_gemm_group(problems):
    """problems: list of (C, M, N, K, alpha, A, tA, B, tB, beta, bias) as for debug_gemm; one grouped launch
    (lstmp_b200_debug_gemm_group).  Returns the number of kernel launches."""
    import torch
    L = load_library()
    L.lstmp_b200_debug_gemm_group.argtypes = [ctypes.c_int, ctypes.POINTER(GemmDesc), ctypes.c_void_p]
    L.lstmp_b200_debug_gemm_group.restype = ctypes.c_int
    arr = (GemmDesc * len(problems))()
    for d, (C, M, N, K, alpha, A, tA, B, tB, beta, bias) in zip(arr, problems):
        d.C, d.ldc, d.M, d.N, d.K, d.alpha = C.data_ptr(), C.stride(0), M, N, K, float(alpha)
        d.A, d.lda, d.tA, d.B, d.ldb, d.tB = A.data_ptr(), A.stride(0), int(tA), B.data_ptr(), B.stride(0), int(tB)
        d.beta, d.bias = float(beta), (bias.data_ptr() if bias is not None else None)
    rc = L.lstmp_b200_debug_gemm_group(len(problems), arr, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc < 0:
        _chk(rc)
    return rc


def time_shift(x, out, shift):
    """out row dst = x row clamp(dst + shift, 0, rows - 1) on the device (TimeShift::PropagateFnc)."""
    px, ldx = Engine._mat(x, x.shape[1], "in")
    po, ldo = Engine._mat(out, x.shape[1], "out")
    if out.shape[0] != x.shape[0]:
        raise EngineError(EINVAL, "out rows != in rows")
    _chk(load_library().lstmp_b200_time_shift(px, ldx, po, ldo, x.shape[0], x.shape[1], int(shift),
                                              _cur_stream(x.device.index or 0)))


class XentEngine:
    """Device side of the masked cross-entropy with sparse targets (lstmp_b200_xent_*)."""

    def __init__(self, max_frames, device=0):
        L = load_library()
        h = ctypes.c_void_p()
        _chk(L.lstmp_b200_xent_create(int(max_frames), device, ctypes.byref(h)))
        self._h = h
        self.max_frames, self.device = int(max_frames), device

    def close(self):
        if getattr(self, "_h", None):
            load_library().lstmp_b200_xent_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def eval_masked(self, frame_mask_host, net_out, row_ptr, pdf, weight, diff):
        """frame_mask_host / row_ptr / pdf / weight: host numpy arrays (CSR posterior); net_out, diff: CUDA tensors."""
        import numpy as np
        rows, num_pdf = net_out.shape
        mask = np.ascontiguousarray(frame_mask_host, np.float32)
        rp = np.ascontiguousarray(row_ptr, np.int32)
        pd = np.ascontiguousarray(pdf, np.int32)
        wt = np.ascontiguousarray(weight, np.float32)
        if mask.size != rows or rp.size != rows + 1 or pd.size != wt.size:
            raise EngineError(EINVAL, "mask / posterior sizes do not match the %d frames of net_out" % rows)
        po, ldo = Engine._mat(net_out, num_pdf, "net_out")
        pdiff, ldd = Engine._mat(diff, num_pdf, "diff")
        if diff.shape[0] != rows:
            raise EngineError(EINVAL, "diff rows != net_out rows")
        _chk(load_library().lstmp_b200_xent_eval_masked(
            self._h, ctypes.c_void_p(mask.ctypes.data), po, ldo, rows, num_pdf, ctypes.c_void_p(rp.ctypes.data),
            ctypes.c_void_p(pd.ctypes.data) if pd.size else None, ctypes.c_void_p(wt.ctypes.data) if wt.size else None,
            pdiff, ldd, _cur_stream(self.device)))

    def stats(self):
        st = XentStats()
        _chk(load_library().lstmp_b200_xent_get_stats(self._h, ctypes.byref(st), _cur_stream(self.device)))
        return {"loss": st.loss, "entropy": st.entropy, "correct": int(st.correct), "frames": int(st.frames),
                "kernel_launches": int(st.kernel_launches)}

    def reset_stats(self):
        _chk(load_library().lstmp_b200_xent_reset_stats(self._h, _cur_stream(self.device)))
