"""ctypes binding of oracle/_ref -- THE REFERENCE ITSELF, compiled from its own sources.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__ and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.

oracle/_ref/libkaldi_lstm_ref.so is built by `make -C oracle ref` on a machine that has the reference tree
(/root/reference): the unmodified headers google/nnet/bd-nnet-lstm-projected-streams.h, google/nnet/nnet-loss.h,
standard/nnet/nnet-lstm-projected.h, standard/nnet/nnet-time-shift.h plus the method bodies of
kaldi-matrix.cc / cu-matrix.cc / nnet-loss.cc on the path, against a CPU Kaldi surface
(oracle/ref_build/shim/kaldi-ref-shim.h).  The .so travels to the GPU box; the sources do not.

`RefLstm` has the same Python interface as oracle_py.Oracle so a test can run either.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libkaldi_lstm_ref.so")
CUK_LIB_PATH = os.path.join(_HERE, "_ref", "libbd_cu_kernels_ref.so")
REFERENCE_ROOT = "/root/reference"


def build(force=False):
    """Build oracle/_ref when the reference tree is present; returns True when the library exists afterwards."""
    if os.path.isdir(REFERENCE_ROOT) and (force or not os.path.exists(LIB_PATH) or not os.path.exists(CUK_LIB_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return os.path.exists(LIB_PATH)


def available():
    return os.path.exists(LIB_PATH)


_lib = None
_keep = []
_vp, _i, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref is not built (needs /root/reference: make -C oracle ref)")
        L = ctypes.CDLL(LIB_PATH)
        for pfx in ("lstmp_ref_", "lstmp_std_ref_"):
            g = lambda n: getattr(L, pfx + n)
            g("last_error").restype = ctypes.c_char_p
            g("create").restype = _vp
            g("create").argtypes = [_i] * (4 if pfx == "lstmp_ref_" else 3)
            g("destroy").argtypes = [_vp]
            g("num_params").restype = ctypes.c_long
            g("num_params").argtypes = [_vp]
            for n in ("get_params", "set_params", "get_corr", "set_corr"):
                g(n).argtypes = [_vp, _vp]
            g("propagate").argtypes = [_vp, _vp, _i, _vp, _i, _i]
            g("backpropagate").argtypes = [_vp, _vp, _i, _vp, _i, _vp, _i, _i, _f]
            g("update").argtypes = [_vp, _f]
        L.lstmp_ref_get_state.argtypes = [_vp, _vp]
        L.lstmp_ref_set_state.argtypes = [_vp, _vp]
        L.lstmp_ref_buf_rows.argtypes = [_vp, _i]
        L.lstmp_ref_get_buf.argtypes = [_vp, _i, _vp]
        L.lstmp_ref_reset.argtypes = [_vp, _vp, _i]
        L.lstmp_ref_set_sgemm.argtypes = [_vp]
        L.lstmp_ref_write.restype = ctypes.c_long
        L.lstmp_ref_write.argtypes = [_vp, _i, _vp, ctypes.c_long]
        L.lstmp_ref_read.argtypes = [_vp, _i, _vp, ctypes.c_long]
        L.lstmp_ref_sources.restype = ctypes.c_char_p
        L.timeshift_ref_propagate.argtypes = [_i, _vp, _i, _vp, _i, _i, _i]
        L.xent_ref_create.restype = _vp
        L.xent_ref_destroy.argtypes = [_vp]
        L.xent_ref_eval_masked.argtypes = [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i]
        L.xent_ref_get_stats.argtypes = [_vp, _vp]
        L.xent_ref_report.restype = ctypes.c_long
        L.xent_ref_report.argtypes = [_vp, _vp, ctypes.c_long]
        _lib = L
    return _lib


def use_openblas(num_threads=None):
    """Route the reference's cblas_Xgemm (kaldi-matrix.cc:172) through scipy's OpenBLAS; returns the thread count
    (0: not found, the shim's own triple loop runs)."""
    import glob
    try:
        import scipy  # noqa: F401
        base = os.path.join(os.path.dirname(scipy.__file__), os.pardir, "scipy.libs")
        cands = glob.glob(os.path.join(base, "libscipy_openblas*.so"))
        if not cands:
            return 0
        blas = ctypes.CDLL(cands[0])
        fn = getattr(blas, "scipy_cblas_sgemm")
        if num_threads is not None:
            blas.scipy_openblas_set_num_threads(int(num_threads))
        nthr = int(blas.scipy_openblas_get_num_threads())
        lib().lstmp_ref_set_sgemm(ctypes.cast(fn, _vp))
        _keep.append(blas)
        return nthr
    except Exception:
        return 0


def use_builtin_gemm():
    lib().lstmp_ref_set_sgemm(None)


def _p(a):
    return a.ctypes.data_as(_vp)


class _RefBase:
    _pfx = None

    def _f(self, name):
        return getattr(lib(), self._pfx + name)

    def _check(self, rc, what):
        if rc != 0:
            raise ValueError("reference %s: %s" % (what, self._f("last_error")().decode()))

    def __del__(self):
        try:
            if self._h:
                self._f("destroy")(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def num_params(self):
        return int(self._f("num_params")(self._h))

    def set_params(self, flat):
        a = np.ascontiguousarray(flat, np.float32)
        assert a.size == self.num_params
        self._f("set_params")(self._h, _p(a))

    def get_params(self):
        out = np.empty(self.num_params, np.float32)
        self._check(self._f("get_params")(self._h, _p(out)), "GetParams")
        return out

    def get_grads(self):
        """the `*_corr_` members (momentum-accumulated gradients) in GetParams order"""
        out = np.empty(self.num_params, np.float32)
        self._f("get_corr")(self._h, _p(out))
        return out

    def set_grads(self, flat):
        a = np.ascontiguousarray(flat, np.float32)
        self._f("set_corr")(self._h, _p(a))

    def propagate(self, x):
        x = np.ascontiguousarray(x, np.float32)
        rows = x.shape[0]
        out = np.empty((rows, self.R), np.float32)
        self._check(self._f("propagate")(self._h, _p(x), x.shape[1], _p(out), self.R, rows), "Propagate")
        self.T = rows // self.S
        return out

    def backpropagate(self, x, out_diff, momentum, want_in_diff=True):
        x = np.ascontiguousarray(x, np.float32)
        od = np.ascontiguousarray(out_diff, np.float32)
        rows = x.shape[0]
        in_diff = np.empty((rows, self.I), np.float32)
        self._check(self._f("backpropagate")(self._h, _p(x), x.shape[1], _p(od), od.shape[1], _p(in_diff), self.I,
                                              rows, float(momentum)), "Backpropagate")
        return in_diff if want_in_diff else None

    def update(self, lr):
        self._check(self._f("update")(self._h, float(lr)), "Update")


class RefLstm(_RefBase):
    """The reference's LstmProjectedStreams (google/nnet/bd-nnet-lstm-projected-streams.h) on its CPU matrix path."""
    _pfx = "lstmp_ref_"

    def __init__(self, I, C, R, S, dtype=np.float32):
        assert np.dtype(dtype) == np.float32, "the reference is BaseFloat = float"
        self.I, self.C, self.R, self.S = int(I), int(C), int(R), int(S)
        self.W = 7 * self.C + self.R
        self.dtype = np.dtype(np.float32)
        self.T = 0
        self._h = _vp(lib().lstmp_ref_create(self.I, self.C, self.R, self.S))
        if not self._h:
            raise ValueError("reference InitData: %s" % lib().lstmp_ref_last_error().decode())

    def get_state(self):
        out = np.empty((self.S, self.W), np.float32)
        lib().lstmp_ref_get_state(self._h, _p(out))
        return out

    def set_state(self, st):
        a = np.ascontiguousarray(st, np.float32)
        assert a.shape == (self.S, self.W)
        lib().lstmp_ref_set_state(self._h, _p(a))

    def reset(self, flags):
        f = np.ascontiguousarray(flags, np.int32)
        self._check(lib().lstmp_ref_reset(self._h, _p(f), int(f.size)), "Reset")

    def _buf(self, which):
        rows = lib().lstmp_ref_buf_rows(self._h, which)
        out = np.empty((rows, self.W), np.float32)
        lib().lstmp_ref_get_buf(self._h, which, _p(out))
        return out

    def prop_buf(self):
        return self._buf(0)

    def bprop_buf(self):
        return self._buf(1)

    def write(self, binary=True):
        n = lib().lstmp_ref_write(self._h, int(binary), None, 0)
        buf = ctypes.create_string_buffer(n)
        lib().lstmp_ref_write(self._h, int(binary), buf, n)
        return buf.raw[:n]

    def read(self, data, binary=True):
        self._check(lib().lstmp_ref_read(self._h, int(binary), data, len(data)), "ReadData")


class RefStdLstm(_RefBase):
    """The reference's standard/ LstmProjected (one utterance per call, zero initial state, gradient clip in Update)."""
    _pfx = "lstmp_std_ref_"

    def __init__(self, I, C, R):
        self.I, self.C, self.R, self.S = int(I), int(C), int(R), 1
        self.T = 0
        self._h = _vp(lib().lstmp_std_ref_create(self.I, self.C, self.R))
        if not self._h:
            raise ValueError("reference InitData: %s" % lib().lstmp_std_ref_last_error().decode())


def time_shift(x, shift):
    """TimeShift::PropagateFnc (standard/nnet/nnet-time-shift.h:42-51)."""
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    rc = lib().timeshift_ref_propagate(int(shift), _p(x), x.shape[1], _p(out), x.shape[1], x.shape[0], x.shape[1])
    if rc != 0:
        raise ValueError("reference TimeShift failed")
    return out


class RefXent:
    """The reference's Xent::EvalMasked / Report (google/nnet/nnet-loss.cc:76-164, 293-307)."""

    def __init__(self):
        self._h = _vp(lib().xent_ref_create())

    def __del__(self):
        try:
            if self._h:
                lib().xent_ref_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def eval_masked(self, mask, net_out, post):
        """post: list (per frame) of lists of (pdf, weight).  Returns diff."""
        net_out = np.ascontiguousarray(net_out, np.float32)
        frames, num_pdf = net_out.shape
        mask = np.ascontiguousarray(mask, np.float32)
        row_ptr = np.zeros(frames + 1, np.int32)
        pdf, w = [], []
        for t, row in enumerate(post):
            for (p, v) in row:
                pdf.append(p)
                w.append(v)
            row_ptr[t + 1] = len(pdf)
        pdf = np.asarray(pdf, np.int32).reshape(-1)
        w = np.asarray(w, np.float32).reshape(-1)
        if pdf.size == 0:
            pdf, w = np.zeros(1, np.int32), np.zeros(1, np.float32)
        diff = np.empty((frames, num_pdf), np.float32)
        rc = lib().xent_ref_eval_masked(self._h, _p(mask), _p(net_out), num_pdf, frames, num_pdf, _p(row_ptr), _p(pdf),
                                        _p(w), _p(diff), num_pdf)
        if rc != 0:
            raise ValueError("reference EvalMasked: %s" % lib().lstmp_ref_last_error().decode())
        return diff

    def stats(self):
        out = np.zeros(4, np.float64)
        lib().xent_ref_get_stats(self._h, _p(out))
        return {"frames": out[0], "correct": out[1], "loss": out[2], "entropy": out[3]}

    def report(self):
        buf = ctypes.create_string_buffer(4096)
        lib().xent_ref_report(self._h, buf, 4096)
        return buf.value.decode()
