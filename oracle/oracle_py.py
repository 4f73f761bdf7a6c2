"""ctypes binding of the CPU oracle (oracle/lstmp_streams_oracle.c).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product package
(kaldi-lstm_b200/) never does.

The oracle restates google/nnet/bd-nnet-lstm-projected-streams.h:212-512 of the
reference on the CPU; see the header of the C file for the line-by-line map.
"""
import ctypes
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liblstmp_oracle.so")


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "lstmp_streams_oracle.c"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None
_blas_keepalive = []


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        for pfx, ct in (("lstmp_oracle_f32_", ctypes.c_float), ("lstmp_oracle_f64_", ctypes.c_double)):
            g = lambda n: getattr(_lib, pfx + n)
            g("create").restype = ctypes.c_void_p
            g("create").argtypes = [ctypes.c_int] * 4
            g("destroy").argtypes = [ctypes.c_void_p]
            g("num_params").restype = ctypes.c_long
            g("num_params").argtypes = [ctypes.c_void_p]
            for n in ("set_params", "get_params", "set_grads", "get_grads", "get_state", "set_state"):
                g(n).argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            g("prop_buf").restype = ctypes.c_void_p
            g("prop_buf").argtypes = [ctypes.c_void_p]
            g("bprop_buf").restype = ctypes.c_void_p
            g("bprop_buf").argtypes = [ctypes.c_void_p]
            g("reset").argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
            g("propagate").restype = ctypes.c_int
            g("propagate").argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_int]
            g("backpropagate").restype = ctypes.c_int
            g("backpropagate").argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ct]
            g("update").argtypes = [ctypes.c_void_p, ct]
            g("clip_grads").argtypes = [ctypes.c_void_p, ct]
            g("set_sgemm").argtypes = [ctypes.c_void_p]
    return _lib


def use_openblas(num_threads=None):
    """Route the fp32 oracle's AddMatMat through the OpenBLAS bundled with scipy
    (the reference's CPU path is cblas_Xgemm, kaldi-matrix.cc:172).  Returns the
    number of BLAS threads in use, or 0 when no OpenBLAS was found (the oracle's
    own loops then run)."""
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
        lib().lstmp_oracle_f32_set_sgemm(ctypes.cast(fn, ctypes.c_void_p))
        _blas_keepalive.append(blas)
        return nthr
    except Exception:
        return 0


def use_builtin_gemm():
    lib().lstmp_oracle_f32_set_sgemm(None)


class Oracle:
    """One LstmProjectedStreams component on the CPU (fp32 or fp64)."""

    def __init__(self, I, C, R, S, dtype=np.float32):
        self.I, self.C, self.R, self.S = int(I), int(C), int(R), int(S)
        self.W = 7 * self.C + self.R
        self.dtype = np.dtype(dtype)
        self._pfx = "lstmp_oracle_f32_" if self.dtype == np.float32 else "lstmp_oracle_f64_"
        self._h = ctypes.c_void_p(self._f("create")(self.I, self.C, self.R, self.S))
        self.T = 0

    def _f(self, name):
        return getattr(lib(), self._pfx + name)

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

    def _arr(self, a):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        return a, a.ctypes.data_as(ctypes.c_void_p)

    def set_params(self, flat):
        a, p = self._arr(flat)
        assert a.size == self.num_params
        self._f("set_params")(self._h, p)

    def get_params(self):
        out = np.empty(self.num_params, self.dtype)
        self._f("get_params")(self._h, out.ctypes.data_as(ctypes.c_void_p))
        return out

    def set_grads(self, flat):
        a, p = self._arr(flat)
        self._f("set_grads")(self._h, p)

    def get_grads(self):
        out = np.empty(self.num_params, self.dtype)
        self._f("get_grads")(self._h, out.ctypes.data_as(ctypes.c_void_p))
        return out

    def get_state(self):
        out = np.empty((self.S, self.W), self.dtype)
        self._f("get_state")(self._h, out.ctypes.data_as(ctypes.c_void_p))
        return out

    def set_state(self, st):
        a, p = self._arr(st)
        assert a.shape == (self.S, self.W)
        self._f("set_state")(self._h, p)

    def reset(self, flags):
        f = np.ascontiguousarray(flags, dtype=np.int32)
        self._f("reset")(self._h, f.ctypes.data_as(ctypes.c_void_p), int(f.size))

    def propagate(self, x):
        x, px = self._arr(x)
        rows = x.shape[0]
        out = np.empty((rows, self.R), self.dtype)
        rc = self._f("propagate")(self._h, px, x.shape[1], out.ctypes.data_as(ctypes.c_void_p), self.R, rows)
        if rc != 0:
            raise ValueError("oracle propagate: bad shape (rows %% S != 0)")
        self.T = rows // self.S
        return out

    def backpropagate(self, x, out_diff, momentum, want_in_diff=True):
        x, px = self._arr(x)
        od, pod = self._arr(out_diff)
        rows = x.shape[0]
        in_diff = np.empty((rows, self.I), self.dtype) if want_in_diff else None
        pid = in_diff.ctypes.data_as(ctypes.c_void_p) if want_in_diff else None
        mm = ctypes.c_float(momentum) if self.dtype == np.float32 else ctypes.c_double(momentum)
        rc = self._f("backpropagate")(self._h, px, x.shape[1], pod, od.shape[1], pid, self.I, rows, mm)
        if rc != 0:
            raise ValueError("oracle backpropagate: rc=%d" % rc)
        return in_diff

    def update(self, lr):
        self._f("update")(self._h, ctypes.c_float(lr) if self.dtype == np.float32 else ctypes.c_double(lr))

    def clip_grads(self, max_grad=50.0):
        v = ctypes.c_float(max_grad) if self.dtype == np.float32 else ctypes.c_double(max_grad)
        self._f("clip_grads")(self._h, v)

    def _buf(self, name):
        ptr = self._f(name)(self._h)
        n = (self.T + 2) * self.S * self.W
        ct = ctypes.c_float if self.dtype == np.float32 else ctypes.c_double
        arr = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ct)), shape=(n,))
        return arr.reshape((self.T + 2) * self.S, self.W).copy()

    def prop_buf(self):
        """propagate_buf_ as the reference lays it out: (T+2)S x [g|i|f|o|c|h|m|r]."""
        return self._buf("prop_buf")

    def bprop_buf(self):
        return self._buf("bprop_buf")


# --------------------------------------------------------------------------
# Flat-parameter helpers shared by tests and bench (GetParams order, LPS.h:162-189)
# --------------------------------------------------------------------------
def param_slices(I, C, R):
    lens = [4 * C * I, 4 * C * R, 4 * C, C, C, C, R * C]
    names = ["w_gifo_x", "w_gifo_r", "bias", "peephole_i_c", "peephole_f_c", "peephole_o_c", "w_r_m"]
    shapes = [(4 * C, I), (4 * C, R), (4 * C,), (C,), (C,), (C,), (R, C)]
    out, off = {}, 0
    for n, l, s in zip(names, lens, shapes):
        out[n] = (off, off + l, s)
        off += l
    return out


def init_params(I, C, R, scale, seed):
    """U(-scale, +scale) like InitMatParam/InitVecParam (LPS.h:41-53); own RNG stream
    (Kaldi's RandUniform sequence is not reproducible outside Kaldi)."""
    rng = np.random.RandomState(seed)
    n = 4 * C * I + 4 * C * R + 4 * C + 3 * C + R * C
    return ((rng.random_sample(n) - 0.5) * 2.0 * scale).astype(np.float32)
