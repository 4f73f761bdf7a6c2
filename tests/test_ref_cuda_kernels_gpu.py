"""Runs the reference's OWN two CUDA kernels on the B200 (google/cudamatrix/bd-cu-kernels.cu:17-49, compiled by
`make -C oracle ref` with nvcc for sm_100a against a stub cu-matrixdim.h -> oracle/_ref/libbd_cu_kernels_ref.so) with
the launch geometry of their call sites (cu-matrix.cc:1014-1068) and checks them against the CPU formulas the oracle
restates (kaldi-matrix.cc:447-497).  This closes the loop "reference GPU path == reference CPU path == oracle" for
the two ops the reference adds to Kaldi; the engine fuses both into its recurrent kernels.
Tolerance 1e-6: nvcc contracts a*b + c into FMA, the CPU path rounds the product first."""
import ctypes
import os

import numpy as np
import pytest

from oracle import ref_py

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(ref_py.CUK_LIB_PATH), reason="oracle/_ref CUDA kernels not built")]

CU2DBLOCK = 16  # upstream cu-matrixdim.h


class Dim3(ctypes.Structure):
    _fields_ = [("x", ctypes.c_uint), ("y", ctypes.c_uint), ("z", ctypes.c_uint)]


class MatrixDim(ctypes.Structure):
    _fields_ = [("rows", ctypes.c_int32), ("cols", ctypes.c_int32), ("stride", ctypes.c_int32)]


def _nb(n):
    return (n + CU2DBLOCK - 1) // CU2DBLOCK


@pytest.fixture(scope="module")
def cuk():
    import torch
    torch.zeros(1, device="cuda")  # primary context before the runtime-API launches in the .so
    L = ctypes.CDLL(ref_py.CUK_LIB_PATH)
    vp, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    L.cudaF_add_mat_diag_vec.argtypes = [Dim3, Dim3, f, vp, MatrixDim, vp, i, i, vp, f]
    L.cudaF_add_mat_dot_mat.argtypes = [Dim3, Dim3, vp, vp, vp, i, i, MatrixDim, i, i, f, f]
    return L


@pytest.mark.parametrize("rows,cols,pad", [(64, 800, 0), (80, 13, 3), (1, 1, 0), (1280, 800, 8)])
def test_reference_add_mat_diag_vec_kernel(cuk, rows, cols, pad):
    """`YI.AddMatDiagVec(1.0, YC(t-1), kNoTrans, peephole_i_c_, 1.0)` (LPS.h:278)."""
    import torch
    rng = np.random.RandomState(rows + cols)
    stride = cols + pad
    d = rng.randn(rows, stride).astype(np.float32)
    m = rng.randn(rows, stride).astype(np.float32)
    v = rng.randn(cols).astype(np.float32)
    alpha, beta = np.float32(0.75), np.float32(1.0)
    dd, dm, dv = torch.from_numpy(d).cuda(), torch.from_numpy(m).cuda(), torch.from_numpy(v).cuda()
    cuk.cudaF_add_mat_diag_vec(Dim3(_nb(rows), _nb(cols), 1), Dim3(CU2DBLOCK, CU2DBLOCK, 1), float(alpha),
                               dd.data_ptr(), MatrixDim(rows, cols, stride), dm.data_ptr(), stride, 1, dv.data_ptr(),
                               float(beta))
    torch.cuda.synchronize()
    want = d.copy()
    want[:, :cols] += alpha * v[None, :] * m[:, :cols]           # kaldi-matrix.cc:466-470
    got = dd.cpu().numpy()
    assert np.array_equal(got[:, cols:], d[:, cols:])            # padding untouched
    assert np.abs(got[:, :cols] - want[:, :cols]).max() <= 1e-6 * np.abs(want).max()


@pytest.mark.parametrize("rows,cols,pad", [(64, 800, 0), (80, 13, 3), (1, 1, 0)])
def test_reference_add_mat_dot_mat_kernel(cuk, rows, cols, pad):
    """`YC(t).AddMatDotMat(1.0, YG(t), kNoTrans, YI(t), kNoTrans, 1.0)` (LPS.h:292-293)."""
    import torch
    rng = np.random.RandomState(rows * 3 + cols)
    stride = cols + pad
    d = rng.randn(rows, stride).astype(np.float32)
    a = rng.randn(rows, stride).astype(np.float32)
    b = rng.randn(rows, stride).astype(np.float32)
    alpha, beta = np.float32(1.0), np.float32(0.5)
    dd, da, db = torch.from_numpy(d).cuda(), torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    cuk.cudaF_add_mat_dot_mat(Dim3(_nb(cols), _nb(rows), 1), Dim3(CU2DBLOCK, CU2DBLOCK, 1), dd.data_ptr(),
                              da.data_ptr(), db.data_ptr(), 0, 0, MatrixDim(rows, cols, stride), stride, stride,
                              float(alpha), float(beta))
    torch.cuda.synchronize()
    want = d.copy()
    want[:, :cols] = beta * d[:, :cols] + alpha * a[:, :cols] * b[:, :cols]   # kaldi-matrix.cc:489-491
    got = dd.cpu().numpy()
    assert np.array_equal(got[:, cols:], d[:, cols:])
    assert np.abs(got[:, :cols] - want[:, :cols]).max() <= 1e-6 * np.abs(want).max()
