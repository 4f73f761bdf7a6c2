"""The engine's GEMM kernels (fp32 SIMT and tcgen05 3xTF32) against an fp64 torch product, on
every operand-order combination the path uses (AddMatMat call sites LPS.h:246,457,468,471,486)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from kaldi_lstm_b200 import engine
    engine.load_library()
    return engine


def _case(eng, backend, M, N, K, tA, tB, alpha=1.0, beta=0.0, bias=False, pad=0, seed=0):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn((K, M + pad) if tA else (M, K + pad), device="cuda", generator=g)[:, :(M if tA else K)]
    B = torch.randn((N, K + pad) if tB else (K, N + pad), device="cuda", generator=g)[:, :(K if tB else N)]
    C = torch.randn((M, N + pad), device="cuda", generator=g)[:, :N]
    bvec = torch.randn(N, device="cuda", generator=g) if bias else None
    opA = (A.t() if tA else A).double()
    opB = (B.t() if tB else B).double()
    ref = alpha * (opA @ opB) + beta * C.double()
    if bias:
        ref = ref + bvec.double()
    out = C.clone() if pad == 0 else C  # strided view is written in place
    if pad == 0:
        out = C.clone()
    eng.debug_gemm(backend, out, M, N, K, alpha, A, tA, B, tB, beta, bvec)
    torch.cuda.synchronize()
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    return err


SHAPES = [
    # (M, N, K, tA, tB)  -- the five call sites at cfg-3 sizes
    (1280, 3200, 40, 0, 1),    # input GEMM  in * W_x^T          (K-major, K-major)
    (1280, 3200, 512, 0, 1),   # layer-2 input GEMM
    (1280, 40, 3200, 0, 0),    # in_diff = DGIFO * W_x           (K-major, MN-major), small N
    (1280, 512, 3200, 0, 0),
    (3200, 40, 1280, 1, 0),    # G(w_gifo_x) = DGIFO^T * in      (MN-major, MN-major), small N
    (3200, 512, 1280, 1, 0),   # G(w_gifo_r)
    (512, 800, 1280, 1, 0),    # G(w_r_m), N not a multiple of the tile
    (80, 3200, 40, 0, 1),      # cfg-2 sizes (M < tile)
    (3200, 512, 80, 1, 0),
    (132, 260, 36, 1, 1),      # ragged everything, (MN-major, K-major)
    (4, 8, 4, 0, 1),
]


@pytest.mark.parametrize("shape", SHAPES)
def test_gemm_simt(eng, shape):
    M, N, K, tA, tB = shape
    assert _case(eng, 0, M, N, K, tA, tB, alpha=0.7, beta=0.3, bias=True) <= 1e-5


@pytest.mark.parametrize("shape", SHAPES)
def test_gemm_tcgen05_3xtf32(eng, shape):
    M, N, K, tA, tB = shape
    err = _case(eng, 1, M, N, K, tA, tB, alpha=0.7, beta=0.3, bias=True)
    assert err <= 5e-5, "3xTF32 error %.3e (single-pass TF32 would be ~1e-3)" % err


def test_gemm_tcgen05_pitched(eng):
    assert _case(eng, 1, 1280, 512, 3200, 0, 0, pad=12) <= 5e-5
    assert _case(eng, 1, 3200, 512, 1280, 1, 0, pad=8) <= 5e-5


# ---- bf16 hi/lo tile-image GEMM (lstmp_gemm_hl.cu): pre-split operands + bulk copies + tcgen05 kind::f16 -----------
@pytest.mark.parametrize("shape", SHAPES + [(640, 16624, 512, 0, 1), (640, 512, 16624, 0, 0), (16624, 512, 640, 1, 0)])
def test_gemm_hl_bf16_split(eng, shape):
    M, N, K, tA, tB = shape
    err = _case(eng, 2, M, N, K, tA, tB, alpha=0.7, beta=0.3, bias=True)
    assert err <= 2e-5, "bf16 hi/lo (4-term) error %.3e (single-pass bf16 would be ~4e-3)" % err


def test_gemm_hl_pitched_and_alpha_beta(eng):
    assert _case(eng, 2, 1280, 512, 3200, 0, 0, pad=12) <= 2e-5
    assert _case(eng, 2, 3200, 512, 1280, 1, 0, pad=8) <= 2e-5
    assert _case(eng, 2, 1280, 3200, 512, 0, 1, alpha=1.0, beta=0.0, bias=False) <= 2e-5
    assert _case(eng, 2, 256, 384, 200, 0, 0, alpha=-1.5, beta=1.0, bias=False, seed=3) <= 2e-5


# ---- grouped launch: the four products after a layer's backward loop as one split + one GEMM + one reduce launch ----
def _group_case(eng, shapes, seed=0, shared_a=()):
    """shapes: list of (M, N, K, tA, tB, alpha, beta, bias); shared_a: pairs (i, j) -- product j reuses product i's A."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    probs, refs = [], []
    for idx, (M, N, K, tA, tB, alpha, beta, bias) in enumerate(shapes):
        A = torch.randn((K, M) if tA else (M, K), device="cuda", generator=g)
        for (i, j) in shared_a:
            if j == idx:
                A = probs[i][5]
        B = torch.randn((N, K) if tB else (K, N), device="cuda", generator=g)
        C = torch.randn((M, N), device="cuda", generator=g)
        bvec = torch.randn(N, device="cuda", generator=g) if bias else None
        ref = alpha * ((A.t() if tA else A).double() @ (B.t() if tB else B).double()) + beta * C.double()
        if bias:
            ref = ref + bvec.double()
        probs.append((C, M, N, K, alpha, A, tA, B, tB, beta, bvec))
        refs.append(ref)
    nl = eng.debug_gemm_group(probs)
    torch.cuda.synchronize()
    errs = [((p[0].double() - r).abs().max() / r.abs().max()).item() for p, r in zip(probs, refs)]
    return nl, errs


def test_gemm_group_backward_layer2(eng):
    # cfg3 layer 2: in_diff, G(w_gifo_x), G(w_gifo_r) (both on DGIFO^T), G(w_r_m)
    shapes = [(1280, 512, 3200, 0, 0, 1.0, 0.0, False), (3200, 512, 1280, 1, 0, 1.0, 0.0, False),
              (3200, 512, 1280, 1, 0, 1.0, 0.0, False), (512, 800, 1280, 1, 0, 1.0, 0.0, False)]
    for rep in range(2):   # second call: cached plan, reused image arena
        nl, errs = _group_case(eng, shapes, seed=rep, shared_a=((1, 2),))
        assert 2 <= nl <= 3
        assert max(errs) <= 2e-5, errs


def test_gemm_group_backward_layer1(eng):
    shapes = [(3200, 40, 1280, 1, 0, 1.0, 0.0, False), (3200, 512, 1280, 1, 0, 1.0, 0.0, False),
              (512, 800, 1280, 1, 0, 1.0, 0.0, False)]
    nl, errs = _group_case(eng, shapes, shared_a=((0, 1),))
    assert max(errs) <= 2e-5, errs


def test_gemm_group_mixed_small_and_ragged(eng):
    # alpha / beta / bias per product, ragged tiles, N not a multiple of 4 (no split-K for that product), tiny products
    shapes = [(132, 260, 36, 1, 1, 0.7, 0.3, True), (4, 8, 4, 0, 1, 1.0, 0.0, False),
              (80, 3200, 40, 0, 1, 0.5, 1.0, True), (200, 50, 1000, 0, 0, 1.0, 0.5, False)]
    nl, errs = _group_case(eng, shapes)
    assert max(errs) <= 2e-5, errs
    nl, errs = _group_case(eng, [(640, 512, 16624, 0, 0, 1.0, 0.0, False)])   # a group of one, deep K (split-K)
    assert max(errs) <= 2e-5, errs
