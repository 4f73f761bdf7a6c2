"""Pins the C oracle (oracle/lstmp_streams_oracle.c).

The reference ships no golden vectors for this path, so the
primary pin is tests/test_ref_pin.py: the oracle against oracle/_ref, the reference's own
headers compiled here.  This file adds the checks that do not depend on the reference's
control flow being right: (a) an independent torch-autograd restatement of the equations,
(b) fp64 central finite differences, and (c) the committed golden fixtures (tests/golden/,
generated from oracle/_ref), so a silent change of the oracle is caught.
"""
import os

import numpy as np
import pytest

import ref_torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rand_case(oracle_py, I, C, R, S, T, seed, scale=0.3, state_scale=0.5):
    rng = np.random.RandomState(seed)
    flat = oracle_py.init_params(I, C, R, scale, seed + 1).astype(np.float64)
    x = rng.randn(T * S, I)
    od = rng.randn(T * S, R) * 0.1
    c0 = rng.randn(S, C) * state_scale
    r0 = rng.randn(S, R) * state_scale
    return flat, x, od, c0, r0


def _state(S, C, R, c0, r0, dtype):
    st = np.zeros((S, 7 * C + R), dtype)
    st[:, 4 * C:5 * C] = c0
    st[:, 7 * C:] = r0
    return st


@pytest.mark.parametrize("shape", [(5, 7, 4, 3, 6), (3, 8, 4, 1, 9), (40, 16, 12, 4, 5)])
def test_oracle_f64_matches_autograd(oracle_mod, shape):
    I, C, R, S, T = shape
    flat, x, od, c0, r0 = _rand_case(oracle_mod, I, C, R, S, T, seed=11)
    o = oracle_mod.Oracle(I, C, R, S, np.float64)
    o.set_params(flat)
    o.set_state(_state(S, C, R, c0, r0, np.float64))
    out = o.propagate(x)
    in_diff = o.backpropagate(x, od, momentum=0.0)
    ref_out, ref_in_diff, ref_grad, ref_cT, ref_rT = ref_torch.fwd_bwd(flat, x, od, c0, r0, S, I, C, R)
    np.testing.assert_allclose(out, ref_out, rtol=0, atol=1e-12)
    np.testing.assert_allclose(in_diff, ref_in_diff, rtol=0, atol=1e-12)
    np.testing.assert_allclose(o.get_grads(), ref_grad, rtol=0, atol=1e-11)
    st = o.get_state()
    np.testing.assert_allclose(st[:, 4 * C:5 * C], ref_cT, atol=1e-12)
    np.testing.assert_allclose(st[:, 7 * C:], ref_rT, atol=1e-12)


def test_oracle_clamp_forward_not_backward(oracle_mod):
    """|c| driven past 50: forward clamps (LPS.h:296-297), backward does not mask (LPS.h:424-428)."""
    I, C, R, S, T = 4, 6, 3, 2, 8
    flat, x, od, c0, r0 = _rand_case(oracle_mod, I, C, R, S, T, seed=5)
    c0 = np.full((S, C), 49.5) * np.sign(np.random.RandomState(0).randn(S, C))
    sl = oracle_mod.param_slices(I, C, R)
    a, b, _ = sl["bias"]
    flat[a:b] = 3.0  # gates wide open: c grows past +-50
    o = oracle_mod.Oracle(I, C, R, S, np.float64)
    o.set_params(flat)
    o.set_state(_state(S, C, R, c0, r0, np.float64))
    out = o.propagate(x)
    cbuf = o.prop_buf()[:, 4 * C:5 * C]
    assert np.abs(cbuf).max() == 50.0
    in_diff = o.backpropagate(x, od, momentum=0.0)
    ref_out, ref_in_diff, ref_grad, _, _ = ref_torch.fwd_bwd(flat, x, od, c0, r0, S, I, C, R)
    np.testing.assert_allclose(out, ref_out, atol=1e-12)
    np.testing.assert_allclose(in_diff, ref_in_diff, atol=1e-11)
    np.testing.assert_allclose(o.get_grads(), ref_grad, atol=1e-10)


def test_oracle_finite_differences(oracle_mod):
    """fp64 central differences of L = sum(out * od) wrt a sample of parameters."""
    I, C, R, S, T = 3, 8, 4, 2, 3
    flat, x, od, c0, r0 = _rand_case(oracle_mod, I, C, R, S, T, seed=3)

    def loss(p):
        o = oracle_mod.Oracle(I, C, R, S, np.float64)
        o.set_params(p)
        o.set_state(_state(S, C, R, c0, r0, np.float64))
        return float((o.propagate(x) * od).sum())

    o = oracle_mod.Oracle(I, C, R, S, np.float64)
    o.set_params(flat)
    o.set_state(_state(S, C, R, c0, r0, np.float64))
    o.propagate(x)
    o.backpropagate(x, od, momentum=0.0)
    g = o.get_grads()
    rng = np.random.RandomState(9)
    idx = rng.choice(flat.size, 60, replace=False)
    eps = 1e-6
    for k in idx:
        p1, p2 = flat.copy(), flat.copy()
        p1[k] += eps
        p2[k] -= eps
        fd = (loss(p1) - loss(p2)) / (2 * eps)
        assert abs(fd - g[k]) <= 1e-6 * max(1.0, abs(g[k])), (k, fd, g[k])


def test_oracle_momentum_update_reset_and_carry(oracle_mod):
    """corr = grad + mmt*corr (LPS.h:465-487); param -= lr*corr (LPS.h:501-512); state carry
    (LPS.h:231,331) and Reset (LPS.h:212-220) across two chunks."""
    I, C, R, S, T = 6, 10, 5, 4, 4
    flat, x1, od1, c0, r0 = _rand_case(oracle_mod, I, C, R, S, T, seed=21)
    _, x2, od2, _, _ = _rand_case(oracle_mod, I, C, R, S, T, seed=22)
    mmt, lr = 0.9, 1e-2
    o = oracle_mod.Oracle(I, C, R, S, np.float64)
    o.set_params(flat)
    o.set_state(_state(S, C, R, c0, r0, np.float64))
    o.propagate(x1)
    o.backpropagate(x1, od1, momentum=mmt)
    g1 = o.get_grads().copy()
    _, _, rg1, cT, rT = ref_torch.fwd_bwd(flat, x1, od1, c0, r0, S, I, C, R)
    np.testing.assert_allclose(g1, rg1, atol=1e-11)  # corr starts at 0
    o.update(lr)
    p1 = o.get_params()
    np.testing.assert_allclose(p1, flat - lr * rg1, atol=1e-13)
    flags = np.array([0, 1, 0, 1], np.int32)
    o.reset(flags)
    keep = (1 - flags)[:, None].astype(np.float64)
    o.propagate(x2)
    o.backpropagate(x2, od2, momentum=mmt)
    _, _, rg2, _, _ = ref_torch.fwd_bwd(p1, x2, od2, cT * keep, rT * keep, S, I, C, R)
    np.testing.assert_allclose(o.get_grads(), rg2 + mmt * rg1, atol=1e-10)


def test_oracle_f32_close_to_f64(oracle_mod):
    I, C, R, S, T = 40, 32, 16, 4, 20
    flat, x, od, c0, r0 = _rand_case(oracle_mod, I, C, R, S, T, seed=7, scale=0.1)
    res = []
    for dt in (np.float32, np.float64):
        o = oracle_mod.Oracle(I, C, R, S, dt)
        o.set_params(flat)
        o.set_state(_state(S, C, R, c0, r0, dt))
        out = o.propagate(x)
        ind = o.backpropagate(x, od, momentum=0.9)
        res.append((out, ind, o.get_grads()))
    for a, b in zip(res[0], res[1]):
        assert np.abs(a - b).max() <= 2e-5 * np.abs(b).max()


def test_oracle_openblas_path_matches_builtin(oracle_mod):
    I, C, R, S, T = 40, 24, 16, 4, 6
    flat, x, od, c0, r0 = _rand_case(oracle_mod, I, C, R, S, T, seed=13, scale=0.1)

    def run():
        o = oracle_mod.Oracle(I, C, R, S, np.float32)
        o.set_params(flat)
        o.set_state(_state(S, C, R, c0, r0, np.float32))
        out = o.propagate(x)
        o.backpropagate(x, od, momentum=0.0)
        return out, o.get_grads()

    a = run()
    nthr = oracle_mod.use_openblas(1)
    try:
        if nthr == 0:
            pytest.skip("no bundled OpenBLAS")
        b = run()
    finally:
        oracle_mod.use_builtin_gemm()
    for u, v in zip(a, b):
        assert np.abs(u - v).max() <= 1e-5 * np.abs(v).max()


def test_oracle_bad_shape(oracle_mod):
    o = oracle_mod.Oracle(4, 4, 4, 3, np.float32)
    with pytest.raises(ValueError):
        o.propagate(np.zeros((4, 4), np.float32))  # 4 rows not a multiple of S=3 (LPS.h:225)


def test_oracle_reproduces_golden(oracle_mod):
    g = np.load(os.path.join(GOLD, "lstmp_small.npz"))
    I, C, R, S, T = [int(v) for v in g["dims"]]
    o = oracle_mod.Oracle(I, C, R, S, np.float32)
    o.set_params(g["params"])
    for n in range(int(g["nchunks"])):
        o.reset(g["flags"][n])
        out = o.propagate(g["x"][n])
        ind = o.backpropagate(g["x"][n], g["out_diff"][n], float(g["momentum"]))
        o.update(float(g["lr"]))
        for a, b in ((out, g["out"][n]), (ind, g["in_diff"][n]), (o.get_grads(), g["corr"][n]),
                     (o.get_params(), g["params_after"][n]), (o.get_state(), g["state"][n])):
            assert np.abs(a - b).max() <= 1e-6 * max(np.abs(b).max(), 1e-30)
    # and the fp64 twin agrees with the stored fp32 vectors to fp32 accuracy
    o64 = oracle_mod.Oracle(I, C, R, S, np.float64)
    o64.set_params(g["params"])
    o64.reset(g["flags"][0])
    out64 = o64.propagate(g["x"][0])
    assert np.abs(out64 - g["out"][0]).max() <= 1e-5 * np.abs(out64).max()
