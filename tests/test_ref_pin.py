"""Pins the restated oracle (oracle/lstmp_streams_oracle.c, oracle/xent_oracle.py, nnet_io.time_shift_rows) to THE
REFERENCE ITSELF: oracle/_ref is the reference's unmodified bd-nnet-lstm-projected-streams.h / nnet-lstm-projected.h /
nnet-time-shift.h / nnet-loss.{h,cc} and the kaldi-matrix.cc / cu-matrix.cc method bodies on the path, compiled from
/root/reference by `make -C oracle ref`.  The .so travels to the GPU box; when it is
absent (no reference tree and no prebuilt library) these tests skip and the golden fixtures -- generated FROM _ref,
tests/golden/make_golden.py -- carry the pin.

Tolerance: 1e-6 relative.  Both sides are routed through the SAME cblas_sgemm (scipy's OpenBLAS, one thread --
the reference's CPU path calls cblas_Xgemm, kaldi-matrix.cc:172), so what is compared is everything the reference
does around the GEMMs: op order, buffer layout, peepholes, clamp, gradient sums, momentum, update.  Without a BLAS
the two private GEMM loops differ in summation order and the bound is 5e-6.
"""
import os

import numpy as np
import pytest

from oracle import oracle_py, ref_py, xent_oracle

pytestmark = pytest.mark.skipif(not ref_py.build(), reason="oracle/_ref not built and /root/reference absent")

TOL = 1e-6


@pytest.fixture(scope="module", autouse=True)
def _same_sgemm_on_both_sides():
    global TOL
    n1 = ref_py.use_openblas(1) if ref_py.available() else 0
    n2 = oracle_py.use_openblas(1)
    if not (n1 and n2):
        TOL = 5e-6
    yield
    if ref_py.available():
        ref_py.use_builtin_gemm()
    oracle_py.use_builtin_gemm()


def _close(a, b, what, tol=None):
    tol = TOL if tol is None else tol
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, what
    err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
    assert err <= tol, "%s: rel_max %.3e > %.0e" % (what, err, tol)


def _pair(I, C, R, S, scale, seed):
    flat = oracle_py.init_params(I, C, R, scale, seed)
    o = oracle_py.Oracle(I, C, R, S, np.float32)
    r = ref_py.RefLstm(I, C, R, S)
    o.set_params(flat)
    r.set_params(flat)
    return flat, o, r


def _run_chunks(o, r, I, R, S, T, nchunks, seed, momentum, lr, resets=None, od_scale=0.1, record=True):
    rng = np.random.RandomState(seed)
    for n in range(nchunks):
        x = rng.randn(T * S, I).astype(np.float32)
        od = (rng.randn(T * S, R) * od_scale).astype(np.float32)
        if resets is not None:
            o.reset(resets[n])
            r.reset(resets[n])
        _close(o.propagate(x), r.propagate(x), "chunk %d out" % n)
        if record:
            _close(o.prop_buf(), r.prop_buf(), "chunk %d propagate_buf_" % n)
        _close(o.get_state(), r.get_state(), "chunk %d prev_nnet_state_" % n)
        _close(o.backpropagate(x, od, momentum), r.backpropagate(x, od, momentum), "chunk %d in_diff" % n)
        if record:
            _close(o.bprop_buf(), r.bprop_buf(), "chunk %d backpropagate_buf_" % n)
        _close(o.get_grads(), r.get_grads(), "chunk %d *_corr_" % n)
        o.update(lr)
        r.update(lr)
        _close(o.get_params(), r.get_params(), "chunk %d params" % n)


def test_ref_sources_are_the_reference():
    src = ref_py.lib().lstmp_ref_sources().decode()
    assert "bd-nnet-lstm-projected-streams.h" in src and "kaldi-matrix.cc" in src


def test_cfg2_recipe_default_three_chunks_resets_momentum():
    """configs[1]: 40 -> 800/512, NumStream 4, 20-frame BPTT (LPS.h:222-512, TRAIN.cc:128-209 flags)."""
    I, C, R, S, T = 40, 800, 512, 4, 20
    _, o, r = _pair(I, C, R, S, 0.05, 7)
    resets = [np.array([1, 1, 1, 1], np.int32), np.array([0, 1, 0, 0], np.int32), np.array([1, 0, 0, 1], np.int32)]
    _run_chunks(o, r, I, R, S, T, 3, 11, momentum=0.9, lr=1e-3, resets=resets)


def test_cfg3_both_layers_s64_one_chunk():
    """configs[2]: layer 1 (40 -> 800/512) and layer 2 (512 -> 800/512) at NumStream 64, T 20; OpenBLAS sgemm."""
    for I in (40, 512):
        C, R, S, T = 800, 512, 64, 20
        _, o, r = _pair(I, C, R, S, 0.05, 3 + I)
        _run_chunks(o, r, I, R, S, T, 1, 5 + I, momentum=0.0, lr=1e-4, record=False)


def test_bit_exact_with_the_same_sgemm():
    """With one cblas_sgemm under both, every forward / backward activation is BIT-identical: the restated oracle
    performs the reference's fp32 operations in the reference's order (incl. its double-literal promotions,
    kaldi-matrix.cc:2561-2593)."""
    if TOL != 1e-6:
        pytest.skip("no shared BLAS")
    I, C, R, S, T = 40, 96, 48, 4, 10
    _, o, r = _pair(I, C, R, S, 0.3, 59)
    rng = np.random.RandomState(61)
    for n in range(3):
        x = rng.randn(T * S, I).astype(np.float32)
        od = (rng.randn(T * S, R) * 0.1).astype(np.float32)
        fl = np.array([n == 1, 0, n == 2, 0], np.int32)
        o.reset(fl)
        r.reset(fl)
        assert np.array_equal(o.propagate(x), r.propagate(x))
        assert np.array_equal(o.prop_buf(), r.prop_buf())
        assert np.array_equal(o.backpropagate(x, od, 0.9), r.backpropagate(x, od, 0.9))
        assert np.array_equal(o.bprop_buf(), r.bprop_buf())
        o.update(1e-3)
        r.update(1e-3)
        _close(o.get_params(), r.get_params(), "params", 1e-7)


def test_single_stream_100_frames():
    """configs[0] shape on the streams component: S = 1, T = 100."""
    I, C, R, S, T = 40, 64, 32, 1, 100
    _, o, r = _pair(I, C, R, S, 0.2, 19)
    _run_chunks(o, r, I, R, S, T, 2, 23, momentum=0.5, lr=1e-2)


@pytest.mark.parametrize("momentum", [0.0, 0.9])
def test_small_odd_shapes_carry_and_reset(momentum):
    I, C, R, S, T = 7, 13, 5, 3, 6
    _, o, r = _pair(I, C, R, S, 0.5, 29)
    resets = [np.array(f, np.int32) for f in ([0, 0, 0], [0, 1, 0], [1, 0, 1], [0, 0, 0])]
    _run_chunks(o, r, I, R, S, T, 4, 31, momentum=momentum, lr=5e-2, resets=resets)


def test_cell_clamp_saturation():
    """c driven past +-50: ApplyFloor/ApplyCeiling forward (LPS.h:296-297), no mask backward (LPS.h:424-428)."""
    I, C, R, S, T = 4, 6, 3, 2, 8
    flat, o, r = _pair(I, C, R, S, 0.3, 37)
    sl = oracle_py.param_slices(I, C, R)
    a, b, _ = sl["bias"]
    flat = flat.copy()
    flat[a:b] = 3.0
    o.set_params(flat)
    r.set_params(flat)
    st = np.zeros((S, 7 * C + R), np.float32)
    st[:, 4 * C:5 * C] = 49.5 * np.sign(np.random.RandomState(0).randn(S, C))
    o.set_state(st)
    r.set_state(st)
    _run_chunks(o, r, I, R, S, T, 2, 41, momentum=0.0, lr=1e-3)
    assert np.abs(r.prop_buf()[:, 4 * C:5 * C]).max() == 50.0


def test_reference_asserts_on_ragged_rows():
    r = ref_py.RefLstm(4, 4, 4, 3)
    with pytest.raises(ValueError):
        r.propagate(np.zeros((4, 4), np.float32))  # KALDI_ASSERT(in.NumRows() % nstream_ == 0), LPS.h:225


def test_reference_write_read_roundtrip_matches_nnet_io():
    """WriteData/ReadData of the reference (LPS.h:101-160) vs this repo's Kaldi text reader (nnet_io.parse_nnet)."""
    import kaldi_lstm_b200 as klb
    I, C, R, S = 5, 6, 4, 2
    flat, _, r = _pair(I, C, R, S, 0.3, 43)
    text = r.write(binary=False).decode()
    body = "<Nnet>\n<LstmProjectedStreams> %d %d\n%s</Nnet>\n" % (R, I, text)
    comps = klb.nnet_io.parse_nnet(body)
    assert comps[0].type == "<LstmProjectedStreams>" and comps[0].attr("<NumStream>") == S
    _close(klb.nnet_io.lstm_flat_params(comps[0]), flat, "text round trip", 1e-5)
    r2 = ref_py.RefLstm(I, C, R, S)
    r2.read(r.write(binary=True), binary=True)
    assert np.array_equal(r2.get_params(), flat)


def test_binary_model_io_matches_the_reference_bytes():
    """nnet_io's BINARY writer emits byte for byte what the reference's WriteData(os, binary=true) emits (LPS.h:133-150
    token order; FM / FV containers of kaldi-matrix.cc:1172-1211), its reader parses the reference's bytes, and the
    reference's ReadData (LPS.h:101-131) accepts what nnet_io writes."""
    import kaldi_lstm_b200 as klb
    nio = klb.nnet_io
    I, C, R, S = 5, 6, 4, 3
    flat, _, r = _pair(I, C, R, S, 0.3, 44)
    ref_bytes = r.write(binary=True)
    comp = nio.lstm_component_from_flat(flat, R, I, C, num_stream=S)
    assert nio.component_data_binary(comp) == ref_bytes
    # the reference's bytes inside a whole binary <Nnet> (header, component marker + dims as upstream Component::Write)
    body = b"\0B<Nnet> <LstmProjectedStreams> " + b"\x04" + np.int32(R).tobytes() + b"\x04" + np.int32(I).tobytes() + \
        ref_bytes + b"</Nnet> "
    comps = nio.parse_nnet_binary(body)
    assert len(comps) == 1 and comps[0].type == "<LstmProjectedStreams>"
    assert comps[0].attr("<CellDim>") == C and comps[0].attr("<NumStream>") == S
    assert np.array_equal(nio.lstm_flat_params(comps[0]), flat)
    assert nio.format_nnet_binary(comps) == body
    # and back into the reference
    flat2 = oracle_py.init_params(I, C, R, 0.2, 45)
    r2 = ref_py.RefLstm(I, C, R, S)
    r2.read(nio.component_data_binary(nio.lstm_component_from_flat(flat2, R, I, C, num_stream=S)), binary=True)
    assert np.array_equal(r2.get_params(), flat2)


def test_standard_lstm_projected_is_the_s1_case_with_grad_clip():
    """standard/nnet/nnet-lstm-projected.h == streams component at S = 1 from zero state, plus the element-wise
    gradient clip at 50 in Update (:480-493) = oracle.clip_grads."""
    I, C, R, T = 9, 12, 6, 15
    flat = oracle_py.init_params(I, C, R, 0.4, 47)
    o = oracle_py.Oracle(I, C, R, 1, np.float32)
    s = ref_py.RefStdLstm(I, C, R)
    o.set_params(flat)
    s.set_params(flat)
    rng = np.random.RandomState(53)
    for n in range(2):
        x = rng.randn(T, I).astype(np.float32)
        od = (rng.randn(T, R) * 400.0).astype(np.float32)  # large enough that the clip is active
        o.reset(np.array([1], np.int32))  # the standard component starts every utterance from zero state
        _close(o.propagate(x), s.propagate(x), "std out %d" % n)
        _close(o.backpropagate(x, od, 0.9), s.backpropagate(x, od, 0.9), "std in_diff %d" % n)
        _close(o.get_grads(), s.get_grads(), "std corr %d" % n, 2e-6)
        gmax = np.abs(s.get_grads()).max()
        assert gmax > 50.0
        o.clip_grads(50.0)
        o.update(1e-4)
        s.update(1e-4)
        assert np.abs(s.get_grads()).max() == 50.0
        assert np.abs(o.get_grads() - s.get_grads()).max() <= 2e-6 * gmax, "std clipped corr %d" % n
        _close(o.get_params(), s.get_params(), "std params %d" % n)


@pytest.mark.parametrize("shift", [-7, -1, 0, 3, 5, 40])
def test_time_shift_matches_reference(shift):
    import kaldi_lstm_b200 as klb
    x = np.random.RandomState(shift + 100).randn(23, 5).astype(np.float32)
    assert np.array_equal(x[klb.nnet_io.time_shift_rows(x.shape[0], shift)], ref_py.time_shift(x, shift))


def _xent_case(frames, num_pdf, seed, soft=False, with_empty=True, with_dup=True, ties=True):
    rng = np.random.RandomState(seed)
    logits = rng.randn(frames, num_pdf).astype(np.float32) * 2
    y = np.exp(logits - logits.max(1, keepdims=True))
    y = (y / y.sum(1, keepdims=True)).astype(np.float32)
    post = []
    for t in range(frames):
        if soft:
            k = rng.randint(1, 4)
            ids = rng.choice(num_pdf, k, replace=False)
            w = rng.dirichlet(np.ones(k)).astype(np.float32)
            post.append([(int(i), float(v)) for i, v in zip(ids, w)])
        else:
            post.append([(int(rng.randint(num_pdf)), 1.0)])
    if with_empty and frames > 2:
        post[1] = []
    if with_dup and frames > 3:
        p = int(rng.randint(num_pdf))
        post[2] = [(p, 0.25), (p, 0.5), ((p + 1) % num_pdf, 0.25)]
    if ties and frames > 4:
        y[3, :] = 1.0 / num_pdf  # arg-max tie across the whole row: first maximum wins (cu-matrix.cc:1333-1343)
    mask = (rng.rand(frames) > 0.3).astype(np.float32)
    return mask, y, post


@pytest.mark.parametrize("soft", [False, True])
def test_xent_eval_masked_matches_reference(soft):
    """Xent::EvalMasked (nnet-loss.cc:76-164): diff bit-exact, accumulators to 1e-6."""
    ref = ref_py.RefXent()
    orc = xent_oracle.XentOracle()
    for n, (frames, num_pdf) in enumerate(((16, 11), (80, 200), (7, 33))):
        mask, y, post = _xent_case(frames, num_pdf, 61 + n, soft=soft)
        d_ref = ref.eval_masked(mask, y, post)
        d_or = orc.eval_masked(mask, y, post)
        assert np.array_equal(d_ref, d_or), "diff chunk %d" % n
    tot = {"frames": orc.frames, "correct": orc.correct, "loss": orc.loss, "entropy": orc.entropy}
    got = ref.stats()
    assert got["frames"] == tot["frames"] and got["correct"] == tot["correct"]
    assert abs(got["loss"] - tot["loss"]) <= 1e-6 * max(1.0, abs(tot["loss"]))
    assert abs(got["entropy"] - tot["entropy"]) <= 1e-6 * max(1.0, abs(tot["entropy"]))
    assert "Xent" in ref.report() or "xent" in ref.report().lower()


def test_xent_reference_rejects_out_of_range_pdf():
    ref = ref_py.RefXent()
    y = np.full((2, 4), 0.25, np.float32)
    with pytest.raises(ValueError):
        ref.eval_masked(np.ones(2, np.float32), y, [[(4, 1.0)], [(0, 1.0)]])  # KALDI_ERR, nnet-loss.cc:88-91
