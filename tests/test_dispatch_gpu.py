"""GPU: device-side chunk assembly (lstmp_b200_dispatch_*), the TimeShift gather and the
standard component's clipped update, each through the C ABI against its CPU statement -- bit-exact where the work is
copies and single fp32 operations."""
import os
import subprocess

import numpy as np
import pytest

from oracle import dispatch_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def klb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import kaldi_lstm_b200 as k
    k.load_library()
    return k


def _utts(rng, n, dim, lo, hi, bad_every=0):
    out = []
    for i in range(n):
        L = int(rng.randint(lo, hi))
        f = rng.randn(L, dim).astype(np.float32)
        t = rng.randint(0, 100, size=L)
        if bad_every and i % bad_every == 1:
            t = t[:-1]
        if bad_every and i % bad_every == 2:
            t = None
        out.append(("utt%d" % i, f, t))
    return out


@pytest.mark.parametrize("S,T,delay,n,transform", [(4, 20, 5, 11, True), (3, 7, 0, 5, False), (8, 5, 9, 30, True),
                                                   (2, 20, 5, 1, True), (64, 20, 5, 200, True)])
def test_device_dispatch_is_bit_exact_with_the_reference_loop(klb, S, T, delay, n, transform):
    D = 40 if S == 64 else 8
    rng = np.random.RandomState(S * 100 + T)
    utts = _utts(rng, n, D, 1, 90, bad_every=4)
    shift = (rng.randn(D) * 10).astype(np.float32) if transform else None
    scale = (0.2 + 0.01 * np.arange(D)).astype(np.float32) if transform else None
    ref = dispatch_oracle.run(utts, S, T, delay, D)
    d = klb.DeviceStreamDispatcher(S, T, delay, D, max_utt_frames=128, shift=shift, scale=scale)
    d.open(utts)
    got = []
    while True:
        c = d.next_chunk()
        if c is None:
            break
        got.append((c[0].cpu().numpy(), c[1], c[2], c[3]))
    assert len(got) == len(ref)
    for (f, m, t, fl), (rf, rm, rt, rfl) in zip(got, ref):
        if transform:
            # AddShift then Rescale (feature_transform.nnet.txt:2-5), applied by the reference to the filled chunk; rows
            # of streams that never got an utterance stay zero in both (the transform of a zero row is not zero, but
            # the reference never reads those rows: their mask is 0 and lent == 0 only when the data ran out)
            filled = np.abs(rf).sum(axis=1) > 0
            rf = np.where(filled[:, None], (rf + shift) * scale, 0).astype(np.float32)
        np.testing.assert_array_equal(f, rf)
        np.testing.assert_array_equal(m, rm)
        np.testing.assert_array_equal(t, rt)
        np.testing.assert_array_equal(fl, rfl)
    st = d.stats()
    assert st["chunks_assembled"] == len(ref) and st["kernel_launches"] == len(ref)
    valid = [u for u in utts if u[2] is not None and len(u[2]) == u[1].shape[0]]
    # every utterance crosses PCIe once; per chunk only 3*S ints
    assert st["h2d_bytes"] == sum(u[1].nbytes for u in valid[:st["utterances_loaded"]]) + len(ref) * 12 * S


@pytest.mark.parametrize("shift", [-7, -1, 0, 3, 5, 40])
def test_time_shift_on_device(klb, shift):
    import torch
    x = torch.randn(23, 12, device="cuda")
    ts = klb.TimeShift(12)
    ts.InitData("<Shift> %d" % shift)
    out = ts.Propagate(x)
    rows = klb.nnet_io.time_shift_rows(23, shift)
    assert torch.equal(out, x[torch.from_numpy(rows).cuda()])
    # pitched views
    buf = torch.full((23, 20), float("nan"), device="cuda")
    view = buf[:, :12]
    ts.PropagateFnc(x, view)
    assert torch.equal(view, x[torch.from_numpy(rows).cuda()])
    with pytest.raises(RuntimeError):
        ts.InitData("<Shfit> 3")


def test_standard_component_clipped_update(klb, oracle_mod):
    """<LstmProjected> (standard/): the S=1 case of the engine + element-wise gradient clip at 50 in Update
    (standard/nnet/nnet-lstm-projected.h:480-493) against the oracle's clip_grads (pinned to the reference's own
    standard component in tests/test_ref_pin.py)."""
    import torch
    from parity_util import assert_close
    I, C, R, T = 8, 16, 8, 12
    flat = oracle_mod.init_params(I, C, R, 0.3, 5)
    comp = klb.LstmProjectedStreams(I, R, max_frames=T)
    comp.InitData("<CellDim> %d <NumStream> 1" % C)
    comp.SetParams(flat)
    comp.max_grad_ = 50.0
    comp.SetTrainOptions(klb.NnetTrainOptions(1e-4, 0.9))
    o = oracle_mod.Oracle(I, C, R, 1, np.float32)
    o.set_params(flat)
    rng = np.random.RandomState(53)
    for n in range(2):
        x = rng.randn(T, I).astype(np.float32)
        od = (rng.randn(T, R) * 400.0).astype(np.float32)   # large enough that the clip is active
        comp.Reset([1])
        o.reset(np.array([1], np.int32))
        xd = torch.from_numpy(x).cuda()
        out = comp.Propagate(xd)
        assert_close(out.cpu().numpy(), o.propagate(x), "out %d" % n)
        comp.Backpropagate(xd, out, torch.from_numpy(od).cuda())
        o.backpropagate(x, od, 0.9)
        assert np.abs(o.get_grads()).max() > 50.0
        o.clip_grads(50.0)
        o.update(1e-4)
        comp.Update()
        corr = comp.GetGradients()
        assert np.abs(corr).max() == 50.0
        assert np.abs(corr - o.get_grads()).max() <= 1e-4 * 50.0
        assert_close(comp.GetParams(), o.get_params(), "params %d" % n)


def test_cpp_dispatch_mirror_on_gpu():
    """kaldi-lstm_b200/kaldi/b200-stream-dispatch.h (B200StreamDispatch::NextChunk, B200TimeShiftPropagate) against a
    literal host restatement of the trainer loop, bit-exact (tests/cpp/dispatch_test.cc)."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s"])
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "_build", "dispatch_test")], capture_output=True, text=True,
                       timeout=120)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
