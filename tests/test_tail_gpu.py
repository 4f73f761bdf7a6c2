"""GPU parity of the fused output tail (AffineTransform + Softmax + Xent::EvalMasked)
through the C ABI against oracle/tail_oracle.py (fp64 restatement; tolerance 1e-4 relative as for the LSTM path), and of
the fused softmax + masked cross-entropy kernel against the unfused pair."""
import numpy as np
import pytest

from oracle import tail_oracle, xent_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def klb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import kaldi_lstm_b200 as k
    k.load_library()
    return k


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


def _post(rng, rows, P, soft_every=0, empty_every=0):
    post = []
    for t in range(rows):
        if empty_every and t % empty_every == 3:
            post.append([])
        elif soft_every and t % soft_every == 1:
            a, b = rng.randint(0, P, 2)
            post.append([(int(a), 0.7), (int(b), 0.3)])
        else:
            post.append([(int(rng.randint(0, P)), 1.0)])
    return post


@pytest.mark.parametrize("I,P,rows", [(16, 40, 12), (64, 1000, 80), (512, 16624, 640), (512, 8000, 80)])
def test_tail_matches_oracle(klb, I, P, rows):
    """Two chunks of PropagateEval / Backpropagate / Update (momentum 0.9): posteriors, diff, in_diff, momentum-accumulated
    gradients, parameters and the loss statistics."""
    import torch
    rng = np.random.RandomState(P + rows)
    tail = klb.AffineSoftmaxXent(I, P, max_frames=rows)
    tail.InitData("<ParamStddev> 0.1 <BiasMean> -2.0 <BiasRange> 2.0", seed=3)
    lr, mmt = 1e-3, 0.9
    tail.SetTrainOptions(klb.NnetTrainOptions(lr, mmt))
    o = tail_oracle.TailOracle(I, P, np.float64)
    o.set_params(tail.GetParams())
    for n in range(2):
        x = rng.randn(rows, I).astype(np.float32)
        mask = (rng.rand(rows) > 0.2).astype(np.float32)
        post = _post(rng, rows, P, soft_every=5, empty_every=11)
        xd = torch.from_numpy(x).cuda()
        y = tail.PropagateEval(xd, mask, post, want_posteriors=True)
        y_ref = o.propagate_eval(x, mask, post)
        assert _rel(y.cpu().numpy(), y_ref) <= 1e-4
        assert _rel(tail.engine.get_diff(), o.diff) <= 1e-4
        ind = tail.Backpropagate(xd)
        ind_ref = o.backpropagate(x, mmt)
        assert _rel(ind.cpu().numpy(), ind_ref) <= 1e-4
        tail.Update()
        o.update(lr)
        assert _rel(tail.GetGradients(), o.get_corr()) <= 1e-4
        assert _rel(tail.GetParams(), o.get_params()) <= 1e-5
    s = tail.Stats()
    assert s["frames"] == o.xent.frames and s["correct"] == o.xent.correct
    assert abs(s["loss"] - o.xent.loss) <= 1e-4 * abs(o.xent.loss)
    assert abs(s["entropy"] - o.xent.entropy) <= 1e-4 * max(abs(o.xent.entropy), 1.0)
    assert "FRAME_ACCURACY" in tail.Report()


def test_fused_softmax_xent_equals_unfused(klb):
    """lstmp_b200_xent_eval_masked_logits (softmax fused, in place) vs torch.softmax + lstmp_b200_xent_eval_masked."""
    import ctypes
    import torch
    from kaldi_lstm_b200 import engine as E
    rows, P = 96, 4000
    rng = np.random.RandomState(5)
    logits = torch.from_numpy((rng.randn(rows, P) * 3).astype(np.float32)).cuda()
    mask = (np.arange(rows) % 4 != 0).astype(np.float32)
    post = _post(rng, rows, P, soft_every=3, empty_every=7)
    rp, pdf, w = klb.posterior_to_csr(post)
    x1 = klb.Xent(rows)
    y = torch.softmax(logits, dim=1)
    d1 = x1.EvalMasked(mask, y, (rp, pdf, w))
    L = E.load_library()
    vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    L.lstmp_b200_xent_eval_masked_logits.argtypes = [vp, vp, vp, sz, ci, ci, vp, vp, vp, vp, sz, vp, sz, vp]
    x2 = E.XentEngine(rows)
    buf = logits.clone()
    post_out = torch.empty_like(logits)
    E._chk(L.lstmp_b200_xent_eval_masked_logits(x2._h, vp(mask.ctypes.data), vp(buf.data_ptr()), P, rows, P, vp(rp.ctypes.data),
                                                vp(pdf.ctypes.data), vp(w.ctypes.data), vp(post_out.data_ptr()), P,
                                                vp(buf.data_ptr()), P, E._cur_stream(0)))     # in place
    torch.cuda.synchronize()
    assert (post_out - y).abs().max().item() <= 1e-6
    assert (buf - d1).abs().max().item() <= 1e-6
    s1, s2 = x1.Stats(), x2.stats()
    assert s1["frames"] == s2["frames"] and s1["correct"] == s2["correct"]
    assert abs(s1["loss"] - s2["loss"]) <= 1e-4 * abs(s1["loss"])
    # the unfused entry point refuses aliasing (its target columns are read after diff is written)
    with pytest.raises(klb.EngineError):
        x1._engine.eval_masked(mask, y, rp, pdf, w, y)


def test_tail_error_behaviour(klb):
    import torch
    with pytest.raises(klb.EngineError):
        klb.TailEngine(10, 40, 8)          # input_dim % 4 != 0
    tail = klb.AffineSoftmaxXent(16, 40, max_frames=8)
    x = torch.zeros(8, 16, device="cuda")
    with pytest.raises(klb.EngineError):
        tail.engine.backpropagate(x)       # no propagate_eval yet
    with pytest.raises(RuntimeError):      # KALDI_ERR nnet-loss.cc:88-91
        tail.PropagateEval(x, np.ones(8, np.float32), [[(40, 1.0)]] * 8)
    with pytest.raises(RuntimeError):
        tail.InitData("<ParamStdev> 0.1")


def test_cpp_tail_mirror_on_gpu():
    """kaldi-lstm_b200/kaldi/b200-affine-softmax-xent.h (B200AffineSoftmaxXent: PropagateEval / Backpropagate / Update /
    Report against Kaldi's types) vs a double-precision host restatement (tests/cpp/tail_test.cc)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "tests", "cpp"), "-s"])
    r = subprocess.run([os.path.join(root, "tests", "cpp", "_build", "tail_test")], capture_output=True, text=True,
                       timeout=120)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
