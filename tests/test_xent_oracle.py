"""CPU tests of the masked cross-entropy oracle (oracle/xent_oracle.py, restating google/nnet/nnet-loss.cc:76-164)
and of the host-side posterior flattening.  The reference has no tests for this function: the oracle is pinned against
an independent torch formulation and its own committed golden vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import xent_oracle


def _torch_reference(mask, y, post):
    """Independent: build logits = log(y), let autograd differentiate the masked soft-target cross-entropy."""
    rows, P = y.shape
    tgt = torch.zeros(rows, P, dtype=torch.float64)
    for t, lst in enumerate(post):
        for p, w in lst:
            tgt[t, p] += float(np.float32(w))
    logits = torch.log(torch.from_numpy(y).double()).requires_grad_(True)
    logp = torch.log_softmax(logits, dim=1)
    m = torch.from_numpy(mask).double()
    loss = -(m[:, None] * tgt * logp).sum()
    loss.backward()
    # d loss / d logits = mask * (softmax * sum(t) - t); equals mask * (y - t) when sum(t) == 1
    return loss.item(), logits.grad.numpy(), tgt.numpy()


def test_diff_and_loss_against_autograd():
    mask, y, post = xent_oracle.random_case(40, 53, seed=3, soft=True, mask_every=4)
    o = xent_oracle.XentOracle()
    diff = o.eval_masked(mask, y, post)
    loss, grad, tgt = _torch_reference(mask, y, post)
    assert abs(o.loss - loss) <= 1e-5 * abs(loss)
    full = np.isclose(tgt.sum(1), 1.0)  # soft posteriors drawn from a Dirichlet sum to one
    assert full.all()
    np.testing.assert_allclose(diff, grad, rtol=0, atol=2e-6)
    ent = -(mask[:, None] * tgt * np.log(tgt + 1e-20)).sum()
    assert abs(o.entropy - ent) <= 1e-5 * max(abs(ent), 1.0)
    assert o.frames == int(mask.sum())


def test_hard_labels_match_nll():
    mask, y, post = xent_oracle.random_case(64, 31, seed=4, soft=False, mask_every=0)
    o = xent_oracle.XentOracle()
    diff = o.eval_masked(mask, y, post)
    labels = np.array([lst[0][0] for lst in post])
    nll = -np.log(y[np.arange(64), labels].astype(np.float64)).sum()
    assert abs(o.loss - nll) <= 1e-5 * nll
    assert abs(o.entropy) <= 1e-12  # t*log(t + 1e-20) with t == 1
    assert o.correct == int((y.argmax(1) == labels).sum())
    onehot = np.zeros_like(y)
    onehot[np.arange(64), labels] = 1
    np.testing.assert_array_equal(diff, y - onehot)


def test_edge_cases():
    y = np.full((4, 5), 0.2, np.float32)          # every column ties: arg-max is column 0 (first maximum)
    post = [[(0, 1.0)], [], [(3, 0.5), (3, 0.5)], [(2, 1.0)]]
    mask = np.array([1, 1, 1, 0], np.float32)
    o = xent_oracle.XentOracle()
    diff = o.eval_masked(mask, y, post)
    # frame 0: target 0 == argmax 0 -> correct; frame 1: empty posterior, dense target row is all zero -> argmax 0 ->
    # counted correct by the reference's rule; frame 2: duplicates accumulate to 1.0 at column 3 -> wrong; frame 3 masked
    assert o.correct == 2 and o.frames == 3
    assert diff[2, 3] == np.float32(0.2) - np.float32(1.0)
    assert np.all(diff[3] == 0)
    assert abs(o.loss - (-2 * np.log(np.float32(0.2)))) < 1e-6
    with pytest.raises(RuntimeError):              # KALDI_ERR nnet-loss.cc:88-91
        o.eval_masked(mask, y, [[(5, 1.0)], [], [], []])
    rep = o.report()
    assert rep["frames"] == 3 and abs(rep["frame_accuracy"] - 100.0 * 2 / 3) < 1e-9


def test_golden_vectors():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "xent_small.npz"))
    o = xent_oracle.XentOracle()
    for n in range(2):
        rp, pdf, w = g["row_ptr%d" % n], g["pdf%d" % n], g["weight%d" % n]
        post = [[(int(pdf[e]), float(w[e])) for e in range(rp[t], rp[t + 1])] for t in range(len(rp) - 1)]
        diff = o.eval_masked(g["mask%d" % n], g["y%d" % n], post)
        np.testing.assert_array_equal(diff, g["diff%d" % n])
        st = g["stats%d" % n]
        assert abs(o.loss - st[0]) <= 1e-9 * abs(st[0]) and abs(o.entropy - st[1]) <= 1e-9 * max(abs(st[1]), 1)
        assert o.correct == int(st[2]) and o.frames == int(st[3])


def test_posterior_to_csr_roundtrip():
    import kaldi_lstm_b200 as klb
    _, _, post = xent_oracle.random_case(17, 11, seed=5, soft=True, empty_every=4, dup_every=3)
    rp, pdf, w = klb.posterior_to_csr(post)
    assert rp[0] == 0 and rp[-1] == len(pdf) == len(w) == sum(len(x) for x in post)
    back = [[(int(pdf[e]), float(w[e])) for e in range(rp[t], rp[t + 1])] for t in range(len(post))]
    assert back == [[(int(p), float(np.float32(x))) for p, x in lst] for lst in post]


class _FakeXentEngine:
    """CPU stand-in for the CUDA side (lstmp_b200_xent_*) so that kaldi-lstm_b200/loss.py's host logic -- CSR
    flattening, assertions, engine (re)creation, Report formatting -- runs on a box without a GPU."""

    def __init__(self, max_frames, device=0):
        if max_frames <= 0:
            raise ValueError("max_frames")
        self.max_frames = max_frames
        self.o = xent_oracle.XentOracle()
        self.calls = 0

    def eval_masked(self, mask, net_out, row_ptr, pdf, weight, diff):
        assert net_out.shape[0] <= self.max_frames
        post = [[(int(pdf[e]), float(weight[e])) for e in range(row_ptr[t], row_ptr[t + 1])]
                for t in range(len(row_ptr) - 1)]
        diff.copy_(torch.from_numpy(self.o.eval_masked(mask, net_out.numpy(), post)))
        self.calls += 1

    def stats(self):
        return {"loss": self.o.loss, "entropy": self.o.entropy, "correct": self.o.correct, "frames": self.o.frames,
                "kernel_launches": 2 * self.calls}


def test_python_mirror_host_logic(monkeypatch):
    import kaldi_lstm_b200 as klb
    from kaldi_lstm_b200 import loss
    monkeypatch.setattr(loss, "XentEngine", _FakeXentEngine)
    mask, y, post = xent_oracle.random_case(12, 9, seed=8, soft=True, empty_every=5, dup_every=4)
    ref = xent_oracle.XentOracle()
    want = ref.eval_masked(mask, y, post)
    x = klb.Xent()                                   # engine created on first use, sized by the call
    d1 = x.EvalMasked(mask, torch.from_numpy(y), post)
    np.testing.assert_array_equal(d1.numpy(), want)
    d2 = torch.empty(12, 9)
    out = x.EvalMasked(mask, torch.from_numpy(y), klb.posterior_to_csr(post), d2)   # pre-flattened posterior, caller's diff
    assert out is d2
    np.testing.assert_array_equal(d2.numpy(), want)
    s = x.Stats()
    assert s["frames"] == 2 * ref.frames and s["correct"] == 2 * ref.correct and abs(s["loss"] - 2 * ref.loss) < 1e-9
    rep = x.Report()
    assert rep.startswith("AvgLoss: ") and "(Xent), [AvgXent: " in rep and "FRAME_ACCURACY >> " in rep   # :293-307
    with pytest.raises(AssertionError):              # KALDI_ASSERT(num_frames == post.size())   nnet-loss.cc:80
        x.EvalMasked(mask, torch.from_numpy(y), post[:-1])
    with pytest.raises(RuntimeError):                # pdf-id outside the network output            :88-91
        x.EvalMasked(mask, torch.from_numpy(y), [[(9, 1.0)]] + post[1:])
    assert klb.Xent().Stats()["frames"] == 0 and "nan" in klb.Xent().Report()

