"""GPU parity of the sparse masked cross-entropy (lstmp_b200_xent_*, kaldi-lstm_b200/csrc/lstmp_xent.cu) against the
CPU oracle of Xent::EvalMasked (google/nnet/nnet-loss.cc:76-164), through the C ABI."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def klb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import kaldi_lstm_b200 as k
    k.load_library()
    return k


def _run(klb, mask, y, post, pad=0, xent=None):
    import torch
    rows, P = y.shape
    buf = torch.zeros(rows, P + pad, device="cuda")
    buf[:, :P] = torch.from_numpy(y).cuda()
    dbuf = torch.full((rows, P + pad), 7.0, device="cuda")
    x = xent or klb.Xent(rows)
    x.EvalMasked(mask, buf[:, :P], post, dbuf[:, :P])
    torch.cuda.synchronize()
    if pad:
        assert torch.all(dbuf[:, P:] == 7.0)  # pitched caller matrix: nothing written beyond num_pdf columns
    return x, dbuf[:, :P].cpu().numpy()


def _check(klb, rows, P, **kw):
    from oracle import xent_oracle
    pad = kw.pop("pad", 0)
    mask, y, post = xent_oracle.random_case(rows, P, **kw)
    o = xent_oracle.XentOracle()
    ref = o.eval_masked(mask, y, post)
    x, diff = _run(klb, mask, y, post, pad=pad)
    np.testing.assert_array_equal(diff, ref)  # (y - t) * mask: the same fp32 operations in the same order
    s = x.Stats()
    assert abs(s["loss"] - o.loss) <= 1e-6 * max(abs(o.loss), 1.0)
    assert abs(s["entropy"] - o.entropy) <= 1e-6 * max(abs(o.entropy), 1.0)
    assert s["correct"] == o.correct and s["frames"] == o.frames
    return x, o


def test_small_hard_and_soft(klb):
    _check(klb, 24, 37, seed=1, soft=False, empty_every=7, dup_every=5)
    _check(klb, 24, 37, seed=2, soft=True, empty_every=7, dup_every=5)


def test_num_pdf_smaller_than_block_and_odd(klb):
    _check(klb, 9, 5, seed=3, soft=True)
    _check(klb, 33, 257, seed=4, soft=False, pad=3)
    _check(klb, 16, 1, seed=5, soft=False, mask_every=0)


def test_recipe_shapes(klb):
    """cfg 2 chunk (80 frames x 8000 pdfs) and the cfg 4 per-GPU chunk (640 x 16624)."""
    _check(klb, 80, 8000, seed=6, soft=False, pad=8)
    _check(klb, 640, 16624, seed=7, soft=False, empty_every=50)


def test_argmax_ties_and_empty_rows(klb):
    from oracle import xent_oracle
    y = np.full((6, 300), 1.0 / 300, np.float32)
    y[1, 299] = y[1, 17] = 0.5          # two equal maxima: the first one (17) wins
    y[2, 256] = 0.9                     # maximum in the second pass of the column loop
    post = [[(0, 1.0)], [(17, 1.0)], [(256, 0.5), (256, 0.5)], [], [(299, 1.0)], [(0, 0.0)]]
    mask = np.array([1, 1, 1, 1, 1, 1], np.float32)
    o = xent_oracle.XentOracle()
    ref = o.eval_masked(mask, y, post)
    x, diff = _run(klb, mask, y, post)
    np.testing.assert_array_equal(diff, ref)
    s = x.Stats()
    assert s["correct"] == o.correct == 5   # all but frame 4 (argmax 0 vs target 299)
    assert s["frames"] == 6


def test_accumulation_reset_and_determinism(klb):
    import torch
    from oracle import xent_oracle
    o = xent_oracle.XentOracle()
    x = klb.Xent(64)
    for seed in (11, 12, 13):
        mask, y, post = xent_oracle.random_case(64, 500, seed=seed, soft=True)
        o.eval_masked(mask, y, post)
        _run(klb, mask, y, post, xent=x)
    s = x.Stats()
    assert abs(s["loss"] - o.loss) <= 1e-6 * o.loss and s["correct"] == o.correct and s["frames"] == o.frames
    assert "FRAME_ACCURACY" in x.Report() and "AvgLoss" in x.Report()
    x._engine.reset_stats()
    assert x.Stats()["frames"] == 0 and x.Stats()["loss"] == 0.0
    # bitwise repeatability (fixed-order reductions, no atomics)
    mask, y, post = xent_oracle.random_case(64, 500, seed=14, soft=True)
    a, b = klb.Xent(64), klb.Xent(64)
    _, d1 = _run(klb, mask, y, post, xent=a)
    _, d2 = _run(klb, mask, y, post, xent=b)
    np.testing.assert_array_equal(d1, d2)
    assert a.Stats()["loss"] == b.Stats()["loss"] and a.Stats()["entropy"] == b.Stats()["entropy"]


def test_golden_fixture(klb):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "xent_small.npz"))
    x = klb.Xent(24)
    for n in range(2):
        post = (g["row_ptr%d" % n], g["pdf%d" % n], g["weight%d" % n])
        _, diff = _run(klb, g["mask%d" % n], g["y%d" % n], post, xent=x)
        np.testing.assert_array_equal(diff, g["diff%d" % n])
        st, s = g["stats%d" % n], x.Stats()
        assert abs(s["loss"] - st[0]) <= 1e-6 * abs(st[0]) and abs(s["entropy"] - st[1]) <= 1e-6 * max(abs(st[1]), 1)
        assert s["correct"] == int(st[2]) and s["frames"] == int(st[3])


def test_errors(klb):
    import torch
    y = torch.full((4, 8), 0.125, device="cuda")
    x = klb.Xent(4)
    with pytest.raises(RuntimeError):   # KALDI_ERR nnet-loss.cc:88-91
        x.EvalMasked(np.ones(4, np.float32), y, [[(8, 1.0)], [], [], []])
    with pytest.raises(AssertionError):  # KALDI_ASSERT nnet-loss.cc:80
        x.EvalMasked(np.ones(4, np.float32), y, [[(1, 1.0)]])
    with pytest.raises(klb.EngineError):
        klb.XentEngine(0)
    with pytest.raises(klb.EngineError):  # host tensor
        x._engine.eval_masked(np.ones(4, np.float32), torch.zeros(4, 8), np.zeros(5, np.int32), np.zeros(0, np.int32),
                              np.zeros(0, np.float32), torch.zeros(4, 8, device="cuda"))


def test_full_size_properties(klb):
    """Size-independent properties at cfg 4's full single-GPU shape (5120 frames x 16624 pdfs)."""
    import torch
    rows, P = 5120, 16624
    g = torch.Generator(device="cuda").manual_seed(5)
    y = torch.softmax(torch.randn(rows, P, device="cuda", generator=g) * 2, dim=1)
    labels = torch.randint(0, P, (rows,), device="cuda", generator=g)
    mask = (torch.arange(rows) % 5 != 0).float().numpy()
    lab = labels.cpu().numpy().astype(np.int32)
    post = (np.arange(rows + 1, dtype=np.int32), lab, np.ones(rows, np.float32))
    x = klb.Xent(rows)
    diff = x.EvalMasked(mask, y, post)
    m = torch.from_numpy(mask).cuda()
    # row sums: mask * (sum y - 1); masked rows are exactly zero; off-target entries equal mask * y bit for bit
    assert diff[m == 0].abs().max().item() == 0.0
    assert (diff.sum(1).abs() <= 1e-4).all()
    onehot = torch.zeros_like(y)
    onehot[torch.arange(rows), labels] = 1
    assert torch.equal(diff, (y - onehot) * m[:, None])
    s = x.Stats()
    nll = -(torch.log(y[torch.arange(rows), labels]).double() * m.double()).sum().item()
    assert abs(s["loss"] - nll) <= 1e-6 * nll
    assert s["frames"] == int(mask.sum())
    assert s["correct"] == int(((y.argmax(1) == labels) & (m == 1)).sum().item())
