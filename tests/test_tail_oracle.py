"""Pins oracle/tail_oracle.py (numpy restatement of the [upstream] AffineTransform + Softmax + the reference's
Xent::EvalMasked, and of their backward / momentum update) against an independent formulation: torch autograd of the
masked soft-target cross-entropy in fp64.  CPU only."""
import numpy as np
import torch

from oracle import tail_oracle


def _case(rows, I, P, seed):
    rng = np.random.RandomState(seed)
    W = (rng.randn(P, I) * 0.3)
    b = rng.randn(P) * 0.5
    x = rng.randn(rows, I).astype(np.float32)
    mask = (rng.rand(rows) > 0.3).astype(np.float32)
    post = []
    for t in range(rows):
        if t % 4 == 1:
            a, c = rng.randint(0, P, 2)
            post.append([(int(a), 0.6), (int(c), 0.4)])
        elif t % 7 == 3:
            post.append([])
        else:
            post.append([(int(rng.randint(0, P)), 1.0)])
    return W, b, x, mask, post


def test_tail_oracle_matches_autograd():
    rows, I, P = 24, 12, 30
    W, b, x, mask, post = _case(rows, I, P, 3)
    o = tail_oracle.TailOracle(I, P, np.float64)
    o.set_params(np.concatenate([W.ravel(), b]))
    y = o.propagate_eval(x, mask, post)
    in_diff = o.backpropagate(x, 0.0)
    # torch: L = -sum_t mask_t sum_p tgt[t,p] log softmax(x W^T + b)[t,p]; dL/dlogits = mask * (y * sum_p tgt - tgt).
    # Kaldi's diff is mask * (y - tgt): identical when every target row sums to 1 or the row is masked / empty rows are
    # compared separately below.
    Wt = torch.tensor(W, dtype=torch.float64, requires_grad=True)
    bt = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    tgt = torch.zeros(rows, P, dtype=torch.float64)
    for t, lst in enumerate(post):
        for pdf, w in lst:
            tgt[t, pdf] += float(np.float32(w))
    logp = torch.log_softmax(xt @ Wt.T + bt, dim=1)
    np.testing.assert_allclose(y, logp.exp().detach().numpy(), rtol=1e-10, atol=1e-12)
    full = torch.tensor([1.0 if len(l) else 0.0 for l in post], dtype=torch.float64)   # rows whose targets sum to 1
    mt = torch.tensor(mask, dtype=torch.float64)
    loss = -(mt[:, None] * tgt * logp).sum()
    loss.backward()
    # rows with an EMPTY target list: Kaldi's diff is mask * y (no target subtracted), autograd's is 0 -> compare the
    # autograd-covered part and check the empty rows directly
    diff = o.diff
    exp_diff = (mt[:, None] * (logp.exp().detach() * full[:, None] - tgt)).numpy()
    rows_full = full.numpy() == 1
    np.testing.assert_allclose(diff[rows_full], exp_diff[rows_full], rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(diff[~rows_full], (mask[:, None] * y)[~rows_full], rtol=1e-9, atol=1e-12)
    # gradients of the rows autograd covers (the fp32 soft weights 0.6 + 0.4 sum to 1 + 2.4e-8: Kaldi's diff y - t and
    # autograd's y * sum(t) - t differ by that much)
    d_full = diff * full.numpy()[:, None]
    np.testing.assert_allclose(d_full.T @ x.astype(np.float64), Wt.grad.numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(d_full.sum(0), bt.grad.numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(d_full @ W, xt.grad.numpy(), rtol=1e-6, atol=1e-6)
    assert in_diff.shape == (rows, I)
    # the statistics follow the reference's EvalMasked (xent_oracle): frames = masked rows, loss = -sum mask*tgt*log y
    assert o.xent.frames == int(mask.sum())
    assert abs(o.xent.loss - float(loss)) <= 1e-4 * abs(float(loss))


def test_tail_oracle_momentum_update():
    rows, I, P = 10, 8, 12
    W, b, x, mask, post = _case(rows, I, P, 5)
    o = tail_oracle.TailOracle(I, P, np.float64)
    o.set_params(np.concatenate([W.ravel(), b]))
    o.propagate_eval(x, mask, post)
    o.backpropagate(x, 0.9)
    g1w, g1b = o.Wc.copy(), o.bc.copy()
    o.update(0.1)
    np.testing.assert_allclose(o.W, W - 0.1 * g1w)
    np.testing.assert_allclose(o.b, b - 0.1 * g1b)
    o.propagate_eval(x, mask, post)
    d2 = o.diff.copy()
    o.backpropagate(x, 0.9)
    np.testing.assert_allclose(o.Wc, d2.T @ x.astype(np.float64) + 0.9 * g1w)   # corr = G + mmt * corr
    np.testing.assert_allclose(o.bc, d2.sum(0) + 0.9 * g1b)
    assert o.get_params().dtype == np.float32 and o.get_corr().size == P * I + P
