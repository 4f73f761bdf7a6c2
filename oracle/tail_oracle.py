"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the network's output tail:
`<AffineTransform> num_pdf input_dim` + `<Softmax>` (google/nnet.proto:4-5) + Xent::EvalMasked
(google/nnet/nnet-loss.cc:76-164, through oracle/xent_oracle.py) and their backward / update.

AffineTransform and Softmax are [upstream] Kaldi nnet1 components of the reference's vintage (October 2014): the
reference tree does not vendor nnet/nnet-affine-transform.h or nnet/nnet-activation.h, so their published algorithm is
restated here and the EvalMasked half is the one pinned against oracle/_ref (tests/test_ref_pin.py):
  AffineTransform::PropagateFnc      out = in * linearity^T + bias
  Softmax::PropagateFnc              y = ApplySoftMaxPerRow: e = exp(x - max(x)), y = e / sum(e)
  Softmax::BackpropagateFnc          in_diff = out_diff                     (derivative folded into the xent diff)
  AffineTransform::BackpropagateFnc  in_diff = out_diff * linearity
  AffineTransform::Update            linearity_corr = diff^T * in + mmt * linearity_corr ; bias_corr = colsum(diff) + mmt *
                                     bias_corr ; linearity -= lr * linearity_corr ; bias -= lr * bias_corr
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module."""
import numpy as np

from .xent_oracle import XentOracle


class TailOracle:
    def __init__(self, input_dim, num_pdf, dtype=np.float64):
        self.I, self.P, self.dt = int(input_dim), int(num_pdf), np.dtype(dtype)
        self.W = np.zeros((self.P, self.I), self.dt)
        self.b = np.zeros(self.P, self.dt)
        self.Wc, self.bc = np.zeros_like(self.W), np.zeros_like(self.b)
        self.xent = XentOracle()
        self.diff = None

    def set_params(self, flat):
        flat = np.asarray(flat, self.dt)
        self.W = flat[:self.P * self.I].reshape(self.P, self.I).copy()
        self.b = flat[self.P * self.I:].copy()

    def get_params(self):
        return np.concatenate([self.W.ravel(), self.b]).astype(np.float32)

    def get_corr(self):
        return np.concatenate([self.Wc.ravel(), self.bc]).astype(np.float32)

    def propagate_eval(self, x, mask, post):
        x = np.asarray(x, self.dt)
        a = x @ self.W.T + self.b
        a = a - a.max(axis=1, keepdims=True)
        e = np.exp(a)
        y = e / e.sum(axis=1, keepdims=True)
        self.diff = self.xent.eval_masked(mask, y.astype(np.float32), post).astype(self.dt) if self.dt == np.float32 else \
            self._diff64(mask, y, post)
        return y

    def _diff64(self, mask, y, post):
        # statistics through the fp32 restatement (what the reference computes); diff in this oracle's precision
        self.xent.eval_masked(mask, y.astype(np.float32), post)
        tgt = np.zeros_like(y)
        for t, lst in enumerate(post):
            for pdf, w in lst:
                tgt[t, pdf] += np.float32(w)
        return (y - tgt) * np.asarray(mask, self.dt)[:, None]

    def backpropagate(self, x, momentum):
        x = np.asarray(x, self.dt)
        in_diff = self.diff @ self.W
        self.Wc = self.diff.T @ x + momentum * self.Wc
        self.bc = self.diff.sum(axis=0) + momentum * self.bc
        return in_diff

    def update(self, lr):
        self.W -= lr * self.Wc
        self.b -= lr * self.bc
