"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's masked cross-entropy
(`Xent::EvalMasked`, google/nnet/nnet-loss.cc:76-164), dense-matrix formulation exactly as the reference
computes it on its CPU matrix path.  The reference ships no tests for this function; parity is PINNED against the
reference's own nnet-loss.cc compiled into oracle/_ref (tests/test_ref_pin.py::test_xent_eval_masked_matches_reference,
hard and soft targets, out-of-range pdf) and against an independent torch formulation (tests/test_xent_oracle.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this module; the product
(kaldi-lstm_b200/) never does.
"""
import numpy as np


class XentOracle:
    """Accumulators of `class Xent` (nnet-loss.h:60-75): frames_, correct_, loss_, entropy_."""

    def __init__(self):
        self.frames = 0
        self.correct = 0
        self.loss = 0.0
        self.entropy = 0.0

    @staticmethod
    def _find_row_max_id(m):
        # CuMatrixBase::FindRowMaxId, CPU branch (google/cudamatrix/cu-matrix.cc:1327-1345): strict '<' scan from
        # column 0 starting at -1e21, i.e. the FIRST column holding the row maximum (np.argmax has the same rule)
        return np.argmax(m, axis=1).astype(np.int32)

    @staticmethod
    def _cu_sum(m):
        # CuMatrixBase::Sum (cu-matrix.cc:1984-1988): fp32 column sums accumulated row by row (AddRowSumMat), then
        # Vector::Sum in double, returned as Real = float.  Pinned bit-for-bit against oracle/_ref.
        col = np.add.reduce(np.ascontiguousarray(m, np.float32), axis=0, dtype=np.float32)
        return float(np.float32(np.sum(col, dtype=np.float64)))

    def eval_masked(self, frame_mask_host, net_out, post):
        """frame_mask_host: [rows] float32 (1 = valid frame); net_out: [rows x num_pdf] softmax outputs;
        post: Kaldi `Posterior`, a list (per row) of lists of (pdf, weight).  Returns diff [rows x num_pdf]."""
        net_out = np.asarray(net_out, np.float32)
        mask = np.asarray(frame_mask_host, np.float32)
        rows, num_pdf = net_out.shape
        assert rows == len(post)                                            # :80
        # convert posterior to matrix                                        :82-96
        tgt = np.zeros((rows, num_pdf), np.float32)
        for t, lst in enumerate(post):
            for pdf, w in lst:
                if pdf >= num_pdf:
                    raise RuntimeError("Posterior pdf-id out of NN-output dimension: nn-outputs %d, pdf-id %d"
                                       % (num_pdf, pdf))                     # :88-91 KALDI_ERR
                tgt[t, pdf] += np.float32(w)
        # derivative wrt. the activations of the last layer, masked           :103-106
        diff = (net_out - tgt) * mask[:, None]
        # frames where the maxima match, valid frames only                    :108-121
        mo, mt = self._find_row_max_id(net_out), self._find_row_max_id(tgt)
        correct = int(np.sum((mask == 1.0) & (mo == mt)))
        # cross entropy and entropy                                            :123-136
        with np.errstate(divide="ignore", invalid="ignore"):
            xe = (np.log(net_out) * tgt) * mask[:, None]
            en = (np.log(tgt + np.float32(1e-20)) * tgt) * mask[:, None]
        cross_entropy = -self._cu_sum(xe)
        entropy = -self._cu_sum(en)
        self.loss += cross_entropy                                           # :138-142
        self.entropy += entropy
        self.correct += correct
        self.frames += int(np.float32(mask.sum(dtype=np.float64)))
        return diff.astype(np.float32)

    def report(self):
        # Xent::Report (nnet-loss.cc, after EvalMasked): average (loss - entropy) per frame and frame accuracy
        f = max(self.frames, 1)
        return {"avg_loss": (self.loss - self.entropy) / f, "xent": self.loss / f, "entropy": self.entropy / f,
                "frame_accuracy": 100.0 * self.correct / f, "frames": self.frames}


def random_case(rows, num_pdf, seed=0, soft=False, empty_every=0, dup_every=0, mask_every=3):
    """Seeded (mask, softmax output, posterior) triple for tests and the benchmark."""
    rng = np.random.RandomState(seed)
    logits = rng.randn(rows, num_pdf).astype(np.float32) * 2.0
    logits -= logits.max(axis=1, keepdims=True)
    y = np.exp(logits)
    y = (y / y.sum(axis=1, keepdims=True)).astype(np.float32)
    post = []
    for t in range(rows):
        if empty_every and t % empty_every == empty_every - 1:
            post.append([])
            continue
        if soft:
            k = 1 + rng.randint(3)
            pdfs = rng.randint(0, num_pdf, size=k)
            ws = rng.dirichlet(np.ones(k)).astype(np.float32)
            lst = [(int(p), float(w)) for p, w in zip(pdfs, ws)]
        else:
            lst = [(int(rng.randint(0, num_pdf)), 1.0)]
        if dup_every and t % dup_every == 0:
            lst.append((lst[0][0], 0.25))  # duplicate pdf in one frame: weights accumulate (:93)
        post.append(lst)
    mask = np.ones(rows, np.float32)
    if mask_every:
        mask[::mask_every] = 0.0
    return mask, y, post
