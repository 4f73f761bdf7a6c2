"""Host-side mirror of the reference's `class Xent` for the masked, multi-stream trainer
(google/nnet/nnet-loss.h:33-75; `EvalMasked` nnet-loss.cc:76-164, `Report` :293-307), over the C ABI's
lstmp_b200_xent_* entry points.  The Kaldi `Posterior` (list per frame of (pdf, weight) pairs) is flattened to CSR on
the host -- a few bytes per frame -- instead of the reference's dense [frames x num_pdf] host matrix + H2D copy."""
import numpy as np

from .engine import XentEngine


def posterior_to_csr(post):
    """Kaldi Posterior -> (row_ptr int32[frames+1], pdf int32[nnz], weight float32[nnz])."""
    row_ptr = np.zeros(len(post) + 1, np.int32)
    for t, lst in enumerate(post):
        row_ptr[t + 1] = row_ptr[t] + len(lst)
    nnz = int(row_ptr[-1])
    pdf = np.empty(nnz, np.int32)
    weight = np.empty(nnz, np.float32)
    e = 0
    for lst in post:
        for p, w in lst:
            pdf[e] = p
            weight[e] = w
            e += 1
    return row_ptr, pdf, weight


class Xent:
    def __init__(self, max_frames=0, device=0):
        self._device = device
        self._engine = XentEngine(max_frames, device) if max_frames > 0 else None
        self._carried = {"loss": 0.0, "entropy": 0.0, "correct": 0, "frames": 0, "kernel_launches": 0}

    def EvalMasked(self, frame_mask_host, net_out, post, diff=None):
        """diff = frame_mask * (net_out - target); accumulates loss / entropy / correct / frames on the device.
        `post` is a Kaldi Posterior or an already flattened (row_ptr, pdf, weight) triple."""
        import torch
        rows = net_out.shape[0]
        assert rows == (len(post[0]) - 1 if isinstance(post, tuple) else len(post))     # KALDI_ASSERT nnet-loss.cc:80
        if self._engine is None or self._engine.max_frames < rows:
            if self._engine is not None:      # keep the statistics across the re-creation (as B200Xent does in C++)
                old = self._engine.stats()
                for k in self._carried:
                    self._carried[k] += old[k]
                self._engine.close()
            self._engine = XentEngine(rows, self._device)
        row_ptr, pdf, weight = post if isinstance(post, tuple) else posterior_to_csr(post)
        if diff is None:
            diff = torch.empty_like(net_out)                                          # diff->Resize  :103
        try:
            self._engine.eval_masked(frame_mask_host, net_out, row_ptr, pdf, weight, diff)
        except Exception as e:  # KALDI_ERR on a pdf-id outside the network output       :88-91
            if "pdf-id" in str(e):
                raise RuntimeError(str(e))
            raise
        return diff

    def Stats(self):
        s = dict(self._carried)
        if self._engine is not None:
            for k, v in self._engine.stats().items():
                s[k] += v
        return s

    def Report(self):
        """Xent::Report, nnet-loss.cc:293-307 (without the progress vector)."""
        s = self.Stats()
        f = s["frames"] if s["frames"] else float("nan")
        return ("AvgLoss: %g (Xent), [AvgXent: %g, AvgTargetEnt: %g]\n\nFRAME_ACCURACY >> %g%% <<"
                % ((s["loss"] - s["entropy"]) / f, s["loss"] / f, s["entropy"] / f, 100.0 * s["correct"] / f))
