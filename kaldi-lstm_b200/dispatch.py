"""Multi-stream chunk filling -- mirror of the trainer's stream dispatch
(google/nnetbin/bd-nnet-train-lstm-streams.cc:128-209).

Packs ``num_stream`` utterances into time-major BPTT chunks of ``batch_size`` frames
(row = t*S + s, :187-206), applies the targets delay by reading the feature row
``curt + targets_delay`` clamped to the last frame (:198-202), pads exhausted streams with
mask 0 and the last target (:190-196) and raises the per-stream reset flag when a stream takes
a new utterance (:170).  ``shard(rank, world)`` gives rank r of N every N-th utterance so that
N processes (one per GPU) run the same loop on disjoint data.
"""
import numpy as np


class StreamDispatcher:
    def __init__(self, num_stream, batch_size, targets_delay, feat_dim):
        self.S, self.T, self.delay, self.D = int(num_stream), int(batch_size), int(targets_delay), int(feat_dim)
        self.keys = [""] * self.S
        self.feats = [None] * self.S
        self.targets = [None] * self.S
        self.curt = np.zeros(self.S, np.int64)
        self.lent = np.zeros(self.S, np.int64)
        self.new_utt_flags = np.zeros(self.S, np.int32)
        self.num_done = 0
        self.num_skipped = 0
        self._it = None

    @staticmethod
    def shard(utterances, rank, world):
        for n, u in enumerate(utterances):
            if n % world == rank:
                yield u

    def open(self, utterances):
        """utterances: iterable of (key, feats[frames x D] float32, targets[frames])."""
        self._it = iter(utterances)

    def _refill(self):
        # :146-174
        for s in range(self.S):
            if self.curt[s] < self.lent[s]:
                self.new_utt_flags[s] = 0
                continue
            while True:
                try:
                    key, feats, targets = next(self._it)
                except StopIteration:
                    break
                if targets is None:  # missing targets (:156-161)
                    self.num_skipped += 1
                    continue
                if feats.shape[0] != len(targets):  # length mismatch (:163-167)
                    self.num_skipped += 1
                    continue
                self.keys[s] = key
                self.feats[s] = np.asarray(feats, np.float32)
                self.targets[s] = np.asarray(targets)
                self.curt[s] = 0
                self.lent[s] = feats.shape[0]
                self.new_utt_flags[s] = 1
                break

    def next_chunk(self):
        """Returns (feat [T*S x D], frame_mask [T*S], target [T*S], new_utt_flags [S]) or None
        once every stream is exhausted (:177-181)."""
        self._refill()
        if not np.any(self.curt < self.lent):
            return None
        S, T = self.S, self.T
        feat = np.zeros((T * S, self.D), np.float32)
        mask = np.zeros(T * S, np.float32)
        target = np.zeros(T * S, np.int64)
        t_idx = np.arange(T)
        for s in range(S):
            L = int(self.lent[s])
            if L == 0:
                continue  # stream never got an utterance: rows stay zero, mask 0
            cur = self.curt[s] + t_idx                       # curt at each t (:204 increments every t)
            valid = cur < L                                   # :190
            rows = t_idx * S + s
            mask[rows] = valid.astype(np.float32)
            tgt_idx = np.where(valid, cur, L - 1)             # :192,:195
            target[rows] = self.targets[s][tgt_idx]
            f_idx = np.where(cur + self.delay < L, cur + self.delay, L - 1)  # :198-202
            feat[rows] = self.feats[s][f_idx]
            self.curt[s] += T
        self.num_done += int(self.new_utt_flags.sum())
        return feat, mask, target, self.new_utt_flags.copy()
