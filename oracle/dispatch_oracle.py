"""Literal restatement (plain loops) of the reference trainer's multi-stream chunk filling,
google/nnetbin/bd-nnet-train-lstm-streams.cc:128-209.  TEST INFRASTRUCTURE: checks
kaldi-lstm_b200/dispatch.py; small cases only."""
import numpy as np


def run(utterances, num_stream, batch_size, targets_delay, feat_dim):
    """utterances: list of (key, feats, targets-or-None).  Returns the list of chunks
    (feat, frame_mask, target, new_utt_flags) the reference loop would produce."""
    it = iter(utterances)
    done_reading = [False]

    def reader_next():
        try:
            return next(it)
        except StopIteration:
            done_reading[0] = True
            return None

    S = num_stream
    feats = [None] * S
    targets = [None] * S
    curt = [0] * S                     # :132
    lent = [0] * S                     # :133
    new_utt_flags = [0] * S            # :134
    chunks = []
    while True:                        # :143
        for s in range(S):             # :146
            if curt[s] < lent[s]:      # :148
                new_utt_flags[s] = 0
                continue
            while not done_reading[0]:  # :153
                u = reader_next()
                if u is None:
                    break
                key, f, t = u
                if t is None:          # :156 missing targets
                    continue
                if f.shape[0] != len(t):  # :163 length mismatch
                    continue
                feats[s], targets[s] = f, t
                curt[s] = 0            # :168
                lent[s] = f.shape[0]   # :169
                new_utt_flags[s] = 1   # :170
                break
        done = 1                       # :177
        for s in range(S):
            if curt[s] < lent[s]:
                done = 0
        if done:
            break
        feat = np.zeros((batch_size * S, feat_dim), np.float32)   # :139 (kSetZero once; rows are overwritten)
        mask = np.zeros(batch_size * S, np.float32)
        target = np.zeros(batch_size * S, np.int64)
        for t in range(batch_size):    # :187
            for s in range(S):         # :188
                if lent[s] == 0:
                    curt[s] += 1       # never-filled stream: the reference would index targets[s][-1]; skipped here
                    continue
                if curt[s] < lent[s]:  # :190
                    mask[t * S + s] = 1
                    target[t * S + s] = targets[s][curt[s]]
                else:
                    mask[t * S + s] = 0
                    target[t * S + s] = targets[s][lent[s] - 1]
                if curt[s] + targets_delay < lent[s]:   # :198
                    feat[t * S + s] = feats[s][curt[s] + targets_delay]
                else:
                    feat[t * S + s] = feats[s][lent[s] - 1]
                curt[s] += 1           # :204
        chunks.append((feat, mask, target, np.array(new_utt_flags, np.int32)))
    return chunks
