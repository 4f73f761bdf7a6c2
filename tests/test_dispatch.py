"""Host logic: the multi-stream chunk filling mirror (kaldi-lstm_b200/dispatch.py) against a literal
restatement of bd-nnet-train-lstm-streams.cc:128-209 (oracle/dispatch_oracle.py)."""
import numpy as np
import pytest

import kaldi_lstm_b200 as klb
from oracle import dispatch_oracle


def _utts(rng, n, dim, lo, hi, bad_every=0):
    out = []
    for i in range(n):
        L = int(rng.randint(lo, hi))
        f = rng.randn(L, dim).astype(np.float32)
        t = rng.randint(0, 100, size=L)
        if bad_every and i % bad_every == 1:
            t = t[:-1]            # length mismatch -> skipped with a warning (:163-167)
        if bad_every and i % bad_every == 2:
            t = None              # missing targets (:156-161)
        out.append(("utt%d" % i, f, t))
    return out


@pytest.mark.parametrize("S,T,delay,n", [(4, 20, 5, 11), (3, 7, 0, 5), (8, 5, 9, 30), (2, 20, 5, 1), (4, 3, 2, 0)])
def test_dispatch_matches_reference_loop(S, T, delay, n):
    rng = np.random.RandomState(S * 100 + T)
    utts = _utts(rng, n, 6, 1, 60, bad_every=4)
    ref = dispatch_oracle.run(utts, S, T, delay, 6)
    d = klb.StreamDispatcher(S, T, delay, 6)
    d.open(utts)
    got = []
    while True:
        c = d.next_chunk()
        if c is None:
            break
        got.append(c)
    assert len(got) == len(ref)
    for (f, m, t, fl), (rf, rm, rt, rfl) in zip(got, ref):
        np.testing.assert_array_equal(f, rf)
        np.testing.assert_array_equal(m, rm)
        np.testing.assert_array_equal(t, rt)
        np.testing.assert_array_equal(fl, rfl)
    # every valid frame is seen exactly once
    valid = sum(u[1].shape[0] for u in utts if u[2] is not None and len(u[2]) == u[1].shape[0])
    assert int(sum(m.sum() for _, m, _, _ in got)) == valid


def test_dispatch_sharding_partitions_utterances():
    utts = [("u%d" % i, np.zeros((3, 2), np.float32), np.zeros(3, np.int64)) for i in range(10)]
    seen = []
    for r in range(4):
        seen += [k for k, _, _ in klb.StreamDispatcher.shard(utts, r, 4)]
    assert sorted(seen) == sorted(k for k, _, _ in utts)
