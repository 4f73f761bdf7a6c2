"""N>1 host logic on CPU: two gloo ranks, each owning half of the streams (the oracle stands in for the GPU
engine behind the same fresh_gradient()/Update protocol), must reproduce the single-process S-stream run:
sum all-reduce of the fresh gradients, then momentum + SGD."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

I, C, R, S, T, NCHUNK = 8, 16, 8, 4, 5, 3
LR, MMT = 1e-2, 0.9


class OracleLayer:
    """The oracle behind the interface StreamShardTrainer drives (tests only)."""

    def __init__(self, oracle_py, S_local, flat):
        self.o = oracle_py.Oracle(I, C, R, S_local, np.float32)
        self.o.set_params(flat)
        self.corr = np.zeros_like(flat)
        self.g = torch.zeros(flat.size, dtype=torch.float32)

    def Reset(self, flags):
        self.o.reset(np.asarray(flags, np.int32))

    def Propagate(self, x):
        return torch.from_numpy(self.o.propagate(x.numpy()))

    def Backpropagate(self, x, out, od, want_in_diff=True):
        self.o.set_grads(np.zeros_like(self.corr))
        ind = self.o.backpropagate(x.numpy(), od.numpy(), 0.0)      # fresh gradient (beta = 0)
        self.g.copy_(torch.from_numpy(self.o.get_grads()))
        return torch.from_numpy(ind) if want_in_diff else None

    def fresh_gradient(self):
        return self.g

    def Update(self):
        self.corr = self.g.numpy() + MMT * self.corr                 # LPS.h:465-487 on the all-reduced sum
        self.o.set_grads(self.corr)
        self.o.update(LR)                                            # LPS.h:501-512


def _data():
    rng = np.random.RandomState(5)
    xs = rng.randn(NCHUNK, T, S, I).astype(np.float32)
    ods = (rng.randn(NCHUNK, T, S, R) * 0.1).astype(np.float32)
    return xs, ods


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_py
    import kaldi_lstm_b200 as klb
    flat = oracle_py.init_params(I, C, R, 0.3, 77)
    lo, hi = klb.parallel.shard_streams(S, rank, world)
    layer = OracleLayer(oracle_py, hi - lo, flat)
    trainer = klb.parallel.StreamShardTrainer([layer])
    xs, ods = _data()
    for n in range(NCHUNK):
        x = torch.from_numpy(np.ascontiguousarray(xs[n][:, lo:hi]).reshape(T * (hi - lo), I))
        od = torch.from_numpy(np.ascontiguousarray(ods[n][:, lo:hi]).reshape(T * (hi - lo), R))
        trainer.train_chunk(x, lambda out: od)
    q.put((rank, layer.o.get_params()))
    dist.destroy_process_group()


def test_two_rank_stream_sharding_matches_single_process():
    sys.path.insert(0, ROOT)
    from oracle import oracle_py
    oracle_py.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process, all S streams, momentum applied the reference's way
    flat = oracle_py.init_params(I, C, R, 0.3, 77)
    o = oracle_py.Oracle(I, C, R, S, np.float32)
    o.set_params(flat)
    xs, ods = _data()
    for n in range(NCHUNK):
        o.propagate(xs[n].reshape(T * S, I))
        o.backpropagate(xs[n].reshape(T * S, I), ods[n].reshape(T * S, R), MMT)
        o.update(LR)
    ref = o.get_params()
    np.testing.assert_array_equal(res[0], res[1])                    # replicas stay bit-identical
    assert np.abs(res[0] - ref).max() <= 1e-5 * np.abs(ref).max()    # == single-GPU S-stream update


def test_shard_streams_requires_divisibility():
    import pytest
    import kaldi_lstm_b200 as klb
    assert klb.parallel.shard_streams(256, 3, 8) == (96, 128)
    with pytest.raises(ValueError):
        klb.parallel.shard_streams(10, 0, 4)


# ---- lock-step termination with uneven shards ----------------------------------------------------
def _uneven_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_py
    import kaldi_lstm_b200 as klb
    flat = oracle_py.init_params(I, C, R, 0.3, 77)
    lo, hi = klb.parallel.shard_streams(S, rank, world)
    Sl = hi - lo
    layer = OracleLayer(oracle_py, Sl, flat)
    trainer = klb.parallel.StreamShardTrainer([layer])
    xs, ods = _data()
    my_chunks = NCHUNK if rank == 0 else NCHUNK - 1          # rank 1 runs out of data one chunk early
    it = iter(range(my_chunks))

    def next_chunk():
        n = next(it, None)
        if n is None:
            return None
        x = torch.from_numpy(np.ascontiguousarray(xs[n][:, lo:hi]).reshape(T * Sl, I))
        od = torch.from_numpy(np.ascontiguousarray(ods[n][:, lo:hi]).reshape(T * Sl, R))
        return x, od, None

    def padding_chunk():                                      # all-padding: zero features, zero loss gradient
        return torch.zeros(T * Sl, I), torch.zeros(T * Sl, R), None

    counts = trainer.run(next_chunk, lambda out, od: od, padding_chunk)
    q.put((rank, layer.o.get_params(), counts))
    dist.destroy_process_group()


def test_uneven_shards_terminate_in_lock_step():
    """Ranks run out of data after different numbers of chunks; every chunk carries a collective.  The rank that is
    done keeps stepping on all-padding chunks until every rank is done -- nobody blocks in the all-reduce, and the
    result equals the single-process run in which the exhausted streams are padded (TRAIN.cc:190-202)."""
    sys.path.insert(0, ROOT)
    from oracle import oracle_py
    oracle_py.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30100 + (os.getpid() % 500)
    procs = [ctx.Process(target=_uneven_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = {r: (prm, cnt) for r, prm, cnt in got}
    assert res[0][1] == (NCHUNK, 0) and res[1][1] == (NCHUNK - 1, 1)
    np.testing.assert_array_equal(res[0][0], res[1][0])
    flat = oracle_py.init_params(I, C, R, 0.3, 77)
    o = oracle_py.Oracle(I, C, R, S, np.float32)
    o.set_params(flat)
    xs, ods = _data()
    half = S // 2
    for n in range(NCHUNK):
        x, od = xs[n].copy(), ods[n].copy()
        if n == NCHUNK - 1:                                   # rank 1's streams are exhausted: padded rows
            x[:, half:] = 0
            od[:, half:] = 0
        o.propagate(x.reshape(T * S, I))
        o.backpropagate(x.reshape(T * S, I), od.reshape(T * S, R), MMT)
        o.update(LR)
    ref = o.get_params()
    assert np.abs(res[0][0] - ref).max() <= 1e-5 * np.abs(ref).max()
