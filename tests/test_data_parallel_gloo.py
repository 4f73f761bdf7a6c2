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
