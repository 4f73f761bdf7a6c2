"""2-GPU parity (skipped on a 1-GPU box): two NCCL ranks with S/2 streams each and one sum all-reduce of the
fresh gradients per Update reproduce the single-GPU S-stream parameters."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
I, C, R, S, T, NCHUNK = 40, 800, 512, 64, 20, 2
LR, MMT = 1e-3, 0.9


def _data():
    rng = np.random.RandomState(11)
    xs = rng.randn(NCHUNK, T, S, I).astype(np.float32)
    ods = (rng.randn(NCHUNK, T, S, R) * 0.1).astype(np.float32)
    return xs, ods


def _run(rank, world, port, q, c_path=False, fused=False):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import kaldi_lstm_b200 as klb
    torch.cuda.set_device(rank)
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    lo, hi = klb.parallel.shard_streams(S, rank, world)
    comp = klb.LstmProjectedStreams(I, R, device=rank, max_frames=T)
    comp.InitData("<CellDim> %d <NumStream> %d <ParamScale> 0.05" % (C, hi - lo), seed=3)
    comp.SetTrainOptions(klb.NnetTrainOptions(LR, MMT))
    # c_path: the engine's own entry point lstmp_b200_allreduce_grads_nccl on an own ncclComm_t and a side stream
    exchange = (klb.parallel.GradientExchange([comp], torch.device("cuda", rank), fused=fused)
                if (c_path and world > 1) else None)
    trainer = klb.parallel.StreamShardTrainer([comp], exchange=exchange)
    xs, ods = _data()
    for n in range(NCHUNK):
        x = torch.from_numpy(np.ascontiguousarray(xs[n][:, lo:hi]).reshape(-1, I)).cuda()
        od = torch.from_numpy(np.ascontiguousarray(ods[n][:, lo:hi]).reshape(-1, R)).cuda()
        trainer.train_chunk(x, lambda out: od)
    torch.cuda.synchronize()
    q.put((rank, comp.GetParams()))
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize("c_path", [False, True, "fused"])
def test_two_gpu_stream_sharding_matches_one_gpu(c_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    fused = c_path == "fused"   # block-wise exchange inside backpropagate (lstmp_b200_set_nccl)
    procs = [ctx.Process(target=_run, args=(r, 2, 29650 + int(bool(c_path)) + int(fused), q, bool(c_path), fused))
             for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    q1 = ctx.Queue()
    p = ctx.Process(target=_run, args=(0, 1, 0, q1))
    p.start()
    ref = q1.get(timeout=300)[1]
    p.join(timeout=60)
    np.testing.assert_array_equal(res[0], res[1])
    assert np.abs(res[0] - ref).max() <= 2e-5 * np.abs(ref).max()
