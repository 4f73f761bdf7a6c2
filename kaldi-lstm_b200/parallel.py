"""Data-parallel training over stream shards.

The path shards naturally over streams: GPU k of N owns NumStream/N streams (its own carried state,
activation record and utterance queue) and a full replica of the parameters.  The only exchange is ONE
sum all-reduce of the fresh gradient arena per layer and Update(): the reference's gradients are plain sums
over all T*S rows (google/nnet/bd-nnet-lstm-projected-streams.h:468-487), so summing the per-shard gradients
reproduces the single-GPU S-stream gradient, after which every rank applies the identical
corr = G_sum + momentum*corr ; param -= lr*corr.  (All-reducing corr instead would scale the momentum
term by N.)  One process per GPU; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the bootstrap.

Two exchange back ends:
  * `allreduce_gradients` -- torch.distributed on the current stream (works with gloo: the CPU tests);
  * `GradientExchange`    -- the engine's own C entry point (lstmp_b200_allreduce_grads_nccl) on an own NCCL
    communicator and a side stream: layer k's all-reduce is issued the moment its weight-gradient GEMMs are done
    (it is final before layer k-1's backward starts) and only that layer's Update() waits for it.

Lock-step termination: ranks own different utterances and run out of data after different numbers of chunks,
but every chunk carries a collective.  `StreamShardTrainer.run` all-reduces a have-data flag per step; a rank
that has run out keeps stepping on an all-padding chunk (zero features, zero loss gradient: a zero
contribution to the gradient sum) until every rank is done.
"""
import ctypes

import torch
import torch.distributed as dist


def allreduce_gradients(layers, group=None):
    """layers: objects with fresh_gradient() -> 1-D tensor viewing the fresh-gradient arena (in place)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for layer in layers:
        dist.all_reduce(layer.fresh_gradient(), op=dist.ReduceOp.SUM, group=group)


def shard_streams(num_stream_total, rank, world):
    """Streams [lo, hi) owned by `rank`; NumStream must divide evenly (configs[3]: 256 -> 32 per GPU)."""
    if num_stream_total % world != 0:
        raise ValueError("NumStream=%d is not divisible by %d ranks" % (num_stream_total, world))
    per = num_stream_total // world
    return rank * per, (rank + 1) * per


class _NcclUniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_char * 128)]


class NcclCommunicator:
    """An own ncclComm_t (ncclCommInitRank through ctypes on the NCCL library torch already loaded), bootstrapped
    over torch.distributed: rank 0's unique id is broadcast through the existing process group.  `ptr` is what
    the C ABI takes (lstmp_b200_allreduce_grads_nccl, include/lstmp_b200.h)."""

    def __init__(self, device, group=None):
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._lib = None
        for name in ("libnccl.so.2", "libnccl.so"):
            try:
                self._lib = ctypes.CDLL(name, mode=ctypes.RTLD_GLOBAL)
                break
            except OSError:
                continue
        if self._lib is None:
            raise RuntimeError("libnccl.so.2 not found")
        L = self._lib
        L.ncclGetUniqueId.argtypes = [ctypes.POINTER(_NcclUniqueId)]
        L.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _NcclUniqueId, ctypes.c_int]
        L.ncclCommDestroy.argtypes = [ctypes.c_void_p]
        uid = _NcclUniqueId()
        if self.rank == 0 and L.ncclGetUniqueId(ctypes.byref(uid)) != 0:
            raise RuntimeError("ncclGetUniqueId failed")
        t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone()
        backend = dist.get_backend(group)
        if backend == "nccl":
            t = t.to(device)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ctypes.memmove(ctypes.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
        comm = ctypes.c_void_p()
        with torch.cuda.device(device):
            rc = L.ncclCommInitRank(ctypes.byref(comm), self.world, uid, self.rank)
        if rc != 0 or not comm.value:
            raise RuntimeError("ncclCommInitRank returned %d" % rc)
        self.ptr = comm.value

    def close(self):
        if getattr(self, "ptr", None):
            self._lib.ncclCommDestroy(ctypes.c_void_p(self.ptr))
            self.ptr = None


class GradientExchange:
    """Overlapped per-layer gradient all-reduce through the engine's C entry point on a side stream."""

    def __init__(self, layers, device, group=None, fused=False):
        self.layers = list(layers)
        self.device = torch.device(device)
        self.comm = NcclCommunicator(self.device, group)
        self.stream = torch.cuda.Stream(device=self.device)
        self._ready = [torch.cuda.Event() for _ in self.layers]   # layer k's gradient is final (compute stream)
        self._done = [torch.cuda.Event() for _ in self.layers]    # layer k's all-reduce finished (side stream)
        # engines that can run the exchange themselves do it block by block INSIDE backpropagate (each gradient block is
        # summed under the GEMMs that follow it, lstmp_b200_set_nccl) and wait for it in update; start / finish are
        # then no-ops for that layer
        # fused=True (or LSTMP_B200_FUSED_EXCHANGE=1): block-wise exchange inside backpropagate.  Measured on 2 B200s
        # (cfg3): 1.286 ms per step against 1.277 ms for one all-reduce per layer after its Backpropagate -- three
        # smaller collectives cost more launch latency and SM contention with the gradient GEMMs than they hide -- so
        # the per-layer exchange is the default (profiles/r2_exchange_ab.txt).
        import os
        if os.environ.get("LSTMP_B200_FUSED_EXCHANGE") is not None:
            fused = os.environ["LSTMP_B200_FUSED_EXCHANGE"] != "0"
        self._fused = []
        for layer in self.layers:
            eng = layer.engine
            ok = bool(fused) and hasattr(eng, "set_nccl")
            if ok:
                eng.set_nccl(self.comm.ptr, self.stream.cuda_stream)
            self._fused.append(ok)

    def start(self, k):
        """Call right after layer k's Backpropagate was enqueued on the current stream."""
        if self._fused[k]:
            return
        self._ready[k].record(torch.cuda.current_stream(self.device))
        self.stream.wait_event(self._ready[k])
        self.layers[k].engine.allreduce_grads_nccl(self.comm.ptr, self.stream.cuda_stream)
        self._done[k].record(self.stream)

    def finish(self, k):
        """Call before layer k's Update(): the current stream waits for that layer's all-reduce only."""
        if self._fused[k]:
            return
        torch.cuda.current_stream(self.device).wait_event(self._done[k])

    def close(self):
        for layer, f in zip(self.layers, self._fused):
            if f:
                try:
                    layer.engine.set_nccl(None, None)
                except Exception:
                    pass
        torch.cuda.synchronize(self.device)
        self.comm.close()


class StreamShardTrainer:
    """One rank's view of a stack of LstmProjectedStreams layers: Reset / Propagate / Backpropagate /
    all-reduce / Update for one BPTT chunk -- the loop body of bd-nnet-train-lstm-streams.cc:209-229."""

    def __init__(self, layers, group=None, exchange=None):
        self.layers = list(layers)
        self.group = group
        self.exchange = exchange   # GradientExchange (GPU) or None (torch.distributed on the current stream)

    def train_chunk(self, feats, out_diff_fn, reset_flags=None):
        """feats: [T*S_local x I] device matrix; out_diff_fn(top_output) -> [T*S_local x R_top] gradient."""
        acts = [feats]
        for layer in self.layers:
            if reset_flags is not None:
                layer.Reset(reset_flags)
            acts.append(layer.Propagate(acts[-1]))
        d = out_diff_fn(acts[-1])
        for li in reversed(range(len(self.layers))):
            d = self.layers[li].Backpropagate(acts[li], acts[li + 1], d, want_in_diff=(li > 0))
            if self.exchange is not None:
                self.exchange.start(li)
        if self.exchange is None:
            allreduce_gradients(self.layers, self.group)
        for li, layer in enumerate(self.layers):
            if self.exchange is not None:
                self.exchange.finish(li)
            layer.Update()
        return acts[-1]

    def _any_rank_has_data(self, have):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return bool(have)
        t = torch.tensor([1 if have else 0], dtype=torch.int32)
        if dist.get_backend(self.group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return bool(int(t.item()))

    def run(self, next_chunk, out_diff_fn, padding_chunk):
        """Lock-step loop over this rank's shard.  next_chunk() -> (feats, aux, reset_flags) or None when the shard is
        exhausted; out_diff_fn(top_output, aux) -> loss gradient; padding_chunk() -> the same triple for an
        all-padding chunk (mask 0 everywhere) used once this rank has run out while others still have data.
        Returns (chunks with data, padding chunks)."""
        n_data = n_pad = 0
        while True:
            chunk = next_chunk()
            if not self._any_rank_has_data(chunk is not None):
                return n_data, n_pad
            if chunk is None:
                chunk = padding_chunk()
                n_pad += 1
            else:
                n_data += 1
            feats, aux, flags = chunk
            self.train_chunk(feats, lambda out: out_diff_fn(out, aux), flags)
