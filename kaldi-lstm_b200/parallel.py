"""Data-parallel training over stream shards.

The path shards naturally over streams: GPU k of N owns NumStream/N streams (its own carried state,
activation record and utterance queue) and a full replica of the parameters.  The only exchange is ONE
sum all-reduce of the fresh gradient arena per Update(): the reference's gradients are plain sums over
all T*S rows (google/nnet/bd-nnet-lstm-projected-streams.h:468-487), so summing the per-shard gradients
reproduces the single-GPU S-stream gradient, after which every rank applies the identical
corr = G_sum + momentum*corr ; param -= lr*corr.  (All-reducing corr instead would scale the momentum
term by N.)  One process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the plumbing.
"""
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


class StreamShardTrainer:
    """One rank's view of a stack of LstmProjectedStreams layers: Reset / Propagate / Backpropagate /
    all-reduce / Update for one BPTT chunk -- the loop body of bd-nnet-train-lstm-streams.cc:209-229."""

    def __init__(self, layers, group=None):
        self.layers = list(layers)
        self.group = group

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
        allreduce_gradients(self.layers, self.group)
        for layer in self.layers:
            layer.Update()
        return acts[-1]
