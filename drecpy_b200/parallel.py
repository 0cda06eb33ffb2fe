"""Data parallelism over user mini-batches (one process per GPU, torch.distributed / NCCL over NVLink).

The reference is single-process (no distributed code exists), so this layer is new.  The
training step shards by user mini-batch with replicated weights; the exchange steps are
  1. all-reduce of the batch label histogram (the reference's loss uses batch-mean labels, SURVEY.md Q1, so the
     mean must be taken over the GLOBAL batch) -- n_items floats,
  2. all-reduce of the gradient arena,
  3. all-reduce of the scalar batch loss.
Every rank then applies the identical dense Adam update, so weights stay bit-identical across ranks.
With dp_sampler='replay' every rank replays the same PointSampler stream for the global batch and takes its own
contiguous slice, which makes an N-rank run consume exactly the pairs a single-process run with batch N*B would.
"""
import numpy as np


class DataParallel:
    def __init__(self, dist=None, group=None):
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist is not None else 0
        self.world = dist.get_world_size(group) if dist is not None else 1

    @property
    def active(self):
        return self.world > 1

    def shard(self, n_global):
        """[lo, hi) of this rank's contiguous slice of a global batch of n_global rows (n_global % world == 0)."""
        assert n_global % self.world == 0, 'global batch must be divisible by the number of ranks'
        per = n_global // self.world
        return self.rank * per, (self.rank + 1) * per

    def all_reduce_sum(self, tensor):
        if self.active:
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM, group=self.group)
        return tensor

    def global_loss(self, loss2):
        """loss2 = [reported loss of this rank, its batch term] -> reported loss of the global batch."""
        if not self.active:
            return loss2[0]
        batch = loss2[1:2].clone()
        self.all_reduce_sum(batch)
        return batch[0] + (loss2[0] - loss2[1])


def shard_slices(n_global, world):
    per = n_global // world
    return [(r * per, (r + 1) * per) for r in range(world)]
