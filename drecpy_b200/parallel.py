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


PREP, GRADS_A, GRADS_B, GRADS_C, UPDATE_V, UPDATE_W2T, UPDATE_REST = 1, 2, 8, 16, 128, 256, 512   # include/drb.h


def data_parallel_step(dp, run_phase, label_count, grads, off_w, off_v, dz1_rows, uids, rows_all, uids_all,
                       add_user_rows):
    """One data-parallel CDAE step: the phases of drb_cdae_step_phases with the collectives between them.

    The collectives run on the communication stream while the next phase computes, and every part of the update starts
    as soon as ITS gradient is final:
      PREP | all-reduce(label histogram) || GRADS_A | GRADS_B | all-reduce(dW'^T) || GRADS_C |
      all-gather(dz1 rows, uids) ; all-reduce(dW, db, db') || add user rows -> UPDATE_V |
      wait dW'^T -> UPDATE_W2T | wait dW, db, db' -> UPDATE_REST (+ loss)
    run_phase(mask) runs the given phases on this rank's slice of the batch; label_count is None for per-user labels;
    grads is the gradient arena ([0, off_w) = dW'^T, [off_w, off_v) = dW, db, db', [off_v, ..) = dV); dz1_rows / uids
    are this rank's B x ld hidden-layer gradient rows and their user ids, rows_all / uids_all the gather buffers;
    add_user_rows(uids_all, rows_all) adds the gathered rows into dV (rows of V never travel).  The same function
    drives the native phases on the GPU and an oracle-backed stand-in in the world-size-2 gloo test."""
    dist = dp.dist
    run_phase(PREP)
    h_lab = None
    if label_count is not None:
        h_lab = dist.all_reduce(label_count, op=dist.ReduceOp.SUM, group=dp.group, async_op=True)
    run_phase(GRADS_A)
    if h_lab is not None:
        h_lab.wait()
    run_phase(GRADS_B)
    h_w2t = dist.all_reduce(grads[:off_w], op=dist.ReduceOp.SUM, group=dp.group, async_op=True)
    run_phase(GRADS_C)
    h_rows = dist.all_gather_into_tensor(rows_all, dz1_rows, group=dp.group, async_op=True)
    h_uids = dist.all_gather_into_tensor(uids_all, uids, group=dp.group, async_op=True)
    h_rest = dist.all_reduce(grads[off_w:off_v], op=dist.ReduceOp.SUM, group=dp.group, async_op=True)
    h_rows.wait()
    h_uids.wait()
    add_user_rows(uids_all, rows_all)
    run_phase(UPDATE_V)              # 72 % of the Adam bytes at the ml-20m shape, under the last all-reduce
    h_w2t.wait()
    run_phase(UPDATE_W2T)
    h_rest.wait()
    run_phase(UPDATE_REST)


def shard_slices(n_global, world):
    per = n_global // world
    return [(r * per, (r + 1) * per) for r in range(world)]
