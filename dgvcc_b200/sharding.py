"""Image sharding of the density-supervision path across the GPUs of one box.

Images are independent units in all three kernels: the Bayesian loss is sum_i L_i / B
(losses/bl.py:62-79), so rank r evaluates its own images with the GLOBAL batch size as divisor and
the only collective is one all-reduce(sum) of the scalar partial loss (NCCL over NVLink on GPUs,
gloo in the CPU tests).  Density gradients stay on the owning rank -- no data-path collective.
"""
import torch
import torch.distributed as dist


def snake_partition(costs, world_size):
    """Assign items to ranks by descending cost in boustrophedon order (balances sum of costs).

    Returns a list of index lists, one per rank.  Cost of a BL image is N_i * M_i pairs.
    """
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    shards = [[] for _ in range(world_size)]
    for pos, idx in enumerate(order):
        lap, off = divmod(pos, world_size)
        rank = off if lap % 2 == 0 else world_size - 1 - off
        shards[rank].append(idx)
    return [sorted(s) for s in shards]


def all_reduce_loss(partial, group=None):
    """Sum of the ranks' partial losses; d(total)/d(partial_local) = 1, so backward stays local."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return partial
    total = partial.detach().clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return partial + (total - partial.detach())


class ShardedLoss(torch.nn.Module):
    """Wrap a per-image-mean loss module (e.g. ``BL``) for image-partitioned data parallelism.

    ``loss_module.global_batch`` is set to the global number of images so that every rank divides
    by the same B (bl.py:79); ``forward`` returns the all-reduced loss on every rank.
    """

    def __init__(self, loss_module, global_batch, group=None):
        super().__init__()
        self.loss_module = loss_module
        self.group = group
        loss_module.global_batch = int(global_batch)

    def forward(self, *args, **kwargs):
        return all_reduce_loss(self.loss_module(*args, **kwargs), self.group)


def sharded_mean_over_batch(local_mean, local_batch, global_batch, group=None):
    """All-reduced value of a loss that is a MEAN over samples, given each rank's mean over its own samples
    (e.g. ``instance_whitening_loss``, which divides by the local B, instance_whitening.py:25):
    ``sum_ranks(local_mean * B_local) / B_global``.  Gradients stay local and are scaled by B_local / B_global."""
    return all_reduce_loss(local_mean * (float(local_batch) / float(global_batch)), group)
