"""Host-side logic of the image sharding (world_size 2 over gloo on CPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dgvcc_b200.sharding import ShardedLoss, all_reduce_loss, sharded_mean_over_batch, snake_partition


def test_snake_partition_balances_and_covers():
    costs = [12000, 6472, 6290, 5222, 3249, 2889, 2619, 2267, 2190, 2137, 1214, 756, 667, 634, 553, 538]
    for world in (1, 2, 4, 8):
        shards = snake_partition(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(len(costs)))
        loads = [sum(costs[i] for i in s) for s in shards]
        if world == 2:
            assert max(loads) / (sum(loads) / world) < 1.15
    assert snake_partition([5, 1], 4) == [[0], [1], [], []]


class _MeanAbs(torch.nn.Module):
    """Stand-in with the BL contract: sum of per-image terms divided by global_batch."""
    global_batch = None

    def forward(self, x):
        return x.abs().sum() / float(self.global_batch or x.shape[0])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(1, 9, dtype=torch.float32) * torch.tensor([1, -1] * 4)
    mine = full[rank * 4:(rank + 1) * 4].clone().requires_grad_(True)
    loss = ShardedLoss(_MeanAbs(), global_batch=8)(mine)
    loss.backward()
    ref = full.clone().requires_grad_(True)
    ref_loss = ref.abs().sum() / 8
    ref_loss.backward()
    ok = torch.allclose(loss.detach(), ref_loss.detach()) and torch.equal(mine.grad, ref.grad[rank * 4:(rank + 1) * 4])
    # plain function form
    part = torch.tensor(float(rank + 1), requires_grad=True)
    tot = all_reduce_loss(part * 2)
    tot.backward()
    ok = ok and float(tot) == 6.0 and float(part.grad) == 2.0
    # a per-sample mean loss (ISW style) on unequal shards: rank 0 holds 3 samples, rank 1 holds 5
    samples = torch.arange(1, 9, dtype=torch.float32)
    mine2 = (samples[:3] if rank == 0 else samples[3:]).clone().requires_grad_(True)
    glob = sharded_mean_over_batch((mine2 ** 2).mean(), len(mine2), 8)
    glob.backward()
    ok = ok and abs(float(glob) - float((samples ** 2).mean())) < 1e-5
    ok = ok and torch.allclose(mine2.grad, 2 * mine2.detach() / 8)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_loss_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]
