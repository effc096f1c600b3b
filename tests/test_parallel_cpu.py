"""World-size-2 gloo test of the batch-sharding host logic (no GPU): shards partition the batch, gather restores
the original order, uneven batches work."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from transformer_latent_diffusion_b200.parallel import shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 5, 8, 64, 67):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


class _FakeGenerator:
    """stands in for DiffusionGenerator: 'sampling' is a deterministic function of (label, seed) so ordering shows"""

    def generate_latents(self, labels, num_imgs, seeds, **kw):
        assert labels.shape[0] == num_imgs == seeds.shape[0]
        return seeds * 2 + labels[:, :1, None, None]


def _worker(rank, world, port, n_imgs, out):
    import torch.distributed as dist

    from transformer_latent_diffusion_b200.parallel import generate_sharded

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    labels = torch.randn(n_imgs, 8, generator=g)
    seeds = torch.randn(n_imgs, 4, 2, 2, generator=g)
    res = generate_sharded(_FakeGenerator(), labels, seeds, dst=0)
    if rank == 0:
        torch.save((res, seeds * 2 + labels[:, :1, None, None]), out)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_imgs", [6, 5])
def test_generate_sharded_gloo_world2(tmp_path, n_imgs):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, port, n_imgs, out), nprocs=2, join=True)
    res, ref = torch.load(out)
    assert torch.equal(res, ref)


def _grad_worker(rank, world, port, out):
    import torch.distributed as dist

    from transformer_latent_diffusion_b200.train import allreduce_gradients

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.LayerNorm(7), torch.nn.Linear(7, 3))
    for i, p in enumerate(m.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    allreduce_gradients(m)
    # a model whose backward already averaged its gradients (overlapped NCCL path, train.py:_overlapped_allreduce) is
    # left alone exactly once, and without NCCL the overlapped path declines so the plain all-reduce above still runs
    from transformer_latent_diffusion_b200.train import _overlapped_allreduce

    m2 = torch.nn.Linear(2, 2)
    for p in m2.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    m2._tld_grads_allreduced = True
    allreduce_gradients(m2)
    skipped = all(torch.equal(p.grad, torch.full_like(p, float(rank + 1))) for p in m2.parameters())
    allreduce_gradients(m2)   # flag consumed: now it averages
    averaged = all(torch.allclose(p.grad, torch.full_like(p, 1.5)) for p in m2.parameters())
    declined = _overlapped_allreduce(m2, None, torch.device("cpu")) is False
    if rank == 0:
        torch.save([p.grad.clone() for p in m.parameters()] + [torch.tensor([skipped, averaged, declined])], out)
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_gradients_gloo_world2(tmp_path):
    """the data-parallel gradient average of the training step (tld/train.py:169 via DDP) on 2 CPU ranks"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "g.pt")
    mp.spawn(_grad_worker, args=(2, port, out), nprocs=2, join=True)
    *grads, flags = torch.load(out)
    assert flags.all(), flags
    for i, g in enumerate(grads):
        assert torch.allclose(g, torch.full_like(g, 1.5 * (i + 1)))  # mean of (1, 2) * (i + 1)


@pytest.mark.parametrize("n,world", [(65, 2), (64, 2), (7, 4), (3, 8), (100, 3)])
def test_shard_indices_equal_length_like_distributed_sampler(n, world):
    """every rank gets ceil(n / world) indices (wrap-around padding), so all ranks run the same number of steps; identical
    to torch.utils.data.DistributedSampler(shuffle=False) over the same permutation"""
    from torch.utils.data import DistributedSampler

    from transformer_latent_diffusion_b200.train import shard_indices

    perm = torch.randperm(n, generator=torch.Generator().manual_seed(n))
    shards = [shard_indices(perm, r, world) for r in range(world)]
    assert len({len(s) for s in shards}) == 1 and len(shards[0]) == -(-n // world)
    assert set(torch.cat(shards).tolist()) == set(range(n))
    for r in range(world):
        ds = DistributedSampler(list(range(n)), num_replicas=world, rank=r, shuffle=False)
        assert [int(perm[i]) for i in ds] == shards[r].tolist()


def test_make_image_grid_single_image_is_unpadded():
    """torchvision.utils.make_grid returns one image as it is (no padding frame): the app's default num_imgs=1"""
    from torchvision.utils import make_grid

    from transformer_latent_diffusion_b200.diffusion import make_image_grid

    g = torch.Generator().manual_seed(0)
    for B, nrow in [(1, 1), (1, 4), (5, 2), (4, 4)]:
        img = torch.rand(B, 3, 8, 6, generator=g) * 2.4 - 1.2
        ref = make_grid((img + 1) / 2, nrow=nrow, padding=4).clip(0, 1)
        assert torch.equal(make_image_grid(img, nrow, 4), ref), (B, nrow)
