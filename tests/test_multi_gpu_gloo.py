"""CPU, world_size 2 over gloo: the multi-rank plumbing of the sampling path (independent batch
shards + one broadcast of rank 0's constants) and of data-parallel calibration (flat SUM all-reduce,
delta averaging, per-interval data sharding)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tfmq_b200 import dist_utils as D
    ok = True
    # batches shard contiguously and exhaustively
    mine = list(D.shard_range(7, rank, world))
    got = [None] * world
    dist.all_gather_object(got, mine)
    ok &= sorted(sum(got, [])) == list(range(7))
    # constants broadcast from rank 0
    t = [torch.full((5,), float(rank + 1)), torch.arange(4, dtype=torch.int32) * (rank + 1)]
    D.broadcast_tensors(t, 0)
    ok &= bool((t[0] == 1).all()) and t[1].tolist() == [0, 1, 2, 3]
    # flat-bucket SUM all-reduce (alpha gradients) keeps shapes and sums over ranks
    g = [torch.ones(2, 3) * (rank + 1), torch.ones(4) * 10 * (rank + 1)]
    r = D.allreduce_flat_(g)
    ok &= r[0].shape == (2, 3) and bool((r[0] == 3).all()) and bool((r[1] == 30).all())
    # activation deltas are averaged
    d = [torch.tensor(float(rank)), torch.tensor(2.0 + rank)]
    D.allaverage_(d)
    ok &= abs(d[0].item() - 0.5) < 1e-6 and abs(d[1].item() - 2.5) < 1e-6
    # calibration data: each rank takes its 1/world slice of every timestep interval
    idx = D.shard_interval_indices(8, 4, rank, world).tolist()
    ok &= idx == ([0, 1, 4, 5] if rank == 0 else [2, 3, 6, 7])
    # runner: the rounds of `sample_batches` are split over the ranks (no collective); the union is the single-rank result
    from types import SimpleNamespace as NS
    from tfmq_b200 import runners as R
    cfg = NS(model=NS(var_type="fixedlarge"), data=NS(channels=3, image_size=4, rescaled=True, logit_transform=False),
             sampling=NS(batch_size=2), diffusion=NS(beta_schedule="linear", beta_start=1e-4, beta_end=0.02,
                                                       num_diffusion_timesteps=1000))
    run = R.Diffusion(NS(skip_type="uniform", timesteps=10, sample_type="generalized", eta=0.0), cfg, device="cpu")
    rounds = []

    def fake_sample_image(x, model, **kw):          # stands in for the GPU sampler: marks every image with its round
        rounds.append(len(rounds))
        return torch.full_like(x, -1.0 + 0.25 * (model["first"] + len(rounds) - 1)), None, None
    run.sample_image = fake_sample_image
    first = list(D.shard_range(4, rank, world))[0]                  # 7 images in batches of 2 = 4 rounds
    imgs = run.sample_batches({"first": first}, total=7)
    allr = [None] * world
    dist.all_gather_object(allr, imgs[:, 0, 0, 0].tolist())
    ok &= imgs.dtype == torch.zeros(1).numpy().astype("uint8").dtype and imgs.shape[1:] == (4, 4, 3)
    ok &= sum(allr, []) == [0, 0, 32, 32, 64, 64, 96]               # round r -> value round(255 * r / 8); last round keeps 1
    # with an identically seeded generator the union over the ranks is bit-identical to the single-rank run
    run.sample_image = lambda x, model, **kw: (torch.tanh(x), None, None)
    part = run.sample_batches(None, total=7, generator=torch.Generator().manual_seed(11))
    parts = [None] * world
    dist.all_gather_object(parts, part)
    saved = (D.rank, D.world)
    D.rank, D.world = (lambda: 0), (lambda: 1)
    single = run.sample_batches(None, total=7, generator=torch.Generator().manual_seed(11))
    D.rank, D.world = saved
    import numpy as np
    ok &= single.shape == (7, 4, 4, 3) and np.array_equal(np.concatenate(parts, axis=0), single) and len(part) in (3, 4)
    ret[rank] = ok
    dist.destroy_process_group()


def test_two_rank_plumbing():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
