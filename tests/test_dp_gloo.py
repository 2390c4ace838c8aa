"""World-size-2 check (gloo, CPU) of the data-parallel recipe TrainEngine implements
(engine.py docstring): grid-row sharding, globally normalised mean losses,
latent consistency on the all-gathered z with only the local slice back-propagated, SUM
all-reduce of the gradients == the single-process gradient."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdvae_b200 import fixtures as fx
from sdvae_b200 import parallel as par

W = dict(kl=1e-3, lc=0.5, lap=0.1, eta1=0.5, eta2=0.5)
BS, LATENT = 4, 9


def _problem():
    from oracle import sdvae_oracle as orc
    tabs = fx.synthetic_tables(203, 2, seq_length=7, n_regions=3, seed=5)
    net = orc.Net(3, [8, 16], LATENT, tabs.spiral_tensors(), tabs.down_tensors(), tabs.up_tensors(),
                  False, True)
    params = orc.xavier_params(net.param_shapes(), seed=3, bias_scale=0.05)
    rng = np.random.RandomState(0)
    x = torch.from_numpy(rng.randn(BS, 203, 3).astype(np.float32))
    eps = torch.from_numpy(rng.randn(BS * BS, LATENT).astype(np.float32))
    xa = orc.swap_features(x, torch.from_numpy(tabs.regions[1][1]))
    region = tabs.latent_regions(LATENT)[tabs.region_keys()[1]]
    lap = tuple(torch.from_numpy(a) for a in tabs.lap)
    return orc, net, params, xa, eps, region, lap


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    orc, net, params, xa, eps, region, lap = _problem()
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    lo, hi = par.local_meshes(BS, world, rank)
    scale = par.mean_loss_scale(world)
    rec, z, mu, lv = net.forward(p, xa[lo:hi], training=True, eps=eps[lo:hi])
    # all-gather z; the local block stays attached to the graph, remote blocks are constants
    blocks = [torch.empty_like(z) for _ in range(world)]
    dist.all_gather(blocks, z.detach())
    blocks[rank] = z
    z_all = torch.cat(blocks, 0)
    loss = scale * (orc.mse_loss(rec, xa[lo:hi]) + W['kl'] * orc.kl_loss(mu, lv)
                    + W['lap'] * orc.laplacian_loss(rec, *lap)) \
        + W['lc'] * orc.latent_consistency_loss(z_all, BS, region[0], region[1], W['eta1'], W['eta2'])
    loss.backward()
    flat = torch.cat([v.grad.reshape(-1) for v in p.values()])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if rank == 0:
        torch.save(flat, os.path.join(out_dir, 'dp_grad.pt'))
    dist.barrier()
    dist.destroy_process_group()


def test_grid_row_sharding_rules():
    assert par.grid_rows(32, 8, 3) == (12, 16)
    assert par.local_meshes(32, 8, 3) == (384, 512)
    assert par.local_meshes(32, 2, 1) == (512, 1024)          # 512 is not a square: never re-square
    assert par.mean_loss_scale(4) == 0.25
    with pytest.raises(ValueError):
        par.grid_rows(6, 4, 0)
    with pytest.raises(ValueError):
        par.grid_rows(4, 2, 2)


def test_dp_recipe_matches_single_process(tmp_path):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    dp = torch.load(os.path.join(str(tmp_path), 'dp_grad.pt'))
    orc, net, params, xa, eps, region, lap = _problem()
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    rec, z, mu, lv = net.forward(p, xa, training=True, eps=eps)
    tot, _ = orc.total_loss(rec, xa, z, mu, lv, lap, BS, region, W)
    tot.backward()
    ref = torch.cat([v.grad.reshape(-1) for v in p.values()])
    err = float((dp - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err
