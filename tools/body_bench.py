#!/usr/bin/env python
"""BASELINE.json configs[3]: body.yaml model section (plain auto-encoder, sampling factors [4,4,4],
channels [32,32,64], latent 33, losses MSE + 1 latent-consistency + 1 Laplacian; body.yaml:19-24,39-49) on a
SYNTHETIC closed template with 6890 vertices (the STAR template is not in the reference mount; tables from
``fixtures.synthetic_tables`` with the same structure: spiral = self + 8 ring neighbours, down = sorted
selection, up = 3-nnz barycentric rows; 11 regions x 3 latents assumed).  Fused TrainEngine step, CUDA graph.

    python tools/body_bench.py [--bs 32] [--steps 20] [--warmup 5]
    torchrun --nproc-per-node N tools/body_bench.py ...          (data parallel over swap-grid rows)"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--bs', type=int, default=32)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--no-tc', action='store_true')
    args = ap.parse_args()
    world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
        pg = dist.group.WORLD
    from sdvae_b200 import fixtures as fx, losses
    from sdvae_b200.engine import StepConfig, TrainEngine
    tabs = fx.synthetic_tables(6890, 3, seq_length=9, n_regions=11, seed=0, name='body-synthetic')
    model = fx.build_model(tabs, 3, [32, 32, 64], 33, False, False, 3, dev)
    cfg = StepConfig(batch_size=args.bs, kl_weight=0.0, latent_consistency_weight=1.0, laplacian_weight=1.0)
    lt = losses.LaplacianTable.build(*tabs.lap, tabs.num_vertices[0], dev)
    lat = tabs.latent_regions(33)
    eng = TrainEngine(model, lt, [r[1] for r in tabs.regions], [lat[k] for k in tabs.region_keys()], cfg,
                      process_group=pg, use_tc=not args.no_tc)
    rng = np.random.RandomState(0)
    x = torch.from_numpy(rng.randn(args.bs, tabs.num_vertices[0], 3).astype(np.float32)).pin_memory()
    regions = [int(r) for r in rng.randint(0, len(tabs.regions), args.steps + args.warmup)]
    eng.load_batch(x)
    eng.prepare(sorted(set(regions)))
    for r in regions[:args.warmup]:
        eng.step(r)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for r in regions[args.warmup:]:
        eng.step(r)
    e.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    eng.step(regions[-1], sync_losses=True)
    if rank == 0:
        print(json.dumps({
            'metric': 'sdvae_train_meshes_per_sec', 'value': args.bs ** 2 * args.steps / (ms / 1e3), 'unit': 'meshes/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
            'config': {'workload': 'body.yaml model section on a synthetic 6890-vertex template '
                                   '(V = %s), global batch %d' % (tabs.num_vertices, args.bs ** 2),
                       'assumption': '11 regions x 3 latents; STAR template not in the reference mount'},
            'tensor_core_passes': sorted('%s:%s' % k for k in eng.tc), 'losses_last_step': eng.loss_dict()}))
    if world > 1:
        del eng
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
