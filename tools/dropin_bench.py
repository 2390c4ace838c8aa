#!/usr/bin/env python
"""Throughput of the DROP-IN path: the reference's own training-step composition
(model_manager.py:274-326: forward, MSE + KL + latent consistency + Laplacian, loss.backward(),
torch.optim.Adam.step()) on the drop-in ``model.py`` modules through autograd -- what an unchanged
``train.py`` gets -- with the tensor-core kernels on and off.  (``bench.py`` times the fused
``TrainEngine`` step instead.)

usage: python tools/dropin_bench.py [bs ...]      grid sides; meshes per step = bs^2 (reference yaml: 4)"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
DEV = 'cuda:0'


def main():
    from sdvae_b200 import fixtures as fx, functional, losses
    tabs = fx.craniofacial_tables()
    lat = tabs.latent_regions(75)
    keys = tabs.region_keys()
    lt = losses.LaplacianTable.build(*tabs.lap, tabs.num_vertices[0], DEV)
    sides = [int(a) for a in sys.argv[1:]] or [4, 16]
    print('| meshes/step | kernels | ms/step | meshes/s |')
    print('|---|---|---|---|')
    for bs in sides:
        rng = np.random.RandomState(0)
        x0 = torch.from_numpy(rng.randn(bs, tabs.num_vertices[0], 3).astype(np.float32))
        x = fx.swap_features_torch(x0, tabs.regions[3][1]).to(DEV)                 # [bs^2, V, 3]
        region = lat[keys[3]]
        for use_tc in (False, True):
            functional.set_tensor_cores(use_tc)
            model = fx.build_model(tabs, 3, [32, 32, 32, 64], 75, False, True, 7, DEV)
            model.train()
            opt = torch.optim.Adam(model.parameters(), lr=1e-4)

            def step():
                opt.zero_grad(set_to_none=True)
                rec, z, mu, lv = model(x)
                mse, lap = losses.mse_and_laplacian(rec, x, lt)
                tot = mse + 1e-4 * losses.kl_divergence(mu, lv) + 0.1 * lap \
                    + 0.5 * losses.latent_consistency(z, bs, region, 0.5, 0.5)
                tot.backward()
                opt.step()
                return tot
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            n = 10
            t0 = time.perf_counter()
            for _ in range(n):
                tot = step()
            float(tot.detach())
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n
            print('| %d | %s | %.2f | %.0f |' % (bs * bs, 'tcgen05 3xTF32' if use_tc else 'fp32 FMA', dt * 1e3,
                                               bs * bs / dt))
            del model, opt


if __name__ == '__main__':
    main()
