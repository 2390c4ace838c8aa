"""Development check: run one fused training step with the fp32-FMA kernels and one with the
tcgen05 kernels on the same weights / batch and compare every intermediate buffer and gradient."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def nerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    from _util import rand
    from sdvae_b200 import fixtures as fx, losses
    from sdvae_b200.engine import StepConfig, TrainEngine
    dev = 'cuda:0'
    tabs = fx.craniofacial_tables()
    bs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    engs = []
    for use_tc in (False, True):
        model = fx.build_model(tabs, 3, [32, 32, 32, 64], 75, False, True, 77, dev)
        lt = losses.LaplacianTable.build(*tabs.lap, tabs.num_vertices[0], dev)
        lat = tabs.latent_regions(75)
        eng = TrainEngine(model, lt, [r[1] for r in tabs.regions], [lat[k] for k in tabs.region_keys()],
                          StepConfig(batch_size=bs, lr=1e-3), use_graph=False, use_tc=use_tc)
        eng.set_fixed_eps(rand((bs * bs, 75), 100).to(dev))
        eng.load_batch(rand((bs, tabs.num_vertices[0], 3), 5).to(dev))
        eng.step(3, sync_losses=True)
        engs.append(eng)
    a, b = engs
    print('tc layers:', sorted(b.tc.keys()))
    for name in ('a', 'u', 'd', 'dd', 'du', 'da'):
        for i, (p, q) in enumerate(zip(getattr(a, name), getattr(b, name))):
            print('%s[%d] %s err %.3e' % (name, i, tuple(p.shape), nerr(q, p)))
    for name in ('recon', 'drecon', 'dh', 'dz', 'mu', 'logvar', 'da_raw'):
        print('%s err %.3e' % (name, nerr(getattr(b, name), getattr(a, name))))
    print('flat grad err %.3e' % nerr(b.flat_g, a.flat_g))
    print('losses', a.loss_dict(), b.loss_dict())


if __name__ == '__main__':
    main()
