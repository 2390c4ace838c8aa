#!/usr/bin/env python
"""BASELINE.json configs[0] on the GPU and on the host: SD-VAE encode+decode inference (eval mode, no grad)
of the craniofacial model through the drop-in ``model.py`` -- batch 8 (the reference's CPU-runnable case;
SURVEY.md 8d config 1), plus larger batches and the decoder-only ``generate`` path
(model_manager.py:248-255, test.py's 10 000-sample diversity runs).  The CPU column is the oracle port of
``model.py`` on this box's host cores (median of 7 after 2 warm-ups).

usage: python tools/infer_bench.py [--no-cpu]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
DEV = 'cuda:0'


def gpu_time(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    from sdvae_b200 import fixtures as fx
    tabs = fx.craniofacial_tables()
    model = fx.build_model(tabs, 3, [32, 32, 32, 64], 75, False, True, 0, DEV)
    model.eval()
    V = tabs.num_vertices[0]
    print('| op | batch | where | ms | meshes/s |')
    print('|---|---|---|---|---|')
    with torch.no_grad():
        for B in (8, 256, 2048):
            x = torch.randn(B, V, 3, device=DEV)
            ms = gpu_time(lambda: model(x))
            print('| encode+decode | %d | 1xB200, drop-in model.py (tcgen05) | %.3f | %.0f |' % (B, ms, B / ms * 1e3))
        for B in (8, 2048):
            z = torch.randn(B, 75, device=DEV)
            ms = gpu_time(lambda: model.decode(z))
            print('| decode (generate) | %d | 1xB200, drop-in model.py (tcgen05) | %.3f | %.0f |' % (B, ms, B / ms * 1e3))
    if '--no-cpu' not in sys.argv:
        # CPU-baseline leg: the oracle port of model.py with the same weights, on this box's host cores
        from oracle import sdvae_oracle as orc
        sp, dn, up = tabs.spiral_tensors(), tabs.down_tensors(), tabs.up_tensors()
        net = orc.Net(3, [32, 32, 32, 64], 75, sp, dn, up, False, True)
        params = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        torch.set_num_threads(os.cpu_count() or 1)
        x = torch.randn(8, V, 3)
        ts = []
        with torch.no_grad():
            for i in range(9):
                t0 = time.perf_counter()
                net.forward(params, x, training=False)
                ts.append(time.perf_counter() - t0)
        ms = float(np.median(ts[2:])) * 1e3
        print('| encode+decode | 8 | CPU oracle port, %d host cores | %.1f | %.0f |' % (os.cpu_count() or 1, ms, 8 / ms * 1e3))


if __name__ == '__main__':
    main()
