#!/usr/bin/env python
"""BASELINE.json configs[0] on the GPU and on the host: SD-VAE encode+decode inference (eval mode, no grad) of the
craniofacial model -- the reference's own CPU-runnable case.

Inputs: the 12 demo OBJ meshes the reference ships (demo_files/meshes, staged into the git-ignored baseline/_ref by
tools/stage_reference.py), normalised with demo_files/norm.pt exactly as data_generation_and_loading.py does
((verts - mean) / std), batch 8 as craniofacial.yaml's demo run uses plus all 12 at once.  When the staged files are
absent the same shapes are filled with synthetic unit-normal vertices and the table says so.

Arms: this repo's drop-in ``model.py`` on one B200 (tcgen05 path) and the UNMODIFIED reference ``model.py`` on this
box's host cores (all threads), the same ``state_dict`` in both; the maximum |difference| of the two reconstructions
on those real meshes is printed under the table.  Larger batches and the decoder-only ``generate`` path
(model_manager.py:248-255, test.py's 10 000-sample diversity runs) follow.

usage: python tools/infer_bench.py [--no-cpu]"""
import glob
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
DEV = 'cuda:0'


def gpu_time(fn, n=20):
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


def cpu_time(fn, n=7, warm=2):
    ts = []
    for _ in range(n + warm):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts[warm:])) * 1e3


def read_obj_vertices(path):
    """'v x y z' lines of a Wavefront OBJ, in file order (what trimesh.load(process=False).vertices holds)."""
    vs = []
    with open(path) as f:
        for line in f:
            if line.startswith('v '):
                vs.append([float(t) for t in line.split()[1:4]])
    return np.asarray(vs, dtype=np.float32)


def demo_batch(V):
    from baseline import refarm
    mdir = os.path.join(refarm.STAGED, 'demo_files', 'meshes')
    files = sorted(glob.glob(os.path.join(mdir, '*.obj')))
    if not files:
        g = torch.Generator().manual_seed(0)
        return torch.randn(12, V, 3, generator=g), 'synthetic unit-normal vertices (demo meshes not staged)'
    norm = torch.load(os.path.join(refarm.STAGED, 'demo_files', 'norm.pt'))
    verts = torch.from_numpy(np.stack([read_obj_vertices(f) for f in files]))
    assert verts.shape[1:] == (V, 3), verts.shape
    x = (verts - norm['mean']) / norm['std']
    return x.float(), '%d demo OBJ meshes, normalised with norm.pt' % len(files)


def main():
    from sdvae_b200 import fixtures as fx
    from baseline import refarm
    tabs = fx.craniofacial_tables()
    model = fx.build_model(tabs, 3, [32, 32, 32, 64], 75, False, True, 0, DEV)
    model.eval()
    V = tabs.num_vertices[0]
    x_demo, what = demo_batch(V)
    print('# Config 0: encode+decode inference, craniofacial.yaml (V = %d), %s\n' % (V, what))
    print('| op | batch | input | where | ms | meshes/s |')
    print('|---|---|---|---|---|---|')
    recon_gpu = {}
    with torch.no_grad():
        for B in (8, 12):
            x = x_demo[:B].to(DEV)
            ms = gpu_time(lambda: model(x))
            recon_gpu[B] = model(x)[0].cpu() if isinstance(model(x), (tuple, list)) else model(x).cpu()
            print('| encode+decode | %d | demo meshes | 1xB200, drop-in model.py (tcgen05) | %.3f | %.0f |'
                  % (B, ms, B / ms * 1e3))
        for B in (256, 2048):
            x = torch.randn(B, V, 3, device=DEV)
            ms = gpu_time(lambda: model(x), 10)
            print('| encode+decode | %d | synthetic | 1xB200, drop-in model.py (tcgen05) | %.3f | %.0f |'
                  % (B, ms, B / ms * 1e3))
        for B in (8, 2048):
            z = torch.randn(B, 75, device=DEV)
            ms = gpu_time(lambda: model.decode(z), 10)
            print('| decode (generate) | %d | z ~ N(0,1) | 1xB200, drop-in model.py (tcgen05) | %.3f | %.0f |'
                  % (B, ms, B / ms * 1e3))
    # generate_for_opt (model_manager.py:253-255, Tester.fit_mesh test.py:395-421): decoder in train mode, loss on the
    # output, gradient w.r.t. z -- with the weights frozen the backward launches no weight-gradient kernel
    model.train()
    for freeze in (False, True):
        model.requires_grad_(not freeze)
        for B in (16, 256):
            z = torch.randn(B, 75, device=DEV, requires_grad=True)
            tgt = torch.randn(B, V, 3, device=DEV)

            def fit_step():
                z.grad = None
                ((model.decode(z) - tgt) ** 2).mean().backward()
            ms = gpu_time(fit_step, 10)
            print('| generate_for_opt: decode + backward to z%s | %d | z ~ N(0,1) | 1xB200, drop-in model.py (tcgen05) | %.3f | %.0f |'
                  % (' (weights frozen: no dW)' if freeze else '', B, ms, B / ms * 1e3))
    model.requires_grad_(True)
    model.eval()
    if '--no-cpu' in sys.argv:
        return
    ref = refarm.find_ref()
    torch.set_num_threads(os.cpu_count() or 1)
    params = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    sp, dn, up = tabs.spiral_tensors(), tabs.down_tensors(), tabs.up_tensors()
    diffs = []
    if ref is not None:
        # the unmodified reference model.py (torch_scatter shimmed, baseline/refarm.py) on the host cores
        mod = refarm.import_reference_module(ref, 'model', '_sdvae_reference_model')
        net = mod.Model(3, [32, 32, 32, 64], 75, sp, dn, up, False, True)
        net.load_state_dict(params)
        net.eval()
        with torch.no_grad():
            for B in (8, 12):
                x = x_demo[:B]
                ms = cpu_time(lambda: net(x))
                out = net(x)
                out = out[0] if isinstance(out, (tuple, list)) else out
                diffs.append((B, float((out - recon_gpu[B]).abs().max()), float(out.abs().max())))
                print('| encode+decode | %d | demo meshes | unmodified reference model.py, %d host cores | %.1f | %.0f |'
                      % (B, os.cpu_count() or 1, ms, B / ms * 1e3))
    else:
        from oracle import sdvae_oracle as orc
        net = orc.Net(3, [32, 32, 32, 64], 75, sp, dn, up, False, True)
        with torch.no_grad():
            x = x_demo[:8]
            ms = cpu_time(lambda: net.forward(params, x, training=False))
            print('| encode+decode | 8 | demo meshes | CPU oracle port, %d host cores | %.1f | %.0f |'
                  % (os.cpu_count() or 1, ms, 8 / ms * 1e3))
    print()
    for B, d, m in diffs:
        print('max |reconstruction(drop-in, B200) - reconstruction(reference, CPU)| on the %d demo meshes: %.3g '
              '(max |value| %.3g)' % (B, d, m))


if __name__ == '__main__':
    main()
