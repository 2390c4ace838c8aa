#!/usr/bin/env python
"""Generate tests/golden/* by EXECUTING THE REFERENCE in the build container.

Run where ``/root/reference`` exists (it does not exist on the GPU box):

    python tools/make_golden.py [--ref /root/reference]

Outputs (committed):
  craniofacialsd-vae_b200/data/craniofacial_tables.npz   index tables + Laplacian + regions derived from
                                         demo_files/{spirals.pkl,transforms.pkl,template.ply}
  tests/golden/reference_vectors.npz     outputs / losses / gradients of the reference's own
                                         ``model.py`` and of the loss functions lifted verbatim
                                         (via ``ast``, at run time, never copied into this repo)
                                         from ``model_manager.py`` / ``utils.py`` and of
                                         ``SwapFeatures`` from ``swap_batch_transform.py``.

The reference's un-vendored dependencies are shimmed exactly as SURVEY.md section 8c
describes: ``torch_scatter.scatter_add(src, index, dim, dim_size)`` =
``zeros(...).scatter_add_(dim, index broadcast, src)`` and a namespace stub for
``torch_geometric.data.Data``.
"""
import argparse
import ast
import importlib
import os
import random
import sys
import textwrap
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sdvae_b200 import fixtures as fx            # noqa: E402
from oracle import sdvae_oracle as orc           # noqa: E402


def install_shims():
    ts = types.ModuleType('torch_scatter')

    def scatter_add(src, index, dim=-1, out=None, dim_size=None):
        shape = list(src.shape)
        shape[dim] = int(dim_size)
        view = [1] * src.dim()
        view[dim] = -1
        return torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(
            dim, index.view(view).expand_as(src), src)
    ts.scatter_add = scatter_add
    sys.modules['torch_scatter'] = ts

    tg = types.ModuleType('torch_geometric')
    tgd = types.ModuleType('torch_geometric.data')

    class Data(types.SimpleNamespace):
        pass
    tgd.Data = Data
    tg.data = tgd
    sys.modules['torch_geometric'] = tg
    sys.modules['torch_geometric.data'] = tgd
    return Data


def lift(path, name, cls=None, extra=None):
    """Compile one function out of a reference file without importing the file."""
    src = open(path).read()
    tree = ast.parse(src)
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    node = next(n for n in body if isinstance(n, ast.FunctionDef) and n.name == name)
    code = textwrap.dedent('\n'.join(src.splitlines()[node.lineno - 1:node.end_lineno]))
    ns = {'torch': torch, 'np': np}
    ns.update(extra or {})
    exec(compile(code, path + ':' + name, 'exec'), ns)
    return ns[name]


def sample(t, stride=97):
    return t.detach().reshape(-1)[::stride].clone().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference')
    args = ap.parse_args()
    ref = args.ref
    demo = os.path.join(ref, 'demo_files')
    gold = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(gold, exist_ok=True)
    torch.set_num_threads(1)                     # make the stored fp32 vectors reproducible

    Data = install_shims()
    sys.path.insert(0, ref)
    ref_model = importlib.import_module('model')
    ref_swap = importlib.import_module('swap_batch_transform')
    sys.path.pop(0)

    mm = os.path.join(ref, 'model_manager.py')
    batch_mm = lift(os.path.join(ref, 'utils.py'), 'batch_mm')
    utils_ns = types.SimpleNamespace(batch_mm=batch_mm)
    ref_mse = lift(mm, 'compute_mse_loss', 'ModelManager')
    ref_kl = lift(mm, '_compute_kl_divergence_loss', 'ModelManager')
    ref_lap = lift(mm, '_compute_laplacian_regularizer', 'ModelManager', {'utils': utils_ns})
    ref_lc = lift(mm, '_compute_latent_consistency', 'ModelManager')

    # ---------------- tables ------------------------------------------------
    tabs = fx.tables_from_reference_files(os.path.join(demo, 'spirals.pkl'),
                                          os.path.join(demo, 'transforms.pkl'),
                                          os.path.join(demo, 'template.ply'))
    tabs.save_npz(fx.default_tables_path())
    out = {}

    # ---------------- case A: craniofacial, bs=2 -> 4 swapped meshes ---------
    norm = torch.load(os.path.join(demo, 'norm.pt'))
    names = sorted(os.listdir(os.path.join(demo, 'meshes')))[:2]
    raw = torch.stack([torch.from_numpy(fx.read_obj_vertices(os.path.join(demo, 'meshes', n)))
                       for n in names])
    x2 = ((raw - norm['mean']) / norm['std']).float()
    out['A_x_unswapped'] = x2.numpy()

    feat = {k: {'feature': idx.tolist()} for k, idx in tabs.regions}
    template = types.SimpleNamespace(feat_and_cont=feat)
    swapper = ref_swap.SwapFeatures(template)
    random.seed(7)
    batch = Data(x=x2, y=['a', 'b'], augmented=torch.zeros(2, 1), gender=['f', 'm'],
                 age=torch.ones(2, 1))
    swapped = swapper(batch)
    xa = swapped.x
    key = swapped.swapped
    out['A_swapped_key'] = np.array(key)
    out['A_x_swapped_sample'] = sample(xa)
    out['A_x_swapped_sum'] = np.array([float(xa.double().sum()), float(xa.double().abs().sum())])

    spirals = tabs.spiral_tensors()
    down, up = tabs.down_tensors(), tabs.up_tensors()
    net_ref = ref_model.Model(3, [32, 32, 32, 64], 75, spirals, down, up, False, True)
    onet = orc.Net(3, [32, 32, 32, 64], 75, spirals, down, up, False, True)
    shapes = onet.param_shapes()
    assert list(shapes) == list(net_ref.state_dict()), 'state-dict key order differs'
    params = orc.xavier_params(shapes, seed=1234, bias_scale=0.05)
    net_ref.load_state_dict(params, strict=True)

    net_ref.eval()
    with torch.no_grad():
        rec_e, z_e, mu_e, lv_e = net_ref(xa)
    out['A_eval_recon0'] = rec_e[0].numpy()
    out['A_eval_recon_sample'] = sample(rec_e)
    out['A_eval_mu'] = mu_e.numpy()
    out['A_eval_logvar'] = lv_e.numpy()

    net_ref.train()
    torch.manual_seed(11)
    eps = torch.randn_like(mu_e)
    torch.manual_seed(11)
    rec, z, mu, lv = net_ref(xa)
    assert torch.equal(z, mu + eps * torch.exp(0.5 * lv))
    out['A_eps'] = eps.numpy()
    out['A_train_z'] = z.detach().numpy()

    lap_t = tabs.laplacian_tensor()
    latent_regions = tabs.latent_regions(75)
    fake_self = types.SimpleNamespace(
        template=types.SimpleNamespace(laplacian=lap_t),
        _optimization_params={'batch_size': 2, 'latent_consistency_eta1': 0.5,
                              'latent_consistency_eta2': 0.5},
        _latent_regions=latent_regions)
    l_rec = ref_mse(rec, xa)
    l_lap = ref_lap(fake_self, rec)
    l_kl = ref_kl(mu, lv)
    l_lc = ref_lc(fake_self, z, key)
    w = {'kl': 1e-4, 'lc': 0.5, 'lap': 0.1}
    tot = l_rec + w['kl'] * l_kl + w['lc'] * l_lc + w['lap'] * l_lap      # model_manager.py:308-312
    tot.backward()
    out['A_losses'] = np.array([float(l_rec), float(l_kl), float(l_lc), float(l_lap), float(tot)],
                               np.float64)
    out['A_region'] = np.array(latent_regions[key])
    for k, prm in net_ref.named_parameters():
        g = prm.grad
        out['A_grad_sum/' + k] = np.array([float(g.double().sum()), float(g.double().abs().sum()),
                                           float(g.abs().max())])
        if g.numel() <= 9216:
            out['A_grad/' + k] = g.numpy().copy()
        else:
            out['A_grad_sample/' + k] = sample(g)

    # ---------------- case B: small synthetic, odd channel counts, plain AE --
    stab = fx.synthetic_tables(203, 2, seq_length=7, n_regions=3, seed=5)
    sp, dn, upm = stab.spiral_tensors(), stab.down_tensors(), stab.up_tensors()
    net_b = ref_model.Model(3, [8, 16], 6, sp, dn, upm, True, False)
    onet_b = orc.Net(3, [8, 16], 6, sp, dn, upm, True, False)
    pb = orc.xavier_params(onet_b.param_shapes(), seed=99, bias_scale=0.1)
    net_b.load_state_dict(pb, strict=True)
    rng = np.random.RandomState(3)
    xb = torch.from_numpy(rng.randn(9, 203, 3).astype(np.float32))
    net_b.train()
    rb, zb, mub, lvb = net_b(xb)
    assert lvb is None
    fake_b = types.SimpleNamespace(
        template=types.SimpleNamespace(laplacian=stab.laplacian_tensor()),
        _optimization_params={'batch_size': 3, 'latent_consistency_eta1': 0.3,
                              'latent_consistency_eta2': 0.7},
        _latent_regions=stab.latent_regions(6))
    kb = stab.region_keys()[1]
    lb = [ref_mse(rb, xb), ref_lc(fake_b, zb, kb), ref_lap(fake_b, rb)]
    totb = lb[0] + 1.0 * lb[1] + 1.0 * lb[2]
    totb.backward()
    out['B_x'] = xb.numpy()
    out['B_recon'] = rb.detach().numpy()
    out['B_z'] = zb.detach().numpy()
    out['B_losses'] = np.array([float(t) for t in lb] + [float(totb)], np.float64)
    out['B_region_key'] = np.array(kb)
    for k, prm in net_b.named_parameters():
        out['B_grad/' + k] = prm.grad.numpy().copy()

    # 2-D input path of SpiralConv (model.py:29-31) and the error message (model.py:36-39)
    conv = ref_model.SpiralConv(3, 5, sp[0])
    wc = torch.from_numpy(rng.uniform(-0.3, 0.3, (5, 21)).astype(np.float32))
    bc = torch.from_numpy(rng.uniform(-0.3, 0.3, (5,)).astype(np.float32))
    conv.load_state_dict({'layer.weight': wc, 'layer.bias': bc})
    out['B_conv2d_w'], out['B_conv2d_b'] = wc.numpy(), bc.numpy()
    out['B_conv2d_out'] = conv(xb[0]).detach().numpy()
    try:
        conv(xb.unsqueeze(0))
    except RuntimeError as e:
        out['B_conv_err'] = np.array(str(e))

    # Pool on both matrix kinds, straight from the reference function
    xp = torch.from_numpy(rng.randn(2, 4260, 32).astype(np.float32))
    out['P_x_seed'] = np.array(3)
    out['P_up0_sample'] = sample(ref_model.Pool(xp, up[0]), 101)
    xq = torch.from_numpy(rng.randn(2, 17039, 4).astype(np.float32))
    out['P_down0_sample'] = sample(ref_model.Pool(xq, down[0]), 11)
    out['P_xp_sample'], out['P_xq_sample'] = sample(xp, 1009), sample(xq, 1009)

    # latent-consistency loss on random latents for several grid sizes
    for bs, d, r0, r1 in ((2, 10, 2, 4), (3, 12, 0, 4), (4, 75, 70, 75), (5, 33, 9, 12)):
        zz = torch.from_numpy(rng.randn(bs * bs, d).astype(np.float32) * 0.7)
        fs = types.SimpleNamespace(
            _optimization_params={'batch_size': bs, 'latent_consistency_eta1': 0.5,
                                  'latent_consistency_eta2': 0.25},
            _latent_regions={'k': [r0, r1]})
        zz.requires_grad_(True)
        l = ref_lc(fs, zz, 'k')
        l.backward()
        tag = 'LC_%d_%d' % (bs, d)
        out[tag + '_z'] = zz.detach().numpy()
        out[tag + '_cfg'] = np.array([bs, r0, r1])
        out[tag + '_loss'] = np.array(float(l))
        out[tag + '_grad'] = zz.grad.numpy().copy()

    np.savez_compressed(os.path.join(gold, 'reference_vectors.npz'), **out)
    for f in sorted(os.listdir(gold)):
        print(f, os.path.getsize(os.path.join(gold, f)))


if __name__ == '__main__':
    main()
