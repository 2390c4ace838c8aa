"""Reference arm: drive the UNMODIFIED reference implementation of the hot path (model.py, the four loss methods of
model_manager.py, SwapFeatures of swap_batch_transform.py) -- none of this repo's kernels, engine or oracle on that
path.  Used by ``bench.py --impl reference`` / ``bench.py``'s ``cpu_baseline`` and ``reference_eager_cuda`` legs and
by the drop-in tests.

The reference is flat Python scripts (no setup.py / pyproject.toml, so ``pip install`` has nothing to install); it is
STAGED file by file into the git-ignored ``baseline/_ref/`` by ``tools/stage_reference.py`` (run by
``__graft_entry__.build()`` wherever /root/reference exists), from where it travels to the GPU box.  Nothing under
``baseline/_ref`` is committed.

Un-vendored dependencies of those files are shimmed exactly as SURVEY.md 8c describes:
``torch_scatter.scatter_add(src, index, dim, dim_size)`` = ``zeros(...).scatter_add_(dim, index broadcast, src)``
(what torch_scatter >= 2.0's scatter_sum does) and a namespace stub for ``torch_geometric.data.Data``.
``model_manager.py`` itself cannot be imported (trimesh, pytorch3d, torch_geometric, ...): its loss methods are
compiled out of the staged file with ``ast`` at run time, and ``_do_iteration``'s composition
(model_manager.py:274-326: forward, four losses, weighted sum :308-312, backward, Adam) is the few lines of
``ReferenceStep.step``.
"""
import ast
import importlib
import os
import pickle
import random
import sys
import textwrap
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, 'baseline', '_ref')
FILES = ['model.py', 'model_manager.py', 'swap_batch_transform.py', 'utils.py']
DEMO = ['spirals.pkl', 'transforms.pkl', 'template.ply', 'norm.pt', 'region_ldas.pkl']


def find_ref():
    """Reference root: $SDVAE_REF, /root/reference, baseline/_ref -- the first that holds model.py; else None."""
    for cand in (os.environ.get('SDVAE_REF'), '/root/reference', STAGED):
        if cand and os.path.exists(os.path.join(cand, 'model.py')):
            return cand
    return None


def install_shims():
    if 'torch_scatter' not in sys.modules:
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
    if 'torch_geometric' not in sys.modules:
        tg = types.ModuleType('torch_geometric')
        tgd = types.ModuleType('torch_geometric.data')

        class Data(types.SimpleNamespace):
            pass
        tgd.Data = Data
        tg.data = tgd
        sys.modules['torch_geometric'] = tg
        sys.modules['torch_geometric.data'] = tgd
    return sys.modules['torch_geometric.data'].Data


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


def import_reference_module(ref, name, alias):
    """Import ``<ref>/<name>.py`` under the module name ``alias`` (so it never shadows, or is shadowed by, this repo's
    drop-in ``model``)."""
    install_shims()
    spec = importlib.util.spec_from_file_location(alias, os.path.join(ref, name + '.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class ReferenceLosses:
    """The reference's own loss methods (model_manager.py:332-334, 343-349, 351-354, 360-393 + utils.batch_mm,
    utils.py:153-165), lifted from the files under ``ref``; ``self_ns`` stands in for the ModelManager instance."""

    def __init__(self, ref, laplacian, latent_regions, batch_size, eta1, eta2):
        mm = os.path.join(ref, 'model_manager.py')
        batch_mm = lift(os.path.join(ref, 'utils.py'), 'batch_mm')
        utils_ns = types.SimpleNamespace(batch_mm=batch_mm)
        self.mse = lift(mm, 'compute_mse_loss', 'ModelManager')
        self.l1 = lift(mm, '_compute_l1_loss', 'ModelManager')
        self.kl = lift(mm, '_compute_kl_divergence_loss', 'ModelManager')
        self._lap = lift(mm, '_compute_laplacian_regularizer', 'ModelManager', {'utils': utils_ns})
        self._lc = lift(mm, '_compute_latent_consistency', 'ModelManager')
        self.self_ns = types.SimpleNamespace(
            template=types.SimpleNamespace(laplacian=laplacian),
            _optimization_params={'batch_size': batch_size, 'latent_consistency_eta1': eta1,
                                  'latent_consistency_eta2': eta2},
            _latent_regions=latent_regions)

    def laplacian(self, recon):
        return self._lap(self.self_ns, recon)

    def latent_consistency(self, z, key):
        return self._lc(self.self_ns, z, key)


class ReferenceStep:
    """``ModelManager._do_iteration`` (model_manager.py:274-326) on the reference's own ``Model`` / losses /
    ``SwapFeatures``: craniofacial.yaml weights by default.  ``net_module``: the module that provides ``Model`` (the
    reference's model.py, or this repo's drop-in for the drop-in tests)."""

    def __init__(self, ref, tabs, device='cpu', bs=4, seed=0, net_module=None, lr=1e-4,
                 w=None, channels=(32, 32, 32, 64), latent=75):
        self.ref, self.tabs, self.dev, self.bs = ref, tabs, torch.device(device), bs
        self.w = dict(kl=1e-4, lc=0.5, lap=0.1, eta1=0.5, eta2=0.5)
        self.w.update(w or {})
        Data = install_shims()
        self.Data = Data
        self.model_mod = net_module or import_reference_module(ref, 'model', '_sdvae_reference_model')
        self.swap_mod = import_reference_module(ref, 'swap_batch_transform', '_sdvae_reference_swap')
        sp = [t.to(self.dev) for t in tabs.spiral_tensors()]
        dn = [t.to(self.dev) for t in tabs.down_tensors()]
        up = [t.to(self.dev) for t in tabs.up_tensors()]
        torch.manual_seed(seed)
        self.net = self.model_mod.Model(3, list(channels), latent, sp, dn, up, False, self.w['kl'] > 0).to(self.dev)
        self.opt = torch.optim.Adam(self.net.parameters(), lr=lr, weight_decay=0.0)     # model_manager.py:69-72
        self.latent_regions = tabs.latent_regions(latent)
        self.losses = ReferenceLosses(ref, tabs.laplacian_tensor().to(self.dev), self.latent_regions, bs,
                                      self.w['eta1'], self.w['eta2'])
        feat = {k: {'feature': idx.tolist()} for k, idx in tabs.regions}
        self.swapper = self.swap_mod.SwapFeatures(types.SimpleNamespace(feat_and_cont=feat))

    def swap(self, x_cpu):
        """The reference's CPU collate transform (swap_batch_transform.py:13-52) on an un-swapped batch [bs, V, 3]."""
        bs = x_cpu.shape[0]
        batch = self.Data(x=x_cpu, y=['a'] * bs, augmented=torch.zeros(bs, 1), gender=['f'] * bs,
                          age=torch.ones(bs, 1))
        out = self.swapper(batch)
        return out.x, out.swapped

    def step(self, x_swapped, key, train=True):
        """model_manager.py:274-326.  Returns the loss dict (python floats)."""
        net, L, w = self.net, self.losses, self.w
        if train:
            self.opt.zero_grad()
            net.train()
        x = x_swapped.to(self.dev)
        recon, z, mu, logvar = net(x)
        l_rec = L.mse(recon, x)
        l_lap = L.laplacian(recon)
        l_kl = L.kl(mu, logvar) if w['kl'] > 0 else torch.zeros((), device=self.dev)
        l_lc = L.latent_consistency(z, key) if w['lc'] > 0 else torch.zeros((), device=self.dev)
        tot = l_rec + w['kl'] * l_kl + w['lc'] * l_lc + w['lap'] * l_lap
        if train:
            tot.backward()
            self.opt.step()
        return {'reconstruction': l_rec.item(), 'kl': l_kl.item(), 'latent_consistency': l_lc.item(),
                'laplacian': l_lap.item(), 'tot': tot.item()}


def time_reference_steps(ref, tabs, device, bs, steps, warmup, seed=0):
    """meshes/s and seconds/step of the reference training step (SwapFeatures collate on the CPU included, as in the
    reference's DataLoader) on ``device``."""
    import time
    st = ReferenceStep(ref, tabs, device=device, bs=bs, seed=seed)
    rng = np.random.RandomState(seed)
    x = torch.from_numpy(rng.randn(bs, tabs.num_vertices[0], 3).astype(np.float32))
    random.seed(seed)
    times = []
    for it in range(warmup + steps):
        if st.dev.type == 'cuda':
            torch.cuda.synchronize(st.dev)
        t0 = time.perf_counter()
        xa, key = st.swap(x)
        st.step(xa, key)
        if st.dev.type == 'cuda':
            torch.cuda.synchronize(st.dev)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = float(np.sum(times))
    return bs * bs * len(times) / total, total / len(times)
