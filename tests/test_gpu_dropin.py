"""Round-2 parity tests (GPU box): measured per-tensor errors of both engines against the fp64 oracle, the
benchmarked configuration, the drop-in used AS the reference's top-level ``model`` module under the reference's own
loss methods, the decoder-only ``generate_for_opt`` path, the L1 reduction."""
import importlib
import os
import random
import sys

import numpy as np
import pytest
import torch

from _util import build_pair, nerr, rand

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W = dict(kl=1e-4, lc=0.5, lap=0.1, eta1=0.5, eta2=0.5)      # craniofacial.yaml:22-27

# Whole-network bars against the oracle evaluated in FLOAT64 (normwise max|a-b| / max|b|), set at <= 2x the values
# measured on B200 (printed by the tests; profiles/r02_parity_tests.log): the fp32-FMA engine meets north_star's 1e-5 on every
# tensor; the tcgen05 engine (error-compensated 3xTF32, accumulators drained every two tiles) is stated separately.
# Measured (B200, this test, craniofacial case A): fp32-FMA engine recon 1.3e-6, z 8.6e-7, losses <= 5.9e-7, gradients max
# 3.6e-6 / median 8.5e-7; tcgen05 engine recon 1.4e-5, z 8.2e-6, losses <= 1.2e-5, gradients max 2.3e-5 / median 1.2e-5
# (per layer it sits at ~2e-6: the tensor core accumulates in fp32 with TRUNCATION, a one-sided error that compounds
# through the ten layers).
BARS = {
    False: dict(recon=3e-6, z=2e-6, loss=2e-6, grad=8e-6),
    True: dict(recon=3e-5, z=2e-5, loss=3e-5, grad=5e-5),
}


def _engine(model, tabs, bs, use_graph, use_tc, **kw):
    from sdvae_b200 import losses
    from sdvae_b200.engine import StepConfig, TrainEngine
    cfg = StepConfig(batch_size=bs, **kw)
    lt = losses.LaplacianTable.build(*tabs.lap, tabs.num_vertices[0], DEV)
    lat = tabs.latent_regions(model.latent_size)
    return TrainEngine(model, lt, [r[1] for r in tabs.regions], [lat[k] for k in tabs.region_keys()],
                       cfg, use_graph=use_graph, use_tc=use_tc)


def _oracle64(cranio, params, xa, bs, region_key, eps):
    """Losses, reconstruction, z and all 24 gradients of the oracle evaluated in float64."""
    from oracle import sdvae_oracle as orc
    sp, dn, up = cranio.spiral_tensors(), cranio.down_tensors(), cranio.up_tensors()
    net = orc.Net(3, [32, 32, 32, 64], 75, sp, [d.double() for d in dn], [u.double() for u in up], False, True)
    p64 = {k: v.double().clone().requires_grad_(True) for k, v in params.items()}
    lap = tuple(torch.from_numpy(a) for a in cranio.lap)
    lap64 = (lap[0], lap[1], lap[2].double())
    recon, z, mu, lv = net.forward(p64, xa.double(), training=True, eps=eps.double())
    region = cranio.latent_regions(75)[region_key]
    tot, parts = orc.total_loss(recon, xa.double(), z, mu, lv, lap64, bs, region, W)
    tot.backward()
    return recon.detach(), z.detach(), parts, {k: v.grad for k, v in p64.items()}


@pytest.mark.parametrize('use_tc', [False, True], ids=['fma', 'tcgen05'])
def test_measured_errors_vs_fp64_oracle(golden, cranio, use_tc):
    """Every tensor of one training step (reconstruction, z, the four losses, all 24 gradients) of the fused
    engine against the oracle in FLOAT64; the measured normwise errors are printed and must stay under BARS."""
    from oracle import sdvae_oracle as orc
    net, params, model = build_pair(cranio, 3, [32, 32, 32, 64], 75, False, True, 77, DEV)
    eng = _engine(model, cranio, 2, False, use_tc, lr=1e-3)
    x2 = torch.from_numpy(golden['A_x_unswapped'])
    keys = cranio.region_keys()
    eps = rand((4, 75), 100)
    xa = orc.swap_features(x2, torch.from_numpy(cranio.regions[3][1]))
    recon64, z64, parts64, grads64 = _oracle64(cranio, params, xa, 2, keys[3], eps)
    eng.set_fixed_eps(eps.to(DEV))
    eng.load_batch(x2.to(DEV))
    got = eng.step(3, sync_losses=True)
    bars = BARS[use_tc]
    e_recon = nerr(eng.recon_template_order(), recon64)
    e_z = nerr(eng.z, z64)
    e_loss = {k: abs(got[k] - float(parts64[k])) / abs(float(parts64[k]))
              for k in ('reconstruction', 'kl', 'latent_consistency', 'laplacian')}
    named = dict(model.named_parameters())
    e_grad = {k: nerr(eng.g(named[k]), g) for k, g in grads64.items()}
    worst = max(e_grad, key=e_grad.get)
    print('\n[%s engine vs fp64 oracle] recon %.2e  z %.2e  losses %s  gradients: max %.2e (%s), median %.2e'
          % ('tcgen05' if use_tc else 'fma', e_recon, e_z, {k: '%.1e' % v for k, v in e_loss.items()},
             e_grad[worst], worst, float(np.median(list(e_grad.values())))))
    assert e_recon < bars['recon'] and e_z < bars['z']
    assert all(v < bars['loss'] for v in e_loss.values()), e_loss
    assert all(v < bars['grad'] for v in e_grad.values()), {k: v for k, v in e_grad.items() if v >= bars['grad']}


def test_benchmarked_configuration_three_steps_vs_oracle(cranio):
    """What bench.py runs -- use_graph=True, use_tc=True, multi-tile CTAs -- at bs = 8 (64 swapped meshes: every CTA
    of the persistent kernels walks several tiles): three consecutive graph-replayed steps (through two Adam
    updates) against the oracle's _do_iteration, losses each step and all 24 gradients on the first."""
    from oracle import sdvae_oracle as orc
    bs = 8
    net, params, model = build_pair(cranio, 3, [32, 32, 32, 64], 75, False, True, 5, DEV)
    eng = _engine(model, cranio, bs, True, True, lr=1e-3)
    lap = tuple(torch.from_numpy(a) for a in cranio.lap)
    trainer = orc.Trainer(net, params, lap, W, lr=1e-3)
    torch.set_num_threads(os.cpu_count() or 1)
    x = rand((bs, cranio.num_vertices[0], 3), 9)
    keys = cranio.region_keys()
    named = dict(model.named_parameters())
    for it, ridx in enumerate((3, 10, 3)):
        eps = rand((bs * bs, 75), 200 + it)
        xa = orc.swap_features(x * (1.0 + 0.1 * it), torch.from_numpy(cranio.regions[ridx][1]))
        want = trainer.step(xa, bs, cranio.latent_regions(75)[keys[ridx]], eps=eps)
        eng.set_fixed_eps(eps.to(DEV))
        eng.load_batch((x * (1.0 + 0.1 * it)).to(DEV))
        got = eng.step(ridx, sync_losses=True)
        # step 0 is a parity statement (same weights, same inputs).  Steps 1 and 2 run on weights that went through Adam:
        # its first update is lr * sign(g) whatever |g| is, so a gradient element whose sign is decided by rounding moves
        # a weight by +lr in one implementation and -lr in the other; the trajectories stay close, not rounding-close
        # (measured here: <= 3.1e-5 on every loss through two updates at lr = 1e-3).
        tol = 3e-5 if it == 0 else 1e-4
        for k in ('reconstruction', 'kl', 'latent_consistency', 'laplacian', 'tot'):
            assert got[k] == pytest.approx(want[k], rel=tol), (it, k, got[k], want[k])
        if it == 0:
            errs = {k: nerr(eng.g(named[k]), v.grad) for k, v in trainer.params.items()}
            assert max(errs.values()) < 5e-5, {k: v for k, v in errs.items() if v >= 5e-5}


# --------------------------------------------------------------------------------------------------------------
def _refarm():
    sys.path.insert(0, ROOT)
    from baseline import refarm
    ref = refarm.find_ref()
    if ref is None:
        pytest.skip('reference not staged (tools/stage_reference.py needs /root/reference)')
    return refarm, ref


def test_dropin_imported_as_toplevel_model_under_the_reference_losses(cranio):
    """The drop-in claim, exercised as a drop-in: ``craniofacialsd-vae_b200/`` first on sys.path, ``import model``
    (top-level, the way model_manager.py:31 imports it), then the REFERENCE's own loss methods (lifted from the staged
    model_manager.py), its SwapFeatures collate and torch autograd drive it on the GPU -- against the same code
    driving the reference's own model.py on the CPU with the same weights (eval mode: z = mu, no RNG)."""
    refarm, ref = _refarm()
    pkg = os.path.join(ROOT, 'craniofacialsd-vae_b200')
    saved = sys.modules.pop('model', None)
    sys.path.insert(0, pkg)
    try:
        dropin = importlib.import_module('model')
        assert os.path.dirname(os.path.abspath(dropin.__file__)) == pkg and dropin.__package__ in ('', None)
        for name in ('SpiralConv', 'Pool', 'SpiralEnblock', 'SpiralDeblock', 'Model', 'MLPClassifier'):
            assert hasattr(dropin, name), name
        ours = refarm.ReferenceStep(ref, cranio, device=DEV, bs=2, seed=3, net_module=dropin)
        theirs = refarm.ReferenceStep(ref, cranio, device='cpu', bs=2, seed=3)
        assert type(theirs.net).__module__ == '_sdvae_reference_model'
        assert list(ours.net.state_dict()) == list(theirs.net.state_dict())          # model_manager.py:693 strict load
        ours.net.load_state_dict({k: v.to(DEV) for k, v in theirs.net.state_dict().items()}, strict=True)
        rng = np.random.RandomState(4)
        x = torch.from_numpy(rng.randn(2, cranio.num_vertices[0], 3).astype(np.float32))
        random.seed(11)
        xa, key = theirs.swap(x)
        out = {}
        for tag, st in (('ours', ours), ('theirs', theirs)):
            st.net.eval()
            xd = xa.to(st.dev)
            recon, z, mu, logvar = st.net(xd)
            L = st.losses
            parts = [L.mse(recon, xd), L.kl(mu, logvar), L.latent_consistency(z, key), L.laplacian(recon),
                     L.l1(recon, xd)]
            tot = parts[0] + 1e-4 * parts[1] + 0.5 * parts[2] + 0.1 * parts[3]
            st.net.zero_grad()
            tot.backward()
            out[tag] = ([float(p) for p in parts], recon.detach().cpu(),
                        {k: p.grad.detach().cpu() for k, p in st.net.named_parameters()})
        for a, b in zip(out['ours'][0], out['theirs'][0]):
            assert a == pytest.approx(b, rel=2e-5)
        assert nerr(out['ours'][1], out['theirs'][1]) < 3e-5        # tcgen05 drop-in vs the reference's fp32 run (measured 1.0e-5)
        errs = {k: nerr(out['ours'][2][k], g) for k, g in out['theirs'][2].items()}
        assert max(errs.values()) < 5e-5, {k: v for k, v in errs.items() if v >= 5e-5}
    finally:
        sys.path.remove(pkg)
        sys.modules.pop('model', None)
        if saved is not None:
            sys.modules['model'] = saved


@pytest.mark.parametrize('freeze', [False, True], ids=['params-require-grad', 'frozen-params'])
def test_generate_for_opt_decoder_gradient_wrt_z(cranio, freeze):
    """ModelManager.generate_for_opt (model_manager.py:253-255) as Tester.fit_mesh uses it (test.py:395-421): decoder in
    TRAIN mode, batch 16, gradient w.r.t. z.  With the parameters frozen (``requires_grad_(False)``) the backward
    functions skip every weight gradient (functional.SpiralConvFn checks needs_input_grad) -- same dz."""
    net, params, model = build_pair(cranio, 3, [32, 32, 32, 64], 75, False, True, 21, DEV)
    model.train()
    if freeze:
        model.requires_grad_(False)
    z = rand((16, 75), 31).to(DEV).requires_grad_(True)
    target = rand((16, cranio.num_vertices[0], 3), 32).to(DEV)
    out = model.decode(z)
    loss = ((out - target) ** 2).mean()
    loss.backward()
    assert z.grad is not None
    if freeze:
        assert all(p.grad is None for p in model.parameters())
    z64 = z.detach().cpu().double().requires_grad_(True)
    sp, dn, up = cranio.spiral_tensors(), cranio.down_tensors(), cranio.up_tensors()
    from oracle import sdvae_oracle as orc
    net64 = orc.Net(3, [32, 32, 32, 64], 75, sp, [d.double() for d in dn], [u.double() for u in up], False, True)
    out64 = net64.decode({k: v.double() for k, v in params.items()}, z64)
    ((out64 - target.cpu().double()) ** 2).mean().backward()
    assert nerr(out, out64) < 1e-5 and nerr(z.grad, z64.grad) < 3e-5


def test_l1_loss_matches_torch(cranio):
    """losses.l1_loss = torch.nn.L1Loss(reduction='mean') (ModelManager._compute_l1_loss, model_manager.py:328-330)."""
    from sdvae_b200 import losses
    a = rand((3, 4260, 3), 41).to(DEV).requires_grad_(True)
    b = rand((3, 4260, 3), 42).to(DEV)
    got = losses.l1_loss(a, b)
    got.backward()
    a2 = a.detach().clone().requires_grad_(True)
    want = torch.nn.L1Loss(reduction='mean')(a2.double(), b.double())
    want.backward()
    assert float(got) == pytest.approx(float(want), rel=1e-6)
    assert torch.equal(a.grad, a2.grad.float())
