"""Pin the CPU oracle (oracle/sdvae_oracle.py) against vectors produced by the
reference's own code (tools/make_golden.py -> tests/golden/reference_vectors.npz)."""
import numpy as np
import pytest
import torch

from oracle import sdvae_oracle as orc
from sdvae_b200 import fixtures as fx

TOL = 1e-5          # normwise max|a-b| / max|b|, SURVEY.md 8d "Parity metric"


def nerr(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def samp(t, stride=97):
    return t.detach().reshape(-1)[::stride]


@pytest.fixture(scope="module")
def case_a(golden, cranio):
    torch.set_num_threads(1)
    spirals, down, up = cranio.spiral_tensors(), cranio.down_tensors(), cranio.up_tensors()
    net = orc.Net(3, [32, 32, 32, 64], 75, spirals, down, up, False, True)
    params = orc.xavier_params(net.param_shapes(), seed=1234, bias_scale=0.05)
    x2 = torch.from_numpy(golden['A_x_unswapped'])
    key = str(golden['A_swapped_key'])
    feat = dict(cranio.regions)[key]
    xa = orc.swap_features(x2, torch.from_numpy(feat))
    return net, params, xa, key


def test_swap_matches_reference(golden, case_a):
    _, _, xa, _ = case_a
    assert torch.equal(samp(xa), torch.from_numpy(golden['A_x_swapped_sample']))
    assert float(xa.double().sum()) == pytest.approx(float(golden['A_x_swapped_sum'][0]), rel=1e-12)


def test_eval_forward_matches_reference(golden, case_a):
    net, params, xa, _ = case_a
    with torch.no_grad():
        rec, z, mu, lv = net.forward(params, xa, training=False)
    assert nerr(rec[0], golden['A_eval_recon0']) < TOL
    assert nerr(samp(rec), golden['A_eval_recon_sample']) < TOL
    assert nerr(mu, golden['A_eval_mu']) < TOL
    assert nerr(lv, golden['A_eval_logvar']) < TOL
    assert torch.equal(z, mu)


def test_train_losses_and_grads_match_reference(golden, case_a, cranio):
    net, params, xa, key = case_a
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    eps = torch.from_numpy(golden['A_eps'])
    rec, z, mu, lv = net.forward(p, xa, training=True, eps=eps)
    assert nerr(z, golden['A_train_z']) < TOL
    lap = tuple(torch.from_numpy(a) for a in cranio.lap)
    region = cranio.latent_regions(75)[key]
    assert list(region) == list(golden['A_region'])
    w = dict(kl=1e-4, lc=0.5, lap=0.1, eta1=0.5, eta2=0.5)
    tot, terms = orc.total_loss(rec, xa, z, mu, lv, lap, 2, region, w)
    ref = golden['A_losses']
    got = [terms['reconstruction'], terms['kl'], terms['latent_consistency'], terms['laplacian'], tot]
    for g, r in zip(got, ref):
        assert float(g) == pytest.approx(float(r), rel=2e-5)
    tot.backward()
    for k, v in p.items():
        g = v.grad
        if 'A_grad/' + k in golden:
            assert nerr(g, golden['A_grad/' + k]) < 5 * TOL, k
        else:
            assert nerr(samp(g), golden['A_grad_sample/' + k]) < 5 * TOL, k
        s = golden['A_grad_sum/' + k]
        assert float(g.double().abs().sum()) == pytest.approx(float(s[1]), rel=1e-4), k


def test_small_ae_case_matches_reference(golden):
    torch.set_num_threads(1)
    stab = fx.synthetic_tables(203, 2, seq_length=7, n_regions=3, seed=5)
    sp, dn, up = stab.spiral_tensors(), stab.down_tensors(), stab.up_tensors()
    net = orc.Net(3, [8, 16], 6, sp, dn, up, True, False)
    p = {k: v.clone().requires_grad_(True)
         for k, v in orc.xavier_params(net.param_shapes(), seed=99, bias_scale=0.1).items()}
    xb = torch.from_numpy(golden['B_x'])
    rec, z, mu, lv = net.forward(p, xb, training=True)
    assert lv is None and nerr(rec, golden['B_recon']) < TOL and nerr(z, golden['B_z']) < TOL
    key = str(golden['B_region_key'])
    region = stab.latent_regions(6)[key]
    lap = tuple(torch.from_numpy(a) for a in stab.lap)
    w = dict(kl=0.0, lc=1.0, lap=1.0, eta1=0.3, eta2=0.7)
    tot, terms = orc.total_loss(rec, xb, z, mu, lv, lap, 3, region, w)
    ref = golden['B_losses']
    assert float(terms['reconstruction']) == pytest.approx(float(ref[0]), rel=2e-5)
    assert float(terms['latent_consistency']) == pytest.approx(float(ref[1]), rel=2e-5)
    assert float(terms['laplacian']) == pytest.approx(float(ref[2]), rel=2e-5)
    tot.backward()
    for k, v in p.items():
        assert nerr(v.grad, golden['B_grad/' + k]) < 5 * TOL, k


def test_conv_2d_input_and_error(golden):
    stab = fx.synthetic_tables(203, 2, seq_length=7, n_regions=3, seed=5)
    idx = stab.spiral_tensors()[0]
    x = torch.from_numpy(golden['B_x'])
    y = orc.spiral_conv(x[0], idx, torch.from_numpy(golden['B_conv2d_w']),
                        torch.from_numpy(golden['B_conv2d_b']))
    assert nerr(y, golden['B_conv2d_out']) < TOL
    with pytest.raises(RuntimeError) as e:
        orc.spiral_conv(x.unsqueeze(0), idx, torch.from_numpy(golden['B_conv2d_w']), None)
    assert str(e.value) == str(golden['B_conv_err'])


def test_pool_matches_reference(golden, cranio):
    rng = np.random.RandomState(3)
    rng.randn(9, 203, 3)                      # consume the stream exactly as make_golden does
    rng.uniform(-0.3, 0.3, (5, 21)); rng.uniform(-0.3, 0.3, (5,))
    xp = torch.from_numpy(rng.randn(2, 4260, 32).astype(np.float32))
    xq = torch.from_numpy(rng.randn(2, 17039, 4).astype(np.float32))
    assert torch.equal(samp(xp, 1009), torch.from_numpy(golden['P_xp_sample']))
    up0, dn0 = cranio.up_tensors()[0], cranio.down_tensors()[0]
    # Pool is add-order sensitive only through fp32 rounding; CPU scatter is sequential
    assert nerr(samp(orc.pool_sparse(xp, up0), 101), golden['P_up0_sample']) < 1e-6
    assert torch.equal(samp(orc.pool_sparse(xq, dn0), 11), torch.from_numpy(golden['P_down0_sample']))


@pytest.mark.parametrize("tag", ["LC_2_10", "LC_3_12", "LC_4_75", "LC_5_33"])
def test_latent_consistency_matches_reference(golden, tag):
    z = torch.from_numpy(golden[tag + '_z']).clone().requires_grad_(True)
    bs, r0, r1 = [int(t) for t in golden[tag + '_cfg']]
    loss = orc.latent_consistency_loss(z, bs, r0, r1, 0.5, 0.25)
    assert float(loss) == pytest.approx(float(golden[tag + '_loss']), rel=2e-6)
    loss.backward()
    assert nerr(z.grad, golden[tag + '_grad']) < TOL


def test_renumbered_tables_give_the_same_network(cranio):
    """``MeshTables.renumbered`` (patch-wise vertex order, groundwork for tile-local staging): the oracle on the
    renumbered tables maps the permuted input to the permuted reconstruction with identical latent codes, and
    the four losses agree."""
    from oracle import sdvae_oracle as orc
    new, orders = cranio.renumbered(128)
    o0 = torch.from_numpy(orders[0])
    assert new.num_vertices == cranio.num_vertices
    assert np.array_equal(orders[-1], np.arange(cranio.num_vertices[-1]))
    for k, (name, idx) in enumerate(cranio.regions):               # same feature vertices, renamed
        assert np.array_equal(np.sort(orders[0][new.regions[k][1]]), np.sort(idx))
    net_a = orc.Net(3, [32, 32, 32, 64], 75, cranio.spiral_tensors(), cranio.down_tensors(), cranio.up_tensors(),
                    False, True)
    net_b = orc.Net(3, [32, 32, 32, 64], 75, new.spiral_tensors(), new.down_tensors(), new.up_tensors(),
                    False, True)
    params = orc.xavier_params(net_a.param_shapes(), seed=3, bias_scale=0.05)
    x = torch.randn(2, cranio.num_vertices[0], 3, generator=torch.Generator().manual_seed(5))
    eps = torch.randn(2, 75, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        ra, za, mua, lva = net_a.forward(params, x, eps=eps, training=True)
        rb, zb, mub, lvb = net_b.forward(params, x[:, o0], eps=eps, training=True)
    assert torch.equal(mua, mub) and torch.equal(lva, lvb) and torch.equal(za, zb)
    assert torch.equal(ra[:, o0], rb)
    lap_a = tuple(torch.from_numpy(a) for a in cranio.lap)
    lap_b = tuple(torch.from_numpy(a) for a in new.lap)
    la = orc.laplacian_loss(ra, *lap_a)
    lb = orc.laplacian_loss(rb, *lap_b)
    assert abs(float(la) - float(lb)) <= 1e-6 * abs(float(la))
