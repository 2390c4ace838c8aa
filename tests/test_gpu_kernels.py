"""Parity of every CUDA kernel against the CPU oracle, through the C ABI (GPU box only)."""
import numpy as np
import pytest
import torch

from _util import TOL, nerr, rand

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def orc():
    from oracle import sdvae_oracle
    return sdvae_oracle


# the nine SpiralConv instances of craniofacial.yaml + odd shapes (generic path)
LAYERS = [  # (level, Cin, Cout)
    (0, 3, 32), (1, 32, 32), (2, 32, 32), (3, 32, 64), (3, 64, 64), (2, 64, 32), (1, 32, 32),
    (0, 32, 32), (0, 32, 3), (3, 5, 7), (2, 16, 48), (3, 32, 96), (3, 64, 2)]


@pytest.mark.parametrize('lvl,cin,cout', LAYERS)
@pytest.mark.parametrize('act', [0, 1])
def test_spiralconv_forward_backward(cranio, orc, lvl, cin, cout, act):
    from sdvae_b200 import cabi, functional as F_
    from sdvae_b200.tables import spiral_table
    B = 3 if lvl else 2
    idx = cranio.spiral_tensors()[lvl]
    V, S = idx.shape
    x = rand((B, V, cin), 10 + lvl)
    w = rand((cout, S * cin), 20 + cin, (2.0 / (S * cin)) ** 0.5)
    b = rand((cout,), 30 + cout, 0.1)
    gy = rand((B, V, cout), 40 + lvl)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    yr = orc.spiral_conv(xr, idx, wr, br)
    yr = orc.elu(yr) if act else yr
    yr.backward(gy)
    xg, wg, bg = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    tab = spiral_table(idx.to(DEV))
    y = F_.spiral_conv(xg, wg, bg, tab, act)
    y.backward(gy.to(DEV))
    assert nerr(y, yr) < TOL
    assert nerr(xg.grad, xr.grad) < TOL
    assert nerr(wg.grad, wr.grad) < 2 * TOL
    assert nerr(bg.grad, br.grad) < 2 * TOL


def test_spiralconv_restricted_rows_equal_conv_then_select(cranio, orc):
    """Fused encoder block: conv on kept vertices == Pool(elu(conv(x)), down)."""
    from sdvae_b200.model import SpiralEnblock
    idx = cranio.spiral_tensors()[1]
    down = cranio.down_tensors()[1]
    blk = SpiralEnblock(32, 32, idx.to(DEV)).to(DEV)
    w = rand((32, 288), 1, 0.08); b = rand((32,), 2, 0.1)
    blk.conv.load_state_dict({'layer.weight': w, 'layer.bias': b})
    x = rand((4, 4260, 32), 3)
    gy = rand((4, 1065, 32), 4)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    yr = orc.pool_sparse(orc.elu(orc.spiral_conv(xr, idx, wr, br)), down)
    yr.backward(gy)
    xg = x.to(DEV).requires_grad_(True)
    y = blk(xg, down.to(DEV))
    y.backward(gy.to(DEV))
    assert nerr(y, yr) < TOL
    assert nerr(xg.grad, xr.grad) < TOL
    assert nerr(blk.conv.layer.weight.grad, wr.grad) < 2 * TOL
    assert nerr(blk.conv.layer.bias.grad, br.grad) < 2 * TOL


def test_spiralconv_2d_input_and_errors(cranio, golden):
    from sdvae_b200 import fixtures as fx
    from sdvae_b200.model import SpiralConv
    stab = fx.synthetic_tables(203, 2, seq_length=7, n_regions=3, seed=5)
    idx = stab.spiral_tensors()[0].to(DEV)
    conv = SpiralConv(3, 5, idx).to(DEV)
    conv.load_state_dict({'layer.weight': torch.from_numpy(golden['B_conv2d_w']),
                          'layer.bias': torch.from_numpy(golden['B_conv2d_b'])})
    x = torch.from_numpy(golden['B_x'])
    y = conv(x[0].to(DEV))
    assert y.shape == (203, 5) and nerr(y, golden['B_conv2d_out']) < TOL
    with pytest.raises(RuntimeError) as e:
        conv(x.unsqueeze(0).to(DEV))
    assert str(e.value) == str(golden['B_conv_err'])
    with pytest.raises(TypeError):
        conv(x.double().to(DEV))


@pytest.mark.parametrize('lvl,C,B', [(0, 32, 3), (0, 32, 6), (1, 32, 5), (2, 64, 3), (3, 64, 9), (3, 5, 3)])
def test_pool_up_down_bit_exact_forward(cranio, orc, lvl, C, B):
    """Pool keeps the reference's storage-order, mul-then-add arithmetic -> identical bits.
    B >= 4 runs the several-meshes-per-thread kernels (with a ragged last group), B = 3 the plain ones."""
    from sdvae_b200.model import Pool
    up, down = cranio.up_tensors()[lvl], cranio.down_tensors()[lvl]
    Vf, Vc = up.shape
    xc = rand((B, Vc, C), 5)
    xf = rand((B, Vf, C), 6)
    for trans, x in ((up, xc), (down, xf)):
        xr = x.clone().requires_grad_(True)
        yr = orc.pool_sparse(xr, trans)
        g = rand(tuple(yr.shape), 7)
        yr.backward(g)
        xg = x.to(DEV).requires_grad_(True)
        y = Pool(xg, trans.to(DEV))
        y.backward(g.to(DEV))
        assert torch.equal(y.detach().cpu(), yr.detach())
        assert nerr(xg.grad, xr.grad) < 1e-6
    # down(up(x)) == x exactly: kept vertices are one-hot rows of the up-transform
    rt = Pool(Pool(xc.to(DEV), up.to(DEV)), down.to(DEV))
    assert torch.equal(rt.cpu(), xc)


@pytest.mark.parametrize('lvl,C,B', [(0, 32, 19), (1, 32, 37), (2, 64, 21), (3, 64, 5), (0, 32, 1)])
def test_pool_staged_forward_equals_ell_gather_bitwise(cranio, lvl, C, B):
    """The shared-memory-staged forward keeps the ELL kernel's arithmetic and order: identical bits
    (several meshes per CTA, ragged last mesh group, last tile shorter than T)."""
    from sdvae_b200 import cabi
    from sdvae_b200.tables import pool_table
    up = cranio.up_tensors()[lvl].to(DEV)
    pt = pool_table(up)
    plan = pt.stage_plan()
    assert plan is not None and cabi.pool_stage_supported(C, pt.width, plan.ucap)
    Vf, Vc = up.shape
    x = rand((B, Vc, C), 11).to(DEV)
    a = torch.empty(B, Vf, C, device=DEV)
    b = torch.full_like(a, float('nan'))
    cabi.pool_ell_fwd(x, pt.ell_col, pt.ell_val, a, B, Vc, Vf, pt.width, C)
    cabi.pool_ell_fwd_staged(x, plan, b, B, Vc, Vf, pt.width, C)
    assert torch.equal(a, b)
    assert plan.T == cabi.load().sdvae_pool_stage_tile()
    # selection matrices (down-transforms) have nothing to stage: no plan, ELL path
    assert pool_table(cranio.down_tensors()[lvl].to(DEV)).stage_plan() is None


@pytest.mark.parametrize('lvl,C,B', [(0, 32, 19), (1, 32, 8), (2, 64, 21), (3, 64, 4)])
@pytest.mark.parametrize('gated', [False, True])
def test_csr_rowsum_row_per_warp_equals_row_per_thread_bitwise(cranio, lvl, C, B, gated):
    """Pool backward: the row-per-warp kernel (large B) against the one-mesh-per-thread kernel (B = 1 calls)."""
    from sdvae_b200 import cabi
    from sdvae_b200.tables import pool_table
    up = cranio.up_tensors()[lvl].to(DEV)
    pt = pool_table(up)
    Vf, Vc = up.shape
    dy = rand((B, Vf, C), 12).to(DEV)
    gate = rand((B, Vc, C), 13).to(DEV) if gated else None
    a = torch.full((B, Vc, C), float('nan'), device=DEV)
    b = torch.empty_like(a)
    cabi.csr_rowsum(dy, pt.t_ptr, pt.t_row, pt.t_val, gate, a, B, Vf, Vc, C)
    for m in range(B):
        cabi.csr_rowsum(dy[m:m + 1], pt.t_ptr, pt.t_row, pt.t_val, None if gate is None else gate[m:m + 1],
                        b[m:m + 1], 1, Vf, Vc, C)
    assert torch.equal(a, b)


def test_pool_empty_rows_and_ragged_width():
    from sdvae_b200.model import Pool
    ind = torch.tensor([[2, 0, 2, 2], [1, 3, 0, 2]])
    val = torch.tensor([1., 2., 3., 4.])
    trans = torch.sparse_coo_tensor(ind, val, (3, 4))
    x = rand((2, 4, 6), 8)
    y = Pool(x.to(DEV), trans.to(DEV)).cpu()
    ref = torch.zeros(2, 3, 6)
    ref[:, 2] = (0 + 1. * x[:, 1]) + 3. * x[:, 0] + 4. * x[:, 2]
    ref[:, 0] = 2. * x[:, 3]
    assert torch.allclose(y, ref, atol=1e-6) and float(y[:, 1].abs().max()) == 0.0


def test_swap_bit_exact(cranio, orc):
    from sdvae_b200 import cabi
    bs, V = 4, cranio.num_vertices[0]
    x = rand((bs, V, 3), 9)
    for k in (0, 10, 13):
        feat = torch.from_numpy(cranio.regions[k][1])
        ref = orc.swap_features(x, feat)
        mask = torch.zeros(V, dtype=torch.uint8)
        mask[feat] = 1
        out = torch.empty(bs * bs, V, 3, device=DEV)
        cabi.swap(x.to(DEV), mask.to(DEV), out, bs, 0, bs, V, 3)
        assert torch.equal(out.cpu(), ref)
        part = torch.empty(2 * bs, V, 3, device=DEV)            # grid rows 1..2 only
        cabi.swap(x.to(DEV), mask.to(DEV), part, bs, 1, 3, V, 3)
        assert torch.equal(part.cpu(), ref[bs:3 * bs])


@pytest.mark.parametrize('tag', ['LC_2_10', 'LC_3_12', 'LC_4_75', 'LC_5_33'])
def test_latent_consistency_vs_reference_golden(golden, tag):
    from sdvae_b200 import losses
    z = torch.from_numpy(golden[tag + '_z']).to(DEV).requires_grad_(True)
    bs, r0, r1 = [int(t) for t in golden[tag + '_cfg']]
    loss = losses.latent_consistency(z, bs, [r0, r1], 0.5, 0.25)
    loss.backward()
    assert float(loss) == pytest.approx(float(golden[tag + '_loss']), rel=1e-5)
    assert nerr(z.grad, golden[tag + '_grad']) < TOL


def test_latent_consistency_bs32(orc):
    from sdvae_b200 import losses
    z = rand((1024, 75), 11, 0.7)
    zr = z.clone().double().requires_grad_(True)
    lr = orc.latent_consistency_loss(zr, 32, 50, 55, 0.5, 0.5)
    lr.backward()
    zg = z.to(DEV).requires_grad_(True)
    l = losses.latent_consistency(zg, 32, [50, 55], 0.5, 0.5)
    l.backward()
    assert float(l) == pytest.approx(float(lr), rel=1e-5)
    assert nerr(zg.grad, zr.grad) < TOL


def test_mse_kl_laplacian(cranio, orc):
    from sdvae_b200 import losses
    V = cranio.num_vertices[0]
    lap = tuple(torch.from_numpy(a) for a in cranio.lap)
    lt = losses.LaplacianTable.build(*cranio.lap, V, DEV)
    lt2 = losses.LaplacianTable.from_sparse(cranio.laplacian_tensor(DEV))
    assert torch.equal(lt.ell_col, lt2.ell_col) and torch.equal(lt.t_val, lt2.t_val)
    p, t = rand((3, V, 3), 12), rand((3, V, 3), 13)
    pr = p.clone().requires_grad_(True)
    ref = orc.mse_loss(pr, t) * 0.7 + orc.laplacian_loss(pr, *lap) * 0.3
    ref.backward()
    pg = p.to(DEV).requires_grad_(True)
    mse, lp = losses.mse_and_laplacian(pg, t.to(DEV), lt)
    (mse * 0.7 + lp * 0.3).backward()
    assert float(mse) == pytest.approx(float(orc.mse_loss(p, t)), rel=1e-5)
    assert float(lp) == pytest.approx(float(orc.laplacian_loss(p, *lap)), rel=1e-5)
    assert nerr(pg.grad, pr.grad) < TOL
    assert float(losses.mse_loss(pg, t.to(DEV))) == pytest.approx(float(mse), rel=1e-6)
    assert float(losses.laplacian_regularizer(pg, lt)) == pytest.approx(float(lp), rel=1e-6)
    mu, lv = rand((16, 75), 14), rand((16, 75), 15, 0.5)
    mr, lr_ = mu.clone().requires_grad_(True), lv.clone().requires_grad_(True)
    kr = orc.kl_loss(mr, lr_)
    kr.backward()
    mg, lg = mu.to(DEV).requires_grad_(True), lv.to(DEV).requires_grad_(True)
    k = losses.kl_divergence(mg, lg)
    k.backward()
    assert float(k) == pytest.approx(float(kr), rel=1e-5)
    assert nerr(mg.grad, mr.grad) < TOL and nerr(lg.grad, lr_.grad) < TOL


def test_adam_matches_torch():
    from sdvae_b200 import cabi
    n = 10007
    p0, g = rand((n,), 16), rand((n,), 17, 0.01)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3, weight_decay=0.01)
    p = p0.to(DEV); m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    for it in range(5):
        gi = g * (it + 1)
        pr.grad = gi.clone()
        opt.step()
        cabi.adam_tick(step)
        cabi.adam_step(p, gi.to(DEV), m, v, step, 0, 1e-3, 0.9, 0.999, 1e-8, 0.01, 1.0)
    assert int(step) == 5
    assert nerr(p, pr) < 1e-6


def test_backward_is_bitwise_reproducible(cranio):
    from sdvae_b200 import functional as F_
    from sdvae_b200.tables import spiral_table
    idx = cranio.spiral_tensors()[1].to(DEV)
    tab = spiral_table(idx)
    x, w, b = rand((8, 4260, 32), 18).to(DEV), rand((32, 288), 19, 0.08).to(DEV), rand((32,), 20).to(DEV)
    gy = rand((8, 4260, 32), 21).to(DEV)
    outs = []
    for _ in range(2):
        xg, wg, bg = (t.clone().requires_grad_(True) for t in (x, w, b))
        F_.spiral_conv(xg, wg, bg, tab, 1).backward(gy)
        outs.append((xg.grad.clone(), wg.grad.clone(), bg.grad.clone()))
    for a, c in zip(*outs):
        assert torch.equal(a, c)


def test_conv_linearity_at_large_batch(cranio):
    """Size-independent property at a batch the oracle cannot check quickly."""
    from sdvae_b200 import functional as F_
    from sdvae_b200.tables import spiral_table
    idx = cranio.spiral_tensors()[0].to(DEV)
    tab = spiral_table(idx)
    B = 64
    g = torch.Generator(device=DEV).manual_seed(0)
    x1 = torch.randn(B, 17039, 32, device=DEV, generator=g)
    x2 = torch.randn(B, 17039, 32, device=DEV, generator=g)
    w = torch.randn(32, 288, device=DEV, generator=g) * 0.08
    zero_b = torch.zeros(32, device=DEV)
    lhs = F_.spiral_conv(2.5 * x1 + x2, w, zero_b, tab, 0)
    rhs = 2.5 * F_.spiral_conv(x1, w, zero_b, tab, 0) + F_.spiral_conv(x2, w, zero_b, tab, 0)
    assert nerr(lhs, rhs) < TOL
    # batch independence: mesh 17 alone gives the same bits as inside the batch
    solo = F_.spiral_conv(x1[17:18].contiguous(), w, zero_b, tab, 0)
    full = F_.spiral_conv(x1, w, zero_b, tab, 0)
    assert torch.equal(solo[0], full[17])
