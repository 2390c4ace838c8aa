"""CPU oracle for the SD-VAE hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional restatement, in plain CPU PyTorch, of the reference's mesh
encoder/decoder path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` (plus the CPU-baseline
column of ``tools/infer_bench.py`` and the fixture generator ``tools/make_golden.py``)
may import this module, and only as the checker or the timed CPU baseline.  Nothing
under ``craniofacialsd-vae_b200/`` imports it.

Every function cites the reference lines it follows (paths are relative to the
reference checkout).  Pinning: the reference has **no** golden vectors or tests
for this path, so the oracle is pinned against outputs of
the reference's own ``model.py`` / ``model_manager.py`` code executed in the
build container by ``tools/make_golden.py`` (committed) and stored in
``tests/golden/reference_vectors.npz``; ``tests/test_oracle_golden.py`` checks
the oracle against those vectors.  Third-party arithmetic the reference delegates
to (``torch_scatter.scatter_add`` = ``zeros().scatter_add_()``, PyG
``get_laplacian('rw')``, ``torch.optim.Adam``) is restated from the published
semantics of those packages (unpinned in the reference's install_env.sh:15-19).

Works in fp32 (the reference's type) or fp64 (the arbiter for tolerances):
the dtype follows the parameters / inputs passed in.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Params = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------
# layer primitives
# --------------------------------------------------------------------------
def spiral_conv(x: torch.Tensor, indices: torch.Tensor, weight: torch.Tensor,
                bias: Optional[torch.Tensor]) -> torch.Tensor:
    """model.py:27-41.  ``y[b,v,:] = W @ concat_s x[b, indices[v,s], :] + bias``.

    ``x`` is ``[B,V,C]`` or ``[V,C]``; any other rank raises RuntimeError with the
    reference's message."""
    n_nodes, seq = indices.shape
    flat = indices.reshape(-1)
    if x.dim() == 2:
        g = x.index_select(0, flat).reshape(n_nodes, -1)
    elif x.dim() == 3:
        g = x.index_select(1, flat).reshape(x.shape[0], n_nodes, -1)
    else:
        raise RuntimeError('x.dim() is expected to be 2 or 3, but received {}'.format(x.dim()))
    y = g @ weight.t()
    return y if bias is None else y + bias


def pool(x: torch.Tensor, row: torch.Tensor, col: torch.Tensor, val: torch.Tensor,
         n_out: int) -> torch.Tensor:
    """model.py:50-55 with torch_scatter.scatter_add == zeros().scatter_add_().

    ``out[b,r,:] = sum_{e: row[e]==r} val[e] * x[b, col[e], :]``, entries visited in
    storage order, product rounded before the add."""
    prod = x.index_select(1, col) * val.unsqueeze(-1)
    out = torch.zeros(x.shape[0], n_out, x.shape[2], dtype=x.dtype)
    return out.scatter_add_(1, row.view(1, -1, 1).expand_as(prod), prod)


def pool_sparse(x: torch.Tensor, trans: torch.Tensor) -> torch.Tensor:
    """``Pool(x, trans)`` for a torch sparse COO ``trans`` (model.py:50-55)."""
    row, col = trans._indices()
    return pool(x, row, col, trans._values().to(x.dtype), trans.size(0))


def elu(x: torch.Tensor) -> torch.Tensor:
    """F.elu, alpha = 1 (model.py:68, 84): ``x > 0 ? x : expm1(x)``."""
    return torch.where(x > 0, x, torch.expm1(x))


# --------------------------------------------------------------------------
# network
# --------------------------------------------------------------------------
class Net:
    """Functional view of ``Model`` (model.py:88-188) over a state-dict-keyed
    parameter dictionary (keys listed in SURVEY.md section 8b)."""

    def __init__(self, in_channels: int, out_channels: Sequence[int], latent_size: int,
                 spirals: Sequence[torch.Tensor], down: Sequence[torch.Tensor],
                 up: Sequence[torch.Tensor], pre_z_sigmoid=False, is_vae=False):
        self.cin = in_channels
        self.chan = list(out_channels)
        self.latent = latent_size
        self.spirals = [s.long() for s in spirals]
        self.down = list(down)
        self.up = list(up)
        self.pre_z_sigmoid = pre_z_sigmoid
        self.is_vae = is_vae
        self.n_blocks = len(self.chan)
        self.num_vert = int(self.down[-1].size(0))                       # model.py:99
        self.seq = [int(s.shape[1]) for s in self.spirals]

    # parameter shapes in the reference's key order (model.py:104-136)
    def param_shapes(self) -> "Dict[str, Tuple[int, ...]]":
        shapes: Dict[str, Tuple[int, ...]] = {}
        L = self.n_blocks
        for i in range(L):
            ci = self.cin if i == 0 else self.chan[i - 1]
            shapes['en_layers.%d.conv.layer.weight' % i] = (self.chan[i], ci * self.seq[i])
            shapes['en_layers.%d.conv.layer.bias' % i] = (self.chan[i],)
        flat = self.num_vert * self.chan[-1]
        n_lin = 2 if self.is_vae else 1
        for j in range(n_lin):
            shapes['en_layers.%d.weight' % (L + j)] = (self.latent, flat)
            shapes['en_layers.%d.bias' % (L + j)] = (self.latent,)
        shapes['de_layers.0.weight'] = (flat, self.latent)
        shapes['de_layers.0.bias'] = (flat,)
        for i in range(L):
            ci = self.chan[-1] if i == 0 else self.chan[-i]
            co = self.chan[-i - 1]
            lvl = L - 1 - i
            shapes['de_layers.%d.conv.layer.weight' % (i + 1)] = (co, ci * self.seq[lvl])
            shapes['de_layers.%d.conv.layer.bias' % (i + 1)] = (co,)
        shapes['de_layers.%d.layer.weight' % (L + 1)] = (self.cin, self.chan[0] * self.seq[0])
        shapes['de_layers.%d.layer.bias' % (L + 1)] = (self.cin,)
        return shapes

    def encode(self, p: Params, x: torch.Tensor):
        """model.py:146-160: enblock = Pool(elu(conv(x)), down) (model.py:67-70);
        mu from the LAST linear, logvar from the one before it."""
        L = self.n_blocks
        for i in range(L):
            h = spiral_conv(x, self.spirals[i], p['en_layers.%d.conv.layer.weight' % i],
                            p['en_layers.%d.conv.layer.bias' % i])
            x = pool_sparse(elu(h), self.down[i])
        flat = x.reshape(-1, self.num_vert * self.chan[-1])
        last = L + (1 if self.is_vae else 0)
        mu = flat @ p['en_layers.%d.weight' % last].t() + p['en_layers.%d.bias' % last]
        if self.is_vae:
            logvar = flat @ p['en_layers.%d.weight' % L].t() + p['en_layers.%d.bias' % L]
        else:
            mu = torch.sigmoid(mu) if self.pre_z_sigmoid else mu
            logvar = None
        return mu, logvar

    def decode(self, p: Params, z: torch.Tensor) -> torch.Tensor:
        """model.py:162-173: deblock = elu(conv(Pool(x, up))) (model.py:82-85); the
        last layer is a bare SpiralConv."""
        L = self.n_blocks
        x = (z @ p['de_layers.0.weight'].t() + p['de_layers.0.bias'])
        x = x.reshape(-1, self.num_vert, self.chan[-1])
        for i in range(L):
            lvl = L - 1 - i
            x = pool_sparse(x, self.up[lvl])
            x = elu(spiral_conv(x, self.spirals[lvl],
                                p['de_layers.%d.conv.layer.weight' % (i + 1)],
                                p['de_layers.%d.conv.layer.bias' % (i + 1)]))
        return spiral_conv(x, self.spirals[0], p['de_layers.%d.layer.weight' % (L + 1)],
                           p['de_layers.%d.layer.bias' % (L + 1)])

    def forward(self, p: Params, x: torch.Tensor, training=False,
                eps: Optional[torch.Tensor] = None):
        """model.py:175-188.  ``eps`` replaces ``randn_like`` so that tests can inject
        the same noise into both implementations."""
        mu, logvar = self.encode(p, x)
        if self.is_vae and training:
            if eps is None:
                eps = torch.randn_like(mu)
            z = mu + eps * torch.exp(0.5 * logvar)
        else:
            z = mu
        return self.decode(p, z), z, mu, logvar


def xavier_params(shapes: "Dict[str, Tuple[int, ...]]", seed: int, dtype=torch.float32,
                  bias_scale: float = 0.0) -> Params:
    """Deterministic xavier-uniform weights from a NumPy stream (the reference
    re-initialises every weight xavier-uniform and every bias to 0, model.py:139-144;
    ``bias_scale > 0`` draws non-zero biases so that tests exercise the bias path)."""
    import numpy as np
    rng = np.random.RandomState(seed)
    out: Params = {}
    for k in shapes:
        shp = shapes[k]
        if len(shp) == 2:
            a = math.sqrt(6.0 / (shp[0] + shp[1]))
            out[k] = torch.from_numpy(rng.uniform(-a, a, shp)).to(dtype)
        else:
            out[k] = torch.from_numpy(rng.uniform(-bias_scale, bias_scale, shp)).to(dtype) \
                if bias_scale > 0 else torch.zeros(shp, dtype=dtype)
    return out


# --------------------------------------------------------------------------
# losses (model_manager.py:274-393)
# --------------------------------------------------------------------------
def mse_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """model_manager.py:332-334: mean over every element."""
    d = pred - target
    return (d * d).sum() / d.numel()


def kl_loss(mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
    """model_manager.py:351-354."""
    per_mesh = -0.5 * (1.0 + logvar - mu * mu - torch.exp(logvar)).sum(dim=1)
    return per_mesh.sum() / mu.shape[0]


def laplacian_loss(pred: torch.Tensor, lap_row: torch.Tensor, lap_col: torch.Tensor,
                   lap_val: torch.Tensor) -> torch.Tensor:
    """model_manager.py:343-349 + utils.py:153-165: ``q_b = L @ pred_b``;
    ``sum_b sum_v ||q_b[v]||_2 / V / B``."""
    b, v, _ = pred.shape
    q = pool(pred, lap_row, lap_col, lap_val.to(pred.dtype), v)
    return q.norm(dim=-1).sum() / v / b


def latent_consistency_loss(z: torch.Tensor, bs: int, r0: int, r1: int,
                            eta1: float, eta2: float) -> torch.Tensor:
    """model_manager.py:360-393, written pair by pair.

    ``z`` is ``[bs*bs, D]`` laid out on the swap grid: element ``i*bs+j`` is base
    mesh ``i`` carrying the swapped region of donor ``j``
    (swap_batch_transform.py:27-38).  ``zf`` = the swapped region's latent slice,
    ``ze`` = all other latents.  For every pair ``a < b`` and every ``t``::

        lg = |zf[b,t]-zf[a,t]|^2   dg = |zf[t,b]-zf[t,a]|^2
        dr = |ze[b,t]-ze[a,t]|^2   lr = |ze[t,b]-ze[t,a]|^2
        loss += max(0, lr - dr + eta2) + max(0, lg - dg + eta1)

    normalised by ``bs^3 - bs^2``."""
    d = z.shape[1]
    zf = z[:, r0:r1].reshape(bs, bs, r1 - r0)
    ze = torch.cat([z[:, :r0], z[:, r1:]], dim=1).reshape(bs, bs, d - (r1 - r0))
    total = z.new_zeros(())
    for a in range(bs):
        for b in range(a + 1, bs):
            lg = ((zf[b] - zf[a]) ** 2).sum(-1)              # [t]
            dg = ((zf[:, b] - zf[:, a]) ** 2).sum(-1)
            dr = ((ze[b] - ze[a]) ** 2).sum(-1)
            lr = ((ze[:, b] - ze[:, a]) ** 2).sum(-1)
            total = total + torch.clamp(lr - dr + eta2, min=0).sum() \
                + torch.clamp(lg - dg + eta1, min=0).sum()
    return total / float(bs ** 3 - bs ** 2)


def total_loss(recon, x, z, mu, logvar, lap, bs, region, weights):
    """model_manager.py:281-312.  ``weights`` = dict(kl, lc, lap, eta1, eta2);
    ``region`` = (r0, r1) latent slice of the swapped feature or None.
    Returns (total, dict of the individual terms)."""
    terms = {'reconstruction': mse_loss(recon, x),
             'laplacian': laplacian_loss(recon, *lap)}
    zero = recon.new_zeros(())
    terms['kl'] = kl_loss(mu, logvar) if weights['kl'] > 0 else zero
    terms['latent_consistency'] = latent_consistency_loss(
        z, bs, region[0], region[1], weights['eta1'], weights['eta2']) \
        if region is not None else zero
    tot = terms['reconstruction'] + weights['kl'] * terms['kl'] + \
        weights['lc'] * terms['latent_consistency'] + weights['lap'] * terms['laplacian']
    terms['tot'] = tot
    return tot, terms


# --------------------------------------------------------------------------
# feature swap (swap_batch_transform.py:13-52)
# --------------------------------------------------------------------------
def swap_features(x: torch.Tensor, feature_idx: torch.Tensor) -> torch.Tensor:
    """``out[i*bs+j] = x[i]`` with the vertices of ``feature_idx`` taken from ``x[j]``;
    the diagonal is the untouched original (swap_batch_transform.py:27-38, 44-52)."""
    bs = x.shape[0]
    out = x.unsqueeze(1).repeat(1, bs, 1, 1)                 # [i, j, V, C] = x[i]
    out[:, :, feature_idx, :] = x[:, feature_idx, :].unsqueeze(0).expand(bs, -1, -1, -1)
    return out.reshape(bs * bs, *x.shape[1:])


# --------------------------------------------------------------------------
# one training iteration (model_manager.py:274-326)
# --------------------------------------------------------------------------
class Trainer:
    """Oracle of ``ModelManager._do_iteration`` for the hot path: forward, the four
    loss terms, backward, ``torch.optim.Adam`` (model_manager.py:69-72, 314-316)."""

    def __init__(self, net: Net, params: Params, lap, weights, lr=1e-4, weight_decay=0.0):
        self.net = net
        self.params = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        self.lap = lap
        self.weights = weights
        self.opt = torch.optim.Adam(list(self.params.values()), lr=lr,
                                    weight_decay=weight_decay)

    def step(self, x, bs, region, eps=None, train=True):
        self.opt.zero_grad()
        recon, z, mu, logvar = self.net.forward(self.params, x, training=train, eps=eps)
        tot, terms = total_loss(recon, x, z, mu, logvar, self.lap, bs, region, self.weights)
        if train:
            tot.backward()
            self.opt.step()
        return {k: float(v.detach()) for k, v in terms.items()}
