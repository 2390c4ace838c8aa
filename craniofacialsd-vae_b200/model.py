"""Drop-in replacement for the reference's ``model.py`` on B200.

Exports exactly what ``model_manager.py:31`` imports -- ``Model`` and
``MLPClassifier`` -- plus the building blocks ``SpiralConv``, ``Pool``,
``SpiralEnblock`` and ``SpiralDeblock`` with the reference's constructor
signatures, attribute names and ``state_dict`` keys, so
``train.py`` / ``test.py`` / ``model_manager.py`` run unchanged when this file is
the ``model`` module on ``sys.path``.

Differences that are invisible to callers:

* every SpiralConv / ELU / Pool runs as a hand-written sm_100a kernel through the
  C ABI (``include/sdvae_b200.h``); the ``[B, V*S, C]`` gather of model.py:34 is
  never materialised;
* when a down-transform is a pure vertex selection (all quadric-collapse
  down-transforms are), an encoder block convolves only the kept vertices --
  ``Pool(elu(conv(x)), down)`` has the same value there;
* CUDA fp32 only: CPU tensors or other dtypes raise instead of falling back.

Reference lines are cited per class.  Dense ``nn.Linear`` layers (0.3 % of the
FLOPs) stay on cuBLAS, as the reference has them.
"""
from __future__ import annotations

import torch
import torch.nn as nn

try:                                        # imported as sdvae_b200.model
    from . import cabi, functional as F_
    from .tables import pool_table, restricted_spiral_table, spiral_table
except ImportError:                         # imported as top-level ``model`` (drop-in use)
    import os as _os
    import sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
    from sdvae_b200 import cabi, functional as F_
    from sdvae_b200.tables import pool_table, restricted_spiral_table, spiral_table

__all__ = ['SpiralConv', 'Pool', 'SpiralEnblock', 'SpiralDeblock', 'Model', 'MLPClassifier']


class SpiralConv(nn.Module):
    """Spiral convolution, reference model.py:11-47.

    ``indices`` is the caller's ``LongTensor[V, S]`` (kept as a plain attribute, not a
    buffer, exactly like the reference, so it is absent from ``state_dict``); the
    learnable part is ``self.layer = nn.Linear(S*Cin, Cout)`` with xavier-uniform
    weight and zero bias."""

    def __init__(self, in_channels, out_channels, indices, dim=1):
        super().__init__()
        if indices.dim() != 2:
            raise ValueError('indices must be a [V, seq_length] tensor')
        self.dim = dim
        self.indices = indices
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.seq_length = indices.size(1)
        self.layer = nn.Linear(in_channels * self.seq_length, out_channels)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.layer.weight)
        nn.init.zeros_(self.layer.bias)

    def _run(self, x, act, table=None):
        squeeze = False
        if x.dim() == 2:                       # model.py:29-31
            x, squeeze = x.unsqueeze(0), True
        elif x.dim() == 3:
            if self.dim not in (1, -2):
                raise RuntimeError('SpiralConv on B200 gathers along dim=1 of [B, V, C]')
        else:
            raise RuntimeError(
                'x.dim() is expected to be 2 or 3, but received {}'.format(x.dim()))
        tab = table if table is not None else spiral_table(self.indices)
        y = F_.spiral_conv(x, self.layer.weight, self.layer.bias, tab, act)
        return y.squeeze(0) if squeeze else y

    def forward(self, x):
        return self._run(x, cabi.ACT_NONE)

    def __repr__(self):
        return '{}({}, {}, seq_length={})'.format(type(self).__name__, self.in_channels,
                                                  self.out_channels, self.seq_length)


def Pool(x, trans, dim=1):
    """Sparse up/down-sampling ``out[:, r] = sum_e val_e * x[:, col_e]``, reference
    model.py:50-55 (``trans`` = torch sparse COO, possibly uncoalesced)."""
    return F_.pool(x, trans, dim)


class SpiralEnblock(nn.Module):
    """``Pool(elu(conv(x)), down_transform)``, reference model.py:58-70."""

    def __init__(self, in_channels, out_channels, indices):
        super().__init__()
        self.conv = SpiralConv(in_channels, out_channels, indices)
        self.reset_parameters()

    def reset_parameters(self):
        self.conv.reset_parameters()

    def forward(self, x, down_transform):
        if x.dim() == 3:
            sub = restricted_spiral_table(self.conv.indices, pool_table(down_transform))
            if sub is not None:                # selection matrix: convolve kept vertices only
                return self.conv._run(x, cabi.ACT_ELU, sub)
        return Pool(self.conv._run(x, cabi.ACT_ELU), down_transform)


class SpiralDeblock(nn.Module):
    """``elu(conv(Pool(x, up_transform)))``, reference model.py:73-85."""

    def __init__(self, in_channels, out_channels, indices):
        super().__init__()
        self.conv = SpiralConv(in_channels, out_channels, indices)
        self.reset_parameters()

    def reset_parameters(self):
        self.conv.reset_parameters()

    def forward(self, x, up_transform):
        return self.conv._run(Pool(x, up_transform), cabi.ACT_ELU)


class Model(nn.Module):
    """SD-VAE encoder/decoder, reference model.py:88-188.

    Encoder: one ``SpiralEnblock`` per entry of ``out_channels`` followed by
    ``Linear(V_last*C_last, latent)`` -- two of them when ``is_vae`` (the LAST module
    yields ``mu``, the one before it ``logvar``, model.py:152-156).  Decoder:
    ``Linear(latent, V_last*C_last)``, the mirrored ``SpiralDeblock`` chain, and a
    final bare ``SpiralConv`` back to ``in_channels``."""

    def __init__(self, in_channels, out_channels, latent_size, spiral_indices, down_transform,
                 up_transform, pre_z_sigmoid=False, is_vae=False):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.latent_size = latent_size
        self.spiral_indices = spiral_indices
        self.down_transform = down_transform
        self.up_transform = up_transform
        self.num_vert = self.down_transform[-1].size(0)
        self.pre_z_sigmoid = pre_z_sigmoid
        self.is_vae = is_vae

        n_blocks = len(out_channels)
        widths = [in_channels] + list(out_channels)
        flat = self.num_vert * out_channels[-1]

        self.en_layers = nn.ModuleList(
            SpiralEnblock(widths[i], widths[i + 1], spiral_indices[i]) for i in range(n_blocks))
        for _ in range(2 if is_vae else 1):
            self.en_layers.append(nn.Linear(flat, latent_size))

        self.de_layers = nn.ModuleList([nn.Linear(latent_size, flat)])
        for lvl in reversed(range(n_blocks)):
            c_in = out_channels[min(lvl + 1, n_blocks - 1)]
            self.de_layers.append(SpiralDeblock(c_in, out_channels[lvl], spiral_indices[lvl]))
        self.de_layers.append(SpiralConv(out_channels[0], in_channels, spiral_indices[0]))
        self.reset_parameters()

    def reset_parameters(self):
        # model.py:139-144: every bias -> 0, every other parameter -> xavier-uniform
        for name, prm in self.named_parameters():
            if 'bias' in name:
                nn.init.zeros_(prm)
            else:
                nn.init.xavier_uniform_(prm)

    @property
    def _n_blocks(self):
        return len(self.out_channels)

    def encode(self, x):
        for i in range(self._n_blocks):
            x = self.en_layers[i](x, self.down_transform[i])
        x = x.reshape(-1, self.en_layers[-1].weight.size(1))
        mu = self.en_layers[-1](x)
        if self.is_vae:
            return mu, self.en_layers[-2](x)
        if self.pre_z_sigmoid:
            mu = torch.sigmoid(mu)
        return mu, None

    def decode(self, x):
        n = self._n_blocks
        x = self.de_layers[0](x).view(-1, self.num_vert, self.out_channels[-1])
        for i in range(1, n + 1):
            x = self.de_layers[i](x, self.up_transform[n - i])
        return self.de_layers[n + 1](x)

    def forward(self, x):
        mu, logvar = self.encode(x)
        z = self._reparameterize(mu, logvar) if (self.is_vae and self.training) else mu
        return self.decode(z), z, mu, logvar

    @staticmethod
    def _reparameterize(mu, logvar):
        # same generator consumption as the reference (one randn_like of [B, latent])
        return F_.reparameterize(mu, logvar, torch.randn_like(mu))


class MLPClassifier(nn.Module):
    """Latent-space classifier head, reference model.py:191-203 (off the hot path,
    plain torch): ``[Linear, ReLU] * n`` (ReLU after the last layer too), returns the
    activations and the arg-max label of their log-softmax."""

    def __init__(self, in_features, hidden_features, out_classes):
        super().__init__()
        sizes = [in_features] + list(hidden_features) + [out_classes]
        blocks = []
        for a, b in zip(sizes[:-1], sizes[1:]):
            blocks += [nn.Linear(a, b), nn.ReLU()]
        self.model = nn.Sequential(*blocks)

    def forward(self, x):
        out = self.model(x)
        return out, torch.log_softmax(out, dim=1).argmax(dim=1)
