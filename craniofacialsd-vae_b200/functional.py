"""Autograd bindings of the CUDA kernels (the layer the drop-in ``model.py`` sits on).

Each ``torch.autograd.Function`` here is the differentiable form of one reference
operator; forward and backward both go through the C ABI (``cabi.py``).  The
backward functions run on the autograd engine thread: they only use the tensors
saved in ``ctx`` and the stream current on that thread -- no module state.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import cabi
from .tables import PoolTable, SpiralTable, pool_table, spiral_table


def _prep(x: torch.Tensor, name: str) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError('sdvae_b200: %s is on %s; the B200 path has no CPU fallback' % (name, x.device))
    if x.dtype != torch.float32:
        raise TypeError('sdvae_b200: %s must be float32 (the reference computes in fp32), got %s'
                        % (name, x.dtype))
    return x.contiguous()


# ---------------------------------------------------------------------------
# SpiralConv (+ optional fused ELU)                        model.py:27-41, :68, :84
# ---------------------------------------------------------------------------
class SpiralConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, table: SpiralTable, act: int):
        B, Vin, Cin = x.shape
        Cout = weight.shape[0]
        y = torch.empty((B, table.n_rows, Cout), device=x.device, dtype=torch.float32)
        cabi.spiralconv_fwd(x, table.idx, weight, bias, y, B, Vin, table.n_rows, table.seq,
                            Cin, Cout, act)
        ctx.table, ctx.act = table, act
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight, y if act == cabi.ACT_ELU else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        table: SpiralTable = ctx.table
        B, Vin, Cin = x.shape
        Cout, S = weight.shape[0], table.seq
        dy = dy.contiguous()
        if ctx.act == cabi.ACT_ELU:
            dpre = torch.empty_like(dy)
            cabi.elu_bwd(dy, y, dpre)
        else:
            dpre = dy
        dx = dw = db = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw = torch.empty_like(weight)
            db = torch.empty(Cout, device=x.device, dtype=torch.float32) if ctx.has_bias else None
            ws = torch.empty(cabi.spiralconv_bwd_w_workspace(B * table.n_rows, S, Cin, Cout) // 4 + 4,
                             device=x.device, dtype=torch.float32)
            cabi.spiralconv_bwd_w(x, table.idx, dpre, dw, db, ws, B, Vin, table.n_rows, S, Cin, Cout)
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if table.n_rows * 2 <= Vin:
                # row-restricted (fused encoder block): per-slot gradients, then an
                # owner-computes row sum -- no work on the vertices that were dropped
                K = S * Cin
                wT = torch.empty((K, Cout), device=x.device, dtype=torch.float32)
                cabi.transpose2d(weight, wT, Cout, K)
                g = torch.empty((B, table.n_rows * S, Cin), device=x.device, dtype=torch.float32)
                cabi.dense_fwd(dpre, wT, None, g, B * table.n_rows, Cout, K, Cout, cabi.ACT_NONE)
                ptr, src = table.inverse_flat()
                cabi.csr_rowsum(g, ptr, src, None, None, dx, B, table.n_rows * S, Vin, Cin)
            else:
                wt = torch.empty((Cin, S * Cout), device=x.device, dtype=torch.float32)
                cabi.weight_transpose(weight, wt, Cout, Cin, S)
                cell_ptr, cell_src = table.inverse()
                cabi.spiralconv_bwd_x(dpre, cell_ptr, cell_src, wt, None, dx, B, table.n_rows,
                                      Vin, S, Cout, Cin)
        return dx, dw, db, None, None


def spiral_conv(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                table: SpiralTable, act: int = cabi.ACT_NONE) -> torch.Tensor:
    """``act(SpiralConv(x))`` for ``x [B, V, Cin]``; rows follow ``table`` (which may be a
    restriction of the full spiral table to the vertices a down-transform keeps)."""
    x = _prep(x, 'x')
    if x.dim() != 3:
        raise RuntimeError('x.dim() is expected to be 2 or 3, but received {}'.format(x.dim()))
    if x.shape[1] != table.n_src:
        raise RuntimeError('SpiralConv: x has %d vertices, spiral table indexes %d'
                           % (x.shape[1], table.n_src))
    if x.shape[2] * table.seq != weight.shape[1]:
        raise RuntimeError('SpiralConv: x has %d channels, weight expects %d'
                           % (x.shape[2], weight.shape[1] // table.seq))
    return SpiralConvFn.apply(x, _prep(weight, 'weight'),
                              None if bias is None else _prep(bias, 'bias'), table, act)


# ---------------------------------------------------------------------------
# Pool                                                                model.py:50-55
# ---------------------------------------------------------------------------
class PoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, table: PoolTable):
        B, Vin, C = x.shape
        out = torch.empty((B, table.n_rows, C), device=x.device, dtype=torch.float32)
        cabi.pool_ell_fwd(x, table.ell_col, table.ell_val, out, B, Vin, table.n_rows,
                          table.width, C)
        ctx.table = table
        return out

    @staticmethod
    def backward(ctx, dout):
        table: PoolTable = ctx.table
        dout = dout.contiguous()
        B, _, C = dout.shape
        dx = torch.empty((B, table.n_cols, C), device=dout.device, dtype=torch.float32)
        cabi.csr_rowsum(dout, table.t_ptr, table.t_row, table.t_val, None, dx, B, table.n_rows,
                        table.n_cols, C)
        return dx, None


def pool(x: torch.Tensor, trans: torch.Tensor, dim: int = 1) -> torch.Tensor:
    """``Pool(x, trans, dim)``: sparse ``trans [Vout, Vin]`` applied along the vertex axis."""
    x = _prep(x, 'x')
    table = pool_table(trans)
    if x.dim() == 2 and dim in (0, -2):
        return PoolFn.apply(x.unsqueeze(0), table).squeeze(0)
    if x.dim() != 3 or dim not in (1, -2):
        raise RuntimeError('Pool expects x [B, V, C] with dim=1, got shape %s dim=%d'
                           % (tuple(x.shape), dim))
    if x.shape[1] != table.n_cols:
        raise RuntimeError('Pool: x has %d vertices, transform expects %d' % (x.shape[1], table.n_cols))
    return PoolFn.apply(x, table)


# ---------------------------------------------------------------------------
# z = mu + eps * exp(logvar / 2)                                    model.py:184-188
# ---------------------------------------------------------------------------
class ReparamFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar, eps):
        z = torch.empty_like(mu)
        cabi.reparam_fwd(mu, logvar, eps, z)
        ctx.save_for_backward(logvar, eps)
        return z

    @staticmethod
    def backward(ctx, dz):
        logvar, eps = ctx.saved_tensors
        dz = dz.contiguous()
        dmu, dlv = torch.empty_like(dz), torch.empty_like(dz)
        cabi.reparam_bwd(dz, logvar, eps, dmu, dlv, False)
        return dmu, dlv, None


def reparameterize(mu, logvar, eps):
    return ReparamFn.apply(_prep(mu, 'mu'), _prep(logvar, 'logvar'), _prep(eps, 'eps'))
