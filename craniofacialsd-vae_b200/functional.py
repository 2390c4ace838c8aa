"""Autograd bindings of the CUDA kernels (the layer the drop-in ``model.py`` sits on).

Each ``torch.autograd.Function`` here is the differentiable form of one reference
operator; forward and backward both go through the C ABI (``cabi.py``).  The
backward functions run on the autograd engine thread: they only use the tensors
saved in ``ctx`` and the stream current on that thread -- no module state.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import cabi
from .tables import PoolTable, SpiralTable, identity_plan, pool_table, spiral_table

# The SpiralConv passes run on the tcgen05 tensor-core kernels (error-compensated 3xTF32, results within
# 1e-5 of fp64 per layer, tests/test_gpu_tc.py) wherever the layer shape is supported, and on the fp32-FMA
# kernels otherwise.  ``set_tensor_cores(False)`` (or SDVAE_NO_TC=1) keeps everything on the FMA kernels.
_USE_TC = os.environ.get('SDVAE_NO_TC', '0') != '1'


def set_tensor_cores(flag: bool) -> None:
    global _USE_TC
    _USE_TC = bool(flag)


def tensor_cores_enabled() -> bool:
    return _USE_TC


def _f32(*shape, like):
    return torch.empty(shape, device=like.device, dtype=torch.float32)


def _aligned(*ts) -> bool:
    return all(t is None or t.data_ptr() % 16 == 0 for t in ts)


def _tc_parts(S, ks, n, rcap):
    """Output-channel passes of a tensor-core convolution: all at once, two passes of 32 for 64 -> 64
    (the weight image of all 64 does not fit in shared memory), or None (unsupported shape)."""
    if ks not in (32, 64):
        return None
    if cabi.tc_supported(S, ks, n, rcap):
        return [(0, n)]
    if n == 64 and cabi.tc_supported(S, ks, 32, rcap):
        return [(0, 32), (32, 32)]
    return None


def _slot_ok(S, narrow, wide, rows) -> bool:
    """3-channel side slot-packed into a dense 32 x 32 contraction (csrc/slot_pack.cuh); ``rows`` = rows of
    the dense problem (identity tile plans carry 16-bit rows)."""
    return S * narrow <= 32 and wide == 32 and rows < 65536 and cabi.tc_supported(1, 32, 32, 128)


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
        R, S = table.n_rows, table.seq
        y = torch.empty((B, R, Cout), device=x.device, dtype=torch.float32)
        packed = None                       # slot-packed input of a 3-channel first layer, kept for dW
        done = False
        tp = table.tile_fwd() if (_USE_TC and B > 0 and act == cabi.ACT_NONE and Cout == 3 and Cin == 32
                                  and _aligned(x, weight, bias)) else None
        if tp is not None and cabi.narrow_out_fwd_tc_supported(S, Cin, Cout, tp.rcap):
            # 3-channel OUTPUT layer on tcgen05 by project-then-gather (csrc/spiral_conv_tile_out.cuh); needs a tile plan
            # with <= 256 distinct source rows per tile (patch-ordered levels)
            cabi.narrow_out_fwd_tc(x, tp, weight, bias, y, B, Vin, R, S, Cin, Cout)
            done = True
        elif _USE_TC and B > 0 and _aligned(x, weight, bias) and act == cabi.ACT_NONE and Cout == 3 and Cin == 32 \
                and S == 9 and cabi.narrow_out_fwd_supported(S, Cin, Cout, table.stage_plan().ucap):
            # 3-channel OUTPUT layer: fp32 FMA over shared-memory-staged source rows (csrc/narrow_conv.cuh)
            cabi.narrow_out_fwd(x, table.stage_plan(), weight, bias, y, B, Vin, R, S, Cin, Cout)
            done = True
        elif _USE_TC and B > 0 and cabi.narrow_in_supported(Vin, S, Cin, Cout):
            # 3-channel INPUT layer: fp32 FMA with the mesh's input resident in shared memory (csrc/narrow_conv.cuh)
            cabi.narrow_in_fwd(x, table.idx, weight, bias, y, B, Vin, R, S, Cin, Cout, act)
            done = True
        elif _USE_TC and B > 0 and _aligned(x, weight, bias):
            plan = table.plan_fwd()
            parts = _tc_parts(S, Cin, Cout, plan.rcap)
            if parts is not None:
                for n0, nc in parts:
                    wimg = _f32(cabi.tc_wimg_floats(S, Cin, nc), like=x)
                    cabi.tc_pack_weights(weight, wimg, S, Cin, Cout, False, n0, nc)
                    full = nc == Cout
                    cabi.spiralconv_fwd_tc(x, plan, wimg, None if bias is None else bias[n0:],
                                           y if full else y.view(-1)[n0:], B, Vin, R, S, Cin, nc, act,
                                           0 if full else Cout)
                done = True
            elif _slot_ok(S, Cin, Cout, R):
                packed = _f32(B, R, 32, like=x)
                cabi.slot_pack(x, None, table.idx, packed, B, Vin, R, S, Cin)
                wd, wimg = _f32(1024, like=x), _f32(cabi.tc_wimg_floats(1, 32, 32), like=x)
                cabi.slot_weight(weight, wd, 0, Cout, S, Cin)
                cabi.tc_pack_weights(wd, wimg, 1, 32, 32, False)
                cabi.dense_tc(packed, identity_plan(R, x.device), wimg, bias, None, y, B, R, act)
                done = True
        if not done:
            cabi.spiralconv_fwd(x, table.idx, weight, bias, y, B, Vin, R, S, Cin, Cout, act)
        ctx.table, ctx.act = table, act
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight, y if act == cabi.ACT_ELU else None, packed)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y, packed = ctx.saved_tensors
        table: SpiralTable = ctx.table
        B, Vin, Cin = x.shape
        Cout, S, R = weight.shape[0], table.seq, table.n_rows
        dy = dy.contiguous()
        if ctx.act == cabi.ACT_ELU:
            dpre = torch.empty_like(dy)
            cabi.elu_bwd(dy, y, dpre)
        else:
            dpre = dy
        dx = dw = db = None
        want_w = ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        want_x = ctx.needs_input_grad[0]
        tc = _USE_TC and B > 0 and _aligned(x, weight, dpre)
        if want_w:
            dw = torch.empty_like(weight)
            db = torch.empty(Cout, device=x.device, dtype=torch.float32) if ctx.has_bias else None
            ws = torch.empty(cabi.spiralconv_bwd_w_workspace(B * max(R, Vin), S, max(Cin, 32), max(Cout, 32)) // 4 + 4,
                             device=x.device, dtype=torch.float32)
        tpb = table.tile_bwd() if (tc and want_w and want_x and ctx.has_bias and Cin == 32 and Cout <= 3) else None
        if tpb is not None and cabi.narrow_out_bwd_tc_supported(S, Cin, Cout, tpb.rcap, tpb.ecap):
            # 3-channel OUTPUT layer, whole backward in one tcgen05 pass (csrc/spiral_conv_tile_out_bw.cuh); needs the
            # inverse tile plan (patch-ordered levels) and computes all three gradients
            dx = torch.empty_like(x)
            nws = _f32(cabi.narrow_out_bwd_tc_workspace(S, Cout) // 4, like=x)
            cabi.narrow_out_bwd_tc(dpre, x, tpb, weight, dx, dw, db, nws, B, R, Vin, S, Cin, Cout, False)
            want_w = want_x = False
        elif tc and (want_w or want_x) and cabi.narrow_out_bwd_supported(R, S, Cin, Cout):
            # 3-channel OUTPUT layer, fused (csrc/narrow_conv.cuh): dx, dW, db from one pass over dpre and x
            cell_ptr, cell_src = table.inverse()
            nws = _f32(cabi.narrow_out_bwd_workspace(S, Cout) // 4, like=x)
            if want_x:
                dx = torch.empty_like(x)
            cabi.narrow_out_bwd(dpre, x, cell_ptr, cell_src, table.inverse_packed(), weight, dx, dw if want_w else None,
                                db if want_w else None, nws, B, R, Vin, S, Cin, Cout, False)
            want_w = want_x = False
        elif tc and _slot_ok(S, Cout, Cin, Vin) and (want_w or want_x):
            # 3-channel OUTPUT layer: G[u, s*C + n] = sum of dpre over the rows that gather u at slot s, then
            # dW = (G^T x) re-indexed and dx = G Wd^T -- two dense 32 x 32 contractions
            cell_ptr, cell_src = table.inverse()
            G = _f32(B, Vin, 32, like=x)
            cabi.slot_pack(dpre, cell_ptr, cell_src, G, B, R, Vin, S, Cout)
            ident = identity_plan(Vin, x.device)
            if want_w:
                dwd, dbd = _f32(32, 32, like=x), _f32(32, like=x)
                cabi.spiralconv_bwd_w_tc(x, ident, G, dwd, dbd, ws, B, Vin, Vin, 1, 32, 32)
                cabi.slot_grad(dwd, dbd, dw, db, 1, Cout, S, Cout)
                want_w = False
            if want_x:
                wd, wimg = _f32(1024, like=x), _f32(cabi.tc_wimg_floats(1, 32, 32), like=x)
                cabi.slot_weight(weight, wd, 1, Cout, S, Cout)
                cabi.tc_pack_weights(wd, wimg, 1, 32, 32, False)
                dx = torch.empty_like(x)
                cabi.dense_tc(G, ident, wimg, None, None, dx, B, Vin, cabi.ACT_NONE)
                want_x = False
        if want_w:
            if _USE_TC and B > 0 and cabi.narrow_in_supported(Vin, S, Cin, Cout):
                nws = _f32(cabi.narrow_in_bwd_w_workspace(S, Cin) // 4, like=x)
                cabi.narrow_in_bwd_w(x, table.idx, dpre, dw, db, nws, B, Vin, R, S, Cin, Cout)
            elif tc and packed is not None:
                # 3-channel INPUT layer: dW = dpre^T P with the slot-packed input of the forward pass
                dwd, dbd = _f32(32, 32, like=x), _f32(32, like=x)
                cabi.spiralconv_bwd_w_tc(packed, identity_plan(R, x.device), dpre, dwd, dbd, ws, B, R, R, 1, 32, 32)
                cabi.slot_grad(dwd, dbd, dw, db, 0, Cout, S, Cin)
            elif tc and cabi.tc_bwd_w_supported(S, Cin, Cout, table.plan_fwd().rcap):
                cabi.spiralconv_bwd_w_tc(x, table.plan_fwd(), dpre, dw, db, ws, B, Vin, R, S, Cin, Cout)
            else:
                cabi.spiralconv_bwd_w(x, table.idx, dpre, dw, db, ws, B, Vin, R, S, Cin, Cout)
        if want_x:
            dx = torch.empty_like(x)
            parts = _tc_parts(S, Cout, Cin, table.plan_bwd().rcap) if tc else None
            if parts is not None:
                # inverse-table plan (of the row-restricted table for a fused encoder block): deterministic
                # scatter-add as an in-order sum inside the A-operand cells
                plan = table.plan_bwd()
                for n0, nc in parts:
                    wimg = _f32(cabi.tc_wimg_floats(S, Cout, nc), like=x)
                    cabi.tc_pack_weights(weight, wimg, S, Cin, Cout, True, n0, nc)
                    full = nc == Cin
                    cabi.spiralconv_bwd_x_tc(dpre, plan, wimg, None, dx if full else dx.view(-1)[n0:], B, R, Vin,
                                             S, Cout, nc, 0 if full else Cin)
            elif table.n_rows * 2 <= Vin:
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
    with torch.cuda.device(x.device):        # launches go to the CURRENT device's stream: make it x's device
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
        cabi.pool_fwd(x, table, out, B, Vin, C)
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
        with torch.cuda.device(x.device):
            return PoolFn.apply(x.unsqueeze(0), table).squeeze(0)
    if x.dim() != 3 or dim not in (1, -2):
        raise RuntimeError('Pool expects x [B, V, C] with dim=1, got shape %s dim=%d'
                           % (tuple(x.shape), dim))
    if x.shape[1] != table.n_cols:
        raise RuntimeError('Pool: x has %d vertices, transform expects %d' % (x.shape[1], table.n_cols))
    with torch.cuda.device(x.device):
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
    with torch.cuda.device(mu.device):
        return ReparamFn.apply(_prep(mu, 'mu'), _prep(logvar, 'logvar'), _prep(eps, 'eps'))
