"""Differentiable loss reductions of the training step on the CUDA kernels.

Mirrors the loss methods of ``ModelManager`` (reference model_manager.py):

=============================  ==========================================  =================
this module                    reference                                   lines
=============================  ==========================================  =================
``l1_loss``                    ``_compute_l1_loss``                        :328-330
``mse_loss``                   ``compute_mse_loss``                        :332-334
``laplacian_regularizer``      ``_compute_laplacian_regularizer``          :343-349
``mse_and_laplacian``          both from one read of the reconstruction    --
``kl_divergence``              ``_compute_kl_divergence_loss``             :351-354
``latent_consistency``         ``_compute_latent_consistency``             :360-393
=============================  ==========================================  =================

Each forward launches the value kernel and (for KL / latent consistency) already
produces the unit gradient, so backward is a scale.  All reductions have a fixed
order: results are bit-identical run to run.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import cabi
from .tables import ell_from_coo, transposed_csr


@dataclass
class LaplacianTable:
    """Random-walk Laplacian (utils.py:87-90) as ELL rows + CSR of its transpose."""
    n_vert: int
    width: int
    ell_col: torch.Tensor
    ell_val: torch.Tensor
    t_ptr: torch.Tensor
    t_row: torch.Tensor
    t_val: torch.Tensor
    coo: Optional[tuple] = None      # the (row, col, val) it was built from, in storage order (host, for renumbered)

    @staticmethod
    def build(row, col, val, n_vert: int, device) -> "LaplacianTable":
        row, col = np.asarray(row, np.int64), np.asarray(col, np.int64)
        val = np.asarray(val, np.float32)
        ec, ev = ell_from_coo(row, col, val, n_vert, n_vert)
        tp, tr, tv = transposed_csr(row, col, val, n_vert)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        return LaplacianTable(int(n_vert), int(ec.shape[1]), d(ec), d(ev), d(tp), d(tr), d(tv), (row, col, val))

    def renumbered(self, order: np.ndarray) -> "LaplacianTable":
        """The same operator with its vertices listed in ``order`` (new position -> old vertex); the entries keep
        their storage order (rows of the operator and columns of its transpose sum in the same order as before)."""
        order = np.asarray(order, np.int64)
        rank = np.empty(order.size, np.int64)
        rank[order] = np.arange(order.size)
        row, col, val = self.coo
        return LaplacianTable.build(rank[row], rank[col], val, self.n_vert, self.ell_col.device)

    @staticmethod
    def from_sparse(lap: torch.Tensor) -> "LaplacianTable":
        """From the reference's ``template.laplacian`` sparse COO tensor."""
        ind = lap._indices().detach().cpu().numpy()
        return LaplacianTable.build(ind[0], ind[1], lap._values().detach().cpu().numpy(),
                                    int(lap.shape[0]), lap.device)


def _check(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda or t.dtype != torch.float32:
        raise RuntimeError('sdvae_b200.losses: %s must be a CUDA float32 tensor' % name)
    return t.contiguous()


class MseLapFn(torch.autograd.Function):
    """(mse, laplacian) of a reconstruction ``[B, V, 3]``; ``lap`` may be None."""

    @staticmethod
    def forward(ctx, recon, target, lap: Optional[LaplacianTable]):
        B, V, C = recon.shape
        if C != 3:
            raise RuntimeError('mse/laplacian kernels expect xyz vertices ([B, V, 3])')
        dev = recon.device
        losses = torch.zeros(8, device=dev, dtype=torch.float32)
        partial = torch.empty(cabi.mse_lap_partial_floats(B, V), device=dev, dtype=torch.float32)
        qn = torch.empty_like(recon) if lap is not None else None
        if lap is not None:
            cabi.mse_lap_fwd(recon, target, lap.ell_col, lap.ell_val, lap.width, qn, partial,
                             losses, B, V)
        else:
            cabi.mse_lap_fwd(recon, target, None, None, 0, None, partial, losses, B, V)
        ctx.lap = lap
        ctx.save_for_backward(recon, target, qn)
        return losses[0].clone(), losses[3].clone()

    @staticmethod
    def backward(ctx, g_mse, g_lap):
        recon, target, qn = ctx.saved_tensors
        lap = ctx.lap
        B, V, _ = recon.shape
        dscale = torch.stack([g_mse, g_lap]).to(torch.float32).contiguous()
        d = torch.empty_like(recon)
        if lap is not None:
            cabi.mse_lap_bwd(recon, target, qn, lap.t_ptr, lap.t_row, lap.t_val, d, B, V,
                             1.0, 1.0, 1.0, dscale)
        else:
            cabi.mse_lap_bwd(recon, target, None, None, None, None, d, B, V, 1.0, 0.0, 1.0, dscale)
        return d, None, None


def mse_and_laplacian(prediction, gt, lap: LaplacianTable):
    """Both reconstruction terms from a single pass over ``prediction``."""
    return MseLapFn.apply(_check(prediction, 'prediction'), _check(gt, 'gt'), lap)


class L1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prediction, gt):
        out = torch.zeros(1, device=prediction.device, dtype=torch.float32)
        partial = torch.empty((prediction.numel() + 255) // 256, device=prediction.device, dtype=torch.float32)
        cabi.l1_fwd(prediction, gt, partial, out)
        ctx.save_for_backward(prediction, gt)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        prediction, gt = ctx.saved_tensors
        d = torch.empty_like(prediction)
        cabi.l1_bwd(prediction, gt, d, 1.0)
        return d * g, None


def l1_loss(prediction, gt):
    """``torch.nn.L1Loss(reduction='mean')`` (model_manager.py:328-330; defined by the reference, never called by
    its training step)."""
    return L1Fn.apply(_check(prediction, 'prediction'), _check(gt, 'gt'))


def mse_loss(prediction, gt):
    """``torch.nn.MSELoss(reduction='mean')`` (model_manager.py:332-334)."""
    return MseLapFn.apply(_check(prediction, 'prediction'), _check(gt, 'gt'), None)[0]


def laplacian_regularizer(prediction, lap: LaplacianTable):
    """``sum_b sum_v |(L pred_b)_v|_2 / V / B`` (model_manager.py:343-349)."""
    p = _check(prediction, 'prediction')
    return MseLapFn.apply(p, p.detach(), lap)[1]


class KlFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar):
        B, D = mu.shape
        dev = mu.device
        losses = torch.zeros(8, device=dev, dtype=torch.float32)
        partial = torch.empty((B * D + 255) // 256, device=dev, dtype=torch.float32)
        dmu, dlv = torch.empty_like(mu), torch.empty_like(mu)
        cabi.kl_fwd_bwd(mu, logvar, dmu, dlv, partial, losses, B, D)
        ctx.save_for_backward(dmu, dlv)
        return losses[1].clone()

    @staticmethod
    def backward(ctx, g):
        dmu, dlv = ctx.saved_tensors
        return dmu * g, dlv * g


def kl_divergence(mu, logvar):
    """``mean_b(-1/2 sum_d(1 + logvar - mu^2 - exp(logvar)))`` (model_manager.py:351-354)."""
    return KlFn.apply(_check(mu, 'mu'), _check(logvar, 'logvar'))


class LatentConsistencyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, bs, r0, r1, eta1, eta2):
        n, D = z.shape
        if n != bs * bs:
            raise RuntimeError('latent consistency expects bs*bs = %d latents, got %d' % (bs * bs, n))
        dev = z.device
        losses = torch.zeros(8, device=dev, dtype=torch.float32)
        nh = bs * (bs - 1) // 2 * bs
        act = torch.empty(2 * nh, device=dev, dtype=torch.uint8)
        partial = torch.empty((nh + 255) // 256, device=dev, dtype=torch.float32)
        dz = torch.empty_like(z)
        cabi.lc_fwd_bwd(z, bs, D, r0, r1, eta1, eta2, act, partial, dz, losses)
        ctx.save_for_backward(dz)
        return losses[2].clone()

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return dz * g, None, None, None, None, None


def latent_consistency(z, batch_size: int, latent_region: Sequence[int], eta1: float, eta2: float):
    """Latent-consistency loss on the ``bs x bs`` swap grid (model_manager.py:360-393).
    ``latent_region = [r0, r1]`` is the latent slice of the swapped feature
    (``ModelManager.latent_regions[data.swapped]``, model_manager.py:232-238, :364)."""
    return LatentConsistencyFn.apply(_check(z, 'z'), int(batch_size), int(latent_region[0]),
                                     int(latent_region[1]), float(eta1), float(eta2))
