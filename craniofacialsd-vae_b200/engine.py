"""Fused training step for the SD-VAE hot path (one process per GPU).

``TrainEngine.step`` is the B200 form of ``ModelManager._do_iteration`` (reference
model_manager.py:274-326): feature swap, forward, the four loss terms, backward,
Adam -- as an explicit sequence of kernel launches through the C ABI on
preallocated buffers, without autograd:

* the feature swap (swap_batch_transform.py:13-52) runs on the device, so only
  the ``bs`` un-swapped meshes cross PCIe, not the ``bs^2`` swapped ones;
* encoder blocks convolve only the vertices their selection down-transform keeps;
* the ELU derivative is applied in the epilogue of the kernel that *produces* a
  gradient (backward-to-input conv or pool-transpose), so every gradient tensor is
  written once, already w.r.t. the pre-activation;
* all parameters / gradients / Adam moments live in flat arenas ordered in
  backward-ready order, so the data-parallel exchange is ONE NCCL all-reduce of a
  contiguous arena (gradients + loss scalars) after the backward pass -- or, opt-in,
  four contiguous buckets on a side stream overlapping the backward pass, which
  measured slower next to the persistent kernels; Adam is one launch over the arena;
* the seven loss scalars come back in ONE 28-byte D2H copy (the reference issues
  seven ``.item()`` syncs, model_manager.py:320-326);
* the whole step -- data-parallel collectives included -- is replayed from a CUDA graph
  (one per swapped region, since the region's latent slice is a kernel argument);
* the 32/64-channel SpiralConv passes run on the tcgen05 tensor-core kernels (``use_tc``)
  through tile plans of the spiral / inverse tables (64 -> 64 layers in two passes); the two
  3-channel layers run on the fp32 FMA units with one mesh resident in shared memory
  (csrc/narrow_conv.cuh; slot packing into dense 32 x 32 contractions, csrc/slot_pack.cuh,
  is their fallback); Pool forward stages each tile's distinct source rows in shared memory.

Data parallelism: rank r owns rows ``[r*bs/N, (r+1)*bs/N)`` of the
``bs x bs`` swap grid.  Mean losses are normalised by the GLOBAL counts, the
latent-consistency loss is evaluated on the all-gathered ``z`` on every rank and
only the local slice of its gradient is back-propagated, so a SUM all-reduce of the
gradient arenas reproduces the single-GPU gradient.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import cabi
from .losses import LaplacianTable
from .parallel import grid_rows, mean_loss_scale
from .tables import (PoolTable, SpiralTable, identity_plan, pool_table, restricted_spiral_table,
                     spiral_table)

LOSS_KEYS = ['reconstruction', 'kl', 'latent_consistency', 'laplacian',
             'classification', 'classification_acc', 'tot']      # model_manager.py:150-154


@dataclass
class StepConfig:
    """The ``optimization`` section of the reference YAML (craniofacial.yaml:16-27)."""
    batch_size: int = 4                  # bs; the step processes bs*bs swapped meshes
    lr: float = 1e-4
    weight_decay: float = 0.0
    laplacian_weight: float = 0.1
    kl_weight: float = 1e-4
    latent_consistency_weight: float = 0.5
    latent_consistency_eta1: float = 0.5
    latent_consistency_eta2: float = 0.5
    betas: tuple = (0.9, 0.999)
    eps: float = 1e-8


class TrainEngine:
    def __init__(self, model, laplacian: Optional[LaplacianTable], region_features: Sequence[np.ndarray],
                 latent_regions: Sequence[Sequence[int]], cfg: StepConfig, process_group=None,
                 use_graph: bool = True, use_tc: bool = True, renumber: Optional[bool] = None):
        """``model``: a ``sdvae_b200.model.Model`` on the CUDA device.
        ``region_features[k]``: vertex ids swapped for region k; ``latent_regions[k]`` = [r0, r1].
        ``use_tc``: run the wide SpiralConv contractions (C_in in {32, 64}) on the tcgen05 tensor-core
        kernels (error-compensated 3xTF32); False keeps every contraction on the fp32-FMA kernels.
        ``renumber`` (default: on with ``use_tc``; env ``SDVAE_RENUMBER=0`` turns it off): run the network on a
        patch-wise renumbering of the internal vertex levels (``tables.renumbered_model_tables``; the batch is
        permuted on load, ``recon_template_order()`` undoes it) -- what the tile-staged tcgen05 kernels
        (csrc/spiral_conv_tile.cuh) need: a tile of 128 consecutive rows then reads ~200 distinct source rows
        instead of ~410.  Losses and parameter gradients do not depend on the vertex order."""
        self.model = model
        self.cfg = cfg
        self.dev = next(model.parameters()).device
        if self.dev.type != 'cuda':
            raise RuntimeError('TrainEngine needs the model on a CUDA device (no CPU fallback)')
        self.pg = process_group
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
            self.rank = dist.get_rank(process_group)
        else:
            self.world, self.rank = 1, 0
        bs = cfg.batch_size
        self.bs = bs
        self.i0, i1 = grid_rows(bs, self.world, self.rank)
        self.rows = i1 - self.i0
        self.B = self.rows * bs                          # local meshes per step
        self.is_vae = bool(model.is_vae)
        self.L = len(model.out_channels)
        # the Laplacian term is always EVALUATED (model_manager.py:283 computes it whatever its weight); with weight 0
        # only its backward is skipped
        self.lap = laplacian
        self.latent_regions = [tuple(int(t) for t in r) for r in latent_regions]
        self.use_lc = cfg.latent_consistency_weight > 0
        # The data-parallel step is captured too (NCCL collectives are graph-capturable; the side stream that
        # carries the bucket all-reduces forks from and joins back into the capture stream).  At 8 GPUs the
        # step is 5 ms, and ~130 eager launches through ctypes were a visible part of it.
        # SDVAE_DP_GRAPH=0 keeps multi-GPU steps eager.
        self.use_graph = use_graph and (self.world == 1 or os.environ.get('SDVAE_DP_GRAPH', '1') != '0')
        self.use_tc = bool(use_tc)
        # data-parallel gradient exchange: ONE all-reduce of the gradient arena after the backward pass on the compute
        # stream (default), or bucketed all-reduces on a side stream overlapped with the backward pass
        # (SDVAE_DP_OVERLAP=1).  The overlapped form measured SLOWER (12.48 vs 12.23 ms at 2 GPUs): the persistent
        # one-CTA-per-SM kernels leave no SM for a concurrent NCCL kernel, which then delays the CTAs of the next
        # kernel by its own duration, while the 4.3 MB all-reduce alone costs ~50 us over NVSwitch
        # (profiles/r02_dp_floor.md)
        self.dp_overlap = os.environ.get('SDVAE_DP_OVERLAP', '0') == '1'
        if renumber is None:
            renumber = self.use_tc and os.environ.get('SDVAE_RENUMBER', '1') != '0'
        self.renumber = bool(renumber)
        self._graphs: Dict[int, torch.cuda.CUDAGraph] = {}
        self.fixed_eps: Optional[torch.Tensor] = None
        self.launches_per_step = 0

        self._build_tables()
        V0 = self.V[0]
        masks = np.zeros((len(region_features), V0), np.uint8)
        for k, idx in enumerate(region_features):
            masks[k, np.asarray(idx, np.int64)] = 1
        if self.order0 is not None:                       # engine-internal vertex order (renumber)
            masks = np.ascontiguousarray(masks[:, self.order0])
            if self.lap is not None:
                self.lap = self.lap.renumbered(self.order0)
        self.masks = torch.from_numpy(masks).to(self.dev)
        with torch.cuda.device(self.dev):
            self._build_arenas()
            self._alloc_buffers()
            self.side = torch.cuda.Stream(device=self.dev) if self.world > 1 else None
            self.side_lc = torch.cuda.Stream(device=self.dev) if self.world > 1 else None
            self.copy_stream = torch.cuda.Stream(device=self.dev)
            self._staged = None                      # event of a batch copied to x_stage and not yet consumed
            self._stage_free = None                  # event: the last step has read x_stage
            if self.world > 1:
                self._sync_replicas()

    # ------------------------------------------------------------------ tables
    def _build_tables(self):
        m = self.model
        spirals, downs, ups = list(m.spiral_indices), list(m.down_transform), list(m.up_transform)
        self.order0 = self.order0_dev = None
        if self.renumber:
            from .tables import renumbered_model_tables
            spirals, downs, ups, orders = renumbered_model_tables(spirals, downs, ups)
            self._renumbered_tables = (spirals, downs, ups)   # derived tables are cached on these tensors: keep them
            self.order0 = orders[0]
            self.order0_dev = torch.from_numpy(orders[0]).to(self.dev)
        self.V = [int(s.shape[0]) for s in spirals] + [int(m.num_vert)]
        self.C = [int(m.in_channels)] + [int(c) for c in m.out_channels]
        self.S = [int(s.shape[1]) for s in spirals]
        self.full: List[SpiralTable] = [spiral_table(s) for s in spirals]
        self.sub: List[SpiralTable] = []
        self.up: List[PoolTable] = [pool_table(u) for u in ups]
        for lvl in range(self.L):
            sub = restricted_spiral_table(spirals[lvl], pool_table(downs[lvl]))
            if sub is None:
                raise RuntimeError(
                    'TrainEngine: down_transform[%d] is not a pure vertex selection; use the '
                    'autograd path (sdvae_b200.model.Model + losses) for general matrices' % lvl)
            self.sub.append(sub)

    # ------------------------------------------------------------------ arenas
    def _build_arenas(self):
        m, L = self.model, self.L
        de_convs = [m.de_layers[L + 1].layer] + [m.de_layers[i].conv.layer for i in range(L, 0, -1)]
        lin = [m.de_layers[0]] + [m.en_layers[i] for i in range(len(m.en_layers) - 1, L - 1, -1)]
        en_convs = [m.en_layers[i].conv.layer for i in range(L - 1, -1, -1)]
        groups = [de_convs, lin[:1], lin[1:], en_convs]       # backward-ready order
        prms = []
        self.buckets = []
        off = 0
        for g in groups:
            start = off
            for mod in g:
                for p in (mod.weight, mod.bias):
                    prms.append(p)
                    off += (p.numel() + 3) // 4 * 4            # keep every tensor 16-byte aligned
            self.buckets.append((start, off))
        n = off
        self.n_params = sum(p.numel() for p in prms)
        self.flat_p = torch.zeros(n, device=self.dev, dtype=torch.float32)
        # 8 more floats behind the gradients: the loss scalars, so that under data parallelism they ride in the last
        # gradient bucket's all-reduce instead of a collective of their own
        self.n_arena = n
        self.flat_g_all = torch.zeros(n + 8, device=self.dev, dtype=torch.float32)
        self.flat_g = self.flat_g_all[:n]
        self.flat_m = torch.zeros(n, device=self.dev, dtype=torch.float32)
        self.flat_v = torch.zeros(n, device=self.dev, dtype=torch.float32)
        self.step_dev = torch.zeros(1, device=self.dev, dtype=torch.int32)
        self.grad: Dict[int, torch.Tensor] = {}
        off = 0
        for p in prms:
            k = p.numel()
            view = self.flat_p[off:off + k].view(p.shape)
            view.copy_(p.data)
            p.data = view                                       # the module now lives in the arena
            self.grad[id(p)] = self.flat_g[off:off + k].view(p.shape)
            off += (k + 3) // 4 * 4

    def g(self, p):
        return self.grad[id(p)]

    def _sync_replicas(self):
        """Data-parallel start-up: every rank takes rank 0's parameters / Adam state (only GRADIENTS are all-reduced
        afterwards, so replicas that start different stay different), and the re-parameterisation noise of each rank
        comes from its own stream of the device generator -- with one seed everywhere every rank would draw the SAME
        eps for its rows of the swap grid, unlike the single-GPU step."""
        import torch.distributed as dist
        for t in (self.flat_p, self.flat_m, self.flat_v):
            dist.broadcast(t, src=dist.get_global_rank(self.pg, 0) if hasattr(dist, 'get_global_rank') else 0, group=self.pg)
        step = self.step_dev.to(torch.int64)
        dist.broadcast(step, src=dist.get_global_rank(self.pg, 0) if hasattr(dist, 'get_global_rank') else 0, group=self.pg)
        self.step_dev.copy_(step.to(torch.int32))
        seed = torch.tensor([torch.cuda.initial_seed()], device=self.dev, dtype=torch.int64)
        dist.broadcast(seed, src=dist.get_global_rank(self.pg, 0) if hasattr(dist, 'get_global_rank') else 0, group=self.pg)
        torch.cuda.manual_seed(int(seed.item()) + 7919 * (self.rank + 1))

    def _check_arena(self):
        """The kernels read the weights through pointers into the flat arena taken at construction: a later
        ``model.to()`` / ``.half()`` / parameter re-assignment would leave them on stale memory -- fail loudly."""
        for p in self.model.parameters():
            v = self.grad.get(id(p))
            if v is None or p.data.dtype != torch.float32 or not (
                    self.flat_p.data_ptr() <= p.data.data_ptr() < self.flat_p.data_ptr() + self.flat_p.numel() * 4):
                raise RuntimeError('TrainEngine: a model parameter no longer lives in the engine arena (the model was '
                                   'moved / cast / re-assigned after the engine was built); rebuild the TrainEngine')

    # ------------------------------------------------------------------ buffers
    def _alloc_buffers(self):
        B, V, C, L, dev = self.B, self.V, self.C, self.L, self.dev
        f = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
        D = int(self.model.latent_size)
        self.D = D
        self.x_in = f(self.bs, V[0], C[0])                      # un-swapped batch (all ranks hold all bs)
        self.x_stage = f(self.bs, V[0], C[0])                   # landing buffer of load_batch (template vertex order)
        self.x0 = f(B, V[0], C[0])
        self.a = [f(B, V[l + 1], C[l + 1]) for l in range(L)]   # encoder block outputs
        self.mu, self.logvar = f(B, D), f(B, D)
        self.eps, self.z = f(B, D), f(B, D)
        self.h = f(B, V[L], C[L])
        # decoder level l (L-1 .. 0): u[l] = Pool(up[l]) output, d[l] = elu(conv) output
        self.cin_de = [C[min(l + 2, L)] for l in range(L)]      # channels entering deblock at level l
        self.u = [f(B, V[l], self.cin_de[l]) for l in range(L)]
        self.d = [f(B, V[l], C[l + 1]) for l in range(L)]
        self.recon = f(B, V[0], C[0])
        # gradients
        self.drecon = f(B, V[0], C[0])
        self.qn = f(B, V[0], C[0]) if self.lap is not None else None
        self.dd = [f(B, V[l], C[l + 1]) for l in range(L)]      # d loss / d pre-activation of deblock l
        self.du = [f(B, V[l], self.cin_de[l]) for l in range(L)]
        self.dh = f(B, V[L], C[L])
        self.dz, self.dmu, self.dlv = f(B, D), f(B, D), f(B, D)
        self.dmu_kl, self.dlv_kl = f(B, D), f(B, D)
        self.da = [f(B, V[l + 1], C[l + 1]) for l in range(L)]  # d loss / d pre-activation of enblock l
        self.da_raw = f(B, V[L], C[L])
        gmax = max(V[l + 1] * self.S[l] * C[l] for l in range(1, L)) if L > 1 else 4
        self.G = f(B * gmax)
        self.wt = [f(self.cin_de[l], self.S[l] * C[l + 1]) for l in range(L)]       # decoder bwd_x weights
        self.wt_out = f(C[1], self.S[0] * C[0])
        self.wT = [f(self.S[l] * C[l], C[l + 1]) for l in range(L)]                 # encoder dense weights
        ws = 16
        for l in range(L):
            ws = max(ws, cabi.spiralconv_bwd_w_workspace(B * V[l + 1], self.S[l], C[l], C[l + 1]))
            ws = max(ws, cabi.spiralconv_bwd_w_workspace(B * V[l], self.S[l], self.cin_de[l], C[l + 1]))
        ws = max(ws, cabi.spiralconv_bwd_w_workspace(B * V[0], self.S[0], C[1], C[0]))
        self.ws = f(ws // 4 + 4)
        self._alloc_tc()
        self.losses = self.flat_g_all[self.n_arena:self.n_arena + 8]
        # two pinned read-back slots: the losses of step k can be waited for after step k+1 has been launched
        self._loss_slots = [(torch.zeros(8, dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(2)]
        self._loss_k = 0
        self.losses_host = self._loss_slots[0][0]
        self.part_mse = f(cabi.mse_lap_partial_floats(B, V[0]))
        self.part_kl = f((B * D + 255) // 256)
        nh = self.bs * (self.bs - 1) // 2 * self.bs
        self.part_lc = f(max(1, (nh + 255) // 256))
        self.act_lc = torch.empty(max(2, 2 * nh), device=dev, dtype=torch.uint8)
        self.z_all = f(self.bs * self.bs, D) if self.world > 1 else None
        self.dz_lc = f(self.bs * self.bs, D)

    # ------------------------------------------------------------------ tensor-core path
    def _alloc_tc(self):
        """Per conv layer: tile plans + packed-weight buffers for the tcgen05 kernels, where
        the layer shape is supported.  Keys: ('f', name) forward, ('b', name) backward-to-input."""
        self.tc = {}
        self._pack_table = None
        self.slot_en0 = self.slot_out = None
        self.narrow_out_ws = None
        self.narrow_out_plan = None
        self.narrow_out_tile = None
        self.narrow_out_tile_bwd = None
        self.narrow_out_tc_ws = None
        self.narrow_in_ws = None
        if not self.use_tc:
            return
        m, L, C, S = self.model, self.L, self.C, self.S
        f = lambda n: torch.empty(n, device=self.dev, dtype=torch.float32)

        def add(kind, name, layer, table, cin, cout):
            if kind == 'f':
                plan, ks, n = table.plan_fwd(), cin, cout
            else:
                plan, ks, n = table.plan_bwd(), cout, cin
            if ks not in (32, 64):
                return
            # output channels in one pass, or (64 -> 64: the weight image of all 64 does not fit in shared
            # memory next to the rings) in two passes of 32, each writing its columns of the output
            if cabi.tc_supported(table.seq, ks, n, plan.rcap):
                parts = [(0, n)]
            elif n == 64 and cabi.tc_supported(table.seq, ks, 32, plan.rcap):
                parts = [(0, 32), (32, 32)]
            else:
                return
            e = dict(layer=layer, plan=plan, cin=cin, cout=cout, seq=table.seq, tile=None,
                     parts=[(n0, nc, f(cabi.tc_wimg_floats(table.seq, ks, nc))) for n0, nc in parts])
            self.tc[(kind, name)] = e
            # 32 -> 32 layers on a patch-ordered level: tile-staged kernels (csrc/spiral_conv_tile.cuh) -- the tile's
            # distinct rows are copied once and gathered from shared memory; their weight image carries the kernel's
            # K permutation (flag bit 1 of the pack entry)
            if ks == 32 and n == 32 and len(parts) == 1 and os.environ.get('SDVAE_TILE', '1') != '0':
                tp = table.tile_fwd() if kind == 'f' else table.tile_bwd()
                if tp is not None and cabi.tile_supported(table.seq, 32, 32, tp.rcap, tp.ecap):
                    e['tile'] = tp

        for l in range(L):
            enc = m.en_layers[l].conv.layer
            add('f', 'en%d' % l, enc, self.sub[l], C[l], C[l + 1])
            if l > 0:
                add('b', 'en%d' % l, enc, self.sub[l], C[l], C[l + 1])
            dec = m.de_layers[L - l].conv.layer
            add('f', 'de%d' % l, dec, self.full[l], self.cin_de[l], C[l + 1])
            add('b', 'de%d' % l, dec, self.full[l], self.cin_de[l], C[l + 1])
        out_layer = m.de_layers[L + 1].layer
        add('f', 'out', out_layer, self.full[0], C[1], C[0])
        # Narrow layers (3 channels on one side, 32 on the other) through slot packing
        # (csrc/slot_pack.cuh): the 27 gathered columns of a vertex are materialised once, every pass of
        # the layer is then a dense 32 x 32 contraction on the tcgen05 kernels (identity plan, S = 1).
        B, V = self.B, self.V
        if S[0] * C[0] <= 32 and C[1] == 32 and V[0] < 65536 and cabi.tc_supported(1, 32, 32, 128):
            R0 = self.sub[0].n_rows
            if not cabi.narrow_in_supported(V[0], S[0], C[0], C[1]):
                self.slot_en0 = dict(plan=identity_plan(R0, self.dev), P=f(B * R0 * 32).view(B, R0, 32),
                                     Wd=f(1024), wimg=f(cabi.tc_wimg_floats(1, 32, 32)),
                                     dWd=f(1024), dbd=f(32))
            # Output layer backward (32 -> 3) fused on the FMA units (csrc/narrow_conv.cuh): G stays on the SM
            # and is never materialised
            narrow = cabi.narrow_out_bwd_supported(V[0], S[0], C[1], C[0])
            self.slot_out = dict(plan=identity_plan(V[0], self.dev),
                                 G=None if narrow else f(B * V[0] * 32).view(B, V[0], 32),
                                 Wd=f(1024), wimg=f(cabi.tc_wimg_floats(1, 32, 32)),
                                 dWd=f(1024), dbd=f(32))
            if narrow:
                self.narrow_out_ws = f(cabi.narrow_out_bwd_workspace(S[0], C[0]) // 4)
        # First encoder block (3 -> 32) on the FMA units with the mesh's input resident in shared memory
        # (csrc/narrow_conv.cuh): no slot-packed P buffer
        if cabi.narrow_in_supported(V[0], S[0], C[0], C[1]):
            self.narrow_in_ws = f(cabi.narrow_in_bwd_w_workspace(S[0], C[0]) // 4)
        sp = self.full[0].stage_plan()
        if self.full[0].n_rows == V[0] and cabi.narrow_out_fwd_supported(S[0], C[1], C[0], sp.ucap):
            self.narrow_out_plan = sp
        # ... or, on the tensor cores, by project-then-gather over the level's forward tile plan (spiral_conv_tile_out.cuh)
        if self.use_tc and self.full[0].n_rows == V[0] and C[1] == 32 and os.environ.get('SDVAE_TILE', '1') != '0':
            tp = self.full[0].tile_fwd()
            if tp is not None and cabi.narrow_out_fwd_tc_supported(S[0], C[1], C[0], tp.rcap):
                self.narrow_out_tile = tp
            # ... and its backward by gather-then-project over the inverse tile plan (spiral_conv_tile_out_bw.cuh)
            tb_ = self.full[0].tile_bwd()
            if tb_ is not None and cabi.narrow_out_bwd_tc_supported(S[0], C[1], C[0], tb_.rcap, tb_.ecap):
                self.narrow_out_tile_bwd = tb_
                self.narrow_out_tc_ws = f(cabi.narrow_out_bwd_tc_workspace(S[0], C[0]) // 4)

    def _pack_tc(self):
        """Re-pack every tensor-core weight image from the current weights (they change each step)."""
        m, L, C, S = self.model, self.L, self.C, self.S
        if self.slot_en0 is not None:
            cabi.slot_weight(m.en_layers[0].conv.layer.weight.data, self.slot_en0['Wd'], 0, C[1], S[0], C[0])
        if self.slot_out is not None and self.narrow_out_ws is None:
            cabi.slot_weight(m.de_layers[L + 1].layer.weight.data, self.slot_out['Wd'], 1, C[0], S[0], C[0])
        if self._pack_table is None:
            # one launch for every image: the weights live in the flat arena and the images are persistent,
            # so a device table of raw pointers stays valid for the life of the engine
            ents = []
            for (kind, _), e in self.tc.items():
                flags = (1 if kind == 'b' else 0) | (2 if e['tile'] is not None else 0)     # sdvae_pack_entry.transposed
                for n0, nc, wimg in e['parts']:
                    ents.append((e['layer'].weight.data, wimg, e['seq'], e['cin'], e['cout'], flags, n0, nc))
            for e in ((self.slot_en0,) if self.slot_en0 is not None else ()) + \
                     ((self.slot_out,) if self.slot_out is not None and self.narrow_out_ws is None else ()):
                ents.append((e['Wd'], e['wimg'], 1, 32, 32, False, 0, 32))
            self._pack_table = (cabi.tc_pack_table(ents, self.dev), len(ents))
        cabi.tc_pack_weights_batch(*self._pack_table)

    # ------------------------------------------------------------------ pieces
    def _conv(self, x, table, layer, out, act, B, Vin, Cin, Cout, name=None):
        e = self.tc.get(('f', name))
        if e is not None and e['tile'] is not None:
            cabi.spiralconv_fwd_tile(x, e['tile'], e['parts'][0][2], layer.bias.data, out, B, Vin, table.n_rows,
                                     table.seq, Cin, Cout, act)
            return
        if e is not None:
            for n0, nc, wimg in e['parts']:
                full = nc == Cout
                cabi.spiralconv_fwd_tc(x, e['plan'], wimg, layer.bias.data[n0:],
                                       out if full else out.view(-1)[n0:], B, Vin, table.n_rows,
                                       table.seq, Cin, nc, act, 0 if full else Cout)
            return
        cabi.spiralconv_fwd(x, table.idx, layer.weight.data, layer.bias.data, out, B, Vin,
                            table.n_rows, table.seq, Cin, Cout, act)

    def forward(self, B=None):
        """x0 -> recon, z, mu, logvar (training mode)."""
        m, L, V, C, S = self.model, self.L, self.V, self.C, self.S
        B = self.B if B is None else B
        self._pack_tc()
        x = self.x0
        for l in range(L):
            if l == 0 and self.narrow_in_ws is not None:
                lay, sub = m.en_layers[0].conv.layer, self.sub[0]
                cabi.narrow_in_fwd(x, sub.idx, lay.weight.data, lay.bias.data, self.a[0], B, V[0], sub.n_rows,
                                   S[0], C[0], C[1], cabi.ACT_ELU)
            elif l == 0 and self.slot_en0 is not None:
                e, sub = self.slot_en0, self.sub[0]
                P = e['P'][:B]
                cabi.slot_pack(x, None, sub.idx, P, B, V[0], sub.n_rows, S[0], C[0])
                cabi.dense_tc(P, e['plan'], e['wimg'], m.en_layers[0].conv.layer.bias.data, None,
                              self.a[0], B, sub.n_rows, cabi.ACT_ELU)
            else:
                self._conv(x, self.sub[l], m.en_layers[l].conv.layer, self.a[l], cabi.ACT_ELU,
                           B, V[l], C[l], C[l + 1], name='en%d' % l)
            x = self.a[l]
        flat = x.view(B, V[L] * C[L])
        lin_mu = m.en_layers[-1]
        torch.addmm(lin_mu.bias.data, flat, lin_mu.weight.data.t(), out=self.mu)
        if self.is_vae:
            lin_lv = m.en_layers[-2]
            torch.addmm(lin_lv.bias.data, flat, lin_lv.weight.data.t(), out=self.logvar)
            if self.fixed_eps is None:
                self.eps.normal_()
            # else: tests injected the oracle's noise into self.eps via set_fixed_eps()
            cabi.reparam_fwd(self.mu, self.logvar, self.eps, self.z)
            z = self.z
        else:
            if m.pre_z_sigmoid:
                torch.sigmoid(self.mu, out=self.mu)
            z = self.mu
        lin0 = m.de_layers[0]
        torch.addmm(lin0.bias.data, z, lin0.weight.data.t(), out=self.h.view(B, V[L] * C[L]))
        x = self.h
        for l in range(L - 1, -1, -1):
            up = self.up[l]
            cabi.pool_fwd(x, up, self.u[l], B, V[l + 1], self.cin_de[l])
            self._conv(self.u[l], self.full[l], m.de_layers[L - l].conv.layer, self.d[l],
                       cabi.ACT_ELU, B, V[l], self.cin_de[l], C[l + 1], name='de%d' % l)
            x = self.d[l]
        out_layer = m.de_layers[L + 1].layer
        if self.narrow_out_tile is not None:
            # 32 -> 3 on tcgen05: the tile's staged rows projected once, nine 3-vectors summed per output row
            cabi.narrow_out_fwd_tc(x, self.narrow_out_tile, out_layer.weight.data, out_layer.bias.data, self.recon,
                                   B, V[0], V[0], S[0], C[1], C[0])
        elif self.narrow_out_plan is not None:
            # 32 -> 3 on the FMA units over shared-memory-staged source rows (csrc/narrow_conv.cuh)
            cabi.narrow_out_fwd(x, self.narrow_out_plan, out_layer.weight.data, out_layer.bias.data, self.recon,
                                B, V[0], V[0], S[0], C[1], C[0])
        else:
            self._conv(x, self.full[0], out_layer, self.recon, cabi.ACT_NONE, B, V[0], C[1], C[0], name='out')
        return z

    def _bwd_w(self, x, table, dpre, layer, B, Vin, Cin, Cout):
        if self.use_tc and Cin in (32, 64) and os.environ.get('SDVAE_TILE', '1') != '0':
            tp = table.tile_fwd()
            if tp is not None and cabi.tile_bwd_w_supported(table.seq, Cin, Cout, tp.rcap):
                cabi.spiralconv_bwd_w_tile(x, tp, dpre, self.g(layer.weight), self.g(layer.bias), self.ws,
                                           B, Vin, table.n_rows, table.seq, Cin, Cout)
                return
        if self.use_tc and Cin in (32, 64):
            plan = table.plan_fwd()
            if cabi.tc_bwd_w_supported(table.seq, Cin, Cout, plan.rcap):
                cabi.spiralconv_bwd_w_tc(x, plan, dpre, self.g(layer.weight), self.g(layer.bias),
                                         self.ws, B, Vin, table.n_rows, table.seq, Cin, Cout)
                return
        cabi.spiralconv_bwd_w(x, table.idx, dpre, self.g(layer.weight), self.g(layer.bias), self.ws,
                              B, Vin, table.n_rows, table.seq, Cin, Cout)

    def _allreduce_bucket(self, k):
        if self.world == 1:
            return
        import torch.distributed as dist
        s, e = self.buckets[k]
        if k == len(self.buckets) - 1:
            e = self.n_arena + 8                        # ... and the loss scalars
        if not self.dp_overlap:
            # one all-reduce of the whole arena (+ losses) on the compute stream after the backward pass: nothing
            # competes with the persistent one-CTA-per-SM kernels for SMs, the collective's latency is exposed
            if k == len(self.buckets) - 1:
                dist.all_reduce(self.flat_g_all[0:e], op=dist.ReduceOp.SUM, group=self.pg)
            return
        self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            dist.all_reduce(self.flat_g_all[s:e], op=dist.ReduceOp.SUM, group=self.pg)

    def losses_and_backward(self, z, region: Optional[int]):
        m, L, V, C, S, cfg = self.model, self.L, self.V, self.C, self.S, self.cfg
        B = self.B
        scale = mean_loss_scale(self.world)            # local means -> global means
        lap = self.lap
        if lap is not None:
            cabi.mse_lap_fwd(self.recon, self.x0, lap.ell_col, lap.ell_val, lap.width, self.qn,
                             self.part_mse, self.losses, B, V[0], scale)
        else:
            cabi.mse_lap_fwd(self.recon, self.x0, None, None, 0, None, self.part_mse, self.losses,
                             B, V[0], scale)
        if self.is_vae and cfg.kl_weight > 0:
            cabi.kl_fwd_bwd(self.mu, self.logvar, self.dmu_kl, self.dlv_kl, self.part_kl,
                            self.losses, B, self.D, scale)
        use_lc = self.use_lc and region is not None
        if use_lc:
            r0, r1 = self.latent_regions[region]
            if self.world > 1:
                # the loss needs the whole grid's z: all-gather + loss + its gradient on a side stream -- only the
                # latent block of the backward pass (after the whole decoder backward) waits for it
                import torch.distributed as dist
                main = torch.cuda.current_stream()
                lc_stream = self.side_lc if self.dp_overlap else main
                lc_stream.wait_stream(main)
                with torch.cuda.stream(lc_stream):
                    dist.all_gather_into_tensor(self.z_all, z.contiguous(), group=self.pg)
                    cabi.lc_fwd_bwd(self.z_all, self.bs, self.D, r0, r1, cfg.latent_consistency_eta1,
                                    cfg.latent_consistency_eta2, self.act_lc, self.part_lc, self.dz_lc,
                                    self.losses)
                    self.losses[2:3].mul_(1.0 / self.world)
            else:
                cabi.lc_fwd_bwd(z, self.bs, self.D, r0, r1, cfg.latent_consistency_eta1,
                                cfg.latent_consistency_eta2, self.act_lc, self.part_lc, self.dz_lc,
                                self.losses)
        # ---- backward: decoder ------------------------------------------------------
        if lap is not None:
            cabi.mse_lap_bwd(self.recon, self.x0, self.qn, lap.t_ptr, lap.t_row, lap.t_val,
                             self.drecon, B, V[0], 1.0, cfg.laplacian_weight, scale, None)      # weight 0: no Laplacian gradient
        else:
            cabi.mse_lap_bwd(self.recon, self.x0, None, None, None, None, self.drecon, B, V[0],
                             1.0, 0.0, scale, None)
        out_layer = m.de_layers[L + 1].layer
        cp, cs = self.full[0].inverse()
        if self.narrow_out_tile_bwd is not None:
            # one fused tcgen05 pass: G = in-order cell sums of drecon, dd0 = (G W') * elu'(d0), dW = G^T d0, db
            cabi.narrow_out_bwd_tc(self.drecon, self.d[0], self.narrow_out_tile_bwd, out_layer.weight.data, self.dd[0],
                                   self.g(out_layer.weight), self.g(out_layer.bias), self.narrow_out_tc_ws,
                                   B, V[0], V[0], S[0], C[1], C[0], True)
        elif self.narrow_out_ws is not None:
            # dd0 = (G Wd^T) * elu'(d0), dW and db from one pass over drecon and d0 (G in registers)
            cabi.narrow_out_bwd(self.drecon, self.d[0], cp, cs, self.full[0].inverse_packed(), out_layer.weight.data, self.dd[0],
                                self.g(out_layer.weight), self.g(out_layer.bias), self.narrow_out_ws,
                                B, V[0], V[0], S[0], C[1], C[0], True)
        elif self.slot_out is not None:
            # G[u, s*3 + n] = sum of drecon over the vertices that gather u at slot s; then
            # dd0 = (G Wd^T) * elu'(d0)  and  dW[n, s*32 + c] = (G^T d0)[s*3 + n, c] -- both dense
            e = self.slot_out
            G = e['G'][:B]
            cabi.slot_pack(self.drecon, cp, cs, G, B, V[0], V[0], S[0], C[0])
            cabi.dense_tc(G, e['plan'], e['wimg'], None, self.d[0], self.dd[0], B, V[0], cabi.ACT_NONE)
            cabi.spiralconv_bwd_w_tc(self.d[0], e['plan'], G, e['dWd'].view(32, 32), e['dbd'], self.ws,
                                     B, V[0], V[0], 1, 32, 32)
            cabi.slot_grad(e['dWd'], e['dbd'], self.g(out_layer.weight), self.g(out_layer.bias), 1,
                           C[0], S[0], C[0])
        else:
            self._bwd_w(self.d[0], self.full[0], self.drecon, out_layer, B, V[0], C[1], C[0])
            cabi.weight_transpose(out_layer.weight.data, self.wt_out, C[0], C[1], S[0])
            cabi.spiralconv_bwd_x(self.drecon, cp, cs, self.wt_out, self.d[0], self.dd[0], B, V[0], V[0],
                                  S[0], C[0], C[1])
        for l in range(L):                               # deblock at level l, fine -> coarse
            layer = m.de_layers[L - l].conv.layer
            cin, cout = self.cin_de[l], C[l + 1]
            self._bwd_w(self.u[l], self.full[l], self.dd[l], layer, B, V[l], cin, cout)
            e = self.tc.get(('b', 'de%d' % l))
            if e is not None and e['tile'] is not None:
                cabi.spiralconv_bwd_x_tile(self.dd[l], e['tile'], e['parts'][0][2], None, self.du[l], B, V[l], V[l],
                                           S[l], cout, cin)
            elif e is not None:
                for n0, nc, wimg in e['parts']:
                    full = nc == cin
                    cabi.spiralconv_bwd_x_tc(self.dd[l], e['plan'], wimg, None,
                                             self.du[l] if full else self.du[l].view(-1)[n0:], B, V[l],
                                             V[l], S[l], cout, nc, 0 if full else cin)
            else:
                cabi.weight_transpose(layer.weight.data, self.wt[l], cout, cin, S[l])
                cp, cs = self.full[l].inverse()
                cabi.spiralconv_bwd_x(self.dd[l], cp, cs, self.wt[l], None, self.du[l], B, V[l], V[l],
                                      S[l], cout, cin)
            up = self.up[l]
            if l + 1 < L:    # gradient w.r.t. the coarser deblock's pre-activation (ELU' fused)
                cabi.csr_rowsum(self.du[l], up.t_ptr, up.t_row, up.t_val, self.d[l + 1],
                                self.dd[l + 1], B, V[l], V[l + 1], cin)
            else:
                cabi.csr_rowsum(self.du[l], up.t_ptr, up.t_row, up.t_val, None, self.dh, B, V[l],
                                V[l + 1], cin)
        self._allreduce_bucket(0)
        # ---- latent block -------------------------------------------------------------
        lin0 = m.de_layers[0]
        dh2 = self.dh.view(B, V[L] * C[L])
        zin = z
        torch.mm(dh2.t(), zin, out=self.g(lin0.weight))
        torch.sum(dh2, dim=0, out=self.g(lin0.bias))
        torch.mm(dh2, lin0.weight.data, out=self.dz)
        self._allreduce_bucket(1)
        if use_lc:
            if self.world > 1:
                if self.dp_overlap:
                    torch.cuda.current_stream().wait_stream(self.side_lc)
            lo = self.i0 * self.bs
            cabi.axpy3(self.dz, self.dz_lc[lo:lo + B], cfg.latent_consistency_weight, None, 0.0,
                       self.dz)
        flat = self.a[L - 1].view(B, V[L] * C[L])
        lin_mu = m.en_layers[-1]
        if self.is_vae:
            lin_lv = m.en_layers[-2]
            cabi.reparam_bwd(self.dz, self.logvar, self.eps, self.dmu, self.dlv, False)
            if cfg.kl_weight > 0:
                cabi.axpy3(self.dmu, self.dmu_kl, cfg.kl_weight, None, 0.0, self.dmu)
                cabi.axpy3(self.dlv, self.dlv_kl, cfg.kl_weight, None, 0.0, self.dlv)
            torch.mm(self.dmu.t(), flat, out=self.g(lin_mu.weight))
            torch.sum(self.dmu, dim=0, out=self.g(lin_mu.bias))
            torch.mm(self.dlv.t(), flat, out=self.g(lin_lv.weight))
            torch.sum(self.dlv, dim=0, out=self.g(lin_lv.bias))
            raw = self.da_raw.view(B, V[L] * C[L])
            torch.mm(self.dmu, lin_mu.weight.data, out=raw)
            raw.addmm_(self.dlv, lin_lv.weight.data)
        else:
            dmu = self.dz
            if m.pre_z_sigmoid:
                dmu = self.dz * self.mu * (1.0 - self.mu)
            torch.mm(dmu.t(), flat, out=self.g(lin_mu.weight))
            torch.sum(dmu, dim=0, out=self.g(lin_mu.bias))
            torch.mm(dmu, lin_mu.weight.data, out=self.da_raw.view(B, V[L] * C[L]))
        self._allreduce_bucket(2)
        cabi.elu_bwd(self.da_raw, self.a[L - 1], self.da[L - 1])
        # ---- backward: encoder ----------------------------------------------------------
        for l in range(L - 1, -1, -1):
            layer = m.en_layers[l].conv.layer
            x_in = self.a[l - 1] if l > 0 else self.x0
            if l == 0 and self.narrow_in_ws is not None:
                sub = self.sub[0]
                cabi.narrow_in_bwd_w(self.x0, sub.idx, self.da[0], self.g(layer.weight), self.g(layer.bias),
                                     self.narrow_in_ws, B, V[0], sub.n_rows, S[0], C[0], C[1])
            elif l == 0 and self.slot_en0 is not None:
                # dW = da0^T P with the slot-packed input of the forward pass: no gather at all
                e = self.slot_en0
                R0 = self.sub[0].n_rows
                cabi.spiralconv_bwd_w_tc(e['P'][:B], e['plan'], self.da[0], e['dWd'].view(32, 32), e['dbd'],
                                         self.ws, B, R0, R0, 1, 32, 32)
                cabi.slot_grad(e['dWd'], e['dbd'], self.g(layer.weight), self.g(layer.bias), 0,
                               C[1], S[0], C[0])
            else:
                self._bwd_w(x_in, self.sub[l], self.da[l], layer, B, V[l], C[l], C[l + 1])
            e = self.tc.get(('b', 'en%d' % l)) if l > 0 else None
            if e is not None and e['tile'] is not None:
                cabi.spiralconv_bwd_x_tile(self.da[l], e['tile'], e['parts'][0][2], self.a[l - 1], self.da[l - 1],
                                           B, self.sub[l].n_rows, V[l], S[l], C[l + 1], C[l])
            elif e is not None:
                # input gradient of the fused block straight from the kept rows (inverse table of the
                # restricted spiral table), ELU' of the previous block fused in the epilogue
                for n0, nc, wimg in e['parts']:
                    full = nc == C[l]
                    cabi.spiralconv_bwd_x_tc(self.da[l], e['plan'], wimg,
                                             self.a[l - 1] if full else self.a[l - 1].view(-1)[n0:],
                                             self.da[l - 1] if full else self.da[l - 1].view(-1)[n0:],
                                             B, self.sub[l].n_rows, V[l], S[l], C[l + 1], nc,
                                             0 if full else C[l])
            elif l > 0:
                K = S[l] * C[l]
                cabi.transpose2d(layer.weight.data, self.wT[l], C[l + 1], K)
                R = self.sub[l].n_rows
                G = self.G[:B * R * K].view(B, R * S[l], C[l])
                cabi.dense_fwd(self.da[l], self.wT[l], None, G, B * R, C[l + 1], K, C[l + 1],
                               cabi.ACT_NONE)
                ptr, src = self.sub[l].inverse_flat()
                cabi.csr_rowsum(G, ptr, src, None, self.a[l - 1], self.da[l - 1], B, R * S[l], V[l],
                                C[l])
        # the last bucket carries the loss scalars too: slots 0/1/3 hold this rank's share of the global means, slot 2
        # (the full latent-consistency loss, identical on every rank) was pre-scaled by 1/world
        self._allreduce_bucket(3)
        self._use_lc_now = use_lc

    def optimizer_step(self):
        cfg = self.cfg
        if self.side is not None and self.dp_overlap:
            torch.cuda.current_stream().wait_stream(self.side)
        cabi.total_loss(self.losses, cfg.kl_weight if self.is_vae else 0.0,
                        cfg.latent_consistency_weight if self._use_lc_now else 0.0,
                        cfg.laplacian_weight if self.lap is not None else 0.0, 0.0)
        cabi.adam_tick(self.step_dev)
        cabi.adam_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.step_dev, 0, cfg.lr,
                       cfg.betas[0], cfg.betas[1], cfg.eps, cfg.weight_decay, 1.0)

    # ------------------------------------------------------------------ public step
    def _body(self, region: Optional[int]):
        if region is not None:
            cabi.swap(self.x_in, self.masks[region], self.x0, self.bs, self.i0, self.i0 + self.rows,
                      self.V[0], self.C[0])
        z = self.forward()
        self.losses.zero_()
        self.losses_and_backward(z, region)
        self.optimizer_step()

    def load_batch(self, x_host_or_dev: torch.Tensor):
        """Hand the next un-swapped batch ``[bs, V, 3]`` (pinned host or device, template vertex order) to the engine.
        The copy runs on the engine's COPY stream into a staging buffer, so the host->device transfer of batch k+1
        overlaps step k; the next ``step`` waits for it and moves it (through the engine's internal vertex order,
        if any) into the swap kernel's input."""
        with torch.cuda.device(self.dev):
            if self._stage_free is not None:                  # the previous consumer of x_stage must be done
                self.copy_stream.wait_event(self._stage_free)
            self.copy_stream.wait_stream(torch.cuda.current_stream())      # (a device-side source may still be written)
            with torch.cuda.stream(self.copy_stream):
                self.x_stage.copy_(x_host_or_dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            self._staged = ev

    def _consume_staged(self):
        """Before the step body: staged batch -> x_in on the compute stream (a 6.5 MB device copy / gather)."""
        if self._staged is None:
            return
        main = torch.cuda.current_stream()
        main.wait_event(self._staged)
        if self.order0_dev is None:
            self.x_in.copy_(self.x_stage, non_blocking=True)
        else:                                                 # template order -> the engine's internal order
            torch.index_select(self.x_stage, 1, self.order0_dev, out=self.x_in)
        ev = torch.cuda.Event()
        ev.record(main)
        self._stage_free = ev
        self._staged = None

    def set_fixed_eps(self, eps: Optional[torch.Tensor]):
        """Parity tests: use this re-parameterisation noise instead of drawing it."""
        self.fixed_eps = eps
        if eps is not None:
            self.eps.copy_(eps)

    def prepare(self, regions):
        """Capture the step graphs for these regions up front (keeps capture out of timed loops)."""
        if not self.use_graph:
            return
        keep = [t.clone() for t in (self.flat_p, self.flat_m, self.flat_v, self.step_dev)]
        for r in regions:
            self.step(r)
        torch.cuda.synchronize(self.dev)
        for dst, src in zip((self.flat_p, self.flat_m, self.flat_v, self.step_dev), keep):
            dst.copy_(src)

    def load_local(self, x_local: torch.Tensor):
        """Copy this rank's ``B`` already-assembled meshes straight into the network input
        (use with ``step(region=None)``: no device-side swap, no latent-consistency term)."""
        if self.order0_dev is None:
            self.x0.copy_(x_local, non_blocking=True)
        else:
            if getattr(self, 'x0_raw', None) is None:
                self.x0_raw = torch.empty_like(self.x0)
            self.x0_raw.copy_(x_local, non_blocking=True)
            torch.index_select(self.x0_raw, 1, self.order0_dev, out=self.x0)

    def recon_template_order(self) -> torch.Tensor:
        """The reconstruction of the last forward pass ``[B, V, 3]`` in the template's vertex order."""
        if self.order0_dev is None:
            return self.recon
        out = torch.empty_like(self.recon)
        out.index_copy_(1, self.order0_dev, self.recon)
        return out

    def step(self, region: Optional[int], sync_losses: bool = False):
        """One training iteration on the batch previously given to ``load_batch``.
        ``region`` = index of the swapped region; ``None``: the ``B`` local meshes were given ready-made through
        ``load_local`` -- no device-side swap and no latent-consistency term."""
        self._check_arena()
        with torch.cuda.device(self.dev):
            return self._step(region, sync_losses)

    def _step(self, region: Optional[int], sync_losses: bool):
        self._consume_staged()
        if self.use_graph:
            key = -1 if region is None else int(region)
            gph = self._graphs.get(key)
            if gph is None:
                n0 = cabi.launch_count()
                # warm-up outside capture (lazy allocations, cuBLAS handles) on a scratch copy
                # of the optimiser state, so that the warm-up is not a training step
                keep = [t.clone() for t in (self.flat_p, self.flat_m, self.flat_v, self.step_dev)]
                self._body(region)
                for dst, src in zip((self.flat_p, self.flat_m, self.flat_v, self.step_dev), keep):
                    dst.copy_(src)
                torch.cuda.synchronize(self.dev)
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph):
                    self._body(region)
                self._graphs[key] = gph
                self.launches_per_step = (cabi.launch_count() - n0) // 2   # warm-up + capture
            gph.replay()
            cabi.add_launches(self.launches_per_step)
        else:
            self._body(region)
        self._loss_k ^= 1
        self.losses_host, ev = self._loss_slots[self._loss_k]
        self.losses_host.copy_(self.losses, non_blocking=True)
        ev.record(torch.cuda.current_stream())
        if sync_losses:
            ev.synchronize()
            return self.loss_dict()
        return None

    def wait_losses(self, lag: int = 0) -> Dict[str, float]:
        """Wait for the 32-byte loss read-back of the last ``step`` (``lag=0``) or of the step before it (``lag=1``:
        a training loop that logs one step late never idles the GPU between steps) and return those losses."""
        buf, ev = self._loss_slots[self._loss_k ^ (lag & 1)]
        ev.synchronize()
        v = buf.tolist()
        return {k: v[i] for i, k in enumerate(LOSS_KEYS)}

    def loss_dict(self) -> Dict[str, float]:
        v = self.losses_host.tolist()
        return {k: v[i] for i, k in enumerate(LOSS_KEYS)}

    def zero_state(self):
        self.flat_m.zero_()
        self.flat_v.zero_()
        self.step_dev.zero_()
