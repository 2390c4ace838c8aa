"""Derived index tables for the kernels (host side, NumPy, bit-exact integer work).

The reference hands the network two kinds of static tables:
``indices`` -- ``LongTensor[V, S]`` spiral neighbour lists -- and ``trans`` --
uncoalesced fp32 ``torch.sparse_coo`` matrices.  The kernels want

* int32 copies (values are range-checked here, once, never at kernel time),
* for the backward-to-input pass the *inverse* spiral table in "cell" form:
  for every input vertex ``u`` and slot ``s`` the ascending list of output rows
  ``r`` with ``idx[r, s] == u`` (CSR over the ``V*S`` cells),
* ELL rows for ``Pool`` that keep each row's entries in **storage order** (the
  reference adds them in that order, model.py:53-54), plus the transposed matrix
  in CSR for the backward pass, again in storage order,
* detection of pure selection matrices (one entry of value 1.0 per row, the
  shape every quadric-collapse down-transform has, mesh_simplification.py:174-188),
  which lets an encoder block convolve only the vertices it keeps.

Derived tables are cached per source tensor (keyed on ``data_ptr``; the source
tensor is kept alive by the cache entry so the pointer cannot be recycled).
Caller tensors are never modified.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

INT32_MAX = 2 ** 31 - 1


# ---------------------------------------------------------------------------
# pure NumPy builders
# ---------------------------------------------------------------------------
def check_indices(idx: np.ndarray, n_src: Optional[int] = None) -> np.ndarray:
    """int64 ``[V,S]`` -> int32 with range checks (0 <= idx < n_src <= 2^31-1)."""
    idx = np.asarray(idx)
    if idx.ndim != 2:
        raise ValueError('spiral indices must be 2-D [V, S], got shape %s' % (idx.shape,))
    if idx.size:
        lo, hi = int(idx.min()), int(idx.max())
        if lo < 0:
            raise IndexError('negative spiral index %d' % lo)
        if n_src is not None and hi >= n_src:
            raise IndexError('spiral index %d out of range for %d vertices' % (hi, n_src))
        if hi > INT32_MAX:
            raise IndexError('spiral index %d does not fit int32' % hi)
    return np.ascontiguousarray(idx.astype(np.int32))


def inverse_cells(idx: np.ndarray, n_src: int) -> Tuple[np.ndarray, np.ndarray]:
    """Inverse of a spiral table in cell-CSR form.

    ``idx`` is ``[R, S]`` (row r gathers ``idx[r, s]`` at slot s).  Returns
    ``cell_ptr [n_src*S + 1]`` and ``cell_src [R*S]`` such that the rows ``r`` with
    ``idx[r, s] == u`` are ``cell_src[cell_ptr[u*S+s] : cell_ptr[u*S+s+1]]``, ascending."""
    idx = np.asarray(idx, np.int64)
    R, S = idx.shape
    cell = (idx * S + np.arange(S, dtype=np.int64)[None, :]).ravel()
    rows = np.repeat(np.arange(R, dtype=np.int64), S)
    order = np.lexsort((rows, cell))                      # by cell, then by row
    counts = np.bincount(cell, minlength=n_src * S)
    ptr = np.zeros(n_src * S + 1, np.int64)
    np.cumsum(counts, out=ptr[1:])
    if ptr[-1] > INT32_MAX:
        raise IndexError('inverse spiral table does not fit int32')
    return ptr.astype(np.int32), rows[order].astype(np.int32)


def pack_cells16(cell_ptr: np.ndarray, cell_src: np.ndarray, n_rows: int, width: int) -> np.ndarray:
    """First four source rows of every cell as 16-bit fields of an 8-byte word (``sdvae_narrow_out_bwd``'s
    ``cell_pack``): ``[n_cells, 2] int32``.  Field k = ELEMENT OFFSET ``row * width`` of row k of the cell inside a
    ``[n_rows, width]`` mesh; an absent row is ``n_rows * width`` (the kernel keeps a zero pad there), and
    ``0xFFFF`` in the fourth field marks a cell with more than four rows (the kernel reads rows 4.. from the
    CSR).  ``n_rows * width`` must be below ``0xFFFF``."""
    ptr = np.asarray(cell_ptr, np.int64)
    src = np.asarray(cell_src, np.int64)
    pad = int(n_rows) * int(width)
    if pad >= 0xFFFF or (src.size and src.max() >= n_rows):
        raise IndexError('pack_cells16: row offsets must fit 16 bits')
    n = ptr.size - 1
    cnt = np.diff(ptr)
    f = np.full((n, 4), pad, np.uint32)
    for k in range(4):
        m = cnt > k
        f[m, k] = src[ptr[:-1][m] + k] * width
    f[cnt > 4, 3] = 0xFFFF
    out = np.empty((n, 2), np.uint32)
    out[:, 0] = f[:, 0] | (f[:, 1] << 16)
    out[:, 1] = f[:, 2] | (f[:, 3] << 16)
    return out.view(np.int32)


def inverse_rows_flat(idx: np.ndarray, n_src: int) -> Tuple[np.ndarray, np.ndarray]:
    """Inverse table as a CSR over input vertices whose entries are flat
    ``r*S + s`` positions (ascending): ``dx[u] = sum_e G[flat_e]`` scatters the
    per-slot input gradients ``G [R*S, Cin]`` of a fused encoder block."""
    idx = np.asarray(idx, np.int64)
    R, S = idx.shape
    flat = np.arange(R * S, dtype=np.int64)
    tgt = idx.ravel()
    order = np.lexsort((flat, tgt))
    counts = np.bincount(tgt, minlength=n_src)
    ptr = np.zeros(n_src + 1, np.int64)
    np.cumsum(counts, out=ptr[1:])
    return ptr.astype(np.int32), flat[order].astype(np.int32)


def ell_from_coo(row: np.ndarray, col: np.ndarray, val: np.ndarray, n_rows: int, n_cols: int):
    """COO in storage order -> (ell_col [n_rows, W] int32 with -1 padding,
    ell_val [n_rows, W] fp32); a row's entries stay in storage order."""
    row = np.asarray(row, np.int64)
    col = np.asarray(col, np.int64)
    val = np.asarray(val, np.float32)
    if row.size and (row.min() < 0 or row.max() >= n_rows or col.min() < 0 or col.max() >= n_cols):
        raise IndexError('sparse matrix index out of range')
    counts = np.bincount(row, minlength=n_rows)
    width = int(counts.max()) if row.size else 1
    width = max(width, 1)
    order = np.argsort(row, kind='stable')                # keeps storage order inside a row
    start = np.zeros(n_rows + 1, np.int64)
    np.cumsum(counts, out=start[1:])
    pos = np.arange(row.size, dtype=np.int64) - start[row[order]]
    ell_col = np.full((n_rows, width), -1, np.int32)
    ell_val = np.zeros((n_rows, width), np.float32)
    ell_col[row[order], pos] = col[order].astype(np.int32)
    ell_val[row[order], pos] = val[order]
    return ell_col, ell_val


def transposed_csr(row: np.ndarray, col: np.ndarray, val: np.ndarray, n_cols: int):
    """CSR of the transposed matrix: for column k the (row, val) entries in storage order."""
    row = np.asarray(row, np.int64)
    col = np.asarray(col, np.int64)
    order = np.argsort(col, kind='stable')
    counts = np.bincount(col, minlength=n_cols)
    ptr = np.zeros(n_cols + 1, np.int64)
    np.cumsum(counts, out=ptr[1:])
    return (ptr.astype(np.int32), row[order].astype(np.int32),
            np.asarray(val, np.float32)[order].copy())


NARROW_FWD_TILE = 64           # output rows per tile of the staged 32 -> 3 forward (= sdvae_narrow_out_fwd_tile())
POOL_STAGE_TILE = 128          # output rows per tile of the staged Pool forward (= sdvae_pool_stage_tile())


def gather_stage_plan(cols: np.ndarray, tile: int):
    """Stage plan of a gather table ``cols [n_rows, W]`` (-1 = padding) for the shared-memory-staged kernels:
    for every tile of ``tile`` consecutive rows the ascending list of DISTINCT source rows it reads.  Returns
    ``(tile_ptr [L+1], stage_src [sum], loc [n_rows, W], ucap)`` with ``loc[r, j]`` = position of ``cols[r, j]``
    in the list of r's tile (-1 for padding) and ``ucap`` the longest list."""
    cols = np.asarray(cols, np.int32)
    n_rows = cols.shape[0]
    n_tiles = (n_rows + tile - 1) // tile
    tile_ptr = np.zeros(n_tiles + 1, np.int64)
    loc = np.full(cols.shape, -1, np.int32)
    lists = []
    for t in range(n_tiles):
        blk = cols[t * tile:(t + 1) * tile]
        valid = blk >= 0
        uniq = np.unique(blk[valid])
        lt = np.full(blk.shape, -1, np.int32)
        lt[valid] = np.searchsorted(uniq, blk[valid]).astype(np.int32)
        loc[t * tile:(t + 1) * tile] = lt
        lists.append(uniq.astype(np.int32))
        tile_ptr[t + 1] = tile_ptr[t] + uniq.size
    stage_src = np.concatenate(lists) if lists else np.zeros(0, np.int32)
    if stage_src.size == 0:
        stage_src = np.zeros(1, np.int32)
    ucap = int(np.diff(tile_ptr).max()) if n_tiles else 0
    return tile_ptr.astype(np.int32), stage_src, loc, ucap


def distinct_rows_per_row(cols: np.ndarray, tile: int) -> float:
    """Mean number of DISTINCT source rows a tile of ``tile`` consecutive rows of the gather table ``cols`` reads,
    per output row: what a shared-memory-staged kernel moves through L2 (the plain gather moves ``cols.shape[1]``)."""
    cols = np.asarray(cols)
    n = cols.shape[0]
    tot = 0
    for r0 in range(0, n, tile):
        blk = cols[r0:r0 + tile]
        tot += np.unique(blk[blk >= 0]).size
    return tot / max(n, 1)


def patch_order(idx: np.ndarray, tile: int = 128) -> np.ndarray:
    """A vertex order in which every run of ``tile`` consecutive vertices is a compact patch of the mesh, from the
    spiral table alone (no geometry): greedy region growing -- breadth-first balls among the unassigned vertices,
    each seeded at the unassigned vertex with the most assigned neighbours (fills holes, keeps patches adjacent).
    Returns ``order`` (new position -> old vertex).

    Why: the template numbers its vertices in thin strips, so a 128-row tile of the level-0 spiral table reads 412
    distinct rows (3.2 per output row); in patch order it reads ~200 (1.6), the up-sampling Pool 0.44 instead of
    0.91, and the orders of the coarser levels follow by rank of the kept vertices.  All tables of the network
    are static and only the input, the output and the latent level are visible outside the engine, so the
    internal levels can be renumbered freely -- the lever that makes tile-local staging pay for the tensor-core
    gather kernels.  Host-side groundwork: not applied by the kernels of this round."""
    idx = np.asarray(idx, np.int64)
    V = idx.shape[0]
    nb = [set() for _ in range(V)]
    for v in range(V):
        for u in idx[v, 1:]:
            u = int(u)
            if u != v:
                nb[v].add(u)
                nb[u].add(v)
    assigned = np.zeros(V, bool)
    touched = np.zeros(V, np.int64)            # assigned neighbours of an unassigned vertex
    cand = set()                               # unassigned vertices next to an assigned one
    order = []

    def next_seed(exclude):
        pool = [v for v in cand if v not in exclude]
        if pool:
            return max(pool, key=lambda v: (touched[v], -v))
        rest = [int(v) for v in np.flatnonzero(~assigned) if int(v) not in exclude]
        return rest[0]

    while len(order) < V:
        run, seen = [], set()
        frontier = []
        while len(run) < tile and len(order) + len(run) < V:
            if not frontier:                   # first ball of the run, or the ball ran out of free vertices
                s = next_seed(seen)
                frontier = [s]
                seen.add(s)
            nxt = []
            for v in frontier:
                if len(run) >= tile:
                    break
                run.append(v)
                for u in sorted(nb[v]):
                    if not assigned[u] and u not in seen:
                        seen.add(u)
                        nxt.append(u)
            frontier = nxt
        for v in run:
            assigned[v] = True
            cand.discard(v)
        for v in run:
            for u in nb[v]:
                if not assigned[u]:
                    touched[u] += 1
                    cand.add(u)
        order.extend(run)
    return np.asarray(order, np.int64)


def renumber_table(cols: np.ndarray, row_order: np.ndarray, src_order: np.ndarray) -> np.ndarray:
    """Gather table ``cols [n_rows, W]`` (-1 = padding) with rows listed in ``row_order`` and source rows renamed by
    ``src_order`` (both: new position -> old index)."""
    cols = np.asarray(cols, np.int64)
    rank = np.empty(len(src_order), np.int64)
    rank[np.asarray(src_order, np.int64)] = np.arange(len(src_order))
    out = cols[np.asarray(row_order, np.int64)]
    return np.where(out >= 0, rank[np.clip(out, 0, None)], -1)


def pack_rows_loader_order(rows: np.ndarray) -> np.ndarray:
    """``rows [L, rcap]`` (rcap a multiple of 32, values < 65536) -> ``[L, rcap/2] int32``: 16-bit pairs in the
    order the tcgen05 loader lanes consume them (``umma::plan_fetch`` / ``sdvae_tc_plan_build``): staged row
    ``e = 32*j + 4*t + rsub`` sits in word ``16*j + 4*rsub + (t >> 1)``, low half for even ``t``."""
    rows = np.asarray(rows, np.int64)
    L, rcap = rows.shape
    if rcap % 32 or (rows.size and (rows.min() < 0 or rows.max() >= 65536)):
        raise ValueError('pack_rows_loader_order: rcap must be a multiple of 32 and rows must fit 16 bits')
    r = rows.reshape(L, rcap // 32, 8, 4).astype(np.uint32)          # [L, j, t, rsub]
    lo, hi = r[:, :, 0::2, :], r[:, :, 1::2, :]                      # [L, j, t>>1, rsub]
    w = (lo | (hi << 16)).transpose(0, 1, 3, 2)                      # [L, j, rsub, t>>1]
    return np.ascontiguousarray(w.reshape(L, rcap // 2)).view(np.int32)


TILE_MAX_RCAP = 288            # distinct rows per tile the tile-staged tcgen05 kernels take (csrc/spiral_conv_tile.cuh)


def _two_colour(n: int, a: np.ndarray, b: np.ndarray, sweeps: int = 12, seed: int = 0) -> np.ndarray:
    """Local-search max-cut: colours in {0, 1} for ``n`` vertices so that as many pairs ``(a[i], b[i])`` as possible
    get different colours (a triangulated surface is far from bipartite: ~75 % of the pairs end up cut)."""
    colour = np.zeros(n, np.int64)
    if n == 0 or a.size == 0:
        return colour
    keep = a != b
    a, b = a[keep], b[keep]
    order = np.argsort(np.concatenate([a, b]), kind='stable')
    nbr = np.concatenate([b, a])[order]
    ptr = np.concatenate([[0], np.cumsum(np.bincount(np.concatenate([a, b]), minlength=n))])
    rng = np.random.RandomState(seed)
    colour = rng.randint(0, 2, n).astype(np.int64)
    for _ in range(sweeps):
        changed = 0
        for v in rng.permutation(n):
            nb = nbr[ptr[v]:ptr[v + 1]]
            if nb.size and 2 * int((colour[nb] == colour[v]).sum()) > nb.size:
                colour[v] ^= 1
                changed += 1
        if not changed:
            break
    return colour


def tile_plan(cell_ptr: np.ndarray, cell_src: np.ndarray, out_rows: int, seq: int, max_rcap: int = TILE_MAX_RCAP):
    """Tile plan of a cell-form table for the tile-staged tcgen05 kernels (``sdvae_spiralconv_fwd_tile`` /
    ``sdvae_spiralconv_bwd_x_tile``, layout in include/sdvae_b200.h): cell ``(r, s)`` = rows
    ``cell_src[cell_ptr[r*seq+s] : cell_ptr[r*seq+s+1]]`` (in that order -- the kernel sums them in order).
    Returns ``(cnt [L], src [L, rcap/2], cell [L, seq*128] uint32, ext [L, ecap] uint16, rcap, ecap)``;
    ``ecap == 0`` for forward tables (one row per cell).  Raises ``RuntimeError`` when a tile reads more than
    ``max_rcap`` distinct rows (number the level patch-wise first: ``patch_order``).

    Position parity: the kernel's two 64-byte halves of a staged row sit in the order (low, high) at even
    positions and (high, low) at odd ones, and the 8 lanes of a shared-memory read phase serve two consecutive tile
    rows: the read is conflict-free when their two staged rows sit at positions of different parity.  The
    positions are therefore chosen by a max-cut over the pairs (first rows of the cells of tile rows 2i, 2i+1)."""
    cell_ptr = np.asarray(cell_ptr, np.int64)
    cell_src = np.asarray(cell_src, np.int64)
    R, S = int(out_rows), int(seq)
    if cell_ptr.size != R * S + 1:
        raise ValueError('tile_plan: cell_ptr must have out_rows*seq + 1 entries')
    L = (R + 127) // 128
    counts = np.diff(cell_ptr)                                   # [R*S]
    if counts.size and counts.max() > 31:
        raise RuntimeError('tile_plan: a cell holds more than 31 rows')
    per_tile = []
    ucap = 0
    for t in range(L):
        e0, e1 = cell_ptr[t * 128 * S], cell_ptr[min(R, (t + 1) * 128) * S]
        uniq, inv = np.unique(cell_src[e0:e1], return_inverse=True)
        per_tile.append((uniq, inv.astype(np.int64)))
        ucap = max(ucap, uniq.size)
    rcap = max(32, (ucap + 31) // 32 * 32)
    if rcap > max_rcap:
        raise RuntimeError('tile_plan: a tile reads %d distinct rows (limit %d)' % (ucap, max_rcap))
    rows = np.zeros((L, rcap), np.int64)
    cnts = np.zeros(L, np.int32)
    cell = np.zeros((L, S, 128), np.uint32)
    cell[:, :, :] = np.uint32(1 << 16)                           # rows past the table: one (valid) row, never stored
    ext_l = []
    r_loc = np.arange(128)
    word_pos = (r_loc >> 5) * 32 + (r_loc & 7) * 4 + ((r_loc >> 3) & 3)      # position of tile row r inside a slot
    half = rcap // 2
    for t in range(L):
        uniq, loc = per_tile[t]
        r0, r1 = t * 128, min(R, (t + 1) * 128)
        n = r1 - r0
        c = counts[r0 * S:r1 * S]                                # [n*S], cell (r, s) at r*S + s
        start = cell_ptr[r0 * S:r1 * S] - cell_ptr[r0 * S]
        first = np.where(c > 0, loc[np.minimum(start, max(loc.size - 1, 0))] if loc.size else 0, -1).reshape(n, S)
        # pairs of staged rows read by one shared-memory phase: tile rows (2i, 2i+1), same slot
        ne = n - (n & 1)
        pa, pb = first[0:ne:2].ravel(), first[1:ne:2].ravel()
        ok = (pa >= 0) & (pb >= 0)
        colour = _two_colour(uniq.size, pa[ok], pb[ok], seed=t)
        # positions: colour 0 -> even, colour 1 -> odd; the overflow of the larger class takes what is left
        pos = np.full(uniq.size, -1, np.int64)
        free = [list(range(0, rcap, 2)), list(range(1, rcap, 2))]
        for col in (0, 1):
            ids = np.flatnonzero(colour == col)
            k = min(ids.size, half)
            pos[ids[:k]] = free[col][:k]
            free[col] = free[col][k:]
        rest = np.flatnonzero(pos < 0)
        spare = sorted(free[0] + free[1])
        pos[rest] = spare[:rest.size]
        rows[t, pos] = uniq
        cnts[t] = int(pos.max()) + 1 if pos.size else 0
        off = (pos << 7) | ((pos & 1) << 6)                      # byte offset of the LOW half of the staged row
        floc = np.where(c > 0, off[np.maximum(first.ravel(), 0)] if off.size else 0, 0)
        extra = np.maximum(c - 1, 0)
        eoff = np.concatenate([[0], np.cumsum(extra)[:-1]]) if c.size else np.zeros(0, np.int64)
        if extra.sum():
            sel = np.ones(loc.size, bool)
            sel[start[c > 0]] = False                            # drop the first row of every non-empty cell
            ext_l.append(off[loc[sel]].astype(np.uint16))
        else:
            ext_l.append(np.zeros(0, np.uint16))
        w = (floc.astype(np.uint64) | (c.astype(np.uint64) << 16) | (eoff.astype(np.uint64) << 21))
        if w.size and w.max() >= (1 << 32):
            raise RuntimeError('tile_plan: cell word overflow')
        w = w.astype(np.uint32).reshape(n, S)
        cell[t][:, word_pos[:n]] = w.T
    ecap = max((e.size for e in ext_l), default=0)
    ecap = (ecap + 63) // 64 * 64                               # keeps every tile stage 128-byte aligned
    ext = np.zeros((L, max(ecap, 8)), np.uint16)
    for t, e in enumerate(ext_l):
        ext[t, :e.size] = e
    return (cnts, pack_rows_loader_order(rows), cell.reshape(L, S * 128), ext, rcap, ecap)


def renumber_levels(spirals, downs, ups, tile: int = 128):
    """Patch-wise renumbering of every level but the coarsest (``patch_order``), applied consistently to the spiral
    tables ``spirals[l] [V_l, S]`` and to the transforms ``downs[l]`` (``[V_{l+1}, V_l]``) / ``ups[l]``
    (``[V_l, V_{l+1}]``) given as ``(row, col, val, shape)`` in storage order.  Entries keep their storage order
    (so every per-vertex sum keeps its terms and their order); the coarsest level keeps its numbering (the latent
    ``Linear`` layers flatten it, model.py:151,166).  Returns ``(spirals', downs', ups', orders)`` with
    ``orders[l]`` = new position -> old vertex."""
    nlev = len(spirals)
    n_last = int(downs[-1][3][0])
    orders = [patch_order(np.asarray(s, np.int64), tile) for s in spirals] + [np.arange(n_last, dtype=np.int64)]
    ranks = []
    for o in orders:
        r = np.empty(o.size, np.int64)
        r[o] = np.arange(o.size)
        ranks.append(r)
    sp = [ranks[l][np.asarray(spirals[l], np.int64)[orders[l]]] for l in range(nlev)]
    dn = [(ranks[l + 1][np.asarray(r, np.int64)], ranks[l][np.asarray(c, np.int64)], v, shape)
          for l, (r, c, v, shape) in enumerate(downs)]
    up = [(ranks[l][np.asarray(r, np.int64)], ranks[l + 1][np.asarray(c, np.int64)], v, shape)
          for l, (r, c, v, shape) in enumerate(ups)]
    return sp, dn, up, orders


def renumbered_model_tables(spiral_indices, down_transform, up_transform, tile: int = 128):
    """``renumber_levels`` on a model's own tensors (``LongTensor[V,S]`` spirals, sparse COO transforms on any
    device): returns new tensors of the same kinds on the same device, and ``orders`` (NumPy)."""
    dev = spiral_indices[0].device

    def coo(t):
        ind = t._indices().detach().cpu().numpy()
        return ind[0], ind[1], t._values().detach().cpu().numpy().astype(np.float32), tuple(int(x) for x in t.shape)

    sp, dn, up, orders = renumber_levels([s.detach().cpu().numpy() for s in spiral_indices],
                                         [coo(t) for t in down_transform], [coo(t) for t in up_transform], tile)

    def sparse(e):
        r, c, v, shape = e
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            t = torch.sparse_coo_tensor(torch.from_numpy(np.stack([r, c]).astype(np.int64)),
                                        torch.from_numpy(np.asarray(v, np.float32)), torch.Size(shape))
        return t.to(dev)

    return ([torch.from_numpy(np.ascontiguousarray(s)).long().to(dev) for s in sp],
            [sparse(e) for e in dn], [sparse(e) for e in up], orders)


def pool_stage_plan(ell_col: np.ndarray, ell_val: np.ndarray, tile: int = POOL_STAGE_TILE):
    """Stage plan of the shared-memory Pool forward (include/sdvae_b200.h, ``sdvae_pool_ell_fwd_staged``):
    ``gather_stage_plan`` of the ELL columns, with each entry's position and value interleaved:
    ``(tile_ptr [L+1], stage_src [sum], ent [n_rows, W, 2], ucap)``,
    ``ent[r, j] = (position of ell_col[r, j] in the list of r's tile, or -1 for padding; bits of ell_val[r, j])``.
    Entry order inside a row is untouched (the reference adds in storage order, model.py:53-54)."""
    ell_val = np.asarray(ell_val, np.float32)
    tile_ptr, stage_src, loc, ucap = gather_stage_plan(ell_col, tile)
    ent = np.empty(loc.shape + (2,), np.int32)
    ent[:, :, 0] = loc
    ent[:, :, 1] = ell_val.view(np.int32)
    return tile_ptr, stage_src, ent, ucap


def selection_columns(row, col, val, n_rows) -> Optional[np.ndarray]:
    """If the matrix is a pure row selection (exactly one entry per row, value 1.0)
    return ``kept[r] = col of row r``; otherwise ``None``."""
    row = np.asarray(row, np.int64)
    val = np.asarray(val, np.float32)
    if row.size != n_rows or not np.all(val == np.float32(1.0)):
        return None
    if not np.array_equal(np.sort(row), np.arange(n_rows)):
        return None
    kept = np.empty(n_rows, np.int64)
    kept[row] = np.asarray(col, np.int64)
    return kept


# ---------------------------------------------------------------------------
# device-side bundles
# ---------------------------------------------------------------------------
def _dev(a: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


@dataclass
class TilePlan:
    """Tile plan of a cell-form table for the tcgen05 kernels (include/sdvae_b200.h,
    ``sdvae_tc_plan_build``): per tile of 128 output rows and spiral slot, the source rows
    to stage and, per output row, the range of staged rows summed into its cell."""
    cnt: torch.Tensor            # int32 [L, S]
    src: torch.Tensor            # int32 [L, S, rcap/2]  two 16-bit source rows per word, loader-lane order
    cell: torch.Tensor           # int32 [L, S, 128]   start | count << 16
    rcap: int
    out_rows: int
    seq: int

    @staticmethod
    def try_build(cell_ptr, cell_src, out_rows: int, seq: int, device) -> "TilePlan":
        """``build``, or the ``NO_PLAN`` marker (empty arrays, ``rcap == 0``) when the table is outside what
        the tensor-core kernels support."""
        try:
            return TilePlan.build(cell_ptr, cell_src, out_rows, seq, device)
        except RuntimeError:
            z = torch.zeros(1, dtype=torch.int32, device=device)
            return TilePlan(z, z, z, 0, int(out_rows), int(seq))

    @staticmethod
    def build(cell_ptr: np.ndarray, cell_src: np.ndarray, out_rows: int, seq: int, device) -> "TilePlan":
        from . import cabi
        cnt, src, cell, rcap = cabi.tc_plan_build(cell_ptr, cell_src, out_rows, seq)
        return TilePlan(_dev(cnt, device), _dev(src, device), _dev(cell, device), int(rcap),
                        int(out_rows), int(seq))


@dataclass
class SpiralTable:
    """Device copies of one (possibly row-restricted) spiral table."""
    idx: torch.Tensor            # int32 [R, S]
    n_rows: int
    n_src: int
    seq: int
    _np_idx: np.ndarray
    _inv: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
    _inv_flat: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
    _inv_pack: Optional[dict] = None
    _stage: Optional[dict] = None
    _plan_fwd: Optional["TilePlan"] = None
    _plan_bwd: Optional["TilePlan"] = None
    _tile_fwd: Optional[tuple] = None
    _tile_bwd: Optional[tuple] = None

    @staticmethod
    def build(idx_np: np.ndarray, n_src: int, device) -> "SpiralTable":
        i32 = check_indices(idx_np, n_src)
        return SpiralTable(_dev(i32, device), int(i32.shape[0]), int(n_src), int(i32.shape[1]), i32)

    def inverse(self):
        """(cell_ptr, cell_src) on the device, built on first use."""
        if self._inv is None:
            ptr, src = inverse_cells(self._np_idx, self.n_src)
            self._inv = (_dev(ptr, self.idx.device), _dev(src, self.idx.device))
        return self._inv

    def inverse_packed(self, width: int = 3) -> torch.Tensor:
        """``pack_cells16`` of the inverse table for ``width``-channel rows, on the device."""
        if self._inv_pack is None:
            self._inv_pack = {}
        if width not in self._inv_pack:
            ptr, src = inverse_cells(self._np_idx, self.n_src)
            self._inv_pack[width] = _dev(pack_cells16(ptr, src, self.n_rows, width), self.idx.device)
        return self._inv_pack[width]

    def stage_plan(self, tile: int = NARROW_FWD_TILE) -> GatherStagePlan:
        """``gather_stage_plan`` of this table (forward gather), on the device."""
        if self._stage is None:
            self._stage = {}
        if tile not in self._stage:
            tp, ss, loc, ucap = gather_stage_plan(self._np_idx, tile)
            dev = self.idx.device
            self._stage[tile] = GatherStagePlan(_dev(tp, dev), _dev(ss, dev), _dev(loc, dev), int(tile), ucap)
        return self._stage[tile]

    def inverse_flat(self):
        if self._inv_flat is None:
            ptr, src = inverse_rows_flat(self._np_idx, self.n_src)
            self._inv_flat = (_dev(ptr, self.idx.device), _dev(src, self.idx.device))
        return self._inv_flat

    def plan_fwd(self) -> "TilePlan":
        """Tile plan of the forward gather (one source row per cell: cell_ptr[i] = i).  A table the
        tensor-core kernels cannot take (>= 65536 source rows: plans carry 16-bit rows) gets the
        ``NO_PLAN`` marker (``rcap == 0``); every ``*_supported`` query rejects it, so callers fall
        back to the fp32-FMA kernels."""
        if self._plan_fwd is None:
            n = self.n_rows * self.seq
            self._plan_fwd = TilePlan.try_build(np.arange(n + 1, dtype=np.int32), self._np_idx.ravel(),
                                                self.n_rows, self.seq, self.idx.device)
        return self._plan_fwd

    def plan_bwd(self) -> "TilePlan":
        """Tile plan of the backward-to-input gather (inverse table, rows = source vertices); ``NO_PLAN``
        marker as for ``plan_fwd`` (also when one (tile, slot) cell group stages more rows than the kernels'
        ring stage holds)."""
        if self._plan_bwd is None:
            ptr, src = inverse_cells(self._np_idx, self.n_src)
            self._plan_bwd = TilePlan.try_build(ptr, src, self.n_src, self.seq, self.idx.device)
        return self._plan_bwd

    def tile_fwd(self) -> Optional["TileStagePlan"]:
        """Tile-staged plan of the forward gather, or ``None`` when a 128-row tile reads too many distinct rows
        (template strip order: renumber the level patch-wise, ``patch_order``)."""
        if self._tile_fwd is None:
            n = self.n_rows * self.seq
            self._tile_fwd = (TileStagePlan.try_build(np.arange(n + 1, dtype=np.int64), self._np_idx.ravel(),
                                                      self.n_rows, self.seq, self.idx.device),)
        return self._tile_fwd[0]

    def tile_bwd(self) -> Optional["TileStagePlan"]:
        """Tile-staged plan of the backward-to-input gather (inverse table), or ``None``."""
        if self._tile_bwd is None:
            ptr, src = inverse_cells(self._np_idx, self.n_src)
            self._tile_bwd = (TileStagePlan.try_build(ptr, src, self.n_src, self.seq, self.idx.device),)
        return self._tile_bwd[0]

    def restrict(self, kept: np.ndarray) -> "SpiralTable":
        """Table of the rows in ``kept`` only (fused conv + selection pooling)."""
        return SpiralTable.build(self._np_idx[np.asarray(kept, np.int64)], self.n_src,
                                 self.idx.device)


@dataclass
class TileStagePlan:
    """Device copy of ``tile_plan`` (tile-staged tcgen05 kernels, csrc/spiral_conv_tile.cuh)."""
    cnt: torch.Tensor            # int32 [L]
    src: torch.Tensor            # int32 [L, rcap/2]
    cell: torch.Tensor           # int32 view of uint32 [L, S*128]
    ext: torch.Tensor            # int16 view of uint16 [L, max(ecap, 8)]
    rcap: int
    ecap: int
    out_rows: int
    seq: int

    @staticmethod
    def build(cell_ptr, cell_src, out_rows: int, seq: int, device) -> "TileStagePlan":
        cnt, src, cell, ext, rcap, ecap = tile_plan(cell_ptr, cell_src, out_rows, seq)
        return TileStagePlan(_dev(cnt, device), _dev(src, device), _dev(cell.view(np.int32), device),
                             _dev(ext.view(np.int16), device), int(rcap), int(ecap), int(out_rows), int(seq))

    @staticmethod
    def try_build(cell_ptr, cell_src, out_rows: int, seq: int, device) -> Optional["TileStagePlan"]:
        """``build``, or ``None`` when a tile reads more distinct rows than a tile stage holds."""
        try:
            return TileStagePlan.build(cell_ptr, cell_src, out_rows, seq, device)
        except RuntimeError:
            return None


@dataclass
class GatherStagePlan:
    """Device copy of ``gather_stage_plan`` of a spiral table."""
    tile_ptr: torch.Tensor       # int32 [L + 1]
    stage_src: torch.Tensor      # int32 [tile_ptr[L]]
    loc: torch.Tensor            # int32 [n_rows, S]
    T: int
    ucap: int


@dataclass
class PoolStagePlan:
    """Device copy of ``pool_stage_plan``."""
    tile_ptr: torch.Tensor       # int32 [L + 1]
    stage_src: torch.Tensor      # int32 [tile_ptr[L]]
    ent: torch.Tensor            # int32 [n_rows, W, 2]
    T: int
    ucap: int


@dataclass
class PoolTable:
    """Device copies of one sparse transform: ELL forward rows, transposed CSR."""
    n_rows: int
    n_cols: int
    width: int
    ell_col: torch.Tensor        # int32 [n_rows, W]
    ell_val: torch.Tensor        # fp32  [n_rows, W]
    t_ptr: torch.Tensor          # int32 [n_cols + 1]
    t_row: torch.Tensor          # int32 [nnz]
    t_val: torch.Tensor          # fp32  [nnz]
    kept: Optional[np.ndarray]   # selection columns or None
    _stage: Optional[object] = None

    def stage_plan(self) -> Optional[PoolStagePlan]:
        """Stage plan of the shared-memory forward, built on first use; ``None`` when staging cannot pay
        (a tile reads at least as many distinct source rows as it has entries, e.g. selection matrices)."""
        if self._stage is None:
            ec = self.ell_col.cpu().numpy()
            tp, ss, ent, ucap = pool_stage_plan(ec, self.ell_val.cpu().numpy())
            entries = int((ec >= 0).sum())
            if ucap == 0 or int(tp[-1]) * 2 > entries:
                self._stage = False
            else:
                dev = self.ell_col.device
                self._stage = PoolStagePlan(_dev(tp, dev), _dev(ss, dev), _dev(ent, dev),
                                            POOL_STAGE_TILE, ucap)
        return self._stage or None

    @staticmethod
    def build(row, col, val, shape, device) -> "PoolTable":
        n_rows, n_cols = int(shape[0]), int(shape[1])
        ec, ev = ell_from_coo(row, col, val, n_rows, n_cols)
        tp, tr, tv = transposed_csr(row, col, val, n_cols)
        return PoolTable(n_rows, n_cols, int(ec.shape[1]), _dev(ec, device), _dev(ev, device),
                         _dev(tp, device), _dev(tr, device), _dev(tv, device),
                         selection_columns(row, col, val, n_rows))


# ---------------------------------------------------------------------------
# caches keyed on the caller's tensors
# ---------------------------------------------------------------------------
_spiral_cache: Dict[tuple, tuple] = {}
_pool_cache: Dict[tuple, tuple] = {}
_restricted_cache: Dict[tuple, SpiralTable] = {}


_CACHE_LIMIT = 64        # derived tables of at most this many caller tensors stay cached (oldest evicted first)


def _key(t: torch.Tensor):
    # _version: an in-place edit of the caller's tensor after first use must not hit stale derived tables
    return (t.data_ptr(), tuple(t.shape), str(t.device), t.dtype, t._version)


def _remember(cache: dict, k, value):
    while len(cache) >= _CACHE_LIMIT:
        cache.pop(next(iter(cache)))
    cache[k] = value


def spiral_table(indices: torch.Tensor, n_src: Optional[int] = None) -> SpiralTable:
    """``SpiralTable`` for a caller-owned ``LongTensor[V,S]`` (model.py:12-17).
    ``n_src`` defaults to V (every use in the reference's Model has V_out == V_in)."""
    k = _key(indices) + (n_src,)
    hit = _spiral_cache.get(k)
    if hit is not None:
        return hit[1]
    if indices.dim() != 2:
        raise ValueError('indices must be [V, S]')
    n_src_eff = int(indices.shape[0]) if n_src is None else int(n_src)
    tab = SpiralTable.build(indices.detach().cpu().numpy(), n_src_eff, indices.device)
    _remember(_spiral_cache, k, (indices, tab))
    return tab


def pool_table(trans: torch.Tensor) -> PoolTable:
    """``PoolTable`` for a caller-owned sparse COO matrix (model.py:50-52); only
    ``_indices()``, ``_values()`` and the shape are read, as in the reference."""
    ind, val = trans._indices(), trans._values()
    k = (ind.data_ptr(), val.data_ptr(), tuple(trans.shape), str(ind.device), ind._version, val._version)
    hit = _pool_cache.get(k)
    if hit is not None:
        return hit[1]
    ind_np = ind.detach().cpu().numpy()
    tab = PoolTable.build(ind_np[0], ind_np[1], val.detach().cpu().numpy().astype(np.float32),
                          (trans.size(0), trans.size(1)), ind.device)
    _remember(_pool_cache, k, ((ind, val), tab))
    return tab


_identity_plans: Dict[Tuple[int, str], "TilePlan"] = {}


def identity_plan(n_rows: int, device) -> "TilePlan":
    """Forward tile plan of the identity table (S = 1, row r gathers row r): the dense 32-wide
    contractions of the slot-packed narrow layers run on the tcgen05 kernels with it."""
    k = (int(n_rows), str(device))
    hit = _identity_plans.get(k)
    if hit is None:
        hit = TilePlan.build(np.arange(n_rows + 1, dtype=np.int32), np.arange(n_rows, dtype=np.int32),
                             int(n_rows), 1, device)
        _identity_plans[k] = hit
    return hit


def restricted_spiral_table(indices: torch.Tensor, pool: PoolTable) -> Optional[SpiralTable]:
    """Spiral table restricted to the rows a selection down-transform keeps, or None
    if ``pool`` is not a pure selection."""
    if pool.kept is None or int(indices.shape[0]) != pool.n_cols:
        return None
    k = _key(indices) + (id(pool),)
    hit = _restricted_cache.get(k)
    if hit is None or hit[0] is not pool:          # (id() of a collected PoolTable can be reused: keep and compare the object)
        hit = (pool, spiral_table(indices).restrict(pool.kept))
        _remember(_restricted_cache, k, hit)
    return hit[1]


def clear_caches():
    """Drop every cached derived table (they hold device memory): call after replacing a model's index tensors."""
    _spiral_cache.clear()
    _pool_cache.clear()
    _restricted_cache.clear()
