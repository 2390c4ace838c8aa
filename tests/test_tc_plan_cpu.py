"""Host-side tile plans of the tcgen05 path (sdvae_tc_plan_build): bit-exact index work, no GPU."""
import numpy as np
import pytest

from sdvae_b200 import cabi
from sdvae_b200.tables import inverse_cells


def _unpack(packed, rcap):
    """Undo the loader-lane packing (include/sdvae_b200.h): staged row e = 32*j + 4*t + rsub sits in word
    16*j + 4*rsub + (t >> 1), low half for even t."""
    w = packed.astype(np.uint32)
    src = np.zeros(packed.shape[:2] + (rcap,), np.int64)
    for j in range(rcap // 32):
        for rsub in range(4):
            for t in range(8):
                word = w[:, :, 16 * j + 4 * rsub + (t >> 1)]
                src[:, :, 32 * j + 4 * t + rsub] = (word >> 16) if (t & 1) else (word & 0xffff)
    return src


def _check_plan(cell_ptr, cell_src, out_rows, S):
    cnt, packed, cell, rcap = cabi.tc_plan_build(cell_ptr, cell_src, out_rows, S)
    L = (out_rows + 127) // 128
    assert cnt.shape == (L, S) and packed.shape == (L, S, rcap // 2) and cell.shape == (L, S, 128)
    assert rcap % 32 == 0 and cnt.max() <= rcap
    src = _unpack(packed, rcap)
    start = cell.astype(np.uint32) & 0xffff
    count = cell.astype(np.uint32) >> 16
    for jt in range(L):
        for s in range(S):
            n = 0
            for lr in range(128):
                r = jt * 128 + lr
                if r >= out_rows:
                    assert count[jt, s, lr] == 0
                    continue
                e0, e1 = cell_ptr[r * S + s], cell_ptr[r * S + s + 1]
                assert start[jt, s, lr] == n and count[jt, s, lr] == e1 - e0
                assert np.array_equal(src[jt, s, n:n + e1 - e0], cell_src[e0:e1])   # storage order kept
                n += e1 - e0
            assert cnt[jt, s] == n
            assert not src[jt, s, n:].any()
    return rcap


def test_forward_plan_is_the_table_itself(cranio):
    idx = cranio.spiral_tensors()[2].numpy().astype(np.int32)
    V, S = idx.shape
    rcap = _check_plan(np.arange(V * S + 1, dtype=np.int32), idx.ravel(), V, S)
    assert rcap == 128


def test_backward_plan_of_inverse_table(cranio):
    for lvl in (2, 3):
        idx = cranio.spiral_tensors()[lvl].numpy().astype(np.int32)
        V, S = idx.shape
        ptr, src = inverse_cells(idx, V)
        _check_plan(ptr, src, V, S)


def test_ragged_tail_and_empty_cells():
    rng = np.random.RandomState(0)
    out_rows, S = 131, 3                               # second tile holds 3 rows
    counts = rng.randint(0, 3, size=out_rows * S)
    ptr = np.zeros(out_rows * S + 1, np.int32)
    ptr[1:] = np.cumsum(counts)
    src = rng.randint(0, 500, size=int(ptr[-1])).astype(np.int32)
    _check_plan(ptr, src, out_rows, S)


def test_shape_support_matrix():
    assert cabi.tc_supported(9, 32, 32, 128) and cabi.tc_supported(9, 32, 32, 192)
    assert not cabi.tc_supported(9, 64, 64, 128)       # 288 KB weight image does not fit next to the rings
    assert cabi.tc_supported(9, 32, 3, 128) and cabi.tc_supported(9, 64, 32, 160)
    assert not cabi.tc_supported(9, 3, 32, 128) and not cabi.tc_supported(9, 32, 128, 128)
    assert cabi.tc_wimg_floats(9, 32, 32) == 9 * 2 * 32 * 32


def test_output_layer_tc_support_matrix():
    """Host-side shape predicates of the tcgen05 output-layer kernels (csrc/spiral_conv_tile_out*.cuh): 32 input
    channels, S*Cout <= 27, forward plans with at most 256 distinct rows per tile (two M = 128 blocks), inverse plans
    with at most 288 and a 128-byte aligned extension list; the workspace holds one partial per SM."""
    assert cabi.narrow_out_fwd_tc_supported(9, 32, 3, 256) and cabi.narrow_out_fwd_tc_supported(9, 32, 3, 224)
    assert not cabi.narrow_out_fwd_tc_supported(9, 32, 3, 288)          # a third block of staged rows
    assert not cabi.narrow_out_fwd_tc_supported(9, 64, 3, 256) and not cabi.narrow_out_fwd_tc_supported(9, 32, 4, 256)
    assert not cabi.narrow_out_fwd_tc_supported(10, 32, 3, 256) and not cabi.narrow_out_fwd_tc_supported(9, 32, 3, 250)
    assert cabi.narrow_out_bwd_tc_supported(9, 32, 3, 288, 384) and cabi.narrow_out_bwd_tc_supported(9, 32, 3, 256, 0)
    assert not cabi.narrow_out_bwd_tc_supported(9, 32, 3, 320, 384) and not cabi.narrow_out_bwd_tc_supported(9, 32, 3, 256, 100)
    assert not cabi.narrow_out_bwd_tc_supported(9, 32, 4, 256, 384) and not cabi.narrow_out_bwd_tc_supported(9, 64, 3, 256, 384)
    assert cabi.narrow_out_bwd_tc_workspace(9, 3) == 4 * 148 * (3 * 9 * 32 + 3)


def test_identity_plan_for_slot_packed_layers():
    """S = 1 identity table (row r gathers row r): the plan the dense 32 x 32 contractions of the slot-packed
    3-channel layers run with."""
    R = 300
    rcap = _check_plan(np.arange(R + 1, dtype=np.int32), np.arange(R, dtype=np.int32), R, 1)
    assert rcap == 128
    cnt, packed, cell, _ = cabi.tc_plan_build(np.arange(R + 1, dtype=np.int32), np.arange(R, dtype=np.int32), R, 1)
    assert cnt.ravel().tolist() == [128, 128, 44]
    assert np.array_equal(_unpack(packed, 128)[0, 0], np.arange(128))


def test_plan_rejects_rows_that_do_not_fit_16_bits():
    import pytest
    ptr = np.arange(5, dtype=np.int32)
    src = np.array([0, 1, 70000, 3], dtype=np.int32)      # packed plans carry 16-bit source rows
    with pytest.raises(RuntimeError):
        cabi.tc_plan_build(ptr, src, 4, 1)


def _unpack_loader_rows(src, rcap):
    """Inverse of tables.pack_rows_loader_order: [L, rcap/2] packed words -> [L, rcap] staged-row list."""
    L = src.shape[0]
    w = src.view(np.uint32).reshape(L, rcap // 32, 4, 4)          # [L, j, rsub, t>>1]
    rows = np.zeros((L, rcap // 32, 8, 4), np.int64)              # [L, j, t, rsub]
    rows[:, :, 0::2, :] = (w & 0xFFFF).transpose(0, 1, 3, 2)
    rows[:, :, 1::2, :] = (w >> 16).transpose(0, 1, 3, 2)
    return rows.reshape(L, rcap)


def _decode_tile_plan(cnt, src, cell, ext, rcap, out_rows, S):
    """Cells of a tile plan exactly as gt_kernel reads them (csrc/spiral_conv_tile.cuh): per (row, slot) the list
    of source rows, plus the number of same-parity first-row pairs (the kernel's 2-way bank conflicts)."""
    rows = _unpack_loader_rows(src, rcap)
    cells, same, pairs = [], 0, 0
    for r in range(out_rows):
        t, lr = divmod(r, 128)
        for s in range(S):
            w = int(cell[t, s * 128 + (lr >> 5) * 32 + (lr & 7) * 4 + ((lr >> 3) & 3)])
            off, c, eo = w & 0xffff, (w >> 16) & 31, w >> 21
            got = []
            if c:
                p0 = off >> 7
                assert (off & 64) == 64 * (p0 & 1) and (off & 63) == 0 and p0 < cnt[t]
                got.append(int(rows[t, p0]))
                for e in range(c - 1):
                    o2 = int(ext[t, eo + e])
                    assert (o2 & 64) == 64 * ((o2 >> 7) & 1) and (o2 >> 7) < cnt[t]
                    got.append(int(rows[t, o2 >> 7]))
            cells.append(got)
    return cells


def test_tile_plan_reproduces_the_table_forward_and_inverse():
    """tables.tile_plan (tile-staged tcgen05 kernels): decoding the plan the way the kernel does gives back every
    cell of the table, in order -- forward table (one row per cell) and inverse table (ragged cells)."""
    from sdvae_b200 import fixtures as fx, tables as tb
    tabs = fx.craniofacial_tables()
    for lvl in (3, 2):
        idx = tabs.spiral_tensors()[lvl].numpy()
        o = tb.patch_order(idx, 128)
        idx = tb.renumber_table(idx, o, o)
        R, S = idx.shape
        for ptr, src_rows in ((np.arange(R * S + 1), idx.ravel()), tb.inverse_cells(idx.astype(np.int32), R)):
            cnt, src, cell, ext, rcap, ecap = tb.tile_plan(ptr, src_rows, R, S)
            L = (R + 127) // 128
            assert cnt.shape == (L,) and src.shape == (L, rcap // 2) and cell.shape == (L, S * 128)
            assert rcap % 32 == 0 and rcap <= tb.TILE_MAX_RCAP and ecap % 64 == 0 and cell.dtype == np.uint32
            cells = _decode_tile_plan(cnt, src, cell, ext, rcap, R, S)
            for i, got in enumerate(cells):
                assert got == [int(v) for v in src_rows[ptr[i]:ptr[i + 1]]], i
            # rows of the last tile past the table: one valid row, never stored
            lr0 = R - (L - 1) * 128
            if lr0 < 128:
                w = cell[L - 1].reshape(S, 128)
                pos = np.array([(r >> 5) * 32 + (r & 7) * 4 + ((r >> 3) & 3) for r in range(lr0, 128)])
                assert np.all(w[:, pos] == np.uint32(1 << 16))


def test_tile_plan_position_parity_cuts_most_phase_pairs():
    """The two tile rows (2i, 2i+1) of a shared-memory read phase should find their staged rows at positions of
    different parity (conflict-free); the max-cut placement leaves well under half of the pairs conflicting."""
    from sdvae_b200 import fixtures as fx, tables as tb
    tabs = fx.craniofacial_tables()
    idx = tabs.spiral_tensors()[2].numpy()
    o = tb.patch_order(idx, 128)
    idx = tb.renumber_table(idx, o, o)
    R, S = idx.shape
    cnt, src, cell, ext, rcap, ecap = tb.tile_plan(np.arange(R * S + 1), idx.ravel(), R, S)
    same = total = 0
    for t in range((R + 127) // 128):
        n = min(128, R - t * 128)
        for s in range(S):
            pos = np.array([int(cell[t, s * 128 + (r >> 5) * 32 + (r & 7) * 4 + ((r >> 3) & 3)]) & 0xffff for r in range(n - (n & 1))]) >> 7
            a, b = pos[0::2], pos[1::2]
            ok = a != b
            same += int(((a & 1) == (b & 1))[ok].sum()); total += int(ok.sum())
    assert same / total < 0.35, same / total


def test_tile_plan_rejects_the_template_strip_order():
    """At level 0 the template's own numbering makes a 128-row tile read up to 547 distinct rows: no tile plan
    (callers fall back to the per-slot-gather kernels); the patch order fits."""
    from sdvae_b200 import fixtures as fx, tables as tb
    tabs = fx.craniofacial_tables()
    idx0 = tabs.spiral_tensors()[0].numpy()
    R, S = idx0.shape
    with pytest.raises(RuntimeError):
        tb.tile_plan(np.arange(R * S + 1), idx0.ravel(), R, S)
    order = tb.patch_order(idx0, 128)
    idx = tb.renumber_table(idx0, order, order)
    assert tb.tile_plan(np.arange(R * S + 1), idx.ravel(), R, S)[4] <= tb.TILE_MAX_RCAP


def test_kperm_is_a_permutation_matching_the_kernel():
    """K position kk of a 32-wide chunk holds channel kperm(kk) (csrc/spiral_conv_umma.cuh): thread q = (kk>>1)&3
    of a row owns the 16-byte pieces q and q+4."""
    kperm = lambda kk: ((kk >> 4) << 4) + (((kk >> 1) & 3) << 2) + (((kk >> 3) & 1) << 1) + (kk & 1)
    p = [kperm(k) for k in range(32)]
    assert sorted(p) == list(range(32))
    for kk in range(32):
        q, n, b = (kk >> 1) & 3, kk >> 3, kk & 1
        assert p[kk] == (16 if n >= 2 else 0) + 4 * q + 2 * (n & 1) + b
