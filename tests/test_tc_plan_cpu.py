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


def test_staged_tile_plan_matches_the_c_packer_and_the_table():
    """Plan of the EXPERIMENTAL staged tcgen05 forward: the numpy packer reproduces sdvae_tc_plan_build's
    loader-lane order bit for bit, and (cnt, src, loc) re-address every gather of the table."""
    from sdvae_b200 import cabi, fixtures as fx, tables as tb
    tabs = fx.craniofacial_tables()
    idx = tabs.spiral_tensors()[3].numpy()                        # 267 rows
    o3 = tb.patch_order(idx, 128)
    idx = tb.renumber_table(idx, o3, o3)                          # in patch order a tile reads <= 192 distinct rows
    R, S = idx.shape
    cnt, src, loc, rcap = tb.staged_tile_plan(idx)
    L = (R + 127) // 128
    assert cnt.shape == (L,) and src.shape == (L, rcap // 2) and loc.shape == (L, S, 128) and rcap % 32 == 0
    # unpack the loader order again
    w = src.view(np.uint32).reshape(L, rcap // 32, 4, 4)          # [L, j, rsub, t>>1]
    rows = np.zeros((L, rcap // 32, 8, 4), np.int64)              # [L, j, t, rsub]
    rows[:, :, 0::2, :] = (w & 0xFFFF).transpose(0, 1, 3, 2)
    rows[:, :, 1::2, :] = (w >> 16).transpose(0, 1, 3, 2)
    rows = rows.reshape(L, rcap)
    for t in range(L):
        lst = rows[t, :cnt[t]]
        assert np.all(np.diff(lst) > 0) and not rows[t, cnt[t]:].any()
        blk = idx[t * 128:(t + 1) * 128]
        assert np.array_equal(lst[loc[t, :, :blk.shape[0]].T], blk)
        assert not loc[t, :, blk.shape[0]:].any()
    # the same lists through the C builder: one pseudo-row per tile whose only cell holds the tile's rows
    assert rcap <= 192
    if True:
        ptr = np.zeros(L * 128 + 1, np.int32)
        starts = np.concatenate([[0], np.cumsum(cnt)])
        for t in range(L):
            ptr[t * 128 + 1:(t + 1) * 128 + 1] = starts[t + 1]     # row 0 of tile t owns all its rows
        flat = np.concatenate([rows[t, :cnt[t]] for t in range(L)]).astype(np.int32)
        c_cnt, c_src, _, c_rcap = cabi.tc_plan_build(ptr, flat, L * 128, 1)
        assert c_rcap == rcap
        assert np.array_equal(c_cnt.ravel(), cnt) and np.array_equal(c_src.reshape(L, -1), src)
    # the template's strip order does not fit at level 0; patch order does
    idx0 = tabs.spiral_tensors()[0].numpy()
    with pytest.raises(RuntimeError):
        tb.staged_tile_plan(idx0)
    order = tb.patch_order(idx0, 128)
    assert tb.staged_tile_plan(tb.renumber_table(idx0, order, order))[3] <= 288
