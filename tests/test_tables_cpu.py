"""Host-side table derivation (bit-exact integer work) and package plumbing -- no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from sdvae_b200 import cabi, fixtures as fx, tables as tb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_inverse_cells_roundtrip(cranio):
    for lvl, idx in enumerate(cranio.spirals):
        V, S = idx.shape
        ptr, src = tb.inverse_cells(idx, V)
        assert ptr[0] == 0 and ptr[-1] == V * S and np.all(np.diff(ptr) >= 0)
        cell = np.repeat(np.arange(V * S), np.diff(ptr))
        u, s = cell // S, cell % S
        assert np.array_equal(idx[src, s], u)                 # every entry points back
        # ascending rows inside each cell
        same = cell[1:] == cell[:-1]
        assert np.all(src[1:][same] > src[:-1][same])
        # slot 0 is the vertex itself: exactly one entry, itself
        assert np.array_equal(np.diff(ptr)[::S], np.ones(V, np.int32))


def test_inverse_flat_matches_cells(cranio):
    idx = cranio.spirals[1]
    V, S = idx.shape
    ptr, flat = tb.inverse_rows_flat(idx, V)
    assert ptr[-1] == V * S
    tgt = np.repeat(np.arange(V), np.diff(ptr))
    assert np.array_equal(idx.ravel()[flat], tgt)


def test_restricted_table_is_row_subset(cranio):
    dn = cranio.down[0]
    kept = tb.selection_columns(dn[0], dn[1], dn[2], dn[3][0])
    assert kept is not None and np.array_equal(kept, dn[1])
    assert np.all(np.diff(kept) > 0)
    up = cranio.up[0]
    assert tb.selection_columns(up[0], up[1], up[2], up[3][0]) is None


def test_ell_keeps_storage_order(cranio):
    row, col, val, shape = cranio.up[0]
    ec, ev = tb.ell_from_coo(row, col, val, *shape)
    assert ec.shape == (shape[0], 3) and np.all(ec >= 0)
    # rebuild the COO row by row and compare with a stable sort of the original
    order = np.argsort(row, kind='stable')
    assert np.array_equal(ec.ravel(), col[order].astype(np.int32))
    assert np.array_equal(ev.ravel(), val[order])
    tp, tr, tv = tb.transposed_csr(row, col, val, shape[1])
    deg = np.diff(tp)
    assert deg.min() >= 1 and deg.max() == 96 and tp[-1] == row.size
    # storage is column-major, so the transposed CSR is the storage order itself
    assert np.array_equal(tr, row.astype(np.int32)) and np.array_equal(tv, val)


def test_ell_ragged_padding():
    row = np.array([2, 0, 2, 2]); col = np.array([1, 3, 0, 2]); val = np.array([1., 2., 3., 4.], np.float32)
    ec, ev = tb.ell_from_coo(row, col, val, 3, 4)
    assert ec.tolist() == [[3, -1, -1], [-1, -1, -1], [1, 0, 2]]
    assert ev[2].tolist() == [1., 3., 4.]


def test_index_range_checks():
    with pytest.raises(IndexError):
        tb.check_indices(np.array([[0, 5]]), n_src=5)
    with pytest.raises(IndexError):
        tb.check_indices(np.array([[-1, 0]]), n_src=5)
    with pytest.raises(ValueError):
        tb.check_indices(np.arange(4))


def test_laplacian_rows_sum_to_zero(cranio):
    row, col, val = cranio.lap
    V = cranio.num_vertices[0]
    assert row.size == 118595
    sums = np.bincount(row, weights=val.astype(np.float64), minlength=V)
    assert np.abs(sums).max() < 1e-6


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'sdvae_b200.h')).read()
    declared = set(re.findall(r'\b(sdvae_[a-z0-9_]+)\s*\(', hdr))
    assert declared == set(cabi.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert cabi.load().sdvae_abi_version() == 1


def test_model_state_dict_keys_match_reference_layout(cranio):
    from oracle import sdvae_oracle as orc
    from sdvae_b200.model import Model, MLPClassifier
    sp, dn, up = cranio.spiral_tensors(), cranio.down_tensors(), cranio.up_tensors()
    m = Model(3, [32, 32, 32, 64], 75, sp, dn, up, False, True)
    shapes = orc.Net(3, [32, 32, 32, 64], 75, sp, dn, up, False, True).param_shapes()
    sd = m.state_dict()
    assert list(sd) == list(shapes)
    assert all(tuple(sd[k].shape) == shapes[k] for k in shapes)
    assert sum(v.numel() for v in sd.values()) == 1081881
    assert all(float(v.abs().max()) == 0.0 for k, v in sd.items() if k.endswith('bias'))
    assert repr(m.en_layers[0].conv) == 'SpiralConv(3, 32, seq_length=9)'
    clf = MLPClassifier(75, [16, 8], 5)
    out, lab = clf(torch.randn(4, 75))
    assert out.shape == (4, 5) and lab.shape == (4,) and float(out.min()) >= 0.0


def test_cpu_tensors_are_rejected(cranio):
    from sdvae_b200.model import SpiralConv, Pool
    conv = SpiralConv(3, 8, cranio.spiral_tensors()[3])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        conv(torch.randn(2, 267, 3))
    with pytest.raises(RuntimeError, match='expected to be 2 or 3'):
        conv(torch.randn(1, 2, 267, 3))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        Pool(torch.randn(2, 67, 4), cranio.up_tensors()[3])


def test_pool_stage_plan_reproduces_ell_columns():
    """Stage plan of the shared-memory Pool forward: per tile the ascending distinct source rows, entries
    re-addressed into that list, values and entry order untouched."""
    from sdvae_b200 import fixtures as fx, tables as tb
    tabs = fx.craniofacial_tables()
    for trans in list(tabs.up_tensors()) + list(tabs.down_tensors()):
        ind = trans._indices().numpy()
        ec, ev = tb.ell_from_coo(ind[0], ind[1], trans._values().numpy(), trans.shape[0], trans.shape[1])
        for T in (32, tb.POOL_STAGE_TILE):
            tp, ss, ent, ucap = tb.pool_stage_plan(ec, ev, T)
            assert tp[0] == 0 and ucap == np.diff(tp).max() and np.array_equal(ent[:, :, 1].view(np.float32), ev)
            for t in range(len(tp) - 1):
                lst = ss[tp[t]:tp[t + 1]]
                assert np.all(np.diff(lst) > 0)                        # distinct, ascending
                blk, loc = ec[t * T:(t + 1) * T], ent[t * T:(t + 1) * T, :, 0]
                assert np.array_equal(loc >= 0, blk >= 0)
                assert np.array_equal(lst[loc[loc >= 0]], blk[blk >= 0])
                assert set(lst.tolist()) == set(blk[blk >= 0].tolist())


def test_pack_cells16_first_rows_and_overflow_marker():
    """``cell_pack`` of the fused output-layer backward: first four rows of each cell as 16-bit element offsets
    row*3, ``n_rows*3`` = none (zero pad), 0xFFFF in the fourth field = longer cell (rest read from the CSR)."""
    from sdvae_b200 import fixtures as fx, tables as tb
    idx = fx.craniofacial_tables().spiral_tensors()[0].numpy()
    V = idx.shape[0]
    ptr, src = tb.inverse_cells(idx, V)
    pk = tb.pack_cells16(ptr, src, V, 3).view(np.uint32)
    f = np.stack([pk[:, 0] & 0xFFFF, pk[:, 0] >> 16, pk[:, 1] & 0xFFFF, pk[:, 1] >> 16], 1).astype(np.int64)
    cnt = np.diff(ptr)
    assert cnt.max() > 4                                   # the template has cells that overflow
    for k in range(4):
        has = cnt > k
        if k == 3:
            assert np.all(f[cnt > 4, 3] == 0xFFFF)
            has = cnt == 4
            assert np.all(f[cnt < 4, 3] == V * 3)
        else:
            assert np.all(f[~has, k] == V * 3)
        assert np.array_equal(f[has, k], 3 * src[ptr[:-1][has] + k])
    with pytest.raises(IndexError):
        tb.pack_cells16(np.array([0, 1]), np.array([0]), 30000, 3)


def test_patch_order_halves_the_distinct_rows_of_a_tile():
    """Groundwork for tile-local staging in the tensor-core gather kernels: in patch order a 128-row
    tile of the level-0 spiral table reads about half the distinct rows it reads in the template's strip order."""
    from sdvae_b200 import fixtures as fx, tables as tb
    idx = fx.craniofacial_tables().spiral_tensors()[0].numpy()
    V = idx.shape[0]
    order = tb.patch_order(idx, 128)
    assert order.shape == (V,) and np.array_equal(np.sort(order), np.arange(V))
    new = tb.renumber_table(idx, order, order)
    # same graph: row p of the new table is row order[p] of the old one, renamed
    assert np.array_equal(order[new], idx[order])
    old_d, new_d = tb.distinct_rows_per_row(idx, 128), tb.distinct_rows_per_row(new, 128)
    assert old_d > 3.0 and new_d < 1.8, (old_d, new_d)
    small = np.array([[0, 1, 2], [1, 0, 2], [2, 1, 3], [3, 2, -1]])
    assert tb.distinct_rows_per_row(small, 2) == 1.5
    assert np.array_equal(tb.renumber_table(small, [3, 2, 1, 0], [3, 2, 1, 0]),
                          np.array([[0, 1, -1], [1, 2, 0], [2, 3, 1], [3, 2, 1]]))


def test_renumbered_model_tables_and_laplacian_are_consistent():
    """Host side of TrainEngine(renumber=True): ``renumbered_model_tables`` on the model's own tensors equals
    ``MeshTables.renumbered``; ``LaplacianTable.renumbered`` is the same operator on the new numbering."""
    from sdvae_b200 import fixtures as fx, tables as tb
    from sdvae_b200.losses import LaplacianTable
    tabs = fx.craniofacial_tables()
    new, orders = tabs.renumbered(128)
    sp, dn, up, orders2 = tb.renumbered_model_tables(tabs.spiral_tensors(), tabs.down_tensors(), tabs.up_tensors())
    assert all(np.array_equal(a, b) for a, b in zip(orders, orders2))
    for a, b in zip(sp, new.spiral_tensors()):
        assert torch.equal(a, b)
    for got, want in zip(dn + up, new.down_tensors() + new.up_tensors()):
        assert got.shape == want.shape and torch.equal(got._indices(), want._indices())
        assert torch.equal(got._values(), want._values())
    V = tabs.num_vertices[0]
    lt = LaplacianTable.build(*tabs.lap, V, 'cpu')
    ln = lt.renumbered(orders[0])
    want = LaplacianTable.build(*new.lap, V, 'cpu')
    for f in ('ell_col', 'ell_val', 't_ptr', 't_row', 't_val'):
        assert torch.equal(getattr(ln, f), getattr(want, f)), f


def _staged_demo():
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from baseline import refarm
    ref = refarm.find_ref()
    if ref is None or not os.path.exists(os.path.join(ref, 'demo_files', 'region_ldas.pkl')):
        import pytest
        pytest.skip('reference not staged (tools/stage_reference.py needs /root/reference)')
    return os.path.join(ref, 'demo_files')


def test_region_extraction_is_pinned_to_region_ldas_keys(cranio):
    """fixtures.extract_regions (restating utils.py:93-135 without trimesh / networkx) must reproduce the region
    names AND their order that the reference's own run stored in demo_files/region_ldas.pkl -- the order defines the
    latent slices [5k, 5k+5) of model_manager.py:232-238."""
    import pickle
    demo = _staged_demo()

    class _Any:                                   # the pickled values are sklearn LDA objects: only the keys matter
        def __init__(self, *a, **k): pass
        def __setstate__(self, st): pass

    class U(pickle.Unpickler):
        def find_class(self, module, name):
            try:
                return super().find_class(module, name)
            except Exception:
                return _Any
    with open(os.path.join(demo, 'region_ldas.pkl'), 'rb') as f:
        ldas = U(f).load()
    assert list(ldas.keys()) == cranio.region_keys()
    lat = cranio.latent_regions(75)
    assert [lat[k] for k in ldas.keys()] == [[5 * i, 5 * i + 5] for i in range(15)]


def test_packaged_tables_equal_the_reference_pkl_files(cranio):
    """data/craniofacial_tables.npz (what the product path loads) == tables read straight from the reference's own
    spirals.pkl / transforms.pkl / template.ply."""
    from sdvae_b200 import fixtures as fx
    live = fx.craniofacial_tables(_staged_demo())
    for a, b in zip(live.spirals, cranio.spirals):
        assert np.array_equal(a, b)
    for la, lb in ((live.down, cranio.down), (live.up, cranio.up)):
        for (r1, c1, v1, s1), (r2, c2, v2, s2) in zip(la, lb):
            assert np.array_equal(r1, r2) and np.array_equal(c1, c2) and np.array_equal(v1, v2) and tuple(s1) == tuple(s2)
    assert live.region_keys() == cranio.region_keys()
    for (k1, i1), (k2, i2) in zip(live.regions, cranio.regions):
        assert k1 == k2 and np.array_equal(i1, i2)
    for a, b in zip(live.lap, cranio.lap):
        assert np.array_equal(a, b)
