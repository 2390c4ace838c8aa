"""Readers for the reference's on-disk inputs and a synthetic table generator.

The hot path consumes three kinds of precomputed tables:

* ``spirals.pkl``     -- ``list[LongTensor[V_l, S]]``            (model_manager.py:211-230)
* ``transforms.pkl``  -- ``[low_res_templates, down, up]`` where ``down``/``up``
  are *uncoalesced* fp32 ``torch.sparse_coo`` matrices                (model_manager.py:176-209)
* the coloured template ``.ply`` from which the reference derives the
  random-walk Laplacian (utils.py:87-90) and the per-region vertex lists used
  by the feature swap and the latent-consistency loss (utils.py:93-144,
  model_manager.py:232-238).

Nothing here needs torch_geometric / trimesh: the pickles are opened with a
stub for ``torch_geometric.data.data.Data`` and the PLY is parsed by hand.
Everything is host-side preprocessing; no kernel is involved.
"""
from __future__ import annotations

import io
import os
import pickle
import struct
from collections import Counter, OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# colour -> anatomical name, as printed by numpy for a uint8 RGBA row
# (reference: utils.py:43-59; used only for pretty names, never for indexing).
REGION_NAMES = {
    '[232 129 166 255]': 'upper lip', '[194 109  97 255]': 'chin',
    '[133 169 172 255]': 'nasolabial', '[237 109  93 255]': 'nose',
    '[ 89  51 139 255]': 'cheeks', '[245 158  40 255]': 'zygomatic',
    '[ 26  81  82 255]': 'eyes', '[164  78 123 255]': 'jaw',
    '[238 206  74 255]': 'supraorbital', '[ 18  78 129 255]': 'neck',
    '[245 160 106 255]': 'ears', '[116 192 194 255]': 'frontal',
    '[ 90  97 115 255]': 'occipital', '[164 184 207 255]': 'temporal',
    '[219 203 190 255]': 'parietal',
}


# --------------------------------------------------------------------------
# pickles
# --------------------------------------------------------------------------
class _DataStub:
    """Stand-in for ``torch_geometric.data.data.Data`` when unpickling."""

    def __setstate__(self, state):
        self.__dict__.update(state)

    def __getattr__(self, item):  # missing attrs read as None, like PyG
        if item.startswith('__'):
            raise AttributeError(item)
        return None


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith('torch_geometric'):
            return _DataStub
        return super().find_class(module, name)


def load_spirals_pkl(path: str) -> List[torch.Tensor]:
    """``spirals.pkl`` -> list of int64 ``[V_l, S]`` tensors (model_manager.py:213-216)."""
    with open(path, 'rb') as fh:
        spirals = pickle.load(fh)
    return [s.contiguous() for s in spirals]


def load_transforms_pkl(path: str):
    """``transforms.pkl`` -> ``(low_res_templates, down, up)`` (model_manager.py:179-182).

    The sparse matrices are returned exactly as stored (uncoalesced COO)."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        with open(path, 'rb') as fh:
            low, down, up = _Unpickler(fh).load()
    return low, list(down), list(up)


# --------------------------------------------------------------------------
# meshes
# --------------------------------------------------------------------------
def read_ply(path: str):
    """Binary little-endian PLY with ``float x,y,z; uchar r,g,b,a`` vertices and
    ``uchar n; int idx[n]`` triangle faces (the layout of demo_files/template.ply)."""
    with open(path, 'rb') as fh:
        blob = fh.read()
    end = blob.index(b'end_header\n') + len(b'end_header\n')
    header = blob[:end].decode('ascii', 'replace').splitlines()
    n_vert = n_face = 0
    vprops: List[Tuple[str, str]] = []
    section = None
    for line in header:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == 'format' and tok[1] != 'binary_little_endian':
            raise ValueError('only binary_little_endian PLY is supported')
        if tok[0] == 'element':
            section = tok[1]
            if section == 'vertex':
                n_vert = int(tok[2])
            elif section == 'face':
                n_face = int(tok[2])
        elif tok[0] == 'property' and section == 'vertex':
            vprops.append((tok[1], tok[2]))
    np_types = {'float': '<f4', 'double': '<f8', 'uchar': 'u1', 'int': '<i4',
                'uint': '<u4', 'short': '<i2', 'ushort': '<u2', 'char': 'i1'}
    vdtype = np.dtype([(name, np_types[typ]) for typ, name in vprops])
    verts = np.frombuffer(blob, dtype=vdtype, count=n_vert, offset=end)
    pos = np.stack([verts['x'], verts['y'], verts['z']], 1).astype(np.float32)
    if 'red' in verts.dtype.names:
        alpha = verts['alpha'] if 'alpha' in verts.dtype.names else \
            np.full(n_vert, 255, np.uint8)
        colors = np.stack([verts['red'], verts['green'], verts['blue'], alpha], 1)
    else:
        colors = np.full((n_vert, 4), 255, np.uint8)
    fdtype = np.dtype([('n', 'u1'), ('idx', '<i4', (3,))])
    faces = np.frombuffer(blob, dtype=fdtype, count=n_face,
                          offset=end + n_vert * vdtype.itemsize)
    if not np.all(faces['n'] == 3):
        raise ValueError('non-triangular face in PLY')
    return pos, colors.astype(np.uint8), faces['idx'].astype(np.int64)


def read_obj_vertices(path: str) -> np.ndarray:
    """``v x y z`` lines of a Wavefront OBJ -> float32 ``[V, 3]``."""
    rows = []
    with open(path, 'r') as fh:
        for line in fh:
            if line.startswith('v '):
                rows.append([float(t) for t in line.split()[1:4]])
    return np.asarray(rows, np.float32)


def unique_edges(faces: np.ndarray) -> np.ndarray:
    """Undirected unique edges ``[E, 2]`` (a < b), sorted lexicographically."""
    f = np.asarray(faces, np.int64)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
    e.sort(axis=1)
    return np.unique(e, axis=0)


def one_rings(edges: np.ndarray, n_vert: int) -> List[List[int]]:
    """Adjacency lists in edge-insertion order (what ``nx.from_edgelist`` keeps,
    utils.py:106-107)."""
    rings: List[List[int]] = [[] for _ in range(n_vert)]
    for a, b in edges.tolist():
        rings[a].append(b)
        rings[b].append(a)
    return rings


def extract_regions(colors: np.ndarray, faces: np.ndarray) -> "OrderedDict[str, dict]":
    """Restatement of ``extract_feature_and_contour_from_colour`` (utils.py:93-135).

    A vertex is *contour* if any one-ring neighbour has a different colour, else
    *feature*.  Colours with fewer than 3 feature vertices are interpolation
    artefacts: their feature vertices are handed to the most common neighbouring
    colour and the colour is dropped.  Key order = first appearance by vertex
    index, which fixes the latent-region order (model_manager.py:232-238)."""
    colors = np.asarray(colors)
    n_vert = colors.shape[0]
    rings = one_rings(unique_edges(faces), n_vert)
    keys = [str(c) for c in colors]
    regions: "OrderedDict[str, dict]" = OrderedDict()
    for v in range(n_vert):
        k = keys[v]
        if k not in regions:
            regions[k] = {'feature': [], 'contour': []}
        is_contour = any(keys[r] != k for r in rings[v])
        regions[k]['contour' if is_contour else 'feature'].append(v)
    drop = []
    for k, reg in regions.items():
        if len(reg['feature']) < 3:
            drop.append(k)
            for v in reg['feature']:
                most = Counter(keys[r] for r in rings[v]).most_common(1)[0][0]
                if most == k:
                    break
                regions[most]['feature'].append(v)
                regions[most]['contour'].append(v)
    for k in drop:
        regions.pop(k, None)
    return regions


def rw_laplacian(faces: np.ndarray, n_vert: int):
    """Random-walk Laplacian ``L = I - D^-1 A`` as COO ``(row, col, val)`` in the
    entry order PyG produces (utils.py:87-90: ``FaceToEdge`` -> coalesced
    directed edges sorted by (row, col); ``get_laplacian(..., 'rw')`` emits the
    off-diagonal ``-1/deg[row]`` entries first and appends the unit diagonal)."""
    e = unique_edges(faces)
    both = np.concatenate([e, e[:, ::-1]], 0)
    order = np.lexsort((both[:, 1], both[:, 0]))
    both = both[order]
    row, col = both[:, 0], both[:, 1]
    deg = np.bincount(row, minlength=n_vert).astype(np.float32)
    with np.errstate(divide='ignore'):
        inv = np.where(deg > 0, np.float32(1.0) / deg, np.float32(0.0)).astype(np.float32)
    off = -(inv[row] * np.float32(1.0))
    diag = np.arange(n_vert, dtype=np.int64)
    return (np.concatenate([row, diag]).astype(np.int64),
            np.concatenate([col, diag]).astype(np.int64),
            np.concatenate([off, np.ones(n_vert, np.float32)]).astype(np.float32))


# --------------------------------------------------------------------------
# table bundle
# --------------------------------------------------------------------------
@dataclass
class MeshTables:
    """Everything static the hot path needs for one template.

    ``down[l]`` / ``up[l]`` are ``(row, col, val, (n_rows, n_cols))`` in *storage
    order*; ``regions`` is an ordered list of ``(key, feature_vertex_ids)``."""
    spirals: List[np.ndarray]
    down: List[Tuple[np.ndarray, np.ndarray, np.ndarray, Tuple[int, int]]]
    up: List[Tuple[np.ndarray, np.ndarray, np.ndarray, Tuple[int, int]]]
    lap: Tuple[np.ndarray, np.ndarray, np.ndarray]
    regions: List[Tuple[str, np.ndarray]] = field(default_factory=list)
    name: str = 'unnamed'

    @property
    def num_vertices(self) -> List[int]:
        return [int(s.shape[0]) for s in self.spirals] + [int(self.down[-1][3][0])]

    # ---- torch views in the reference's own input formats -----------------
    def spiral_tensors(self, device='cpu') -> List[torch.Tensor]:
        return [torch.from_numpy(np.ascontiguousarray(s)).long().to(device)
                for s in self.spirals]

    @staticmethod
    def _coo(entry, device):
        row, col, val, shape = entry
        ind = torch.from_numpy(np.stack([row, col]).astype(np.int64))
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            t = torch.sparse_coo_tensor(ind, torch.from_numpy(val.astype(np.float32)),
                                        torch.Size(shape))
        return t.to(device)

    def down_tensors(self, device='cpu'):
        return [self._coo(e, device) for e in self.down]

    def up_tensors(self, device='cpu'):
        return [self._coo(e, device) for e in self.up]

    def laplacian_tensor(self, device='cpu'):
        v = self.num_vertices[0]
        return self._coo((self.lap[0], self.lap[1], self.lap[2], (v, v)), device)

    def region_keys(self) -> List[str]:
        return [k for k, _ in self.regions]

    def latent_regions(self, latent_size: int) -> Dict[str, List[int]]:
        """model_manager.py:232-238."""
        n = len(self.regions)
        if n == 0 or latent_size % n:
            raise ValueError('latent_size %d is not divisible by %d regions' % (latent_size, n))
        w = latent_size // n
        return {k: [i * w, (i + 1) * w] for i, (k, _) in enumerate(self.regions)}

    # ---- the same template with its vertices renumbered patch-wise ----------
    def renumbered(self, tile: int = 128) -> Tuple["MeshTables", List[np.ndarray]]:
        """The same meshes and operators with the vertices of every level but the coarsest listed in patch order
        (``tables.patch_order``: each run of ``tile`` consecutive vertices is a compact patch; the coarsest level
        keeps its order so that the latent ``Linear`` layers see the same flattening, model.py:151,166).
        Returns the new tables and ``orders[l]`` (new position -> old vertex; identity for the coarsest level).
        A network on the new tables maps ``x[:, orders[0]]`` to ``recon[:, orders[0]]`` with the same latent
        codes: every per-vertex sum keeps its terms and their order (spiral slots, storage order of the
        transform entries).  Groundwork for tile-local staging; nothing applies it by default."""
        from .tables import renumber_levels
        spirals, down, up, orders = renumber_levels(self.spirals, self.down, self.up, tile)
        rank0 = np.empty(orders[0].size, np.int64)
        rank0[orders[0]] = np.arange(orders[0].size)
        lap = (rank0[np.asarray(self.lap[0], np.int64)], rank0[np.asarray(self.lap[1], np.int64)], self.lap[2])
        regions = [(k, np.sort(rank0[np.asarray(idx, np.int64)])) for k, idx in self.regions]
        return MeshTables(spirals, down, up, lap, regions, self.name + '-patch%d' % tile), orders

    # ---- compact on-disk bundle (tests/golden/*.npz) -----------------------
    def save_npz(self, path: str) -> None:
        out = {'name': np.array(self.name), 'n_levels': np.array(len(self.spirals))}
        for i, s in enumerate(self.spirals):
            out['spiral%d' % i] = s.astype(np.int32)
        for tag, mats in (('down', self.down), ('up', self.up)):
            for i, (r, c, v, shape) in enumerate(mats):
                out['%s%d_row' % (tag, i)] = r.astype(np.int32)
                out['%s%d_col' % (tag, i)] = c.astype(np.int32)
                out['%s%d_val' % (tag, i)] = v.astype(np.float32)
                out['%s%d_shape' % (tag, i)] = np.asarray(shape, np.int64)
        out['lap_row'] = self.lap[0].astype(np.int32)
        out['lap_col'] = self.lap[1].astype(np.int32)
        out['lap_val'] = self.lap[2].astype(np.float32)
        out['region_keys'] = np.array([k for k, _ in self.regions])
        for i, (_, idx) in enumerate(self.regions):
            out['region%d' % i] = np.asarray(idx, np.int32)
        np.savez_compressed(path, **out)

    @staticmethod
    def load_npz(path: str) -> "MeshTables":
        z = np.load(path, allow_pickle=False)
        n = int(z['n_levels'])
        spirals = [z['spiral%d' % i].astype(np.int64) for i in range(n)]

        def mats(tag):
            return [(z['%s%d_row' % (tag, i)].astype(np.int64),
                     z['%s%d_col' % (tag, i)].astype(np.int64),
                     z['%s%d_val' % (tag, i)].astype(np.float32),
                     tuple(int(t) for t in z['%s%d_shape' % (tag, i)])) for i in range(n)]
        keys = [str(k) for k in z['region_keys']]
        regions = [(k, z['region%d' % i].astype(np.int64)) for i, k in enumerate(keys)]
        lap = (z['lap_row'].astype(np.int64), z['lap_col'].astype(np.int64),
               z['lap_val'].astype(np.float32))
        return MeshTables(spirals, mats('down'), mats('up'), lap, regions, str(z['name']))


def tables_from_reference_files(spirals_pkl: str, transforms_pkl: str, template_ply: str,
                                name='craniofacial') -> MeshTables:
    """Build a :class:`MeshTables` from the reference's own demo files."""
    spirals = [s.numpy().astype(np.int64) for s in load_spirals_pkl(spirals_pkl)]
    _, down, up = load_transforms_pkl(transforms_pkl)

    def unpack(m):
        ind = m._indices().numpy()
        return (ind[0].astype(np.int64), ind[1].astype(np.int64),
                m._values().numpy().astype(np.float32), (int(m.size(0)), int(m.size(1))))
    pos, colors, faces = read_ply(template_ply)
    regs = extract_regions(colors, faces)
    regions = [(k, np.asarray(v['feature'], np.int64)) for k, v in regs.items()]
    lap = rw_laplacian(faces, pos.shape[0])
    return MeshTables(spirals, [unpack(d) for d in down], [unpack(u) for u in up],
                      lap, regions, name)


# --------------------------------------------------------------------------
# synthetic tables (body.yaml has no template in the mount; unit tests need
# small meshes)
# --------------------------------------------------------------------------
def synthetic_tables(n_vertices: int, n_levels: int, seq_length: int = 9,
                     n_regions: int = 3, factor: int = 4, seed: int = 0,
                     name: str = 'synthetic') -> MeshTables:
    """Tables with the same *structure* as the reference's:

    * spiral row = the vertex itself followed by ``seq_length-1`` distinct lattice
      neighbours (a triangulated-torus style ring lattice, so locality is realistic);
    * ``down[l]`` = sorted selection of ``ceil(V/factor)`` vertices, 1 nnz/row, value 1;
    * ``up[l]`` = 3 nnz/row barycentric (kept vertices one-hot ``(1,0,0)``), stored
      column-major like the reference's CSC->COO conversion (utils.py:147-150);
    * Laplacian from the level-0 lattice graph; ``n_regions`` disjoint vertex sets.
    """
    rng = np.random.RandomState(seed)
    sizes = [int(n_vertices)]
    for _ in range(n_levels):
        sizes.append(-(-sizes[-1] // factor))

    def lattice_offsets(v):
        w = max(2, int(round(np.sqrt(v))))
        return [1, -1, w, -w, w + 1, -(w + 1), 2, -2, 2 * w, -2 * w, w - 1, -(w - 1),
                2 * w + 1, -(2 * w + 1), 3, -3]

    spirals = []
    for lvl in range(n_levels):
        v = sizes[lvl]
        rows = np.empty((v, seq_length), np.int64)
        ids = np.arange(v)
        rows[:, 0] = ids
        offs = lattice_offsets(v)
        for r in range(v):
            seen = {r}
            out = []
            start = int(rng.randint(0, 6))          # spirals start at an arbitrary neighbour
            cand = offs[start:6] + offs[:start] + offs[6:]
            for o in cand:
                t = (r + o) % v
                if t not in seen:
                    seen.add(t)
                    out.append(t)
                if len(out) == seq_length - 1:
                    break
            while len(out) < seq_length - 1:         # tiny meshes: fall back to any unseen vertex
                t = int(rng.randint(0, v))
                if t not in seen or v <= seq_length:
                    seen.add(t)
                    out.append(t)
            rows[r, 1:] = out
        spirals.append(rows)

    down, up = [], []
    for lvl in range(n_levels):
        v_in, v_out = sizes[lvl], sizes[lvl + 1]
        kept = np.sort(rng.choice(v_in, v_out, replace=False)).astype(np.int64)
        down.append((np.arange(v_out, dtype=np.int64), kept,
                     np.ones(v_out, np.float32), (v_out, v_in)))
        coarse_of = np.full(v_in, -1, np.int64)
        coarse_of[kept] = np.arange(v_out)
        cols = np.empty((v_in, 3), np.int64)
        vals = np.empty((v_in, 3), np.float32)
        for r in range(v_in):
            if coarse_of[r] >= 0:
                base = int(coarse_of[r])
                others = [(base + 1) % v_out, (base + 2) % v_out] if v_out >= 3 else [base, base]
                trio = sorted([base] + others)
                w = [1.0 if t == base else 0.0 for t in trio]
                if v_out < 3:
                    trio, w = [base, base, base], [1.0, 0.0, 0.0]
            else:
                near = int(np.searchsorted(kept, r)) % v_out
                trio = sorted({near, (near + 1) % v_out, (near - 1) % v_out})
                while len(trio) < 3:
                    trio = sorted(set(trio) | {int(rng.randint(0, v_out))}) if v_out >= 3 \
                        else (trio + [trio[0]])[:3]
                a, b = rng.uniform(-0.2, 0.9, 2)
                w = [a, b, 1.0 - a - b]
            cols[r] = trio[:3]
            vals[r] = np.asarray(w[:3], np.float32)
        r_idx = np.repeat(np.arange(v_in, dtype=np.int64), 3)
        c_idx, v_flat = cols.ravel(), vals.ravel()
        order = np.argsort(c_idx, kind='stable')      # column-major storage
        up.append((r_idx[order], c_idx[order], v_flat[order], (v_in, v_out)))

    v0 = sizes[0]
    w0 = max(2, int(round(np.sqrt(v0))))
    ids = np.arange(v0)
    e = np.concatenate([np.stack([ids, (ids + o) % v0], 1) for o in (1, w0, w0 + 1)], 0)
    e = e[e[:, 0] != e[:, 1]]
    faces_like = np.concatenate([e, e[:, :1]], 1)      # degenerate "faces": only edges matter
    lap = rw_laplacian(faces_like, v0)

    perm = rng.permutation(v0)
    n_feat = max(1, v0 // (2 * n_regions))
    regions = [('region%02d' % i, np.sort(perm[i * n_feat:(i + 1) * n_feat]).astype(np.int64))
               for i in range(n_regions)]
    return MeshTables(spirals, down, up, lap, regions, name)


def default_tables_path() -> str:
    """Package data (not a test fixture): the reference's index tables re-serialised by tools/make_golden.py."""
    here = os.path.dirname(os.path.abspath(__file__))
    return os.path.join(here, 'data', 'craniofacial_tables.npz')


def craniofacial_tables(ref_demo_dir: Optional[str] = None) -> MeshTables:
    """The reference's craniofacial index tables.  With ``ref_demo_dir`` (a ``demo_files`` directory): straight from
    the reference's own input files ``spirals.pkl`` / ``transforms.pkl`` / ``template.ply``
    (``tables_from_reference_files``); otherwise from the packaged bundle ``data/craniofacial_tables.npz`` that
    ``tools/make_golden.py`` derived from exactly those files (tests/test_tables_cpu.py checks the two agree
    wherever the reference is staged)."""
    if ref_demo_dir is not None:
        return tables_from_reference_files(os.path.join(ref_demo_dir, 'spirals.pkl'),
                                           os.path.join(ref_demo_dir, 'transforms.pkl'),
                                           os.path.join(ref_demo_dir, 'template.ply'))
    p = default_tables_path()
    if not os.path.exists(p):
        raise FileNotFoundError(p + ' is missing; run tools/make_golden.py where /root/reference exists')
    return MeshTables.load_npz(p)


def build_model(tabs: "MeshTables", in_channels, out_channels, latent_size, pre_z_sigmoid, is_vae,
                seed: int, device, bias_scale: float = 0.05):
    """The drop-in ``Model`` for these tables on ``device`` with a seeded initialisation: the reference's
    xavier-uniform weights (model.py:139-144) and small random biases (all-zero biases would hide bias
    handling in benchmarks and checks).  Deterministic for a given seed, identical on every rank."""
    from .model import Model
    sp = [s.to(device) for s in tabs.spiral_tensors()]
    dn = [d.to(device) for d in tabs.down_tensors()]
    up = [u.to(device) for u in tabs.up_tensors()]
    gen = torch.Generator().manual_seed(int(seed))
    model = Model(in_channels, out_channels, latent_size, sp, dn, up, pre_z_sigmoid, is_vae)
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            if p.dim() >= 2:
                bound = (6.0 / (p.shape[0] + p.shape[1])) ** 0.5
                p.copy_((torch.rand(p.shape, generator=gen) * 2 - 1) * bound)
            else:
                p.copy_(torch.randn(p.shape, generator=gen) * bias_scale)
    return model.to(device)


def swap_features_torch(x: torch.Tensor, feature_vertices) -> torch.Tensor:
    """``[bs, V, C] -> [bs*bs, V, C]``: element i*bs + j is base mesh i with the feature's vertices taken from
    mesh j (swap_batch_transform.py:27-38), by plain indexing -- input preparation for tools."""
    bs = x.shape[0]
    out = x.unsqueeze(1).repeat(1, bs, 1, 1)                 # [i, j] = x[i]
    idx = torch.as_tensor(feature_vertices, dtype=torch.long, device=x.device)
    out[:, :, idx] = x[:, idx].unsqueeze(0).expand(bs, bs, idx.numel(), x.shape[2])
    return out.reshape(bs * bs, x.shape[1], x.shape[2])

