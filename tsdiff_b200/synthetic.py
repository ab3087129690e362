"""Seeded synthetic Grambow-shaped reaction graphs.

The real inputs come from utils/datasets.py:407-519 `generate_ts_data2` (rdkit, absent
here): per reaction a condensed graph whose `edge_index` is the symmetric union of the
reactant and product bonds sorted by row*N+col (datasets.py:495-498) and whose
`edge_type = r_type * 22 + p_type` (datasets.py:507); `r_feat` / `p_feat` are (N, 25)
one-hot blocks of sizes (2,3,4,3,4,4,3,2) (preprocessing.py:151-164).  This module draws
graphs with the same encoding and size statistics; it is input plumbing for tests and
bench.py, not part of the measured path.
"""
import math

import numpy as np
import torch

NUM_BOND_TYPES = 22
FEAT_GROUPS = (2, 3, 4, 3, 4, 4, 3, 2)  # sums to 25
_VALENCE = {6: 4, 7: 3, 8: 2, 1: 1}


def _one_reaction(rng, n_atoms):
    n_heavy = max(2, int(math.ceil(0.4 * n_atoms)))
    z = np.ones(n_atoms, dtype=np.int64)
    z[:n_heavy] = rng.choice([6, 7, 8], size=n_heavy, p=[0.75, 0.10, 0.15])
    z[0] = 6
    free = np.array([_VALENCE[int(a)] for a in z])
    bonds = {}

    def add(a, b, order):
        a, b = (a, b) if a < b else (b, a)
        if a == b or (a, b) in bonds or free[a] < order or free[b] < order:
            return False
        bonds[(a, b)] = order
        free[a] -= order
        free[b] -= order
        return True

    # random tree on the heavy atoms
    for v in range(1, n_heavy):
        for _ in range(8):
            u = int(rng.randint(0, v))
            order = int(rng.choice([1, 2, 3], p=[0.8, 0.17, 0.03]))
            if add(u, v, order) or add(u, v, 1):
                break
        else:
            cand = [u for u in range(v) if free[u] >= 1]
            if cand:
                add(cand[0], v, 1)
            else:  # saturated skeleton: bond to atom 0 anyway (over-valent but connected)
                bonds[(0, v)] = 1
    # at most one ring closure
    if n_heavy >= 4 and rng.rand() < 0.5:
        a, b = rng.choice(n_heavy, size=2, replace=False)
        add(int(a), int(b), 1)
    # hydrogens onto heavy atoms with free valence (round-robin), leftovers onto atom 0's neighbours
    for hyd in range(n_heavy, n_atoms):
        cand = [u for u in range(n_heavy) if free[u] >= 1 and z[u] != 1]
        if cand:
            add(int(cand[int(rng.randint(0, len(cand)))]), hyd, 1)
        else:
            # no free valence left: attach anyway to a random heavy atom (over-valent but a valid graph)
            u = int(rng.randint(0, n_heavy))
            bonds[(u, hyd)] = 1
    reactant = dict(bonds)

    # product: 1-3 edits (break / form / re-order)
    product = dict(reactant)
    keys = list(reactant.keys())
    for _ in range(int(rng.randint(1, 4))):
        kind = rng.randint(0, 3)
        if kind == 0 and len(product) > 1:
            k = keys[int(rng.randint(0, len(keys)))]
            product.pop(k, None)
        elif kind == 1:
            a, b = rng.choice(n_atoms, size=2, replace=False)
            a, b = (int(a), int(b)) if a < b else (int(b), int(a))
            product.setdefault((a, b), 1)
        else:
            k = keys[int(rng.randint(0, len(keys)))]
            if k in product:
                product[k] = 1 + (product[k] % 3)
    pairs = sorted(set(reactant) | set(product))
    rows, cols, types = [], [], []
    for (a, b) in pairs:
        t = reactant.get((a, b), 0) * NUM_BOND_TYPES + product.get((a, b), 0)
        rows += [a, b]
        cols += [b, a]
        types += [t, t]
    feats = []
    for _ in range(2):
        f = np.zeros((n_atoms, sum(FEAT_GROUPS)), dtype=np.int64)
        off = 0
        for g in FEAT_GROUPS:
            f[np.arange(n_atoms), off + rng.randint(0, g, size=n_atoms)] = 1
            off += g
        feats.append(f)
    return z, np.array(rows), np.array(cols), np.array(types), feats[0], feats[1]


def make_batch(num_graphs, seed=0, min_atoms=10, max_atoms=25, sizes=None):
    """Returns a dict of CPU tensors shaped like a torch_geometric Batch of TS graphs:
    atom_type (N,), r_feat/p_feat (N,25) int64, bond_index (2,E_b) int64 sorted by
    row*N+col, bond_type (E_b,) int64, batch (N,) int64 sorted, num_graphs, pos_init (N,3)
    ~ N(0,1) (sampling.py:190), num_nodes_per_graph (G,)."""
    rng = np.random.RandomState(seed)
    if sizes is None:
        sizes = rng.randint(min_atoms, max_atoms + 1, size=num_graphs)
    zs, rows, cols, types, rf, pf, batch = [], [], [], [], [], [], []
    off = 0
    for g, n in enumerate(sizes):
        z, r, c, t, fr, fp = _one_reaction(rng, int(n))
        zs.append(z)
        rows.append(r + off)
        cols.append(c + off)
        types.append(t)
        rf.append(fr)
        pf.append(fp)
        batch.append(np.full(int(n), g, dtype=np.int64))
        off += int(n)
    row, col, typ = np.concatenate(rows), np.concatenate(cols), np.concatenate(types)
    order = np.argsort(row * off + col, kind="stable")
    gen = torch.Generator().manual_seed(seed + 1)
    return {
        "atom_type": torch.from_numpy(np.concatenate(zs)),
        "r_feat": torch.from_numpy(np.concatenate(rf)),
        "p_feat": torch.from_numpy(np.concatenate(pf)),
        "bond_index": torch.from_numpy(np.stack([row[order], col[order]])),
        "bond_type": torch.from_numpy(typ[order]),
        "batch": torch.from_numpy(np.concatenate(batch)),
        "num_graphs": int(len(sizes)),
        "num_nodes_per_graph": torch.as_tensor(np.asarray(sizes), dtype=torch.long),
        "pos_init": torch.randn(off, 3, generator=gen),
    }


def shard_batch(data, rank, world_size):
    """Contiguous shard of the reaction list for one rank: graphs
    [lo, hi) with node / bond indices rebased to the shard.  Also returns the global atom
    offset so in-kernel Philox noise is identical for any world size."""
    g = data["num_graphs"]
    lo = (g * rank) // world_size
    hi = (g * (rank + 1)) // world_size
    batch = data["batch"]
    node_mask = (batch >= lo) & (batch < hi)
    node_idx = node_mask.nonzero(as_tuple=True)[0]
    n0 = int(node_idx[0]) if node_idx.numel() else 0
    bi = data["bond_index"]
    bmask = node_mask[bi[0]]
    out = {
        "atom_type": data["atom_type"][node_mask],
        "r_feat": data["r_feat"][node_mask],
        "p_feat": data["p_feat"][node_mask],
        "bond_index": bi[:, bmask] - n0,
        "bond_type": data["bond_type"][bmask],
        "batch": batch[node_mask] - lo,
        "num_graphs": hi - lo,
        "num_nodes_per_graph": data["num_nodes_per_graph"][lo:hi],
        "pos_init": data["pos_init"][node_mask],
        "atom_offset": n0,
        "graph_range": (lo, hi),
    }
    return out
