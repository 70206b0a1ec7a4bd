"""Varlen plan: how one batch of molecules is laid out for the kernels.

The reference turns the padded dense batch into a sparse edge list on every forward
(``adj_mask.nonzero`` + ``dense_to_sparse``, reference models/mol_gnn.py:512-514, two host syncs).
Here the layout depends only on the node mask, so it is built once per mask and reused for every
denoiser call of a sampling run:

* atoms are packed (padding removed) in (molecule, atom) order -> node index in [0, Nn);
* directed edges live in tiles of 128 rows (= one tcgen05 M tile).  A *group* is the set of all
  (n-1) partners of one atom; groups are packed greedily into tiles and never split, so every
  per-atom reduction over partners (softmax over sources, message sum, coordinate sum) is local to
  a tile.  Row (g, j) stands for the ordered pair whose group atom is g and partner is j; because
  every edge feature is symmetric the same stored row serves as edge
  (r=j -> c=g) in the attention pass and as (r=g, c=j) in the coordinate update.
"""
from __future__ import annotations

import numpy as np
import torch

TILE = 128
MAX_GROUPS = 64          # groups per tile (the attention kernel's group-sum MMA has 64 output slots)
WINDOW = 256             # open tiles considered by the best-fit packer
MAX_LOOSE = 255          # largest group of a loose plan (the wide attention kernel keeps one group's logits in shared memory)


class Plan:
    def __init__(self, node_mask: torch.Tensor, device=None, loose=None):
        """node_mask: [B, N, 1] or [B, N] (0/1).  Built on the host (one D2H copy of the mask).

        loose: groups are laid out back to back and may straddle tiles -- the layout of the wide path
        (jodo_b200/wide.py), whose row kernels do not need a group inside one tile; it lifts the 129-atom limit of the
        fused edge-tile kernels.  None = loose only when some molecule has more than TILE + 1 atoms."""
        m = node_mask.detach()
        if m.dim() == 3:
            m = m[..., 0]
        m = (m > 0).cpu().numpy()
        B, N = m.shape
        self.B, self.N = B, N
        n = m.sum(1).astype(np.int64)
        self.n_nodes = n
        if loose is None:
            loose = bool(n.max(initial=0) - 1 > TILE)
        self.loose = loose
        if n.max(initial=0) - 1 > (MAX_LOOSE if loose else TILE):
            raise ValueError(f'molecules with more than {(MAX_LOOSE if loose else TILE) + 1} atoms are not supported '
                             f'(got {int(n.max())})')
        if n.min(initial=1) < 1:
            raise ValueError('every molecule needs at least one atom')
        bb, ii = np.nonzero(m)                                   # (b, i) lexicographic == packed order
        self.Nn = int(bb.shape[0])
        node_mol = bb.astype(np.int32)
        node_dense = (bb * N + ii).astype(np.int32)
        mol_start = np.zeros(B + 1, dtype=np.int64)
        np.cumsum(n, out=mol_start[1:])
        # ---- group -> tile assignment: best fit over a window of open tiles (groups of a molecule share gl, so they
        # are placed in runs).  Molecules stay roughly in order (locality of the per-atom gathers), but a tile that a
        # molecule's groups leave half empty is topped up by groups of following, differently sized molecules
        # (row utilisation 94 -> 97 % on the QM9 histogram, 81 -> 94 % on GEOM-Drugs n <= 80).
        gl_node = (n[bb] - 1).astype(np.int64)                   # group length per packed atom
        g_tile = np.zeros(self.Nn, dtype=np.int64)
        g_start = np.zeros(self.Nn, dtype=np.int64)
        g_idx = np.zeros(self.Nn, dtype=np.int64)
        open_tiles = []                                          # [tile id, rows used, groups]
        ngroups = []
        n_tiles = 0
        if loose:
            g_off = np.zeros(self.Nn + 1, dtype=np.int64)
            np.cumsum(gl_node, out=g_off[1:])
            g_tile[:] = g_off[:-1] // TILE
            g_start[:] = g_off[:-1] % TILE                       # the group continues into the next tile(s)
            n_tiles = int((g_off[-1] + TILE - 1) // TILE)
            ngroups = [0] * n_tiles
        for b in range(0 if not loose else B, B):
            gl = int(n[b]) - 1
            if gl <= 0:
                continue
            v = int(mol_start[b])
            left = int(n[b])
            while left:
                t, k = None, 0
                for c in open_tiles:                             # the fullest tile that still takes a group
                    if c[1] + gl <= TILE and c[2] < MAX_GROUPS and (t is None or c[1] > t[1]):
                        t = c
                if t is not None:
                    k = min(left, (TILE - t[1]) // gl, MAX_GROUPS - t[2])
                else:
                    if len(open_tiles) == WINDOW:                # close the fullest open tile
                        open_tiles.pop(max(range(WINDOW), key=lambda i: open_tiles[i][1]))
                    t = [n_tiles, 0, 0]
                    open_tiles.append(t)
                    ngroups.append(0)
                    n_tiles += 1
                    k = min(left, TILE // gl, MAX_GROUPS)
                idx = np.arange(k)
                g_tile[v:v + k] = t[0]
                g_start[v:v + k] = t[1] + idx * gl
                g_idx[v:v + k] = t[2] + idx
                t[1] += k * gl
                t[2] += k
                ngroups[t[0]] = t[2]
                v += k
                left -= k
        if n_tiles == 0:
            n_tiles, ngroups = 1, [0]
        self.n_tiles = n_tiles
        R = self.n_tiles * TILE
        row_g = np.full(R, -1, dtype=np.int32)
        row_j = np.full(R, -1, dtype=np.int32)
        row_meta = np.zeros(R, dtype=np.uint32)
        row_mol = np.zeros(R, dtype=np.int32)
        has = gl_node > 0
        v_ids = np.nonzero(has)[0]
        gl_v = gl_node[v_ids]
        tot = int(gl_v.sum())
        self.n_edges = tot
        if tot:
            grp = np.repeat(np.arange(v_ids.shape[0]), gl_v)     # group id per row
            first = np.zeros(v_ids.shape[0], dtype=np.int64)
            np.cumsum(gl_v[:-1], out=first[1:])
            k = np.arange(tot) - first[grp]                      # partner ordinal inside the group
            v = v_ids[grp]
            s = mol_start[bb[v]]
            j = s + k + (k >= (v - s))                           # skip the atom itself
            rows = g_tile[v] * TILE + g_start[v] + k
            row_g[rows] = v
            row_j[rows] = j
            if not loose:                                        # (start row, length, group index) inside the tile
                row_meta[rows] = (g_start[v] | (gl_node[v] << 8) | (g_idx[v] << 16)).astype(np.uint32)
            row_mol[rows] = node_mol[v]
        self.utilization = tot / float(R)
        dev = device if device is not None else node_mask.device
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.node_mol = t(node_mol)
        self.node_dense = t(node_dense)
        self.mol_start = t(mol_start.astype(np.int32))
        self.row_g = t(row_g)
        self.row_j = t(row_j)
        self.row_meta = t(row_meta.view(np.int32))
        self.tile_ngroups = t(np.asarray(ngroups, dtype=np.int32))
        self.row_mol = t(row_mol)
        self.grp_row0 = t((g_tile * TILE + g_start).astype(np.int32))      # first edge row / partner count per packed atom
        self.grp_len = t(gl_node.astype(np.int32))
        self.max_group = int(gl_node.max(initial=0))
        self.device = dev
        # ---- pair rows: the edge STATE (e32 / e16 / eh / adjacency bits) is symmetric in (g, j), so
        # it is stored and updated once per unordered pair {i < j}: pairs of a molecule in (i, j) lexicographic order,
        # molecules back to back, 128 rows per tile with no grouping constraint (the kernels that own the state --
        # edge_embed, edge_update, edge_head -- are row-local).  The directed rows above point at their pair row
        # through row_pair; attention and the coordinate update gather the fp16 operand rows through it.
        npair = n * (n - 1) // 2
        pair_start = np.zeros(B + 1, dtype=np.int64)
        np.cumsum(npair, out=pair_start[1:])
        P = int(pair_start[-1])
        self.n_pairs = P
        self.n_pair_tiles = max(1, (P + TILE - 1) // TILE)
        RP = self.n_pair_tiles * TILE
        pair_i = np.full(RP, -1, dtype=np.int32)
        pair_j = np.full(RP, -1, dtype=np.int32)
        pair_mol = np.zeros(RP, dtype=np.int32)
        if P:
            pm = np.repeat(np.arange(B), npair)                  # molecule of every pair row
            q = np.arange(P) - pair_start[pm]                    # pair ordinal inside the molecule
            nn_ = n[pm].astype(np.float64)
            # q = i n - i (i + 1) / 2 + (j - i - 1)  ->  i = floor(((2 n - 1) - sqrt((2 n - 1)^2 - 8 q)) / 2)
            ii_ = np.floor(((2 * nn_ - 1) - np.sqrt((2 * nn_ - 1) ** 2 - 8 * q)) / 2).astype(np.int64)
            first_of = lambda i_: i_ * n[pm] - i_ * (i_ + 1) // 2
            ii_ = np.where(first_of(ii_) > q, ii_ - 1, ii_)      # guard the floating-point floor
            ii_ = np.where(first_of(ii_ + 1) <= q, ii_ + 1, ii_)
            jj_ = q - first_of(ii_) + ii_ + 1
            assert ((ii_ >= 0) & (jj_ > ii_) & (jj_ < n[pm])).all()
            pair_i[:P] = mol_start[pm] + ii_
            pair_j[:P] = mol_start[pm] + jj_
            pair_mol[:P] = pm
        row_pair = np.full(R, -1, dtype=np.int32)
        if tot:
            vg, vj = row_g[row_g >= 0].astype(np.int64), row_j[row_g >= 0].astype(np.int64)
            mb = node_mol[vg].astype(np.int64)
            lo = np.minimum(vg, vj) - mol_start[mb]
            hi = np.maximum(vg, vj) - mol_start[mb]
            row_pair[row_g >= 0] = (pair_start[mb] + lo * n[mb] - lo * (lo + 1) // 2 + (hi - lo - 1)).astype(np.int32)
        self.row_pair = t(row_pair)
        self.pair_i, self.pair_j, self.pair_mol = t(pair_i), t(pair_j), t(pair_mol)
        self.pair_meta = t(np.zeros(RP, dtype=np.int32))
        self.pair_ngroups = t(np.zeros(self.n_pair_tiles, dtype=np.int32))

    # ---- pair-row helpers (tests) ------------------------------------------------------------------
    def pairs_to_dense(self, rows: torch.Tensor) -> torch.Tensor:
        """[n_pair_tiles * 128, C] pair rows -> dense symmetric [B, N, N, C] (diagonal and padding 0)."""
        B, N = self.B, self.N
        C = rows.shape[-1]
        out = torch.zeros(B * N * N, C, dtype=rows.dtype, device=rows.device)
        valid = self.pair_i >= 0
        i = self.node_dense[self.pair_i[valid].long()].long()
        j = self.node_dense[self.pair_j[valid].long()].long()
        b = i // N
        ii, jj = i % N, j % N
        out[b * N * N + ii * N + jj] = rows[valid]
        out[b * N * N + jj * N + ii] = rows[valid]
        return out.reshape(B, N, N, C)

    # ---- helpers used by tests (pure index bookkeeping) -------------------------------------------
    def dense_to_rows(self, dense_bnn: torch.Tensor, group_first=True) -> torch.Tensor:
        """Gather a dense [B,N,N,C] tensor into tile-row order [R, C] (padding rows = 0).
        group_first: row (g, j) reads dense[b, i_g, i_j]; else dense[b, i_j, i_g]."""
        B, N = self.B, self.N
        C = dense_bnn.shape[-1]
        flat = dense_bnn.reshape(B * N * N, C)
        valid = self.row_g >= 0
        g = self.node_dense[self.row_g.clamp(min=0).long()].long()
        j = self.node_dense[self.row_j.clamp(min=0).long()].long()
        b = g // N
        ig, ij = g % N, j % N
        idx = b * N * N + (ig * N + ij if group_first else ij * N + ig)
        out = flat[idx] * valid[:, None].to(flat.dtype)
        return out

    def rows_to_dense(self, rows: torch.Tensor, group_first=True) -> torch.Tensor:
        B, N = self.B, self.N
        C = rows.shape[-1]
        out = torch.zeros(B * N * N, C, dtype=rows.dtype, device=rows.device)
        valid = self.row_g >= 0
        g = self.node_dense[self.row_g[valid].long()].long()
        j = self.node_dense[self.row_j[valid].long()].long()
        b = g // N
        ig, ij = g % N, j % N
        idx = b * N * N + (ig * N + ij if group_first else ij * N + ig)
        out[idx] = rows[valid]
        return out.reshape(B, N, N, C)
