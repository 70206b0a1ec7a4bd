"""Algorithmic FLOP / byte model of the DGT denoiser, used by bench.py for the
roofline fractions.  Real atoms and real directed edges only; elementwise / transcendental work is not
counted; mm(i, o) = 2*i*o FLOP per row."""
from __future__ import annotations


def mm(i, o):
    return 2 * i * o


# Kernels that run once per UNORDERED pair (the edge state is symmetric): their per-row figures below
# apply to edges / 2 rows.  The whole-step model (flops_alg / bytes_alg) stays the reference's per-directed-edge count.
PAIR_KERNELS = frozenset({
    'jodo_edge_embed', 'jodo_edge_update', 'jodo_edge_head',
    'jodo_imglinear:emb', 'jodo_imglinear:g01', 'jodo_imglinear:ff3', 'jodo_imglinear:ff4', 'jodo_imglinear:equi_in',
    'jodo_wide_ln:e2', 'jodo_wide_ln:e1', 'jodo_wide_dist', 'jodo_wide_put', 'jodo_wide_edge_ffn'})


def per_edge_kernel_flops(d):
    """FLOP per real directed edge for ONE launch of each edge-tile kernel."""
    D, ed, qk, r, L, ce, ch, X = d.D, d.ed, d.qk, d.r, d.L, d.ce, d.ch, d.X
    return {
        'jodo_edge_embed': mm(2 * ch + ed, ed),
        'jodo_attn': mm(2 * ed, ed) + mm(ed, qk) + mm(ed, D) + 3 * qk + 2 * D,
        'jodo_edge_update': mm(ed, r * ed) + mm(r * ed, ed) + mm(ed, ce),
        'jodo_equi': mm(2 * ed, D) + mm(D, D) + mm(D, 1 + X),
        'jodo_edge_head': 2 * (mm(ce * L + ed, ed) + mm(ed, ed // 2)) + mm(ed // 2, 1) + mm(ed // 2, ch - 1),
    }


def per_edge_kernel_bytes(d):
    """Algorithmic HBM bytes per real directed edge for one launch of each edge-tile kernel (fp32
    edge state crossing HBM between kernels; per-atom operands are L2-resident and not counted)."""
    ed, ce, ch, L = d.ed, d.ce, d.ch, d.L
    return {
        'jodo_edge_embed': 4 * (2 * ch + ed) + 1,
        'jodo_attn': 4 * ed + 1,
        'jodo_edge_update': 4 * (2 * ed + ce),
        'jodo_equi': 4 * ed + 1,
        'jodo_edge_head': 4 * (ed + ce * L + ch),
    }


def wide_kernel_flops(d):
    """Wide path (nf = 384): algorithmic FLOP per real directed edge for one launch of each per-edge GEMM."""
    D, ed, qk, r, X = d.D, d.ed, d.qk, d.r, d.X
    return {
        'jodo_imglinear:emb': mm(2 * ed, ed),
        'jodo_imglinear:g01': mm(ed, qk) + mm(ed, D),
        'jodo_imglinear:ff3': mm(ed, r * ed),
        'jodo_imglinear:ff4': mm(r * ed, ed),
        'jodo_imglinear:equi_in': mm(2 * ed, D),
        'jodo_imglinear:c0': mm(D, D),
        'jodo_imglinear:c2': mm(D, 1 + X),
        'jodo_wide_edge_ffn': mm(ed, r * ed) + mm(r * ed, ed),
    }


def wide_kernel_bytes(d):
    """Wide path: algorithmic HBM bytes per real directed edge for one launch of each row kernel -- the rows it must
    read and write once (per-atom operands and per-molecule tables are L2-resident and not counted)."""
    D, ed, qk = d.D, d.ed, d.qk
    return {
        'jodo_wide_ln:equi': 2 * D + 2 * D,                 # fp16 pre-LayerNorm rows (the pair's) in, fp16 operand image out
        'jodo_wide_ln:e2': 4 * ed + 4 * ed + 2 * ed,        # fp32 edge state in, fp32 e2 + fp16 image out
        'jodo_wide_ln:e1': 4 * ed + 2 * ed,
        'jodo_wide_attn': 2 * (qk + D) + 1,                 # tanh(lin_edge0 | lin_edge1) rows + adjacency bits
        'jodo_wide_dist': 2 * 2 * ed,
        'jodo_wide_put': 4 * ed + 3 * 2 * ed,
        'jodo_wide_edge_ffn': 4 * ed + 4 * ed + 2 * 2 * ed,  # fp32 edge state in and out, two fp16 copies of the new state
    }


def flops_alg(n, d):
    """F_alg(n) of SURVEY.md §8d: algorithmic FLOP of one denoiser call on one molecule of n atoms."""
    D, ed, T, L, r, qk, inn, ch, cn, ce, X = d.D, d.ed, d.T, d.L, d.r, d.qk, d.inn, d.ch, d.cn, d.ce, d.X
    per_mol = mm(17, T) + mm(T, T) + mm(T, 2)
    if d.cond_ch:
        per_mol += d.cond_ch * (mm(1, D) + mm(D, D)) + mm(d.cond_ch * D, T)
    per_mol_layer = mm(T, 6 * D) + mm(T, 6 * ed) + mm(T, 2 * D) + mm(T, 2)
    per_node = mm(2 * inn, D) + mm(cn * L + D, D) + mm(D, D // 2) + mm(D // 2, inn)
    per_node_layer = 2 * mm(D, qk) + mm(D, D) + mm(D, r * D) + mm(r * D, D) + mm(D, cn) + mm(D, ed) + 2 * mm(D, D)
    k = per_edge_kernel_flops(d)
    per_edge = k['jodo_edge_embed'] + k['jodo_edge_head']
    per_edge_layer = k['jodo_attn'] + k['jodo_edge_update'] + k['jodo_equi']
    return per_mol + L * per_mol_layer + n * (per_node + L * per_node_layer) + n * (n - 1) * (per_edge + L * per_edge_layer)


def bytes_alg(n, d):
    """bytes_alg(n) of SURVEY.md §8d: one kernel per block, fp32 state crossing HBM between blocks."""
    D, ed, L, inn, ch, cn, ce = d.D, d.ed, d.L, d.inn, d.ch, d.cn, d.ce
    io = 4 * (3 * n * (3 + inn) + 3 * n * n * ch + n + n * n)
    per_layer = 4 * (2 * n * D + 2 * n * n * ed + 6 * n + n * cn + n * n * ce)
    head = 4 * (n * (cn * L + D) + n * n * (ce * L + ed))
    return io + L * per_layer + head


def batch_totals(n_nodes, d):
    ns = [int(v) for v in n_nodes]
    return dict(flops=sum(flops_alg(n, d) for n in ns), bytes=sum(bytes_alg(n, d) for n in ns),
                edges=sum(n * (n - 1) for n in ns), atoms=sum(ns))
