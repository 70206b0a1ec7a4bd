"""Parameter tree of the DGT denoiser: names, shapes and registration ORDER.

The drop-in contract is that a checkpoint written by the reference loads
with ``load_state_dict(strict=True)`` (reference utils.py:15-19) and that EMA shadow
parameters, which are matched to ``model.parameters()`` by position
(reference models/ema.py:20-21,52-55), land on the right tensors.  So the names, shapes
and order below restate what ``DGT_concat.__init__`` / ``Cond_DGT_concat.__init__``
(reference models/mol_gnn.py:414-489, 601-684), ``EquivariantMixBlock.__init__``
(:214-260), ``MultiCondEquiUpdate.__init__`` (:54-69), ``TransMixLayer.__init__``
(models/layers.py:98-120) and ``CondGaussianLayer.__init__`` (:316-326) register.
tests/golden/param_tree_*.json holds the reference's own list for comparison.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
from torch import nn


@dataclass(frozen=True)
class Dims:
    """Derived sizes (notation of SURVEY.md §8)."""
    D: int          # model.nf
    ed: int         # D // 4
    T: int          # 4 * D
    L: int          # n_layers
    r: int          # mlp_ratio
    H: int          # n_heads
    X: int          # n_extra_heads
    S: int          # H - X
    sc: int         # D // S
    qk: int         # S * sc
    C: int          # D // H
    inn: int        # atom_types + include_fc_charge
    ch: int         # edge_ch
    cn: int         # 2D // L
    ce: int         # 2ed // L
    cond_ch: int    # 0 for the unconditional model
    two_d: bool = False   # DGT_concat_2D: no coordinates (xh = atom features only)

    @property
    def node_cat(self):
        return self.cn * self.L + self.D

    @property
    def edge_cat(self):
        return self.ce * self.L + self.ed


VARIANTS = ('uncond', 'cond', '2d', 'sim')
_NAME_TO_VARIANT = {'DGT_concat': 'uncond', 'cond_DGT_concat': 'cond', 'DGT_concat_2D': '2d', 'DGT_concat_sim': 'sim'}


def variant_of(config, variant=None):
    """Which reference class the config selects.  A model class passes its own ``variant``; stand-alone callers
    (tests, tools) get it from ``config.model.name``, with any registry suffix such as ``_b200`` ignored
    (reference models/utils.py:24-25 resolves the class by that name)."""
    if variant is not None:
        if variant not in VARIANTS:
            raise ValueError(f'unknown variant {variant!r}')
        return variant
    name = str(config.model.name)
    for base in sorted(_NAME_TO_VARIANT, key=len, reverse=True):
        if name == base or (name.startswith(base + '_') and name[len(base) + 1:] not in ('2D', 'sim')):
            return _NAME_TO_VARIANT[base]
    raise ValueError(f'unsupported model.name {name!r}')


def dims_from_config(config, variant=None) -> Dims:
    m, d = config.model, config.data
    variant = variant_of(config, variant)
    D = int(m.nf)
    H = int(m.n_heads)
    # DGT_concat_sim (reference models/mol_gnn.py:949, EquivariantBlock :97 + Trans_Layer, models/layers.py:13): every head is a
    # learned head of D / H channels, no adjacency heads, coord_mlp.2 has one output -- whatever config.model.n_extra_heads says
    X = 0 if variant == 'sim' else int(m.n_extra_heads)
    S = H - X
    L = int(m.n_layers)
    ed = D // 4
    cond = variant == 'cond'
    two_d = variant == '2d'
    if two_d and int(getattr(m, 'time_dim', 4 * D)) != 4 * D:
        raise NotImplementedError('DGT_concat_2D: model.time_dim must be 4 * model.nf')
    return Dims(two_d=two_d, D=D, ed=ed, T=4 * D, L=L, r=int(m.mlp_ratio), H=H, X=X, S=S, sc=D // S,
                qk=S * (D // S), C=D // H, inn=int(d.atom_types) + int(m.include_fc_charge),
                ch=int(m.edge_ch), cn=(2 * D) // L, ce=(2 * ed) // L,
                cond_ch=int(m.cond_ch) if cond else 0)


def check_supported(config, variant=None):
    """Variants of the reference model this implementation covers.  The variant is the
    CLASS's (the registry may know it under any name, e.g. ``DGT_concat_b200``), not ``config.model.name``."""
    m = config.model
    variant = variant_of(config, variant)
    two_d = variant == '2d'
    need = dict(cond_time=True, softmax_inf=True, pred_data=True)
    if not two_d:
        need.update(dist_gbf=True, gbf_name='CondGaussianLayer', CoM=True)
    for k, v in need.items():
        if getattr(m, k) != v:
            raise ValueError(f'unsupported config.model.{k}={getattr(m, k)!r} (hot path covers {v!r})')
    if getattr(m, 'trans_name', 'TransMixLayer') != 'TransMixLayer':
        raise ValueError('unsupported trans_name')
    if variant != 'sim' and int(m.n_extra_heads) != (1 if two_d else 2):
        raise ValueError('hot path covers n_extra_heads == 2 (1 for DGT_concat_2D), as in all reference configs')


def _lin(name, out_f, in_f, bias=True):
    out = [(f'{name}.weight', (out_f, in_f))]
    if bias:
        out.append((f'{name}.bias', (out_f,)))
    return out


def _gbf(name, d: Dims):
    return [(f'{name}.means.weight', (1, d.ed - 1)), (f'{name}.stds.weight', (1, d.ed - 1))] + \
        _lin(f'{name}.time_mlp.1', 2, d.T)


def param_spec_2d(d: Dims):
    """DGT_concat_2D.__init__ / EquivariantMixBlock_2D.__init__ (reference models/mol_gnn.py:801-866, 328-363)."""
    D, ed, T = d.D, d.ed, d.T
    spec = _lin('node_emb', D, 2 * d.inn) + _lin('edge_emb', ed, 2 * d.ch)
    for i in range(d.L):
        b = f'e_block_{i}'
        spec += _lin(f'{b}.node2edge_lin', ed, D)
        spec += _lin(f'{b}.attn_mpnn.lin_key', d.qk, D)
        spec += _lin(f'{b}.attn_mpnn.lin_query', d.qk, D)
        spec += _lin(f'{b}.attn_mpnn.lin_value', D, D)
        spec += _lin(f'{b}.attn_mpnn.lin_edge0', d.qk, ed, bias=False)
        spec += _lin(f'{b}.attn_mpnn.lin_edge1', D, ed, bias=False)
        spec += _lin(f'{b}.ff_linear1', D * d.r, D)
        spec += _lin(f'{b}.ff_linear2', D, D * d.r)
        spec += _lin(f'{b}.ff_linear3', ed * d.r, ed)
        spec += _lin(f'{b}.ff_linear4', ed, ed * d.r)
        spec += _lin(f'{b}.node_time_mlp.1', 6 * D, T)
        spec += _lin(f'{b}.edge_time_mlp.1', 6 * ed, T)
        spec += _lin(f'node_{i}', d.cn, D)
        spec += _lin(f'edge_{i}', d.ce, ed)
    spec += _lin('node_pred_mlp.0', D, d.node_cat) + _lin('node_pred_mlp.2', D // 2, D) + \
        _lin('node_pred_mlp.4', d.inn, D // 2)
    spec += _lin('edge_type_mlp.0', ed, d.edge_cat) + _lin('edge_type_mlp.2', ed // 2, ed) + \
        _lin('edge_type_mlp.4', d.ch - 1, ed // 2)
    spec += _lin('edge_exist_mlp.0', ed, d.edge_cat) + _lin('edge_exist_mlp.2', ed // 2, ed) + \
        _lin('edge_exist_mlp.4', 1, ed // 2)
    spec += [('time_mlp.0.weights', (8,))]
    spec += _lin('time_mlp.1', T, 17) + _lin('time_mlp.3', T, T)
    return spec


def param_spec(config, variant=None):
    """Ordered [(name, shape)] exactly as the reference module registers them."""
    d = dims_from_config(config, variant)
    if d.two_d:
        return param_spec_2d(d)
    D, ed, T = d.D, d.ed, d.T
    spec = []
    spec += _lin('node_emb', D, 2 * d.inn)
    spec += _lin('edge_emb', ed, 2 * d.ch + ed)
    spec += _gbf('dist_layer', d)
    for i in range(d.L):
        b = f'e_block_{i}'
        spec += _lin(f'{b}.edge_emb', ed, 2 * ed)
        spec += _lin(f'{b}.node2edge_lin', ed, D)
        spec += _lin(f'{b}.attn_mpnn.lin_key', d.qk, D)
        spec += _lin(f'{b}.attn_mpnn.lin_query', d.qk, D)
        spec += _lin(f'{b}.attn_mpnn.lin_value', D, D)
        spec += _lin(f'{b}.attn_mpnn.lin_edge0', d.qk, ed, bias=False)
        spec += _lin(f'{b}.attn_mpnn.lin_edge1', D, ed, bias=False)
        spec += _lin(f'{b}.ff_linear1', D * d.r, D)
        spec += _lin(f'{b}.ff_linear2', D, D * d.r)
        spec += _lin(f'{b}.ff_linear3', ed * d.r, ed)
        spec += _lin(f'{b}.ff_linear4', ed, ed * d.r)
        spec += [(f'{b}.equi_update.coord_norm.scale', (1,))]
        spec += _lin(f'{b}.equi_update.time_mlp.1', 2 * D, T)
        spec += _lin(f'{b}.equi_update.input_lin', D, 2 * D + 2 * ed)
        spec += _lin(f'{b}.equi_update.coord_mlp.0', D, D)
        spec += _lin(f'{b}.equi_update.coord_mlp.2', 1 + d.X, D, bias=False)
        spec += _lin(f'{b}.node_time_mlp.1', 6 * D, T)
        spec += _lin(f'{b}.edge_time_mlp.1', 6 * ed, T)
        spec += _gbf(f'{b}.dist_layer', d)
        spec += _lin(f'node_{i}', d.cn, D)
        spec += _lin(f'edge_{i}', d.ce, ed)
    spec += _lin('node_pred_mlp.0', D, d.node_cat) + _lin('node_pred_mlp.2', D // 2, D) + \
        _lin('node_pred_mlp.4', d.inn, D // 2)
    spec += _lin('edge_type_mlp.0', ed, d.edge_cat) + _lin('edge_type_mlp.2', ed // 2, ed) + \
        _lin('edge_type_mlp.4', d.ch - 1, ed // 2)
    spec += _lin('edge_exist_mlp.0', ed, d.edge_cat) + _lin('edge_exist_mlp.2', ed // 2, ed) + \
        _lin('edge_exist_mlp.4', 1, ed // 2)
    spec += [('time_mlp.0.weights', (8,))]
    spec += _lin('time_mlp.1', T, 17) + _lin('time_mlp.3', T, T)
    if d.cond_ch:
        spec += _lin('cond_mlp.0', D, 1) + _lin('cond_mlp.2', D, D) + _lin('cond_lin', T, d.cond_ch * D)
    return spec


def build_param_tree(root: nn.Module, spec):
    """Register every (dotted name, shape) of `spec` on `root`, creating bare nn.Module
    containers for the intermediate path components, in order."""
    for name, shape in spec:
        mod = root
        parts = name.split('.')
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, nn.Module())
            mod = mod._modules[p]
        mod.register_parameter(parts[-1], nn.Parameter(torch.empty(shape, dtype=torch.float32)))


def synth_state_dict(spec, seed=0, gain=1.0, perturb=False, dtype=torch.float32):
    """Deterministic weights from (name, shape, seed) only, so the reference module (here) and
    our module (on the GPU box, where the reference does not exist) hold identical values.

    Linear weights/biases ~ U(+-gain/sqrt(fan_in)) as nn.Linear's default init does; GBF
    means/stds ~ U(0,3) (reference layers.py:325-326); Fourier weights ~ N(0,1) (:281);
    coord_norm.scale = 1e-2 (mol_gnn.py:56) or, with perturb=True, 0.3 so that the coordinate
    branch is not numerically inert (SURVEY.md §8c "Weights").
    """
    out = {}
    for i, (name, shape) in enumerate(spec):
        g = torch.Generator().manual_seed(1_000_003 * (seed + 1) + i)
        if name.endswith('coord_norm.scale'):
            v = torch.full(shape, 0.3 if perturb else 1e-2, dtype=torch.float64)
        elif name.endswith('means.weight') or name.endswith('stds.weight'):
            v = torch.rand(shape, generator=g, dtype=torch.float64) * 3.0
        elif name.endswith('time_mlp.0.weights'):
            v = torch.randn(shape, generator=g, dtype=torch.float64)
        else:
            fan_in = shape[1] if len(shape) == 2 else None
            if fan_in is None:  # bias: fan_in of its weight = previous entry
                fan_in = spec[i - 1][1][1]
            bound = gain / math.sqrt(fan_in)
            v = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound
        out[name] = v.to(dtype)
    return out
