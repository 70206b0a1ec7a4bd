"""ORACLE (test infrastructure, not product code): CPU restatement of the reference DGT forward.

A dense, per-molecule restatement of ``DGT_concat.forward`` / ``Cond_DGT_concat.forward``
of GRAPH-0/JODO in plain torch CPU ops, any float dtype (fp64 for pinning, fp32 for timing).
It is the checker for the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs may import it; nothing under jodo_b200/ does.

Parity status: PINNED against outputs of the unmodified reference modules executed in the build
container through oracle/shim (generator: oracle/make_golden.py, fixtures: tests/golden/*.pt,
test: tests/test_oracle_golden.py).  The reference itself ships no tests or golden vectors
 and torch_geometric 2.1.0.post1 / torch_scatter 2.0.9 are not installable
offline, so the pin is "reference code + shimmed PyG/scatter primitives", see oracle/README.md.

Each function cites the reference lines it follows (paths relative to the reference root).
The sparse edge list of the reference (all ordered pairs r != c of real atoms, models/mol_gnn.py:
512-514) becomes a dense [B,N,N] grid with edge mask `em`; index convention: first grid axis is
r = edge_index[0] ("row", PyG source j), second is c = edge_index[1] ("col", PyG target i).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _lin(sd, name, x):
    w = sd[name + '.weight'].to(x.dtype)
    b = sd.get(name + '.bias')
    return F.linear(x, w, None if b is None else b.to(x.dtype))


def _ln(x):
    # nn.LayerNorm(elementwise_affine=False, eps=1e-6): models/mol_gnn.py:234-235,240,245,64
    return F.layer_norm(x, (x.shape[-1],), eps=1e-6)


def _modulate(x, shift, scale):
    # models/mol_gnn.py:12-13
    return x * (1 + scale) + shift


def remove_mean_with_mask(x, node_mask):
    # models/utils.py:38-45
    n = node_mask.sum(1, keepdim=True)
    return x - (x.sum(1, keepdim=True) / n) * node_mask


def time_embedding(sd, noise_level, context=None):
    """LearnedSinusodialposEmb + time_mlp (models/layers.py:283-288, models/mol_gnn.py:481-489,534)
    and, for the conditional model, cond_mlp/cond_lin (:679-684, 728-734)."""
    x = noise_level.unsqueeze(-1)
    w = sd['time_mlp.0.weights'].to(x.dtype)
    freqs = x * w.unsqueeze(0) * 2 * math.pi
    four = torch.cat((x, freqs.sin(), freqs.cos()), dim=-1)
    temb = _lin(sd, 'time_mlp.3', F.gelu(_lin(sd, 'time_mlp.1', four)))
    if 'cond_lin.weight' in sd:
        ctx = context.unsqueeze(-1)                                   # [B, cond_ch, 1]
        ctx = _lin(sd, 'cond_mlp.2', F.gelu(_lin(sd, 'cond_mlp.0', ctx)))
        temb = temb + _lin(sd, 'cond_lin', ctx.reshape(ctx.shape[0], -1))
    return temb


def cond_gbf(sd, prefix, d, st):
    """CondGaussianLayer.forward + gaussian (models/layers.py:291-295,328-334).
    d: [B,N,N,1]; st = SiLU(temb) [B,T].  scale comes first in the chunk (:330)."""
    ss = _lin(sd, prefix + '.time_mlp.1', st)                         # [B,2]
    scale, shift = ss[:, 0], ss[:, 1]
    x = d * (scale[:, None, None, None] + 1) + shift[:, None, None, None]
    mean = sd[prefix + '.means.weight'].float().view(-1)              # cast to fp32: :332-333
    std = sd[prefix + '.stds.weight'].float().view(-1).abs() + 1e-5
    # mean/std stay fp32 tensors and a*std is an fp32 product, as in the TorchScript `gaussian`
    # (type promotion does the rest); this only matters when pinning in fp64.
    a = (2 * 3.14159) ** 0.5                                          # literal pi: :293
    g = torch.exp(-0.5 * (((x - mean) / std) ** 2)) / (a * std)
    return torch.cat([x, g], dim=-1)


def trans_mix(sd, prefix, hn, en, extra, em, dims, trace=None):
    """TransMixLayer.forward/message (models/layers.py:131-186) on a dense grid.
    hn [B,N,D], en [B,N,N,ed], extra [B,N,N,X], em [B,N,N,1] -> [B,N,D]."""
    B, N, D = hn.shape
    S, sc, H, C = dims['S'], dims['sc'], dims['H'], dims['C']
    q = _lin(sd, prefix + '.lin_query', hn).reshape(B, N, S, sc)      # indexed by target c (:147)
    k = _lin(sd, prefix + '.lin_key', hn).reshape(B, N, S, sc)        # indexed by source r (:148)
    v = _lin(sd, prefix + '.lin_value', hn).reshape(B, N, H, C)       # source r (:149)
    g0 = torch.tanh(_lin(sd, prefix + '.lin_edge0', en)).reshape(B, N, N, S, sc)   # :165-166
    a = (q[:, None, :, :, :] * k[:, :, None, :, :] * g0).sum(-1) / math.sqrt(C)    # :167 (sqrt(out_channels))
    xi = torch.where(extra == 0, torch.full_like(extra, -1e10), extra)             # :171-173
    A = torch.cat([xi, a], dim=-1)                                    # extra heads first (:174)
    valid = em > 0
    Am = torch.where(valid, A, torch.full_like(A, float('-inf')))
    mx = Am.max(dim=1, keepdim=True).values                           # over sources r, per target c
    mx = torch.where(torch.isfinite(mx), mx, torch.zeros_like(mx))
    ex = torch.where(valid, (A - mx).exp(), torch.zeros_like(A))
    al = ex / (ex.sum(dim=1, keepdim=True) + 1e-16)                   # PyG softmax (:178)
    g1 = torch.tanh(_lin(sd, prefix + '.lin_edge1', en)).reshape(B, N, N, H, C)    # :183
    msg = v[:, :, None, :, :] * g1 * al[..., None]                    # :182-184
    if trace is not None:
        trace.update(q=q.reshape(B, N, -1), k=k.reshape(B, N, -1), v=v.reshape(B, N, -1), alpha=al)
    return msg.sum(dim=1).reshape(B, N, D)                            # aggr='add' onto target c (:101)


def equi_update(sd, prefix, hout, pos, eout, df, st, extra, em):
    """MultiCondEquiUpdate.forward (models/mol_gnn.py:71-94), CoorsNorm (models/layers.py:344-347)."""
    B, N, D = hout.shape
    u = torch.cat([hout[:, :, None, :].expand(B, N, N, D), hout[:, None, :, :].expand(B, N, N, D),
                   eout, df], dim=-1)                                 # cat[h[row], h[col], e, dist] (:73)
    diff = pos[:, :, None, :] - pos[:, None, :, :]                    # pos[row] - pos[col] (:74)
    nrm = diff.norm(dim=-1, keepdim=True)
    dl = diff / nrm.clamp(min=1e-8) * sd[prefix + '.coord_norm.scale'].to(pos.dtype)
    ss = _lin(sd, prefix + '.time_mlp.1', st)
    shift, scale = ss.chunk(2, dim=1)                                 # :78
    inv = _modulate(_ln(_lin(sd, prefix + '.input_lin', u)), shift[:, None, None, :], scale[:, None, None, :])
    inv = torch.tanh(F.linear(F.silu(_lin(sd, prefix + '.coord_mlp.0', inv)),
                              sd[prefix + '.coord_mlp.2.weight'].to(inv.dtype)))    # :82
    adjs = torch.cat([torch.ones_like(inv[..., :1]), extra], dim=-1)               # :85-86 (CondEquiUpdate, :16-49: just the 1)
    inv = (inv * adjs).mean(-1, keepdim=True)                         # :87
    return pos + (dl * inv * em).sum(dim=2)                           # scatter-add on row (:90-92)


def mix_block(sd, prefix, pos, h, e, extra, m, em, st, dims, trace=None):
    """EquivariantMixBlock.forward (models/mol_gnn.py:270-322)."""
    diff = pos[:, :, None, :] - pos[:, None, :, :]
    d = (diff ** 2).sum(-1, keepdim=True)                             # coord2dist (models/utils.py:122-126)
    df = cond_gbf(sd, prefix + '.dist_layer', d, st)                  # :285-286
    e1 = _lin(sd, prefix + '.edge_emb', torch.cat([df, e], dim=-1))   # :287
    nt = _lin(sd, prefix + '.node_time_mlp.1', st)                    # :291-292
    et = _lin(sd, prefix + '.edge_time_mlp.1', st)                    # :293-294
    nsm, ncm, ngm, nsf, ncf, ngf = [t[:, None, :] for t in nt.chunk(6, dim=1)]
    esm, ecm, egm, esf, ecf, egf = [t[:, None, None, :] for t in et.chunk(6, dim=1)]
    hn = _modulate(_ln(h), nsm, ncm)                                  # :296
    en = _modulate(_ln(e1), esm, ecm)                                 # :297
    hnode = trans_mix(sd, prefix + '.attn_mpnn', hn, en, extra, em, dims, trace)   # :303
    hedge = _lin(sd, prefix + '.node2edge_lin', hnode[:, :, None, :] + hnode[:, None, :, :])  # :304-305
    h1 = h + ngm * hnode                                              # :307
    h2 = _modulate(_ln(h1), nsf, ncf) * m                             # :308
    hout = (h2 + ngf * _lin(sd, prefix + '.ff_linear2', F.silu(_lin(sd, prefix + '.ff_linear1', h2)))) * m  # :310
    e2 = _modulate(_ln(e + egm * hedge), esf, ecf)                    # :313-314 (residual from block input e)
    eout = e2 + egf * _lin(sd, prefix + '.ff_linear4', F.silu(_lin(sd, prefix + '.ff_linear3', e2)))  # :316
    pos = equi_update(sd, prefix + '.equi_update', hout, pos, eout, df, st, extra, em)       # :320
    if trace is not None:
        trace.update(hn=hn, en=en, e1=e1, hnode=hnode, h2=h2, hedge=hedge, df=df)
    return hout, eout, pos


def mix_block_2d(sd, prefix, h, e, extra, m, em, st, dims):
    """EquivariantMixBlock_2D.forward (models/mol_gnn.py:372-408): the block without coordinates."""
    nt = _lin(sd, prefix + '.node_time_mlp.1', st)
    et = _lin(sd, prefix + '.edge_time_mlp.1', st)
    nsm, ncm, ngm, nsf, ncf, ngf = [t[:, None, :] for t in nt.chunk(6, dim=1)]
    esm, ecm, egm, esf, ecf, egf = [t[:, None, None, :] for t in et.chunk(6, dim=1)]
    hn = _modulate(_ln(h), nsm, ncm)                                  # :390
    en = _modulate(_ln(e), esm, ecm)                                  # :391 (norm1_edge on the block input itself)
    hnode = trans_mix(sd, prefix + '.attn_mpnn', hn, en, extra, em, dims)          # :394
    hedge = _lin(sd, prefix + '.node2edge_lin', hnode[:, :, None, :] + hnode[:, None, :, :])  # :395-396
    h1 = h + ngm * hnode
    h2 = _modulate(_ln(h1), nsf, ncf) * m                             # :399
    hout = (h2 + ngf * _lin(sd, prefix + '.ff_linear2', F.silu(_lin(sd, prefix + '.ff_linear1', h2)))) * m
    e2 = _modulate(_ln(e + egm * hedge), esf, ecf)                    # :402-403
    eout = e2 + egf * _lin(sd, prefix + '.ff_linear4', F.silu(_lin(sd, prefix + '.ff_linear3', e2)))
    return hout, eout


@torch.no_grad()
def dgt2d_forward(sd, config, t, xh, node_mask, edge_mask, context=None, edge_x=None, noise_level=None,
                  cond_x=None, cond_edge_x=None):
    """DGT_concat_2D.forward (models/mol_gnn.py:868-947): atom features and bonds only."""
    dims = dims_of(config)
    B, N, _ = xh.shape
    dt = xh.dtype
    m = node_mask.to(dt)
    em = edge_mask.reshape(B, N, N, 1).to(dt)
    if cond_x is None:                                                # :887-890
        cond_x = torch.zeros_like(xh)
        cond_edge_x = torch.zeros_like(edge_x)
        adj2d = torch.ones_like(em)
    else:                                                             # :892-895
        adj2d = (cond_edge_x[..., 0:1] >= config.model.edge_quan_th).to(dt)
    temb = time_embedding(sd, noise_level.to(dt))                     # :902-904
    st = F.silu(temb)
    extra = adj2d * em
    h = _lin(sd, 'node_emb', torch.cat([xh, cond_x], dim=-1))          # :898-899, 918
    e = _lin(sd, 'edge_emb', torch.cat([edge_x, cond_edge_x], dim=-1))             # :915, 919
    atom_hids, edge_hids = [h], [e]
    for i in range(dims['L']):                                        # :924-928
        h, e = mix_block_2d(sd, f'e_block_{i}', h, e, extra, m, em, st, dims)
        atom_hids.append(_lin(sd, f'node_{i}', h))
        edge_hids.append(_lin(sd, f'edge_{i}', e))
    atom_pred = _mlp3(sd, 'node_pred_mlp', torch.cat(atom_hids, dim=-1)) * m       # :933
    eh = torch.cat(edge_hids, dim=-1)
    ep = torch.cat([_mlp3(sd, 'edge_exist_mlp', eh), _mlp3(sd, 'edge_type_mlp', eh)], dim=-1) * em   # :934-939
    return atom_pred, 0.5 * (ep + ep.permute(0, 2, 1, 3))             # :940


def _mlp3(sd, name, x):
    return _lin(sd, name + '.4', F.silu(_lin(sd, name + '.2', F.silu(_lin(sd, name + '.0', x)))))


def dims_of(config):
    m, d = config.model, config.data
    D, H, X = int(m.nf), int(m.n_heads), int(m.n_extra_heads)
    if str(m.name).startswith('DGT_concat_sim'):          # Trans_Layer / CondEquiUpdate: no adjacency heads (mol_gnn.py:949, :97, :16)
        X = 0
    S = H - X
    return dict(D=D, H=H, X=X, S=S, sc=D // S, C=D // H, L=int(m.n_layers), ed=D // 4)


@torch.no_grad()
def dgt_forward(sd, config, t, xh, node_mask, edge_mask, context=None, edge_x=None, noise_level=None,
                cond_x=None, cond_edge_x=None, collect=None, trace=None):
    """DGT_concat.forward (models/mol_gnn.py:491-594) / Cond_DGT_concat.forward (:687-794).
    `collect`, if a list, receives (h, e, pos) after every block for stage-level debugging."""
    if str(config.model.name) == 'DGT_concat_2D':
        return dgt2d_forward(sd, config, t, xh, node_mask, edge_mask, context, edge_x, noise_level, cond_x, cond_edge_x)
    dims = dims_of(config)
    B, N, _ = xh.shape
    dt = xh.dtype
    m = node_mask.to(dt)
    em = edge_mask.reshape(B, N, N, 1).to(dt)
    pos = xh[..., :3].clone()
    h = xh[..., 3:]
    first = cond_x is None
    if first:                                                         # :517-520
        cond_x = torch.zeros_like(xh)
        cond_edge_x = torch.zeros_like(edge_x)
        adj2d = torch.ones_like(em)
    else:                                                             # :523-525
        adj2d = (cond_edge_x[..., 0:1] >= config.model.edge_quan_th).to(dt)
    cpos = cond_x[..., :3]
    h = torch.cat([h, cond_x[..., 3:]], dim=-1)                       # :528-530
    temb = time_embedding(sd, noise_level.to(dt), None if context is None else context.to(dt))
    st = F.silu(temb)
    cdiff = cpos[:, :, None, :] - cpos[:, None, :, :]
    d0 = (cdiff ** 2).sum(-1, keepdim=True)                           # coord2diff_adj (models/utils.py:111-119)
    adjsp = (d0 <= config.model.spatial_cut_off).to(dt)
    if float((d0 * em).sum()) == 0:                                   # batch-global branch (:544-545)
        dist0 = torch.zeros(B, N, N, dims['ed'], dtype=dt)
    else:
        dist0 = cond_gbf(sd, 'dist_layer', d0, st)                    # :547-548
    extra = torch.cat([adj2d, adjsp], dim=-1) * em                    # :552 (only real edges exist)
    if dims['X'] == 0:                                                # DGT_concat_sim (:1030-1124): no structural heads at all
        extra = extra[..., :0]
    e = _lin(sd, 'edge_emb', torch.cat([edge_x, cond_edge_x, dist0], dim=-1))      # :553,557
    h = _lin(sd, 'node_emb', h)                                       # :556
    atom_hids, edge_hids = [h], [e]
    if trace is not None:
        trace.update(temb=temb, h0=h, e0=e, extra=extra, dist0=dist0)
    for i in range(dims['L']):                                        # :562-568
        bt = {} if trace is not None else None
        h, e, pos = mix_block(sd, f'e_block_{i}', pos, h, e, extra, m, em, st, dims, bt)
        pos = remove_mean_with_mask(pos, m)                           # config.model.CoM (:565-566)
        atom_hids.append(_lin(sd, f'node_{i}', h))
        edge_hids.append(_lin(sd, f'edge_{i}', e))
        if collect is not None:
            collect.append((h.clone(), e.clone(), pos.clone()))
        if trace is not None:
            bt.update(h=h, e=e, pos=pos)
            trace.setdefault('blocks', []).append(bt)
    ah = torch.cat(atom_hids, dim=-1)
    eh = torch.cat(edge_hids, dim=-1)
    if trace is not None:
        trace.update(ah=ah, eh=eh)
    atom_pred = _mlp3(sd, 'node_pred_mlp', ah) * m                    # :573
    ep = torch.cat([_mlp3(sd, 'edge_exist_mlp', eh), _mlp3(sd, 'edge_type_mlp', eh)], dim=-1) * em  # :574-578
    ef = 0.5 * (ep + ep.permute(0, 2, 1, 3))                          # :579
    pos = pos * m                                                     # pred_data (:582-583)
    if bool(torch.any(torch.isnan(pos))):                             # :587-589
        pos = torch.zeros_like(pos)
    pos = remove_mean_with_mask(pos, m)                               # :592
    return torch.cat([pos, atom_pred], dim=2), ef
