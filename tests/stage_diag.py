"""Stage-by-stage comparison of the CUDA forward against the oracle (debugging tool + used by the GPU
parity test to print where a mismatch starts)."""
import torch

from helpers import golden_weights, load_golden, oracle_forward
from jodo_b200.pack import image_to_matrix, image_to_matrix_h


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def tiles_to_dense(plan, img, k, tile_floats=None, group_first=True):
    """tile images [n_tiles][k/32][128][32] -> dense [B,N,N,k]"""
    nt = plan.n_tiles
    tile_floats = (k // 32) * 128 * 32 if tile_floats is None else tile_floats
    img = img.reshape(nt, tile_floats)[:, :(k // 32) * 4096]
    rows = torch.stack([image_to_matrix(img[t], 128, k) for t in range(nt)]).reshape(nt * 128, k)
    return plan.rows_to_dense(rows, group_first=group_first)


def tiles_to_dense_h(plan, img, k, tile_halves, group_first=True):
    """fp16 PAIR-tile images [n_pair_tiles][k/64][128][64] -> dense symmetric [B,N,N,k] (fp32)"""
    nt = plan.n_pair_tiles
    img = img.reshape(nt, tile_halves)[:, :(k // 64) * 8192]
    rows = torch.stack([image_to_matrix_h(img[t], 128, k) for t in range(nt)]).reshape(nt * 128, k)
    return plan.pairs_to_dense(rows)


def act_image_to_rows(img, k):
    """fp16 activation image [mt][k/64][128][64] -> fp32 rows [mt*128, k]"""
    mt = img.numel() // (128 * k)
    return torch.cat([image_to_matrix_h(img.reshape(mt, -1)[t], 128, k) for t in range(mt)])


def packed_to_dense(plan, x):
    B, N = plan.B, plan.N
    out = torch.zeros(B * N, x.shape[1], dtype=x.dtype, device=x.device)
    out[plan.node_dense.long()] = x
    return out.reshape(B, N, -1)


def _fused_stages(model, dbg, trace, plan, inp, em, rep, D):
    eh = dbg['eh']
    rep.append(('e0', rel(tiles_to_dense_h(plan, eh, 64, model._plans[next(iter(model._plans))][1].eh_tile_bytes // 2,
                                           group_first=False), trace['e0'] * em)))
    for l, (b, ob) in enumerate(zip(dbg['blocks'], trace['blocks'])):
        m = inp['node_mask'].double()
        hn = act_image_to_rows(b['hn_img'], D)[:plan.Nn]
        rep.append((f'b{l}.hn', rel(packed_to_dense(plan, hn), ob['hn'] * m)))
        b['qkv'] = b['qkv'].permute(1, 0, 2).reshape(b['qkv'].shape[1], -1)      # piece-major -> rows
        qk = ob['q'].shape[-1]
        hq = qk // 2                      # q / k are stored as two head halves at columns [0, qk/2) and [D/2, D/2 + qk/2)
        unsplit = lambda x: torch.cat([x[:, :hq], x[:, D // 2:D // 2 + hq]], dim=1)
        rep.append((f'b{l}.q', rel(packed_to_dense(plan, unsplit(b['qkv'].float()[:, :D])), ob['q'] * m)))
        rep.append((f'b{l}.k', rel(packed_to_dense(plan, unsplit(b['qkv'].float()[:, D:2 * D])), ob['k'] * m)))
        rep.append((f'b{l}.v', rel(packed_to_dense(plan, b['qkv'].float()[:, 2 * D:]), ob['v'] * m)))
        rep.append((f'b{l}.hnode', rel(packed_to_dense(plan, b['hnode']), ob['hnode'] * m)))
        rep.append((f'b{l}.h', rel(packed_to_dense(plan, b['h']), ob['h'])))
        npt = plan.n_pair_tiles
        e_rows = b['e'].reshape(npt, 16, 128, 4).permute(0, 2, 1, 3).reshape(npt * 128, 64)   # piece-major pair tiles
        rep.append((f'b{l}.e', rel(plan.pairs_to_dense(e_rows), ob['e'] * em)))
        rep.append((f'b{l}.pos', rel(packed_to_dense(plan, b['pos'][:, :3]), ob['pos'])))


def run_case(name, device='cuda', verbose=True):
    from jodo_b200.model import MODELS
    g, cfg = load_golden(name)
    sd = golden_weights(g, cfg)
    trace = {}
    ox, oe = oracle_forward(sd, cfg, g['inputs'], torch.float64)
    oracle_forward(sd, cfg, g['inputs'], torch.float64, collect=None) if False else None
    from oracle.dgt_dense import dgt_forward
    c = lambda x: None if x is None else x.double()
    inp = g['inputs']
    dgt_forward({k: v.double() for k, v in sd.items()}, cfg, c(inp['t']), c(inp['xh']), c(inp['node_mask']),
                c(inp['edge_mask']), context=c(inp['context']), edge_x=c(inp['edge_x']),
                noise_level=c(inp['noise_level']), cond_x=c(inp['cond_x']), cond_edge_x=c(inp['cond_edge_x']), trace=trace)
    model = MODELS[cfg.model.name](cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(device).eval()
    model.debug = {}
    dev = lambda x: None if x is None else x.to(device)
    x, e = model(dev(inp['t']), dev(inp['xh']), dev(inp['node_mask']), dev(inp['edge_mask']), context=dev(inp['context']),
                 edge_x=dev(inp['edge_x']), noise_level=dev(inp['noise_level']), cond_x=dev(inp['cond_x']),
                 cond_edge_x=dev(inp['cond_edge_x']))
    torch.cuda.synchronize()
    dbg = model.debug
    plan = dbg['plan']
    rep = []
    D = model.dims.D
    if model.dims.two_d:                       # the 2-D oracle keeps no trace: outputs only
        rx, re_ = g['ref_fp64']
        rep = [('out.atom', rel(x, rx)), ('out.edge', rel(e, re_))]
        if verbose:
            for k, v in rep:
                print(f'{name:24s} {k:12s} {v:.3e}')
        return x, e, rep
    rep.append(('temb', rel(dbg['temb'], trace['temb'])))
    rep.append(('h0', rel(packed_to_dense(plan, dbg['ah'][:, :D]), trace['h0'] * inp['node_mask'].double())))
    em = inp['edge_mask'].reshape(plan.B, plan.N, plan.N, 1).double()
    if model.wide:
        ed = model.dims.ed
        m = inp['node_mask'].double()
        for l, (b, ob) in enumerate(zip(dbg['blocks'], trace['blocks'])):
            rep.append((f'b{l}.e1', rel(plan.pairs_to_dense(b['e1'][:, :ed]), ob['e1'] * em)))
            rep.append((f'b{l}.hnode', rel(packed_to_dense(plan, b['hnode']), ob['hnode'] * m)))
            rep.append((f'b{l}.h', rel(packed_to_dense(plan, b['h']), ob['h'])))
            rep.append((f'b{l}.e', rel(plan.pairs_to_dense(b['e'][:, :ed]), ob['e'] * em)))
            rep.append((f'b{l}.pos', rel(packed_to_dense(plan, b['pos'][:, :3]), ob['pos'])))
    else:
        _fused_stages(model, dbg, trace, plan, inp, em, rep, D)
    rx, re_ = g['ref_fp64']
    rep.append(('ah', rel(packed_to_dense(plan, dbg['ah'][:, :D]), trace['ah'][..., :D] * inp['node_mask'].double())))
    rep.append(('out.pos', rel(x[..., :3], rx[..., :3])))
    rep.append(('out.atom', rel(x[..., 3:], rx[..., 3:])))
    rep.append(('out.edge', rel(e, re_)))
    if verbose:
        for k, v in rep:
            print(f'{name:24s} {k:12s} {v:.3e}')
    return x, e, rep


if __name__ == '__main__':
    import sys
    for n in sys.argv[1:] or ['qm9_first']:
        run_case(n)
