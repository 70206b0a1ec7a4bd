"""Wide path of the DGT denoiser: hidden sizes the fused edge-tile kernels are not built for
(``model.nf = 384``, the reference's "large" GEOM-Drugs model, reference README.md:156,168).

Same boundary and the same plan / packed-atom layout as the fused path (jodo_b200/model.py), but the
per-edge work is a sequence of persistent tcgen05 GEMMs (``jodo_imglinear``) over the plan's edge rows
with the row kernels of csrc/wide.cu between them; per-edge intermediates live in HBM.  Reference lines:
models/mol_gnn.py:491-594 (forward), :270-322 (block), :71-94 (coordinate update),
models/layers.py:131-186 (attention).
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib
from .pack import TAB_HEAD, ceil_to, tab_layer_stride

_c = ctypes.c_int
# A/B switch: the coordinate branch as ONE kernel (LayerNorm warps producing coord_mlp.0's operand tile in shared memory,
# csrc/wide_equi.cu).  Correct (parity-green), measured SLOWER than the LayerNorm row kernel + GEMM pair on B200 (1.4 vs 1.0 ms
# per block at GEOM nf = 384, B = 512): eight LayerNorm warps per SM cannot keep enough gathers in flight, where the row kernel
# runs at full occupancy.  Off by default.
FUSED_EQUI = os.environ.get('JODO_WIDE_EQUI_FUSED') == '1'
# column tile of coord_mlp.0 (fused-row-dots GEMM, csrc/imglinear.cu k_imglinear_dot2: two row tiles share every W chunk, so the
# widest tile that divides D and fits two accumulators in tensor memory moves the fewest L2 bytes per FLOP); JODO_C0_NT: A/B
C0_NT = int(os.environ.get('JODO_C0_NT', '0'))


def c0_tile(D):
    if C0_NT or FUSED_EQUI:                   # the fused variant (csrc/wide_equi.cu) is built for 128-column tiles
        return C0_NT or 128
    return 256 if D % 256 == 0 else (192 if D % 192 == 0 else 128)


EDP = 128            # row stride (floats) of the per-edge fp32 buffers and K of the per-edge images (ed <= 128)


# A/B switch: the edge FFN as LayerNorm row kernel + two streaming GEMMs (the path before csrc/wide_ffn.cu)
FFN_UNFUSED = os.environ.get('JODO_WIDE_FFN_UNFUSED') == '1'


# A/B switch: block edge_emb as a plain GEMM followed by the LayerNorm row kernel (before the JODO_EPI_LN_MOD epilogue)
EMB_LN_UNFUSED = os.environ.get('JODO_WIDE_EMB_LN_UNFUSED') == '1'


def ffn_fused(d) -> bool:
    """Sizes csrc/wide_ffn.cu is built for (both weight images + two tiles in shared memory, r ed <= 256 TMEM columns)."""
    return d.r in (2, 4) and d.ed in (32, 64, 96)


def supported(d) -> str | None:
    """None when the wide path covers these sizes, else the reason."""
    if d.D % 128 or d.D > 512:
        return f'model.nf must be a multiple of 128 up to 512 (got {d.D})'
    if d.ed % 8 or d.ed > EDP:
        return f'edge width nf/4 must be a multiple of 8 up to {EDP} (got {d.ed})'
    if d.H > 32 or d.D % d.H or (d.D // d.H) % 4:
        return f'n_heads must divide nf into head widths that are multiples of 4, at most 32 heads (got {d.H})'
    if 2 * d.ch + d.ed > EDP:
        return f'edge_ch too large for the embedding image (got {d.ch})'
    return None


def pack_wide(pk, sd, d, add_lin, lin):
    """Edge-level and per-block weights of the wide path, recorded on the packer of pack.pack_model (which packs the
    molecule / atom level pieces).  Every GEMM runs with 128-column tiles; N and K are zero padded."""
    from .pack import _gbf_all
    D, ed, L, r = d.D, d.ed, d.L, d.r
    dev = pk.device
    W = lambda n: sd[n + '.weight']
    Bv = lambda n: sd[n + '.bias']
    qkp = ceil_to(d.qk, 128)
    pk.meta.update(qkp=qkp, wide=True)
    # ---- model level: edge_emb on [dist0 (ed) | edge_x (ch) | cond_edge_x (ch)]; 2-D model: [edge_x | cond_edge_x]
    we = W('edge_emb')                                         # [ed, 2ch + ed]: [edge_x | cond_edge_x | dist]
    if d.two_d:
        pk.mat('gbf', 3, EDP, [])
        lin('edge_emb', 'edge_emb', 128)                       # K = 64
    else:
        # {mu, sqrt(0.5 log2 e) / sg, 1 / (a sg)} x EDP indexed by feature COLUMN: Gaussian k sits at entry k + 1 (column 0 of
        # the features is the raw x)
        mu, c1, c2 = _gbf_all(sd, ['dist_layer'] + [f'e_block_{l}.dist_layer' for l in range(L)], dev)
        pk._keep += [mu, c1, c2]
        pk.mat('gbf', 3, EDP, [(mu[0], 0, 1), (c1[0], 1, 1), (c2[0], 2, 1)])
        add_lin('edge_emb', [(we[:, 2 * d.ch:], 0, 0), (we[:, :2 * d.ch], 0, ed)], [(Bv('edge_emb'), 0)], 128, ed, EDP)
    # ---- edge heads: layer 0 of edge_exist_mlp | edge_type_mlp is linear in the concatenated edge hiddens
    # cat[e0, edge_0(e_1), ..] (reference models/mol_gnn.py:567-574), so the edge_i projections are folded into it:
    # H = W0[:, :ed] e0 + sum_l (W0[:, s_l] We_l) e_l + b is ONE GEMM over the operand [e0 | e_1 | .. | e_L] (slots of
    # EDP columns, written block by block), K = (L + 1) EDP.  The L small products run as two batched einsums.
    f32 = lambda t: t.detach().to(dev, torch.float32)
    w0 = torch.cat([f32(W('edge_exist_mlp.0')), f32(W('edge_type_mlp.0'))], dim=0)            # [2ed, ed + L ce]
    sl = w0[:, ed:ed + L * d.ce].reshape(2 * ed, L, d.ce)
    wl = torch.stack([f32(W(f'edge_{l}')) for l in range(L)])                                 # [L, ce, ed]
    bl = torch.stack([f32(Bv(f'edge_{l}')) for l in range(L)])                                # [L, ce]
    prod = torch.einsum('nlc,lce->nle', sl, wl).contiguous()                                  # [2ed, L, ed]
    b0 = torch.cat([f32(Bv('edge_exist_mlp.0')), f32(Bv('edge_type_mlp.0'))]) + torch.einsum('nlc,lc->n', sl, bl)
    pk._keep += [w0, prod, b0]
    hp = ceil_to(2 * ed, 128)
    add_lin('hcat', [(w0[:, :ed], 0, 0)] + [(prod[:, l, :], 0, (l + 1) * EDP) for l in range(L)], [(b0, 0)], 128, 2 * ed,
            (L + 1) * EDP, n_pad=hp)
    add_lin('ehead2', [(W('edge_exist_mlp.2'), 0, 0), (W('edge_type_mlp.2'), ed // 2, ed)],
            [(Bv('edge_exist_mlp.2'), 0), (Bv('edge_type_mlp.2'), ed // 2)], 128, ed, hp)
    pk.mat('ehead4.w', d.ch, ed // 2, [(W('edge_exist_mlp.4'), 0, 0), (W('edge_type_mlp.4'), 1, 0)])     # [ch, ed / 2]
    pk.vec('ehead4.b', d.ch, [(Bv('edge_exist_mlp.4'), 0), (Bv('edge_type_mlp.4'), 1)])
    pk.meta['hp'] = hp
    # ---- blocks
    f3p = ceil_to(ed * r, 128)
    pk.meta['f3p'] = f3p
    for l in range(L):
        b = f'e_block_{l}'
        p = f'b{l}.'
        add_lin(p + 'qkv',
                [(W(f'{b}.attn_mpnn.lin_query'), 0, 0), (W(f'{b}.attn_mpnn.lin_key'), qkp, 0), (W(f'{b}.attn_mpnn.lin_value'), 2 * qkp, 0)],
                [(Bv(f'{b}.attn_mpnn.lin_query'), 0), (Bv(f'{b}.attn_mpnn.lin_key'), qkp), (Bv(f'{b}.attn_mpnn.lin_value'), 2 * qkp)],
                128, 2 * qkp + D, D)
        lin(p + 'n2e', f'{b}.node2edge_lin', 128, bias=False)
        pk.vec(p + 'n2e.bias', EDP, [(Bv(f'{b}.node2edge_lin'), 0)])
        lin(p + 'ff1', f'{b}.ff_linear1', 128)
        lin(p + 'ff2', f'{b}.ff_linear2', 128)
        lin(p + 'node_l', f'node_{l}', 128)
        if not d.two_d:
            wi, bi = W(f'{b}.equi_update.input_lin'), Bv(f'{b}.equi_update.input_lin')    # [D, 2D + 2ed]: [h_row | h_col | e | dist]
            add_lin(p + 'ab', [(wi[:, :D], 0, 0), (wi[:, D:2 * D], D, 0)], [(bi, 0)], 128, 2 * D, D)   # bias rides on the h[row] part
            pk.mat(p + 'gbf', 3, EDP, [(mu[1 + l], 0, 1), (c1[1 + l], 1, 1), (c2[1 + l], 2, 1)])
            # block edge_emb reads cat[dist, e] (mol_gnn.py:287), input_lin's edge part cat[e, dist] (:72): the columns of the
            # former are reordered so that both GEMMs read ONE operand image [e | dist] (K = 2ed)
            we_l = W(f'{b}.edge_emb')
            add_lin(p + 'emb', [(we_l[:, ed:], 0, 0), (we_l[:, :ed], 0, ed)], [(Bv(f'{b}.edge_emb'), 0)], 128, ed, 2 * ed)
            add_lin(p + 'equi_in', [(wi[:, 2 * D:], 0, 0)], [], 128, D, 2 * ed)             # K = 2ed: [e | dist]
            lin(p + 'c0', f'{b}.equi_update.coord_mlp.0', c0_tile(D))
            # coord_mlp.2 (1 + X outputs, no bias) rides on coord_mlp.0's epilogue as three fp32 row dots (rows beyond 1 + X zero)
            pk.mat(p + 'c2.w32', 3, D, [(W(f'{b}.equi_update.coord_mlp.2'), 0, 0)])
        # 256-column tiles where they divide the width: 243 vs 266 us per launch at nf = 384 (tools/bench_gemm_wide.py)
        add_lin(p + 'g01', [(W(f'{b}.attn_mpnn.lin_edge0'), 0, 0), (W(f'{b}.attn_mpnn.lin_edge1'), qkp, 0)], [],
                256 if (qkp + D) % 256 == 0 else 128, qkp + D, EDP)
        add_lin(p + 'ff3', [(W(f'{b}.ff_linear3'), 0, 0)], [(Bv(f'{b}.ff_linear3'), 0)], 128, f3p, EDP, n_pad=f3p)
        add_lin(p + 'ff4', [(W(f'{b}.ff_linear4'), 0, 0)], [(Bv(f'{b}.ff_linear4'), 0)], 128, ed, f3p)
        if ffn_fused(d):
            # the fused edge FFN (csrc/wide_ffn.cu) keeps both weights resident as single-tile images: ff_linear3 [r ed, ed] and
            # its bias pre-scaled by 1/2 (SiLU(x) = h + h tanh h, h = x / 2; exact in fp16), ff_linear4 [ed, r ed]
            pk.image_h(p + 'ff3f.img', r * ed, ceil_to(ed, 64), r * ed, [(W(f'{b}.ff_linear3'), 0, 0, 0.5)])
            pk.vec(p + 'ff3f.b', r * ed, [(Bv(f'{b}.ff_linear3'), 0, 0.5)])
            pk.image_h(p + 'ff4f.img', ed, r * ed, ed, [(W(f'{b}.ff_linear4'), 0, 0)])
            pk.vec(p + 'ff4f.b', ed, [(Bv(f'{b}.ff_linear4'), 0)])
    if not d.two_d:
        cs = torch.stack([f32(sd[f'e_block_{l}.equi_update.coord_norm.scale']).reshape(()) for l in range(L)])
        pk.add_host('coord_scale', cs)


class WideWorkspace:
    """Device buffers of the wide path for one plan."""

    def __init__(self, plan, d, meta, dev):
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        zf = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        B, Nn, nt = plan.B, plan.Nn, plan.n_tiles
        D, T, R = d.D, d.T, plan.n_tiles * 128
        self.R = R
        self.feat, self.t1, self.temb = f(B, 64), f(B, T), f(B, T)
        if d.cond_ch:
            self.c1, self.c2, self.ctx = f(B * d.cond_ch, D), f(B * d.cond_ch, D), f(B, T)
        self.tab = f(B, meta['ld_tab'])
        bt = (B + 127) // 128 * 128
        self.temb_img = torch.zeros(bt * T, device=dev, dtype=torch.float16)
        self.kin = meta['node_emb']['K']
        self.xin = f(Nn, self.kin)
        self.pos = [zf(Nn, 4), zf(Nn, 4)]
        self.ah = zf(Nn, meta['ld_ah'])
        self.h = [f(Nn, D), f(Nn, D)]
        mt = (Nn + 127) // 128
        nimg = lambda k: torch.zeros(mt * 128 * k, device=dev, dtype=torch.float16)
        eimg = lambda k: torch.zeros(R * k, device=dev, dtype=torch.float16)
        self.hn_img, self.h2_img, self.hnode_img, self.hout_img = nimg(D), nimg(D), nimg(D), nimg(D)
        self.ff_img = nimg(d.r * D)
        self.ldq = 2 * meta['qkp'] + D
        self.qkv = torch.zeros(Nn, self.ldq, device=dev, dtype=torch.float16)
        self.hnode, self.h2 = zf(Nn, D), f(Nn, D)
        self.P = zf(Nn, EDP)
        if not d.two_d:
            self.AB = torch.zeros(Nn, 2 * D, device=dev, dtype=torch.float16)    # hoisted input_lin parts, gathered per edge
        self.n1, self.n2, self.ap = f(Nn, D), f(Nn, meta['npred2']['N']), f(Nn, meta['npred4']['N'])
        # The edge state and everything computed from it alone is symmetric in (r, c): those buffers
        # have one row per UNORDERED pair (plan.pair_*), RP rows; the directed rows (R) read them through plan.row_pair.
        # Only the coordinate branch behind input_lin (whose h[row] / h[col] parts are ordered) runs per directed row.
        RP = plan.n_pair_tiles * 128
        self.RP = RP
        pimg = lambda k: torch.zeros(RP * k, device=dev, dtype=torch.float16)
        self.A0 = pimg(64 if d.two_d else EDP)
        if not d.two_d:
            self.ED = pimg(2 * d.ed)                              # [e | dist]: operand of block edge_emb and of input_lin's edge part
            self.e1 = zf(RP, EDP)
        self.e32, self.e2 = zf(RP, EDP), zf(RP, EDP)
        self.en_img, self.e2_img = pimg(EDP), pimg(EDP)
        self.ldg = meta['qkp'] + D
        self.G = torch.zeros(RP, self.ldg, device=dev, dtype=torch.float16)
        self.f3_img = pimg(meta['f3p'])
        self.EH = pimg((d.L + 1) * EDP)                       # operand of the edge heads: [e0 | e_1 | .. | e_L]
        self.H_img = pimg(meta['hp'])
        self.X2 = zf(RP, EDP)
        if not d.two_d:
            self.U = torch.zeros(RP, D, device=dev, dtype=torch.float16)     # input_lin edge part (pre-LayerNorm), per pair
            self.u_img = eimg(D)                                              # per directed row
            self.c3 = zf(R, 64)                                               # coord_mlp.2 partial outputs: [slot][4]
        self.node_dense_l = plan.node_dense.long()
        self.extra = torch.zeros(RP, device=dev, dtype=torch.uint8)
        self.flags = torch.zeros(4, device=dev, dtype=torch.int32)            # [0] dist flag, [1] nan flag
        self.grp_row0, self.grp_len = plan.grp_row0, plan.grp_len


def forward_wide(self, pk, plan, ws, ps, pps, xh, edge_x, noise_level, cond_x, cond_edge_x, context):
    """The wide-path launch sequence of one denoiser evaluation (called by model._DGTBase.forward)."""
    d, meta = self.dims, pk.meta
    B, N, Nn, R, RP = plan.B, plan.N, plan.Nn, ws.R, ws.RP
    D, T, ed, ld_tab = d.D, d.T, d.ed, meta['ld_tab']
    st = _lib.stream_ptr()
    P, dp = _lib.ptr, _lib.dp
    dbg = self.debug

    def lin(name, A, C, M=None, **kw):
        m = meta[name]
        _lib.rowlinear(A, m['K'], pk[name + '.img'], pk[name + '.b'], C, m['N'], m['NT'], M=M, stream=st,
                       tag='jodo_rowlinear:' + name.split('.')[-1], **kw)

    def ilin(name, Aimg, M, bias=True, **kw):
        m = meta[name]
        _lib.imglinear(Aimg, M, m['K'], pk[name + '.img'], pk[name + '.b'] if bias else None, m['N'], m['NT'], stream=st,
                       tag='jodo_imglinear:' + name.split('.')[-1], **kw)

    def ln(M, W, K, x, tab_off, row_mol, out_img=None, out32=None, y=None, yi=None, y2=None, y2i=None, ybias=None,
           gate=-1, valid=None, y_img=None, xi=None, tag=''):
        shift, scale = tab_off
        a = _lib.WideLnArgs(M, W, K, dp(x), x.stride(0), dp(xi), dp(y), 0 if y is None else y.stride(0), dp(yi),
                            dp(y2), 0 if y2 is None else y2.stride(0), dp(y2i), dp(ybias), dp(ws.tab), ld_tab, dp(row_mol),
                            gate, shift, scale, dp(valid), dp(out32), 0 if out32 is None else out32.stride(0),
                            dp(out_img), dp(y_img), int(x.dtype == torch.float16),
                            int(y is not None and y.dtype == torch.float16), nonuni)
        _lib.call('jodo_wide_ln', ctypes.byref(a), st, tag='jodo_wide_ln:' + tag)

    # ---- per molecule: noise-level embedding (+ context) and all AdaLN tables
    _lib.call('jodo_time_features', P(noise_level), P(pk['time.w8']), P(ws.feat), _c(B), st)
    lin('time1', ws.feat, ws.t1, epi=_lib.EPI_ACT, act_out=_lib.ACT_GELU)
    if d.cond_ch:
        ctx = context.contiguous().float().reshape(B * d.cond_ch)
        _lib.call('jodo_cond_in', P(ctx), P(pk['cond0.w']), P(pk['cond0.b']), P(ws.c1), _c(B * d.cond_ch), _c(D), st)
        lin('cond2', ws.c1, ws.c2)
        lin('condlin', ws.c2.view(B, d.cond_ch * D), ws.ctx)
        lin('time3', ws.t1, ws.temb, epi=_lib.EPI_ADD, aux=ws.ctx)
    else:
        lin('time3', ws.t1, ws.temb)
    # flags[2] = 1 unless every molecule carries the same conditioning row (the samplers broadcast one noise level): the row
    # kernels then read table row 0 for every row (L1-resident, and the loads no longer wait for the row's molecule index)
    nonuni = ws.flags.data_ptr() + 8
    _lib.call('jodo_uniform_flag', P(ws.temb), _c(B), _c(T), ctypes.c_void_p(nonuni), st)
    _lib.call('jodo_act_image', P(ws.temb), _c(T), _c(B), _c(T), _c(_lib.ACT_SILU), P(ws.temb_img), st)
    ilin('tab', ws.temb_img, B, C32=ws.tab)
    # ---- per atom
    if d.two_d:                                   # no coordinates: xh = atom features [B, N, in] (mol_gnn.py:883, 897-899)
        ws.xin.zero_()
        ws.xin[:, :d.inn] = xh.reshape(B * N, d.inn)[ws.node_dense_l]
        if cond_x is not None:
            ws.xin[:, d.inn:2 * d.inn] = cond_x.reshape(B * N, d.inn)[ws.node_dense_l]
    else:
        _lib.call('jodo_gather_nodes', P(xh), P(cond_x), ctypes.byref(ps), _c(d.inn), _c(ws.kin), P(ws.xin), P(ws.pos[0]), None, st)
    lin('node_emb', ws.xin, ws.ah[:, :D])
    # ---- per edge: model-level embedding, adjacency heads
    # ---- per unordered pair: model-level embedding, adjacency heads
    ea = _lib.WideEmbedArgs(pps, dp(edge_x), dp(cond_edge_x), 0 if d.two_d else dp(cond_x), d.ch, d.inn,
                            0 if d.two_d else ed, self.edge_th, self.spatial_cut_off, dp(ws.flags), dp(ws.tab), ld_tab,
                            pk.ptr('gbf'), EDP, dp(ws.A0), 64 if d.two_d else EDP, dp(ws.extra))
    _lib.call('jodo_wide_embed_in', ctypes.byref(ea), st)
    K2 = 2 * ed
    KH = (d.L + 1) * EDP
    # the fp16 copies of the edge state are written by the GEMM that produces it: the e columns of the [e | dist] operand
    # and the block's slot of the edge heads' operand (no separate conversion pass)
    if d.two_d:
        ilin('edge_emb', ws.A0, RP, C32=ws.e32, Cimg=ws.EH, cimg_place=(KH, 0, ed))
    else:
        ilin('edge_emb', ws.A0, RP, C32=ws.e32, Cimg=ws.ED, cimg_place=(K2, 0, ed), Cimg2=ws.EH, cimg2_place=(KH, 0, ed))

    h = ws.ah[:, :D]
    stride = tab_layer_stride(D)
    for l in range(d.L):
        p = f'b{l}.'
        o = TAB_HEAD + l * stride
        oe, oq, og = o + 6 * D, o + 6 * D + 6 * ed, o + 6 * D + 6 * ed + 2 * D
        pin, pout = ws.pos[l & 1], ws.pos[(l + 1) & 1]
        hout = ws.h[l & 1]
        # distance features into the [dist | e] and [e | dist] operands; block edge_emb; norm1_edge; g0 | g1
        if d.two_d:                               # EquivariantMixBlock_2D: norm1_edge on the block input (mol_gnn.py:391)
            ln(RP, ed, EDP, ws.e32, (oe, oe + ed), plan.pair_mol, out_img=ws.en_img, valid=plan.pair_i, tag='e1')
        else:
            _lib.call('jodo_wide_dist', ctypes.byref(pps), P(pin), P(ws.tab), _c(ld_tab), _c(og), P(pk[p + 'gbf']), _c(EDP),
                      _c(ed), P(ws.ED), _c(K2), _c(ed), None, _c(0), _c(0), st)
            me = meta[p + 'emb']
            if me['N'] == 128 and me['NT'] == 128 and EDP == 128 and dbg is None and not EMB_LN_UNFUSED:
                # block edge_emb with norm1_edge + modulation as its epilogue: the fp32 edge_emb output never goes to HBM
                ilin(p + 'emb', ws.ED, RP, epi=_lib.EPI_LN_MOD, Cimg=ws.en_img, gate=ws.tab, row_mol=plan.pair_mol, nonuni=nonuni,
                     ln=(ed, oe, oe + ed), ln_valid=plan.pair_i)
            else:
                ilin(p + 'emb', ws.ED, RP, C32=ws.e1)
                ln(RP, ed, EDP, ws.e1, (oe, oe + ed), plan.pair_mol, out_img=ws.en_img, valid=plan.pair_i, tag='e1')
        ilin(p + 'g01', ws.en_img, RP, bias=False, epi=_lib.EPI_ACT, act_out=_lib.ACT_TANH, C16=ws.G)
        # attention
        ln(Nn, D, D, h, (o, o + D), plan.node_mol, out_img=ws.hn_img, tag='h1')
        ilin(p + 'qkv', ws.hn_img, Nn, C16=ws.qkv)
        aa = _lib.WideAttnArgs(Nn, D, d.H, d.X, d.sc, dp(ws.grp_row0), dp(ws.grp_len), dp(plan.row_j), dp(ws.qkv), ws.ldq,
                               meta['qkp'], 2 * meta['qkp'], dp(ws.G), ws.ldg, meta['qkp'], dp(ws.extra), dp(plan.row_pair),
                               dp(ws.hnode), max(plan.max_group, 1), dp(plan.mol_start), B, int(plan.n_nodes.max()))
        _lib.call('jodo_wide_attn', ctypes.byref(aa), st)
        # node path
        ln(Nn, D, D, h, (o + 3 * D, o + 4 * D), plan.node_mol, out_img=ws.h2_img, out32=ws.h2, y=ws.hnode, gate=o + 2 * D,
           y_img=ws.hnode_img, tag='h2')
        ilin(p + 'n2e', ws.hnode_img, Nn, bias=False, C32=ws.P)
        ilin(p + 'ff1', ws.h2_img, Nn, epi=_lib.EPI_ACT, act_out=_lib.ACT_SILU, Cimg=ws.ff_img)
        ilin(p + 'ff2', ws.ff_img, Nn, epi=_lib.EPI_GATED_RES, aux=ws.h2, gate=ws.tab[:, o + 5 * D:], row_mol=plan.node_mol,
             C32=hout, Cimg=ws.hout_img, nonuni=nonuni)
        if not d.two_d:
            ilin(p + 'ab', ws.hout_img, Nn, C16=ws.AB)
        ilin(p + 'node_l', ws.hout_img, Nn, C32=ws.ah[:, D + l * meta['cnp']:])
        # edge path: e2 = norm2(e + gate * node2edge(hnode[r] + hnode[c])), e_out = e2 + gate * FFN(e2); the fp16 copies of the new
        # state go to the e columns of the [e | dist] operand (3-D models) and to the block's slot of the edge heads' operand
        imgs = [(ws.EH, KH, (l + 1) * EDP)] if d.two_d else [(ws.ED, K2, 0), (ws.EH, KH, (l + 1) * EDP)]
        if (p + 'ff3f.img') in pk._off and not FFN_UNFUSED and dbg is None:
            # one kernel on pair tiles, weights resident: e32 is read and written once, e2 and the hidden rows stay on chip
            (i1, k1, c1), (i2, k2, c2) = imgs[0], (imgs[1] if len(imgs) > 1 else (None, 0, 0))
            fa = _lib.WideFfnArgs(RP, ed, d.r * ed, dp(ws.e32), ws.e32.stride(0), dp(ws.P), ws.P.stride(0), dp(plan.pair_i),
                                  dp(plan.pair_j), dp(plan.pair_mol), pk.ptr(p + 'n2e.bias'), dp(ws.tab), ld_tab, oe + 2 * ed,
                                  oe + 3 * ed, oe + 4 * ed, oe + 5 * ed, pk.ptr(p + 'ff3f.img'), pk.ptr(p + 'ff3f.b'),
                                  pk.ptr(p + 'ff4f.img'), pk.ptr(p + 'ff4f.b'), dp(i1), k1, c1, dp(i2), k2, c2, nonuni)
            _lib.call('jodo_wide_edge_ffn', ctypes.byref(fa), st)
        else:
            ln(RP, ed, EDP, ws.e32, (oe + 3 * ed, oe + 4 * ed), plan.pair_mol, out_img=ws.e2_img, out32=ws.e2, y=ws.P,
               yi=plan.pair_i, y2=ws.P, y2i=plan.pair_j, ybias=pk[p + 'n2e.bias'], gate=oe + 2 * ed, valid=plan.pair_i, tag='e2')
            ilin(p + 'ff3', ws.e2_img, RP, epi=_lib.EPI_ACT, act_out=_lib.ACT_SILU, Cimg=ws.f3_img)
            kw = dict(Cimg=imgs[0][0], cimg_place=(imgs[0][1], imgs[0][2], ed))
            if len(imgs) > 1:
                kw.update(Cimg2=imgs[1][0], cimg2_place=(imgs[1][1], imgs[1][2], ed))
            ilin(p + 'ff4', ws.f3_img, RP, epi=_lib.EPI_GATED_RES, aux=ws.e2, gate=ws.tab[:, oe + 5 * ed:], row_mol=plan.pair_mol,
                 C32=ws.e32, nonuni=nonuni, **kw)
        if d.two_d:
            if dbg is not None:
                dbg.setdefault('blocks', []).append(dict(hnode=ws.hnode.clone(), h=hout.clone(), e=ws.e32.clone()))
            h = hout
            continue
        # coordinate update: the [e | dist] part of input_lin per pair, everything behind the LayerNorm per directed row
        ilin(p + 'equi_in', ws.ED, RP, bias=False, C16=ws.U)
        mc0 = meta[p + 'c0']
        if FUSED_EQUI and D in (256, 384) and mc0['NT'] == 128 and mc0['N'] == D:
            # LayerNorm + modulation as the producer of coord_mlp.0's operand tile in shared memory, SiLU and the three
            # coord_mlp.2 dots in its epilogue: no operand image in HBM (csrc/wide_equi.cu)
            wa = _lib.WideEquiArgs(R, D, dp(ws.U), ws.U.stride(0), dp(plan.row_pair), dp(ws.AB), ws.AB.stride(0), dp(plan.row_g),
                                   dp(plan.row_j), dp(plan.row_mol), dp(ws.tab), ld_tab, oq, oq + D, pk.ptr(p + 'c0.img'),
                                   pk.ptr(p + 'c0.b'), pk.ptr(p + 'c2.w32'), dp(ws.c3), 64)
            _lib.call('jodo_wide_equi', ctypes.byref(wa), st)
            nslots = 2
        else:
            ln(R, D, D, ws.U, (oq, oq + D), plan.row_mol, out_img=ws.u_img, y=ws.AB, yi=plan.row_g, y2=ws.AB[:, D:],
               y2i=plan.row_j, valid=plan.row_g, xi=plan.row_pair, tag='equi')
            ilin(p + 'c0', ws.u_img, R, epi=_lib.EPI_ACT, act_out=_lib.ACT_SILU, dot_w=pk[p + 'c2.w32'], dot_out=ws.c3)
            nslots = 2 * mc0['N'] // mc0['NT']
        _lib.call('jodo_wide_equi_out', P(ws.grp_row0), P(ws.grp_len), P(plan.row_j), P(ws.c3), _c(64), _c(nslots), P(ws.extra),
                  P(plan.row_pair), _c(d.X), ctypes.c_float(meta['coord_scale'][l]), P(pin), P(pout), _c(Nn), st)
        _lib.call('jodo_com', P(pout), ctypes.byref(ps), st)
        if dbg is not None:
            dbg.setdefault('blocks', []).append(dict(hnode=ws.hnode.clone(), h=hout.clone(), e=ws.e32.clone(),
                                                     pos=pout.clone(), e1=ws.e1.clone(), e2=ws.e2.clone()))
        h = hout
    # ---- heads
    lin('npred0', ws.ah, ws.n1, epi=_lib.EPI_ACT, act_out=_lib.ACT_SILU)
    lin('npred2', ws.n1, ws.n2, epi=_lib.EPI_ACT, act_out=_lib.ACT_SILU)
    lin('npred4', ws.n2, ws.ap)
    if d.two_d:                                   # atom_pred * node_mask (mol_gnn.py:933): padding stays zero
        out_x = torch.zeros(B * N, d.inn, device=xh.device, dtype=torch.float32)
        out_x[ws.node_dense_l] = ws.ap[:, :d.inn]
        out_x = out_x.reshape(B, N, d.inn)
    else:
        out_x = torch.zeros(B, N, 3 + d.inn, device=xh.device, dtype=torch.float32)
        _lib.call('jodo_node_out', P(ws.pos[d.L & 1]), P(ws.ap), _c(ws.ap.stride(0)), ctypes.byref(ps),
                  ctypes.c_void_p(ws.flags.data_ptr() + 4), None, _c(d.inn), P(out_x), st)
    ilin('hcat', ws.EH, RP, epi=_lib.EPI_ACT, act_out=_lib.ACT_SILU, Cimg=ws.H_img)
    ilin('ehead2', ws.H_img, RP, epi=_lib.EPI_ACT, act_out=_lib.ACT_SILU, C32=ws.X2)
    # every pair row writes both orientations: the reference's 0.5 (e + e^T) (mol_gnn.py:579) of two identical values
    out_e = torch.zeros(B, N, N, d.ch, device=xh.device, dtype=torch.float32)
    _lib.call('jodo_wide_head_out', ctypes.byref(pps), P(ws.X2), _c(EDP), _c(ed // 2), P(pk['ehead4.w']), P(pk['ehead4.b']),
              _c(d.ch), _c(1), P(out_e), st)
    if dbg is not None:
        dbg.update(tab=ws.tab.clone(), temb=ws.temb.clone(), ah=ws.ah.clone(), extra=ws.extra.clone(),
                   plan=plan, flags=ws.flags.clone())
    return out_x, out_e
