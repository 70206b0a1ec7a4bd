"""GPU unit tests of the persistent TMA-fed GEMM (jodo_imglinear) and of the LayerNorm/modulate kernel that writes its
fp16 operand images (jodo_ln_mod_img), through the C ABI."""
import ctypes

import pytest
import torch

from jodo_b200 import _lib
from jodo_b200.pack import image_to_matrix_h, weight_image_h
from jodo_b200.plan import Plan

pytestmark = pytest.mark.gpu


def act_image(A):
    """fp32 rows [M, K] -> fp16 operand image [ceil(M/128)][K/64][128][64] (same swizzle as the weight images)."""
    M, K = A.shape
    mt = (M + 127) // 128
    Ap = torch.zeros(mt * 128, K, device=A.device)
    Ap[:M] = A
    return weight_image_h(Ap, 128).view(torch.float16)


def image_rows(img, K):
    mt = img.numel() // (128 * K)
    return torch.cat([image_to_matrix_h(img.reshape(mt, -1)[t], 128, K) for t in range(mt)])


def h(x):
    return x.float().half().double()


@pytest.mark.parametrize('M,K,N,NT', [(128, 64, 64, 64), (300, 256, 768, 256), (1000, 256, 64, 64), (45105, 256, 512, 256),
                                     (777, 512, 256, 256), (129, 1024, 256, 128), (5000, 256, 128, 128)])
def test_imglinear_matches_fp64(M, K, N, NT):
    g = torch.Generator(device='cuda').manual_seed(M + K + N)
    A = torch.randn(M, K, device='cuda', generator=g)
    W = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    b = torch.randn(N, device='cuda', generator=g)
    C32 = torch.full((M, N), float('nan'), device='cuda')
    C16 = torch.zeros(M, N, device='cuda', dtype=torch.float16)
    Cimg = torch.zeros(((M + 127) // 128) * 128 * max(N, 64), device='cuda', dtype=torch.float16) if N % 64 == 0 else None
    _lib.imglinear(act_image(A), M, K, weight_image_h(W, NT), b, N, NT, C32=C32, C16=C16, Cimg=Cimg)
    torch.cuda.synchronize()
    ref = (h(A) @ h(W).t() + b.double())
    assert float((C32.double() - ref).abs().max()) < 2e-4
    assert float((C16.double() - ref).abs().max()) < 1e-2
    if Cimg is not None:
        rows = image_rows(Cimg, N)
        assert float((rows[:M].double() - ref).abs().max()) < 1e-2
        assert float(rows[M:].abs().sum()) == 0.0          # padding rows of the last tile are zero


def test_imglinear_epilogues():
    g = torch.Generator(device='cuda').manual_seed(3)
    M, K, N, NT = 1500, 512, 256, 256
    A = torch.randn(M, K, device='cuda', generator=g)
    W = torch.randn(N, K, device='cuda', generator=g) / 16
    b = torch.randn(N, device='cuda', generator=g)
    Ai, Wi = act_image(A), weight_image_h(W, NT)
    base = h(A) @ h(W).t() + b.double()
    # SiLU -> image (the ff1 -> ff2 hand-off)
    Cimg = torch.zeros(((M + 127) // 128) * 128 * N, device='cuda', dtype=torch.float16)
    _lib.imglinear(Ai, M, K, Wi, b, N, NT, epi=_lib.EPI_ACT, act_out=_lib.ACT_SILU, Cimg=Cimg)
    ref = torch.nn.functional.silu(base)
    assert float((image_rows(Cimg, N)[:M].double() - ref).abs().max()) < 2e-2      # fp16 output + hardware tanh
    # gated residual -> fp32 rows (strided view) + image
    mol = torch.randint(0, 9, (M,), device='cuda', dtype=torch.int32, generator=g)
    gate = torch.randn(9, N + 64, device='cuda', generator=g)
    res = torch.randn(M, N, device='cuda', generator=g)
    Cbig = torch.zeros(M, N + 32, device='cuda')
    _lib.imglinear(Ai, M, K, Wi, b, N, NT, epi=_lib.EPI_GATED_RES, aux=res, gate=gate[:, 64:], row_mol=mol,
                   C32=Cbig[:, 16:16 + N], Cimg=Cimg)
    ref = res.double() + gate[:, 64:][mol.long()].double() * base
    assert float((Cbig[:, 16:16 + N].double() - ref).abs().max()) < 5e-4
    assert float(Cbig[:, :16].abs().max()) == 0 and float(Cbig[:, 16 + N:].abs().max()) == 0
    assert float((image_rows(Cimg, N)[:M].double() - ref).abs().max()) < 2e-2


def test_imglinear_piece_major_output():
    """fp16 output in the piece-major layout the edge kernels gather from: [N/8][rows][8]."""
    g = torch.Generator(device='cuda').manual_seed(11)
    M, K, N, NT = 1000, 256, 512, 256
    A = torch.randn(M, K, device='cuda', generator=g)
    W = torch.randn(N, K, device='cuda', generator=g) / 16
    b = torch.randn(N, device='cuda', generator=g)
    Cpm = torch.zeros(N // 8, M + 24, 8, device='cuda', dtype=torch.float16)
    Crm = torch.zeros(M, N, device='cuda', dtype=torch.float16)
    _lib.imglinear(act_image(A), M, K, weight_image_h(W, NT), b, N, NT, C16=Cpm)
    _lib.imglinear(act_image(A), M, K, weight_image_h(W, NT), b, N, NT, C16=Crm)
    torch.cuda.synchronize()
    back = Cpm[:, :M].permute(1, 0, 2).reshape(M, N)
    assert torch.equal(back, Crm)
    assert float(Cpm[:, M:].abs().sum()) == 0.0


def test_imglinear_rejects_bad_args():
    A = torch.zeros(128 * 64, device='cuda', dtype=torch.float16)
    with pytest.raises(_lib.JodoError):
        _lib.imglinear(A, 128, 40, A, None, 64, 64, C16=A.view(128, 64))
    with pytest.raises(_lib.JodoError):
        _lib.imglinear(A, 128, 64, A, None, 64, 64)                     # no output


def test_ln_mod_img_matches_torch():
    g = torch.Generator(device='cuda').manual_seed(9)
    n = torch.tensor([3, 17, 29, 1, 8] * 40)
    B, N, D = len(n), int(n.max()), 256
    mask = (torch.arange(N)[None] < n[:, None]).float().cuda()
    plan = Plan(mask)
    ps = _lib.plan_struct(plan)
    Nn = plan.Nn
    x = torch.randn(Nn, D, device='cuda', generator=g)
    y = torch.randn(Nn, D, device='cuda', generator=g)
    tab = torch.randn(B, 1024, device='cuda', generator=g)
    mt = (Nn + 127) // 128
    out32 = torch.empty(Nn, D, device='cuda')
    oimg = torch.full((mt * 128 * D,), 7.0, device='cuda', dtype=torch.float16)
    yimg = torch.full((mt * 128 * D,), 7.0, device='cuda', dtype=torch.float16)
    c = ctypes.c_int
    _lib.call('jodo_ln_mod_img', _lib.ptr(x), c(D), _lib.ptr(y), c(D), _lib.ptr(tab), c(1024), c(0), c(256), c(512),
              ctypes.byref(ps), _lib.ptr(out32), c(D), _lib.ptr(oimg), _lib.ptr(yimg), None, _lib.stream_ptr())
    torch.cuda.synchronize()
    t = tab[plan.node_mol.long()]
    z = x + t[:, :256] * y
    ref = torch.nn.functional.layer_norm(z, (D,), eps=1e-6) * t[:, 512:768] + t[:, 256:512]      # the table holds 1 + scale
    assert float((out32 - ref).abs().max()) < 2e-5
    rows = image_rows(oimg, D)
    assert float((rows[:Nn] - ref).abs().max()) < 4e-3 * float(ref.abs().max())
    assert float(rows[Nn:].abs().sum()) == 0.0
    yr = image_rows(yimg, D)
    assert torch.equal(yr[:Nn], y.half().float())
    assert float(yr[Nn:].abs().sum()) == 0.0


@pytest.mark.parametrize('M,K,N,NT,epi', [(1000, 256, 128, 128, 'gated'), (700, 128, 128, 128, 'store'), (260, 384, 384, 128, 'act'),
                                            (700, 384, 384, 192, 'act'), (390, 256, 512, 256, 'act'), (128, 128, 256, 256, 'act')])
def test_imglinear_placed_images_and_row_dots(M, K, N, NT, epi):
    """Placed image outputs (the output columns written at an offset inside wider operand images, only the first
    `ncols` columns, everything else untouched) and the fused row dot products of the activated output
    (dot_out[row, 4 slot + k], slots = column tile x column half), against fp64 on the fp16-rounded operands."""
    g = torch.Generator(device='cuda').manual_seed(M + N)
    A = torch.randn(M, K, device='cuda', generator=g)
    W = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    b = torch.randn(N, device='cuda', generator=g)
    Ai, Wi = act_image(A), weight_image_h(W, NT)
    mt = (M + 127) // 128
    ref = h(A) @ h(W).t() + b.double()
    if epi == 'act':
        dw = torch.randn(3, N, device='cuda', generator=g)
        nslots = 2 * N // NT
        dot = torch.full((M, 64), float('nan'), device='cuda')
        _lib.imglinear(Ai, M, K, Wi, b, N, NT, epi=_lib.EPI_ACT, act_out=_lib.ACT_SILU, dot_w=dw, dot_out=dot)
        torch.cuda.synchronize()
        act = ref * torch.sigmoid(ref)
        want = act @ dw.double().t()
        got = dot[:, :4 * nslots].double().reshape(M, nslots, 4)
        assert float(got[..., 3].abs().max()) == 0.0
        assert float((got[..., :3].sum(1) - want).abs().max()) < 2e-3 * float(want.abs().max())
        # every slot is the dot over its own NT / 2 columns
        cw = NT // 2
        for s in range(nslots):
            ws_ = act[:, s * cw:(s + 1) * cw] @ dw.double()[:, s * cw:(s + 1) * cw].t()
            assert float((got[:, s, :3] - ws_).abs().max()) < 2e-3 * float(want.abs().max())
        return
    k1, c1, n1 = 192, 0, 96                     # the [e | dist] operand: e columns only
    k2, c2, n2 = 384, 128, 96                   # a slot of a wider operand
    img1 = torch.full((mt * 128 * k1,), 7.0, device='cuda', dtype=torch.float16)
    img2 = torch.full((mt * 128 * k2,), 7.0, device='cuda', dtype=torch.float16)
    C32 = torch.empty(M, N, device='cuda')
    kw = {}
    if epi == 'gated':
        aux = torch.randn(M, N, device='cuda', generator=g)
        gate = torch.randn(5, N, device='cuda', generator=g)
        rm = torch.randint(0, 5, (M,), device='cuda', generator=g, dtype=torch.int32)
        kw = dict(epi=_lib.EPI_GATED_RES, aux=aux, gate=gate, row_mol=rm)
        ref = aux.double() + gate.double()[rm.long()] * ref
    _lib.imglinear(Ai, M, K, Wi, b, N, NT, C32=C32, Cimg=img1, cimg_place=(k1, c1, n1), Cimg2=img2, cimg2_place=(k2, c2, n2), **kw)
    torch.cuda.synchronize()
    assert float((C32.double() - ref).abs().max()) < 3e-4 * max(1.0, float(ref.abs().max()))
    for img, k, c0, nc in ((img1, k1, c1, n1), (img2, k2, c2, n2)):
        rows = image_rows(img, k)
        assert float((rows[:M, c0:c0 + nc].double() - ref[:, :nc]).abs().max()) < 1e-2 * max(1.0, float(ref.abs().max()))
        keep = torch.ones(k, dtype=torch.bool)
        keep[c0:c0 + nc] = False
        assert bool((rows[:, keep] == 7.0).all())            # nothing outside the placed columns is touched
    # bad placements are refused before any launch
    with pytest.raises(_lib.JodoError):
        _lib.imglinear(Ai, M, K, Wi, b, N, NT, C32=C32, Cimg=img1, cimg_place=(k1, 4, n1))
    with pytest.raises(_lib.JodoError):
        _lib.imglinear(Ai, M, K, Wi, b, N, NT, C32=C32, Cimg=img1, cimg_place=(k1, 128, 96))


@pytest.mark.parametrize('M,K,W', [(1000, 256, 96), (128, 192, 64), (389, 128, 128)])
def test_imglinear_layernorm_modulate_epilogue(M, K, W):
    """JODO_EPI_LN_MOD: LayerNorm (eps 1e-6) over the first W columns of acc + bias, modulation by the per-molecule table
    row, written as the fp16 operand image; columns >= W and padding rows are zeros; uniform flag -> row 0 for every row."""
    g = torch.Generator(device='cuda').manual_seed(M + W)
    N = NT = 128
    A = torch.randn(M, K, device='cuda', generator=g)
    Wt = torch.zeros(N, K, device='cuda')
    Wt[:W] = torch.randn(W, K, device='cuda', generator=g) / K ** 0.5
    b = torch.zeros(N, device='cuda')
    b[:W] = torch.randn(W, device='cuda', generator=g)
    B, ld = 6, 16 + 2 * 128
    tab = torch.randn(B, ld, device='cuda', generator=g)
    mol = torch.randint(0, B, (M,), device='cuda', generator=g, dtype=torch.int32)
    valid = torch.zeros(M, device='cuda', dtype=torch.int32)
    valid[::11] = -1
    Ai, Wi = act_image(A), weight_image_h(Wt, NT)
    mt = (M + 127) // 128
    x = (h(A) @ h(Wt).t() + b.double())[:, :W]
    mu = x.mean(1, keepdim=True)
    var = (x * x).mean(1, keepdim=True) - mu * mu
    nrm = (x - mu) / torch.sqrt(var.clamp(min=0) + 1e-6)
    for uni in (False, True):
        flag = torch.tensor([0 if uni else 1], device='cuda', dtype=torch.int32)
        t = tab.double()[torch.zeros_like(mol).long() if uni else mol.long()]
        want = nrm * t[:, 16 + 128:16 + 128 + W] + t[:, 16:16 + W]
        want[valid < 0] = 0
        img = torch.full((mt * 128 * N,), 7.0, device='cuda', dtype=torch.float16)
        _lib.imglinear(Ai, M, K, Wi, b, N, NT, epi=_lib.EPI_LN_MOD, Cimg=img, gate=tab, row_mol=mol, nonuni=flag.data_ptr(),
                       ln=(W, 16, 16 + 128), ln_valid=valid)
        torch.cuda.synchronize()
        rows = image_rows(img, N)
        assert float((rows[:M, :W].double() - want).abs().max()) < 2e-3 * max(1.0, float(want.abs().max()))
        assert float(rows[:M, W:].abs().max()) == 0.0 if W < N else True
        assert float(rows[:M][valid < 0].abs().max()) == 0.0
    with pytest.raises(_lib.JodoError):
        _lib.imglinear(Ai, M, K, Wi, b, N, NT, epi=_lib.EPI_LN_MOD, Cimg=img, gate=tab, row_mol=mol, ln=(W + 4, 16, 16 + 128))
