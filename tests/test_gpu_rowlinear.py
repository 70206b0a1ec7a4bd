"""GPU unit test of the tcgen05 tile-GEMM primitive (fp16 operands, fp32 accumulation) through jodo_rowlinear (C ABI)."""
import pytest
import torch

from jodo_b200 import _lib
from jodo_b200.pack import weight_image_h

pytestmark = pytest.mark.gpu


def _ref(A, W, b, act_in=None):
    A = A.double()
    if act_in == 'silu':
        A = torch.nn.functional.silu(A)
    h = lambda x: x.float().half().double()          # operands are rounded to fp16 (the mantissa of tf32)
    return (h(A) @ h(W).t() + (0 if b is None else b.double())).float()


@pytest.mark.parametrize('M,K,N,NT', [(128, 64, 16, 16), (128, 64, 64, 64), (300, 256, 768, 256), (77, 1024, 512, 128),
                                     (1000, 128, 32, 32), (257, 768, 256, 256)])
def test_rowlinear_matches_fp64(M, K, N, NT):
    g = torch.Generator(device='cuda').manual_seed(M * 7 + K)
    A = torch.randn(M, K, device='cuda', generator=g)
    W = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    b = torch.randn(N, device='cuda', generator=g)
    C = torch.full((M, N), float('nan'), device='cuda')
    _lib.rowlinear(A, K, weight_image_h(W, NT), b, C, N, NT)
    torch.cuda.synchronize()
    ref = _ref(A, W, b)
    err = float((C - ref).abs().max())
    assert err < 2e-4, err          # operands are identically fp16-rounded: only fp32 accumulation order differs


def test_rowlinear_epilogues_and_strides():
    g = torch.Generator(device='cuda').manual_seed(5)
    M, K, N, NT = 333, 256, 256, 128
    Abig = torch.randn(M, K + 64, device='cuda', generator=g)
    A = Abig[:, 32:32 + K]                                   # lda != K, offset view
    W = torch.randn(N, K, device='cuda', generator=g) / 16
    b = torch.randn(N, device='cuda', generator=g)
    Wi = weight_image_h(W, NT)
    # SiLU on input, GELU on output, strided output
    Cbig = torch.zeros(M, N + 128, device='cuda')
    _lib.rowlinear(A, K, Wi, b, Cbig[:, 64:64 + N], N, NT, act_in=_lib.ACT_SILU, epi=_lib.EPI_ACT, act_out=_lib.ACT_GELU)
    ref = torch.nn.functional.gelu(_ref(A, W, b, 'silu'))
    assert float((Cbig[:, 64:64 + N] - ref).abs().max()) < 5e-4
    assert float(Cbig[:, :64].abs().max()) == 0 and float(Cbig[:, 64 + N:].abs().max()) == 0
    # gated residual
    mol = torch.randint(0, 7, (M,), device='cuda', dtype=torch.int32, generator=g)
    gate = torch.randn(7, N, device='cuda', generator=g)
    res = torch.randn(M, N, device='cuda', generator=g)
    C = torch.empty(M, N, device='cuda')
    _lib.rowlinear(A, K, Wi, b, C, N, NT, epi=_lib.EPI_GATED_RES, aux=res, gate=gate, row_mol=mol)
    ref = res + gate[mol.long()] * _ref(A, W, b)
    assert float((C - ref).abs().max()) < 5e-4
    # add
    _lib.rowlinear(A, K, Wi, None, C, N, NT, epi=_lib.EPI_ADD, aux=res)
    assert float((C - (res + _ref(A, W, None))).abs().max()) < 5e-4


def test_rowlinear_rejects_bad_args():
    A = torch.zeros(8, 40, device='cuda')
    with pytest.raises(_lib.JodoError):
        _lib.rowlinear(A, 40, A, None, A, 16, 16)


@pytest.mark.parametrize('K,N,NT,act', [(1024, 19712, 128, 'silu'), (64, 256, 256, None), (256, 768, 192, 'silu')])
def test_row0_linear_matches_fp64_and_obeys_the_flag(K, N, NT, act):
    """jodo_row0_linear: row 0 of the rowlinear product on the same weight image (fp16-rounded operands, fp32 accumulation), run only
    while the device flag reads 0."""
    import ctypes
    g = torch.Generator(device='cuda').manual_seed(K + N)
    A = torch.randn(5, K, device='cuda', generator=g)
    W = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    b = torch.randn(N, device='cuda', generator=g)
    Wi = weight_image_h(W, NT)
    out = torch.full((3, N), 7.0, device='cuda')
    gelu = K == 64                                            # the time_mlp.1 shape: GELU on the output, plus an added row
    aux = torch.randn(N, device='cuda', generator=g)
    flag = torch.zeros(1, device='cuda', dtype=torch.int32)
    c = ctypes.c_int

    def run():
        _lib.call('jodo_row0_linear', _lib.ptr(A), c(K), _lib.ptr(Wi), c(NT), c(N), _lib.ptr(b), c(_lib.ACT_SILU if act else 0),
                  c(_lib.ACT_GELU if gelu else 0), _lib.ptr(aux) if gelu else None, _lib.ptr(out), _lib.ptr(flag), _lib.stream_ptr())
        torch.cuda.synchronize()
    run()
    x = A[0].double()
    if act:
        x = torch.nn.functional.silu(x)
    ref = W.half().double() @ x.float().half().double() + b.double()                  # both operands fp16-rounded, as in the GEMM
    if gelu:
        ref = torch.nn.functional.gelu(ref) + aux.double()
    assert float((out[0].double() - ref).abs().max()) < 1e-4 * max(1.0, float(ref.abs().max()))
    assert bool((out[1:] == 7.0).all())
    out.fill_(7.0)
    flag.fill_(1)                                             # conditioning not uniform: the all-rows GEMM owns the table
    run()
    assert bool((out == 7.0).all())
