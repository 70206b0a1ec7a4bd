"""ctypes binding of libjodo_b200.so (the C ABI declared in include/jodo_b200.h).

There is no fallback: if the library is missing or a call fails, this raises."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libjodo_b200.so')

_lib = None

c_fp = ctypes.c_void_p
c_int = ctypes.c_int

ACT_NONE, ACT_SILU, ACT_GELU = 0, 1, 2
EPI_STORE, EPI_ACT, EPI_ADD, EPI_GATED_RES = 0, 1, 2, 3


class JodoError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise JodoError(f'{LIB_PATH} not built: run `python -c "import __graft_entry__ as g; g.build()"` '
                            f'(there is no CPU or PyTorch fallback for the DGT hot path)')
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.jodo_last_error_string.restype = ctypes.c_char_p
        if _lib.jodo_abi_version() != 1:
            raise JodoError('libjodo_b200.so ABI version mismatch; rebuild')
    return _lib


def check(rc, what):
    if rc != 0:
        raise JodoError(f'{what} failed ({rc}): {lib().jodo_last_error_string().decode()}')


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def rowlinear(A, K, Wimg, bias, C, N, NT, act_in=ACT_NONE, epi=EPI_STORE, act_out=ACT_NONE, aux=None, gate=None,
              row_mol=None, M=None, stream=None):
    """C[:, :N] = epi(act_in(A[:, :K]) W^T + bias); A, C, aux, gate are 2-D row-major views (stride(1)==1)."""
    M = A.shape[0] if M is None else M
    rc = lib().jodo_rowlinear(ptr(A), c_int(A.stride(0)), c_int(M), c_int(K), ptr(Wimg), ptr(bias), ptr(C),
                              c_int(C.stride(0)), c_int(N), c_int(NT), c_int(act_in), c_int(epi), c_int(act_out),
                              ptr(aux), c_int(0 if aux is None else aux.stride(0)), ptr(gate),
                              c_int(0 if gate is None else gate.stride(0)), ptr(row_mol),
                              stream if stream is not None else stream_ptr())
    check(rc, 'jodo_rowlinear')
