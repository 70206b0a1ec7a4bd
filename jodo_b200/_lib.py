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

ACT_NONE, ACT_SILU, ACT_GELU, ACT_TANH = 0, 1, 2, 3
EPI_STORE, EPI_ACT, EPI_ADD, EPI_GATED_RES, EPI_LN_MOD = 0, 1, 2, 3, 4


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
        if _lib.jodo_abi_version() != 17:
            raise JodoError('libjodo_b200.so ABI version mismatch; rebuild')
    return _lib


def check(rc, what):
    if rc != 0:
        raise JodoError(f'{what} failed ({rc}): {lib().jodo_last_error_string().decode()}')


# ---- launch accounting / per-call device timing (bench.py, profiling) ------------------------------
# kernels launched per C-ABI call (everything else launches exactly one)
KERNELS_PER_CALL = {'jodo_edge_embed': 2, 'jodo_wide_embed_in': 2, 'jodo_node_out': 2, 'jodo_ancestral_update': 2, 'jodo_ancestral_update_philox': 2, 'jodo_dpm_update': 2}     # (memsets / D2D constant uploads are not kernels)
LAUNCHES = 0           # kernels launched through this binding since import
TRACE = None           # set to a list to record (name, start_event, end_event) around every call


def _account(name, fn):
    global LAUNCHES
    LAUNCHES += KERNELS_PER_CALL.get(name.split(':')[0], 1)
    if TRACE is None:
        return fn()
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn()
    e1.record()
    TRACE.append((name, e0, e1))
    return rc


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def rowlinear(A, K, Wimg, bias, C, N, NT, act_in=ACT_NONE, epi=EPI_STORE, act_out=ACT_NONE, aux=None, gate=None,
              row_mol=None, M=None, stream=None, tag=None, out_f16=False, skip_if_zero=None):
    """C[:, :N] = epi(act_in(A[:, :K]) W^T + bias); A, C, aux, gate are 2-D row-major views (stride(1)==1)."""
    M = A.shape[0] if M is None else M
    f = lib().jodo_rowlinear
    st = stream if stream is not None else stream_ptr()
    rc = _account(tag or 'jodo_rowlinear', lambda: f(
        ptr(A), c_int(A.stride(0)), c_int(M), c_int(K), ptr(Wimg), ptr(bias), ptr(C), c_int(C.stride(0)), c_int(N),
        c_int(NT), c_int(act_in), c_int(epi), c_int(act_out), ptr(aux), c_int(0 if aux is None else aux.stride(0)),
        ptr(gate), c_int(0 if gate is None else gate.stride(0)), ptr(row_mol), c_int(1 if out_f16 else 0),
        ctypes.c_void_p(skip_if_zero), st))
    check(rc, 'jodo_rowlinear')


# ---- argument blocks (mirror include/jodo_b200.h field by field) -----------------------------------
_P = ctypes.c_void_p
_I = ctypes.c_int
_F = ctypes.c_float
_Z = ctypes.c_size_t


PACK_ELEMS_PER_BLOCK = 2048        # JODO_PACK_ELEMS_PER_BLOCK


class PackItem(ctypes.Structure):
    _fields_ = [('src', _P), ('src_ld', _I), ('rows', _I), ('cols', _I), ('dst', _P), ('kind', _I), ('dst_ld', _I),
                ('nt', _I), ('k_pad', _I), ('row0', _I), ('col0', _I), ('scale', _F), ('add', _F)]


class PlanStruct(ctypes.Structure):
    _fields_ = [('B', _I), ('Nn', _I), ('n_tiles', _I), ('N', _I), ('node_mol', _P), ('node_dense', _P),
                ('mol_start', _P), ('row_g', _P), ('row_j', _P), ('row_meta', _P), ('tile_ngroups', _P), ('row_mol', _P),
                ('row_pair', _P)]


class EdgeEmbedArgs(ctypes.Structure):
    _fields_ = [('p', PlanStruct), ('edge_x', _P), ('cond_edge_x', _P), ('cond_x', _P), ('ch', _I), ('inn', _I),
                ('edge_th', _F), ('spatial_cut', _F), ('dist_flag', _P), ('tab', _P), ('ld_tab', _I), ('gbf', _P),
                ('w_img', _P), ('bias', _P), ('e32', _P), ('e16', _P), ('eh', _P), ('eh_tile_bytes', _Z), ('extra', _P),
                ('nonuni', _P), ('mol_bad', _P)]


class AttnArgs(ctypes.Structure):
    _fields_ = [('p', PlanStruct), ('e16', _P), ('pos', _P), ('qkv', _P), ('ldq', _I),
                ('tab', _P), ('ld_tab', _I), ('tab_off', _I), ('extra', _P), ('w_emb_img', _P),
                ('w0_img', _P), ('w1_img', _P), ('hnode', _P), ('nonuni', _P), ('gbf4', _F * 256), ('b_emb', _F * 64)]


class EdgeUpdateArgs(ctypes.Structure):
    _fields_ = [('p', PlanStruct), ('e32', _P), ('e16', _P), ('P', _P), ('ldp', _I),
                ('tab', _P), ('ld_tab', _I), ('tab_off', _I), ('r', _I), ('w3_img', _P), ('w4_img', _P), ('wl_img', _P),
                ('eh', _P), ('eh_tile_bytes', _Z), ('eh_col', _I), ('ce', _I), ('nonuni', _P),
                ('b_n2e', _F * 64), ('b3', _F * 256), ('b4', _F * 64), ('bl', _F * 16)]


class EquiArgs(ctypes.Structure):
    _fields_ = [('p', PlanStruct), ('e16', _P), ('pos_in', _P), ('pos_out', _P), ('AB', _P),
                ('ldab', _I), ('tab', _P), ('ld_tab', _I), ('tab_off', _I), ('extra', _P),
                ('win_img', _P), ('wc0_img', _P), ('w2_img', _P), ('w2_img32', _P), ('coord_scale', _F), ('nonuni', _P),
                ('gbf4', _F * 256), ('b0h', _F * 256), ('skip_if_uniform', _I)]


class EquiComposeItem(ctypes.Structure):
    _fields_ = [('w0', _P), ('b0', _P), ('wi', _P), ('bi', _P), ('w2', _P), ('tab_off', _I), ('wce_img', _P), ('ab_img', _P),
                ('ab_bias', _P), ('consts', _P)]


class EquiLinArgs(ctypes.Structure):
    _fields_ = [('p', PlanStruct), ('e16', _P), ('pos_in', _P), ('pos_out', _P), ('AB', _P), ('ldab', _I), ('extra', _P),
                ('win_img', _P), ('wce_img', _P), ('consts', _P), ('coord_scale', _F), ('nonuni', _P), ('gbf4', _F * 256)]


class EdgeHeadArgs(ctypes.Structure):
    _fields_ = [('p', PlanStruct), ('eh', _P), ('eh_tile_bytes', _Z), ('keh', _I), ('w0_img', _P), ('b0', _P),
                ('w2_img', _P), ('b2', _P), ('w4', _P), ('b4', _P), ('ch', _I), ('out_dense', _P), ('mol_bad', _P)]


class ImgLinearArgs(ctypes.Structure):
    _fields_ = [('Aimg', _P), ('M', _I), ('K', _I), ('Wimg', _P), ('bias', _P), ('N', _I), ('NT', _I), ('epi', _I),
                ('act_out', _I), ('aux', _P), ('ld_aux', _I), ('gate', _P), ('ld_gate', _I), ('row_mol', _P), ('nonuni', _P),
                ('skip_if_zero', _P), ('C32', _P), ('ldc32', _I), ('C16', _P), ('ldc16', _I), ('c16_piece_major', _I), ('Cimg', _P),
                ('cimg_k', _I), ('cimg_col0', _I), ('cimg_ncols', _I), ('Cimg2', _P), ('cimg2_k', _I), ('cimg2_col0', _I),
                ('cimg2_ncols', _I), ('dot_w', _P), ('dot_out', _P), ('ld_dot', _I), ('ln_valid', _P), ('ln_cols', _I),
                ('ln_off_shift', _I), ('ln_off_scale', _I)]


class WideEmbedArgs(ctypes.Structure):
    _fields_ = [('p', PlanStruct), ('edge_x', _P), ('cond_edge_x', _P), ('cond_x', _P), ('ch', _I), ('inn', _I), ('ed', _I),
                ('edge_th', _F), ('spatial_cut', _F), ('dist_flag', _P), ('tab', _P), ('ld_tab', _I), ('gbf', _P),
                ('ld_gbf', _I), ('img', _P), ('K', _I), ('extra', _P)]


class WideLnArgs(ctypes.Structure):
    _fields_ = [('M', _I), ('W', _I), ('Kimg', _I), ('x', _P), ('ldx', _I), ('xi', _P), ('y', _P), ('ldy', _I), ('yi', _P),
                ('y2', _P), ('ldy2', _I), ('y2i', _P), ('ybias', _P), ('tab', _P), ('ld_tab', _I), ('row_mol', _P),
                ('off_gate', _I), ('off_shift', _I), ('off_scale', _I), ('valid', _P), ('out32', _P), ('ldo', _I),
                ('out_img', _P), ('y_img', _P), ('x_f16', _I), ('y_f16', _I), ('nonuni', _P)]


class WideEquiArgs(ctypes.Structure):
    _fields_ = [('M', _I), ('D', _I), ('U', _P), ('ldu', _I), ('xi', _P), ('AB', _P), ('ldab', _I), ('row_g', _P), ('row_j', _P),
                ('row_mol', _P), ('tab', _P), ('ld_tab', _I), ('off_shift', _I), ('off_scale', _I), ('Wimg', _P), ('bias', _P),
                ('dot_w', _P), ('out', _P), ('ld_out', _I)]


class WideFfnArgs(ctypes.Structure):
    _fields_ = [('M', _I), ('ed', _I), ('H', _I), ('e32', _P), ('lde', _I), ('P', _P), ('ldp', _I), ('pair_i', _P), ('pair_j', _P),
                ('pair_mol', _P), ('n2e_bias', _P), ('tab', _P), ('ld_tab', _I), ('off_gate', _I), ('off_shift', _I),
                ('off_scale', _I), ('off_gate2', _I), ('w3_img', _P), ('b3', _P), ('w4_img', _P), ('b4', _P), ('img1', _P),
                ('k1', _I), ('col1', _I), ('img2', _P), ('k2', _I), ('col2', _I), ('nonuni', _P)]


class WideAttnArgs(ctypes.Structure):
    _fields_ = [('Nn', _I), ('D', _I), ('H', _I), ('X', _I), ('sc', _I), ('grp_row0', _P), ('grp_len', _P), ('row_j', _P),
                ('qkv', _P), ('ldq', _I), ('k_off', _I), ('v_off', _I), ('G', _P), ('ldg', _I), ('g1_off', _I),
                ('extra', _P), ('row_pair', _P), ('hnode', _P), ('max_gl', _I), ('mol_start', _P), ('B', _I), ('n_max', _I)]


def imglinear(Aimg, M, K, Wimg, bias, N, NT, epi=EPI_STORE, act_out=ACT_NONE, aux=None, gate=None, row_mol=None,
              C32=None, C16=None, Cimg=None, stream=None, tag=None, nonuni=0, skip_if_zero=0, cimg_place=None, Cimg2=None,
              cimg2_place=None, dot_w=None, dot_out=None, ln=None, ln_valid=None):
    """Persistent TMA-fed GEMM on an fp16 activation image (include/jodo_b200.h: jodo_imglinear).
    C32 / C16 are 2-D row-major views (stride(1) == 1) -- or C16 a contiguous 3-D [N/8, rows, 8] tensor for the
    piece-major layout the edge kernels gather from; Cimg a flat fp16 image buffer."""
    pm = C16 is not None and C16.dim() == 3          # piece-major fp16 output: tensor [N/8, rows, 8]
    a = ImgLinearArgs(dp(Aimg), M, K, dp(Wimg), dp(bias), N, NT, epi, act_out, dp(aux),
                      0 if aux is None else aux.stride(0), dp(gate), 0 if gate is None else gate.stride(0), dp(row_mol), nonuni,
                      skip_if_zero, dp(C32), 0 if C32 is None else C32.stride(0), dp(C16),
                      0 if C16 is None else (C16.shape[1] if pm else C16.stride(0)), 1 if pm else 0, dp(Cimg),
                      *(cimg_place or (0, 0, 0)), dp(Cimg2), *(cimg2_place or (0, 0, 0)), dp(dot_w), dp(dot_out),
                      0 if dot_out is None else dot_out.stride(0), dp(ln_valid), *(ln or (0, 0, 0)))      # ln = (cols, off_shift, off_scale)
    st = stream if stream is not None else stream_ptr()
    f = lib().jodo_imglinear
    check(_account(tag or 'jodo_imglinear', lambda: f(ctypes.byref(a), st)), 'jodo_imglinear')


def saturation_count(reset=True):
    """fp16 operand stores clamped at +-65504 since the last reset (include/jodo_b200.h: jodo_saturation_count); synchronises."""
    out = ctypes.c_ulonglong(0)
    check(lib().jodo_saturation_count(ctypes.byref(out), ctypes.c_int(1 if reset else 0)), 'jodo_saturation_count')
    return int(out.value)


def dp(t):
    """raw device address (int) of a tensor, 0 for None"""
    return 0 if t is None else t.data_ptr()


def plan_struct(plan):
    """The directed-edge plan (groups inside tiles), with the map onto the pair rows."""
    return PlanStruct(plan.B, plan.Nn, plan.n_tiles, plan.N, dp(plan.node_mol), dp(plan.node_dense),
                      dp(plan.mol_start), dp(plan.row_g), dp(plan.row_j), dp(plan.row_meta), dp(plan.tile_ngroups),
                      dp(plan.row_mol), dp(plan.row_pair))


def pair_plan_struct(plan):
    """The pair plan: rows = unordered pairs i < j (row_g = i, row_j = j), no groups, no pair map.  What the kernels
    that own the symmetric edge state run on (jodo_edge_embed, jodo_edge_update, jodo_edge_head)."""
    return PlanStruct(plan.B, plan.Nn, plan.n_pair_tiles, plan.N, dp(plan.node_mol), dp(plan.node_dense),
                      dp(plan.mol_start), dp(plan.pair_i), dp(plan.pair_j), dp(plan.pair_meta), dp(plan.pair_ngroups),
                      dp(plan.pair_mol), 0)


def call(name, *args, tag=None):
    f = getattr(lib(), name)
    check(_account(tag or name, lambda: f(*args)), name)
