"""The C-ABI shared library loads without a GPU and exports every entry point include/jodo_b200.h declares; the ctypes
argument blocks mirror the header's structs; argument validation fails loudly before any launch."""
import ctypes
import os
import re

import pytest

from jodo_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'jodo_b200.h')


@pytest.fixture(scope='module')
def lib():
    build.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(jodo_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/jodo_b200.h but not exported'


def test_abi_version_and_error_string(lib):
    src = open(HEADER).read()
    ver = int(re.search(r'#define JODO_ABI_VERSION (\d+)', src).group(1))
    assert lib.jodo_abi_version() == ver
    lib.jodo_last_error_string.restype = ctypes.c_char_p
    assert isinstance(lib.jodo_last_error_string(), bytes)


def _struct_fields(name):
    src = open(HEADER).read()
    body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (name, name), src, flags=re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    out = []
    for stmt in body.split(';'):
        stmt = stmt.strip()
        if not stmt:
            continue
        for decl in stmt.split(','):
            m = re.search(r'([A-Za-z_][A-Za-z0-9_]*)\s*(\[\d+\])?\s*$', decl.strip())
            out.append(m.group(1))
    return out


@pytest.mark.parametrize('cname,pytype', [('jodo_plan', 'PlanStruct'), ('jodo_edge_embed_args', 'EdgeEmbedArgs'),
                                          ('jodo_attn_args', 'AttnArgs'), ('jodo_edge_update_args', 'EdgeUpdateArgs'),
                                          ('jodo_equi_args', 'EquiArgs'), ('jodo_edge_head_args', 'EdgeHeadArgs'),
                                          ('jodo_imglinear_args', 'ImgLinearArgs'),
                                          ('jodo_wide_embed_args', 'WideEmbedArgs'), ('jodo_wide_ln_args', 'WideLnArgs'),
                                          ('jodo_wide_attn_args', 'WideAttnArgs'), ('jodo_wide_equi_args', 'WideEquiArgs'),
                                          ('jodo_wide_ffn_args', 'WideFfnArgs'),
                                          ('jodo_equi_lin_args', 'EquiLinArgs'), ('jodo_equi_compose_item', 'EquiComposeItem'),
                                          ('jodo_pack_item', 'PackItem')])
def test_ctypes_structs_mirror_header(cname, pytype):
    assert [f[0] for f in getattr(_lib, pytype)._fields_] == _struct_fields(cname)


def test_bad_arguments_fail_before_any_launch(lib):
    """No GPU here: these calls must be rejected by validation (JODO_ERR_ARG), never reach the CUDA runtime."""
    lib.jodo_last_error_string.restype = ctypes.c_char_p
    assert lib.jodo_time_features(None, None, None, ctypes.c_int(0), None) == 1
    assert b'B <= 0' in lib.jodo_last_error_string()
    assert lib.jodo_attn(None, None) == 1
    assert lib.jodo_imglinear(None, None) == 1
    a = _lib.ImgLinearArgs()
    a.M, a.K, a.N, a.NT = 128, 40, 64, 64
    assert lib.jodo_imglinear(ctypes.byref(a), None) == 1 and b'multiple of 64' in lib.jodo_last_error_string()
    e = _lib.EdgeUpdateArgs()
    assert lib.jodo_edge_update(ctypes.byref(e), None) == 1
    # round-2 entry points: classifier row kernels, in-kernel noise, opt-in coordinate variants
    assert lib.jodo_egnn_edge_in(None, None, None, 0, 0, None, None, None) == 1
    assert lib.jodo_egnn_agg(None, None, None, 0, 0, None, ctypes.c_float(0), None, 0, 0, None) == 1
    assert lib.jodo_mol_sum(None, 0, 0, None, 0, None, 0, None) == 1
    assert lib.jodo_philox_normal(ctypes.c_ulonglong(0), ctypes.c_ulonglong(0), 0, 0, None, None) == 1
    assert lib.jodo_ancestral_update_philox(None, None, None, None, None, None, 0, 0, 0, 0, ctypes.c_float(0), ctypes.c_float(0),
                                            ctypes.c_float(0), None, ctypes.c_ulonglong(0), 0, None, None, None, None, None) == 1
    assert lib.jodo_equi_lin(None, None) == 1 and lib.jodo_equi_compose(None, 0, None, None, None) == 1
    assert lib.jodo_wide_equi(None, None) == 1
    assert lib.jodo_wide_edge_ffn(None, None) == 1
    f = _lib.WideFfnArgs()
    f.M, f.ed, f.H = 256, 128, 256              # ed = 128 is served by the unfused kernels
    assert lib.jodo_wide_edge_ffn(ctypes.byref(f), None) == 1 and b'built for ed' in lib.jodo_last_error_string()
    q = _lib.WideEquiArgs()
    q.M, q.D = 128, 320
    assert lib.jodo_wide_equi(ctypes.byref(q), None) == 1 and b'D = 256 and D = 384' in lib.jodo_last_error_string()
    i = _lib.ImgLinearArgs()
    i.M, i.K, i.N, i.NT = 128, 64, 128, 128
    i.Aimg = i.Wimg = 128
    i.Cimg, i.cimg_k, i.cimg_col0, i.cimg_ncols = 128, 192, 4, 96
    assert lib.jodo_imglinear(ctypes.byref(i), None) == 1 and b'placement' in lib.jodo_last_error_string()
    # the LayerNorm + modulation epilogue is built for N = NT = 128 with the image output alone; NT = 192 for the dots-only mode
    n = _lib.ImgLinearArgs()
    n.M, n.K, n.N, n.NT, n.epi = 128, 64, 256, 128, _lib.EPI_LN_MOD
    n.Aimg = n.Wimg = n.Cimg = 128
    assert lib.jodo_imglinear(ctypes.byref(n), None) == 1 and b'JODO_EPI_LN_MOD' in lib.jodo_last_error_string()
    n.N, n.NT, n.epi = 192, 192, _lib.EPI_STORE
    assert lib.jodo_imglinear(ctypes.byref(n), None) == 1 and b'192: fused row dots only' in lib.jodo_last_error_string()
    assert lib.jodo_row0_linear(None, 64, None, 64, 64, None, 0, 0, None, None, None, None) == 1
    # wide path (nf = 384)
    assert lib.jodo_wide_ln(None, None) == 1 and lib.jodo_wide_attn(None, None) == 1 and lib.jodo_wide_embed_in(None, None) == 1
    w = _lib.WideLnArgs()
    w.M, w.W, w.Kimg = 128, 100, 128
    assert lib.jodo_wide_ln(ctypes.byref(w), None) == 1 and b'bad sizes' in lib.jodo_last_error_string()
    t = _lib.WideAttnArgs()
    t.Nn, t.D, t.H, t.X, t.sc, t.max_gl = 10, 384, 16, 2, 27, 300
    assert lib.jodo_wide_attn(ctypes.byref(t), None) == 1 and b'max_gl' in lib.jodo_last_error_string()
    t.max_gl, t.H = 80, 64
    assert lib.jodo_wide_attn(ctypes.byref(t), None) == 1 and b'bad sizes' in lib.jodo_last_error_string()
    assert lib.jodo_wide_put(None, 0, 0, 0, None, None, 0, 0, None, 0, 0, None, 0, 0, None) == 1
    assert lib.jodo_wide_dist(None, None, None, 0, 0, None, 0, 0, None, 0, 0, None, 0, 0, None) == 1
    assert lib.jodo_wide_equi_out(None, None, None, None, 0, 0, None, None, 0, ctypes.c_float(0), None, None, 0, None) == 1
    assert lib.jodo_wide_head_out(None, None, 0, 0, None, None, 0, 0, None, None) == 1


def test_unsupported_sizes_raise_by_name():
    """nf = 256 -> fused kernels, multiples of 128 up to 512 -> wide path, anything else names the key."""
    from jodo_b200 import configs
    from jodo_b200.model import MODELS
    cfg = configs.NAMED['geom_large']()
    assert MODELS[cfg.model.name](cfg).wide
    cfg.model.nf = 320
    with pytest.raises(NotImplementedError, match='model.nf'):
        MODELS[cfg.model.name](cfg)


def test_product_path_has_no_cpu_fallback():
    """The drop-in module refuses CPU tensors instead of silently computing elsewhere."""
    import torch
    from jodo_b200 import configs, synth
    from jodo_b200.model import MODELS
    cfg = configs.NAMED['qm9_uncond']()
    model = MODELS[cfg.model.name](cfg).eval()
    b = synth.make_batch(cfg, 2, seed=0)
    with pytest.raises(_lib.JodoError):
        model(b['t'], b['xh'], b['node_mask'], b['edge_mask'], edge_x=b['edge_x'], noise_level=b['noise_level'])
    model.train()
    with pytest.raises(RuntimeError):
        model(b['t'], b['xh'], b['node_mask'], b['edge_mask'], edge_x=b['edge_x'], noise_level=b['noise_level'])
    # nothing under the product package imports the oracle
    import glob
    for f in glob.glob(os.path.join(ROOT, 'jodo_b200', '*.py')):
        assert 'oracle' not in open(f).read().replace('the oracle', '').replace('against the oracle', ''), f
