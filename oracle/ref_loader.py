"""ORACLE tooling (test infrastructure): import the UNMODIFIED reference modules -- from /root/reference in
the build container, else from the byte-identical staged copy under oracle/_ref/reference
(oracle/stage_ref.py; git-ignored, travels to the GPU box).  Only tests/, smoke() and the reference arm
of bench.py import this; the product package never does.

The reference needs torch_geometric / torch_scatter / ml_collections, none installable offline;
oracle/shim provides pure-torch stand-ins for the handful of entry points the DGT hot path uses
(see oracle/README.md).  Nothing is copied out of the reference tree.
"""
import importlib
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(_HERE, '_ref', 'reference')
SHIM = os.path.join(_HERE, 'shim')


def _pick_root():
    env = os.environ.get('JODO_REFERENCE_ROOT')
    for cand in ([env] if env else []) + ['/root/reference', STAGED]:
        if cand and os.path.isdir(os.path.join(cand, 'models')):
            return cand
    return env or '/root/reference'


REF_ROOT = _pick_root()


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'models'))


def is_staged_copy():
    return os.path.abspath(REF_ROOT) == os.path.abspath(STAGED)


def load():
    """Returns a namespace with the reference's `models`, `sampling`, `mix_dpm_solver`,
    `diffusion.noise_schedule`, `utils` modules."""
    if not available():
        raise RuntimeError(f'reference tree not found at {REF_ROOT}')
    for p in (SHIM, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    ns = type('Ref', (), {})()
    ns.models = importlib.import_module('models')
    ns.mol_gnn = importlib.import_module('models.mol_gnn')
    ns.model_utils = importlib.import_module('models.utils')
    ns.noise_schedule = importlib.import_module('diffusion.noise_schedule')
    ns.utils = importlib.import_module('utils')
    ns.sampling = importlib.import_module('sampling')
    ns.mix_dpm_solver = importlib.import_module('mix_dpm_solver')
    ns.ema = importlib.import_module('models.ema')
    ns.node_distribution = importlib.import_module('models.node_distribution')
    return ns


def load_cond_gen():
    """The reference's property classifier modules by file path: cond_gen/model.py (EGNN) and cond_gen/utils.py
    (get_adj_matrix_fn).  Both import nothing but torch; the package __init__ is bypassed (it pulls in the property
    distribution, which needs the datasets)."""
    ns = type('CondGen', (), {})()
    for name in ('model', 'utils'):
        spec = importlib.util.spec_from_file_location('ref_cond_gen_' + name, os.path.join(REF_ROOT, 'cond_gen', name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        setattr(ns, name, mod)
    return ns


def load_datasets_config():
    """datasets/datasets_config.py by file path (the `datasets` package itself needs PyG + rdkit)."""
    spec = importlib.util.spec_from_file_location('ref_datasets_config', os.path.join(REF_ROOT, 'datasets', 'datasets_config.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_config(fname):
    """configs/<fname>.py -> ConfigDict (through the ml_collections shim)."""
    if SHIM not in sys.path:
        sys.path.insert(0, SHIM)
    spec = importlib.util.spec_from_file_location('refcfg_' + fname, os.path.join(REF_ROOT, 'configs', fname + '.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.get_config()
