"""ORACLE tooling (test infrastructure): import the UNMODIFIED reference modules in the build
container.  /root/reference does not exist on the GPU box; nothing that runs there imports this.

The reference needs torch_geometric / torch_scatter / ml_collections, none installable offline;
oracle/shim provides pure-torch stand-ins for the handful of entry points the DGT hot path uses
(see oracle/README.md).  Nothing is copied out of the reference tree.
"""
import importlib
import importlib.util
import os
import sys

REF_ROOT = os.environ.get('JODO_REFERENCE_ROOT', '/root/reference')
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shim')


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'models'))


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
    return ns


def load_config(fname):
    """configs/<fname>.py -> ConfigDict (through the ml_collections shim)."""
    if SHIM not in sys.path:
        sys.path.insert(0, SHIM)
    spec = importlib.util.spec_from_file_location('refcfg_' + fname, os.path.join(REF_ROOT, 'configs', fname + '.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.get_config()
