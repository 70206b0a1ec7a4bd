"""Shared test helpers: golden loading and oracle invocation."""
import os

import torch

from jodo_b200 import configs
from jodo_b200.params import param_spec, synth_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
FORWARD_CASES = ['qm9_first', 'qm9_first_default_init', 'qm9_selfcond', 'qm9_cond_ctx', 'geom_l8',
                 'geom_l10_first', 'geom_large', 'moses_2d', 'moses_2d_first', 'qm9_cond_multi', 'qm9_sim']


def load_golden(name):
    g = torch.load(os.path.join(GOLDEN, name + '.pt'), weights_only=False)
    cfg = configs.NAMED[g['config']]()
    return g, cfg


def golden_weights(g, cfg, dtype=torch.float32):
    return synth_state_dict(param_spec(cfg), dtype=dtype, **g['weights'])


def oracle_forward(sd, cfg, inp, dtype=torch.float64, collect=None):
    from oracle.dgt_dense import dgt_forward
    c = lambda x: None if x is None else x.to(dtype)
    sd = {k: v.to(dtype) for k, v in sd.items()}
    return dgt_forward(sd, cfg, c(inp['t']), c(inp['xh']), c(inp['node_mask']), c(inp['edge_mask']),
                       context=c(inp['context']), edge_x=c(inp['edge_x']), noise_level=c(inp['noise_level']),
                       cond_x=c(inp['cond_x']), cond_edge_x=c(inp['cond_edge_x']), collect=collect)
