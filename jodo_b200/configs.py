"""Config objects carrying the reference's config keys for the DGT hot path.

The reference reads ``config.model.*`` / ``config.data.*`` attributes of an
``ml_collections.ConfigDict`` (configs/vpsde_qm9_uncond_jodo.py:38-64,
configs/vpsde_geom_uncond_jodo.py:38-64, configs/vpsde_qm9_cond_jodo.py:38-65 of the
reference).  Our modules only use attribute access, so they accept either the
reference's ConfigDict or the plain ``Config`` below (which is what bench.py / tests
use on the GPU box, where the reference tree does not exist).
"""
from __future__ import annotations


class Config(dict):
    """Minimal attribute dict (same access pattern as ml_collections.ConfigDict)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def copy(self):
        out = Config()
        for k, v in self.items():
            out[k] = v.copy() if isinstance(v, Config) else v
        return out


def _base(name, atom_types, edge_ch, nf, n_layers, mlp_ratio, spatial_cut_off, max_node, fc_scale,
          info_name, eval_batch):
    c = Config()
    c.exp_type = 'vpsde_edge'
    c.pred_edge = True
    c.only_2D = False
    c.seed = 42
    c.data = Config(atom_types=atom_types, max_node=max_node, compress_edge=True, centered=True,
                    fc_scale=fc_scale, info_name=info_name)
    c.sde = Config(schedule='cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
    c.model = Config(
        name=name, pred_data=True, include_fc_charge=True, normalize_factors='1, 4, 4, 1',
        edge_ch=edge_ch, nf=nf, n_layers=n_layers, n_heads=16, dropout=0.1, cond_time=True,
        dist_gbf=True, gbf_name='CondGaussianLayer', self_cond=True, self_cond_type='ori',
        edge_quan_th=0., n_extra_heads=2, CoM=True, mlp_ratio=mlp_ratio,
        spatial_cut_off=spatial_cut_off, softmax_inf=True, trans_name='TransMixLayer')
    c.sampling = Config(method='ancestral', steps=1000, dpm_solver_method='singlestep_fixed',
                        dpm_solver_order=2)
    c.eval = Config(batch_size=eval_batch)
    return c


def qm9_uncond():
    """configs/vpsde_qm9_uncond_jodo.py (BASELINE configs 1 and 2)."""
    return _base('DGT_concat', 5, 2, 256, 8, 2, 2., 29, [-1., 1.], 'qm9_with_h', 2500)


def qm9_cond():
    """configs/vpsde_qm9_cond_jodo.py (BASELINE config 5): cond_DGT_concat, cond_ch=1."""
    c = _base('cond_DGT_concat', 5, 2, 256, 8, 2, 2., 29, [-1., 1.], 'qm9_second_half', 2500)
    c.model.cond_ch = 1
    return c


def qm9_cond_multi():
    """configs/vpsde_qm9_cond_multi_jodo.py: cond_DGT_concat conditioned on two properties (cond_ch = 2)."""
    c = qm9_cond()
    c.model.cond_ch = 2
    return c


def qm9_sim():
    """DGT_concat_sim (reference models/mol_gnn.py:949) on the QM9 config: the variant without adjacency heads."""
    c = qm9_uncond()
    c.model.name = 'DGT_concat_sim'
    return c


def geom_uncond(n_layers=8, nf=256):
    """configs/vpsde_geom_uncond_jodo.py; BASELINE config 3 quotes n_layers=8 (the file's
    default is 10), config 4 quotes nf=384."""
    return _base('DGT_concat', 16, 3, nf, n_layers, 4, 3., 181, [-2., 3.], 'geom_with_h_1', 512)


def moses_2d():
    """configs/vpsde_moses_2d_jodo.py: the 2-D-only model DGT_concat_2D (no coordinates: no distance features, no
    coordinate update, one adjacency head), reference models/mol_gnn.py:797-947."""
    c = _base('DGT_concat_2D', 7, 3, 256, 8, 2, 0., 27, None, 'moses', 2000)
    c.exp_type = 'vpsde'
    c.only_2D = True
    c.data = Config(atom_types=7, max_node=27, compress_edge=True, centered=True, info_name='moses')
    m = c.model
    m.include_fc_charge = False
    m.normalize_factors = '1, 2, 2, 1'
    m.time_dim = 1024
    m.n_extra_heads = 1
    for k in ('dist_gbf', 'gbf_name', 'CoM', 'spatial_cut_off'):
        del m[k]
    return c


def tiny(nf=64, n_layers=2, atom_types=5, edge_ch=2, mlp_ratio=2, cond=False):
    """Small architecture for fast CPU tests of the oracle (not a reference config)."""
    c = _base('cond_DGT_concat' if cond else 'DGT_concat', atom_types, edge_ch, nf, n_layers,
              mlp_ratio, 2., 29, [-1., 1.], 'qm9_with_h', 8)
    if cond:
        c.model.cond_ch = 1
    return c


NAMED = {
    'qm9_uncond': qm9_uncond,
    'qm9_cond': qm9_cond,
    'qm9_cond_multi': qm9_cond_multi,
    'qm9_sim': qm9_sim,
    'geom_l8': lambda: geom_uncond(8, 256),
    'geom_l10': lambda: geom_uncond(10, 256),
    'geom_large': lambda: geom_uncond(10, 384),
    'moses_2d': moses_2d,
}
