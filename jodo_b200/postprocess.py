"""Post-processing of the sampler's final state into integer molecules.

Mirrors the reference's ``post_process`` (sampling.py:53-97, with the inverse data scaler of utils.py:71-105) and
``mol_process`` (sampling.py:12-32).  ``post_process`` is a handful of tensor ops and runs wherever its inputs live;
``mol_process`` in the reference issues three to four ``.cpu()`` copies PER MOLECULE (10 000 device syncs for a
2500-molecule batch) -- here the four tensors cross to the host once and are sliced there.  Integer outputs are
bit-exact against the reference (tests/test_postprocess.py)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def normalize_factors(config):
    nf = config.model.normalize_factors
    nf = [int(v) for v in nf.split(',')] if isinstance(nf, str) else list(nf)
    if len(nf) == 3:
        nf.append(1)
    return nf                                   # pos, atom type, formal charge, edge


def inverse_scale(config, pos, atom_type, fc_charge, node_mask, edge_type=None, edge_mask=None):
    """utils.get_data_inverse_scaler(config)(...) of the reference (utils.py:88-103)."""
    pos_norm, atom_norm, fc_norm, edge_norm = normalize_factors(config)
    centered = config.data.centered
    if pos is not None:                         # 2-D models carry no coordinates (utils.py:96-97)
        pos = pos * pos_norm * node_mask
    atom_type = atom_type * atom_norm
    fc_charge = fc_charge * fc_norm * node_mask
    if centered:
        atom_type = (atom_type + 1.) / 2. * node_mask
    if edge_type is None:
        return pos, atom_type, fc_charge
    edge_type = edge_type * edge_norm
    if centered:
        edge_type = (edge_type + 1.) / 2.
    B, N = node_mask.shape[0], node_mask.shape[1]
    edge_type = edge_type * edge_mask.reshape(B, N, N, 1)
    return pos, atom_type, fc_charge, edge_type


def post_process(config, xh, node_mask, edge_x=None, edge_mask=None):
    """sampling.post_process (sampling.py:53-97): unnormalise, one-hot atom types, rounded charges, bond orders
    0..4 (0 none, 1-3 single/double/triple, 4 aromatic) from the compressed (exist, order[, aromatic]) channels."""
    atom_types = int(config.data.atom_types)
    include_charge = bool(config.model.include_fc_charge)
    pos = xh[:, :, :3]
    if include_charge:
        h_int, h_cat = xh[:, :, -1:], xh[:, :, 3:-1]
    else:
        h_int, h_cat = torch.zeros(0, device=xh.device), xh[:, :, 3:]
    assert h_cat.shape[-1] == atom_types
    if edge_x is not None:
        pos, h_cat, h_int, h_edge = inverse_scale(config, pos, h_cat, h_int, node_mask, edge_x, edge_mask)
    else:
        pos, h_cat, h_int = inverse_scale(config, pos, h_cat, h_int, node_mask)
    h_cat = F.one_hot(torch.argmax(h_cat, dim=2), atom_types) * node_mask
    h_int = torch.round(h_int).long() * node_mask
    if edge_x is None:
        return pos, h_cat, h_int
    return pos, h_cat, h_int, _bond_orders(config, h_edge)


def _bond_orders(config, h_edge):
    """Unnormalised edge channels -> bond orders 0..4 (sampling.py:75-95 = :120-142)."""
    if config.data.compress_edge:
        exist = (h_edge[..., 0] >= 0.5).to(h_edge.dtype)
        t = h_edge[..., 1] * 3.
        order = torch.zeros_like(t)
        order = torch.where(t >= 0.5, torch.ones_like(t), order)
        order = torch.where(t >= 1.5, torch.full_like(t, 2.), order)
        order = torch.where(t >= 2.5, torch.full_like(t, 3.), order)
        order = exist * order
        if h_edge.size(-1) == 3:
            arom = exist * (h_edge[..., 2] >= 0.5).to(h_edge.dtype)
            order = torch.where((arom > 0.) & (order == 0.), torch.full_like(order, 4.), order)
        h_edge = order
    else:
        any_on = torch.sum(h_edge > 0.5, dim=-1) != 0
        h_edge = any_on * (torch.argmax(h_edge, dim=-1) + 1.0)
    return h_edge


def post_process_2d(config, xh, node_mask, edge_x, edge_mask):
    """sampling.post_process_2D (sampling.py:100-144): the same without coordinates (xh = atom features [+ charge])."""
    atom_types = int(config.data.atom_types)
    if bool(config.model.include_fc_charge):
        h_int, h_cat = xh[:, :, -1:], xh[:, :, :-1]
    else:
        h_int, h_cat = torch.zeros(0, device=xh.device), xh
    assert h_cat.shape[-1] == atom_types
    _, h_cat, h_int, h_edge = inverse_scale(config, None, h_cat, h_int, node_mask, edge_x, edge_mask)
    h_cat = F.one_hot(torch.argmax(h_cat, dim=2), atom_types) * node_mask
    h_int = torch.round(h_int).long() * node_mask
    return h_cat, h_int, _bond_orders(config, h_edge)


def mol_process_2d(one_hot, formal_charges, n_nodes, edge_types):
    """sampling.mol_process_2D (sampling.py:35-50) with one device-to-host transfer per tensor: list of
    (None, atom_type, edge_type, fc)."""
    atom_type_all = one_hot.argmax(2).detach().cpu()
    edge_all = edge_types.detach().cpu()
    n = [int(v) for v in n_nodes]
    if formal_charges.shape[-1] != 0:
        fc_all = formal_charges[..., 0].long().detach().cpu()
        fcs = [fc_all[i, :n[i]] for i in range(len(n))]
    else:
        fc_all = formal_charges.detach().cpu()
        fcs = [fc_all[i][:n[i]] for i in range(len(n))]
    return [(None, atom_type_all[i, :n[i]], edge_all[i, :n[i], :n[i]], fcs[i]) for i in range(len(n))]


def mol_process(one_hot, x, formal_charges, n_nodes, edge_types=None):
    """sampling.mol_process (sampling.py:12-32) with ONE device-to-host transfer per tensor instead of one per
    molecule.  Returns the same list of (pos, atom_type, edge_type, fc) CPU tensors."""
    atom_type_all = one_hot.argmax(2).detach().cpu()
    pos_all = x.detach().cpu()
    n = [int(v) for v in n_nodes]
    if edge_types is None:
        return [(pos_all[i, :n[i]], atom_type_all[i, :n[i]]) for i in range(len(n))]
    edge_all = edge_types.detach().cpu()
    if formal_charges.shape[-1] != 0:
        fc_all = formal_charges[..., 0].long().detach().cpu()
    else:
        fc_all = formal_charges.detach().cpu()
    out = []
    for i in range(len(n)):
        fc = fc_all[i, :n[i]] if formal_charges.shape[-1] != 0 else fc_all[i][:n[i]]
        out.append((pos_all[i, :n[i]], atom_type_all[i, :n[i]], edge_all[i, :n[i], :n[i]], fc))
    return out
