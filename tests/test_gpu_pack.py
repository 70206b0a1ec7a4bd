"""jodo_pack_weights (C ABI, one kernel for the whole model) against the CPU emulation of the same recorded items:
every packed piece bit-identical, by-value tables identical, a (re)pack costs a handful of launches and ~1 ms."""
import time

import pytest
import torch

from jodo_b200 import _lib, configs
from jodo_b200.pack import pack_model
from jodo_b200.params import dims_from_config, param_spec, synth_state_dict

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name,fused', [('qm9_uncond', True), ('qm9_cond', True), ('geom_l10', True), ('geom_large', False),
                                        ('moses_2d', False)])
def test_device_packer_equals_emulation(name, fused):
    cfg = configs.NAMED[name]()
    d = dims_from_config(cfg)
    sd = synth_state_dict(param_spec(cfg), seed=5, perturb=True)
    sd['e_block_0.ff_linear1.weight'][0, :4] = torch.tensor([7e4, -7e4, 1e-9, 65504.0])      # saturation / underflow / max
    ref = pack_model(sd, d, 'cpu', fused=fused)
    dev = {k: v.cuda() for k, v in sd.items()}
    l0 = _lib.LAUNCHES
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pk = pack_model(dev, d, 'cuda', fused=fused)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert _lib.LAUNCHES - l0 == 1                                  # ONE jodo_pack_weights launch through the C ABI
    print(f'{name}: {len(pk._off)} pieces, {pk.buf.numel() * 4 / 2 ** 20:.1f} MiB, packed in {dt * 1e3:.2f} ms')
    assert set(pk._off) == set(ref._off) and pk.meta == ref.meta
    for k in ref._off:
        a, b = ref[k], pk[k].cpu()
        if k == 'hcat.b':                                           # batched einsum: summation order differs CPU vs GPU
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
            continue
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), k
    for k, v in ref.host.items():
        assert list(v) == list(pk.host[k]), k


def test_repack_after_weight_update_is_cheap():
    from jodo_b200.model import MODELS
    cfg = configs.NAMED['qm9_uncond']()
    m = MODELS[cfg.model.name](cfg).cuda().eval()
    m._weights()
    with torch.no_grad():
        next(m.parameters()).mul_(1.0)                              # bumps the version counter -> repack
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m._weights()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'repack after an in-place update: {dt * 1e3:.1f} ms')
    assert dt < 0.5
