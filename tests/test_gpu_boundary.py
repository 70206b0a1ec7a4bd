"""Boundary contract of the drop-in module on the GPU: in-place weight mutation (optimizer / EMA,
reference models/ema.py:52-55) invalidates the packed images; the call works on a non-default stream; the reference's
single-device DataParallel wrapper (models/utils.py:27) passes the call through; several node masks can alternate;
outputs are fresh tensors that later calls do not overwrite (the sampler keeps them as cond_x, sampling.py:555)."""
import pytest
import torch

from helpers import golden_weights, load_golden, oracle_forward
from jodo_b200.model import MODELS

pytestmark = pytest.mark.gpu
TOL = 5e-3


def _setup(name='qm9_selfcond'):
    g, cfg = load_golden(name)
    sd = golden_weights(g, cfg)
    model = MODELS[cfg.model.name](cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in g['inputs'].items()}
    return g, cfg, sd, model, inp


def _call(model, inp):
    return model(inp['t'], inp['xh'], inp['node_mask'], inp['edge_mask'], context=inp['context'], edge_x=inp['edge_x'],
                 noise_level=inp['noise_level'], cond_x=inp['cond_x'], cond_edge_x=inp['cond_edge_x'])


def _rel(a, b):
    return float((a.double().cpu() - b).abs().max() / b.abs().max())


def _scale_some(model, f, through_data):
    with torch.no_grad():
        for n, p in model.named_parameters():
            if 'ff_linear1.weight' in n or 'coord_norm.scale' in n:
                (p.data if through_data else p).mul_(f)


def test_in_place_weight_update_repacks():
    g, cfg, sd, model, inp = _setup()
    x0, e0 = _call(model, inp)
    _scale_some(model, 1.5, through_data=False)               # what an optimizer step / load_state_dict do
    x1, e1 = _call(model, inp)
    assert float((x1 - x0).abs().max()) > 1e-4                # the new weights are in effect
    sd2 = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ox, oe = oracle_forward(sd2, cfg, g['inputs'], torch.float64)
    assert _rel(x1, ox) < TOL and _rel(e1, oe) < TOL


def test_data_writes_need_refresh_and_are_caught():
    """Reference models/ema.py:55 writes through param.data, which bypasses the version counter."""
    from jodo_b200 import _lib
    from jodo_b200.model import watch_data_writers
    g, cfg, sd, model, inp = _setup()
    x0, _ = _call(model, inp)
    _scale_some(model, 1.5, through_data=True)
    model.refresh_weights()
    x1, _ = _call(model, inp)
    assert float((x1 - x0).abs().max()) > 1e-4

    class Ema:                                                # same shape as the reference's writer
        def copy_to(self, parameters, f):
            for p in parameters:
                p.data.mul_(f)

        def restore(self, parameters):
            pass
    watch_data_writers(Ema)
    Ema().copy_to([p for n, p in model.named_parameters() if 'ff_linear1.weight' in n], 1 / 1.5)
    x2, _ = _call(model, inp)
    assert float((x2 - x1).abs().max()) > 1e-4

    # backstop: an unannounced .data write is reported by a later call instead of being used silently forever
    _scale_some(model, 1.25, through_data=True)
    with pytest.raises(_lib.JodoError):
        for _ in range(40):
            _call(model, inp)
            torch.cuda.synchronize()
    x3, _ = _call(model, inp)                                 # the error dropped the stale images
    assert float((x3 - x2).abs().max()) > 1e-4


def test_non_default_stream_and_fresh_outputs():
    g, cfg, sd, model, inp = _setup()
    x_ref, e_ref = _call(model, inp)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        x1, e1 = _call(model, inp)
        keep_x, keep_e = x1, e1                               # the sampler keeps these as the next cond_x / cond_edge_x
        inp2 = dict(inp, xh=inp['xh'] * 0.5)
        x2, e2 = _call(model, inp2)                            # a later call must not overwrite the earlier outputs
    s.synchronize()
    assert torch.equal(keep_x, x_ref) and torch.equal(keep_e, e_ref)
    assert x2.data_ptr() != keep_x.data_ptr() and float((x2 - keep_x).abs().max()) > 0


def test_data_parallel_wrapper_single_device():
    g, cfg, sd, model, inp = _setup()
    x_ref, e_ref = _call(model, inp)
    dp = torch.nn.DataParallel(model, device_ids=[0])
    x, e = _call(dp, inp)
    assert torch.equal(x, x_ref) and torch.equal(e, e_ref)
    # the reference's checkpoints carry the 'module.' prefix of this wrapper (utils.py:17, strict=True)
    dp.load_state_dict({'module.' + k: v for k, v in sd.items()}, strict=True)


def test_alternating_node_masks():
    g, cfg, sd, model, inp = _setup()
    ga, _ = load_golden('qm9_first')
    inp_a = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in ga['inputs'].items()}
    outs = []
    for _ in range(3):
        outs.append((_call(model, inp), _call(model, inp_a)))
    for (x, e), (xa, ea) in outs[1:]:
        assert torch.equal(x, outs[0][0][0]) and torch.equal(e, outs[0][0][1])
        assert torch.equal(xa, outs[0][1][0]) and torch.equal(ea, outs[0][1][1])


def test_rejects_inconsistent_edge_mask():
    g, cfg, sd, model, inp = _setup()
    bad = dict(inp, edge_mask=torch.ones_like(inp['edge_mask']))
    with pytest.raises(ValueError):
        _call(model, bad)
