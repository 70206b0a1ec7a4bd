#!/usr/bin/env python
"""Per-phase cycle counts of CTA 0 of an edge kernel built with -DJODO_PHASE_TIMING (debug builds only).
usage: JODO_NVCC_EXTRA=-DJODO_PHASE_TIMING python tools/phase_timing.py [qm9|geom]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jodo_b200 import build
build.build()
import torch
from jodo_b200 import _lib, configs, sampler as S, synth
from jodo_b200.model import create_model
wl = sys.argv[1] if len(sys.argv) > 1 else 'qm9'
cfg = configs.NAMED['qm9_uncond' if wl == 'qm9' else 'geom_l8']()
model = create_model(cfg, 'cuda')
b = synth.make_batch(cfg, 2500 if wl == 'qm9' else 512, seed=42, max_n=None if wl == 'qm9' else 80)
dev = 'cuda'
nm, em = b['node_mask'].to(dev), b['edge_mask'].to(dev)
smp = S.AncestralSampler(S.CosineVP(), torch.linspace(0.9946, 1e-3, 1000), generator=torch.Generator(device=dev).manual_seed(1))
x, ex, cx, cex = b['xh'].to(dev), b['edge_x'].to(dev), None, None
for i in range(3):
    x, ex, _, _, cx, cex = smp.step(model, i, x, ex, nm, em, cx, cex)
L = _lib.lib()
for name in ('equi_lin', 'equi2', 'equi', 'attn', 'edge_update'):
    fn = getattr(L, f'jodo_debug_{name}_phases', None)
    if fn is None:
        continue
    buf = (ctypes.c_longlong * 16)()
    fn(None, 1)
    x, ex, _, _, cx, cex = smp.step(model, 3, x, ex, nm, em, cx, cex)
    fn(buf, 0)
    v = list(buf)
    tot = sum(v)
    print(name, 'total cycles CTA0 over', cfg.model.n_layers, 'launches:', tot)
    for i, c in enumerate(v):
        if c:
            print(f'   phase {i:2d}: {c:10d}  {100.0 * c / tot:5.1f}%')
