#!/usr/bin/env python
"""Small-batch reverse chain: eager loop vs the CUDA-graph-captured step (sampler.sampling(graph=True)).
usage: python tools/bench_graph.py [B ...]   (QM9 uncond architecture, 200 reverse steps; the first graphed chain of a
process pays the one-time graph / RNG-state setup, so each mode runs twice and the second run is reported)"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jodo_b200 import configs, synth, sampler as S
from jodo_b200.model import MODELS

cfg = configs.NAMED['qm9_uncond']()
model = MODELS[cfg.model.name](cfg).cuda().eval()
grid = torch.linspace(0.9946, 1e-3, 1000)[::5]
for B in [int(a) for a in sys.argv[1:]] or [16, 64, 256, 1024]:
    b = synth.make_batch(cfg, B, seed=21)
    d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    out = {}
    for graph in (False, True):
        for rep in range(2):
            torch.manual_seed(1)
            smp = S.AncestralSampler(S.CosineVP(), grid)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            smp.sampling(model, d['xh'], d['node_mask'], d['edge_mask'], d['edge_x'], None, graph=graph)
            torch.cuda.synchronize()
            out[graph] = (time.perf_counter() - t0) / len(grid)
    print(f'B={B:5d}  eager {out[False] * 1e3:7.3f} ms/step ({B / out[False]:9.0f} mol-steps/s)   '
          f'graphed {out[True] * 1e3:7.3f} ms/step ({B / out[True]:9.0f} mol-steps/s)   x{out[False] / out[True]:.2f}')
