#!/usr/bin/env python
"""Microbenchmark of jodo_imglinear on the per-edge GEMM shapes of the wide path (nf = 384).  usage: python tools/bench_gemm_wide.py [M]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jodo_b200 import _lib
from jodo_b200.pack import weight_image_h

M = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
mt = (M + 127) // 128
dev = 'cuda'
SETS = 3
shapes = {  # name: (K, N, NT, outputs, act)
    'g01 nt128 c16 tanh': (128, 768, 128, ('C16',), _lib.ACT_TANH),
    'g01 nt256 c16 tanh': (128, 768, 256, ('C16',), _lib.ACT_TANH),
    'g01 nt128 c16 none': (128, 768, 128, ('C16',), None),
    'g01 nt128 c16 silu': (128, 768, 128, ('C16',), _lib.ACT_SILU),
    'g01 nt128 c32 tanh': (128, 768, 128, ('C32',), _lib.ACT_TANH),
    'g01 nt128 img tanh': (128, 768, 128, ('Cimg',), _lib.ACT_TANH),
    'equi_in c32': (192, 384, 128, ('C32',), None),
    'equi_in c16': (192, 384, 128, ('C16',), None),
    'c0 img silu': (384, 384, 128, ('Cimg',), _lib.ACT_SILU),
    'ff3 img silu': (128, 384, 128, ('Cimg',), _lib.ACT_SILU),
    'emb c32': (192, 128, 128, ('C32',), None),
    'c0 dots nt128': (384, 384, 128, ('dot',), _lib.ACT_SILU),
    'c0 dots nt192': (384, 384, 192, ('dot',), _lib.ACT_SILU),
    'c0 dots D256 nt256': (256, 256, 256, ('dot',), _lib.ACT_SILU),
}
only = os.environ.get('ONLY')
if only:
    shapes = {k: v for k, v in shapes.items() if only in k}
dw = None
for name, (K, N, NT, outs, act) in shapes.items():
    W = weight_image_h(torch.randn(N, K, device=dev) / K ** 0.5, NT)
    b = torch.randn(N, device=dev)
    sets = []
    for s in range(SETS):
        d = dict(A=torch.randn(mt * 128 * K, device=dev).half())
        if 'C16' in outs: d['C16'] = torch.empty(M, N, device=dev, dtype=torch.float16)
        if 'C32' in outs: d['C32'] = torch.empty(M, N, device=dev)
        if 'Cimg' in outs: d['Cimg'] = torch.empty(mt * 128 * N, device=dev, dtype=torch.float16)
        if 'dot' in outs: d['dot_out'] = torch.zeros(M, 64, device=dev); d['dot_w'] = torch.randn(3, N, device=dev)
        sets.append(d)

    def run(d):
        kw = {k: d[k] for k in ('C16', 'C32', 'Cimg', 'dot_out', 'dot_w') if k in d}
        if act is not None:
            kw.update(epi=_lib.EPI_ACT, act_out=act)
        _lib.imglinear(d['A'], M, K, W, b, N, NT, **kw)
    try:
        for d in sets: run(d)
    except _lib.JodoError as e:
        print(f'{name:20s} refused: {e}')
        continue
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        for d in sets: run(d)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * SETS)
    byt = mt * 128 * K * 2 + sum({'C16': 2, 'C32': 4, 'Cimg': 2}.get(o, 0) * M * N for o in outs)
    fl = 2.0 * M * K * N
    print(f'{name:20s} K={K:4d} N={N:4d} NT={NT:3d}  {us:8.1f} us  {byt / us / 1e3:7.1f} GB/s  {fl / us / 1e6:7.1f} TFLOP/s')
