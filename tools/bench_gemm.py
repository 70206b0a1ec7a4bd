#!/usr/bin/env python
"""Microbenchmark of jodo_imglinear on the per-atom GEMM shapes of one DGT block (QM9 B=2500: 45105 atoms).
Rotates over buffer sets larger than L2.  usage: python tools/bench_gemm.py [M]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jodo_b200 import _lib
from jodo_b200.pack import weight_image_h

M = int(sys.argv[1]) if len(sys.argv) > 1 else 45105
mt = (M + 127) // 128
dev = 'cuda'
SETS = 6
shapes = {  # name: (K, N, NT, outputs)
    'qkv': (256, 768, 256, ('C16',)),
    'n2e': (256, 64, 64, ('C16',)),
    'ff1': (256, 512, 256, ('Cimg',)),
    'ff2': (512, 256, 256, ('C32', 'Cimg', 'gated')),
    'ff2_nt128': (512, 256, 128, ('C32', 'Cimg', 'gated')),
    'qkv_nt128': (256, 768, 128, ('C16',)),
    'ff1_nt128': (256, 512, 128, ('Cimg',)),
    'ab_nt128': (256, 512, 128, ('C16',)),
    'ab': (256, 512, 256, ('C16',)),
    'node_l': (256, 64, 64, ('C32',)),
}
for name, (K, N, NT, outs) in shapes.items():
    W = weight_image_h(torch.randn(N, K, device=dev) / K ** 0.5, NT)
    b = torch.randn(N, device=dev)
    sets = []
    for s in range(SETS):
        d = dict(A=torch.randn(mt * 128 * K, device=dev).half())
        if 'C16' in outs: d['C16'] = torch.empty(M, N, device=dev, dtype=torch.float16)
        if 'C32' in outs: d['C32'] = torch.empty(M, N, device=dev)
        if 'Cimg' in outs: d['Cimg'] = torch.empty(mt * 128 * N, device=dev, dtype=torch.float16)
        if 'gated' in outs:
            d['aux'] = torch.randn(M, N, device=dev)
        sets.append(d)
    gate = torch.randn(2500, N, device=dev)
    mol = torch.randint(0, 2500, (M,), device=dev, dtype=torch.int32).sort().values.int()

    def run(d):
        kw = {k: d[k] for k in ('C16', 'C32', 'Cimg') if k in d}
        if 'gated' in outs:
            kw.update(epi=_lib.EPI_GATED_RES, aux=d['aux'], gate=gate, row_mol=mol)
        _lib.imglinear(d['A'], M, K, W, b, N, NT, **kw)
    for d in sets: run(d)
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        for d in sets: run(d)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * SETS)
    byt = mt * 128 * K * 2 + sum({'C16': 2, 'C32': 4, 'Cimg': 2}.get(o, 0) * M * N for o in outs) + (4 * M * N if 'gated' in outs else 0)
    fl = 2.0 * M * K * N
    print(f'{name:8s} K={K:4d} N={N:4d} NT={NT:3d}  {us:8.1f} us  {byt / us / 1e3:7.1f} GB/s  {fl / us / 1e6:7.1f} TFLOP/s')
