"""Weight packing: reference parameter tensors -> operand images the kernels consume.

An operand image is the K-major SWIZZLE_128B layout described in csrc/common.cuh: 32 fp32 columns
(128 bytes) per row per chunk, the 16-byte piece p of row r stored at slot p ^ (r & 7)."""
from __future__ import annotations

import torch


def round_tf32(w: torch.Tensor) -> torch.Tensor:
    """Round fp32 to the nearest tf32 (10 explicit mantissa bits), ties away from zero like cvt.rna."""
    bits = w.contiguous().view(torch.int32)
    bits = (bits + 0x1000) & ~0x1FFF
    return bits.view(torch.float32)


def pad2(w: torch.Tensor, n: int, k: int) -> torch.Tensor:
    out = torch.zeros((n, k), dtype=torch.float32, device=w.device)
    out[:w.shape[0], :w.shape[1]] = w
    return out


def ceil_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def weight_image(w: torch.Tensor, nt: int | None = None, tf32: bool = True) -> torch.Tensor:
    """W [N, K] (N % nt == 0, K % 32 == 0) -> flat image [N/nt][K/32][nt][8 slots][4]."""
    n, k = w.shape
    nt = n if nt is None else nt
    assert n % nt == 0 and k % 32 == 0 and nt % 8 == 0, (n, k, nt)
    w = w.to(torch.float32)
    if tf32:
        w = round_tf32(w)
    x = w.reshape(n // nt, nt, k // 32, 8, 4).permute(0, 2, 1, 3, 4)          # [tile, chunk, row, piece, 4]
    rows = torch.arange(nt, device=w.device)
    slots = torch.arange(8, device=w.device)
    src_piece = slots[None, :] ^ (rows[:, None] & 7)                           # piece stored in (row, slot)
    idx = src_piece[None, None, :, :, None].expand(n // nt, k // 32, nt, 8, 4)
    return torch.gather(x, 3, idx).contiguous().reshape(-1)


def image_to_matrix(img: torch.Tensor, rows: int, k: int) -> torch.Tensor:
    """Inverse of the activation/edge image layout for ONE tile: [k/32][rows][8 slots][4] -> [rows, k]."""
    x = img.reshape(k // 32, rows, 8, 4)
    r = torch.arange(rows, device=img.device)
    slots = torch.arange(8, device=img.device)
    slot_of_piece = slots[None, :] ^ (r[:, None] & 7)                          # slot holding piece p of row r
    idx = slot_of_piece[None, :, :, None].expand(k // 32, rows, 8, 4)
    y = torch.gather(x, 2, idx)                                                # [chunk, row, piece, 4]
    return y.permute(1, 0, 2, 3).reshape(rows, k)


def matrix_to_image(m: torch.Tensor) -> torch.Tensor:
    """[rows, k] -> one tile image [k/32][rows][8 slots][4] (no tf32 rounding)."""
    rows, k = m.shape
    return weight_image(m, rows, tf32=False)
