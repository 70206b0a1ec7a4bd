"""ORACLE (test infrastructure): Philox4x32-10 counter-based generator + Box-Muller, numpy restatement of what
csrc/sampler_kernels.cu computes when the sampler draws its noise in-kernel (``noise='philox'``).

The reference draws its noise with torch.randn (models/utils.py:67-99); an in-kernel generator is a different random
stream by construction, so this mode has its own parity chain: (1) this file against the published known-answer vectors
of Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 kat_vectors), (2) the
kernel's normals against this file, (3) the fused update with in-kernel noise against the same update fed with these
normals, (4) distribution checks.  Only tests/ import this."""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter, key):
    """counter [..., 4] uint32, key [..., 2] uint32 (broadcastable) -> [..., 4] uint32."""
    c = np.asarray(counter, dtype=np.uint64)
    k = np.asarray(key, dtype=np.uint64)
    c0, c1, c2, c3 = (c[..., i] for i in range(4))
    k0, k1 = np.broadcast_to(k[..., 0], c0.shape).copy(), np.broadcast_to(k[..., 1], c0.shape).copy()
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def normals4(idx, step, stream, seed):
    """Four standard normals per counter (idx, 0, step, stream) under key = seed (64 bit): the kernel's transform,
    u = r 2^-32 + 2^-33 in (0, 1], z = sqrt(-2 ln u1) (cos, sin)(2 pi u2), evaluated in float64."""
    idx = np.asarray(idx, dtype=np.uint64)
    ctr = np.stack([idx & np.uint64(0xFFFFFFFF), idx >> np.uint64(32), np.full_like(idx, step), np.full_like(idx, stream)], axis=-1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64)
    r = philox4x32_10(ctr, key).astype(np.float64)
    u = r * 2.0 ** -32 + 2.0 ** -33
    ra, rb = np.sqrt(-2.0 * np.log(u[..., 0])), np.sqrt(-2.0 * np.log(u[..., 2]))
    ta, tb = 2.0 * np.pi * u[..., 1], 2.0 * np.pi * u[..., 3]
    return np.stack([ra * np.cos(ta), ra * np.sin(ta), rb * np.cos(tb), rb * np.sin(tb)], axis=-1)


STREAM_POS, STREAM_FEAT, STREAM_EDGE = 0, 1, 2


def node_raw(B, N, nfeat, step, seed):
    """raw_pos [B, N, 3], raw_feat [B, N, nfeat] as the kernel draws them."""
    a = np.arange(B * N, dtype=np.uint64)
    pos = normals4(a, step, STREAM_POS, seed)[:, :3].reshape(B, N, 3)
    q4 = (nfeat + 3) // 4
    f = normals4((a[:, None] * np.uint64(q4) + np.arange(q4, dtype=np.uint64)[None, :]).reshape(-1), step, STREAM_FEAT, seed)
    return pos, f.reshape(B * N, q4 * 4)[:, :nfeat].reshape(B, N, nfeat)


def edge_raw(B, ch, N, step, seed):
    """raw_edge [B, ch, N, N]: element (b, c, hi, lo) = normal (linear & 3) of counter linear >> 2."""
    lin = np.arange(B * ch * N * N, dtype=np.uint64)
    z = normals4(lin >> np.uint64(2), step, STREAM_EDGE, seed)
    return z[np.arange(lin.shape[0]), (lin & np.uint64(3)).astype(np.int64)].reshape(B, ch, N, N)
