"""Pure-Python model of the run-based labelling scheme of csrc/ccl.cu (test infrastructure, CPU only).

It mirrors the kernels phase by phase -- 32 x 128 tiles (64 x 128 for the overlay's background), runs clipped at the
word width, union-find over run starts
with min-index roots, the packed per-tile accumulators, border merges, kept-root numbering by popcount prefix -- so
that the DESIGN of the GPU algorithm (bit tricks, field widths, offsets) is checked against the oracle on the CPU,
where there is no GPU to run the kernels themselves.  Sequential: no atomics, no barriers."""
from __future__ import annotations

import numpy as np

TILE_H = 128
BITS = 64                      # word width of the model (set by label_stats: 32 = droplet path, 64 = overlay path)
M64 = (1 << 64) - 1            # mask of one word (kept under its old name)


def bits_below(b):
    return M64 if b >= BITS else (1 << b) - 1


def bits_upto(b):
    return M64 if b >= BITS - 1 else (2 << b) - 1


def ffs(w):            # 0-based index of the lowest set bit
    return (w & -w).bit_length() - 1


def run_start(w, b):
    z = ~w & M64 & bits_below(b)
    return z.bit_length() if z else 0          # BITS - clz(z)


def run_end(w, b):
    z = ~w & M64 & ~bits_upto(b) & M64
    return ffs(z) - 1 if z else BITS - 1


def merge_rows(dn, up, unite):
    ov = dn & up
    while ov:
        b = ffs(ov)
        unite(run_start(dn, b), run_start(up, b))
        ov &= ~bits_upto(min(run_end(dn, b), run_end(up, b))) & M64


def nz4(v):
    h = (v | (((v | 0x80808080) - 0x01010101) & 0xFFFFFFFF)) & 0x80808080
    return ((((h >> 7) * 0x01020408) & 0xFFFFFFFF) >> 24) & 0xF


def find(P, x):
    while P[x] != x:
        x = P[x]
    return x


def union_min(P, a, b):
    a, b = find(P, a), find(P, b)
    if a < b:
        P[b] = a
    elif b < a:
        P[a] = b


def label_stats(mask: np.ndarray, min_area: int = 1, invert: bool = False, word_bits: int = 32):
    """Returns (labels int32 [H,W], area, sum_row, sum_col) following csrc/ccl.cu with `word_bits`-pixel words."""
    global BITS, M64
    BITS, M64 = word_bits, (1 << word_bits) - 1
    TW = word_bits
    H, W = mask.shape
    WW = (W + TW - 1) // TW
    fgm = (mask != 0) != invert
    bits = [[0] * WW for _ in range(H)]
    for y in range(H):
        for wx in range(WW):
            seg = fgm[y, wx * TW:min(W, wx * TW + TW)]
            w = 0
            for k, v in enumerate(seg):
                if v:
                    w |= 1 << k
            bits[y][wx] = w
    P, ACC, AUX = {}, {}, {}
    rootbits = [[0] * WW for _ in range(H)]
    # K1: tiles
    for ty0 in range(0, H, TILE_H):
        for wx in range(WW):
            x0 = wx * TW
            par = {}
            rows = [bits[ty0 + t][wx] if ty0 + t < H else 0 for t in range(TILE_H)]
            for t, w in enumerate(rows):
                s = w & ~(w << 1) & M64
                while s:
                    b = ffs(s)
                    par[t * TW + b] = t * TW + b
                    s &= s - 1
            for t in range(1, TILE_H):
                merge_rows(rows[t], rows[t - 1], lambda sd, su, t=t: union_min(par, t * TW + sd, (t - 1) * TW + su))
            for t, w in enumerate(rows):
                s = w & ~(w << 1) & M64
                gbase = (ty0 + t) * W + x0
                while s:
                    b = ffs(s)
                    e = run_end(w, b)
                    ln = e - b + 1
                    pk = ln | (((b + e) * ln // 2) << 16) | ((t * ln) << 36)
                    r = find(par, t * TW + b)
                    groot = (ty0 + r // TW) * W + x0 + (r % TW)
                    if r == t * TW + b:
                        ACC[gbase + b] = ACC.get(gbase + b, 0) + pk
                        P[gbase + b] = gbase + b
                        AUX[gbase + b] = 0
                        rootbits[ty0 + t][wx] |= 1 << b
                    else:
                        P[gbase + b] = groot
                        ACC[groot] = ACC.get(groot, 0) + pk
                    s &= s - 1
    for g, pk in ACC.items():      # field widths: no carries between the packed fields
        assert (pk & 0xFFFF) <= 8192 and ((pk >> 16) & 0xFFFFF) <= 258048 and (pk >> 36) <= 520192
    # K2: borders
    for y in range(TILE_H, H, TILE_H):
        for wx in range(WW):
            g = y * W + wx * TW
            merge_rows(bits[y][wx], bits[y - 1][wx], lambda sd, su, g=g: union_min(P, g + sd, g - W + su))
    for y in range(H):
        for wx in range(1, WW):
            a, b = bits[y][wx - 1], bits[y][wx]
            if (a >> (TW - 1)) & b & 1:
                g = y * W + wx * TW
                union_min(P, g, g - TW + run_start(a, TW - 1))
    # K2b: areas
    roots = [(y, wx) for y in range(H) for wx in range(WW) if rootbits[y][wx]]
    if min_area > 1:
        for y, wx in roots:
            s = rootbits[y][wx]
            while s:
                gi = y * W + wx * TW + ffs(s)
                AUX[find(P, gi)] += ACC[gi] & 0xFFFF
                s &= s - 1
    # K3-K5: kept roots, ids in raster order
    n = 0
    for y, wx in roots:
        s = rootbits[y][wx]
        while s:
            gi = y * W + wx * TW + ffs(s)
            if P[gi] == gi:
                if min_area <= 1 or AUX[gi] >= min_area:
                    n += 1
                    AUX[gi] = n
                else:
                    AUX[gi] = 0
            s &= s - 1
    area = np.zeros(n, np.int64)
    s0 = np.zeros(n, np.int64)
    s1 = np.zeros(n, np.int64)
    # K6
    for y, wx in roots:
        s = rootbits[y][wx]
        while s:
            gi = y * W + wx * TW + ffs(s)
            i = AUX[find(P, gi)]
            if i:
                pk = ACC[gi]
                a, sc, sr = pk & 0xFFFF, (pk >> 16) & 0xFFFFF, pk >> 36
                area[i - 1] += a
                s0[i - 1] += sr + a * ((y // TILE_H) * TILE_H)
                s1[i - 1] += sc + a * (wx * TW)
            s &= s - 1
    # K8
    labels = np.zeros((H, W), np.int32)
    for y in range(H):
        for wx in range(WW):
            w = bits[y][wx]
            s = w & ~(w << 1) & M64
            while s:
                b = ffs(s)
                e = run_end(w, b)
                labels[y, wx * TW + b:wx * TW + e + 1] = AUX[find(P, y * W + wx * TW + b)]
                s &= s - 1
    return labels, area, s0, s1
