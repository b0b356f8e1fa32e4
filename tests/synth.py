#!/usr/bin/env python3
"""Integer-only deterministic synthetic YUV generator.

This numpy listing is the *specification* of the synthetic content used by every
config in BASELINE.json; `oracle/synth.c` (CPU, fast) and the CUDA generator in
the product library must match it bit for bit (checked in tests/test_synth.py).

usage: synth.py W H NFRAMES {444|422|420|411} SEED CUT out.yuv
"""
import sys
import numpy as np

np.seterr(over='ignore')
U = np.uint32


def h32(x, y, s):                      # all arithmetic mod 2^32
    h = (x.astype(U) * U(0x9E3779B1)) ^ (y.astype(U) * U(0x85EBCA77)) ^ U((s * 0xC2B2AE3D) & 0xFFFFFFFF)
    h ^= h >> U(15); h *= U(0x2C1B3C6D); h ^= h >> U(12); h *= U(0x297A2D39); h ^= h >> U(15)
    return h


def vnoise(X, Y, P, seed):             # bilinear value noise, lattice period P, output 0..255 (int64)
    x0 = X // P; y0 = Y // P; fx = X % P; fy = Y % P
    a = (h32(x0, y0, seed) & U(255)).astype(np.int64); b = (h32(x0 + 1, y0, seed) & U(255)).astype(np.int64)
    c = (h32(x0, y0 + 1, seed) & U(255)).astype(np.int64); d = (h32(x0 + 1, y0 + 1, seed) & U(255)).astype(np.int64)
    return ((a * (P - fx) + b * fx) * (P - fy) + (c * (P - fx) + d * fx) * fy) // (P * P)


def plane(w, h, ox2, oy2, seed):
    # 2x-resolution texture sampled at half-pel offset (ox2,oy2), then 2x2 box-averaged with rounding
    ys, xs = np.mgrid[0:2 * h, 0:2 * w].astype(np.int64)
    X = xs + ox2 + (1 << 20); Y = ys + oy2 + (1 << 20)
    t = (vnoise(X, Y, 256, seed) * 5 + vnoise(X, Y, 32, seed + 1) * 4 + vnoise(X, Y, 8, seed + 2) * 4
         + vnoise(X, Y, 4, seed + 3) * 3) // 16 + (h32(X, Y, seed + 4) & U(63)).astype(np.int64) - 32
    return (t[0::2, 0::2] + t[0::2, 1::2] + t[1::2, 0::2] + t[1::2, 1::2] + 2) >> 2


def frame(w, h, hs, vs, t, seed, cut):
    sc = seed + (1000 if (cut > 0 and t >= cut) else 0)          # scene cut switches texture seed
    ox2 = 3 * t; oy2 = t                                          # global pan: +1.5 px/frame x, +0.5 px/frame y
    Y = plane(w, h, ox2, oy2, sc)
    if cut > 0 and t >= cut:
        Y = (Y * 3) // 4 + 60                                     # brightness jump so the avg-luma SCD fires
    ow, oh = max(16, w // 6), max(16, h // 6)                     # foreground object, +2.5 px/frame x, +1 px/frame y
    px = (w // 5 + (5 * t) // 2) % (w - ow); py = (h // 4 + t) % (h - oh)
    O = plane(ow, oh, (5 * t) % 2, 0, sc + 7); Y[py:py + oh, px:px + ow] = (O + Y[py:py + oh, px:px + ow]) // 2 + 20
    Y[h // 16:h // 16 + h // 12, w // 16:w // 16 + w // 8] = 200  # static flat overlay ("logo")
    ys, xs = np.mgrid[0:h, 0:w].astype(np.int64)
    Y = Y + (h32(xs, ys, sc * 977 + t) & U(7)).astype(np.int64) - 3     # per-frame noise -3..+4
    Y = np.clip(Y, 0, 255).astype(np.uint8)
    cw, ch = (w + (1 << hs) - 1) >> hs, (h + (1 << vs) - 1) >> vs
    Uc = np.clip(64 + plane(cw, ch, ox2 >> hs, oy2 >> vs, sc + 11) // 2, 0, 255).astype(np.uint8)
    Vc = np.clip(192 - plane(cw, ch, ox2 >> hs, oy2 >> vs, sc + 13) // 2, 0, 255).astype(np.uint8)
    return Y, Uc, Vc


SHIFTS = {'444': (0, 0), '422': (1, 0), '420': (1, 1), '411': (2, 0)}


def sequence(w, h, n, fmt, seed, cut=0, start=0):
    """Return the n frames as one bytes object in planar YUV file order."""
    hs, vs = SHIFTS[fmt]
    out = bytearray()
    for t in range(start, start + n):
        for p in frame(w, h, hs, vs, t, seed, cut):
            out += p.tobytes()
    return bytes(out)


if __name__ == '__main__':
    w, h, n, fmt, seed, cut, out = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4],
                                    int(sys.argv[5]), int(sys.argv[6]), sys.argv[7])
    with open(out, 'wb') as f:
        f.write(sequence(w, h, n, fmt, seed, cut))
