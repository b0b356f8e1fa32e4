/*
 * oracle/synth.c -- TEST INFRASTRUCTURE (not product code).
 *
 * Plain-C twin of tests/synth.py: the deterministic,
 * integer-only synthetic YUV content every BASELINE.json config is quoted on.
 * The numpy listing is the specification; this file exists only because numpy
 * needs ~1.5 s per 1080p frame.  tests/test_synth.py checks both against each
 * other and against the md5 recorded in SURVEY.md Appendix C.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static uint32_t h32(uint32_t x, uint32_t y, uint32_t s)
{
    uint32_t h = (x * 0x9E3779B1u) ^ (y * 0x85EBCA77u) ^ (s * 0xC2B2AE3Du);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}

/* bilinear value noise on lattice period P; X, Y are non-negative */
static int64_t vnoise(int64_t X, int64_t Y, int64_t P, uint32_t seed)
{
    int64_t x0 = X / P, y0 = Y / P, fx = X % P, fy = Y % P;
    int64_t a = h32((uint32_t) x0, (uint32_t) y0, seed) & 255;
    int64_t b = h32((uint32_t) (x0 + 1), (uint32_t) y0, seed) & 255;
    int64_t c = h32((uint32_t) x0, (uint32_t) (y0 + 1), seed) & 255;
    int64_t d = h32((uint32_t) (x0 + 1), (uint32_t) (y0 + 1), seed) & 255;
    return ((a * (P - fx) + b * fx) * (P - fy) + (c * (P - fx) + d * fx) * fy) / (P * P);
}

static int64_t tex(int64_t xs, int64_t ys, int64_t ox2, int64_t oy2, uint32_t seed)
{
    int64_t X = xs + ox2 + (1 << 20), Y = ys + oy2 + (1 << 20);
    int64_t t = (vnoise(X, Y, 256, seed) * 5 + vnoise(X, Y, 32, seed + 1) * 4
               + vnoise(X, Y, 8, seed + 2) * 4 + vnoise(X, Y, 4, seed + 3) * 3) / 16;
    return t + (int64_t) (h32((uint32_t) X, (uint32_t) Y, seed + 4) & 63) - 32;
}

/* one sample of plane(w,h,ox2,oy2,seed) at (i,j); floor semantics of >> on negatives */
static int64_t plane_px(int i, int j, int64_t ox2, int64_t oy2, uint32_t seed)
{
    int64_t s = tex(2 * i, 2 * j, ox2, oy2, seed) + tex(2 * i + 1, 2 * j, ox2, oy2, seed)
              + tex(2 * i, 2 * j + 1, ox2, oy2, seed) + tex(2 * i + 1, 2 * j + 1, ox2, oy2, seed);
    return (s + 2) >> 2;
}

static uint8_t clip8(int64_t v) { return v < 0 ? 0 : v > 255 ? 255 : (uint8_t) v; }

/* writes Y then U then V, tightly packed, into out; returns bytes written */
long synth_frame(int w, int h, int hs, int vs, int t, int seed, int cut, uint8_t *out)
{
    int sc = seed + ((cut > 0 && t >= cut) ? 1000 : 0);
    int64_t ox2 = 3 * (int64_t) t, oy2 = t;
    int ow = w / 6 > 16 ? w / 6 : 16, oh = h / 6 > 16 ? h / 6 : 16;
    int px = (w / 5 + (5 * t) / 2) % (w - ow), py = (h / 4 + t) % (h - oh);
    int lx0 = w / 16, lx1 = w / 16 + w / 8, ly0 = h / 16, ly1 = h / 16 + h / 12;
    int cw = (w + (1 << hs) - 1) >> hs, ch = (h + (1 << vs) - 1) >> vs;
    int i, j;
    uint8_t *Y = out, *Uc = out + (long) w * h, *Vc = Uc + (long) cw * ch;

    for (j = 0; j < h; j++) {
        for (i = 0; i < w; i++) {
            int64_t v = plane_px(i, j, ox2, oy2, (uint32_t) sc);
            if (cut > 0 && t >= cut) {
                v = ((v * 3) >> 2) + 60;
            }
            if (i >= px && i < px + ow && j >= py && j < py + oh) {
                int64_t o = plane_px(i - px, j - py, (5 * t) % 2, 0, (uint32_t) (sc + 7));
                v = ((o + v) >> 1) + 20;
            }
            if (i >= lx0 && i < lx1 && j >= ly0 && j < ly1) {
                v = 200;
            }
            v += (int64_t) (h32((uint32_t) i, (uint32_t) j, (uint32_t) (sc * 977 + t)) & 7) - 3;
            Y[(long) j * w + i] = clip8(v);
        }
    }
    for (j = 0; j < ch; j++) {
        for (i = 0; i < cw; i++) {
            int64_t u = plane_px(i, j, ox2 >> hs, oy2 >> vs, (uint32_t) (sc + 11));
            int64_t v = plane_px(i, j, ox2 >> hs, oy2 >> vs, (uint32_t) (sc + 13));
            Uc[(long) j * cw + i] = clip8(64 + (u >> 1));
            Vc[(long) j * cw + i] = clip8(192 - (v >> 1));
        }
    }
    return (long) w * h + 2L * cw * ch;
}

/* n frames starting at frame index `start`, file order */
long synth_sequence(int w, int h, int hs, int vs, int start, int n, int seed, int cut, uint8_t *out)
{
    long off = 0;
    int t;
    for (t = start; t < start + n; t++) {
        off += synth_frame(w, h, hs, vs, t, seed, cut, out + off);
    }
    return off;
}
