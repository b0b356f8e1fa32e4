/*
 * synth.cu -- SURVEY.md Appendix-C synthetic YUV content generated on the GPU (bench / test utility, not
 * part of the codec path): multi-octave integer value noise sampled at 2x resolution and box filtered,
 * half-pel global pan, a moving foreground rectangle, a static flat overlay and per-frame noise.
 * Integer-only and bit-identical to the numpy specification (tests/synth.py) and its C twin
 * (oracle/synth.c); tests/test_gpu_batch.py compares them.
 */
#include "common.cuh"

namespace dsv {

DSV_D uint32_t h32(uint32_t x, uint32_t y, uint32_t s)
{
    uint32_t h = (x * 0x9E3779B1u) ^ (y * 0x85EBCA77u) ^ (s * 0xC2B2AE3Du);
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 12;
    h *= 0x297A2D39u;
    h ^= h >> 15;
    return h;
}

/* bilinear value noise on lattice period P (power of two, P = 1 << LP); X, Y non-negative */
DSV_D int vnoise(int X, int Y, int LP, uint32_t seed)
{
    const int P = 1 << LP;
    const int x0 = X >> LP, y0 = Y >> LP, fx = X & (P - 1), fy = Y & (P - 1);
    const int a = (int) (h32((uint32_t) x0, (uint32_t) y0, seed) & 255);
    const int b = (int) (h32((uint32_t) (x0 + 1), (uint32_t) y0, seed) & 255);
    const int c = (int) (h32((uint32_t) x0, (uint32_t) (y0 + 1), seed) & 255);
    const int d = (int) (h32((uint32_t) (x0 + 1), (uint32_t) (y0 + 1), seed) & 255);
    return ((a * (P - fx) + b * fx) * (P - fy) + (c * (P - fx) + d * fx) * fy) >> (2 * LP);
}

DSV_D int tex(int xs, int ys, int ox2, int oy2, uint32_t seed)
{
    const int X = xs + ox2 + (1 << 20), Y = ys + oy2 + (1 << 20);
    const int t = (vnoise(X, Y, 8, seed) * 5 + vnoise(X, Y, 5, seed + 1) * 4 + vnoise(X, Y, 3, seed + 2) * 4 +
                   vnoise(X, Y, 2, seed + 3) * 3) / 16;
    return t + (int) (h32((uint32_t) X, (uint32_t) Y, seed + 4) & 63) - 32;
}

DSV_D int plane_px(int i, int j, int ox2, int oy2, uint32_t seed)
{
    const int s = tex(2 * i, 2 * j, ox2, oy2, seed) + tex(2 * i + 1, 2 * j, ox2, oy2, seed) +
                  tex(2 * i, 2 * j + 1, ox2, oy2, seed) + tex(2 * i + 1, 2 * j + 1, ox2, oy2, seed);
    return (s + 2) >> 2;
}

struct SynthArgs {
    int w, h, hs, vs, cw, ch, start, seed, cut;
    size_t frame_bytes;
    uint8_t *out;
};

__global__ void __launch_bounds__(256) synth_kernel(SynthArgs a)
{
    const int t = a.start + (int) blockIdx.z;
    uint8_t *out = a.out + (size_t) blockIdx.z * a.frame_bytes;
    const int sc = a.seed + ((a.cut > 0 && t >= a.cut) ? 1000 : 0);
    const int ox2 = 3 * t, oy2 = t;
    const int i = (int) (blockIdx.x * blockDim.x + threadIdx.x), j = (int) blockIdx.y;
    if (j < a.h) {
        if (i >= a.w) {
            return;
        }
        const int w = a.w, h = a.h;
        const int ow = w / 6 > 16 ? w / 6 : 16, oh = h / 6 > 16 ? h / 6 : 16;
        const int px = (w / 5 + (5 * t) / 2) % (w - ow), py = (h / 4 + t) % (h - oh);
        const int lx0 = w / 16, lx1 = w / 16 + w / 8, ly0 = h / 16, ly1 = h / 16 + h / 12;
        int v = plane_px(i, j, ox2, oy2, (uint32_t) sc);
        if (a.cut > 0 && t >= a.cut) {
            v = ((v * 3) >> 2) + 60;
        }
        if (i >= px && i < px + ow && j >= py && j < py + oh) {
            const int o = plane_px(i - px, j - py, (5 * t) % 2, 0, (uint32_t) (sc + 7));
            v = ((o + v) >> 1) + 20;
        }
        if (i >= lx0 && i < lx1 && j >= ly0 && j < ly1) {
            v = 200;
        }
        v += (int) (h32((uint32_t) i, (uint32_t) j, (uint32_t) (sc * 977 + t)) & 7) - 3;
        out[(size_t) j * w + i] = clamp_u8(v);
    } else {
        const int cj = j - a.h;
        if (cj >= a.ch || i >= a.cw) {
            return;
        }
        const int u = plane_px(i, cj, ox2 >> a.hs, oy2 >> a.vs, (uint32_t) (sc + 11));
        const int v = plane_px(i, cj, ox2 >> a.hs, oy2 >> a.vs, (uint32_t) (sc + 13));
        uint8_t *U = out + (size_t) a.w * a.h, *V = U + (size_t) a.cw * a.ch;
        U[(size_t) cj * a.cw + i] = clamp_u8(64 + (u >> 1));
        V[(size_t) cj * a.cw + i] = clamp_u8(192 - (v >> 1));
    }
}

void synth_launch(int w, int h, int hs, int vs, int start, int n, int seed, int cut, uint8_t *d_out, cudaStream_t st)
{
    SynthArgs a;
    a.w = w;
    a.h = h;
    a.hs = hs;
    a.vs = vs;
    a.cw = ceil_shift(w, hs);
    a.ch = ceil_shift(h, vs);
    a.start = start;
    a.seed = seed;
    a.cut = cut;
    a.frame_bytes = (size_t) w * h + 2 * (size_t) a.cw * a.ch;
    a.out = d_out;
    DSV_LAUNCH(synth_kernel, dim3(ceil_div(w, 256), h + a.ch, n), dim3(256), 0, st, a);
    KERNEL_CHECK();
}

} // namespace dsv
