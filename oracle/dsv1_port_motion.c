/*
 * oracle/dsv1_port_motion.c -- TEST INFRASTRUCTURE (checker only; never linked into or called from the
 * product path).
 *
 * Plain-C restatement of the motion half of the DSV1 hot path, written from the algorithm (SURVEY.md
 * section 3.3, Appendix D) in closed form over "sample access functions" instead of the reference's
 * pointer walks: luma pyramid (frame.c:240-327), hierarchical motion estimation with the level-0 half-pel
 * refinement, block statistics and intra decision (hme.c:32-741), and half-pel block motion compensation
 * with residual formation / reconstruction (bmc.c:29-346).  Each function cites the lines it restates.
 * Parity is PINNED: tests/test_oracle_motion.py compares every function here with
 * oracle/_ref/libdsv1ref.so (the unmodified reference) on Appendix-C content incl. the scene-cut pair with
 * 269 intra blocks, forced vectors and partial intra masks.
 *
 * Exported flat API (prefix port_) = the one in oracle/ref_harness.c.
 */
#include <limits.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BORDER 64
#define IMIN(a, b) ((a) < (b) ? (a) : (b))
#define IMAX(a, b) ((a) > (b) ? (a) : (b))
#define ICLAMP(x, a, b) ((x) < (a) ? (a) : ((x) > (b) ? (b) : (x)))
#define CEIL_SHIFT(x, s) (((x) + (1 << (s)) - 1) >> (s))

static int u8c(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* ---- bordered planes in the reference's layout (frame.c:63-120): the three planes of a frame are contiguous,
 * so a tap that steps past one plane's last border byte reads the next plane's (zero or replicated) border ---- */
typedef struct {
    uint8_t *base; /* allocation incl. guard */
    uint8_t *p[3]; /* sample (0,0) */
    int stride[3], w[3], h[3];
} frame_t;

#define GUARD 16384

static void frame_alloc(frame_t *f, int w, int h, int hs, int vs)
{
    size_t len[3], tot = 0;
    int c;
    f->w[0] = w;
    f->h[0] = h;
    f->w[1] = f->w[2] = CEIL_SHIFT(w, hs);
    f->h[1] = f->h[2] = CEIL_SHIFT(h, vs);
    for (c = 0; c < 3; c++) {
        f->stride[c] = (f->w[c] + 2 * BORDER + 15) & ~15;
        len[c] = (size_t) f->stride[c] * (f->h[c] + 2 * BORDER);
        tot += len[c];
    }
    f->base = (uint8_t *) calloc(1, tot + 2 * GUARD);
    tot = GUARD;
    for (c = 0; c < 3; c++) {
        f->p[c] = f->base + tot + (size_t) f->stride[c] * BORDER + BORDER;
        tot += len[c];
    }
}

static void frame_free(frame_t *f) { free(f->base); }

/* replicate the border of plane c (dsv_extend_frame, frame.c:263-327) */
static void plane_extend(frame_t *f, int c)
{
    int x, y, w = f->w[c], h = f->h[c], s = f->stride[c];
    uint8_t *p = f->p[c];
    for (y = -BORDER; y < h + BORDER; y++) {
        int sy = ICLAMP(y, 0, h - 1);
        for (x = -BORDER; x < w + BORDER; x++) {
            if (y < 0 || y >= h || x < 0 || x >= w) {
                p[(ptrdiff_t) y * s + x] = p[(ptrdiff_t) sy * s + ICLAMP(x, 0, w - 1)];
            }
        }
    }
}

static void frame_load(frame_t *f, const uint8_t *yuv, int w, int h, int hs, int vs)
{
    int c, y;
    frame_alloc(f, w, h, hs, vs);
    for (c = 0; c < 3; c++) {
        for (y = 0; y < f->h[c]; y++) {
            memcpy(f->p[c] + (ptrdiff_t) y * f->stride[c], yuv, (size_t) f->w[c]);
            yuv += f->w[c];
        }
        plane_extend(f, c);
    }
}

static void frame_store(const frame_t *f, uint8_t *yuv)
{
    int c, y;
    for (c = 0; c < 3; c++) {
        for (y = 0; y < f->h[c]; y++) {
            memcpy(yuv, f->p[c] + (ptrdiff_t) y * f->stride[c], (size_t) f->w[c]);
            yuv += f->w[c];
        }
    }
}

#define PX(f, c, x, y) ((int) (f)->p[c][(ptrdiff_t) (y) * (f)->stride[c] + (x)])

/* ======================================================================== */
/* luma pyramid (dsv_ds2x_frame_luma + dsv_extend_frame_luma, frame.c:240-327) */
/* ======================================================================== */
static void pyr_level(frame_t *dst, const frame_t *src, int hs, int vs)
{
    int x, y;
    frame_alloc(dst, CEIL_SHIFT(src->w[0], 1), CEIL_SHIFT(src->h[0], 1), hs, vs);
    for (y = 0; y < dst->h[0]; y++) {
        for (x = 0; x < dst->w[0]; x++) {
            dst->p[0][(ptrdiff_t) y * dst->stride[0] + x] =
                (uint8_t) ((PX(src, 0, 2 * x, 2 * y) + PX(src, 0, 2 * x + 1, 2 * y) + PX(src, 0, 2 * x, 2 * y + 1) +
                            PX(src, 0, 2 * x + 1, 2 * y + 1) + 2) >> 2);
        }
    }
    plane_extend(dst, 0);
}

int port_pyramid(const uint8_t *yuv, int w, int h, int subsamp, int levels, uint8_t *out, int *out_w, int *out_h)
{
    frame_t f[6];
    int hs = (subsamp >> 2) & 3, vs = subsamp & 3, l, y;
    frame_load(&f[0], yuv, w, h, hs, vs);
    for (l = 0; l < levels; l++) {
        pyr_level(&f[l + 1], &f[l], hs, vs);
        for (y = 0; y < f[l + 1].h[0]; y++) {
            memcpy(out, f[l + 1].p[0] + (ptrdiff_t) y * f[l + 1].stride[0], (size_t) f[l + 1].w[0]);
            out += f[l + 1].w[0];
        }
        out_w[l] = f[l + 1].w[0];
        out_h[l] = f[l + 1].h[0];
    }
    for (l = 0; l <= levels; l++) {
        frame_free(&f[l]);
    }
    return 0;
}

/* ======================================================================== */
/* hierarchical motion estimation (hme.c:378-741)                            */
/* ======================================================================== */
typedef struct {
    int16_t x, y;
    uint8_t mode, submask, lo_var, lo_tex, high_detail, pad[3];
} mv_t; /* == DSV_MV (dsv.h:137-150) */

static int sad_at(const frame_t *s, int sx, int sy, const frame_t *r, int rx, int ry, int bw, int bh)
{
    int i, j, acc = 0;
    for (j = 0; j < bh; j++) {
        for (i = 0; i < bw; i++) {
            acc += abs(PX(s, 0, sx + i, sy + j) - PX(r, 0, rx + i, ry + j));
        }
    }
    return acc;
}

/* variance / texture of a w x h block (block_analysis, hme.c:211-246); uint32 wrap-around kept */
static unsigned blk_analysis(const frame_t *f, int c, int x0, int y0, int w, int h, unsigned *tex)
{
    unsigned s = 0, ss = 0, sh = 0, sv = 0;
    int x, y;
    for (y = 0; y < h; y++) {
        for (x = 0; x < w; x++) {
            int p = PX(f, c, x0 + x, y0 + y);
            int right = x == w - 1 ? p : PX(f, c, x0 + x + 1, y0 + y);
            int up = y == 0 ? p : PX(f, c, x0 + x, y0 + y - 1);
            sh += (unsigned) abs(p - right);
            sv += (unsigned) abs(p - up);
            s += (unsigned) p;
            ss += (unsigned) (p * p);
        }
    }
    if (tex) {
        *tex = ((sh + sv) / 2) / (unsigned) (w * h);
    }
    return ss - (s * s) / (unsigned) (w * h);
}

/* 14x14 patch statistics (block_texture, hme.c:179-209) on an arbitrary sample getter */
static int patch_stats(const uint8_t *p, int stride, int *avg, int *var)
{
    unsigned sh = 0, sv = 0, av = 0, avs = 0;
    int x, y;
    for (y = 0; y < 14; y++) {
        for (x = 0; x < 14; x++) {
            int v = p[y * stride + x];
            int right = x == 13 ? v : p[y * stride + x + 1];
            int up = y == 0 ? v : p[(y - 1) * stride + x];
            sh += (unsigned) abs(v - right);
            sv += (unsigned) abs(v - up);
            av += (unsigned) v;
            avs += (unsigned) (v * v);
        }
    }
    *avg = (int) (av / 196u);
    *var = (int) (avs - (av * av) / 196u);
    return (int) (((sh + sv) / 2) / 196u);
}

/* "does the zero-vector reference do more good than evil" (intra_metric, hme.c:87-134) */
static int good_vs_evil(const frame_t *s, const frame_t *r, int x0, int y0, int w, int h)
{
    unsigned good = 0, evil = 0;
    int x, y;
    for (y = 0; y < h; y++) {
        for (x = 0; x < w; x++) {
            int a = PX(s, 0, x0 + x, y0 + y), b = PX(r, 0, x0 + x, y0 + y);
            int al = x == 0 ? a : PX(s, 0, x0 + x - 1, y0 + y), bl = x == 0 ? b : PX(r, 0, x0 + x - 1, y0 + y);
            int au = y == 0 ? a : PX(s, 0, x0 + x, y0 + y - 1), bu = y == 0 ? b : PX(r, 0, x0 + x, y0 + y - 1);
            int d = abs(a - b);
            good += (unsigned) (abs(a - al) + abs(a - au) + abs(b - bl) + abs(b - bu));
            if (d == 0) {
                good += 192;
            } else if (d == 1) {
                good += 128;
            } else if (d == 2) {
                good += 96;
            } else {
                evil += (unsigned) d;
            }
        }
    }
    return good >= (unsigned) ((w + h) >> 1) * evil;
}

/* D.3: a sample the reduced-range intra path cannot represent keeps the block inter (hme.c:141-177) */
static int intra_unrepresentable(const frame_t *s, const frame_t *r, int x0, int y0, int w, int h)
{
    int x, y, avg = 0;
    for (y = 0; y < h; y++) {
        for (x = 0; x < w; x++) {
            avg += PX(r, 0, x0 + x, y0 + y);
        }
    }
    avg /= w * h;
    for (y = 0; y < h; y++) {
        for (x = 0; x < w; x++) {
            int p = PX(s, 0, x0 + x, y0 + y);
            if (u8c(avg + u8c(p - avg + 128) - 128) != p) {
                return 1;
            }
        }
    }
    return 0;
}

static unsigned chroma_maxvar(const frame_t *f, int x0, int y0, int w, int h)
{
    unsigned vu = blk_analysis(f, 1, x0, y0, w, h, NULL), vv = blk_analysis(f, 2, x0, y0, w, h, NULL);
    return vu > vv ? vu : vv;
}

/* luma half-pel taps (hme.c:340-349, bmc.c:113-122) */
static int tap_h(const frame_t *f, int c, int x, int y) { return 9 * (PX(f, c, x, y) + PX(f, c, x + 1, y)) - (PX(f, c, x - 1, y) + PX(f, c, x + 2, y)); }
static int tap_v(const frame_t *f, int c, int x, int y) { return 9 * (PX(f, c, x, y) + PX(f, c, x, y + 1)) - (PX(f, c, x, y - 1) + PX(f, c, x, y + 2)); }
/* half-pel sample at integer position (x, y) + (xh, yh)/2, xh,yh in {0,1} (hme.c:350-376, bmc.c:124-174) */
static int hp_luma(const frame_t *f, int x, int y, int xh, int yh)
{
    if (!xh && !yh) {
        return PX(f, 0, x, y);
    }
    if (xh && !yh) {
        return u8c((tap_h(f, 0, x, y) + 8) >> 4);
    }
    if (!xh) {
        return u8c((tap_v(f, 0, x, y) + 8) >> 4);
    }
    return u8c((9 * (tap_h(f, 0, x, y) + tap_h(f, 0, x, y + 1)) - (tap_h(f, 0, x, y - 1) + tap_h(f, 0, x, y + 2)) + 128) >> 8);
}

static int refine_levels(const frame_t *src, const frame_t *ref, int nlevels, int subsamp, int blk_w, int blk_h,
                         int nbh, int nbv, mv_t *out)
{
    static const int xf[9] = {0, 1, -1, 0, 0, -1, 1, -1, 1}, yf[9] = {0, 0, 0, 1, -1, -1, -1, 1, 1};
    static const int xh[8] = {1, -1, 0, 0, -1, 1, -1, 1}, yh[8] = {0, 0, 1, -1, -1, -1, 1, 1};
    static const int ptx[5] = {0, -2, 2, 0, 0}, pty[5] = {0, 0, 0, -2, 2};
    const int hs = (subsamp >> 2) & 3, vs = subsamp & 3;
    mv_t *field[7] = {0};
    int level, nintra = 0, i, j, k, m;
    for (level = 0; level <= nlevels; level++) {
        field[level] = (mv_t *) calloc((size_t) nbh * nbv, sizeof(mv_t));
    }
    for (level = nlevels; level >= 0; level--) {
        const frame_t *S = &src[level], *R = &ref[level];
        const mv_t *parent = level < nlevels ? field[level + 1] : NULL;
        mv_t *mf = field[level];
        const int step = 1 << level, W = R->w[0], H = R->h[0];
        nintra = 0;
        for (j = 0; j < nbv; j += step) {
            for (i = 0; i < nbh; i += step) {
                mv_t *mv = &mf[i + j * nbh];
                int bx = (i * blk_w) >> level, by = (j * blk_h) >> level, bw, bh;
                int cand[8][2], n = 0, best_k, dx, dy, best, fx, fy;
                memset(mv, 0, sizeof(*mv));
                if (bx >= S->w[0] || by >= S->h[0]) {
                    continue; /* hme.c:442-445 */
                }
                bw = IMIN(S->w[0] - bx, blk_w);
                bh = IMIN(S->h[0] - by, blk_h);
                /* candidates: zero, then unique non-zero parents (hme.c:452-480) */
                cand[n][0] = cand[n][1] = 0;
                n++;
                if (parent) {
                    int pi = i & ~((step << 1) - 1), pj = j & ~((step << 1) - 1);
                    for (m = 0; m < 5; m++) {
                        int x = pi + ptx[m] * step, y = pj + pty[m] * step, dup = 0;
                        if (x < 0 || x >= nbh || y < 0 || y >= nbv) {
                            continue;
                        }
                        if (parent[x + y * nbh].x == 0 && parent[x + y * nbh].y == 0) {
                            continue;
                        }
                        for (k = 0; k < n; k++) {
                            dup |= cand[k][0] == parent[x + y * nbh].x && cand[k][1] == parent[x + y * nbh].y;
                        }
                        if (!dup) {
                            cand[n][0] = parent[x + y * nbh].x;
                            cand[n][1] = parent[x + y * nbh].y;
                            n++;
                        }
                    }
                }
                /* best inherited vector: first minimum over the valid ones, default = last (hme.c:482-510) */
                best_k = n - 1;
                if (n > 1) {
                    int best_score = INT_MAX;
                    for (k = 0; k < n; k++) {
                        int cx = bx + (cand[k][0] >> level), cy = by + (cand[k][1] >> level), sc;
                        if (cx < -BORDER || cy < -BORDER || cx + bw > W + BORDER || cy + bh > H + BORDER) {
                            continue;
                        }
                        sc = sad_at(S, bx, by, R, cx, cy, bw, bh);
                        if (best_score > sc) {
                            best_score = sc;
                            best_k = k;
                        }
                    }
                }
                dx = ICLAMP(cand[best_k][0] >> level, -bw - bx, W - bx);
                dy = ICLAMP(cand[best_k][1] >> level, -bh - by, H - by);
                /* 9-point full-pel search, first minimum (hme.c:522-541) */
                best = INT_MAX;
                m = 0;
                for (k = 0; k < 9; k++) {
                    int sc = sad_at(S, bx, by, R, bx + dx + xf[k], by + dy + yf[k], bw, bh);
                    if (best > sc) {
                        best = sc;
                        m = k;
                    }
                }
                fx = dx + xf[m];
                fy = dy + yf[m];
                mv->x = (int16_t) (fx << level);
                mv->y = (int16_t) (fy << level);
                if (level != 0) {
                    continue;
                }
                /* ---- level 0: half-pel refinement on the 14x14 centre patch (hme.c:551-598) ---- */
                {
                    const unsigned area = (unsigned) (bw * bh), areasq = area * area;
                    const int cx = bx + (bw >> 1) - 7, cy = by + (bh >> 1) - 7;
                    uint8_t refblk[14 * 14];
                    int hx = 0, hy = 0, found = 0, x, y;
                    unsigned luma_tex, luma_var, thresh_intra;
                    int src_tex, src_avg, src_var, ref_tex, ref_avg, ref_var, intra = 0;
                    uint8_t srcpatch[14 * 14];
                    if (best > blk_w * blk_h) {
                        int best_hp = (int) ((unsigned) (best * 196) / area);
                        for (k = 0; k < 8; k++) {
                            int sc = 0;
                            for (y = 0; y < 14; y++) {
                                for (x = 0; x < 14; x++) {
                                    /* position (cx + fx + x, cy + fy + y) + (xh, yh)/2: negative halves step one sample back */
                                    int ix = cx + fx + x + (xh[k] < 0 ? -1 : 0), iy = cy + fy + y + (yh[k] < 0 ? -1 : 0);
                                    sc += abs(PX(S, 0, cx + x, cy + y) - hp_luma(R, ix, iy, xh[k] != 0, yh[k] != 0));
                                }
                            }
                            if (best_hp > sc) {
                                best_hp = sc;
                                hx = xh[k];
                                hy = yh[k];
                                found = 1;
                            }
                        }
                        if (found) {
                            best = (int) ((unsigned) best_hp * area / 196u);
                        }
                    }
                    mv->x = (int16_t) (2 * fx + hx);
                    mv->y = (int16_t) (2 * fy + hy);
                    for (y = 0; y < 14; y++) {
                        for (x = 0; x < 14; x++) {
                            srcpatch[y * 14 + x] = (uint8_t) PX(S, 0, cx + x, cy + y);
                            if (found) {
                                int ix = cx + fx + x + (hx < 0 ? -1 : 0), iy = cy + fy + y + (hy < 0 ? -1 : 0);
                                refblk[y * 14 + x] = (uint8_t) hp_luma(R, ix, iy, hx != 0, hy != 0);
                            } else {
                                refblk[y * 14 + x] = (uint8_t) PX(R, 0, cx + (mv->x >> 1) + x, cy + (mv->y >> 1) + y);
                            }
                        }
                    }
                    /* ---- block statistics and the intra cascade (hme.c:599-720) ---- */
                    luma_var = blk_analysis(S, 0, bx, by, bw, bh, &luma_tex);
                    mv->lo_tex = luma_tex <= 2;
                    mv->lo_var = luma_var < areasq;
                    src_tex = patch_stats(srcpatch, 14, &src_avg, &src_var);
                    ref_tex = patch_stats(refblk, 14, &ref_avg, &ref_var);
                    {
                        unsigned thresh_tex = 1;
                        int thresh_var = 196;
                        const mv_t *nb;
                        if (i > 0 && (nb = &mf[j * nbh + i - 1])->mode == 0 && !nb->lo_tex && !nb->lo_var) {
                            thresh_var *= 14;
                            thresh_tex++;
                        }
                        if (j > 0 && (nb = &mf[(j - 1) * nbh + i])->mode == 0 && !nb->lo_tex && !nb->lo_var) {
                            thresh_var *= 14;
                            thresh_tex++;
                        }
                        if (i > 0 && j > 0 && (nb = &mf[(j - 1) * nbh + i - 1])->mode == 0 && !nb->lo_tex && !nb->lo_var) {
                            thresh_var *= 14 / 4;
                            thresh_tex++;
                        }
                        mv->high_detail = luma_tex > thresh_tex && src_var > thresh_var;
                    }
                    thresh_intra = areasq / 16;
                    if (src_tex < 2 && blk_analysis(R, 0, bx, by, bw, bh, NULL) > luma_var * 2) {
                        intra = 1;
                    } else if (ref_var > src_var * 2) {
                        intra = 1;
                    } else if (src_tex == 0 && ref_tex != 0) {
                        intra = 1;
                    } else if (abs(src_avg - ref_avg) > 8) {
                        intra = 1;
                    } else if (luma_tex <= 10 && (unsigned) best > thresh_intra) {
                        intra = 1;
                    } else {
                        int cbx = i * (blk_w >> hs), cby = j * (blk_h >> vs), cbw = bw >> hs, cbh = bh >> vs;
                        intra = chroma_maxvar(R, cbx, cby, cbw, cbh) > 4 * chroma_maxvar(S, cbx, cby, cbw, cbh);
                    }
                    if (intra && !intra_unrepresentable(S, R, bx, by, bw, bh)) {
                        int mask = 15;
                        if (src_tex > 1) {
                            int sbw = bw / 2, sbh = bh / 2, q;
                            for (q = 0; q < 4; q++) {
                                if (good_vs_evil(S, R, bx + (q & 1) * sbw, by + (q >> 1) * sbh, sbw, sbh)) {
                                    mask &= ~(1 << q);
                                }
                            }
                        }
                        mv->submask = (uint8_t) mask;
                        if (mask) {
                            mv->mode = 1;
                            nintra++;
                        }
                    }
                }
            }
        }
    }
    memcpy(out, field[0], sizeof(mv_t) * (size_t) nbh * nbv);
    for (level = 0; level <= nlevels; level++) {
        free(field[level]);
    }
    return nintra;
}

int port_hme(const uint8_t *src_yuv, const uint8_t *ref_yuv, int w, int h, int subsamp, int blk_w, int blk_h, int levels,
             void *mv_out)
{
    frame_t s[6], r[6];
    int hs = (subsamp >> 2) & 3, vs = subsamp & 3, l, nintra;
    int nbh = (w + blk_w - 1) / blk_w, nbv = (h + blk_h - 1) / blk_h;
    frame_load(&s[0], src_yuv, w, h, hs, vs);
    frame_load(&r[0], ref_yuv, w, h, hs, vs);
    for (l = 0; l < levels; l++) {
        pyr_level(&s[l + 1], &s[l], hs, vs);
        pyr_level(&r[l + 1], &r[l], hs, vs);
    }
    nintra = refine_levels(s, r, levels, subsamp, blk_w, blk_h, nbh, nbv, (mv_t *) mv_out);
    for (l = 0; l <= levels; l++) {
        frame_free(&s[l]);
        frame_free(&r[l]);
    }
    return nintra * 100 / (nbh * nbv);
}

/* ======================================================================== */
/* block motion compensation (bmc.c:204-346)                                 */
/* ======================================================================== */

/* prediction of plane c into pred (w x h only); pred starts zeroed like the reference's fresh frame */
static void predict_plane(const mv_t *mvs, int nbh, int nbv, int blk_w, int blk_h, int hs, int vs, int c, const frame_t *ref,
                          frame_t *pred)
{
    const int sh = c ? hs : 0, sv = c ? vs : 0;
    const int bw = blk_w >> sh, bh = blk_h >> sv, W = pred->w[c], H = pred->h[c];
    int i, j, x, y;
    for (j = 0; j < nbv; j++) {
        for (i = 0; i < nbh; i++) {
            const mv_t *mv = &mvs[i + j * nbh];
            const int x0 = i * bw, y0 = j * bh;
            const int cw = x0 + bw >= W ? W - x0 : bw, ch = y0 + bh >= H ? H - y0 : bh;
            uint8_t *dst = pred->p[c];
            const int ds = pred->stride[c];
            if (mv->mode == 0) { /* inter (bmc.c:240-254) */
                const int dx = mv->x >> sh, dy = mv->y >> sv;
                const int px = ICLAMP(x0 + (dx >> 1), -BORDER, W - bw + BORDER - 1);
                const int py = ICLAMP(y0 + (dy >> 1), -BORDER, H - bh + BORDER - 1);
                const int fx = dx & 1, fy = dy & 1;
                for (y = 0; y < ch; y++) {
                    for (x = 0; x < cw; x++) {
                        int v;
                        if (c == 0) {
                            v = hp_luma(ref, px + x, py + y, fx, fy);
                        } else if (fx && fy) { /* chroma: bilinear (bmc.c:58-110) */
                            v = (PX(ref, c, px + x, py + y) + PX(ref, c, px + x + 1, py + y) + PX(ref, c, px + x, py + y + 1) +
                                 PX(ref, c, px + x + 1, py + y + 1) + 2) >> 2;
                        } else if (fx) {
                            v = (PX(ref, c, px + x, py + y) + PX(ref, c, px + x + 1, py + y) + 1) >> 1;
                        } else if (fy) {
                            v = (PX(ref, c, px + x, py + y) + PX(ref, c, px + x, py + y + 1) + 1) >> 1;
                        } else {
                            v = PX(ref, c, px + x, py + y);
                        }
                        dst[(ptrdiff_t) (y0 + y) * ds + x0 + x] = (uint8_t) v;
                    }
                }
            } else { /* intra: mean of the co-located reference (sub)block (bmc.c:255-298) */
                const int whole = mv->submask == 15;
                const int sbw = whole ? cw : cw / 2, sbh = whole ? ch : ch / 2;
                int q;
                for (q = 0; q < (whole ? 1 : 4); q++) {
                    const int qx = x0 + (q & 1) * sbw, qy = y0 + (q >> 1) * sbh;
                    int avg = 0;
                    if (sbw <= 0 || sbh <= 0) {
                        continue;
                    }
                    if (whole || (mv->submask & (1 << q))) {
                        for (y = 0; y < sbh; y++) {
                            for (x = 0; x < sbw; x++) {
                                avg += PX(ref, c, qx + x, qy + y);
                            }
                        }
                        avg /= sbw * sbh;
                    }
                    for (y = 0; y < sbh; y++) {
                        for (x = 0; x < sbw; x++) {
                            dst[(ptrdiff_t) (qy + y) * ds + qx + x] =
                                (uint8_t) ((whole || (mv->submask & (1 << q))) ? avg : PX(ref, c, qx + x, qy + y));
                        }
                    }
                }
            }
        }
    }
}

int port_sub_pred(const void *mvs, int w, int h, int subsamp, int blk_w, int blk_h, const uint8_t *inp_yuv,
                  const uint8_t *ref_yuv, uint8_t *pred_out, uint8_t *resid_out)
{
    frame_t inp, ref, pred;
    int hs = (subsamp >> 2) & 3, vs = subsamp & 3, c, x, y;
    int nbh = (w + blk_w - 1) / blk_w, nbv = (h + blk_h - 1) / blk_h;
    frame_load(&inp, inp_yuv, w, h, hs, vs);
    frame_load(&ref, ref_yuv, w, h, hs, vs);
    frame_alloc(&pred, w, h, hs, vs);
    for (c = 0; c < 3; c++) {
        predict_plane((const mv_t *) mvs, nbh, nbv, blk_w, blk_h, hs, vs, c, &ref, &pred);
        for (y = 0; y < inp.h[c]; y++) { /* subf, bmc.c:43-55 */
            for (x = 0; x < inp.w[c]; x++) {
                uint8_t *p = &inp.p[c][(ptrdiff_t) y * inp.stride[c] + x];
                *p = (uint8_t) u8c(*p - PX(&pred, c, x, y) + 128);
            }
        }
    }
    frame_store(&pred, pred_out);
    frame_store(&inp, resid_out);
    frame_free(&inp);
    frame_free(&ref);
    frame_free(&pred);
    return 0;
}

int port_add_pred(const void *mvs, int w, int h, int subsamp, int blk_w, int blk_h, const uint8_t *resid_yuv,
                  const uint8_t *ref_yuv, uint8_t *out_yuv)
{
    frame_t res, ref, out;
    int hs = (subsamp >> 2) & 3, vs = subsamp & 3, c, x, y;
    int nbh = (w + blk_w - 1) / blk_w, nbv = (h + blk_h - 1) / blk_h;
    frame_load(&res, resid_yuv, w, h, hs, vs);
    frame_load(&ref, ref_yuv, w, h, hs, vs);
    frame_alloc(&out, w, h, hs, vs);
    for (c = 0; c < 3; c++) {
        predict_plane((const mv_t *) mvs, nbh, nbv, blk_w, blk_h, hs, vs, c, &ref, &out);
        for (y = 0; y < out.h[c]; y++) { /* addf, bmc.c:29-41 */
            for (x = 0; x < out.w[c]; x++) {
                uint8_t *p = &out.p[c][(ptrdiff_t) y * out.stride[c] + x];
                *p = (uint8_t) u8c(*p + PX(&res, c, x, y) - 128);
            }
        }
    }
    frame_store(&out, out_yuv);
    frame_free(&res);
    frame_free(&ref);
    frame_free(&out);
    return 0;
}
