/*
 * oracle/dsv1_port.c -- TEST INFRASTRUCTURE (checker only; never linked into or
 * called from the product path).
 *
 * Plain-C restatement of the DSV1 per-frame hot path, written from the
 * algorithm (SURVEY.md section 8 / appendices) in "closed form": every output
 * element is described by a formula over the inputs, which is also the shape
 * the CUDA kernels take.  Each function cites the reference lines it restates.
 * Parity is PINNED: tests/test_oracle_*.py compare every function here with
 * oracle/_ref/libdsv1ref.so (the unmodified reference compiled from
 * /root/reference) and with the committed fixtures under tests/golden/.
 *
 * Exported flat API (prefix port_) = the one in oracle/ref_harness.c.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define CEIL_SHIFT(x, s) (((x) + (1 << (s)) - 1) >> (s))
#define IMIN(a, b) ((a) < (b) ? (a) : (b))
#define IMAX(a, b) ((a) > (b) ? (a) : (b))
#define ICLAMP(x, a, b) ((x) < (a) ? (a) : ((x) > (b) ? (b) : (x)))

typedef int32_t sbc_t;

/* ======================================================================== */
/* Quantiser derivation (hzcc.c:50-92, 437-447)                             */
/* ======================================================================== */

/* smallest L with 2^L >= n  (hzcc.c:437-447) */
int port_lb2(unsigned n)
{
    int l = 0;
    while ((1u << l) < n) {
        l++;
    }
    return l;
}

/* per-level base quantiser (hzcc.c:77-92) */
int port_get_quant(int q, int isP, int level)
{
    if (isP) {
        q = q * 3 / 2;
    }
    if (level == 1) {
        q = q * 2 / 3;
    } else if (level == 2) {
        q = q * 3 / 2;
    }
    return q < 16 ? 16 : q;
}

/* number of transform levels: ceil(log2(max(w,h)))  (sbt.c:617-628) */
static int sbt_levels(int w, int h)
{
    return port_lb2((unsigned) IMAX(w, h));
}

/* ======================================================================== */
/* (1) Subband transform                                                    */
/* ======================================================================== */

/* round-half-away-from-zero division by 2^s, s = 1,2,3 (sbt.c:63-88) */
static int rnd_shift(int v, int s)
{
    int half = 1 << (s - 1);
    return v < 0 ? -((-v + half) >> s) : ((v + half) >> s);
}

/* LL scaling (sbt.c:20-22): C division truncates toward zero */
static int ll_down(int v) { return v * 4 / 5; }
static int ll_up(int v) { return v * 5 / 4; }
static int ll_scaled(int isI, int lvl) { return isI ? 1 : lvl > 1; }

/*
 * Forward B4T on one line of n samples with stride s (sbt.c:91-126,166-201).
 * L[k] = round2(3a + 3b - p - n), H[k] = round2(p - 3a + 3b - n) with
 * (p,a,b,n) = x[2k-1], x[2k], x[2k+1], x[2k+2]; x[-1] := x[1]; x[n] := x[n-1].
 * L lands at index k, H at n/2 + k.
 */
static void b4t_fwd_line(sbc_t *out, const sbc_t *in, int n, int s)
{
    int k, half = n / 2;
    for (k = 0; k < half; k++) {
        int p = in[(k == 0 ? 1 : 2 * k - 1) * s];
        int a = in[(2 * k) * s];
        int b = in[(2 * k + 1) * s];
        int nx = in[(k == half - 1 ? n - 1 : 2 * k + 2) * s];
        out[k * s] = rnd_shift(3 * a + 3 * b - p - nx, 1);
        out[(half + k) * s] = rnd_shift(p - 3 * a + 3 * b - nx, 1);
    }
}

/*
 * Inverse B4T on one line (sbt.c:129-163,204-238).
 * out[2k]   = round8(L[k-1] + 3L[k] + H[k-1] - 3H[k])
 * out[2k+1] = round8(3L[k] + L[k+1] + 3H[k] - H[k+1]),  index clamped to [0, n/2-1].
 */
static void b4t_inv_line(sbc_t *out, const sbc_t *in, int n, int s)
{
    int k, half = n / 2;
    for (k = 0; k < half; k++) {
        int km = k == 0 ? 0 : k - 1, kp = k == half - 1 ? k : k + 1;
        int L0 = in[km * s], L1 = in[k * s], L2 = in[kp * s];
        int H0 = in[(half + km) * s], H1 = in[(half + k) * s], H2 = in[(half + kp) * s];
        out[(2 * k) * s] = rnd_shift(L0 + 3 * L1 + H0 - 3 * H1, 3);
        out[(2 * k + 1) * s] = rnd_shift(3 * L1 + L2 + 3 * H1 - H2, 3);
    }
}

/* rows then columns, whole w x h array (sbt.c:240-251) */
static void b4t_fwd_2d(sbc_t *a, sbc_t *tmp, int w, int h)
{
    int i, j;
    for (j = 0; j < h; j++) {
        b4t_fwd_line(tmp + j * w, a + j * w, w, 1);
    }
    for (i = 0; i < w; i++) {
        b4t_fwd_line(a + i, tmp + i, h, w);
    }
}

/* columns then rows (sbt.c:253-265) */
static void b4t_inv_2d(sbc_t *a, sbc_t *tmp, int w, int h)
{
    int i, j;
    for (i = 0; i < w; i++) {
        b4t_inv_line(tmp + i, a + i, h, w);
    }
    for (j = 0; j < h; j++) {
        b4t_inv_line(a + j * w, tmp + j * w, w, 1);
    }
}

/*
 * One forward Haar level (sbt.c:268-349), pair-indexed.  The level-lvl input is the
 * ws x hs top-left region; pair (ix,iy) covers samples (2ix..2ix+1, 2iy..2iy+1).
 * Missing samples of edge pairs (odd ws / hs) are handled by the "doubling" rules.
 */
static void haar_fwd_level(sbc_t *a, sbc_t *tmp, int w, int h, int lvl, int isI)
{
    int ws = CEIL_SHIFT(w, lvl - 1), hs = CEIL_SHIFT(h, lvl - 1);
    int wo = CEIL_SHIFT(w, lvl), ho = CEIL_SHIFT(h, lvl);
    int scale = ll_scaled(isI, lvl);
    int ix, iy, y;

    for (iy = 0; iy < ho; iy++) {
        int has_row2 = 2 * iy + 1 < hs;
        for (ix = 0; ix < wo; ix++) {
            int has_col2 = 2 * ix + 1 < ws;
            const sbc_t *p = a + (2 * iy) * w + 2 * ix;
            int x0 = p[0], ll;
            if (has_col2 && has_row2) {
                int x1 = p[1], x2 = p[w], x3 = p[w + 1];
                ll = x0 + x1 + x2 + x3;
                tmp[iy * w + wo + ix] = x0 - x1 + x2 - x3;            /* LH */
                tmp[(ho + iy) * w + ix] = x0 + x1 - x2 - x3;          /* HL */
                tmp[(ho + iy) * w + wo + ix] = x0 - x1 - x2 + x3;     /* HH */
            } else if (has_row2) { /* last, unpaired column */
                int x2 = p[w];
                ll = 2 * (x0 + x2);
                tmp[(ho + iy) * w + ix] = 2 * (x0 - x2);              /* HL */
            } else if (has_col2) { /* last, unpaired row */
                int x1 = p[1];
                ll = 2 * (x0 + x1);
                tmp[iy * w + wo + ix] = 2 * (x0 - x1);                /* LH */
            } else {
                ll = 4 * x0;
            }
            tmp[iy * w + ix] = scale ? ll_down(ll) : ll;
        }
    }
    for (y = 0; y < hs; y++) {
        memcpy(a + y * w, tmp + y * w, sizeof(sbc_t) * ws);
    }
}

/*
 * Smoothing nudge of the filtered inverse (sbt.c:480-503 / 505-527):
 * c = this pair's (scaled) LL, lp/ln = previous/next LL along the axis,
 * hb = the high band coefficient that is adjusted, bound = hqp.
 */
static int smooth_nudge(int c, int lp, int ln, int hb, int bound)
{
    int mx = c - ln, mn = lp - c, t;
    if (mn > mx) {
        t = mn; mn = mx; mx = t;
    }
    if (mx > 0) mx = 0;
    if (mn < 0) mn = 0;
    if (mx == mn) {
        return hb;
    }
    t = rnd_shift(lp - ln, 2);
    t = ICLAMP(t, mx, mn);
    t = rnd_shift(t - 2 * hb, 1);
    return hb + ICLAMP(t, -bound, bound);
}

/*
 * One inverse Haar level.  filtered = 0: sbt.c:352-435 (chroma); filtered = 1:
 * sbt.c:438-574 (luma) where every full pair with ix > 0 (iy > 0) gets its LH (HL)
 * nudged from the neighbouring LL values.  NOTE the neighbour reads are plain array
 * reads at offset +-1 / +-w from the pair's LL in the PRE-LEVEL array, so for the last
 * pair of an even-sized level the "next LL" is really LH[0] of that row / HL row 0 of
 * that column.  All divisions by 4 truncate toward zero.
 */
static void haar_inv_level(sbc_t *a, sbc_t *tmp, int w, int h, int lvl, int isI, int filtered, int hqp)
{
    int ws = CEIL_SHIFT(w, lvl - 1), hs = CEIL_SHIFT(h, lvl - 1);
    int wo = CEIL_SHIFT(w, lvl), ho = CEIL_SHIFT(h, lvl);
    int scale = ll_scaled(isI, lvl);
    int ix, iy, y;
#define LLAT(off) (scale ? ll_up(a[off]) : a[off])

    for (iy = 0; iy < ho; iy++) {
        int has_row2 = 2 * iy + 1 < hs;
        for (ix = 0; ix < wo; ix++) {
            int has_col2 = 2 * ix + 1 < ws;
            int o = iy * w + ix;
            int LL = LLAT(o);
            sbc_t *d = tmp + (2 * iy) * w + 2 * ix;
            if (has_col2 && has_row2) {
                int LH = a[iy * w + wo + ix];
                int HL = a[(ho + iy) * w + ix];
                int HH = a[(ho + iy) * w + wo + ix];
                if (filtered) {
                    if (ix > 0) {
                        LH = smooth_nudge(LL, LLAT(o - 1), LLAT(o + 1), LH, hqp);
                    }
                    if (iy > 0) {
                        HL = smooth_nudge(LL, LLAT(o - w), LLAT(o + w), HL, hqp);
                    }
                }
                d[0] = (LL + LH + HL + HH) / 4;
                d[1] = (LL - LH + HL - HH) / 4;
                d[w] = (LL + LH - HL - HH) / 4;
                d[w + 1] = (LL - LH - HL + HH) / 4;
            } else if (has_row2) {
                int HL = a[(ho + iy) * w + ix];
                d[0] = (LL + HL) / 4;
                d[w] = (LL - HL) / 4;
            } else if (has_col2) {
                int LH = a[iy * w + wo + ix];
                d[0] = (LL + LH) / 4;
                d[1] = (LL - LH) / 4;
            } else {
                d[0] = LL / 4;
            }
        }
    }
#undef LLAT
    for (y = 0; y < hs; y++) {
        memcpy(a + y * w, tmp + y * w, sizeof(sbc_t) * ws);
    }
}

/* nudge bound per level for the luma inverse (sbt.c:677-696) */
static int inv_hqp(int q, int isP, int lvl)
{
    int v;
    if (lvl > 3) {
        return port_get_quant(q, isP, 0) / 2;
    }
    v = port_get_quant(q, isP, 3 - lvl);
    if (lvl == 1) {
        v = port_lb2((unsigned) v) - (isP ? 1 : 3);
        v = ICLAMP(v, 1, 24);
        v = (1 << v) >> 1;
    }
    return v / 2;
}

/* dsv_fwd_sbt (sbt.c:576-592, 630-651) */
int port_fwd_sbt(const uint8_t *pix, int stride, int pw, int ph, int cw, int ch, int isP, int32_t *coef)
{
    int x, y, l, lvls = sbt_levels(cw, ch);
    sbc_t *tmp = (sbc_t *) calloc((size_t) (cw + 2) * (ch + 2), sizeof(sbc_t));
    (void) pw;
    memset(coef, 0, sizeof(sbc_t) * cw * ch);
    for (y = 0; y < ph; y++) { /* rows >= ph stay 0; columns run to cw (may read 1 past pw) */
        for (x = 0; x < cw; x++) {
            coef[y * cw + x] = (int) pix[y * stride + x] - 128;
        }
    }
    for (l = 1; l <= lvls; l++) {
        if (!isP && l == 1) {
            b4t_fwd_2d(coef, tmp, cw, ch);
        } else {
            haar_fwd_level(coef, tmp, cw, ch, l, !isP);
        }
    }
    free(tmp);
    return 0;
}

/* dsv_inv_sbt (sbt.c:594-614, 653-714) */
int port_inv_sbt(int32_t *coef, int cw, int ch, int q, int isP, int c, uint8_t *pix, int stride, int pw, int ph)
{
    int x, y, l, lvls = sbt_levels(cw, ch);
    sbc_t *tmp = (sbc_t *) calloc((size_t) (cw + 2) * (ch + 2), sizeof(sbc_t));
    for (l = lvls; l >= 1; l--) {
        if (!isP && l == 1) {
            b4t_inv_2d(coef, tmp, cw, ch);
        } else {
            haar_inv_level(coef, tmp, cw, ch, l, !isP, c == 0, c == 0 ? inv_hqp(q, isP, l) : 0);
        }
    }
    for (y = 0; y < ph; y++) {
        for (x = 0; x < pw; x++) {
            int v = coef[y * cw + x] + 128;
            pix[y * stride + x] = (uint8_t) ICLAMP(v, 0, 255);
        }
    }
    free(tmp);
    return 0;
}

/* ======================================================================== */
/* Bit I/O (bs.c) -- MSB first; the writer ORs into zeroed memory            */
/* ======================================================================== */

typedef struct {
    uint8_t *buf;
    uint64_t pos; /* bit position */
} bitw_t;

static void bw_align(bitw_t *b) { b->pos = (b->pos + 7) & ~(uint64_t) 7; }

static void bw_bits(bitw_t *b, unsigned n, uint32_t v) /* bs.c:76-91 */
{
    while (n--) {
        if ((v >> n) & 1) {
            b->buf[b->pos >> 3] |= (uint8_t) (0x80 >> (b->pos & 7));
        }
        b->pos++;
    }
}

/* interleaved exp-Golomb (bs.c:128-145): for x = v+1 with top bit n: n pairs (0, x_bit) then 1 */
static void bw_ueg(bitw_t *b, uint32_t v)
{
    uint32_t x = v + 1;
    int n = 31, i;
    while (!(x >> n)) {
        n--;
    }
    for (i = n - 1; i >= 0; i--) {
        bw_bits(b, 2, (x >> i) & 1);
    }
    bw_bits(b, 1, 1);
}

static void bw_seg(bitw_t *b, int v) /* bs.c:159-175 */
{
    unsigned m = (unsigned) (v < 0 ? -v : v);
    bw_ueg(b, m);
    if (m) {
        bw_bits(b, 1, v < 0);
    }
}

static void bw_neg(bitw_t *b, int v) /* bs.c:190-206; v != 0 */
{
    unsigned m = (unsigned) (v < 0 ? -v : v);
    bw_ueg(b, m - 1);
    bw_bits(b, 1, v < 0);
}

typedef struct {
    const uint8_t *buf;
    uint64_t pos;
} bitr_t;

static unsigned br_bit(bitr_t *b)
{
    unsigned r = (b->buf[b->pos >> 3] >> (7 - (b->pos & 7))) & 1;
    b->pos++;
    return r;
}
static void br_align(bitr_t *b) { b->pos = (b->pos + 7) & ~(uint64_t) 7; }
static uint32_t br_bits(bitr_t *b, unsigned n)
{
    uint32_t v = 0;
    while (n--) {
        v = (v << 1) | br_bit(b);
    }
    return v;
}
static uint32_t br_ueg(bitr_t *b) /* bs.c:147-157 */
{
    uint32_t v = 1;
    while (!br_bit(b)) {
        v = (v << 1) | br_bit(b);
    }
    return v - 1;
}
static int br_seg(bitr_t *b) /* bs.c:177-188 */
{
    int v = (int) br_ueg(b);
    return (v && br_bit(b)) ? -v : v;
}
static int br_neg(bitr_t *b) /* bs.c:208-219 */
{
    int v = (int) br_ueg(b) + 1;
    return (v && br_bit(b)) ? -v : v;
}

/* ======================================================================== */
/* (4) Quantiser + HZCC coefficient coder (hzcc.c)                          */
/* ======================================================================== */

/* dead-zone quantiser and its reconstruction (hzcc.c:94-128) */
static int dz_quant(int v, int q)
{
    int m = (v < 0 ? -v : v) * 2;
    if (m <= q) {
        return 0;
    }
    m = (m + 1) / (2 * q);
    return v < 0 ? -m : m;
}
static int dz_dequant(int v, int q)
{
    int m = ((v < 0 ? -v : v) * (2 * q) + q) >> 1;
    return v < 0 ? -m : m;
}
/* top level: power-of-two quantiser on the magnitude (hzcc.c:114-135) */
static int p2_quant(int v, int s) { return v < 0 ? -((-v) >> s) : (v >> s); }
static int p2_dequant(int v, int s) { return v * (1 << s); }

/*
 * The ten scan regions (hzcc.c:29-48,158-281).
 * kind 0: "LL" (plain quantiser), 1: adaptive dead-zone, 2: adaptive power-of-two.
 */
typedef struct {
    int x0, y0, sw, sh, kind, level;
} region_t;

static int build_regions(region_t *r, int w, int h)
{
    int n = 0, l, s;
    r[n].x0 = 0; r[n].y0 = 0; r[n].sw = CEIL_SHIFT(w, 3); r[n].sh = CEIL_SHIFT(h, 3);
    r[n].kind = 0; r[n].level = 0;
    n++;
    for (l = 0; l < 3; l++) {
        int sw = CEIL_SHIFT(w, 3 - l), sh = CEIL_SHIFT(h, 3 - l);
        for (s = 1; s < 4; s++) {
            r[n].x0 = (s & 1) ? sw : 0;
            r[n].y0 = (s & 2) ? sh : 0;
            r[n].sw = sw; r[n].sh = sh;
            r[n].kind = l == 2 ? 2 : 1;
            r[n].level = l;
            n++;
        }
    }
    return n;
}

typedef struct {
    int q, isP, c, nbh, nbv;
    const uint8_t *stable;
} qctx_t;

/* quantiser that applies to element (x,y) of a region (hzcc.c:196-222,255-258) */
static int region_quant(const qctx_t *qc, const region_t *r, int x, int y)
{
    int q = qc->q, base, flags;
    if (qc->c > 0 && q > 512) {
        q = 512; /* chroma limit, hzcc.c:50-57 */
    }
    base = port_get_quant(q, qc->isP, r->level);
    if (r->kind == 0) {
        return base;
    }
    flags = qc->stable[((y * ((qc->nbv << 14) / r->sh)) >> 14) * qc->nbh + ((x * ((qc->nbh << 14) / r->sw)) >> 14)];
    if (r->kind == 2) {
        int s = port_lb2((unsigned) base);
        if (flags) {
            s = ICLAMP(s - (qc->isP ? 1 : 3), 1, 24);
        }
        return s;
    }
    if (flags & 2) {
        base >>= 2;
    } else if (flags) {
        base >>= 1;
    }
    return base < 16 ? 16 : base;
}

/* dsv_encode_plane + hzcc_enc (hzcc.c:137-293, 449-476) */
int port_encode_plane(int32_t *coef, int cw, int ch, int q, int isP, int c,
                      const uint8_t *stable, int nbh, int nbv, uint8_t *out, int out_cap)
{
    region_t reg[10];
    qctx_t qc;
    bitw_t bw;
    int nreg = build_regions(reg, cw, ch), ri, x, y;
    int dc = coef[0], run = 0, nruns = 0, pending = 0;
    uint64_t plen_at, nruns_at, end;
    (void) out_cap;

    qc.q = q; qc.isP = isP; qc.c = c; qc.nbh = nbh; qc.nbv = nbv; qc.stable = stable;
    bw.buf = out; bw.pos = 0;

    plen_at = bw.pos >> 3;
    bw.pos += 32;
    bw_seg(&bw, dc);
    bw_align(&bw);
    nruns_at = bw.pos >> 3;
    bw.pos += 32;

    coef[0] = 0;
    for (ri = 0; ri < nreg; ri++) {
        const region_t *r = &reg[ri];
        for (y = 0; y < r->sh; y++) {
            for (x = 0; x < r->sw; x++) {
                int32_t *p = &coef[(r->y0 + y) * cw + r->x0 + x];
                int qq = region_quant(&qc, r, x, y);
                int v = r->kind == 2 ? p2_quant(*p, qq) : dz_quant(*p, qq);
                if (v) {
                    *p = r->kind == 2 ? p2_dequant(v, qq) : dz_dequant(v, qq);
                    bw_ueg(&bw, (uint32_t) run);
                    if (pending) {
                        bw_neg(&bw, pending);
                    }
                    pending = v;
                    nruns++;
                    run = 0;
                } else {
                    *p = 0;
                    run++;
                }
            }
        }
    }
    if (pending) {
        bw_neg(&bw, pending);
    }
    bw_align(&bw);
    end = bw.pos;
    bw.pos = nruns_at * 8;
    bw_bits(&bw, 32, (uint32_t) nruns);
    bw.pos = end;
    coef[0] = dc;

    bw_bits(&bw, 8, 0x55);
    bw_align(&bw);
    end = bw.pos;
    bw.pos = plen_at * 8;
    bw_bits(&bw, 32, (uint32_t) ((end >> 3) - plen_at - 4));
    return (int) (end >> 3);
}

/* dsv_decode_plane + hzcc_dec (hzcc.c:295-435, 478-496); `in` starts after the plen field */
int port_decode_plane(const uint8_t *in, int plen, int cw, int ch, int q, int isP, int c,
                      const uint8_t *stable, int nbh, int nbv, int32_t *coef)
{
    region_t reg[10];
    qctx_t qc;
    bitr_t br;
    int nreg = build_regions(reg, cw, ch), ri, x, y;
    int dc, runs, run;

    qc.q = q; qc.isP = isP; qc.c = c; qc.nbh = nbh; qc.nbv = nbv; qc.stable = stable;
    memset(coef, 0, sizeof(int32_t) * cw * ch);
    br.buf = in; br.pos = 0;
    dc = br_seg(&br);
    br_align(&br);
    runs = (int) br_bits(&br, 32);
    br_align(&br);
    run = runs-- > 0 ? (int) br_ueg(&br) : INT_MAX;

    for (ri = 0; ri < nreg; ri++) {
        const region_t *r = &reg[ri];
        for (y = 0; y < r->sh; y++) {
            for (x = 0; x < r->sw; x++) {
                if (run-- == 0) {
                    int v, qq;
                    run = runs-- > 0 ? (int) br_ueg(&br) : INT_MAX;
                    v = br_neg(&br);
                    if ((br.pos >> 3) >= (uint64_t) plen) {
                        goto done; /* truncated plane: keep what was decoded (hzcc.c:337-339) */
                    }
                    qq = region_quant(&qc, r, x, y);
                    coef[(r->y0 + y) * cw + r->x0 + x] = r->kind == 2 ? p2_dequant(v, qq) : dz_dequant(v, qq);
                }
            }
        }
    }
done:
    coef[0] = dc;
    return 0;
}
