/* sbt_host.cu -- host-side geometry / quantiser tables for the SBT kernels. */
#include "sbt.cuh"

namespace dsv {

int sbt_num_levels(int cw, int ch) { return lb2((unsigned) imax(cw, ch)); } /* sbt.c:617-628 */

static size_t llx_head_elems(int cw, int ch)
{
    int nlt = imin(SBT_NLT, sbt_num_levels(cw, ch));
    return ((size_t) sbt_wo(cw, nlt) * sbt_wo(ch, nlt) + 3) & ~(size_t) 3;
}

size_t sbt_llx_elems(int cw, int ch) { return llx_head_elems(cw, ch) + (size_t) sbt_wo(cw, SBT_HI) * sbt_wo(ch, SBT_HI) + 4; }

SbtDims sbt_assign_tiles(SbtJob *jobs, int n)
{
    SbtDims d;
    d.njobs = n;
    for (int i = 0; i < n; i++) {
        jobs[i].tile_base = d.tiles;
        d.tiles += jobs[i].tiles_x * jobs[i].tiles_y;
        jobs[i].mtile_base = d.mtiles;
        d.mtiles += jobs[i].mtiles_x * jobs[i].mtiles_y;
        d.any_intra |= !jobs[i].isP;
        d.any_inter |= jobs[i].isP != 0;
    }
    /* uniform groups? */
    const int gsz = (n % 3 == 0) ? 3 : 1;
    bool uni = true;
    for (int i = gsz; i < n && uni; i++) {
        uni = jobs[i].tiles_x * jobs[i].tiles_y == jobs[i % gsz].tiles_x * jobs[i % gsz].tiles_y &&
              jobs[i].mtiles_x * jobs[i].mtiles_y == jobs[i % gsz].mtiles_x * jobs[i % gsz].mtiles_y;
    }
    if (uni && n > 0) {
        d.gsz = gsz;
        d.c0 = jobs[0].tiles_x * jobs[0].tiles_y;
        d.mc0 = jobs[0].mtiles_x * jobs[0].mtiles_y;
        d.c1 = gsz == 3 ? jobs[1].tiles_x * jobs[1].tiles_y : 0;
        d.mc1 = gsz == 3 ? jobs[1].mtiles_x * jobs[1].mtiles_y : 0;
        d.tg = gsz == 3 ? d.c0 + d.c1 + jobs[2].tiles_x * jobs[2].tiles_y : d.c0;
        d.mtg = gsz == 3 ? d.mc0 + d.mc1 + jobs[2].mtiles_x * jobs[2].mtiles_y : d.mc0;
        d.tg_fd = make_fastdiv(d.tg > 0 ? d.tg : 1);
        d.mtg_fd = make_fastdiv(d.mtg > 0 ? d.mtg : 1);
    }
    return d;
}

size_t sbt_dv_elems(int cw, int ch) { return (size_t) 2 * (cw + ch) + 16; }

/* dynamic shared memory of the lo kernels: LL_nlt plus the next level's LL (ping-pong) */
size_t sbt_lo_smem_bytes(int cw, int ch)
{
    int nlt = imin(SBT_NLT, sbt_num_levels(cw, ch));
    size_t a = (size_t) sbt_wo(cw, nlt) * sbt_wo(ch, nlt);
    size_t b = (size_t) sbt_wo(cw, nlt + 1) * sbt_wo(ch, nlt + 1);
    return (a + b + 8) * sizeof(int32_t);
}

void sbt_fill_geometry(SbtJob *j, int pw, int ph, int cw, int ch, int isP, int plane)
{
    j->pw = pw;
    j->ph = ph;
    j->cw = cw;
    j->ch = ch;
    j->isP = isP;
    j->plane = plane;
    j->lvls = sbt_num_levels(cw, ch);
    j->nlt = imin(SBT_NLT, j->lvls);
    j->tiles_x = ceil_div(cw, SBT_TW);
    j->tiles_y = ceil_div(ch, SBT_TH);
    j->mtiles_x = ceil_div(sbt_wo(cw, SBT_HI), SBT_TW);
    j->mtiles_y = ceil_div(sbt_wo(ch, SBT_HI), SBT_TH);
    j->tiles_x_fd = make_fastdiv(j->tiles_x);
    j->mtiles_x_fd = make_fastdiv(j->mtiles_x);
    j->ll2_off = (int) llx_head_elems(cw, ch);

    /* double-visited positions: level l (2,1) elements that level l+1's hzcc scan also covers */
    DvGeom &g = j->dg;
    int off = 0;
    for (int l = 0; l < 3; l++) {
        g.dvx[l] = g.dvy[l] = -1;
        g.dvex[l] = g.dvey[l] = 0;
        g.col_base[l] = g.row_base[l] = 0;
    }
    for (int l = 2; l >= 1; l--) {
        int wsU = sbt_ws(cw, l + 1), hsU = sbt_ws(ch, l + 1);
        int woU = sbt_wo(cw, l + 1), hoU = sbt_wo(ch, l + 1);
        g.dvex[l] = 2 * woU;
        g.dvey[l] = 2 * hoU;
        g.dvx[l] = (wsU & 1) ? wsU : -1;
        g.dvy[l] = (hsU & 1) ? hsU : -1;
        g.col_base[l] = off;
        off += g.dvey[l];
        g.row_base[l] = off;
        off += g.dvex[l];
    }
    g.total = off;
}

void sbt_fill_quant(SbtJob *j, int q, int isP, int plane, int nbh, int nbv)
{
    PlaneQ &p = j->pq;
    int qe = (plane > 0 && q > 512) ? 512 : q; /* chroma limit, hzcc.c:50-57 */
    j->quant = q;
    p.nbh = nbh;
    p.nbv = nbv;
    p.ll_q = get_quant(qe, isP, 0);
    p.ll_fd = make_fastdiv(2 * p.ll_q);
    for (int l = 0; l < 2; l++) {
        int base = get_quant(qe, isP, l);
        int v[3] = {base, base >> 1, base >> 2}; /* tmq4pos, hzcc.c:63-74 */
        for (int k = 0; k < 3; k++) {
            p.lv[l].q[k] = v[k] < 16 ? 16 : v[k];
            p.lv[l].fd[k] = make_fastdiv(2 * p.lv[l].q[k]);
        }
    }
    {
        int s = lb2((unsigned) get_quant(qe, isP, 2));
        p.sh_plain = s;
        p.sh_hq = iclamp(s - (isP ? 1 : 3), 1, 24); /* DSV_QP_P / DSV_QP_I */
    }
    for (int l = 1; l <= 3; l++) {
        p.dbx[l] = (nbh << 14) / sbt_wo(j->cw, l);
        p.dby[l] = (nbv << 14) / sbt_wo(j->ch, l);
    }
    p.dbx[0] = p.dby[0] = 0;

    /* nudge bounds of the filtered inverse (sbt.c:677-696); uses the frame quant, NOT chroma-limited */
    for (int l = 0; l < 16; l++) {
        int v;
        if (l > 3) {
            v = get_quant(q, isP, 0) / 2;
        } else if (l >= 1) {
            v = get_quant(q, isP, 3 - l);
            if (l == 1) {
                v = iclamp(lb2((unsigned) v) - (isP ? 1 : 3), 1, 24);
                v = (1 << v) >> 1;
            }
            v /= 2;
        } else {
            v = 0;
        }
        j->hqp[l] = v;
    }
}

} // namespace dsv
