/*
 * quant.cuh -- the adaptive quantiser shared by the SBT epilogue (sbt_fwd.cu) and the stand-alone
 * quantise kernel (hzcc_enc.cu).  JT is any job type with cw, ch, coef, dv, pq, dg members.
 * Reference: hzcc.c:50-135 (quant/dequant/quantH/dequantH, tmq4pos), hzcc.c:190-281 (which
 * quantiser applies where), SURVEY.md Appendix B-1 (positions visited twice).
 */
#pragma once
#include "sbt.cuh"

namespace dsv {

template <class JT> DSV_D int stab_flags(const JT &J, const uint8_t *stab, int lvl, int bx, int by)
{
    return stab[((by * J.pq.dby[lvl]) >> 14) * J.pq.nbh + ((bx * J.pq.dbx[lvl]) >> 14)];
}

/* quantise + dequantise one coefficient of transform level lvl at band-local (bx,by) */
template <class JT> DSV_D int requant(const JT &J, const uint8_t *stab, int lvl, int bx, int by, int v)
{
    if (lvl >= 4) {
        int s = dz_quant(v, J.pq.ll_q, J.pq.ll_fd);
        return s ? dz_dequant(s, J.pq.ll_q) : 0;
    }
    int f = stab_flags(J, stab, lvl, bx, by);
    if (lvl == 1) {
        int sh = f ? J.pq.sh_hq : J.pq.sh_plain;
        return p2_dequant(p2_quant(v, sh), sh);
    }
    int sel = (f & 2) ? 2 : (f ? 1 : 0);
    const LevelQ &L = J.pq.lv[3 - lvl];
    int s = dz_quant(v, L.q[sel], L.fd[sel]);
    return s ? dz_dequant(s, L.q[sel]) : 0;
}

/* store one high-band coefficient (band: 1 = LH, 2 = HL, 3 = HH) */
template <class JT> DSV_D void emit_h(const JT &J, bool do_quant, const uint8_t *stab, int lvl, int band, int bx, int by, int v)
{
    int wo = sbt_wo(J.cw, lvl), ho = sbt_wo(J.ch, lvl);
    int ax = bx + ((band & 1) ? wo : 0), ay = by + ((band & 2) ? ho : 0);
    if (do_quant) {
        if (lvl <= 2) {
            /* Position also scanned (first) by hzcc level of transform level lvl+1: the reference
             * quantises it there, writes the dequantised value back, then quantises THAT again at
             * this level.  Keep the first symbol for the entropy coder. */
            const DvGeom &g = J.dg;
            bool col = (ax == g.dvx[lvl]) && (ay < g.dvey[lvl]);
            bool row = (ay == g.dvy[lvl]) && (ax < g.dvex[lvl]);
            if (col || row) {
                int U = lvl + 1;
                int woU = sbt_wo(J.cw, U), hoU = sbt_wo(J.ch, U);
                int lx = ax >= woU ? ax - woU : ax, ly = ay >= hoU ? ay - hoU : ay;
                int f = stab_flags(J, stab, U, lx, ly);
                int sel = (f & 2) ? 2 : (f ? 1 : 0);
                const LevelQ &L = J.pq.lv[3 - U];
                int s1 = dz_quant(v, L.q[sel], L.fd[sel]);
                J.dv[col ? g.col_base[lvl] + ay : g.row_base[lvl] + ax] = s1;
                v = s1 ? dz_dequant(s1, L.q[sel]) : 0;
            }
        }
        v = requant(J, stab, lvl, bx, by, v);
    }
    J.coef[(size_t) ay * J.cw + ax] = v;
}

} // namespace dsv
