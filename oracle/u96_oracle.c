/*
 * u96_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see u96_oracle.h).
 *
 * Plain scalar C, written for clarity and fidelity to the cited reference
 * lines, not for speed.  Never linked into the product library.
 */
#include "u96_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* diven.v:26-177 -- pipelined non-restoring divider, emulated stage by stage */
/* ------------------------------------------------------------------------ */
uint64_t orc_diven(int DW, int VW, int QW, int MSB_INV, uint64_t dividend, uint64_t divisor)
{
    /* diven.v:35-43 secondary parameters */
    const int EXT_REM = VW - MSB_INV;
    const int EXT_DIV = DW - MSB_INV - 1;
    const int RW = (EXT_DIV < 0) ? DW - EXT_DIV : DW + EXT_REM;
    const int EVW = (EXT_DIV < 0) ? VW : VW + EXT_DIV;
    const uint64_t RMASK = (RW >= 64) ? ~0ull : ((1ull << RW) - 1);
    const uint64_t EMASK = (EVW >= 64) ? ~0ull : ((1ull << EVW) - 1);
    const uint64_t E1MASK = (1ull << (EVW + 1)) - 1;

    dividend &= (1ull << DW) - 1;
    divisor &= (1ull << VW) - 1;

    /* diven.v:113-123 operand extension */
    uint64_t ext_div = (EXT_DIV <= 0) ? divisor : (divisor << EXT_DIV);
    ext_div &= EMASK;
    uint64_t ext_dvd;
    if (EXT_DIV < 0) {
        ext_dvd = dividend << (-EXT_DIV);
    } else {
        ext_dvd = dividend;
        if ((dividend >> (DW - 1)) & 1) ext_dvd |= RMASK & ~((1ull << DW) - 1); /* sign extend */
    }
    ext_dvd &= RMASK;

    uint64_t rem = ext_dvd, quot = 0;
    const uint64_t div_msb = (ext_div >> (EVW - 1)) & 1;
    /* stage 0 (diven.v:139-151) computes a remainder but no quotient bit;
     * stages 1..QW (diven.v:154-175) append ~op each.                        */
    for (int i = 0; i <= QW; i++) {
        uint64_t op = div_msb ^ ((rem >> (RW - 1)) & 1);             /* 1: add, 0: sub */
        /* update(): {rem[RW-2:0], ~op} + ({EVW+1{~op}} ^ {div, 1'b0})  (diven.v:81-88) */
        uint64_t a = ((rem << 1) | (op ^ 1)) & RMASK;
        uint64_t b = ((op ? 0ull : E1MASK) ^ (ext_div << 1)) & E1MASK;
        rem = (a + b) & RMASK;
        if (i >= 1) quot = (quot << 1) | (op ^ 1);
    }
    /* diven.v:177: add 1 in case of a negative divisor */
    quot = (quot + div_msb) & ((1ull << QW) - 1);
    return quot;
}

/* arithmetic helpers */
static inline int64_t floordiv64(int64_t a, int64_t b)
{
    int64_t q = a / b, r = a % b;
    if (r != 0 && ((r < 0) != (b < 0))) q--;
    return q;
}

/* ------------------------------------------------------------------------ */
/* rectification map: fpga.c:303-366 (== rect_rmp.v:366-585)                  */
/* ------------------------------------------------------------------------ */
void orc_rect_remap32(const orc_rect_params *p, int lr, int W, int H, int32_t *xs, int32_t *ys);

void orc_rect_remap(const orc_rect_params *p, int lr, int W, int H, int16_t *xs, int16_t *ys)
{
    /* the reference stores the map as shorts (struct MAT2S, fpga.c:361-362) */
    int32_t *x32 = (int32_t *)malloc(sizeof(int32_t) * (size_t)W * H), *y32 = (int32_t *)malloc(sizeof(int32_t) * (size_t)W * H);
    orc_rect_remap32(p, lr, W, H, x32, y32);
    for (int i = 0; i < W * H; i++) { xs[i] = (int16_t)x32[i]; ys[i] = (int16_t)y32[i]; }
    free(x32); free(y32);
}

/* RTL-extended variant for frames beyond the RTL counter widths (W > 1023 or H > 511): no 16-bit wrap */
void orc_rect_remap32(const orc_rect_params *p, int lr, int W, int H, int32_t *xs, int32_t *ys)
{
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            /* (u10.0)*(u-8.32) -> (u1.24); constants shared by both cameras (fpga.c:295-300) */
            int64_t xd = (((int64_t)x * p->f2inv[0]) >> 8) - p->c2_f2[0];
            int64_t yd = (((int64_t)y * p->f2inv[1]) >> 8) - p->c2_f2[1];
            /* rotate: each product truncated separately (fpga.c:325-330) */
            int64_t lx = (((int64_t)p->rot[lr][0][0] * xd) >> 24) + (((int64_t)p->rot[lr][1][0] * yd) >> 24) + p->rot[lr][2][0];
            int64_t ly = (((int64_t)p->rot[lr][0][1] * xd) >> 24) + (((int64_t)p->rot[lr][1][1] * yd) >> 24) + p->rot[lr][2][1];
            int64_t lw = (((int64_t)p->rot[lr][0][2] * xd) >> 24) + (((int64_t)p->rot[lr][1][2] * yd) >> 24) + p->rot[lr][2][2];
            /* fpga.c:343: (1ull<<48)/lw evaluates in UNSIGNED 64-bit; lw>0 for any sane rig. */
            int64_t winv = (int64_t)((1ull << 48) / (uint64_t)lw);
            int64_t x2 = (lx * winv) >> 24;
            int64_t y2 = (ly * winv) >> 24;
            int64_t xf = ((x2 * p->f[lr][0]) >> 34) + ((int64_t)p->c[0] << 6);
            int64_t yf = ((y2 * p->f[lr][1]) >> 34) + ((int64_t)p->c[1] << 6);
            xs[y * W + x] = (int32_t)((xf + 1) >> 1);
            ys[y * W + x] = (int32_t)((yf + 1) >> 1);
        }
    }
}

/* rect_intp.v:288-412 */
void orc_rect_interp32(const uint8_t *src, int W, int H, int src_stride,
                       const int32_t *xs, const int32_t *ys, uint8_t *dst);

void orc_rect_interp(const uint8_t *src, int W, int H, int src_stride,
                     const int16_t *xs, const int16_t *ys, uint8_t *dst)
{
    int32_t *x32 = (int32_t *)malloc(sizeof(int32_t) * (size_t)W * H), *y32 = (int32_t *)malloc(sizeof(int32_t) * (size_t)W * H);
    for (int i = 0; i < W * H; i++) { x32[i] = xs[i]; y32[i] = ys[i]; }
    orc_rect_interp32(src, W, H, src_stride, x32, y32, dst);
    free(x32); free(y32);
}

void orc_rect_interp32(const uint8_t *src, int W, int H, int src_stride,
                       const int32_t *xs, const int32_t *ys, uint8_t *dst)
{
    for (int i = 0; i < W * H; i++) {
        int xi = xs[i] >> 5, xf = xs[i] & 31;
        int yi = ys[i] >> 5, yf = ys[i] & 31;
        int tap[2][2];
        for (int dy = 0; dy < 2; dy++)
            for (int dx = 0; dx < 2; dx++) {
                int sx = xi + dx, sy = yi + dy;
                tap[dy][dx] = (sx >= 0 && sx < W && sy >= 0 && sy < H) ? src[sy * src_stride + sx] : 0;
            }
        /* u1.5*u1.5 -> u1.10 weights, u8*u1.10 -> u8.10 (rect_intp.v:337-360) */
        int s = tap[0][0] * (32 - xf) * (32 - yf) + tap[0][1] * xf * (32 - yf)
              + tap[1][0] * (32 - xf) * yf + tap[1][1] * xf * yf;
        int r = ((s >> 9) + 1) >> 1;                 /* rect_intp.v:389-405 */
        dst[i] = (uint8_t)(r > 255 ? 255 : r);
    }
}

/* ------------------------------------------------------------------------ */
/* x-Sobel                                                                    */
/* ------------------------------------------------------------------------ */
void orc_xsobel_rtl(const uint8_t *src, int W, int H, uint8_t *dst)
{
    memset(dst, 0, (size_t)W * H);                  /* rows 0, H-1 never written */
    for (int y = 1; y < H - 1; y++) {
        uint8_t *o = dst + (size_t)y * W;
        o[0] = 32; o[W - 1] = 32;                   /* xsbl2.v:869-872 */
        for (int x = 1; x < W - 1; x++) {
            const uint8_t *a = src + (size_t)(y - 1) * W + x;
            const uint8_t *b = a + W, *c = b + W;
            int s = (a[1] - a[-1]) + 2 * (b[1] - b[-1]) + (c[1] - c[-1]);   /* xsbl2.v:682-698, 826-857 */
            if (s < -32) s = -32;                   /* limit(): xsbl2.v:185-198 */
            if (s > 31) s = 31;
            o[x] = (uint8_t)(s + 32);
        }
    }
}

void orc_xsobel_cv(const uint8_t *src, int W, int H, int cap, uint8_t *dst)
{
    for (int y = 0; y < H; y++) {
        uint8_t *o = dst + (size_t)y * W;
        if ((H & 1) && y == H - 1) {                /* odd H: last row = cap (Appendix A.1) */
            memset(o, cap, W);
            continue;
        }
        int ym = (y > 0) ? y - 1 : 1;               /* reflect-101 */
        int yp = (y < H - 1) ? y + 1 : H - 2;
        const uint8_t *a = src + (size_t)ym * W, *b = src + (size_t)y * W, *c = src + (size_t)yp * W;
        o[0] = (uint8_t)cap; o[W - 1] = (uint8_t)cap;
        for (int x = 1; x < W - 1; x++) {
            int s = (a[x + 1] - a[x - 1]) + 2 * (b[x + 1] - b[x - 1]) + (c[x + 1] - c[x - 1]);
            if (s < -cap) s = -cap;
            if (s > cap) s = cap;
            o[x] = (uint8_t)(s + cap);
        }
    }
}

/* ------------------------------------------------------------------------ */
/* BM, RTL profile                                                            */
/* ------------------------------------------------------------------------ */
typedef struct { uint16_t min1, min2; uint8_t idx1, idx2; uint16_t l, r; } det_t;

/* bm_calc_det.v:124-426 -- 5-level tournament over lanes 1..32 of sad[0..33].
 * Each node carries (L,C,R) = (sad[k-1], sad[k], sad[k+1]) of its winner.    */
static det_t rtl_det(const uint16_t sad[34])
{
    uint16_t L[32], C[32], R[32];
    uint8_t idx[32];
    for (int k = 0; k < 32; k++) { L[k] = sad[k]; C[k] = sad[k + 1]; R[k] = sad[k + 2]; idx[k] = (uint8_t)k; }
    int n = 32;
    uint16_t l4_min[2] = {0, 0};  uint8_t l4_idx[2] = {0, 0};   /* min2_r4 / idx2_r4 */
    uint16_t fin_min = 0;  uint8_t fin_idx = 0;                 /* min2_r5[0] / idx2_r5[0] */
    for (int level = 1; level <= 5; level++) {
        for (int i = 0; i < n / 2; i++) {
            int lo = 2 * i, hi = 2 * i + 1;
            int pick_hi = C[hi] < C[lo];            /* strict <: ties keep the lower index */
            int w = pick_hi ? hi : lo, s = pick_hi ? lo : hi;
            if (level == 4) { l4_min[i] = C[s]; l4_idx[i] = idx[s]; }   /* bm_calc_det.v:286-303 */
            if (level == 5) { fin_min = C[s]; fin_idx = idx[s]; }       /* bm_calc_det.v:335-350 */
            L[i] = L[w]; C[i] = C[w]; R[i] = R[w]; idx[i] = idx[w];
        }
        n /= 2;
    }
    /* bm_calc_det.v:362-377: the smaller of the two level-4 losers, ties -> half 0 */
    uint16_t c1_min; uint8_t c1_idx;
    if (l4_min[1] < l4_min[0]) { c1_min = l4_min[1]; c1_idx = l4_idx[1]; }
    else                       { c1_min = l4_min[0]; c1_idx = l4_idx[0]; }
    /* bm_calc_det.v:382-389 adjacency with 6-bit increments (no wrap) */
    int i1 = idx[0];
    int adj0 = (fin_idx == i1 + 1) || (i1 == fin_idx + 1);
    int adj1 = (c1_idx == i1 + 1) || (i1 == c1_idx + 1);
    det_t d;
    d.min1 = C[0]; d.idx1 = idx[0]; d.l = L[0]; d.r = R[0];
    if (((c1_min < fin_min) && !adj1) || adj0) { d.min2 = c1_min; d.idx2 = c1_idx; }   /* :404-411 */
    else                                       { d.min2 = fin_min; d.idx2 = fin_idx; }
    return d;
}

/* bm_calc_frac.v:63-173; returns the 8-bit two's-complement fraction */
static uint8_t rtl_frac(uint16_t l, uint16_t c, uint16_t r, int bitserial)
{
    int32_t dif_lr = (int32_t)l - r, dif_lc = (int32_t)l - c, dif_rc = (int32_t)r - c;
    int cmp = l < r;
    int neg_val = (dif_lc < 0) || (dif_rc < 0);
    int32_t dividend = neg_val ? 0 : dif_lr;                     /* 18-bit signed */
    int32_t divisor = 2 * (cmp ? dif_rc : dif_lc);               /* {dif[16:0],1'b0}: 18-bit signed */
    if ((divisor & 0x3FFFF) == 0) return cmp ? 0x40 : 0xC0;      /* :156-163 */
    if (bitserial)
        return (uint8_t)orc_diven(18, 18, 8, 17, (uint64_t)(uint32_t)dividend, (uint64_t)(uint32_t)divisor);
    return (uint8_t)(floordiv64((int64_t)dividend * 128, divisor) & 0xFF);
}

/* bm_calc_uni.v:120-134 */
static int rtl_uni_ratio(uint16_t min1, uint16_t min2, int bitserial)
{
    if (bitserial) return (int)(orc_diven(17, 17, 11, 16, min1, min2) & 0x3FF);
    if (min2 == 0) return 2047 & 0x3FF;
    return (int)(((int64_t)min1 * 1024 / min2) & 0x3FF);
}

typedef struct { uint16_t min1, min2; uint8_t disp1, disp2, frac; } rec_t;

static int64_t g_sat_events = 0;
int64_t orc_bm_rtl_last_sat_events(void) { return g_sat_events; }

int orc_bm_rtl(const uint8_t *xl, const uint8_t *xr, int W, int H,
               const orc_bm_rtl_params *p, int16_t *disp)
{
    const int wsz = p->wsz, D = p->ndisp;
    if (wsz < 1 || wsz > 31 || !(wsz & 1) || D < 32 || D > 256 || (D & 31)) return -1;
    /* bm.v:232-259 secondary parameters */
    const int hwsz = wsz >> 1;
    const int hsad_wdt = W - D - 1;               /* columns x in [D, W-2] */
    const int sad_wdt = hsad_wdt - 2 * hwsz;
    const int sad_hgt = H - 2 * hwsz;
    for (int i = 0; i < W * H; i++) disp[i] = -1; /* fpga.c:105-106 memset 0xFF */
    g_sat_events = 0;
    if (sad_wdt <= 0 || sad_hgt <= 0) return 0;

    uint16_t *col = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)hsad_wdt * 34);
    rec_t *rec = (rec_t *)calloc((size_t)sad_wdt * sad_hgt, sizeof(rec_t));
    if (!col || !rec) { free(col); free(rec); return -2; }

    const int npass = D / 32;
    for (int ph = 0; ph < npass; ph++) {          /* bm_ibuf.v:143-189 dphase loop */
        const int last = (ph == npass - 1);
        for (int i = 0; i < sad_hgt; i++) {       /* output row index; centre row = hwsz + i */
            /* bm_ibuf.v:195-248 line schedule: first output row adds lines 0..wsz-1
             * (first one assigns), every later row subtracts line i-1 then adds line i+wsz-1. */
            int nops = (i == 0) ? wsz : 2;
            for (int o = 0; o < nops; o++) {
                int line, sub, first;
                if (i == 0) { line = o; sub = 0; first = (o == 0); }
                else        { line = (o == 0) ? i - 1 : i + wsz - 1; sub = (o == 0); first = 0; }
                const uint8_t *rl = xl + (size_t)line * W, *rr = xr + (size_t)line * W;
                for (int k = 0; k < hsad_wdt; k++) {
                    const int x = D + k;
                    const int lpix = rl[x] & 63;
                    uint16_t *c = col + (size_t)k * 34;
                    for (int j = 0; j < 34; j++) {    /* lane j <-> d = 32*ph + j - 1 (bm_calc_sad.v:353-418) */
                        const int d = 32 * ph + j - 1;
                        const int rpix = rr[x - d] & 63;
                        const int ad = abs(lpix - rpix);           /* dif6/abs7: bm_calc_sad.v:82-101 */
                        int v;
                        if (first) v = ad;                          /* :449 first_line */
                        else if (!sub) { v = c[j] + ad; if (v > 1023) { v = 1023; g_sat_events++; } }   /* upper_lim10 */
                        else           { v = c[j] - ad; if (v < 0) v = 0; }                             /* lower_lim10 */
                        c[j] = (uint16_t)v;
                    }
                }
            }
            /* horizontal window: bm_calc_sad.v:501-605 */
            uint32_t run[34];                         /* sliding add/sub, reset at line start (:571-587) */
            for (int j = 0; j < 34; j++) {
                run[j] = 0;
                for (int k = 0; k < 2 * hwsz; k++) run[j] += col[(size_t)k * 34 + j];
            }
            for (int xc = 0; xc < sad_wdt; xc++) {    /* centre column x = D + hwsz + xc */
                uint16_t sad[34];
                for (int j = 0; j < 34; j++) {
                    run[j] += col[(size_t)(xc + 2 * hwsz) * 34 + j];
                    if (xc > 0) run[j] -= col[(size_t)(xc - 1) * 34 + j];
                    sad[j] = (uint16_t)(run[j] > 0xFFFF ? 0xFFFF : run[j]);   /* limit16 (unreachable) */
                }
                det_t dt = rtl_det(sad);
                uint8_t fr = rtl_frac(dt.l, dt.min1, dt.r, p->bitserial_div);
                uint8_t d1 = (uint8_t)(((ph & 7) << 5) | dt.idx1), d2 = (uint8_t)(((ph & 7) << 5) | dt.idx2);
                rec_t *s = rec + (size_t)i * sad_wdt + xc;
                rec_t n;
                int upd;
                if (ph == 0) {                        /* bm_calc_upd.v:150-157 initial SAD */
                    n.min1 = dt.min1; n.disp1 = d1; n.min2 = dt.min2; n.disp2 = d2; upd = 1;
                } else {                              /* bm_calc_upd.v:125-207 */
                    int d1_lt_s1 = dt.min1 < s->min1, d2_lt_s1 = dt.min2 < s->min1;
                    int d1_lt_s2 = dt.min1 < s->min2, d2_lt_s2 = dt.min2 < s->min2;
                    int adj = (d1 == (uint8_t)(s->disp1 + 1));
                    if (d1_lt_s1 && d2_lt_s1) {
                        n.min1 = dt.min1; n.disp1 = d1; n.min2 = dt.min2; n.disp2 = d2; upd = 1;
                    } else if (d1_lt_s1 && !d2_lt_s1 && d2_lt_s2) {
                        n.min1 = dt.min1; n.disp1 = d1; upd = 1;
                        if (!adj) { n.min2 = s->min1; n.disp2 = s->disp1; } else { n.min2 = dt.min2; n.disp2 = d2; }
                    } else if (d1_lt_s1 && !d2_lt_s1 && !d2_lt_s2) {
                        n.min1 = dt.min1; n.disp1 = d1; upd = 1;
                        if (!adj) { n.min2 = s->min1; n.disp2 = s->disp1; } else { n.min2 = s->min2; n.disp2 = s->disp2; }
                    } else if (!d1_lt_s1 && d1_lt_s2 && d2_lt_s2) {
                        n.min1 = s->min1; n.disp1 = s->disp1; upd = 0;
                        if (!adj) { n.min2 = dt.min1; n.disp2 = d1; } else { n.min2 = dt.min2; n.disp2 = d2; }
                    } else if (!d1_lt_s1 && d1_lt_s2 && !d2_lt_s2) {
                        n.min1 = s->min1; n.disp1 = s->disp1; upd = 0;
                        if (!adj) { n.min2 = dt.min1; n.disp2 = d1; } else { n.min2 = s->min2; n.disp2 = s->disp2; }
                    } else {
                        n = *s; upd = 0;
                    }
                }
                n.frac = upd ? fr : s->frac;          /* bm_calc.v:313 */
                if (!last) { *s = n; continue; }

                /* last dphase: uniqueness filter + output formatter */
                uint8_t od = n.disp1, of = n.frac;
                if (p->uni_enb) {                     /* bm_calc.v:315-328 */
                    int ratio = rtl_uni_ratio(n.min1, n.min2, p->bitserial_div);
                    if (ratio > (p->uni_thr & 0x3FF)) {
                        if (!p->uni_mode) { od = 0x00; of = 0x00; } else { od = 0xFF; of = 0xFF; }
                    }
                }
                /* bm_obuf2.v:122-154 */
                int32_t depth = ((int32_t)od << 8) + (int8_t)of;
                int16_t out;
                if (depth <= 0) out = -1;
                else if (p->rtl_extended) out = (int16_t)(depth >> 4);
                else out = (int16_t)(((depth >> 4) & 0x0FFF) | ((depth & 0x8000) ? 0xF000 : 0));
                int xo = D + hwsz + xc + p->x_store_offset;   /* bm_obuf2.v:125-127, 288-301 (A1) */
                if (xo >= 0 && xo < W) disp[(size_t)(hwsz + i) * W + xo] = out;
            }
        }
    }
    free(col); free(rec);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* BM, cv::StereoBM profile (generic CV_16S path)           */
/* ------------------------------------------------------------------------ */
int orc_bm_cv(const uint8_t *pl, const uint8_t *pr, int W, int H,
              const orc_bm_cv_params *p, int16_t *disp)
{
    return orc_bm_cv_cost(pl, pr, W, H, p, disp, NULL);
}

/* same, also returning the winning SAD of every VALID pixel (the `cost` image cv::StereoBM hands to
 * validateDisparity); entries of invalid pixels are left untouched, as in OpenCV. */
int orc_bm_cv_cost(const uint8_t *pl, const uint8_t *pr, int W, int H,
                   const orc_bm_cv_params *p, int16_t *disp, int16_t *cost)
{
    const int wsz = p->wsz, D = p->ndisp, h = wsz >> 1, cap = p->prefilter_cap;
    if (wsz < 5 || !(wsz & 1) || D < 16 || (D & 15)) return -1;
    for (int i = 0; i < W * H; i++) disp[i] = -16;
    const int x0 = D - 1 + h, x1 = W - h;          /* valid x in [x0, x1) */
    if (x1 <= x0 || H <= 2 * h) return 0;
    int32_t *sad = (int32_t *)malloc(sizeof(int32_t) * (D + 2));
    /* per-row column sums (exact, no saturation): colsum[x][d], coltex[x], x in [D-1, W) */
    int32_t *colsum = (int32_t *)malloc(sizeof(int32_t) * (size_t)W * D);
    int32_t *coltex = (int32_t *)malloc(sizeof(int32_t) * (size_t)W);
    if (!sad || !colsum || !coltex) { free(sad); free(colsum); free(coltex); return -2; }
    for (int y = h; y < H - h; y++) {
        for (int x = D - 1; x < W; x++) {
            int32_t *cs = colsum + (size_t)x * D;
            for (int d = 0; d < D; d++) cs[d] = 0;
            coltex[x] = 0;
            for (int j = -h; j <= h; j++) {
                const int lv = pl[(size_t)(y + j) * W + x];
                const uint8_t *rp = pr + (size_t)(y + j) * W + x;
                coltex[x] += abs(lv - cap);
                for (int d = 0; d < D; d++) cs[d] += abs(lv - rp[-d]);
            }
        }
        for (int x = x0; x < x1; x++) {
            int32_t tex = 0;
            for (int d = 0; d < D; d++) sad[d] = 0;
            for (int i = -h; i <= h; i++) {
                const int32_t *cs = colsum + (size_t)(x + i) * D;
                tex += coltex[x + i];
                for (int d = 0; d < D; d++) sad[d] += cs[d];
            }
            int mind = -1; int32_t minsad = INT32_MAX;
            for (int d = D - 1; d >= 0; d--)          /* ties -> larger d */
                if (sad[d] < minsad) { minsad = sad[d]; mind = d; }
            if (tex < p->texture_threshold) continue;
            if (p->uniqueness_ratio > 0) {
                int32_t thresh = minsad + minsad * p->uniqueness_ratio / 100;
                int d;
                for (d = 0; d < D; d++)
                    if ((d < mind - 1 || d > mind + 1) && sad[d] <= thresh) break;
                if (d < D) continue;
            }
            int32_t pp = sad[mind > 0 ? mind - 1 : mind + 1];
            int32_t nn = sad[mind < D - 1 ? mind + 1 : mind - 1];
            int32_t den = pp + nn - 2 * minsad + abs(pp - nn);
            int32_t frac = den ? ((pp - nn) * 256) / den : 0;      /* C division: toward zero */
            disp[(size_t)y * W + x] = (int16_t)((mind * 256 + frac + 15) >> 4);
            if (cost) cost[(size_t)y * W + x] = (int16_t)minsad;
        }
    }
    free(sad); free(colsum); free(coltex);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* cv::StereoBM post filters as enabled at slam/src/core/main.cpp:210-212      */
/* (OpenCV calib3d: validateDisparity, filterSpeckles; restated from the       */
/*  published algorithm, pinned against cv2 4.13 in tests)                      */
/* ------------------------------------------------------------------------ */
void orc_validate_disparity(int16_t *disp, const int16_t *cost, int W, int H, int min_d, int ndisp, int disp12_max_diff)
{
    const int maxD = min_d + ndisp;
    const int minX1 = maxD > 0 ? maxD : 0, maxX1 = W + (min_d < 0 ? min_d : 0);
    const int INVALID = (min_d - 1) * 16;
    const int maxdiff = disp12_max_diff * 16;
    int *d2 = (int *)malloc(sizeof(int) * 2 * (size_t)W), *c2 = d2 + W;
    for (int y = 0; y < H; y++) {
        int16_t *dp = disp + (size_t)y * W;
        const int16_t *cp = cost + (size_t)y * W;
        for (int x = 0; x < W; x++) { d2[x] = INVALID; c2[x] = 0x7FFFFFFF; }
        for (int x = minX1; x < maxX1; x++) {          /* right-image disparity: cheapest match wins, first wins ties */
            const int d = dp[x], c = cp[x];
            if (d == INVALID) continue;
            const int x2 = x - ((d + 8) >> 4);
            if (c2[x2] > c) { c2[x2] = c; d2[x2] = d; }
        }
        for (int x = minX1; x < maxX1; x++) {          /* rounded towards -inf and +inf: invalid only if both disagree */
            const int d = dp[x];
            if (d == INVALID) continue;
            const int d0 = d >> 4, d1 = (d + 15) >> 4, x0 = x - d0, x1 = x - d1;
            if ((0 <= x0 && x0 < W && d2[x0] > INVALID && abs(d2[x0] - d) > maxdiff) &&
                (0 <= x1 && x1 < W && d2[x1] > INVALID && abs(d2[x1] - d) > maxdiff))
                dp[x] = (int16_t)INVALID;
        }
    }
    free(d2);
}

/* 4-connected regions of pixels != new_val whose neighbours differ by <= max_diff; regions of at most
 * max_size pixels become new_val.  (Region membership is the transitive closure of the pairwise relation,
 * so the result does not depend on the traversal order.) */
void orc_filter_speckles(int16_t *img, int W, int H, int new_val, int max_size, int max_diff)
{
    int *label = (int *)calloc((size_t)W * H, sizeof(int));
    int *stack = (int *)malloc(sizeof(int) * (size_t)W * H);
    int cur = 0;
    for (int i0 = 0; i0 < W * H; i0++) {
        if (img[i0] == new_val || label[i0]) continue;
        cur++;
        int sp = 0, count = 0;
        stack[sp++] = i0; label[i0] = cur;
        int first = i0; (void)first;
        /* pass 1: flood fill and count */
        int *members = stack;                          /* members are recorded in place: visited order */
        int head = 0;
        while (head < sp) {
            const int i = members[head++];
            count++;
            const int x = i % W, y = i / W, v = img[i];
            const int nb[4] = {x > 0 ? i - 1 : -1, x < W - 1 ? i + 1 : -1, y > 0 ? i - W : -1, y < H - 1 ? i + W : -1};
            for (int k = 0; k < 4; k++) {
                const int j = nb[k];
                if (j < 0 || label[j] || img[j] == new_val) continue;
                if (abs((int)img[j] - v) <= max_diff) { label[j] = cur; members[sp++] = j; }
            }
        }
        if (count <= max_size)
            for (int k = 0; k < sp; k++) img[members[k]] = (int16_t)new_val;
    }
    free(label); free(stack);
}

/* ------------------------------------------------------------------------ */
/* reprojection: Stereo.cpp:53-117, 157-198 + main.cpp:522-551 (+SensorData.cpp:50-58) */
/* ------------------------------------------------------------------------ */
/* projectDisparityTo3D (Stereo.cpp:157-182) for disp > 0.  The reference mixes float and double: every store into a
 * float variable rounds once; the volatiles pin that down under any optimisation level. */
static void orc_project_one(float u, float v, float d, const double *P_l, const double *P_r, float *out)
{
    const double fx_l = P_l[0], fy_l = P_l[5], cx_l = P_l[2], cy_l = P_l[6], Tx_l = P_l[3];
    const double fx_r = P_r[0], fy_r = P_r[5], cx_r = P_r[2], Tx_r = P_r[3];
    volatile float c = (float)(cx_r - cx_l);
    volatile float dc = d + c;                                     /* float add */
    volatile double nx = Tx_l / fx_l - Tx_r / fx_r;
    volatile double ny = Tx_l / fy_l - Tx_r / fy_r;
    volatile float Wx = (float)(nx / (double)dc);
    volatile float Wy = (float)(ny / (double)dc);
    volatile double ax = (double)u - cx_l, ay = (double)v - cy_l;
    out[0] = (float)(ax * (double)Wx);
    out[1] = (float)(ay * (double)Wy);
    out[2] = (float)(fx_l * (double)Wx);
}

/* transformPoint (Stereo.cpp:189-198): float products summed left to right, no contraction (built -ffp-contract=off) */
static void orc_transform_point(float *p, const float *t)
{
    volatile float x = p[0], y = p[1], z = p[2];
    volatile float a, b, c;
    a = t[0] * x; b = t[1] * y; a = a + b; b = t[2] * z; a = a + b; c = a + t[3];  p[0] = c;
    a = t[4] * x; b = t[5] * y; a = a + b; b = t[6] * z; a = a + b; c = a + t[7];  p[1] = c;
    a = t[8] * x; b = t[9] * y; a = a + b; b = t[10] * z; a = a + b; c = a + t[11]; p[2] = c;
}

/* StereoCameraModel.cpp:9-14 */
static const float ORC_LOCAL_T[12] = {0.0f, 0.0f, 1.0f, 0.0f, -1.0f, 0.0f, 0.0f, 0.0f, 0.0f, -1.0f, 0.0f, 0.0f};

/* Transform::isNull (Transform.cpp:88-95) */
static int orc_transform_is_null(const float *t)
{
    if (!t) return 1;
    for (int i = 0; i < 12; i++) if (t[i] != 0.0f) return 0;
    return 1;
}

/* dense consumer (main.cpp:522-551): local_T / pose = 3x4 row-major float transforms or NULL */
void orc_reproject_ex(const int16_t *disp, int W, int H, const double *P_l, const double *P_r,
                      int decim, const float *local_T, const float *pose, float *xyz)
{
    const int ow = W / decim, oh = H / decim;
    for (int row = 0; row < oh; row++) {
        for (int colx = 0; colx < ow; colx++) {
            const int16_t s = disp[(size_t)(row * decim) * W + colx * decim];
            volatile float d = (float)(s / 16.0f);                 /* main.cpp:529 */
            float p[3] = {NAN, NAN, NAN};
            if (d > 0.0f) {
                orc_project_one((float)(colx * decim), (float)(row * decim), d, P_l, P_r, p);
                if (isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2])) {      /* main.cpp:535-538 */
                    if (local_T) orc_transform_point(p, local_T);
                    if (pose) orc_transform_point(p, pose);
                } else if (local_T || pose) { p[0] = p[1] = p[2] = NAN; }      /* the consumer drops non-finite points */
            }
            float *o = xyz + ((size_t)row * ow + colx) * 3;
            o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
        }
    }
}

void orc_reproject(const int16_t *disp, int W, int H, const double *P_l, const double *P_r,
                   int decim, int apply_local, float *xyz)
{
    orc_reproject_ex(disp, W, H, P_l, P_r, decim, apply_local ? ORC_LOCAL_T : NULL, NULL, xyz);
}

/* generateKeypoints3DStereo (Stereo.cpp:53-117) with a dense-map depth method: uv = n (x, y) float pairs, mask NULL or n
 * bytes, local_T NULL (or all zero: Transform::isNull) = no transform.  A keypoint whose integer pixel lies outside the map
 * is undefined behaviour in the reference (cv::Mat::at without a check); the build defines it as a bad point (NaN). */
void orc_reproject_points(const int16_t *disp, int W, int H, const double *P_l, const double *P_r,
                          const float *uv, int n, const uint8_t *mask, float min_depth, float max_depth,
                          const float *local_T, float *xyz)
{
    const int have_T = !orc_transform_is_null(local_T);
    for (int i = 0; i < n; i++) {
        float *o = xyz + (size_t)3 * i;
        o[0] = o[1] = o[2] = NAN;
        if (mask && !mask[i]) continue;
        const float x = uv[2 * i], y = uv[2 * i + 1];
        if (!(x > -1.0f && x < (float)W && y > -1.0f && y < (float)H)) continue;   /* also rejects NaN coordinates */
        const int xi = (int)x, yi = (int)y;                        /* truncation toward zero, :79 */
        const int16_t s = disp[(size_t)yi * W + xi];
        volatile float d = (float)(s / 16.0f);
        if (d < 0) d = 0;                                          /* :81-83 */
        if (d != 0.0f) {
            float p[3];
            orc_project_one(x, y, d, P_l, P_r, p);                 /* the FLOAT keypoint coordinates, :94-97 */
            if (isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]) &&
                (min_depth < 0.0f || p[2] > min_depth) && (max_depth <= 0.0f || p[2] <= max_depth)) {
                if (have_T) orc_transform_point(p, local_T);
                o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * UVC payload: Xusb_ReceiveData, StereoBM/src/xusb_main.c:293-376.  The firmware's loops hard-code 640x480 and the
 * FPGA's DDR layouts (RECT: planar, XSBL: interleaved {L,R} words read back to front); here the inputs are the
 * planar images of the C ABI, the OUTPUT layout is the firmware's: dst[(row*2W + col + lr*W)*2] = Y, +1 = 0x80
 * (:313-327), disparity mode Y = (u8)(s16 >> 4) on the left half and 0x00 on the right (:356-372). */
void orc_pack_uvc(int mode, const uint8_t *L, const uint8_t *R, const int16_t *disp, int W, int H, uint8_t *frame)
{
    for (int row = 0; row < H; row++)
        for (int col = 0; col < W; col++) {
            size_t dl = ((size_t)row * 2 * W + col) * 2, dr = ((size_t)row * 2 * W + col + W) * 2;
            if (mode == 3) {
                int16_t t = disp[(size_t)row * W + col];
                t = (int16_t)(t >> 4);
                frame[dl] = (uint8_t)t; frame[dr] = 0x00;
            } else {
                frame[dl] = L[(size_t)row * W + col]; frame[dr] = R[(size_t)row * W + col];
            }
            frame[dl + 1] = 0x80; frame[dr + 1] = 0x80;
        }
}

/* ---- GFTT min-eigenvalue map: dvp/rtl/gftt*.v ------------------------------------------------
 * Stage by stage, with the RTL's bit widths:
 *  gftt_ibuf.v:385-415   three-line window; the Sobel stream covers image rows 1..H-2 (row2==2 gate, :385-398)
 *  gftt_sbl.v:118-205    dx = x-Sobel, dy = y-Sobel (s11.0); 0 at columns 0 and W-1 (first/last sample, :151, :198)
 *  gftt_eig.v:104-124    |dx|,|dy| (abs11) ; dx2=|dx|^2>>6, dy2=|dy|^2>>6, dxdy=|dx||dy|>>6  (the SIGN of dx*dy is dropped)
 *  gftt_box.v:185,232-262  3x3 box: horizontal 3-sum forced to 0 at columns 0 and W-1, three line sums, limit to 0xFFFF;
 *                        box rows are image rows 2..H-3 (line_ready, :131-140)
 *  gftt_eig.v:196-287    apc=a+c (17 bit) ; amc=|a-c| ; amc2=(amc^2)>>10 (22 bit) ; b2=(b^2)>>8 (24 bit) ;
 *                        s=min(amc2+b2, 0x3FFFFF) ; root = CORDIC sqrt IP, UnsignedFraction 32 -> 17 bit, Truncate
 *                        (ip/gftt_sqrt/gftt_sqrt.xci) on {0, s, 9'b0}:  in = s*2^9/2^31, out = sqrt(in)*2^16
 *                        = floor(sqrt(s * 2^10))  [IP modelled as an exact truncating square root: PARITY UNPINNED,
 *                        the encrypted IP model cannot be run here and the reference ships no eigen dump]
 *  gftt_eig.v:293-310    eig = apc - root[15:0] (18 bit signed): <0 -> 0 ; bit 16 -> 0xFFFF ; else eig[15:0]
 *  gftt_obuf.v:90-118, 262-275  per-frame maximum of the stream ; rows 2..H-3 written at a two-line offset, the bank is
 *                        zeroed by the firmware (fpga.c:107-108) so rows 0,1,H-2,H-1 read 0.
 * eig: W*H u16 row-major (what Fpga::receiveEigen returns, FPGA.cpp:281-296); *max_out = reg gftt.Max of the bank. */
static uint32_t orc_isqrt32(uint32_t x)
{
    uint32_t r = 0;
    for (int b = 15; b >= 0; b--) {
        const uint32_t t = r | (1u << b);
        if ((uint64_t)t * t <= x) r = t;
    }
    return r;
}

void orc_gftt_eig(const uint8_t *src, int W, int H, int src_stride, uint16_t *eig, uint16_t *max_out)
{
    const size_t np = (size_t)W * H;
    uint16_t *v[3];                       /* dx2, dy2, dxdy (u16) on the Sobel rows */
    uint32_t *hs[3];                      /* horizontal 3-sums (18 bit)             */
    for (int k = 0; k < 3; k++) { v[k] = (uint16_t *)calloc(np, sizeof(uint16_t)); hs[k] = (uint32_t *)calloc(np, sizeof(uint32_t)); }
    memset(eig, 0, np * sizeof(uint16_t));
    for (int y = 1; y <= H - 2; y++)
        for (int x = 1; x <= W - 2; x++) {
            const uint8_t *p0 = src + (size_t)(y - 1) * src_stride, *p1 = src + (size_t)y * src_stride, *p2 = src + (size_t)(y + 1) * src_stride;
            const int dx = (p0[x + 1] - p0[x - 1]) + 2 * (p1[x + 1] - p1[x - 1]) + (p2[x + 1] - p2[x - 1]);
            const int dy = (p2[x - 1] - p0[x - 1]) + 2 * (p2[x] - p0[x]) + (p2[x + 1] - p0[x + 1]);
            const uint32_t ax = (uint32_t)(dx < 0 ? -dx : dx) & 0x7FF, ay = (uint32_t)(dy < 0 ? -dy : dy) & 0x7FF;
            v[0][(size_t)y * W + x] = (uint16_t)((ax * ax) >> 6);
            v[1][(size_t)y * W + x] = (uint16_t)((ay * ay) >> 6);
            v[2][(size_t)y * W + x] = (uint16_t)((ax * ay) >> 6);
        }
    for (int k = 0; k < 3; k++)
        for (int y = 1; y <= H - 2; y++)
            for (int x = 1; x <= W - 2; x++)
                hs[k][(size_t)y * W + x] = (uint32_t)v[k][(size_t)y * W + x - 1] + v[k][(size_t)y * W + x] + v[k][(size_t)y * W + x + 1];
    uint16_t mx = 0;
    for (int y = 2; y <= H - 3; y++)
        for (int x = 0; x < W; x++) {
            uint32_t box[3];
            for (int k = 0; k < 3; k++) {
                const uint32_t s = hs[k][(size_t)(y - 1) * W + x] + hs[k][(size_t)y * W + x] + hs[k][(size_t)(y + 1) * W + x];
                box[k] = s > 0xFFFFu ? 0xFFFFu : s;
            }
            const uint32_t a = box[0], c = box[1], b = box[2];
            const uint32_t apc = a + c;
            const uint32_t amc = a > c ? a - c : c - a;
            const uint32_t amc2 = (amc * amc) >> 10, b2 = (b * b) >> 8;
            uint32_t s = amc2 + b2;
            if (s > 0x3FFFFFu) s = 0x3FFFFFu;
            const uint32_t root = orc_isqrt32(s << 10) & 0xFFFFu;
            const int32_t e = (int32_t)apc - (int32_t)root;
            const uint16_t o = e < 0 ? 0 : (e & 0x10000) ? 0xFFFF : (uint16_t)e;
            eig[(size_t)y * W + x] = o;
            if (o > mx) mx = o;
        }
    if (max_out) *max_out = mx;
    for (int k = 0; k < 3; k++) { free(v[k]); free(hs[k]); }
}
