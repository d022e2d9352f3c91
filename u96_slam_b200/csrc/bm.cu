// bm.cu -- SAD block matching + WTA + uniqueness/texture + sub-pixel, sm_100a.
//
// One kernel family, two semantic profiles selected at compile time:
//   RTL    : dvp/rtl/bm.v:232-259, bm_calc_sad.v:353-605 (AD, 10-bit saturating column sums,
//            horizontal window), bm_calc_det.v:124-426 (32-lane tournament, approximate min2),
//            bm_calc_frac.v:63-173 (sub-pixel, floor(128*num/den)), bm_calc_upd.v:119-211
//            (cross-dphase merge), bm_calc_uni.v:120-134 + bm_calc.v:315-328 (uniqueness),
//            bm_obuf2.v:122-154, 257-301 (s11.4 output, +1 column store offset).
//   OPENCV : cv::StereoBM as called from slam/src/core/main.cpp:197-217.
//
// Mapping.  A CTA owns (frame, x-tile of TX centre columns, y-band) and sweeps its band top to
// bottom.  The cost volume never touches HBM: per row step
//   phase 1  every thread owns column cx and a set of 8-disparity groups: VABSDIFF4 on the newest and
//            the oldest row of the window, widen to 2x16-bit, update the running COLUMN sums
//            (sub oldest, add newest; VIADDMNMX/VIMNMX.U16x2 give the RTL's [0,1023] saturation for
//            free), column sums live in shared memory [column][disparity slot] (u16).
//   phase 2  threads own (16-column segment, 8-disparity group): sliding horizontal window sum from
//            the column sums, packed 2x16-bit, and the minimum key (SAD<<16 | tie-break) of each
//            8-group -- exactly the level-3 winners of the RTL tournament.
//   phase 3  one thread per pixel finishes the tournament / cross-dphase merge / uniqueness /
//            sub-pixel division and stores the s16 disparity.
// HBM traffic is 2 B/px in + 2 B/px out; the kernel is bound by the integer pipe.
//
// Disparity slots: s in [0,D) <-> d = s.  RTL adds the two guard lanes of the first/last dphase
// (d = -1 -> slot D, d = D -> slot D+1; bm_calc_sad.v:353-418 lanes 0 and 33) which only feed the
// sub-pixel stage.  OPENCV adds the texture lane (|L - cap| -> slot D).  Slots are padded to DP = D+8.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace u96 {

constexpr int BM_NC = 128;        // column sums per CTA (centre columns + 2*hwsz halo)
constexpr int BM_THREADS = 256;
constexpr int BM_LS = 16;         // horizontal sliding segment length

struct BmArgs {
    const uint8_t *xl, *xr;
    int16_t *disp;
    int16_t *cost;                      // OPENCV: winning SAD of valid pixels (validateDisparity input) or null
    int pitch; size_t frame;            // input bytes
    int dpitch; size_t dframe;          // output elements
    int W, H, D, DP, NG;                // NG = DP/8 groups (last one = special lanes)
    int wsz, h, TX, ntx;
    int band_h, nbands;
    int col_lo, col_hi;                 // image x range in which column sums exist
    int ctr_lo, ctr_hi;                 // centre x range (inclusive)
    int y_lo, y_hi;                     // centre y range (inclusive)
    int x_store_offset, uni_enable, uni_mode, uni_thr, rtl_extended;
    int cap, tex_thr, uniq;
    int rlw;                            // words per shifted R copy
};

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }

// ---- shared memory carve-up ----
struct BmSmem {
    uint16_t *col;      // [BM_NC][DP]
    uint16_t *sad;      // [TX][DP]
    uint32_t *key;      // [TX][NG-1]
    uint32_t *rcp;      // [2 rows][4 copies][rlw] words, reversed R rows (index i <-> x = xr_max - i)
    uint8_t *lrow;      // [2][BM_NC]
};
__host__ __device__ inline size_t bm_smem_layout(int DP, int NG, int TX, int rlw, size_t *o_col, size_t *o_sad,
                                                 size_t *o_key, size_t *o_rcp, size_t *o_lrow)
{
    size_t o = 0;
    *o_col = o; o += (size_t)BM_NC * DP * 2;
    *o_sad = o; o += (size_t)TX * DP * 2;
    *o_key = o; o += (size_t)TX * (NG - 1) * 4;
    *o_rcp = o; o += (size_t)2 * 4 * rlw * 4;
    *o_lrow = o; o += 2 * BM_NC;
    return (o + 15) & ~(size_t)15;
}

// floor(a / b) for b > 0 or b < 0, exact
__device__ __forceinline__ int floordiv(int a, int b)
{
    int q = a / b, r = a - q * b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

template <int PROFILE, bool SAT>
__global__ void __launch_bounds__(BM_THREADS) k_bm(const BmArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    size_t o_col, o_sad, o_key, o_rcp, o_lrow;
    bm_smem_layout(a.DP, a.NG, a.TX, a.rlw, &o_col, &o_sad, &o_key, &o_rcp, &o_lrow);
    uint16_t *s_col = reinterpret_cast<uint16_t *>(smem_raw + o_col);
    uint16_t *s_sad = reinterpret_cast<uint16_t *>(smem_raw + o_sad);
    uint32_t *s_key = reinterpret_cast<uint32_t *>(smem_raw + o_key);
    uint32_t *s_rcp = reinterpret_cast<uint32_t *>(smem_raw + o_rcp);
    uint8_t *s_lrow = smem_raw + o_lrow;

    const int tid = threadIdx.x;
    const int tile = blockIdx.x, band = blockIdx.y, f = blockIdx.z;
    const int D = a.D, DP = a.DP, NG = a.NG, h = a.h, wsz = a.wsz;
    const int NGK = NG - 1;                      // groups that take part in the WTA

    const int ctr0 = a.ctr_lo + tile * a.TX;     // first centre column of this tile
    const int ntx = min(a.TX, a.ctr_hi - ctr0 + 1);
    const int xs = ctr0 - h;                     // image x of column index 0
    const int xr_max = xs + BM_NC;               // reversed R row: index i <-> x = xr_max - i
    const int yb0 = a.y_lo + band * a.band_h;
    const int yb1 = min(a.y_hi + 1, yb0 + a.band_h);   // exclusive

    const uint8_t *gl = a.xl + (size_t)f * a.frame;
    const uint8_t *gr = a.xr + (size_t)f * a.frame;
    int16_t *gout = a.disp + (size_t)f * a.dframe;

    // zero the column sums
    for (int i = tid; i < BM_NC * DP / 2; i += BM_THREADS) reinterpret_cast<uint32_t *>(s_col)[i] = 0;

    // phase-1 identity of this thread
    const int cx = tid & (BM_NC - 1);
    const int tg0 = tid / BM_NC;                 // 0 .. BM_THREADS/BM_NC-1
    const int x = xs + cx;
    const bool col_ok = (x >= a.col_lo && x <= a.col_hi);
    const int ibase = BM_NC - cx;                // reversed index of d = 0
    const int cpy = ibase & 3;
    const uint32_t *rc_new_base = s_rcp + cpy * a.rlw + ((ibase - cpy) >> 2);
    const int rl_bytes = a.rlw * 4;

    const int nsteps = (wsz - 1) + (yb1 - yb0);
    for (int r = 0; r < nsteps; r++) {
        const int y_add = yb0 - h + r;
        const int y_sub = y_add - wsz;
        const bool has_sub = (r >= wsz);
        __syncthreads();                          // previous step's readers of rows / sad / key are done
        // ---- stage the newest and the oldest row: L bytes, and 4 byte-shifted copies of reversed R ----
        {
            const uint8_t *rl_new = gl + (size_t)y_add * a.pitch, *rr_new = gr + (size_t)y_add * a.pitch;
            const uint8_t *rl_old = gl + (size_t)(has_sub ? y_sub : y_add) * a.pitch;
            const uint8_t *rr_old = gr + (size_t)(has_sub ? y_sub : y_add) * a.pitch;
            if (tid < BM_NC) {
                const int xx = xs + tid;
                const bool in = (xx >= 0 && xx < a.W);
                uint8_t vn = in ? rl_new[xx] : 0, vo = in ? rl_old[xx] : 0;
                if (PROFILE == U96_PROFILE_RTL) { vn &= 63; vo &= 63; }     // lr_din[13:8] is 6 bit
                s_lrow[tid] = vn;
                s_lrow[BM_NC + tid] = vo;
            }
            uint8_t *rcb = reinterpret_cast<uint8_t *>(s_rcp);
            for (int i = tid; i < rl_bytes; i += BM_THREADS) {
                const int xx = xr_max - i;
                const bool in = (xx >= 0 && xx < a.W);
                uint8_t vn = in ? rr_new[xx] : 0, vo = in ? rr_old[xx] : 0;
                if (PROFILE == U96_PROFILE_RTL) { vn &= 63; vo &= 63; }
#pragma unroll
                for (int c = 0; c < 4; c++) {     // copy c holds Rrev[m + c] at byte m
                    const int m = i - c;
                    if (m >= 0) {
                        rcb[(0 * 4 + c) * rl_bytes + m] = vn;
                        rcb[(1 * 4 + c) * rl_bytes + m] = vo;
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase 1: column sums ----
        if (col_ok) {
            const uint32_t ln = s_lrow[cx], lo = s_lrow[BM_NC + cx];
            const uint32_t ln4 = ln * 0x01010101u, lo4 = lo * 0x01010101u;
            uint16_t *colp = s_col + (size_t)cx * DP;
            for (int g = tg0; g < NG; g += BM_THREADS / BM_NC) {
                uint32_t rn0, rn1, ro0, ro1;
                if (g < NGK) {
                    const uint32_t *pn = rc_new_base + 2 * g;
                    rn0 = pn[0]; rn1 = pn[1];
                    ro0 = pn[4 * a.rlw]; ro1 = pn[4 * a.rlw + 1];
                } else {
                    const uint8_t *rb = reinterpret_cast<const uint8_t *>(s_rcp);    // copy 0 = plain reversed row
                    if (PROFILE == U96_PROFILE_RTL) {
                        // guard lanes d = -1 (slot D) and d = D (slot D+1); pad slots see R = L -> AD 0
                        const uint32_t gn = rb[ibase - 1] | ((uint32_t)rb[ibase + D] << 8);
                        const uint32_t go = rb[4 * rl_bytes + ibase - 1] | ((uint32_t)rb[4 * rl_bytes + ibase + D] << 8);
                        rn0 = prmt(gn, ln4, 0x5410); ro0 = prmt(go, lo4, 0x5410);
                    } else {
                        // texture lane: |L - cap| (slot D)
                        rn0 = prmt((uint32_t)a.cap, ln4, 0x5440); ro0 = prmt((uint32_t)a.cap, lo4, 0x5440);
                    }
                    rn1 = ln4; ro1 = lo4;
                }
                const uint32_t an0 = __vabsdiffu4(ln4, rn0), an1 = __vabsdiffu4(ln4, rn1);
                const uint32_t ao0 = has_sub ? __vabsdiffu4(lo4, ro0) : 0u, ao1 = has_sub ? __vabsdiffu4(lo4, ro1) : 0u;
                uint4 c = *reinterpret_cast<uint4 *>(colp + 8 * g);
                if (SAT) {
                    // sub oldest with floor 0, then add newest with ceiling 1023 (bm_calc_sad.v:449-466)
                    uint32_t w;
                    w = prmt(ao0, 0, 0x4140); c.x -= __vminu2(c.x, w);
                    w = prmt(ao0, 0, 0x4342); c.y -= __vminu2(c.y, w);
                    w = prmt(ao1, 0, 0x4140); c.z -= __vminu2(c.z, w);
                    w = prmt(ao1, 0, 0x4342); c.w -= __vminu2(c.w, w);
                    c.x = __viaddmin_u16x2(c.x, prmt(an0, 0, 0x4140), 0x03FF03FFu);
                    c.y = __viaddmin_u16x2(c.y, prmt(an0, 0, 0x4342), 0x03FF03FFu);
                    c.z = __viaddmin_u16x2(c.z, prmt(an1, 0, 0x4140), 0x03FF03FFu);
                    c.w = __viaddmin_u16x2(c.w, prmt(an1, 0, 0x4342), 0x03FF03FFu);
                } else {
                    // exact sums: one biased byte-wise delta, widened once
                    const uint32_t t0 = an0 + 0x80808080u - ao0, t1 = an1 + 0x80808080u - ao1;
                    c.x += prmt(t0, 0, 0x4140) - 0x00800080u;
                    c.y += prmt(t0, 0, 0x4342) - 0x00800080u;
                    c.z += prmt(t1, 0, 0x4140) - 0x00800080u;
                    c.w += prmt(t1, 0, 0x4342) - 0x00800080u;
                }
                *reinterpret_cast<uint4 *>(colp + 8 * g) = c;
            }
        }
        if (r < wsz - 1) continue;                // window not complete yet (uniform branch)
        __syncthreads();

        // ---- phase 2: horizontal window sums + 8-group minima ----
        {
            const int nseg = (ntx + BM_LS - 1) / BM_LS;
            for (int it = tid; it < nseg * NG; it += BM_THREADS) {
                const int g = it % NG, seg = it / NG;
                const int p0 = seg * BM_LS;                         // first centre (tile-relative)
                const int p1 = min(ntx, p0 + BM_LS);
                // window of centre p spans column indices [p, p + 2h]
                uint4 s = make_uint4(0, 0, 0, 0);
                const uint16_t *cg = s_col + 8 * g;
                for (int k = p0; k <= p0 + 2 * h; k++) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(cg + (size_t)k * DP);
                    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                }
                for (int p = p0; p < p1; p++) {
                    if (p > p0) {
                        const uint4 vn = *reinterpret_cast<const uint4 *>(cg + (size_t)(p + 2 * h) * DP);
                        const uint4 vo = *reinterpret_cast<const uint4 *>(cg + (size_t)(p - 1) * DP);
                        s.x += vn.x - vo.x; s.y += vn.y - vo.y; s.z += vn.z - vo.z; s.w += vn.w - vo.w;
                    }
                    *reinterpret_cast<uint4 *>(s_sad + (size_t)p * DP + 8 * g) = s;
                    if (g < NGK) {
                        // key = SAD<<16 | tie-break.  RTL: lower d wins (bm_calc_det.v:130-137 strict <);
                        // OPENCV: higher d wins (reverse scan).
                        const uint32_t d0 = 8 * g;
                        uint32_t t[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) t[k] = (PROFILE == U96_PROFILE_RTL) ? (d0 + k) : (0xFFFFu - (d0 + k));
                        const uint32_t k0 = (s.x << 16) | t[0], k1 = (s.x & 0xFFFF0000u) | t[1];
                        const uint32_t k2 = (s.y << 16) | t[2], k3 = (s.y & 0xFFFF0000u) | t[3];
                        const uint32_t k4 = (s.z << 16) | t[4], k5 = (s.z & 0xFFFF0000u) | t[5];
                        const uint32_t k6 = (s.w << 16) | t[6], k7 = (s.w & 0xFFFF0000u) | t[7];
                        uint32_t m = __vimin3_u32(k0, k1, k2);
                        m = __vimin3_u32(m, k3, k4);
                        m = __vimin3_u32(m, k5, k6);
                        m = min(m, k7);
                        s_key[(size_t)p * NGK + g] = m;
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase 3: per-pixel decision ----
        const int yc = y_add - h;
        for (int p = tid; p < ntx; p += BM_THREADS) {
            const uint32_t *kp = s_key + (size_t)p * NGK;
            const uint16_t *sp = s_sad + (size_t)p * DP;
            int out;
            if (PROFILE == U96_PROFILE_RTL) {
                uint32_t s_min1 = 0, s_min2 = 0, s_d1 = 0, s_d2 = 0;
                int s_frac = 0;
                const int npass = D >> 5;
                for (int ph = 0; ph < npass; ph++) {
                    // level-3 winners -> levels 4 and 5 of the tournament (bm_calc_det.v:268-377)
                    const uint32_t k0 = kp[4 * ph], k1 = kp[4 * ph + 1], k2 = kp[4 * ph + 2], k3 = kp[4 * ph + 3];
                    const uint32_t w0 = min(k0, k1), l0 = max(k0, k1);
                    const uint32_t w1 = min(k2, k3), l1 = max(k2, k3);
                    const uint32_t win = min(w0, w1), fin = max(w0, w1);
                    const uint32_t c1 = ((l1 >> 16) < (l0 >> 16)) ? l1 : l0;
                    const int d1 = win & 0xFFFF, dfin = fin & 0xFFFF, dc1 = c1 & 0xFFFF;
                    const bool adj0 = (dfin == d1 + 1) || (d1 == dfin + 1);
                    const bool adj1 = (dc1 == d1 + 1) || (d1 == dc1 + 1);
                    const bool pick1 = (((c1 >> 16) < (fin >> 16)) && !adj1) || adj0;     // bm_calc_det.v:404-411
                    const uint32_t m2k = pick1 ? c1 : fin;
                    const uint32_t min1 = win >> 16, min2 = m2k >> 16;
                    const uint32_t d2 = m2k & 0xFFFF;
                    // neighbours for the sub-pixel stage (guard lanes at the ends of the range)
                    const int L = sp[(d1 == 0) ? D : d1 - 1];
                    const int R = sp[(d1 == D - 1) ? D + 1 : d1 + 1];
                    const int C = (int)min1;
                    // bm_calc_frac.v:63-173
                    int q;
                    {
                        const bool cmp = L < R;
                        const bool neg = (L < C) || (R < C);
                        const int num = neg ? 0 : (L - R);
                        const int den = 2 * (cmp ? (R - C) : (L - C));
                        if (den == 0) q = cmp ? 64 : -64;
                        else q = floordiv(num * 128, den);
                    }
                    if (ph == 0) {
                        s_min1 = min1; s_d1 = d1; s_min2 = min2; s_d2 = d2; s_frac = q;
                    } else {
                        // bm_calc_upd.v:125-207
                        const bool d1_lt_s1 = min1 < s_min1, d2_lt_s1 = min2 < s_min1;
                        const bool d1_lt_s2 = min1 < s_min2, d2_lt_s2 = min2 < s_min2;
                        const bool adj = ((uint32_t)d1 & 0xFF) == ((s_d1 + 1) & 0xFF);
                        if (d1_lt_s1) {
                            if (d2_lt_s1)      { s_min2 = min2; s_d2 = d2; }
                            else if (d2_lt_s2) { if (!adj) { s_min2 = s_min1; s_d2 = s_d1; } else { s_min2 = min2; s_d2 = d2; } }
                            else               { if (!adj) { s_min2 = s_min1; s_d2 = s_d1; } }
                            s_min1 = min1; s_d1 = d1; s_frac = q;
                        } else if (d1_lt_s2) {
                            if (d2_lt_s2) { if (!adj) { s_min2 = min1; s_d2 = d1; } else { s_min2 = min2; s_d2 = d2; } }
                            else          { if (!adj) { s_min2 = min1; s_d2 = d1; } }
                        }
                    }
                }
                int od = (int)s_d1, of = s_frac;
                if (a.uni_enable) {
                    // bm_calc_uni.v:120-134: floor(1024*min1/min2) & 0x3FF, min2 == 0 -> 2047 & 0x3FF
                    const uint32_t ratio = (s_min2 == 0) ? 1023u : ((s_min1 * 1024u) / s_min2) & 0x3FFu;
                    if (ratio > (uint32_t)a.uni_thr) { od = a.uni_mode ? 0xFF : 0; of = a.uni_mode ? -1 : 0; }
                }
                // bm_obuf2.v:122-154
                const int depth = od * 256 + of;
                if (depth <= 0) out = -1;
                else if (a.rtl_extended) out = depth >> 4;
                else out = (int)(int16_t)(((depth >> 4) & 0x0FFF) | ((depth & 0x8000) ? 0xF000 : 0));
            } else {
                uint32_t best = 0xFFFFFFFFu;
                for (int g = 0; g < NGK; g++) best = min(best, kp[g]);
                const int mind = 0xFFFF - (int)(best & 0xFFFF);
                const int minsad = (int)(best >> 16);
                bool valid = (int)sp[D] >= a.tex_thr;                 // texture lane
                if (valid && a.uniq > 0) {
                    const int thresh = minsad + minsad * a.uniq / 100;
                    const int ga = max(0, mind - 1) >> 3, gb = min(D - 1, mind + 1) >> 3;
                    for (int g = 0; g < NGK && valid; g++) {
                        if (g == ga || g == gb) {
                            for (int k = 0; k < 8; k++) {
                                const int d = 8 * g + k;
                                if ((d < mind - 1 || d > mind + 1) && (int)sp[d] <= thresh) valid = false;
                            }
                        } else if ((int)(kp[g] >> 16) <= thresh) valid = false;
                    }
                }
                if (valid) {
                    const int pp = sp[mind > 0 ? mind - 1 : mind + 1];
                    const int nn = sp[mind < D - 1 ? mind + 1 : mind - 1];
                    const int den = pp + nn - 2 * minsad + abs(pp - nn);
                    const int frac = den ? ((pp - nn) * 256) / den : 0;       // C division, toward zero
                    out = (mind * 256 + frac + 15) >> 4;
                    if (a.cost) a.cost[(size_t)f * a.dframe + (size_t)yc * a.dpitch + ctr0 + p] = (int16_t)minsad;
                } else out = -16;
            }
            const int xo = ctr0 + p + ((PROFILE == U96_PROFILE_RTL) ? a.x_store_offset : 0);
            if (xo < a.W) gout[(size_t)yc * a.dpitch + xo] = (int16_t)out;
        }
    }
}

// Write the profile's invalid code into the frame border that no CTA of the BM kernels touches:
// rows [0,y_lo) and (y_hi,H), and in the valid rows the columns left of x_lo / right of x_hi.
// (bm_obuf2.v skips these bursts and the DISP bank keeps its 0xFF memset, fpga.c:105-106; OpenCV writes -16.)
constexpr int FB_ROWS = 16;
__global__ void __launch_bounds__(256) k_fill_border(int16_t *p, int dpitch, size_t dframe, int W, int H,
                                                     int y_lo, int y_hi, int x_lo, int x_hi, int16_t v)
{
    const int y0 = blockIdx.x * FB_ROWS, f = blockIdx.y;
    const int nright = W - 1 - x_hi, nside = x_lo + nright;        // border pixels of a valid row
    int16_t *base = p + (size_t)f * dframe;
    for (int yy = 0; yy < FB_ROWS; yy++) {
        const int y = y0 + yy;
        if (y >= H) break;
        int16_t *row = base + (size_t)y * dpitch;
        if (y < y_lo || y > y_hi) {
            for (int x = threadIdx.x; x < W; x += blockDim.x) row[x] = v;
        } else {
            for (int i = threadIdx.x; i < nside; i += blockDim.x) row[i < x_lo ? i : x_hi + 1 + (i - x_lo)] = v;
        }
    }
}

static bool bm_fill_args(const BmConfig &c, BmArgs &a)
{
    a.W = c.W; a.H = c.H; a.D = c.D; a.wsz = c.wsz; a.h = c.wsz >> 1;
    a.DP = c.D + 8; a.NG = a.DP / 8;
    a.TX = BM_NC - 2 * a.h;
    if (c.profile == U96_PROFILE_RTL) {
        a.col_lo = c.D; a.col_hi = c.W - 2;                    // bm.v:246 hsad_wdt = W - ndisp - 1
        a.ctr_lo = c.D + a.h; a.ctr_hi = c.W - 2 - a.h;
    } else {
        a.col_lo = c.D - 1; a.col_hi = c.W - 1;
        a.ctr_lo = c.D - 1 + a.h; a.ctr_hi = c.W - 1 - a.h;
    }
    a.y_lo = a.h; a.y_hi = c.H - 1 - a.h;
    if (a.ctr_hi < a.ctr_lo || a.y_hi < a.y_lo) return false;
    a.ntx = (a.ctr_hi - a.ctr_lo + 1 + a.TX - 1) / a.TX;
    const bool sat = (c.profile == U96_PROFILE_RTL) && (c.wsz * 63 > 1023);
    const int rows = a.y_hi - a.y_lo + 1;
    if (sat) { a.band_h = rows; a.nbands = 1; }                // saturating chain is sequential in y
    else { a.band_h = min(rows, 96); a.nbands = (rows + a.band_h - 1) / a.band_h; }
    a.x_store_offset = c.x_store_offset; a.uni_enable = c.uni_enable; a.uni_mode = c.uni_mode;
    a.uni_thr = c.uni_thr & 0x3FF; a.rtl_extended = c.rtl_extended;
    a.cap = c.cap; a.tex_thr = c.tex_thr; a.uniq = c.uniq;
    // reversed R row needs indices [0, NC + D]; words per copy == 8 (mod 32) spreads the 4 copies over banks
    int words = (BM_NC + c.D + 1 + 3) / 4 + 1;
    while ((words & 31) != 8) words++;
    a.rlw = words;
    return true;
}

// The fast path (bm_fast.cuh) covers numDisparities 64/128/256 (one 64-disparity slice per CTA of a cluster) for both profiles.
bool bm_fast_supported(const BmConfig &c)
{
    if (c.D != 64 && c.D != 128 && c.D != 256) return false;
    if (c.wsz < 3 || c.wsz > 31) return false;
    if (c.profile == U96_PROFILE_OPENCV) return c.wsz >= 5 && c.cap >= 1 && c.cap <= 63;
    return c.profile == U96_PROFILE_RTL;
}

// does launch_bm_fast hand this configuration to the fused-role kernel?
static bool bm_takes_fused(const BmConfig &c)
{
    static const int fused_env = getenv("U96_BM_FUSED") ? atoi(getenv("U96_BM_FUSED")) : -1;
    const bool sat = (c.profile == U96_PROFILE_RTL) && (c.wsz * 63 > 1023);
    return bm_fused_ok(c) && (fused_env == 1 || (fused_env != 0 && bm_fused_preferred(c, sat)));
}

int launch_bm_fast(const uint8_t *xl, const uint8_t *xr, int pitch, size_t frame, Img16 disp,
                   const BmConfig &c, int n, cudaStream_t s)
{
    // the fused-role kernel (bm_fused.cuh) wherever it is the faster one (profiles/r02_summary.md); U96_BM_FUSED = 0 / 1 forces
    // k_bm_fast / k_bm_fused where both apply (developer switch)
    // (a handful of pairs of the saturating 64-disparity chain, which cannot be cut into y-bands, is a latency problem: one CTA per SM,
    //  ~480 sequential rows; there the V/H-role kernel's row is 15 % shorter -- single pair 0.46 ms against 0.54 ms)
    static const int fused_env = getenv("U96_BM_FUSED") ? atoi(getenv("U96_BM_FUSED")) : -1;
    const bool few_sat64 = fused_env != 1 && c.profile == U96_PROFILE_RTL && c.D == 64 && c.wsz * 63 > 1023 && n <= 18;
    // ... unless the chain can be cut into y-bands (bm_fused.cuh: band functions, <= 8 pairs, scratch provided by the caller)
    const size_t band_need = bm_sat_scratch_bytes(c, n);
    const bool banded = fused_env != 0 && band_need > 0 && c.sat_scratch && c.sat_scratch_bytes >= band_need;
    if (bm_takes_fused(c) && (banded || !few_sat64)) return launch_bm_fused_rtl(xl, xr, pitch, frame, disp, c, n, s);
    if (c.D == 64) return launch_bm_fast_cs1(xl, xr, pitch, frame, disp, c, n, s);
    if (c.D == 128) return launch_bm_fast_cs2(xl, xr, pitch, frame, disp, c, n, s);
    return launch_bm_fast_cs4(xl, xr, pitch, frame, disp, c, n, s);
}

// Frames whose BM grid fills the device exactly once: the chunk pipeline of the C ABI cuts host batches at multiples of this so
// that no chunk runs the SMs half empty.  k_bm_fast: 2 CTAs per SM, one CTA per tile and 64-disparity slice; k_bm_fused: 4 / 2 / 1
// CTAs per SM at 64 / 128 / 256 disparities, one CTA per tile.
int bm_wave_frames(const BmConfig &c)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int h = c.wsz >> 1;
    const int ncen = (c.profile == U96_PROFILE_RTL) ? (c.W - 2 - h) - (c.D + h) + 1 : (c.W - 1 - h) - (c.D - 1 + h) + 1;
    if (ncen <= 0) return 1;
    const int tx5 = 160 - 2 * h, tx4 = 128 - 2 * h;
    if (bm_takes_fused(c)) {
        const int tiles = (ncen + tx5 - 1) / tx5, occ = (c.D == 64) ? 4 : (c.D == 128) ? 2 : 1;
        return std::max(1, (occ * sms + tiles - 1) / tiles);
    }
    const int tiles = std::min((ncen + tx5 - 1) / tx5, (ncen + tx4 - 1) / tx4) * std::max(1, c.D / 64);
    return std::max(1, (2 * sms + tiles - 1) / tiles);
}

int bm_smem_bytes(const BmConfig &c)
{
    BmArgs a;
    if (!bm_fill_args(c, a)) return 0;
    size_t o[5];
    return (int)bm_smem_layout(a.DP, a.NG, a.TX, a.rlw, &o[0], &o[1], &o[2], &o[3], &o[4]);
}

template <int PROFILE, bool SAT>
static void bm_launch_t(const BmArgs &a, dim3 grid, int smem, cudaStream_t s)
{
    cudaFuncSetAttribute(k_bm<PROFILE, SAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_bm<PROFILE, SAT><<<grid, BM_THREADS, smem, s>>>(a);
}

// The border of the disparity map (everything outside the valid rectangle) holds the invalid code.  The BM kernels never write
// it, so -- like the firmware, which memsets the DISP banks once at start-up (fpga.c:105-106) and lets bm_obuf2.v skip those bursts
// -- the C ABI fills it once per bank and configuration instead of once per submit.
int launch_bm_border(Img16 disp, const BmConfig &c, int n, cudaStream_t s)
{
    const int16_t inv = (c.profile == U96_PROFILE_RTL) ? (int16_t)-1 : (int16_t)-16;
    {
        BmArgs b;
        const bool ok = bm_fill_args(c, b);
        const int off = (c.profile == U96_PROFILE_RTL) ? c.x_store_offset : 0;
        // without a valid region the whole frame is border
        const int y_lo = ok ? b.y_lo : c.H, y_hi = ok ? b.y_hi : c.H, x_lo = ok ? b.ctr_lo + off : c.W, x_hi = ok ? b.ctr_hi + off : c.W;
        k_fill_border<<<dim3((c.H + FB_ROWS - 1) / FB_ROWS, n), 256, 0, s>>>(disp.p, disp.pitch, disp.frame, c.W, c.H, y_lo, y_hi, x_lo, x_hi, inv);
    }
    return 1;
}

int launch_bm(const uint8_t *xl, const uint8_t *xr, int pitch, size_t frame, Img16 disp,
              const BmConfig &c, int n, cudaStream_t s)
{
    if (bm_fast_supported(c) && !getenv("U96_BM_GENERIC")) {
        const int k = launch_bm_fast(xl, xr, pitch, frame, disp, c, n, s);
        if (k) return k;
    }
    BmArgs a;
    if (!bm_fill_args(c, a)) return 0;
    a.xl = xl; a.xr = xr; a.disp = disp.p; a.cost = c.cost; a.pitch = pitch; a.frame = frame;
    a.dpitch = disp.pitch; a.dframe = disp.frame;
    const int smem = bm_smem_bytes(c);
    dim3 grid(a.ntx, a.nbands, n);
    const bool sat = (c.profile == U96_PROFILE_RTL) && (c.wsz * 63 > 1023);
    if (c.profile == U96_PROFILE_RTL) {
        if (sat) bm_launch_t<U96_PROFILE_RTL, true>(a, grid, smem, s);
        else bm_launch_t<U96_PROFILE_RTL, false>(a, grid, smem, s);
    } else bm_launch_t<U96_PROFILE_OPENCV, false>(a, grid, smem, s);
    return 1;
}

}  // namespace u96
