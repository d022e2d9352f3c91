// bm_fast.cu -- warp-specialised, row-pipelined SAD block matching for the shipped configuration
// family (RTL profile, 64 disparities = two 32-lane dphases, any odd window <= 31), sm_100a.
//
// Same arithmetic as bm.cu (which stays the generic path); the difference is the mapping:
//
//   CTA = 8 warps, one (frame, 108-column tile, y-band).  Every loop iteration handles one image row
//   and ends in ONE __syncthreads; three pipeline stages are in flight on different buffers:
//
//   V warps 0-3 (thread = column, 128 columns): the running COLUMN sums of all 64(+2 guard) disparities
//     live in REGISTERS for the whole sweep (33 x u16x2).  Per row: 8-byte LDS of the byte-shifted R-row
//     copy, VABSDIFF4 against the broadcast L pixel for the newest and the oldest row of the window,
//     widen (PRMT), sub-oldest with floor 0 (VIMNMX.U16x2 + IADD) and add-newest with ceiling 1023
//     (VIADDMNMX.U16x2) = bm_calc_sad.v:449-466, then one conflict-free STS.128 per 8-disparity group.
//     They also prefetch the next two image rows (global -> registers at the top of the iteration,
//     registers -> 8 byte-shifted shared copies at the bottom), so no warp ever waits for HBM.
//   H warps 4-7 (lane = 7/8-pixel segment x 8-disparity group): sliding horizontal window sums from the
//     previous row's column sums (two LDS.128 per pixel step, packed 2x16 adds), group minimum key
//     (SAD<<16 | d) = the level-3 winners of the RTL tournament (bm_calc_det.v), results into a
//     warp-private shared buffer; after a __syncwarp the same warp finishes its own 28-32 pixels
//     (levels 4-5, approximate min2, cross-dphase merge bm_calc_upd.v, sub-pixel bm_calc_frac.v,
//     uniqueness bm_calc_uni.v, s11.4 output bm_obuf2.v) and stores the disparity row segment.
//
// Disparity slot order inside a group is DESCENDING (slot 8g+k <-> d = 8g+7-k) because the R window
// is read in natural memory order (x-d grows as d shrinks).  Guard lanes d=-1 / d=D (sub-pixel only)
// sit in slots 64/65 and their window sums are formed lazily by the few pixels whose winner is d=0 or
// d=D-1.
#include <cstdlib>

#include "common.cuh"

namespace u96 {

constexpr int F_D = 64;
constexpr int F_NGR = 8;           // regular 8-disparity groups
constexpr int F_DPS = 72;          // u16 slots per column in shared memory (64 + 2 guards + pad) -> 144 B rows
constexpr int F_CS = 272;          // bytes per byte-shifted R copy: >= NC + D + 16 (NC <= 192) and == 16 (mod 128)

struct FastArgs {
    const uint8_t *xl, *xr;
    int16_t *disp;
    int pitch; size_t frame;
    int dpitch; size_t dframe;
    int W, H, wsz, h, TX, LS, ntx_tiles;
    int nblk, fix_lo, fix_hi, fix_add;     // window = nblk whole blocks +/- columns [fix_lo, fix_hi)
    int band_h, nbands;
    int ctr_lo, ctr_hi, y_lo, y_hi;
    int x_store_offset, uni_enable, uni_mode, uni_thr, rtl_extended;
};

template <int NCW>
struct FastSmem {
    static constexpr int F_NC = 32 * NCW, F_NSEG = 4 * NCW;
    uint16_t col[2][F_NC][F_DPS];          // 36864 B   column sums, double buffered (V -> H)
    uint16_t sad[F_NC][F_DPS];             // 18432 B   window sums of the row in flight (H warp private rows)
    uint32_t key[2][F_NC * 4 + 16];        //  4224 B   group minima: [group half][pixel][4], halves 16 banks apart
    uint32_t guard[2][F_NC];               //  1024 B   column sums of the guard lanes (d=-1 | d=D << 16), double buffered
    uint8_t rcp[2][2][8][F_CS];            //  8704 B   [buffer][newest/oldest][byte shift][..] R row copies
    uint8_t lrow[2][2][F_NC];              //   512 B
    uint4 rec[2][F_NC];                    //  4096 B   per-pixel winner records (H -> V), double buffered
    uint16_t blk[F_NSEG + 4][F_DPS];       //  2880 B   per-segment block sums of the column sums (H warps)
};

__device__ __forceinline__ uint32_t fprmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }

// one 8-disparity group of one column: AD of newest/oldest row, saturating (or exact) column-sum update
template <bool SAT>
__device__ __forceinline__ void col_update(uint4 &c, uint32_t ln4, uint32_t lo4, uint2 rn, uint2 ro)
{
    const uint32_t an0 = __vabsdiffu4(ln4, rn.x), an1 = __vabsdiffu4(ln4, rn.y);
    const uint32_t ao0 = __vabsdiffu4(lo4, ro.x), ao1 = __vabsdiffu4(lo4, ro.y);
    if (SAT) {
        uint32_t w;
        w = fprmt(ao0, 0, 0x4140); c.x -= __vminu2(c.x, w);
        w = fprmt(ao0, 0, 0x4342); c.y -= __vminu2(c.y, w);
        w = fprmt(ao1, 0, 0x4140); c.z -= __vminu2(c.z, w);
        w = fprmt(ao1, 0, 0x4342); c.w -= __vminu2(c.w, w);
        c.x = __viaddmin_u16x2(c.x, fprmt(an0, 0, 0x4140), 0x03FF03FFu);
        c.y = __viaddmin_u16x2(c.y, fprmt(an0, 0, 0x4342), 0x03FF03FFu);
        c.z = __viaddmin_u16x2(c.z, fprmt(an1, 0, 0x4140), 0x03FF03FFu);
        c.w = __viaddmin_u16x2(c.w, fprmt(an1, 0, 0x4342), 0x03FF03FFu);
    } else {
        const uint32_t t0 = an0 + 0x80808080u - ao0, t1 = an1 + 0x80808080u - ao1;
        c.x += fprmt(t0, 0, 0x4140) - 0x00800080u;
        c.y += fprmt(t0, 0, 0x4342) - 0x00800080u;
        c.z += fprmt(t1, 0, 0x4140) - 0x00800080u;
        c.w += fprmt(t1, 0, 0x4342) - 0x00800080u;
    }
}

// slot position of disparity d inside a column / pixel record
__device__ __forceinline__ int slot_of(int d) { return (d & ~7) | (7 - (d & 7)); }

// NCW = number of V warps = number of H warps; the tile has 32*NCW column sums and 4*NCW horizontal segments
template <bool SAT, int LS, int NCW>
__global__ void __launch_bounds__(64 * NCW, (NCW <= 4) ? 3 : 2) k_bm_rtl64(const FastArgs a)
{
    constexpr int F_NC = 32 * NCW, F_NSEG = 4 * NCW, F_RWORDS = (F_NC + F_D + 16) / 8, F_LWORDS = F_NC / 4;
    static_assert(F_NC + F_D + 16 <= F_CS, "R copy stride too small");
    static_assert(2 * F_RWORDS + 2 * F_LWORDS <= F_NC, "not enough V threads to stage the rows");
    extern __shared__ __align__(16) unsigned char fsm_raw[];
    FastSmem<NCW> &sm = *reinterpret_cast<FastSmem<NCW> *>(fsm_raw);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, band = blockIdx.y, f = blockIdx.z;
    const int h = a.h, wsz = a.wsz;
    const int ctr0 = a.ctr_lo + tile * a.TX;
    const int ntx = min(a.TX, a.ctr_hi - ctr0 + 1);
    const int xs = ctr0 - h;                      // image x of column 0
    const int xr0 = xs - F_D - 7;                 // image x of staged R byte 0 (makes the byte shift of column cx equal cx & 7)
    const int yb0 = a.y_lo + band * a.band_h;
    const int yb1 = min(a.y_hi + 1, yb0 + a.band_h);
    const int nsteps = (wsz - 1) + (yb1 - yb0);   // rows fed to the column sums

    const uint8_t *gl = a.xl + (size_t)f * a.frame;
    const uint8_t *gr = a.xr + (size_t)f * a.frame;
    int16_t *gout = a.disp + (size_t)f * a.dframe;
    const int pw = a.pitch >> 2;                  // row pitch in 32-bit words

    if (warp < NCW) {
        // ======================================================================================
        // V role
        // ======================================================================================
        const int cx = tid;                                           // 0..127
        const int sh = cx & 7;                                        // byte shift of this column's R window
        const int qb = (cx >> 3) + F_D / 8;                           // 64-bit word of group 0
        uint4 c[F_NGR];
#pragma unroll
        for (int g = 0; g < F_NGR; g++) c[g] = make_uint4(0, 0, 0, 0);
        uint32_t cg = 0;                                              // guard lanes (d=-1 | d=D<<16)
        const bool v_active = (warp * 32 < ntx + 2 * h);              // partial last tile: idle warps only keep the barriers

        // ---- row staging: the first 2*RWORDS threads stage one 64-bit word of an R row each, the next 2*LWORDS one word of an L row ----
        const int lt = tid - 2 * F_RWORDS;
        const bool st_r = (tid < 2 * F_RWORDS), st_l = (lt >= 0 && lt < 2 * F_LWORDS);
        const int st_rt = st_r ? (tid / F_RWORDS) : (lt / F_LWORDS);     // 0 = newest row, 1 = oldest row
        const int st_q = st_r ? (tid % F_RWORDS) : (lt % F_LWORDS);
        uint32_t sw[5];                                               // prefetched aligned words
        auto stage_load = [&](int it) {
            const int y_add = yb0 - h + it;
            const bool has_sub = (it >= wsz);
            const int y = st_rt ? (y_add - wsz) : y_add;
            const bool live = (it < nsteps) && (st_rt == 0 || has_sub);
#pragma unroll
            for (int k = 0; k < 5; k++) sw[k] = 0;
            if (!live) return;
            if (st_r) {
                const uint32_t *row = reinterpret_cast<const uint32_t *>(gr + (size_t)y * a.pitch);
                const int x0 = xr0 + 8 * st_q;                        // image x of the first byte
                const int w0 = (x0 - (x0 & 3)) >> 2;                  // arithmetic shift: floor for negatives
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const int wi = w0 + k;
                    sw[k] = (wi >= 0 && wi < pw) ? __ldg(row + wi) : 0u;
                }
            } else if (st_l) {
                const uint32_t *row = reinterpret_cast<const uint32_t *>(gl + (size_t)y * a.pitch);
                const int x0 = xs + 4 * st_q;
                const int w0 = (x0 - (x0 & 3)) >> 2;
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const int wi = w0 + k;
                    sw[k] = (wi >= 0 && wi < pw) ? __ldg(row + wi) : 0u;
                }
            }
        };
        auto stage_store = [&](int it) {
            const int b = it & 1;
            if (st_r) {
                const int m = ((xr0 + 8 * st_q) & 3) * 8;             // misalignment of the global row segment
                uint32_t A[4];
#pragma unroll
                for (int k = 0; k < 4; k++) A[k] = __funnelshift_r(sw[k], sw[k + 1], m) & 0x3F3F3F3Fu;   // lr_din is 6 bit
#pragma unroll
                for (int s = 0; s < 8; s++) {                         // copy s holds bytes [8q+s, 8q+s+8)
                    const int k0 = s >> 2, sb = (s & 3) * 8;
                    uint2 v;
                    v.x = __funnelshift_r(A[k0], A[k0 + 1], sb);
                    v.y = __funnelshift_r(A[k0 + 1], (k0 + 2 < 4) ? A[k0 + 2] : 0u, sb);
                    *reinterpret_cast<uint2 *>(&sm.rcp[b][st_rt][s][8 * st_q]) = v;
                }
            } else if (st_l) {
                const int m = ((xs + 4 * st_q) & 3) * 8;
                const uint32_t v = __funnelshift_r(sw[0], sw[1], m) & 0x3F3F3F3Fu;
                *reinterpret_cast<uint32_t *>(&sm.lrow[b][st_rt][4 * st_q]) = v;
            }
        };

        stage_load(0);
        stage_store(0);
        __syncthreads();                                              // (A) rows of iteration 0 are staged

        for (int it = 0; it <= nsteps + 1; it++) {
            stage_load(it + 1);                                       // global loads in flight during the math
            if (it < nsteps && v_active) {
                const int b = it & 1;
                const uint32_t ln4 = (uint32_t)sm.lrow[b][0][cx] * 0x01010101u;
                const uint32_t lo4 = (uint32_t)sm.lrow[b][1][cx] * 0x01010101u;
                const uint2 *pn = reinterpret_cast<const uint2 *>(&sm.rcp[b][0][sh][0]) + qb;
                const uint2 *po = reinterpret_cast<const uint2 *>(&sm.rcp[b][1][sh][0]) + qb;
                uint16_t *colp = &sm.col[b][cx][0];
#pragma unroll
                for (int g = 0; g < F_NGR; g++) {
                    col_update<SAT>(c[g], ln4, lo4, pn[-g], po[-g]);
                    *reinterpret_cast<uint4 *>(colp + 8 * g) = c[g];
                }
                // guard lanes: d=-1 reads R(x+1), d=D reads R(x-D)          (bm_calc_sad.v lanes 0 and 33)
                {
                    const uint8_t *r0n = &sm.rcp[b][0][0][0], *r0o = &sm.rcp[b][1][0][0];
                    const uint32_t gn = r0n[cx + F_D + 8] | ((uint32_t)r0n[cx + 7] << 8);
                    const uint32_t go = r0o[cx + F_D + 8] | ((uint32_t)r0o[cx + 7] << 8);
                    const uint32_t an = __vabsdiffu4(ln4, fprmt(gn, ln4, 0x5410));
                    const uint32_t ao = __vabsdiffu4(lo4, fprmt(go, lo4, 0x5410));
                    if (SAT) {
                        cg -= __vminu2(cg, fprmt(ao, 0, 0x4140));
                        cg = __viaddmin_u16x2(cg, fprmt(an, 0, 0x4140), 0x03FF03FFu);
                    } else {
                        cg += fprmt(an, 0, 0x4140) - fprmt(ao, 0, 0x4140);
                    }
                    sm.guard[b][cx] = cg;
                }
            }
            // ---- finish the pixels of row it-2 from the H warps' records: sub-pixel, uniqueness, s11.4 output ----
            {
                const int r2 = it - 2;
                if (r2 >= wsz - 1 && cx < ntx) {
                    const uint4 rc = sm.rec[r2 & 1][cx];
                    const int L = (int)(rc.x & 0xFFFFu), R = (int)(rc.x >> 16);
                    const int C = (int)(rc.y & 0xFFFFu);
                    const uint32_t s_min1 = rc.y & 0xFFFFu, s_min2 = rc.y >> 16;
                    const int d1 = (int)rc.z;
                    int q;                                             // bm_calc_frac.v:63-173, floor(128*num/den)
                    {
                        const bool cmp = L < R;
                        const bool neg = (L < C) || (R < C);
                        const int num = neg ? 0 : (L - R);
                        const int den = 2 * (cmp ? (R - C) : (L - C));
                        if (den == 0) q = cmp ? 64 : -64;
                        else q = (int)floorf(__fdiv_rn((float)(num * 128), (float)den));   // exact: |q|<=64, den<2^17
                    }
                    int od = d1, of = q;
                    if (a.uni_enable) {                                // bm_calc_uni.v:120-134
                        const uint32_t ratio = (s_min2 == 0) ? 1023u : ((s_min1 * 1024u) / s_min2) & 0x3FFu;
                        if (ratio > (uint32_t)a.uni_thr) { od = a.uni_mode ? 0xFF : 0; of = a.uni_mode ? -1 : 0; }
                    }
                    const int depth = od * 256 + of;                   // bm_obuf2.v:122-154
                    int out;
                    if (depth <= 0) out = -1;
                    else if (a.rtl_extended) out = depth >> 4;
                    else out = (int)(int16_t)(((depth >> 4) & 0x0FFF) | ((depth & 0x8000) ? 0xF000 : 0));
                    const int yc = yb0 + (r2 - (wsz - 1));
                    const int xo = ctr0 + cx + a.x_store_offset;
                    if (xo < a.W) gout[(size_t)yc * a.dpitch + xo] = (int16_t)out;
                }
            }
            stage_store(it + 1);
            asm volatile("bar.sync 1, %0;" ::"n"(64 * NCW) : "memory");               // (B) one barrier per row
        }
    } else {
        // ======================================================================================
        // H role: horizontal sums + WTA for the row whose column sums were finished last iteration
        // ======================================================================================
        const int hw = warp - NCW;
        const int g = lane & 7;
        const int seg = hw * 4 + (lane >> 3);
        const int p0 = seg * LS;
        // tie-break constants: slot k of group g <-> d = 8g + 7 - k; lower d wins (bm_calc_det.v strict <)
        uint32_t t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = 8 * g + 7 - k;
        // pixel this lane finishes after the segment sweep
        const int fin_seg = hw * 4 + lane / LS, fin_j = lane % LS;
        const int fp = fin_seg * LS + fin_j;
        const bool fin_ok = (lane < 4 * LS) && (fp < ntx);
        const bool h_active = (hw * 4 * LS < ntx);                    // partial last tile: this warp has no pixel
        const int blk_last = (ntx - 1) / LS + a.nblk - 1;             // last block any active segment needs

        __syncthreads();                                              // (A)
        for (int it = 0; it <= nsteps + 1; it++) {
            const int r = it - 1;                                     // row index whose column sums are complete
            if (r >= wsz - 1 && r < nsteps) {
                const int cb = r & 1;
                const uint16_t *cg0 = &sm.col[cb][0][8 * g];
                // ---- block sums: every lane adds up the LS columns of its own segment (they are the "oldest"
                //      operands of its sweep anyway and stay in registers); the last H warp also covers blocks NSEG..NSEG+3 ----
                uint4 ov[LS];
                uint4 s = make_uint4(0, 0, 0, 0);
                if (hw * 4 <= blk_last) {
#pragma unroll
                    for (int j = 0; j < LS; j++) {
                        ov[j] = *reinterpret_cast<const uint4 *>(cg0 + (size_t)(p0 + j) * F_DPS);
                        s.x += ov[j].x; s.y += ov[j].y; s.z += ov[j].z; s.w += ov[j].w;
                    }
                    *reinterpret_cast<uint4 *>(&sm.blk[seg][8 * g]) = s;
                }
                if (hw == NCW - 1 && F_NSEG <= blk_last) {
                    const int eb = F_NSEG + (lane >> 3);
                    uint4 e = make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int j = 0; j < LS; j++) {
                        const uint4 v = *reinterpret_cast<const uint4 *>(cg0 + (size_t)(eb * LS + j) * F_DPS);
                        e.x += v.x; e.y += v.y; e.z += v.z; e.w += v.w;
                    }
                    *reinterpret_cast<uint4 *>(&sm.blk[eb][8 * g]) = e;
                }
                asm volatile("bar.sync 2, %0;" ::"n"(32 * NCW) : "memory");   // H warps only
                if (h_active) {
                // ---- window sum of the first pixel = whole blocks +/- a few single columns ----
                for (int k = 1; k < a.nblk; k++) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(&sm.blk[seg + k][8 * g]);
                    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                }
                for (int k = a.fix_lo; k < a.fix_hi; k++) {           // single columns added (fix_sign=+1) or removed
                    const uint4 v = *reinterpret_cast<const uint4 *>(cg0 + (size_t)(p0 + k) * F_DPS);
                    if (a.fix_add) { s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
                    else           { s.x -= v.x; s.y -= v.y; s.z -= v.z; s.w -= v.w; }
                }
                // sliding sweep, fully unrolled (LS is 7 or 8): the newest operand of step j+1 is fetched before
                // the key arithmetic of step j; the fetch past the last step stays inside the shared struct.
                const uint16_t *pn = cg0 + (size_t)(p0 + 2 * h + 1) * F_DPS;
                uint4 vn = *reinterpret_cast<const uint4 *>(pn);
                uint16_t *sp = &sm.sad[p0][8 * g];
                uint32_t *kp = &sm.key[g >> 2][p0 * 4 + (g & 3)];
#pragma unroll
                for (int j = 0; j < LS; j++) {
                    const uint4 sc = s;
                    s.x += vn.x - ov[j].x; s.y += vn.y - ov[j].y; s.z += vn.z - ov[j].z; s.w += vn.w - ov[j].w;
                    if (j + 1 < LS) vn = *reinterpret_cast<const uint4 *>(pn + (j + 1) * F_DPS);
                    const uint32_t k0 = (sc.x << 16) | t[0], k1 = (sc.x & 0xFFFF0000u) | t[1];
                    const uint32_t k2 = (sc.y << 16) | t[2], k3 = (sc.y & 0xFFFF0000u) | t[3];
                    const uint32_t k4 = (sc.z << 16) | t[4], k5 = (sc.z & 0xFFFF0000u) | t[5];
                    const uint32_t k6 = (sc.w << 16) | t[6], k7 = (sc.w & 0xFFFF0000u) | t[7];
                    uint32_t m = __vimin3_u32(k0, k1, k2);
                    m = __vimin3_u32(m, k3, k4);
                    m = __vimin3_u32(m, k5, k6);
                    m = min(m, k7);
                    *reinterpret_cast<uint4 *>(sp + j * F_DPS) = sc;
                    kp[j * 4] = m;
                }
                __syncwarp();
                // ---- per-pixel decision for this warp's own pixels ----
                if (fin_ok) {
                    const uint4 ka = *reinterpret_cast<const uint4 *>(&sm.key[0][fp * 4]);
                    const uint4 kb = *reinterpret_cast<const uint4 *>(&sm.key[1][fp * 4]);
                    uint32_t s_min1, s_min2, s_d1, s_d2;
                    {   // dphase 0: levels 4-5 of the tournament and the approximate min2 (bm_calc_det.v:268-411)
                        const uint32_t w0 = min(ka.x, ka.y), l0 = max(ka.x, ka.y);
                        const uint32_t w1 = min(ka.z, ka.w), l1 = max(ka.z, ka.w);
                        const uint32_t win = min(w0, w1), fin = max(w0, w1);
                        const uint32_t c1 = min(l0, l1);            // value tie -> l0 (lower d)
                        const int d1 = win & 0xFFFF, dfin = fin & 0xFFFF, dc1 = c1 & 0xFFFF;
                        const bool adj0 = (dfin == d1 + 1) || (d1 == dfin + 1);
                        const bool adj1 = (dc1 == d1 + 1) || (d1 == dc1 + 1);
                        const uint32_t m2 = ((((c1 >> 16) < (fin >> 16)) && !adj1) || adj0) ? c1 : fin;
                        s_min1 = win >> 16; s_d1 = d1; s_min2 = m2 >> 16; s_d2 = m2 & 0xFFFF;
                    }
                    {   // dphase 1, merged into the stored record (bm_calc_upd.v:125-207)
                        const uint32_t w0 = min(kb.x, kb.y), l0 = max(kb.x, kb.y);
                        const uint32_t w1 = min(kb.z, kb.w), l1 = max(kb.z, kb.w);
                        const uint32_t win = min(w0, w1), fin = max(w0, w1);
                        const uint32_t c1 = min(l0, l1);            // value tie -> l0 (lower d)
                        const int d1 = win & 0xFFFF, dfin = fin & 0xFFFF, dc1 = c1 & 0xFFFF;
                        const bool adj0 = (dfin == d1 + 1) || (d1 == dfin + 1);
                        const bool adj1 = (dc1 == d1 + 1) || (d1 == dc1 + 1);
                        const uint32_t m2 = ((((c1 >> 16) < (fin >> 16)) && !adj1) || adj0) ? c1 : fin;
                        const uint32_t min1 = win >> 16, min2 = m2 >> 16, d2 = m2 & 0xFFFF;
                        const bool d1_lt_s1 = min1 < s_min1, d2_lt_s1 = min2 < s_min1;
                        const bool d1_lt_s2 = min1 < s_min2, d2_lt_s2 = min2 < s_min2;
                        const bool adj = ((uint32_t)d1 == s_d1 + 1);
                        if (d1_lt_s1) {
                            if (d2_lt_s1)      { s_min2 = min2; s_d2 = d2; }
                            else if (d2_lt_s2) { if (!adj) { s_min2 = s_min1; s_d2 = s_d1; } else { s_min2 = min2; s_d2 = d2; } }
                            else               { if (!adj) { s_min2 = s_min1; s_d2 = s_d1; } }
                            s_min1 = min1; s_d1 = d1;
                        } else if (d1_lt_s2) {
                            if (d2_lt_s2) { if (!adj) { s_min2 = min1; s_d2 = d1; } else { s_min2 = min2; s_d2 = d2; } }
                            else          { if (!adj) { s_min2 = min1; s_d2 = d1; } }
                        }
                    }
                    (void)s_d2;
                    // neighbours of the final winner (the fraction follows min1: bm_calc.v:313)
                    const int d1 = (int)s_d1;
                    int L, R;
                    if (d1 == 0) {                                     // guard lane d=-1: lazy window sum
                        uint32_t acc = 0;
                        for (int k = 0; k <= 2 * h; k++) acc += sm.guard[cb][fp + k] & 0xFFFFu;
                        L = (int)acc;
                    } else L = sm.sad[fp][slot_of(d1 - 1)];
                    if (d1 == F_D - 1) {                               // guard lane d=D
                        uint32_t acc = 0;
                        for (int k = 0; k <= 2 * h; k++) acc += sm.guard[cb][fp + k] >> 16;
                        R = (int)acc;
                    } else R = sm.sad[fp][slot_of(d1 + 1)];
                    // winner record for the V warps, which finish the pixel one iteration later
                    sm.rec[r & 1][fp] = make_uint4((uint32_t)L | ((uint32_t)R << 16), s_min1 | (s_min2 << 16), (uint32_t)d1, 0u);
                }
                }   // h_active
            }
            asm volatile("bar.sync 1, %0;" ::"n"(64 * NCW) : "memory");               // (B)
        }
    }
}

bool bm_fast_supported(const BmConfig &c)
{
    return c.profile == U96_PROFILE_RTL && c.D == F_D && c.wsz >= 3 && c.wsz <= 31;
}

template <int NCW>
static int launch_bm_fast_t(const uint8_t *xl, const uint8_t *xr, int pitch, size_t frame, Img16 disp,
                            const BmConfig &c, int n, cudaStream_t s)
{
    constexpr int F_NC = 32 * NCW, F_NSEG = 4 * NCW;
    FastArgs a;
    a.xl = xl; a.xr = xr; a.disp = disp.p; a.pitch = pitch; a.frame = frame; a.dpitch = disp.pitch; a.dframe = disp.frame;
    a.W = c.W; a.H = c.H; a.wsz = c.wsz; a.h = c.wsz >> 1;
    a.TX = F_NC - 2 * a.h;
    a.LS = (a.TX + F_NSEG - 1) / F_NSEG;
    {   // window [0, 2h] of the first pixel of a segment in units of LS-column blocks
        const int wlen = 2 * a.h + 1, mfull = wlen / a.LS, rem = wlen % a.LS;
        if (rem <= a.LS - rem) { a.nblk = mfull; a.fix_lo = mfull * a.LS; a.fix_hi = wlen; a.fix_add = 1; }
        else                   { a.nblk = mfull + 1; a.fix_lo = wlen; a.fix_hi = (mfull + 1) * a.LS; a.fix_add = 0; }
        if (a.nblk == 0) { a.nblk = 1; a.fix_lo = wlen; a.fix_hi = a.LS; a.fix_add = 0; }      // window shorter than a block
    }
    a.ctr_lo = c.D + a.h; a.ctr_hi = c.W - 2 - a.h;                    // bm.v:246-252
    a.y_lo = a.h; a.y_hi = c.H - 1 - a.h;
    if (a.ctr_hi < a.ctr_lo || a.y_hi < a.y_lo) return 0;
    a.ntx_tiles = (a.ctr_hi - a.ctr_lo + 1 + a.TX - 1) / a.TX;
    const bool sat = c.wsz * 63 > 1023;
    const int rows = a.y_hi - a.y_lo + 1;
    if (sat) { a.band_h = rows; a.nbands = 1; }                        // saturating chain: sequential in y
    else { a.band_h = min(rows, 120); a.nbands = (rows + a.band_h - 1) / a.band_h; }
    a.x_store_offset = c.x_store_offset; a.uni_enable = c.uni_enable; a.uni_mode = c.uni_mode;
    a.uni_thr = c.uni_thr & 0x3FF; a.rtl_extended = c.rtl_extended;
    const int smem = (int)sizeof(FastSmem<NCW>);
    dim3 grid(a.ntx_tiles, a.nbands, n);
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        kern<<<grid, 64 * NCW, smem, s>>>(a);
    };
    if (a.LS == 7) { if (sat) go(k_bm_rtl64<true, 7, NCW>); else go(k_bm_rtl64<false, 7, NCW>); }
    else if (a.LS == 8) { if (sat) go(k_bm_rtl64<true, 8, NCW>); else go(k_bm_rtl64<false, 8, NCW>); }
    else return 0;
    return 1;
}

// tile width: 4 or 5 warps of columns, whichever wastes fewer column slots on this image width
int launch_bm_fast(const uint8_t *xl, const uint8_t *xr, int pitch, size_t frame, Img16 disp,
                   const BmConfig &c, int n, cudaStream_t s)
{
    const int h = c.wsz >> 1, ncen = c.W - 2 - h - (c.D + h) + 1;
    auto slots = [&](int ncw) { const int tx = 32 * ncw - 2 * h; return (ncen + tx - 1) / tx * 32 * ncw; };
    const char *force = getenv("U96_BM_NCW");
    int ncw = (slots(5) * 100 < slots(4) * 92) ? 5 : 4;              // the wider tile runs at lower occupancy: needs > 8 % less work
    if (force) ncw = atoi(force);
    if (ncw == 5) return launch_bm_fast_t<5>(xl, xr, pitch, frame, disp, c, n, s);
    return launch_bm_fast_t<4>(xl, xr, pitch, frame, disp, c, n, s);
}

}  // namespace u96
