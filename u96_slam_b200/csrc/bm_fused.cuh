// bm_fused.cuh -- SAD block matching with ONE compute role, sm_100a (RTL profile, 64 / 128 / 256 disparities, uniqueness off:
// the shipped register set, fpga.c:150-160).  Same arithmetic as bm_fast.cuh / bm.cu; different mapping.
//
// Why: ncu shows k_bm_fast bound by the shared-memory pipe (l1tex__data_pipe_lsu_wavefronts_mem_shared 77-80 % of peak, ALU pipe
// 62 %): its V warps (thread = column) and H warps (lane = segment x disparity group) hold the column sums in two different
// layouts, so every column sum crosses shared memory once as a store and twice as a load, on top of 16 byte-shifted R loads per
// column; beyond 64 disparities the slices of a tile are separate CTAs of a cluster that exchange records through DSMEM.
// Here ONE thread owns (segment of 8 columns) x (group of 8 disparities) for both steps, and a CTA holds ALL groups of its tile
// (NG = 8 / 16 / 32 groups: 160 / 320 / 640 compute threads) -- no cluster, no record ring:
//
//   * the 64 running column sums of its 8x8 patch never leave the thread's registers (bm_calc_sad.v:449-466);
//   * the R bytes of all 8 columns come from two aligned 8-byte words (the windows of neighbouring columns overlap by 7 bytes;
//     each column's window is funnel-shifted out of them) -- no byte-shifted copies, the rows are staged as they are;
//   * what crosses shared memory is the per-segment INCLUSIVE PREFIX SUM of the column sums (8 x 16 B per thread): the window sum
//     of pixel p over columns [p, p+2h] is   own block sum - own prefix before p   (registers)
//                                          + whole blocks in between                (their last prefix entry)
//                                          + prefix entry of column p+2h            (one 16-byte load per pixel);
//   * winner search on PACKED 16-bit minima (3 VIMNMX.U16x2 per 8 disparities instead of 8 key builds + 4 VIMNMX3); the exact
//     "lowest disparity among the minima" (bm_calc_det.v strict <, bm_calc_upd.v strict < across dphases) is recovered once per
//     pixel: a warp-local pass reduces the packed minima of 64 disparities to one key per (pixel, 64-disparity chunk), the thread
//     that owns the pixel takes the smallest chunk key, rescans the winning group's eight sums, fetches the winner's neighbours,
//     forms the sub-pixel fraction, formats and stores -- no record ring.
//
// CTA = 5*NG/8 compute warps + 1 staging warp + 1 guard warp (disparities -1 and D: bm_calc_sad.v lanes 0 and 33 of the first and
// last dphase, column-parallel).  Two barriers per image row: all warps after the column-sum step, the compute warps after the
// window-sum step (the prefix buffer is single: four 64-disparity CTAs fit an SM).  ~700 shared-memory wavefronts per tile row and
// 64 disparities against ~1010 of k_bm_fast.
#pragma once
#include "bm_fast.cuh"

namespace u96 {

constexpr int U_NC = 160, U_NSEG = 20;                 // tile columns, segments of 8 columns

template <int NG>
struct FusedSmem {
    static constexpr int D = 8 * NG, RLEN = U_NC + D + 16, SADP = D + 8, PMS = 9 * NG, NCH = NG / 8;
    uint4 pre[U_NC + 8][NG];               // inclusive prefix sums inside a segment: [column][group] = 8 x u16; 8 never-written pad columns: the
                                           // unrolled sweep of the last segment reads up to 7 columns past the tile for pixels nobody finishes
    uint16_t sad[U_NC][SADP];              // window sums of the row in flight; pitch 2D+16 B: rows skew over the banks
    uint32_t pmin[U_NSEG][PMS];            // packed minima [segment][pixel j][group] (odd | even disparities); 9*NG-word segment stride: the
                                           // segments of a warp store to disjoint banks
    uint32_t ckey[U_NC][NCH];              // per pixel: one key (min SAD << 8 | group) per 64-disparity chunk
    uint16_t guard[2][2][U_NC + 8];        // [buffer][d=-1 / d=D][column] column sums of the guard lanes
    uint8_t rrow[2][2][RLEN];              // [buffer][newest / oldest] R row segment: R[xs - D - 8 .. xs + 168)
    uint32_t lrow4[2][2][U_NC];            // L row segment, every pixel replicated into the four bytes of a word (VABSDIFF4 operand)
    uint8_t lrow[2][2][U_NC];              // L row segment (guard warp: 8 columns per word pair)
};

// bytes (i+1)..(i+8) of the 16-byte pair (a, b): the R window of column i of a segment (i = 7: b itself)
__device__ __forceinline__ uint2 r_window(uint2 a, uint2 b, int i)
{
    const int s = i + 1;                                   // 1..8, compile-time after unrolling
    uint32_t w0, w1, w2;
    if (s < 4) { w0 = a.x; w1 = a.y; w2 = b.x; }
    else if (s < 8) { w0 = a.y; w1 = b.x; w2 = b.y; }
    else return b;
    const int sh = (s & 3) * 8;
    if (sh == 0) return make_uint2(w0, w1);
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}

// resident CTAs per SM: 64 disparities 4 (54.5 KB, 72 registers), 128: 2, 256: 1
__host__ __device__ constexpr int fused_occupancy(int ng) { return ng == 8 ? 4 : ng == 16 ? 2 : 1; }

template <bool SAT, int NG>
__global__ void __launch_bounds__(U_NSEG * NG + 64, fused_occupancy(NG)) k_bm_fused(const FastArgs a)
{
    using SM = FusedSmem<NG>;
    constexpr int D = SM::D, RLEN = SM::RLEN, SADP = SM::SADP, NCH = SM::NCH;
    constexpr int NCT = U_NSEG * NG, CW = NCT / 32, NT = NCT + 64;    // compute threads / warps | + staging warp + guard warp
    constexpr int LG = (NG == 8) ? 3 : (NG == 16) ? 4 : 5;
    extern __shared__ __align__(16) unsigned char usm_raw[];
    SM &sm = *reinterpret_cast<SM *>(usm_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, band = blockIdx.y, f = blockIdx.z;
    const int h = a.h, wsz = a.wsz;
    const int ctr0 = a.ctr_lo + tile * a.TX;
    const int ntx = min(a.TX, a.ctr_hi - ctr0 + 1);
    const int xs = ctr0 - h;                      // image x of column 0
    const int xr0 = xs - (D + 8);                 // image x of staged R byte 0
    const int yb0 = a.y_lo + band * a.band_h;
    const int yb1 = min(a.y_hi + 1, yb0 + a.band_h);
    const int nsteps = (wsz - 1) + (yb1 - yb0);   // rows fed to the column sums
    const uint8_t *gl = a.xl + (size_t)f * a.frame;
    const uint8_t *gr = a.xr + (size_t)f * a.frame;
    int16_t *gout = a.disp + (size_t)f * a.dframe;
    const int pw = a.pitch >> 2;

    if (warp < CW) {
        // ======================================================================================
        // compute role: thread = (segment s of 8 columns, group g of 8 disparities)
        // ======================================================================================
        const int s = tid >> LG, g = tid & (NG - 1);
        uint4 c[8];                                                   // column sums: c[i] = column 8s+i, slots k <-> d = 8g+7-k
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = make_uint4(0, 0, 0, 0);
        const int aoff = 8 * (s - g) + D;                             // rrow index of word A (word B = +8)
        const int two_h = 2 * h;
        const int q0 = two_h >> 3;                                    // blocks fully inside the window of the segment's first pixel: s+1 .. s+q0-1
        const int jt = 8 - (two_h & 7);                               // pixels j >= jt reach one block further
        const bool seg_px = (8 * s < ntx);                            // this segment holds at least one pixel
        // chunk pass: the warp's 32 / NG segments x 8 pixels x NCH chunks are exactly 32 (pixel, 64-disparity chunk) items
        const int it_px = (warp * (32 >> LG) * 8) + (lane / NCH), it_ch = lane % NCH;
        // finishing pass: lane = pixel, on the first five compute warps
        const int px = 32 * warp + lane;
        const bool px_ok = (warp < 5) && (px < ntx);
        const int out_x = ctr0 + px + a.x_store_offset;
        int16_t *out_p = gout + (ptrdiff_t)(yb0 - (wsz - 1)) * (ptrdiff_t)a.dpitch + out_x;   // row of iteration 0 (not dereferenced before wsz-1)

        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");        // rows of iteration 0 are staged
        for (int it = 0; it < nsteps; it++) {
            const int b = it & 1;
            uint4 run;                                                // prefix sums of the 8 columns; after phase 1: the block sum
            // ---- phase 1: the newest and the oldest row enter the 64 column sums of this thread ----
            {
                const uint4 ln_a = *reinterpret_cast<const uint4 *>(&sm.lrow4[b][0][8 * s]), ln_b = *reinterpret_cast<const uint4 *>(&sm.lrow4[b][0][8 * s + 4]);
                const uint4 lo_a = *reinterpret_cast<const uint4 *>(&sm.lrow4[b][1][8 * s]), lo_b = *reinterpret_cast<const uint4 *>(&sm.lrow4[b][1][8 * s + 4]);
                const uint32_t lnw[8] = {ln_a.x, ln_a.y, ln_a.z, ln_a.w, ln_b.x, ln_b.y, ln_b.z, ln_b.w};
                const uint32_t low[8] = {lo_a.x, lo_a.y, lo_a.z, lo_a.w, lo_b.x, lo_b.y, lo_b.z, lo_b.w};
                const uint2 an = *reinterpret_cast<const uint2 *>(&sm.rrow[b][0][aoff]), bn = *reinterpret_cast<const uint2 *>(&sm.rrow[b][0][aoff + 8]);
                const uint2 ao = *reinterpret_cast<const uint2 *>(&sm.rrow[b][1][aoff]), bo = *reinterpret_cast<const uint2 *>(&sm.rrow[b][1][aoff + 8]);
                run = make_uint4(0, 0, 0, 0);
                uint4 *prow = &sm.pre[8 * s][g];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    col_update<SAT>(c[i], lnw[i], low[i], r_window(an, bn, i), r_window(ao, bo, i));
                    run.x += c[i].x; run.y += c[i].y; run.z += c[i].z; run.w += c[i].w;      // <= 8 * 1023 per half
                    prow[NG * i] = run;                               // pre[8s+i][g]
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");    // prefixes, guard sums and the next rows are in place
            if (it >= wsz - 1) {
                // ---- phase 2: window sums of this segment's 8 pixels for this thread's 8 disparities ----
                if (seg_px) {
                    // W(j) = own block sum - own columns before j  +  whole blocks s+1 .. s+Q(j)-1  +  prefix entry of column 8s+j+2h
                    uint4 bs = run;
                    for (int t = 1; t < q0; t++) {
                        const uint4 v = sm.pre[8 * (s + t) + 7][g];
                        bs.x += v.x; bs.y += v.y; bs.z += v.z; bs.w += v.w;
                    }
                    uint4 bq = make_uint4(0, 0, 0, 0);                // block s+q0: joins at pixel jt (the windows from there on reach past it)
                    if (jt < 8) bq = sm.pre[8 * (s + q0) + 7][g];
                    const uint4 *pe = &sm.pre[8 * s + two_h][g];      // prefix entry of the window's last column, pixel j: pe[NG*j]
                    uint16_t *sp = &sm.sad[8 * s][8 * g];
                    uint32_t *mp = &sm.pmin[s][g];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const uint4 e = pe[NG * j];
                        if (j == jt) { bs.x += bq.x; bs.y += bq.y; bs.z += bq.z; bs.w += bq.w; }
                        uint4 w;
                        w.x = bs.x + e.x; w.y = bs.y + e.y; w.z = bs.z + e.z; w.w = bs.w + e.w;
                        *reinterpret_cast<uint4 *>(sp + j * SADP) = w;
                        mp[NG * j] = __vminu2(__vminu2(w.x, w.y), __vminu2(w.z, w.w));       // low half: odd d, high half: even d
                        bs.x -= c[j].x; bs.y -= c[j].y; bs.z -= c[j].z; bs.w -= c[j].w;      // next pixel starts one column later
                    }
                }
                __syncwarp();
                // ---- chunk pass (warp-local): 8 packed minima -> one key per (pixel, 64-disparity chunk); lowest group wins ties ----
                if (it_px < ntx) {
                    const uint32_t *pm = &sm.pmin[it_px >> 3][NG * (it_px & 7) + 8 * it_ch];
                    const uint4 pa = *reinterpret_cast<const uint4 *>(pm), pb = *reinterpret_cast<const uint4 *>(pm + 4);
                    const uint32_t g0 = 8u * it_ch;
                    auto gk = [](uint32_t m, uint32_t gi) { return (min(m & 0xFFFFu, m >> 16) << 8) | gi; };
                    uint32_t best = __vimin3_u32(gk(pa.x, g0), gk(pa.y, g0 + 1), gk(pa.z, g0 + 2));
                    best = __vimin3_u32(best, gk(pa.w, g0 + 3), gk(pb.x, g0 + 4));
                    best = __vimin3_u32(best, gk(pb.y, g0 + 5), gk(pb.z, g0 + 6));
                    sm.ckey[it_px][it_ch] = min(best, gk(pb.w, g0 + 7));
                }
                asm volatile("bar.sync 2, %0;" ::"n"(NCT) : "memory");  // every prefix entry has been read (the next row may overwrite them); chunk keys and window sums are visible
                // ---- finishing pass: lane = pixel: winner, neighbours, sub-pixel fraction, output (bm_calc_det / upd / frac / obuf2) ----
                if (px_ok) {
                    uint32_t best = sm.ckey[px][0];
#pragma unroll
                    for (int k = 1; k < NCH; k++) best = min(best, sm.ckey[px][k]);
                    const uint32_t mv = best >> 8; const int gs = best & 0xFF;
                    const uint16_t *srow = &sm.sad[px][0];
                    const uint4 v = *reinterpret_cast<const uint4 *>(srow + 8 * gs);
                    // lowest disparity with SAD == mv inside the group = highest slot
                    auto sk = [&](uint32_t val, uint32_t k) { return (val << 3) | (7u - k); };
                    uint32_t bk = __vimin3_u32(sk(v.x & 0xFFFFu, 0), sk(v.x >> 16, 1), sk(v.y & 0xFFFFu, 2));
                    bk = __vimin3_u32(bk, sk(v.y >> 16, 3), sk(v.z & 0xFFFFu, 4));
                    bk = __vimin3_u32(bk, sk(v.z >> 16, 5), sk(v.w & 0xFFFFu, 6));
                    bk = min(bk, sk(v.w >> 16, 7));
                    const int d1 = 8 * gs + (int)(bk & 7u);            // 7 - k  ==  d - 8g
                    int L, R;
                    if (d1 == 0) { uint32_t acc = 0; for (int k = 0; k <= two_h; k++) acc += sm.guard[b][0][px + k]; L = (int)acc; }
                    else L = srow[slot_of(d1 - 1)];
                    if (d1 == D - 1) { uint32_t acc = 0; for (int k = 0; k <= two_h; k++) acc += sm.guard[b][1][px + k]; R = (int)acc; }
                    else R = srow[slot_of(d1 + 1)];
                    const int q = rtl_frac(L, R, (int)mv);
                    const int depth = d1 * 256 + q;                    // bm_obuf2.v:122-154
                    int out;
                    if (depth <= 0) out = -1;
                    else if (a.rtl_extended) out = depth >> 4;
                    else out = (int)(int16_t)(((depth >> 4) & 0x0FFF) | ((depth & 0x8000) ? 0xF000 : 0));
                    if (out_x < a.W) *out_p = (int16_t)out;
                }
            }
            out_p += a.dpitch;
        }
    } else if (warp == CW) {
        // ======================================================================================
        // staging role: rows of iteration it+1 (newest, oldest) -> shared memory as they are (6-bit masked: lr_din, bm_calc_sad.v:82-101)
        // ======================================================================================
        constexpr int RW = RLEN / 4, LW = U_NC / 4, ITEMS = 2 * RW + 2 * LW, NI = (ITEMS + 31) / 32;
        const uint32_t *p[NI]; bool ok0[NI], ok1[NI], on[NI]; int rt[NI], m[NI]; uint32_t so[NI];
#pragma unroll
        for (int j = 0; j < NI; j++) {
            const int item = lane + 32 * j;
            on[j] = item < ITEMS;
            const bool isr = item < 2 * RW;
            const int k = isr ? item : item - 2 * RW, wl = isr ? RW : LW;
            rt[j] = k / wl;                                           // 0 = newest row, 1 = oldest row
            const int q = k % wl;
            const int x0 = (isr ? xr0 : xs) + 4 * q;
            const int w0 = (x0 - (x0 & 3)) >> 2;
            m[j] = (x0 & 3) * 8;
            ok0[j] = on[j] && w0 >= 0 && w0 < pw; ok1[j] = on[j] && w0 + 1 >= 0 && w0 + 1 < pw;
            p[j] = reinterpret_cast<const uint32_t *>(isr ? gr : gl) + ((ptrdiff_t)(yb0 - h - (rt[j] ? wsz : 0)) * pw + w0);
            so[j] = (uint32_t)((isr ? offsetof(SM, rrow) + (size_t)rt[j] * RLEN : offsetof(SM, lrow) + (size_t)rt[j] * U_NC) + 4 * q);
        }
        uint32_t w0r[NI], w1r[NI];
        auto load = [&](int it) {
            const bool live_n = it < nsteps, live_o = live_n && it >= wsz;
#pragma unroll
            for (int j = 0; j < NI; j++) {
                const bool live = rt[j] ? live_o : live_n;
                w0r[j] = (live && ok0[j]) ? __ldg(p[j]) : 0u;
                w1r[j] = (live && ok1[j]) ? __ldg(p[j] + 1) : 0u;
                p[j] += pw;
            }
        };
        auto store = [&](int it) {
            const uint32_t boff = (it & 1) ? 2u : 0u;                 // buffer stride: 2 rows of the respective array
#pragma unroll
            for (int j = 0; j < NI; j++)
                if (on[j]) {
                    const bool isr = (lane + 32 * j) < 2 * RW;
                    const uint32_t v = __funnelshift_r(w0r[j], w1r[j], m[j]) & 0x3F3F3F3Fu;
                    *reinterpret_cast<uint32_t *>(usm_raw + so[j] + boff * (isr ? RLEN : U_NC)) = v;
                    if (!isr) {                                       // L pixels once more, replicated for the compute threads
                        const uint32_t o4 = (uint32_t)offsetof(SM, lrow4) + 4u * (so[j] - (uint32_t)offsetof(SM, lrow)) + boff * 4u * U_NC;
                        *reinterpret_cast<uint4 *>(usm_raw + o4) = make_uint4((v & 0xFFu) * 0x01010101u, ((v >> 8) & 0xFFu) * 0x01010101u,
                                                                              ((v >> 16) & 0xFFu) * 0x01010101u, (v >> 24) * 0x01010101u);
                    }
                }
        };
        load(0); store(0);
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        for (int it = 0; it < nsteps; it++) {
            load(it + 1);
            store(it + 1);
            asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        }
    } else {
        // ======================================================================================
        // guard role: column sums of d = -1 and d = D (bm_calc_sad.v lanes 0 and 33), 8 columns per item, column-parallel
        // ======================================================================================
        uint4 cg[2];                                                  // item = lane + 32*j: segment = item >> 1, which = item & 1
        cg[0] = cg[1] = make_uint4(0, 0, 0, 0);
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        for (int it = 0; it < nsteps; it++) {
            const int b = it & 1;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int item = lane + 32 * j;
                if (item < 2 * U_NSEG) {
                    const int s = item >> 1, which = item & 1;
                    uint2 rv[2];
#pragma unroll
                    for (int t = 0; t < 2; t++) {                     // newest, oldest
                        if (which) rv[t] = *reinterpret_cast<const uint2 *>(&sm.rrow[b][t][8 * s + 8]);          // R(x - D)
                        else {                                        // R(x + 1): bytes 1..8 of the pair at 8s+D+8
                            const uint2 u0 = *reinterpret_cast<const uint2 *>(&sm.rrow[b][t][8 * s + D + 8]);
                            const uint32_t u2 = (8 * s + D + 16 < RLEN) ? *reinterpret_cast<const uint32_t *>(&sm.rrow[b][t][8 * s + D + 16]) : 0u;
                            rv[t] = make_uint2(__funnelshift_r(u0.x, u0.y, 8), __funnelshift_r(u0.y, u2, 8));
                        }
                    }
                    const uint2 ln = *reinterpret_cast<const uint2 *>(&sm.lrow[b][0][8 * s]);
                    const uint2 lo = *reinterpret_cast<const uint2 *>(&sm.lrow[b][1][8 * s]);
                    const uint32_t an0 = __vabsdiffu4(ln.x, rv[0].x), an1 = __vabsdiffu4(ln.y, rv[0].y);
                    const uint32_t ao0 = __vabsdiffu4(lo.x, rv[1].x), ao1 = __vabsdiffu4(lo.y, rv[1].y);
                    uint4 &cc = cg[j];
                    if (SAT) {
                        cc.x -= __vminu2(cc.x, fprmt(ao0, 0, 0x4140)); cc.y -= __vminu2(cc.y, fprmt(ao0, 0, 0x4342));
                        cc.z -= __vminu2(cc.z, fprmt(ao1, 0, 0x4140)); cc.w -= __vminu2(cc.w, fprmt(ao1, 0, 0x4342));
                        cc.x = __viaddmin_u16x2(cc.x, fprmt(an0, 0, 0x4140), 0x03FF03FFu); cc.y = __viaddmin_u16x2(cc.y, fprmt(an0, 0, 0x4342), 0x03FF03FFu);
                        cc.z = __viaddmin_u16x2(cc.z, fprmt(an1, 0, 0x4140), 0x03FF03FFu); cc.w = __viaddmin_u16x2(cc.w, fprmt(an1, 0, 0x4342), 0x03FF03FFu);
                    } else {
                        cc.x += fprmt(an0, 0, 0x4140) - fprmt(ao0, 0, 0x4140); cc.y += fprmt(an0, 0, 0x4342) - fprmt(ao0, 0, 0x4342);
                        cc.z += fprmt(an1, 0, 0x4140) - fprmt(ao1, 0, 0x4140); cc.w += fprmt(an1, 0, 0x4342) - fprmt(ao1, 0, 0x4342);
                    }
                    *reinterpret_cast<uint4 *>(&sm.guard[b][which][8 * s]) = cc;             // columns 8s .. 8s+7 as u16
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        }
    }
}

template <bool SAT, int NG>
static inline void fused_go(const FastArgs &a, int n, cudaStream_t s)
{
    const int smem = (int)sizeof(FusedSmem<NG>);
    cudaFuncSetAttribute(k_bm_fused<SAT, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_bm_fused<SAT, NG><<<dim3(a.ntx_tiles, a.nbands, n), U_NSEG * NG + 64, smem, s>>>(a);
}

// RTL profile, 64 / 128 / 256 disparities, uniqueness filter off, window 9..31
static inline bool bm_fused_supported(const BmConfig &c)
{
    return c.profile == U96_PROFILE_RTL && (c.D == 64 || c.D == 128 || c.D == 256) && !c.uni_enable && c.wsz >= 9 && c.wsz <= 31;
}

static inline int launch_bm_fused(const uint8_t *xl, const uint8_t *xr, int pitch, size_t frame, Img16 disp,
                                  const BmConfig &c, int n, cudaStream_t s)
{
    FastArgs a;
    a.xl = xl; a.xr = xr; a.disp = disp.p; a.pitch = pitch; a.frame = frame; a.dpitch = disp.pitch; a.dframe = disp.frame;
    const int ng = c.D / 8;
    fast_fill_args<5>(a, c, n, 1, fused_occupancy(ng) * fast_sm_count());    // same tile (160 columns) and valid rectangle as k_bm_fast<NCW=5>
    if (a.ctr_hi < a.ctr_lo || a.y_hi < a.y_lo) return 0;
    const bool sat = c.wsz * 63 > 1023;
    if (ng == 8)       { if (sat) fused_go<true, 8>(a, n, s);  else fused_go<false, 8>(a, n, s); }
    else if (ng == 16) { if (sat) fused_go<true, 16>(a, n, s); else fused_go<false, 16>(a, n, s); }
    else               { if (sat) fused_go<true, 32>(a, n, s); else fused_go<false, 32>(a, n, s); }
    return 1;
}

}  // namespace u96
