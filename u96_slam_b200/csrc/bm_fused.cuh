// bm_fused.cuh -- SAD block matching with ONE compute role, sm_100a: 64 / 128 / 256 disparities, RTL profile with the uniqueness
// filter off (the shipped register set, fpga.c:150-160) and the cv::StereoBM profile (texture, exact uniqueness, mirrored
// sub-pixel neighbours).  Same arithmetic as bm_fast.cuh / bm.cu; different mapping.  This is the kernel that runs the headline
// configuration (profiles/r02_summary.md).
//
// Why: ncu shows k_bm_fast bound by the shared-memory pipe (l1tex__data_pipe_lsu_wavefronts_mem_shared 77-85 % of peak, ALU pipe
// 58-62 %): its V warps (thread = column) and H warps (lane = segment x disparity group) hold the column sums in two different
// layouts, so every column sum crosses shared memory once as a store and twice as a load, on top of 16 byte-shifted R loads per
// column; beyond 64 disparities the slices of a tile are separate CTAs of a cluster that exchange records through DSMEM.
// Here ONE thread owns (segment of 8 columns) x (group of 8 disparities) for both steps, and a CTA holds ALL groups of its tile
// (NG = 8 / 16 / 32 groups: 160 / 320 / 640 compute threads) -- no cluster, no record ring:
//
//   * the 64 running column sums of its 8x8 patch never leave the thread's registers (bm_calc_sad.v:449-466);
//   * the R pixels of all 8 columns come from two aligned 16-byte loads (the windows of neighbouring columns overlap by 7 pixels):
//     RTL variants and the cv::StereoBM variants whose column sums stay below 2048 stage the rows 16 bit per pixel ("WIDE": windows
//     are words in the lane format of the sums, the oldest row runs on the FMA pipe as fp16), the others as bytes (funnel-shifted
//     windows, PRMT widening) -- no byte-shifted copies in either form;
//   * what crosses shared memory is the per-segment INCLUSIVE PREFIX SUM of the column sums (8 x 16 B per thread): the window sum
//     of pixel p over columns [p, p+2h] is   own block sum - own prefix before p   (registers)
//                                          + whole blocks in between                (their last prefix entry)
//                                          + prefix entry of column p+2h            (one 16-byte load per pixel);
//   * winner search on PACKED 16-bit minima (3 VIMNMX.U16x2 per 8 disparities instead of 8 key builds + 4 VIMNMX3); the exact
//     "lowest disparity among the minima" (bm_calc_det.v strict <, bm_calc_upd.v strict < across dphases) is recovered once per
//     pixel: a warp-local pass reduces the packed minima of 64 disparities to one key per (pixel, 64-disparity chunk), the lane
//     that owns the pixel takes the smallest chunk key, rescans the winning group's eight sums, fetches the winner's neighbours,
//     forms the sub-pixel fraction, formats and stores -- no record ring;
//   * a handful of pairs of the saturating chain run in y-bands from exact start states (MODE 1 / 2, k_bm_chain; see below).
//
// CTA = 5*NG/8 compute warps + 1 staging warp + 1 guard warp (disparities -1 and D: bm_calc_sad.v lanes 0 and 33 of the first and
// last dphase, column-parallel).  Two barriers per image row: all warps after the column-sum step, the compute warps after the
// window-sum step (the prefix buffer is single: four 64-disparity CTAs fit an SM).  ~810 shared-memory wavefronts per tile row and
// 64 disparities against ~970 of k_bm_fast.
#pragma once
#include "bm_fast.cuh"

namespace u96 {

constexpr int U_NC = 160, U_NSEG = 20;                 // tile columns, segments of 8 columns

#ifndef U96_FUSED_PF
#define U96_FUSED_PF -1                    // window-sum step: prefix entries fetched this many pixels ahead (0: in place; -1: per variant, measured)
#endif
#ifndef U96_FUSED_KEEP
#define U96_FUSED_KEEP 1                   // 64 disparities: the winner key and the winner's group stay in registers across the second barrier
#endif
#ifndef U96_FUSED_WIDE
#define U96_FUSED_WIDE 1                   // 0: the RTL variants use the byte rows + PRMT widening of the cv::StereoBM variants
#endif

// WIDE (the RTL variants: 6-bit pixels; cv::StereoBM variants with window x 2 cap < 2048: column sums below 2048): the staged R rows
// are 16-bit per pixel, so the 8-disparity window of a column is four words already in the 2 x u16 lane format of the column sums -- aligned words for the odd columns of a segment, and for the even ones
// seven 16-bit funnel shifts shared by all four of them: no PRMT widening (64 per row and thread) on the ALU pipe, which binds
// this step, and the oldest row's |l - r| and saturating subtract run on the FMA pipe as fp16 (satsub_absdiff_u16x2).
template <int NG, bool WIDE, bool SAT, bool CV>
struct FusedSmem {
    static constexpr int D = 8 * NG, RLEN = U_NC + D + 16, RB = WIDE ? 2 * RLEN : RLEN, SADP = D + 8, PMS = 9 * NG, NCH = NG / 8;
    // pixels / segments with window sums: a saturating window is at least 17 wide, so a tile holds at most 144 pixels there
    static constexpr int NPX = (WIDE && SAT) ? 144 : U_NC, NPS = NPX / 8;
    uint4 pre[U_NC + 8][NG];               // inclusive prefix sums inside a segment: [column][group] = 8 x u16; 8 never-written pad columns: the
                                           // unrolled sweep of the last segment reads up to 7 columns past the tile for pixels nobody finishes
    uint16_t sad[NPX][SADP];               // window sums of the row in flight; pitch 2D+16 B: rows skew over the banks
    uint32_t pmin[NPS][PMS];               // packed minima [segment][pixel j][group] (odd | even disparities); 9*NG-word segment stride: the
                                           // segments of a warp store to disjoint banks
    static constexpr bool KEEP = U96_FUSED_KEEP && NG == 8;      // winner key and group stay in registers: ckey is not used
    uint32_t ckey[KEEP ? 4 : NPX][CV ? 2 : 1];     // per pixel: winner key (min SAD << 8 | group code) and, OPENCV, the smallest group minimum outside the
                                           // winner's group and its two neighbours
    uint32_t gt[2][U_NC + 8];              // output of the guard warp, [buffer]: RTL = u16 [d=-1 / d=D][column] column sums of the guard lanes;
                                           // OPENCV = u32 [1 + column] prefix sums over the columns of the texture column sums (entry 0 = 0)
    uint8_t rrow[2][2][RB];                // [buffer][newest / oldest] R row segment: R[xs - D - 8 .. xs + 168); WIDE: u16 per pixel
    uint32_t lrow4[2][2][U_NC];            // L row segment, every pixel replicated into the four bytes of a word (VABSDIFF4 operand); WIDE: into its two halves
    uint8_t lrow[2][2][(WIDE && !CV) ? 4 : U_NC];   // L row segment as bytes (guard warp: byte-row variants, and the OPENCV texture sums)
};

// |l - r| on two u16 lanes below 2048 and max(c - |l - r|, 0), both on the FMA pipe (see satsub_u16x2): HADD2 + HADD2.SAT with -|.| folded
__device__ __forceinline__ uint32_t satsub_absdiff_u16x2(uint32_t c, uint32_t l, uint32_t r)
{
    uint32_t d;
    asm("{\n\t.reg .b32 t, u;\n\tsub.f16x2 t, %2, %3;\n\tabs.f16x2 u, t;\n\tsub.sat.f16x2 %0, %1, u;\n\t}" : "=r"(d) : "r"(c), "r"(l), "r"(r));
    return d;
}

// |l - r| on two u16 lanes below 2048 as an integer bit pattern, on the FMA pipe; and max(c - m, 0) for such a pattern m
__device__ __forceinline__ uint32_t absdiff_f16_u16x2(uint32_t l, uint32_t r)
{
    uint32_t d;
    asm("{\n\t.reg .b32 t;\n\tsub.f16x2 t, %1, %2;\n\tabs.f16x2 %0, t;\n\t}" : "=r"(d) : "r"(l), "r"(r));
    return d;
}

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

// The saturating chain in y-bands (a handful of pairs: the sweep over the image height is a latency problem, 0.41-0.46 ms however idle the GPU).
// One row of the chain is f(c) = min(max(c - o, 0) + n, 1023) (bm_calc_sad.v:449-466): a clamp-add map c -> min(max(c + a, lo), hi).
// Such maps are closed under composition -- (a2, lo2, hi2) o (a1, lo1, hi1) = (a1 + a2, f2(lo1), f2(hi1)) -- so the effect of a whole
// band of rows on ANY start state is three numbers per column sum: the chain run from 0, the chain run from 1023, and the plain sum of
// n - o.  MODE 1 of the kernel computes those for every band in parallel (column-sum step only), k_bm_chain applies them band after
// band (one clamp-add per band and column sum) and MODE 2 then runs every band in parallel from its exact start state (MODE 0: the
// plain kernel).  Bit-exact by
// construction; three times the arithmetic, but spread over the whole GPU instead of four SMs.
__host__ __device__ constexpr size_t fused_state_block(int ng) { return (size_t)U_NC * ng + 2 * U_NSEG; }     // uint4 per (frame, band, tile): column sums + guard lanes

template <int PROFILE, bool SAT, int NG, bool WD, int MODE = 0>
__global__ void __launch_bounds__(U_NSEG * NG + 64, MODE == 1 ? 1 : fused_occupancy(NG)) k_bm_fused(const FastArgs a)
{
    constexpr bool COMPOSE = (MODE == 1);
    constexpr bool CV = (PROFILE == U96_PROFILE_OPENCV);
    constexpr bool WIDE = U96_FUSED_WIDE && WD;                  // RTL: always; OPENCV: when window x 2 cap < 2048 (column sums exact as fp16)
    static_assert(!COMPOSE || (WIDE && SAT && PROFILE == U96_PROFILE_RTL), "band functions exist for the saturating RTL chain only");
    constexpr bool KEEP = FusedSmem<NG, WIDE, SAT, CV>::KEEP;
    // same-box A/B (profiles/r02_summary.md): the 72-register variants without the wide rows lose 4 % to a two-pixel look-ahead
    constexpr int PF = (U96_FUSED_PF >= 0) ? U96_FUSED_PF : (NG == 8) ? ((WIDE && !CV) ? (SAT ? 1 : 2) : 0) : (NG == 16) ? ((WIDE && CV) ? 0 : 2) : 1;
    using SM = FusedSmem<NG, WIDE, SAT, CV>;
    constexpr int D = SM::D, RLEN = SM::RLEN, RB = SM::RB, SADP = SM::SADP, NCH = SM::NCH;
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
    // bands that carry the saturating chain: band > 0 starts behind its window fill, from the state k_bm_chain left for it
    // (MODE 0 = the plain kernel, 1 = band functions, 2 = bands from their start states: the plain kernel carries none of this)
    constexpr bool carry = (MODE != 0);
    const int it0 = (carry && band > 0) ? wsz - 1 : 0;
    const int gofs = carry ? band * a.band_h : 0;                // rows fed before this band's iteration 0 (is the oldest row part of the window?)
    constexpr size_t SBLK = fused_state_block(NG);
    uint4 *fn_blk = COMPOSE ? a.st_fn + (((size_t)f * (a.nbands - 1) + band) * a.ntx_tiles + tile) * 3 * SBLK : nullptr;
    const uint4 *sv_blk = (MODE == 2 && band > 0) ? a.st_val + (((size_t)f * a.nbands + band) * a.ntx_tiles + tile) * SBLK : nullptr;

    if (warp < CW) {
        // ======================================================================================
        // compute role: thread = (segment s of 8 columns, group g of 8 disparities)
        // ======================================================================================
        const int s = tid >> LG, g = tid & (NG - 1);
        uint4 c[8];                                                   // column sums: c[i] = column 8s+i, slots k <-> d = 8g+7-k
        uint4 ch[COMPOSE ? 8 : 1], ac[COMPOSE ? 8 : 1];               // MODE 1: the chain from 1023 and the sum of n - o (c is the chain from 0)
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = sv_blk ? sv_blk[(size_t)(8 * s + i) * NG + g] : make_uint4(0, 0, 0, 0);
        if (COMPOSE) {
#pragma unroll
            for (int i = 0; i < 8; i++) { ch[i] = make_uint4(0x03FF03FFu, 0x03FF03FFu, 0x03FF03FFu, 0x03FF03FFu); ac[i] = make_uint4(0, 0, 0, 0); }
        }
        const int aoff = 8 * (s - g) + D;                             // rrow index of word A (word B = +8)
        const int two_h = 2 * h;
        const int q0 = two_h >> 3;                                    // blocks fully inside the window of the segment's first pixel: s+1 .. s+q0-1
        const int jt = 8 - (two_h & 7);                               // pixels j >= jt reach one block further
        const bool seg_px = (8 * s < ntx);                            // this segment holds at least one pixel
        // chunk pass: the warp's 32 / NG segments x 8 pixels x NCH chunks are exactly 32 (pixel, 64-disparity chunk) items
        const int it_px = (warp * (32 >> LG) * 8) + (lane / NCH), it_ch = lane % NCH;
        // finishing pass: lane = pixel, on the first five compute warps
        const int px = 32 * warp + lane;
        const int out_x = ctr0 + px + (CV ? 0 : a.x_store_offset);
        const bool px_ok = (warp < 5) && (px < ntx) && (out_x < a.W);          // (the RTL's store offset can push the last pixel of a row out of the image)
        int16_t *out_p = gout + (ptrdiff_t)(yb0 - (wsz - 1) + it0) * (ptrdiff_t)a.dpitch + out_x;   // row of iteration it0 (not dereferenced before wsz-1)

        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");        // rows of iteration 0 are staged
        uint4 pf_e0, pf_e1, pf_la, pf_lb;                             // WIDE: oldest-row operands of the coming iteration
        auto prefetch_old = [&](int bb) {
            const uint4 *E = reinterpret_cast<const uint4 *>(&sm.rrow[bb][1][2 * aoff]);
            pf_e0 = E[0]; pf_e1 = E[1];
            pf_la = *reinterpret_cast<const uint4 *>(&sm.lrow4[bb][1][8 * s]); pf_lb = *reinterpret_cast<const uint4 *>(&sm.lrow4[bb][1][8 * s + 4]);
        };
        if (WIDE) prefetch_old(it0 & 1);
        for (int it = it0; it < nsteps; it++) {
            const int b = it & 1;
            uint4 run = make_uint4(0, 0, 0, 0);                       // prefix sums of the 8 columns; after phase 1: the block sum
            // ---- phase 1: the newest and the oldest row enter the 64 column sums of this thread ----
            if constexpr (COMPOSE) {
                // the same two rows enter the chain from 0, the chain from 1023 and the sum of n - o; no window sums in this mode
                {
                    const uint4 e0 = pf_e0, e1 = pf_e1, la = pf_la, lb = pf_lb;
                    const uint32_t ew[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
                    uint32_t ow[7];
#pragma unroll
                    for (int k = 0; k < 7; k++) ow[k] = __funnelshift_r(ew[k], ew[k + 1], 16);
                    const uint32_t lw[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const uint32_t *w = (i & 1) ? &ew[(i + 1) >> 1] : &ow[i >> 1];
                        const uint32_t m0 = absdiff_f16_u16x2(lw[i], w[0]), m1 = absdiff_f16_u16x2(lw[i], w[1]);
                        const uint32_t m2 = absdiff_f16_u16x2(lw[i], w[2]), m3 = absdiff_f16_u16x2(lw[i], w[3]);
                        c[i].x = satsub_u16x2(c[i].x, m0); c[i].y = satsub_u16x2(c[i].y, m1); c[i].z = satsub_u16x2(c[i].z, m2); c[i].w = satsub_u16x2(c[i].w, m3);
                        ch[i].x = satsub_u16x2(ch[i].x, m0); ch[i].y = satsub_u16x2(ch[i].y, m1); ch[i].z = satsub_u16x2(ch[i].z, m2); ch[i].w = satsub_u16x2(ch[i].w, m3);
                        ac[i].x = __vsub2(ac[i].x, m0); ac[i].y = __vsub2(ac[i].y, m1); ac[i].z = __vsub2(ac[i].z, m2); ac[i].w = __vsub2(ac[i].w, m3);
                    }
                }
                {
                    const uint4 *E = reinterpret_cast<const uint4 *>(&sm.rrow[b][0][2 * aoff]);
                    const uint4 e0 = E[0], e1 = E[1];
                    const uint4 la = *reinterpret_cast<const uint4 *>(&sm.lrow4[b][0][8 * s]), lb = *reinterpret_cast<const uint4 *>(&sm.lrow4[b][0][8 * s + 4]);
                    const uint32_t ew[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
                    uint32_t ow[7];
#pragma unroll
                    for (int k = 0; k < 7; k++) ow[k] = __funnelshift_r(ew[k], ew[k + 1], 16);
                    const uint32_t lw[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const uint32_t *w = (i & 1) ? &ew[(i + 1) >> 1] : &ow[i >> 1];
                        const uint32_t n0 = __vabsdiffu4(lw[i], w[0]), n1 = __vabsdiffu4(lw[i], w[1]), n2 = __vabsdiffu4(lw[i], w[2]), n3 = __vabsdiffu4(lw[i], w[3]);
                        c[i].x = __viaddmin_u16x2(c[i].x, n0, 0x03FF03FFu); c[i].y = __viaddmin_u16x2(c[i].y, n1, 0x03FF03FFu);
                        c[i].z = __viaddmin_u16x2(c[i].z, n2, 0x03FF03FFu); c[i].w = __viaddmin_u16x2(c[i].w, n3, 0x03FF03FFu);
                        ch[i].x = __viaddmin_u16x2(ch[i].x, n0, 0x03FF03FFu); ch[i].y = __viaddmin_u16x2(ch[i].y, n1, 0x03FF03FFu);
                        ch[i].z = __viaddmin_u16x2(ch[i].z, n2, 0x03FF03FFu); ch[i].w = __viaddmin_u16x2(ch[i].w, n3, 0x03FF03FFu);
                        ac[i].x = __vadd2(ac[i].x, n0); ac[i].y = __vadd2(ac[i].y, n1); ac[i].z = __vadd2(ac[i].z, n2); ac[i].w = __vadd2(ac[i].w, n3);
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
                prefetch_old(b ^ 1);
                continue;
            } else if constexpr (WIDE) {
                // oldest row first, on the FMA pipe: c = max(c - |l - r|, 0); then the newest row: c = min(c + |l - r|, 1023) on the ALU pipe
                // (exact sums, window <= 16: the same without the ceiling -- the subtraction never clamps and the add is a plain one)
                uint4 *prow = &sm.pre[8 * s][g];
                {
                    const uint4 e0 = pf_e0, e1 = pf_e1, la = pf_la, lb = pf_lb;
                    const uint32_t ew[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
                    uint32_t ow[7];                                   // one pixel ahead: halves 1..14
#pragma unroll
                    for (int k = 0; k < 7; k++) ow[k] = __funnelshift_r(ew[k], ew[k + 1], 16);
                    const uint32_t lw[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const uint32_t *w = (i & 1) ? &ew[(i + 1) >> 1] : &ow[i >> 1];
                        c[i].x = satsub_absdiff_u16x2(c[i].x, lw[i], w[0]); c[i].y = satsub_absdiff_u16x2(c[i].y, lw[i], w[1]);
                        c[i].z = satsub_absdiff_u16x2(c[i].z, lw[i], w[2]); c[i].w = satsub_absdiff_u16x2(c[i].w, lw[i], w[3]);
                    }
                }
                {
                    const uint4 *E = reinterpret_cast<const uint4 *>(&sm.rrow[b][0][2 * aoff]);
                    const uint4 e0 = E[0], e1 = E[1];
                    const uint4 la = *reinterpret_cast<const uint4 *>(&sm.lrow4[b][0][8 * s]), lb = *reinterpret_cast<const uint4 *>(&sm.lrow4[b][0][8 * s + 4]);
                    const uint32_t ew[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
                    uint32_t ow[7];                                   // one pixel ahead: halves 1..14
#pragma unroll
                    for (int k = 0; k < 7; k++) ow[k] = __funnelshift_r(ew[k], ew[k + 1], 16);
                    const uint32_t lw[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
                    run = make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const uint32_t *w = (i & 1) ? &ew[(i + 1) >> 1] : &ow[i >> 1];
                        if (SAT) {
                            c[i].x = __viaddmin_u16x2(c[i].x, __vabsdiffu4(lw[i], w[0]), 0x03FF03FFu); c[i].y = __viaddmin_u16x2(c[i].y, __vabsdiffu4(lw[i], w[1]), 0x03FF03FFu);
                            c[i].z = __viaddmin_u16x2(c[i].z, __vabsdiffu4(lw[i], w[2]), 0x03FF03FFu); c[i].w = __viaddmin_u16x2(c[i].w, __vabsdiffu4(lw[i], w[3]), 0x03FF03FFu);
                        } else {
                            c[i].x += __vabsdiffu4(lw[i], w[0]); c[i].y += __vabsdiffu4(lw[i], w[1]); c[i].z += __vabsdiffu4(lw[i], w[2]); c[i].w += __vabsdiffu4(lw[i], w[3]);
                        }
                        run.x += c[i].x; run.y += c[i].y; run.z += c[i].z; run.w += c[i].w;
                        prow[NG * i] = run;
                    }
                }
            } else {
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
                    // every load is issued ahead of its use: shared loads do not move above the shared stores of the pixel before,
                    // and one load in flight per warp leaves the warp waiting on the scoreboard (ncu: 11 % of the kernel's samples)
                    const uint4 *pe = &sm.pre[8 * s + two_h][g];      // prefix entry of the window's last column, pixel j: pe[NG*j]
                    const uint4 zero4 = make_uint4(0, 0, 0, 0);
                    uint4 e = zero4, e1 = zero4;
                    if (PF >= 1) e = pe[0];
                    if (PF >= 2) e1 = pe[NG];
                    const uint4 v1 = (q0 > 1) ? sm.pre[8 * (s + 1) + 7][g] : zero4;          // whole blocks s+1 .. s+q0-1 (window <= 31: at most two)
                    const uint4 v2 = (q0 > 2) ? sm.pre[8 * (s + 2) + 7][g] : zero4;
                    const uint4 bq = (jt < 8) ? sm.pre[8 * (s + q0) + 7][g] : zero4;         // block s+q0: joins at pixel jt (the windows from there on reach past it)
                    uint4 bs = run;
                    if (q0 > 1) { bs.x += v1.x; bs.y += v1.y; bs.z += v1.z; bs.w += v1.w; }
                    if (q0 > 2) { bs.x += v2.x; bs.y += v2.y; bs.z += v2.z; bs.w += v2.w; }
                    uint16_t *sp = &sm.sad[8 * s][8 * g];
                    uint32_t *mp = &sm.pmin[s][g];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        uint4 e2 = zero4;
                        if (PF == 0) e = pe[NG * j];
                        if (PF == 1 && j + 1 < 8) e1 = pe[NG * (j + 1)];
                        if (PF == 2 && j + 2 < 8) e2 = pe[NG * (j + 2)];
                        if (j == jt) { bs.x += bq.x; bs.y += bq.y; bs.z += bq.z; bs.w += bq.w; }
                        uint4 w;
                        w.x = bs.x + e.x; w.y = bs.y + e.y; w.z = bs.z + e.z; w.w = bs.w + e.w;
                        *reinterpret_cast<uint4 *>(sp + j * SADP) = w;
                        mp[NG * j] = __vminu2(__vminu2(w.x, w.y), __vminu2(w.z, w.w));       // low half: odd d, high half: even d
                        bs.x -= c[j].x; bs.y -= c[j].y; bs.z -= c[j].z; bs.w -= c[j].w;      // next pixel starts one column later
                        if (PF >= 1) e = e1;
                        if (PF >= 2) e1 = e2;
                    }
                }
                __syncwarp();
                uint32_t best_r = 0, um_r = 0;                        // NG == 8: the chunk-pass lane is the finishing lane of the same pixel
                uint4 v_r = make_uint4(0, 0, 0, 0);
                // ---- chunk pass (warp-local, lane = (pixel, 64-disparity chunk)): 8 packed minima -> the chunk's key; the NCH lanes of
                //      a pixel then agree on the winner by shuffles.  Group code = group (RTL: the lower disparity wins a tie,
                //      bm_calc_det.v / bm_calc_upd.v strict <) or 255 - group (OPENCV: the higher one, reverse scan) ----
                {
                    const bool live = it_px < ntx;
                    const int rp = live ? it_px : 0;                           // idle lanes load nothing (their result is dropped)
                    const uint32_t *pm = &sm.pmin[rp >> 3][NG * (rp & 7) + 8 * it_ch];
                    const uint4 none = make_uint4(~0u, ~0u, ~0u, ~0u);
                    const uint4 pa = live ? *reinterpret_cast<const uint4 *>(pm) : none, pb = live ? *reinterpret_cast<const uint4 *>(pm + 4) : none;
                    const uint32_t m8[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
                    uint32_t gm[8], key[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        gm[k] = min(m8[k] & 0xFFFFu, m8[k] >> 16);
                        const uint32_t gi = 8u * it_ch + k;
                        key[k] = (gm[k] << 8) | (CV ? 255u - gi : gi);
                    }
                    uint32_t best = __vimin3_u32(key[0], key[1], key[2]);
                    best = __vimin3_u32(best, key[3], key[4]);
                    best = __vimin3_u32(best, key[5], key[6]);
                    best = min(best, key[7]);
#pragma unroll
                    for (int o = 1; o < NCH; o <<= 1) best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
                    uint32_t um = 0xFFFFu;
                    if (CV) {
                        // smallest group minimum outside the winner's group and its two neighbours (those three are rescanned exactly below)
                        const int gs = 255 - (int)(best & 0xFFu);
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const int gi = 8 * it_ch + k;
                            um = (gi < gs - 1 || gi > gs + 1) ? min(um, gm[k]) : um;
                        }
#pragma unroll
                        for (int o = 1; o < NCH; o <<= 1) um = min(um, __shfl_xor_sync(0xFFFFFFFFu, um, o));
                    }
                    if (KEEP) {
                        // window sums and keys of a warp's 32 pixels come from the warp itself: the winner's group is fetched before the barrier
                        best_r = best; um_r = um;
                        const int gs = CV ? 255 - (int)(best & 0xFFu) : (int)(best & 0xFFu);
                        if (live) v_r = *reinterpret_cast<const uint4 *>(&sm.sad[it_px][8 * gs]);
                    } else if (live && it_ch == 0) { sm.ckey[it_px][0] = best; if (CV) sm.ckey[it_px][CV ? 1 : 0] = um; }
                }
                asm volatile("bar.sync 2, %0;" ::"n"(NCT) : "memory");  // every prefix entry has been read (the next row may overwrite them); keys and window sums are visible
                // ---- finishing pass: lane = pixel: winner, neighbours, sub-pixel fraction, output ----
                if (px_ok) {
                    const uint32_t best = KEEP ? best_r : sm.ckey[px][0];
                    const uint32_t mv = best >> 8;
                    const int gs = CV ? 255 - (int)(best & 0xFFu) : (int)(best & 0xFFu);
                    const uint16_t *srow = &sm.sad[px][0];
                    const uint4 v = KEEP ? v_r : *reinterpret_cast<const uint4 *>(srow + 8 * gs);
                    const uint32_t vv[8] = {v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16, v.z & 0xFFFFu, v.z >> 16, v.w & 0xFFFFu, v.w >> 16};
                    // slot k <-> d = 8g + 7 - k.  RTL: lowest disparity with SAD == mv = highest slot; OPENCV: highest disparity = lowest slot
                    uint32_t sk[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) sk[k] = (vv[k] << 3) | (CV ? (uint32_t)k : 7u - k);
                    uint32_t bk = __vimin3_u32(sk[0], sk[1], sk[2]);
                    bk = __vimin3_u32(bk, sk[3], sk[4]);
                    bk = __vimin3_u32(bk, sk[5], sk[6]);
                    bk = min(bk, sk[7]);
                    const int d1 = 8 * gs + (CV ? 7 - (int)(bk & 7u) : (int)(bk & 7u));
                    int out;
                    if (!CV) {
                        const uint16_t *gd = reinterpret_cast<const uint16_t *>(&sm.gt[b][0]);        // [d=-1 / d=D][column]
                        int L, R;
                        if (d1 == 0) { uint32_t acc = 0; for (int k = 0; k <= two_h; k++) acc += gd[px + k]; L = (int)acc; }
                        else L = srow[slot_of(d1 - 1)];
                        if (d1 == D - 1) { uint32_t acc = 0; for (int k = 0; k <= two_h; k++) acc += gd[(U_NC + 8) + px + k]; R = (int)acc; }
                        else R = srow[slot_of(d1 + 1)];
                        const int q = rtl_frac(L, R, (int)mv);
                        const int depth = d1 * 256 + q;                // bm_obuf2.v:122-154
                        if (depth <= 0) out = -1;
                        else if (a.rtl_extended) out = depth >> 4;
                        else out = (int)(int16_t)(((depth >> 4) & 0x0FFF) | ((depth & 0x8000) ? 0xF000 : 0));
                    } else {
                        // cv::StereoBM
                        const int minsad = (int)mv;
                        bool fail = false;
                        if (a.uniq > 0) {
                            // any d with |d - mind| > 1 and SAD(d) <= thresh: the groups away from the winner through their minima (um), the
                            // winner's group and its two neighbours value by value with mind-1, mind, mind+1 left out
                            const int thresh = minsad + minsad * a.uniq / 100;
                            uint32_t umin = KEEP ? um_r : sm.ckey[px][CV ? 1 : 0];
                            const int dl = d1 & 7;                     // position of the winner inside its group
#pragma unroll
                            for (int k = 0; k < 8; k++) { const int dk = 7 - k; umin = (dk < dl - 1 || dk > dl + 1) ? min(umin, vv[k]) : umin; }
                            if (gs > 0) {                              // group below: its top disparity is mind-1 iff the winner sits at the bottom of its group
                                const uint4 u = *reinterpret_cast<const uint4 *>(srow + 8 * (gs - 1));
                                const uint32_t top = (dl == 0) ? 0xFFFFu : (u.x & 0xFFFFu);           // slot 0 <-> d = 8(g-1)+7
                                umin = min(umin, min(min(top, u.x >> 16), min(__vminu2(u.y, __vminu2(u.z, u.w)) & 0xFFFFu, __vminu2(u.y, __vminu2(u.z, u.w)) >> 16)));
                            }
                            if (gs < NG - 1) {                         // group above: its bottom disparity is mind+1 iff the winner sits at the top
                                const uint4 u = *reinterpret_cast<const uint4 *>(srow + 8 * (gs + 1));
                                const uint32_t bot = (dl == 7) ? 0xFFFFu : (u.w >> 16);               // slot 7 <-> d = 8(g+1)
                                umin = min(umin, min(min(bot, u.w & 0xFFFFu), min(__vminu2(u.x, __vminu2(u.y, u.z)) & 0xFFFFu, __vminu2(u.x, __vminu2(u.y, u.z)) >> 16)));
                            }
                            fail = (int)umin <= thresh;
                        }
                        // texture: sum of |L' - cap| over the window = difference of two entries of the column prefix sums
                        const uint32_t *tx = &sm.gt[b][0];
                        const uint32_t tsum = tx[px + two_h + 1] - tx[px];
                        const bool valid = !fail && ((int)tsum >= a.tex_thr);
                        // neighbours of the winner, mirrored at the ends of the range
                        const int pp = srow[slot_of(d1 == 0 ? 1 : d1 - 1)], nn = srow[slot_of(d1 == D - 1 ? D - 2 : d1 + 1)];
                        const int den = pp + nn - 2 * minsad + abs(pp - nn);
                        // C division toward zero; exact in float: |(pp-nn)*256| < 2^24, den >= 2|pp-nn| so |frac| <= 128, and a
                        // non-integer quotient is more than 1/den > 2^-18 = half an ulp away from the next integer
                        const int frac = (den > 0) ? (int)truncf(fdiv_rn_inrange((float)((pp - nn) * 256), (float)den)) : 0;
                        out = valid ? ((d1 * 256 + frac + 15) >> 4) : -16;
                        if (a.cost && valid) a.cost[(size_t)f * a.dframe + (size_t)(yb0 + it - (wsz - 1)) * a.dpitch + ctr0 + px] = (int16_t)minsad;
                    }
                    *out_p = (int16_t)out;
                }
            }
            out_p += a.dpitch;
            // the rows of iteration it+1 were staged before the barrier in the middle of this one (the staging warp overwrites them only after the
            // next one): their loads are in flight across the back edge instead of at the head of the column-sum step
            if (WIDE) prefetch_old(b ^ 1);
        }
        if (COMPOSE) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const size_t e = (size_t)(8 * s + i) * NG + g;
                fn_blk[e] = c[i]; fn_blk[SBLK + e] = ch[i]; fn_blk[2 * SBLK + e] = ac[i];
            }
        }
    } else if (warp == CW) {
        // ======================================================================================
        // staging role: rows of iteration it+1 (newest, oldest) -> shared memory as they are (RTL: 6-bit masked, lr_din, bm_calc_sad.v:82-101)
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
            p[j] = reinterpret_cast<const uint32_t *>(isr ? gr : gl) + ((ptrdiff_t)(yb0 - h - (rt[j] ? wsz : 0) + it0) * pw + w0);
            so[j] = (uint32_t)((isr ? offsetof(SM, rrow) + (size_t)rt[j] * RB + (WIDE ? 8 : 4) * q : offsetof(SM, lrow4) + (size_t)rt[j] * 4 * U_NC + 16 * q));

        }
        uint32_t w0r[NI], w1r[NI];
        auto load = [&](int it) {
            const bool live_n = it < nsteps, live_o = live_n && it + gofs >= wsz;
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
                    const uint32_t v = __funnelshift_r(w0r[j], w1r[j], m[j]) & (CV ? 0xFFFFFFFFu : 0x3F3F3F3Fu);
                    unsigned char *dst = usm_raw + so[j] + boff * (isr ? RB : 4 * U_NC);
                    if (isr) {
                        if (WIDE) *reinterpret_cast<uint2 *>(dst) = make_uint2(fprmt(v, 0, 0x4140), fprmt(v, 0, 0x4342));      // pixels 4q .. 4q+3 as u16
                        else *reinterpret_cast<uint32_t *>(dst) = v;
                    } else {                                          // L pixels replicated for the compute threads (+ as bytes for the guard warp)
                        constexpr uint32_t REP = WIDE ? 0x00010001u : 0x01010101u;
                        *reinterpret_cast<uint4 *>(dst) = make_uint4((v & 0xFFu) * REP, ((v >> 8) & 0xFFu) * REP, ((v >> 16) & 0xFFu) * REP, (v >> 24) * REP);
                        if (!WIDE || CV) {
                            const uint32_t o1 = (uint32_t)offsetof(SM, lrow) + (so[j] - (uint32_t)offsetof(SM, lrow4)) / 4u + boff * U_NC;
                            *reinterpret_cast<uint32_t *>(usm_raw + o1) = v;
                        }
                    }
                }
        };
        load(it0); store(it0);
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        for (int it = it0; it < nsteps; it++) {
            load(it + 1);
            store(it + 1);
            asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        }
    } else {
        // ======================================================================================
        // guard role.  RTL: column sums of d = -1 and d = D (bm_calc_sad.v lanes 0 and 33), 8 columns per item, column-parallel.
        // OPENCV: the texture column sums |L' - cap| (cv::StereoBM textureThreshold) and their prefix sums over the tile's columns.
        // ======================================================================================
        uint4 cg[2];                                                  // RTL: item = lane + 32*j: segment = item >> 1, which = item & 1
        uint4 cgh[COMPOSE ? 2 : 1], cga[COMPOSE ? 2 : 1];
        cg[0] = cg[1] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int item = lane + 32 * j;
            if (sv_blk && item < 2 * U_NSEG) cg[j] = sv_blk[(size_t)U_NC * NG + (item & 1) * U_NSEG + (item >> 1)];
            if (COMPOSE) { cgh[j] = make_uint4(0x03FF03FFu, 0x03FF03FFu, 0x03FF03FFu, 0x03FF03FFu); cga[j] = make_uint4(0, 0, 0, 0); }
        }
        const uint32_t cap4 = (uint32_t)a.cap * 0x01010101u;
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        for (int it = it0; it < nsteps; it++) {
            const int b = it & 1;
            if (CV) {
                // lane = segment (20 of 32 lanes): 8 columns packed as 4 x u16x2
                const int s = min(lane, U_NSEG - 1);
                const uint2 ln = *reinterpret_cast<const uint2 *>(&sm.lrow[b][0][8 * s]);
                const uint2 lo = *reinterpret_cast<const uint2 *>(&sm.lrow[b][1][8 * s]);
                const uint32_t an0 = __vabsdiffu4(ln.x, cap4), an1 = __vabsdiffu4(ln.y, cap4);
                uint4 &cc = cg[0];
                cc.x += fprmt(an0, 0, 0x4140); cc.y += fprmt(an0, 0, 0x4342); cc.z += fprmt(an1, 0, 0x4140); cc.w += fprmt(an1, 0, 0x4342);
                if (it + gofs >= wsz) {                               // (before that the oldest row is not part of the window: its staged zeros are not pixels)
                    const uint32_t ao0 = __vabsdiffu4(lo.x, cap4), ao1 = __vabsdiffu4(lo.y, cap4);
                    cc.x -= fprmt(ao0, 0, 0x4140); cc.y -= fprmt(ao0, 0, 0x4342); cc.z -= fprmt(ao1, 0, 0x4140); cc.w -= fprmt(ao1, 0, 0x4342);
                }
                uint32_t pf[8];                                       // inclusive prefix inside the segment
                pf[0] = cc.x & 0xFFFFu; pf[1] = pf[0] + (cc.x >> 16); pf[2] = pf[1] + (cc.y & 0xFFFFu); pf[3] = pf[2] + (cc.y >> 16);
                pf[4] = pf[3] + (cc.z & 0xFFFFu); pf[5] = pf[4] + (cc.z >> 16); pf[6] = pf[5] + (cc.w & 0xFFFFu); pf[7] = pf[6] + (cc.w >> 16);
                uint32_t tot = (lane < U_NSEG) ? pf[7] : 0u, sc = tot;                    // exclusive scan of the segment totals
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, sc, o); if (lane >= o) sc += t; }
                const uint32_t base = sc - tot;
                if (lane < U_NSEG) {
                    uint32_t *tx = &sm.gt[b][1 + 8 * s];              // entry 1 + column
#pragma unroll
                    for (int i = 0; i < 8; i++) tx[i] = base + pf[i];
                    if (lane == 0) sm.gt[b][0] = 0u;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int item = lane + 32 * j;
                    if (item < 2 * U_NSEG) {
                        const int s = item >> 1, which = item & 1;
                        uint4 &cc = cg[j];
                        uint16_t *gd = reinterpret_cast<uint16_t *>(&sm.gt[b][0]) + which * (U_NC + 8);
                        if constexpr (WIDE) {
                            // R(x - D): halves 8s+8 ..; R(x + 1): halves 8s+D+9 .. (one pixel past an aligned quad); L pairs widened here
                            const uint4 *lnp = reinterpret_cast<const uint4 *>(&sm.lrow4[b][0][8 * s]), *lop = reinterpret_cast<const uint4 *>(&sm.lrow4[b][1][8 * s]);
                            const uint4 lna = lnp[0], lnb = lnp[1], loa = lop[0], lob = lop[1];       // one word (l, 0, l, 0) per column
                            const int ro = which ? 2 * (8 * s + 8) : 2 * (8 * s + D + 8);
                            uint4 rn = *reinterpret_cast<const uint4 *>(&sm.rrow[b][0][ro]), rold = *reinterpret_cast<const uint4 *>(&sm.rrow[b][1][ro]);
                            if (!which) {
                                const uint32_t tn = *reinterpret_cast<const uint32_t *>(&sm.rrow[b][0][ro + 16]), to = *reinterpret_cast<const uint32_t *>(&sm.rrow[b][1][ro + 16]);
                                rn = make_uint4(__funnelshift_r(rn.x, rn.y, 16), __funnelshift_r(rn.y, rn.z, 16), __funnelshift_r(rn.z, rn.w, 16), __funnelshift_r(rn.w, tn, 16));
                                rold = make_uint4(__funnelshift_r(rold.x, rold.y, 16), __funnelshift_r(rold.y, rold.z, 16), __funnelshift_r(rold.z, rold.w, 16), __funnelshift_r(rold.w, to, 16));
                            }
                            // columns (i, i+1) of the pair: low half of word i, high half of word i+1
                            if constexpr (COMPOSE) {
                                const uint32_t m0 = absdiff_f16_u16x2(fprmt(loa.x, loa.y, 0x7610), rold.x), m1 = absdiff_f16_u16x2(fprmt(loa.z, loa.w, 0x7610), rold.y);
                                const uint32_t m2 = absdiff_f16_u16x2(fprmt(lob.x, lob.y, 0x7610), rold.z), m3 = absdiff_f16_u16x2(fprmt(lob.z, lob.w, 0x7610), rold.w);
                                uint4 &hh = cgh[j], &aa = cga[j];
                                hh.x = satsub_u16x2(hh.x, m0); hh.y = satsub_u16x2(hh.y, m1); hh.z = satsub_u16x2(hh.z, m2); hh.w = satsub_u16x2(hh.w, m3);
                                aa.x = __vsub2(aa.x, m0); aa.y = __vsub2(aa.y, m1); aa.z = __vsub2(aa.z, m2); aa.w = __vsub2(aa.w, m3);
                            }
                            cc.x = satsub_absdiff_u16x2(cc.x, fprmt(loa.x, loa.y, 0x7610), rold.x); cc.y = satsub_absdiff_u16x2(cc.y, fprmt(loa.z, loa.w, 0x7610), rold.y);
                            cc.z = satsub_absdiff_u16x2(cc.z, fprmt(lob.x, lob.y, 0x7610), rold.z); cc.w = satsub_absdiff_u16x2(cc.w, fprmt(lob.z, lob.w, 0x7610), rold.w);
                            const uint32_t n0 = __vabsdiffu4(fprmt(lna.x, lna.y, 0x7610), rn.x), n1 = __vabsdiffu4(fprmt(lna.z, lna.w, 0x7610), rn.y);
                            const uint32_t n2 = __vabsdiffu4(fprmt(lnb.x, lnb.y, 0x7610), rn.z), n3 = __vabsdiffu4(fprmt(lnb.z, lnb.w, 0x7610), rn.w);
                            if (SAT) {
                                cc.x = __viaddmin_u16x2(cc.x, n0, 0x03FF03FFu); cc.y = __viaddmin_u16x2(cc.y, n1, 0x03FF03FFu);
                                cc.z = __viaddmin_u16x2(cc.z, n2, 0x03FF03FFu); cc.w = __viaddmin_u16x2(cc.w, n3, 0x03FF03FFu);
                            } else { cc.x += n0; cc.y += n1; cc.z += n2; cc.w += n3; }
                            if constexpr (COMPOSE) {
                                uint4 &hh = cgh[j], &aa = cga[j];
                                hh.x = __viaddmin_u16x2(hh.x, n0, 0x03FF03FFu); hh.y = __viaddmin_u16x2(hh.y, n1, 0x03FF03FFu);
                                hh.z = __viaddmin_u16x2(hh.z, n2, 0x03FF03FFu); hh.w = __viaddmin_u16x2(hh.w, n3, 0x03FF03FFu);
                                aa.x = __vadd2(aa.x, n0); aa.y = __vadd2(aa.y, n1); aa.z = __vadd2(aa.z, n2); aa.w = __vadd2(aa.w, n3);
                            }
                            *reinterpret_cast<uint4 *>(gd + 8 * s) = cc;
                        } else {
                        uint2 rv[2];
#pragma unroll
                        for (int t = 0; t < 2; t++) {                 // newest, oldest
                            if (which) rv[t] = *reinterpret_cast<const uint2 *>(&sm.rrow[b][t][8 * s + 8]);      // R(x - D)
                            else {                                    // R(x + 1): bytes 1..8 of the pair at 8s+D+8
                                const uint2 u0 = *reinterpret_cast<const uint2 *>(&sm.rrow[b][t][8 * s + D + 8]);
                                const uint32_t u2 = (8 * s + D + 16 < RLEN) ? *reinterpret_cast<const uint32_t *>(&sm.rrow[b][t][8 * s + D + 16]) : 0u;
                                rv[t] = make_uint2(__funnelshift_r(u0.x, u0.y, 8), __funnelshift_r(u0.y, u2, 8));
                            }
                        }
                        const uint2 ln = *reinterpret_cast<const uint2 *>(&sm.lrow[b][0][8 * s]);
                        const uint2 lo = *reinterpret_cast<const uint2 *>(&sm.lrow[b][1][8 * s]);
                        const uint32_t an0 = __vabsdiffu4(ln.x, rv[0].x), an1 = __vabsdiffu4(ln.y, rv[0].y);
                        const uint32_t ao0 = __vabsdiffu4(lo.x, rv[1].x), ao1 = __vabsdiffu4(lo.y, rv[1].y);
                        if (SAT) {
                            cc.x = satsub_u16x2(cc.x, fprmt(ao0, 0, 0x4140)); cc.y = satsub_u16x2(cc.y, fprmt(ao0, 0, 0x4342));
                            cc.z = satsub_u16x2(cc.z, fprmt(ao1, 0, 0x4140)); cc.w = satsub_u16x2(cc.w, fprmt(ao1, 0, 0x4342));
                            cc.x = __viaddmin_u16x2(cc.x, fprmt(an0, 0, 0x4140), 0x03FF03FFu); cc.y = __viaddmin_u16x2(cc.y, fprmt(an0, 0, 0x4342), 0x03FF03FFu);
                            cc.z = __viaddmin_u16x2(cc.z, fprmt(an1, 0, 0x4140), 0x03FF03FFu); cc.w = __viaddmin_u16x2(cc.w, fprmt(an1, 0, 0x4342), 0x03FF03FFu);
                        } else {
                            cc.x += fprmt(an0, 0, 0x4140) - fprmt(ao0, 0, 0x4140); cc.y += fprmt(an0, 0, 0x4342) - fprmt(ao0, 0, 0x4342);
                            cc.z += fprmt(an1, 0, 0x4140) - fprmt(ao1, 0, 0x4140); cc.w += fprmt(an1, 0, 0x4342) - fprmt(ao1, 0, 0x4342);
                        }
                        *reinterpret_cast<uint4 *>(gd + 8 * s) = cc;  // columns 8s .. 8s+7 as u16
                        }
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        }
        if (COMPOSE) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int item = lane + 32 * j;
                if (item < 2 * U_NSEG) {
                    const size_t e = (size_t)U_NC * NG + (item & 1) * U_NSEG + (item >> 1);
                    fn_blk[e] = cg[j]; fn_blk[SBLK + e] = cgh[j]; fn_blk[2 * SBLK + e] = cga[j];
                }
            }
        }
    }
}

// Band start states of the saturating chain: state(b + 1) = min(max(state(b) + a_b, lo_b), hi_b), state(0) = 0, per u16 lane.
// One thread = one uint4 (8 lanes) of one (frame, tile); nb - 1 sequential steps.
__global__ void __launch_bounds__(256) k_bm_chain(const uint4 *__restrict__ fn, uint4 *__restrict__ st, int nb, int ntiles, size_t blk, int n)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int tile = blockIdx.y, f = blockIdx.z;
    if (e >= blk) return;
    auto step = [](uint32_t s, uint32_t a, uint32_t lo, uint32_t hi) {
        uint32_t r = 0;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int v = (int)((s >> (16 * k)) & 0xFFFFu) + (int)(int16_t)((a >> (16 * k)) & 0xFFFFu);
            const int l = (int)((lo >> (16 * k)) & 0xFFFFu), h = (int)((hi >> (16 * k)) & 0xFFFFu);
            r |= (uint32_t)min(max(v, l), h) << (16 * k);
        }
        return r;
    };
    uint4 s = make_uint4(0, 0, 0, 0);
    for (int b = 0; b + 1 < nb; b++) {
        const uint4 *fb = fn + (((size_t)f * (nb - 1) + b) * ntiles + tile) * 3 * blk;
        const uint4 lo = fb[e], hi = fb[blk + e], a = fb[2 * blk + e];
        s = make_uint4(step(s.x, a.x, lo.x, hi.x), step(s.y, a.y, lo.y, hi.y), step(s.z, a.z, lo.z, hi.z), step(s.w, a.w, lo.w, hi.w));
        st[(((size_t)f * nb + b + 1) * ntiles + tile) * blk + e] = s;
    }
}

template <int PROFILE, bool SAT, int NG, bool WD, int MODE = 0>
static inline void fused_go(const FastArgs &a, int n, cudaStream_t s)
{
    const int smem = (int)sizeof(FusedSmem<NG, U96_FUSED_WIDE && WD, SAT, PROFILE == U96_PROFILE_OPENCV>);
    cudaFuncSetAttribute(k_bm_fused<PROFILE, SAT, NG, WD, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_bm_fused<PROFILE, SAT, NG, WD, MODE><<<dim3(a.ntx_tiles, MODE == 1 ? a.nbands - 1 : a.nbands, n), U_NSEG * NG + 64, smem, s>>>(a);
}

// y-bands of the saturating chain (64 / 128 disparities, a handful of pairs): rows per band, or 0 when the chain stays sequential
constexpr int FUSED_BAND_MAX_PAIRS = 8;
static inline int fused_band_rows() { static const int v = getenv("U96_SAT_BAND_ROWS") ? std::max(8, atoi(getenv("U96_SAT_BAND_ROWS"))) : 16; return v; }
#define FUSED_BAND_ROWS fused_band_rows()
static inline int fused_sat_bands(const BmConfig &c, int n, int *ntiles = nullptr)
{
    static const int env = getenv("U96_SAT_BANDS") ? atoi(getenv("U96_SAT_BANDS")) : 1;      // developer switch
    if (!env || c.profile != U96_PROFILE_RTL || c.uni_enable || c.wsz * 63 <= 1023 || !(c.D == 64 || c.D == 128) || n > FUSED_BAND_MAX_PAIRS) return 0;
    const int h = c.wsz >> 1, rows = c.H - 2 * h, ncen = (c.W - 2 - h) - (c.D + h) + 1, tx = U_NC - 2 * h;
    if (rows < 4 * FUSED_BAND_ROWS || ncen <= 0) return 0;
    const int tiles = (ncen + tx - 1) / tx;
    if ((long long)n * tiles * fused_occupancy(c.D / 8) > fast_sm_count()) return 0;             // the sequential sweep already fills the SMs
    if (ntiles) *ntiles = tiles;
    return (rows + FUSED_BAND_ROWS - 1) / FUSED_BAND_ROWS;
}
static inline size_t fused_sat_scratch_bytes(const BmConfig &c, int n)
{
    int tiles = 0;
    const int nb = fused_sat_bands(c, n, &tiles);
    if (!nb) return 0;
    return (size_t)n * tiles * fused_state_block(c.D / 8) * sizeof(uint4) * ((size_t)3 * (nb - 1) + nb);
}

// 64 / 128 / 256 disparities, window 9..31; RTL profile with the uniqueness filter off, or the cv::StereoBM profile
static inline bool bm_fused_supported(const BmConfig &c)
{
    if (!(c.D == 64 || c.D == 128 || c.D == 256) || c.wsz < 9 || c.wsz > 31) return false;
    if (c.profile == U96_PROFILE_RTL) return !c.uni_enable;
    return c.profile == U96_PROFILE_OPENCV && c.cap >= 1 && c.cap <= 63;
}

static inline int launch_bm_fused(const uint8_t *xl, const uint8_t *xr, int pitch, size_t frame, Img16 disp,
                                  const BmConfig &c, int n, cudaStream_t s)
{
    FastArgs a;
    a.xl = xl; a.xr = xr; a.disp = disp.p; a.pitch = pitch; a.frame = frame; a.dpitch = disp.pitch; a.dframe = disp.frame;
    const int ng = c.D / 8;
    fast_fill_args<5>(a, c, n, 1, fused_occupancy(ng) * fast_sm_count());    // same tile (160 columns) and valid rectangle as k_bm_fast<NCW=5>
    if (a.ctr_hi < a.ctr_lo || a.y_hi < a.y_lo) return 0;
    constexpr int R = U96_PROFILE_RTL, V = U96_PROFILE_OPENCV;
    if (c.profile == U96_PROFILE_OPENCV) {
        // 16-bit rows + fp16 oldest-row step when every column sum is exact as fp16: window x max |difference| (= 2 cap, xsobel clip) < 2048
        // (the reference's own configuration, main.cpp:198-212: cap 31, window 21)
        static const int cvw_env = getenv("U96_CV_WIDE") ? atoi(getenv("U96_CV_WIDE")) : 1;      // developer switch
        const bool wide = cvw_env && (c.wsz * 2 * c.cap < 2048);
        if (ng == 8)       { if (wide) fused_go<V, false, 8, true>(a, n, s);  else fused_go<V, false, 8, false>(a, n, s); }   // (same-box A/B: -3.8 % / -2.5 % / -11 % at 64 / 128 / 256)
        else if (ng == 16) { if (wide) fused_go<V, false, 16, true>(a, n, s); else fused_go<V, false, 16, false>(a, n, s); }
        else               { if (wide) fused_go<V, false, 32, true>(a, n, s); else fused_go<V, false, 32, false>(a, n, s); }
        return 1;
    }
    const bool sat = c.wsz * 63 > 1023;
    if (sat && a.TX > 144) return 0;                                             // cannot happen (window >= 17): FusedSmem::NPX
    const int nb = fused_sat_bands(c, n);
    if (nb && c.sat_scratch && c.sat_scratch_bytes >= fused_sat_scratch_bytes(c, n)) {
        // band functions (all bands at once) -> band start states (sequential over the bands, one clamp-add each) -> every band at once
        const size_t blk = fused_state_block(ng);
        a.band_h = FUSED_BAND_ROWS; a.nbands = nb; a.st_mode = 1;
        a.st_fn = reinterpret_cast<uint4 *>(c.sat_scratch);
        a.st_val = a.st_fn + (size_t)n * (nb - 1) * a.ntx_tiles * 3 * blk;
        if (ng == 8) fused_go<R, true, 8, true, 1>(a, n, s); else fused_go<R, true, 16, true, 1>(a, n, s);
        k_bm_chain<<<dim3((unsigned)((blk + 255) / 256), a.ntx_tiles, n), 256, 0, s>>>(a.st_fn, a.st_val, nb, a.ntx_tiles, blk, n);
        if (ng == 8) fused_go<R, true, 8, true, 2>(a, n, s); else fused_go<R, true, 16, true, 2>(a, n, s);
        return 3;
    }
    if (ng == 8)       { if (sat) fused_go<R, true, 8, true>(a, n, s);  else fused_go<R, false, 8, true>(a, n, s); }
    else if (ng == 16) { if (sat) fused_go<R, true, 16, true>(a, n, s); else fused_go<R, false, 16, true>(a, n, s); }
    else               { if (sat) fused_go<R, true, 32, true>(a, n, s); else fused_go<R, false, 32, true>(a, n, s); }
    return 1;
}

}  // namespace u96
