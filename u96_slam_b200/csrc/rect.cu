// rect.cu -- stereo rectification for sm_100a.
//
//   k_rect_build_map : the fixed-point inverse map of StereoBM/src/fpga.c:303-366
//                      (== dvp/rtl/rect_rmp.v:366-585), run once per parameter set.
//   k_rect_remap     : 5-bit-fraction bilinear gather of dvp/rtl/rect_intp.v:288-412.
//
// The map is frame-invariant, so it is materialised once (8 B/px/camera, L2-resident)
// and the per-frame kernel is a pure gather: 1 B/px read + 1 B/px written to HBM.
// The FPGA's run-length command stream (fpga.c:368-605) is a line-buffer scheduling
// artefact and has no GPU counterpart: every destination pixel is written exactly once.
#include <cuda.h>      // CUtensorMap types; cuTensorMapEncodeTiled is fetched with cudaGetDriverEntryPoint (no -lcuda)

#include <algorithm>
#include <climits>
#include <cstdlib>

#include "common.cuh"

namespace u96 {

struct RectConst {
    long long f[2][2], rot[2][3][3];
    long long c[2], f2inv[2], c2_f2[2];
};

__global__ void __launch_bounds__(256) k_rect_build_map(RectConst k, int2 *__restrict__ map, int W, int H, int wrap16)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int lr = blockIdx.z;
    if (x >= W) return;
    // (u10.0)*(u-8.32) -> (u1.24), minus (u0.24)                         fpga.c:317-323
    const long long xd = (((long long)x * k.f2inv[0]) >> 8) - k.c2_f2[0];
    const long long yd = (((long long)y * k.f2inv[1]) >> 8) - k.c2_f2[1];
    // each product truncated separately, then summed                     fpga.c:325-340
    const long long lx = ((k.rot[lr][0][0] * xd) >> 24) + ((k.rot[lr][1][0] * yd) >> 24) + k.rot[lr][2][0];
    const long long ly = ((k.rot[lr][0][1] * xd) >> 24) + ((k.rot[lr][1][1] * yd) >> 24) + k.rot[lr][2][1];
    const long long lw = ((k.rot[lr][0][2] * xd) >> 24) + ((k.rot[lr][1][2] * yd) >> 24) + k.rot[lr][2][2];
    // (1ull << 48) / lw is an unsigned 64-bit division in the reference   fpga.c:343
    const long long winv = (long long)((1ull << 48) / (unsigned long long)lw);
    const long long x2 = (lx * winv) >> 24;
    const long long y2 = (ly * winv) >> 24;
    const long long xf = ((x2 * k.f[lr][0]) >> 34) + (k.c[0] << 6);
    const long long yf = ((y2 * k.f[lr][1]) >> 34) + (k.c[1] << 6);
    long long xs = (xf + 1) >> 1, ys = (yf + 1) >> 1;
    if (wrap16) { xs = (short)xs; ys = (short)ys; }     // MAT2S stores shorts (fpga.c:361-362)
    else {                                               // RTL-extended: saturate far outside
        xs = max(-64ll, min(xs, (long long)(W + 1) * 32));
        ys = max(-64ll, min(ys, (long long)(H + 1) * 32));
    }
    map[((size_t)lr * H + y) * W + x] = make_int2((int)xs, (int)ys);
}

int launch_rect_build_map(const RectMapParams &rp, int2 *map, cudaStream_t s)
{
    RectConst k;
    for (int cam = 0; cam < 2; cam++) {
        for (int i = 0; i < 2; i++) k.f[cam][i] = rp.p.f[cam][i];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) k.rot[cam][i][j] = rp.p.rot[cam][i][j];
    }
    for (int i = 0; i < 2; i++) { k.c[i] = rp.p.c[i]; k.f2inv[i] = rp.p.f2inv[i]; k.c2_f2[i] = rp.p.c2_f2[i]; }
    dim3 grid((rp.W + 255) / 256, rp.H, 2);
    k_rect_build_map<<<grid, 256, 0, s>>>(k, map, rp.W, rp.H, rp.wrap16);
    return 1;
}

// One thread = 4 consecutive destination pixels of one row of one camera; it keeps the 4 map entries and
// bilinear weights in registers and loops over FPB frames of the batch, so the map is read once per FPB frames.
// Taps outside the source read 0.
//
// Word path (taken when the source columns of the 4 pixels span <= 6 bytes and their source rows span <= 2 rows,
// i.e. almost everywhere for a rectifying rotation): the 2 or 3 source rows are fetched as 3 aligned 32-bit
// words each and funnel-shifted into 8 consecutive bytes; the taps are byte-permuted out and interpolated as
//   64*s + 2^15 = sum_rows dp2a({64*wy_row*(32-xf), 64*wy_row*xf}, {left, right}) + 2^15
// i.e. the RTL's sum of four u1.10-weighted taps (rect_intp.v:337-378) scaled by 64, one IDP.2A (16-bit weights x 8-bit
// taps) per source row; the rounding ((s>>9)+1)>>1 == (s+512)>>10 (the 0xFF clamp of rect_intp.v:399-405 is
// unreachable: s <= 255*1024), so the result is byte 2 of the accumulator and four pixels are packed with three PRMTs.
// The only weight that does not fit 16 bits is 64*1024 (xf = yf = 0, all other weights zero); it is stored as 65535,
// which still yields (tap*65535 + 2^15) >> 16 == tap.
// ROWS = 2 when no lane of the warp straddles a source-row step, else 3 (warp-uniform choice, no divergence).
constexpr int RECT_FPB = 8;

// packed 16-bit weight pairs {left | right << 16} of one pixel for the up to three source rows of its group, scaled by 64
__device__ __forceinline__ void row_weights(uint32_t (&wr)[3], int xf, int yf, bool up)
{
    const uint32_t a = 64u * (uint32_t)(32 - yf), b = 64u * (uint32_t)yf;
    const uint32_t w0 = up ? a : 0u, w1 = up ? b : a, w2 = up ? 0u : b;
    const uint32_t xl = (uint32_t)(32 - xf), xr = (uint32_t)xf;
    wr[0] = min(w0 * xl, 65535u) | ((w0 * xr) << 16);
    wr[1] = min(w1 * xl, 65535u) | ((w1 * xr) << 16);
    wr[2] = min(w2 * xl, 65535u) | ((w2 * xr) << 16);
}

template <int ROWS>
__device__ __forceinline__ void remap_words(const uint32_t *__restrict__ w, uint8_t *__restrict__ dst, int spw, size_t sfw, size_t df,
                                            int nf, int mis, const uint32_t (&selw)[4], const uint32_t (&wr)[4][3])
{
#pragma unroll 4
    for (int f = 0; f < nf; f++) {
        uint32_t lo[ROWS], hi[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            const uint32_t a0 = __ldg(w + r * spw), a1 = __ldg(w + r * spw + 1), a2 = __ldg(w + r * spw + 2);
            lo[r] = __funnelshift_r(a0, a1, mis);
            hi[r] = __funnelshift_r(a1, a2, mis);
        }
        uint32_t acc[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            acc[k] = 1u << 15;
#pragma unroll
            for (int r = 0; r < ROWS; r++) acc[k] = __dp2a_lo(wr[k][r], __byte_perm(lo[r], hi[r], selw[k]), acc[k]);
        }
        const uint32_t p01 = __byte_perm(acc[0], acc[1], 0x0062), p23 = __byte_perm(acc[2], acc[3], 0x0062);
        *reinterpret_cast<uint32_t *>(dst) = __byte_perm(p01, p23, 0x5410);      // pitch is a multiple of 128: always in-row
        w += sfw;
        dst += df;
    }
}

__global__ void __launch_bounds__(128, 8) k_rect_remap(const uint8_t *__restrict__ srcL, const uint8_t *__restrict__ srcR,
                                                    int sp, size_t sf, uint8_t *__restrict__ dL, uint8_t *__restrict__ dR,
                                                    int dp, size_t df, const int2 *__restrict__ map, int W, int H, int n)
{
    const int w4 = (W + 3) >> 2;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;       // flattened (row, 4-pixel group)
    const bool live = item < w4 * H;
    const int y = live ? item / w4 : 0;
    const int x4 = live ? (item - y * w4) * 4 : 0;
    const int lr = blockIdx.y & 1;
    const int f0 = (blockIdx.y >> 1) * RECT_FPB;
    const int2 *m = map + ((size_t)lr * H + y) * W + x4;
    int xi[4], yi[4], xf[4], yf[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int2 e = (live && x4 + k < W) ? m[k] : make_int2(-64, -64);
        xi[k] = e.x >> 5; xf[k] = e.x & 31; yi[k] = e.y >> 5; yf[k] = e.y & 31;
    }
    const uint8_t *src = (lr ? srcR : srcL) + (size_t)f0 * sf;
    uint8_t *dst = (lr ? dR : dL) + (size_t)f0 * df + (size_t)y * dp + x4;
    const int nf = live ? min(RECT_FPB, n - f0) : 0;

    int ymin = yi[0], ymax = yi[0];
#pragma unroll
    for (int k = 1; k < 4; k++) { ymin = min(ymin, yi[k]); ymax = max(ymax, yi[k]); }
    bool words = live && (ymin >= 0) && (ymax + 1 < H) && (ymax - ymin <= 1) && (xi[0] >= 0) && (xi[0] + 12 <= sp);
#pragma unroll
    for (int k = 1; k < 4; k++) words = words && (xi[k] >= xi[0]) && (xi[k] - xi[0] <= 6);
#pragma unroll
    for (int k = 0; k < 4; k++) words = words && (xi[k] + 1 < W);
    const bool three = words && (ymax != ymin) && (ymin + 2 < H);
    words = words && (ymax == ymin || three);
    const unsigned any3 = __any_sync(0xFFFFFFFFu, three);

    if (words) {
        const int o0 = ymin * sp + xi[0];
        const int mis = (o0 & 3) * 8;
        uint32_t selw[4], wr[4][3];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t dk = (uint32_t)(xi[k] - xi[0]);
            selw[k] = dk | ((dk + 1) << 4);                           // bytes dk, dk+1 ; dp2a.lo ignores the upper two bytes
            row_weights(wr[k], xf[k], yf[k], yi[k] == ymin);
        }
        const uint32_t *w = reinterpret_cast<const uint32_t *>(src) + (o0 >> 2);
        // the third row is only touched when some lane of the warp needs it; lanes at the bottom edge clamp it
        if (any3) {
            if (ymin + 2 < H) remap_words<3>(w, dst, sp >> 2, sf >> 2, df, nf, mis, selw, wr);
            else              remap_words<2>(w, dst, sp >> 2, sf >> 2, df, nf, mis, selw, wr);
        } else remap_words<2>(w, dst, sp >> 2, sf >> 2, df, nf, mis, selw, wr);
        return;
    }

    // generic path (image borders, exotic maps): four independent byte gathers per pixel
    int off[4];          // byte offset of the upper-left tap (may be outside)
    uint32_t w01[4], w23[4];   // packed weights: w00 | w01<<16, w10 | w11<<16   (u1.10 each)
    uint32_t ok[4];      // validity bits of the four taps
#pragma unroll
    for (int k = 0; k < 4; k++) {
        w01[k] = (uint32_t)((32 - xf[k]) * (32 - yf[k])) | ((uint32_t)(xf[k] * (32 - yf[k])) << 16);
        w23[k] = (uint32_t)((32 - xf[k]) * yf[k]) | ((uint32_t)(xf[k] * yf[k]) << 16);
        const bool x0 = (xi[k] >= 0 && xi[k] < W), x1 = (xi[k] + 1 >= 0 && xi[k] + 1 < W);
        const bool y0 = (yi[k] >= 0 && yi[k] < H), y1 = (yi[k] + 1 >= 0 && yi[k] + 1 < H);
        ok[k] = (x0 && y0 ? 1u : 0u) | (x1 && y0 ? 2u : 0u) | (x0 && y1 ? 4u : 0u) | (x1 && y1 ? 8u : 0u);
        off[k] = yi[k] * sp + xi[k];
    }
    for (int f = 0; f < nf; f++) {
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint8_t *t = src + off[k];
            const uint32_t ul = (ok[k] & 1u) ? __ldg(t) : 0u;
            const uint32_t ur = (ok[k] & 2u) ? __ldg(t + 1) : 0u;
            const uint32_t dl = (ok[k] & 4u) ? __ldg(t + sp) : 0u;
            const uint32_t dr = (ok[k] & 8u) ? __ldg(t + sp + 1) : 0u;
            // u8 * u1.10 summed -> u8.10 ; ((s>>9)+1)>>1 with clamp      rect_intp.v:347-405
            const uint32_t s = ul * (w01[k] & 0xFFFFu) + ur * (w01[k] >> 16) + dl * (w23[k] & 0xFFFFu) + dr * (w23[k] >> 16);
            out |= min(255u, ((s >> 9) + 1u) >> 1) << (8 * k);
        }
        *reinterpret_cast<uint32_t *>(dst) = out;
        src += sf;
        dst += df;
    }
}

// ---------------------------------------------------------------------------------------------
// TMA path.  The map is frame-invariant, so the SOURCE bounding box of every 128x16 destination tile
// is known once the map exists (k_rect_tile_bbox).  Per frame a tile then needs one fixed-size box of
// the source image: a producer warp streams those boxes into a 4-stage shared-memory ring with
// cp.async.bulk.tensor (3-D tensor map x, y, frame; the hardware zero-fills everything outside the
// image = the "taps outside the source read 0" rule), 8 consumer warps interpolate out of shared
// memory with the same word/dp4a arithmetic as above and write 128-byte row segments.  HBM sees each
// source byte once (neighbouring tiles' halo overlap is absorbed by L2) and each destination byte once.
constexpr int RT_TW = 128, RT_TH = 16, RT_CONS = 256, RT_THREADS = RT_CONS + 32, RT_G = RT_TH / 8;

__global__ void __launch_bounds__(256) k_rect_tile_bbox(const int2 *__restrict__ map, int4 *__restrict__ tiles, int W, int H, int tiles_x, int ntiles)
{
    __shared__ int s_mm[4];
    __shared__ int s_row[RT_TH][2];      // per destination row of the tile: min / max source row
    const int tile = blockIdx.x, cam = blockIdx.y;
    const int tx0 = (tile % tiles_x) * RT_TW, ty0 = (tile / tiles_x) * RT_TH;
    if (threadIdx.x == 0) { s_mm[0] = INT_MAX; s_mm[1] = INT_MAX; s_mm[2] = INT_MIN; s_mm[3] = INT_MIN; }
    if (threadIdx.x < RT_TH) { s_row[threadIdx.x][0] = INT_MAX; s_row[threadIdx.x][1] = INT_MIN; }
    __syncthreads();
    int x0 = INT_MAX, y0 = INT_MAX, x1 = INT_MIN, y1 = INT_MIN;
    for (int i = threadIdx.x; i < RT_TW * RT_TH; i += blockDim.x) {
        const int x = tx0 + (i % RT_TW), y = ty0 + (i / RT_TW);
        if (x < W && y < H) {
            const int2 e = map[((size_t)cam * H + y) * W + x];
            const int xi = e.x >> 5, yi = e.y >> 5;
            x0 = min(x0, xi); x1 = max(x1, xi + 1); y0 = min(y0, yi); y1 = max(y1, yi + 1);
            atomicMin(&s_row[i / RT_TW][0], yi); atomicMax(&s_row[i / RT_TW][1], yi);
        }
    }
    atomicMin(&s_mm[0], x0); atomicMin(&s_mm[1], y0); atomicMax(&s_mm[2], x1); atomicMax(&s_mm[3], y1);
    __syncthreads();
    // the box origin is rounded down to a 16-byte boundary: TMA needs 16-byte aligned global row starts
    if (threadIdx.x == 0) {
        tiles[(size_t)cam * ntiles + tile] = make_int4(s_mm[0] & ~15, s_mm[1], s_mm[2], s_mm[3]);
        int span = 0;
        for (int r = 0; r < RT_TH; r++) if (s_row[r][1] >= s_row[r][0]) span = max(span, s_row[r][1] - s_row[r][0]);
        atomicMax(&tiles[2 * (size_t)ntiles].x, span);              // plan-wide: steepest destination row (extra entry behind the boxes)
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

// words of one 4-pixel group out of the shared-memory box (same arithmetic as remap_words)
template <int ROWS>
__device__ __forceinline__ uint32_t remap_group_smem(const uint32_t *w, int bww, uint32_t mis, const uint32_t (&selw)[4],
                                                     const uint32_t (&wr)[4][3])
{
    uint32_t lo[ROWS], hi[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        const uint32_t a0 = w[r * bww], a1 = w[r * bww + 1], a2 = w[r * bww + 2];
        lo[r] = __funnelshift_r(a0, a1, mis);
        hi[r] = __funnelshift_r(a1, a2, mis);
    }
    uint32_t acc[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        acc[k] = 1u << 15;
#pragma unroll
            for (int r = 0; r < ROWS; r++) acc[k] = __dp2a_lo(wr[k][r], __byte_perm(lo[r], hi[r], selw[k]), acc[k]);
    }
    const uint32_t p01 = __byte_perm(acc[0], acc[1], 0x0062), p23 = __byte_perm(acc[2], acc[3], 0x0062);
    return __byte_perm(p01, p23, 0x5410);
}

template <int RT_STAGES>
__global__ void __launch_bounds__(RT_THREADS, 3) k_rect_remap_tma(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR,
                                                               uint8_t *__restrict__ dL, uint8_t *__restrict__ dR, int dp, size_t df,
                                                               const int2 *__restrict__ map, const int4 *__restrict__ tiles,
                                                               int W, int H, int n, int BW, int BH, int stage_bytes, int tiles_x, int ntiles, int fpb,
                                                               int rt_per_lane_base)
{
    extern __shared__ __align__(128) uint8_t rt_smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(rt_smem + (size_t)RT_STAGES * stage_bytes);
    uint64_t *empty = full + RT_STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, cam = blockIdx.y;
    const int f0 = blockIdx.z * fpb, nf = min(fpb, n - f0);
    const int4 tb = tiles[(size_t)cam * ntiles + tile];
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < RT_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], RT_CONS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == RT_CONS / 32) {
        // ---- producer: one lane streams the per-frame source boxes through the ring ----
        if (lane == 0) {
            const CUtensorMap *tm = cam ? &tmR : &tmL;
            for (int f = 0; f < nf; f++) {
                const int s = f % RT_STAGES, k = f / RT_STAGES;
                if (k > 0) mbar_wait(&empty[s], (uint32_t)(k - 1) & 1u);
                mbar_expect_tx(&full[s], (uint32_t)(BW * BH));
                tma_load_3d(rt_smem + (size_t)s * stage_bytes, tm, tb.x, tb.y, f0 + f, &full[s]);
            }
        }
        return;
    }

    // ---- consumers: thread = (row r, 4-pixel group) x RT_G rows ----
    const int tx0 = (tile % tiles_x) * RT_TW, ty0 = (tile / tiles_x) * RT_TH;
    const int x4 = tx0 + lane * 4;
    const int bww = BW >> 2;
    int mode[RT_G];                    // 0 dead, 2/3 word path with that many rows, 1 generic
    int o0[RT_G];
    uint32_t mis[RT_G], selw[RT_G][4], wr[RT_G][4][3];
    size_t doff[RT_G];
#pragma unroll
    for (int g = 0; g < RT_G; g++) {
        const int y = ty0 + warp + 8 * g;
        const bool live = (x4 < W) && (y < H);
        doff[g] = (size_t)y * dp + x4;
        int xi[4], yi[4], xf[4], yf[4];
        const int2 *m = map + ((size_t)cam * H + (live ? y : 0)) * W + (live ? x4 : 0);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // pixels right of the image inside the last 4-group repeat the group's first entry (never stored past the pitch)
            const int2 e = live ? m[(x4 + k < W) ? k : 0] : make_int2(0, 0);
            xi[k] = (e.x >> 5) - tb.x; xf[k] = e.x & 31; yi[k] = (e.y >> 5) - tb.y; yf[k] = e.y & 31;
        }
        int ymin = yi[0], ymax = yi[0];
#pragma unroll
        for (int k = 1; k < 4; k++) { ymin = min(ymin, yi[k]); ymax = max(ymax, yi[k]); }
        bool words = live && (ymax - ymin <= 1);
#pragma unroll
        for (int k = 1; k < 4; k++) words = words && (xi[k] >= xi[0]) && (xi[k] - xi[0] <= 6);
        // Row base.  A warp covers 128 destination pixels of one row; under a rectifying rotation their source rows step
        // once or twice along the way.  If every lane started at its OWN first source row, lanes left and right of a step
        // would read different box rows in the same LDS -- a bank conflict whenever the box pitch is not a multiple of 128 B
        // (61 % of the shared wavefronts of the r01 kernel).  With one base row for the whole warp (REDUX min) every LDS
        // reads consecutive words of ONE row: conflict-free; a pixel whose taps start one row lower just carries a zero
        // weight for the base row.  Maps too steep for that (more than two source rows under a warp) keep per-lane bases.
        const int wmin = __reduce_min_sync(0xFFFFFFFFu, words ? ymin : INT_MAX);
        const int wmax = __reduce_max_sync(0xFFFFFFFFu, words ? ymax : INT_MIN);
        const bool uni = (wmax - wmin <= 1) && !rt_per_lane_base;
        const int base = uni ? wmin : ymin;
        const bool three = words && (uni ? (wmax != wmin) : (ymax != ymin));
        const unsigned any3 = __any_sync(0xFFFFFFFFu, three);
        mode[g] = !live ? 0 : words ? (any3 ? 3 : 2) : 1;
        o0[g] = base * BW + xi[0];
        mis[g] = (uint32_t)(o0[g] & 3) * 8u;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t dk = (uint32_t)(xi[k] - xi[0]) & 7u;
            selw[g][k] = dk | ((dk + 1) << 4);
            row_weights(wr[g][k], xf[k], yf[k], yi[k] == base);
        }
    }
    uint8_t *dbase = (cam ? dR : dL) + (size_t)f0 * df;
    static_assert(RT_G == 2, "the fast loop below is written for two rows per thread");
    if (__all_sync(0xFFFFFFFFu, mode[0] >= 2 && mode[1] >= 2)) {
        // whole warp on the word path (every full tile of a sane map): branch-free frame loop.  The row count is warp-uniform;
        // a third row with zero weights is harmless (the stage keeps a spare row behind the box).
        const bool r3 = (mode[0] == 3) || (mode[1] == 3);
        uint8_t *d0 = dbase + doff[0], *d1 = dbase + doff[1];
        const uint32_t w0 = smem_u32(rt_smem) + (uint32_t)(o0[0] & ~3), w1 = smem_u32(rt_smem) + (uint32_t)(o0[1] & ~3);
        const uint32_t rowb = (uint32_t)BW;
        auto lds = [](uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; };
        // one PRMT gathers the tap pairs of TWO pixels (bytes x0, x0+1, x1, x1+1): dp2a.lo takes the first pair, dp2a.hi the second
        const uint32_t sel2[2][2] = {{selw[0][0] | (selw[0][1] << 8), selw[0][2] | (selw[0][3] << 8)},
                                     {selw[1][0] | (selw[1][1] << 8), selw[1][2] | (selw[1][3] << 8)}};
        auto group = [&](uint32_t a, uint32_t mi, const uint32_t (&sel)[2], const uint32_t (&wg)[4][3], bool three_rows) {
            uint32_t acc[4] = {1u << 15, 1u << 15, 1u << 15, 1u << 15};
#pragma unroll
            for (int r = 0; r < 3; r++) {
                if (r == 2 && !three_rows) break;
                const uint32_t a0 = lds(a + r * rowb), a1 = lds(a + r * rowb + 4), a2 = lds(a + r * rowb + 8);
                const uint32_t lo = __funnelshift_r(a0, a1, mi), hi = __funnelshift_r(a1, a2, mi);
                const uint32_t p01 = __byte_perm(lo, hi, sel[0]), p23 = __byte_perm(lo, hi, sel[1]);
                acc[0] = __dp2a_lo(wg[0][r], p01, acc[0]); acc[1] = __dp2a_hi(wg[1][r], p01, acc[1]);
                acc[2] = __dp2a_lo(wg[2][r], p23, acc[2]); acc[3] = __dp2a_hi(wg[3][r], p23, acc[3]);
            }
            const uint32_t p01 = __byte_perm(acc[0], acc[1], 0x0062), p23 = __byte_perm(acc[2], acc[3], 0x0062);
            return __byte_perm(p01, p23, 0x5410);
        };
        // kept out of the compiler's sight: it otherwise re-derives the leader test and both row pointers from %tid every frame
        uint32_t leader = (lane == 0) ? 1u : 0u;
        asm volatile("" : "+r"(leader));
        asm volatile("" : "+l"(d0));
        asm volatile("" : "+l"(d1));
        // one trip = one turn of the ring: the stage index, its barriers and its shared-memory offset are compile-time constants
        uint32_t ph = 0;
        for (int f = 0; f < nf; f += RT_STAGES) {
#pragma unroll
            for (int s = 0; s < RT_STAGES; s++) {
                if (f + s < nf) {
                    mbar_wait(&full[s], ph);
                    const uint32_t soff = (uint32_t)s * (uint32_t)stage_bytes;
                    uint32_t out0, out1;
                    if (r3) { out0 = group(w0 + soff, mis[0], sel2[0], wr[0], true);  out1 = group(w1 + soff, mis[1], sel2[1], wr[1], true); }
                    else    { out0 = group(w0 + soff, mis[0], sel2[0], wr[0], false); out1 = group(w1 + soff, mis[1], sel2[1], wr[1], false); }
                    __syncwarp();
                    if (leader) mbar_arrive(&empty[s]);
                    asm volatile("st.global.u32 [%0], %1;" ::"l"(d0), "r"(out0) : "memory");      // (the laundered pointers are generic to the compiler)
                    asm volatile("st.global.u32 [%0], %1;" ::"l"(d1), "r"(out1) : "memory");
                    d0 += df; d1 += df;
                }
            }
            ph ^= 1u;
        }
        return;
    }
    for (int f = 0; f < nf; f++) {
        const int s = f % RT_STAGES, kk = f / RT_STAGES;
        mbar_wait(&full[s], (uint32_t)kk & 1u);
        const uint8_t *sb = rt_smem + (size_t)s * stage_bytes;
        uint32_t out[RT_G];
#pragma unroll
        for (int g = 0; g < RT_G; g++) {
            out[g] = 0;
            if (mode[g] >= 2) {
                const uint32_t *w = reinterpret_cast<const uint32_t *>(sb) + (o0[g] >> 2);
                out[g] = (mode[g] == 3) ? remap_group_smem<3>(w, bww, mis[g], selw[g], wr[g])
                                        : remap_group_smem<2>(w, bww, mis[g], selw[g], wr[g]);
            } else if (mode[g] == 1) {
                // generic path (exotic maps): byte gathers; every tap lies inside the box, outside-image taps are TMA zero fill
                const int y = ty0 + warp + 8 * g;
                const int2 *m = map + ((size_t)cam * H + y) * W + x4;
                for (int k = 0; k < 4; k++) {
                    const int2 e = m[(x4 + k < W) ? k : 0];
                    const int xi = (e.x >> 5) - tb.x, yi = (e.y >> 5) - tb.y, xf = e.x & 31, yf = e.y & 31;
                    const uint8_t *t = sb + yi * BW + xi;
                    const uint32_t sum = t[0] * (uint32_t)((32 - xf) * (32 - yf)) + t[1] * (uint32_t)(xf * (32 - yf)) +
                                         t[BW] * (uint32_t)((32 - xf) * yf) + t[BW + 1] * (uint32_t)(xf * yf);
                    out[g] |= min(255u, ((sum >> 9) + 1u) >> 1) << (8 * k);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);           // this warp is done with the stage
        uint8_t *d = dbase + (size_t)f * df;
#pragma unroll
        for (int g = 0; g < RT_G; g++)
            if (mode[g]) *reinterpret_cast<uint32_t *>(d + doff[g]) = out[g];
    }
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled tmap_encoder()
{
    static PFN_tmapEncodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }
    return fn;
}

// Source bounding boxes of the destination tiles; decides whether the TMA path applies to this parameter set.
int rect_plan_build(RectPlan &pl, const int2 *map, int W, int H, cudaStream_t s)
{
    pl.tma = false;
    pl.tiles_x = (W + RT_TW - 1) / RT_TW; pl.tiles_y = (H + RT_TH - 1) / RT_TH;
    const int nt = pl.tiles_x * pl.tiles_y;
    if (cudaMalloc(&pl.d_tiles, sizeof(int4) * (2 * nt + 1)) != cudaSuccess) { pl.d_tiles = nullptr; return 0; }
    cudaMemsetAsync(pl.d_tiles + 2 * nt, 0, sizeof(int4), s);
    k_rect_tile_bbox<<<dim3(nt, 2), 256, 0, s>>>(map, pl.d_tiles, W, H, pl.tiles_x, nt);
    int4 *ht = new int4[2 * nt + 1];
    if (cudaMemcpyAsync(ht, pl.d_tiles, sizeof(int4) * (2 * nt + 1), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) { delete[] ht; return 1; }
    const int row_span = ht[2 * nt].x;
    int bw = 0, bh = 0;
    for (int i = 0; i < 2 * nt; i++) { bw = std::max(bw, ht[i].z - ht[i].x + 1); bh = std::max(bh, ht[i].w - ht[i].y + 1); }
    delete[] ht;
    // Box pitch in shared memory = box width.  A destination row (one warp) whose source rows stay within two box rows reads
    // them through one warp-uniform base row -- conflict-free at any pitch.  Steeper maps (keystone: the shipped set spans up to
    // 7 source rows under 128 pixels at the top and bottom edge) make the lanes of one LDS read different box rows; those only
    // stay on distinct banks when the pitch is a multiple of 128 B, so the box is widened to that (more L2->SM bytes, same HBM).
    const bool steep = row_span > 1;
    pl.BH = bh;
    pl.BW = align_up(bw, 16);
    if ((steep || getenv("U96_RECT_BW128")) && !getenv("U96_RECT_BW16")) {
        const int wide = align_up(bw, 128);
        if (wide <= 256 && align_up(wide * (bh + 1) + 16, 128) <= 16384) pl.BW = wide;
    }
    pl.stage_bytes = align_up(pl.BW * (pl.BH + 1) + 16, 128);     // one spare row + tail for the word path's over-read
    pl.tma = (pl.BW <= 256 && pl.BH <= 255 && pl.stage_bytes <= 16384 && tmap_encoder() != nullptr && !getenv("U96_RECT_LEGACY"));
    return 1;
}

void rect_plan_free(RectPlan &pl) { cudaFree(pl.d_tiles); pl.d_tiles = nullptr; pl.tma = false; }

static bool encode_src_map(CUtensorMap *tm, const uint8_t *src, int pitch, size_t frame, int W, int H, int n, int BW, int BH)
{
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame};
    const cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, 1u};
    const cuuint32_t es[3] = {1u, 1u, 1u};
    return tmap_encoder()(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(src), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int launch_rect_remap(const uint8_t *srcL, const uint8_t *srcR, int src_pitch, size_t src_frame,
                      Img8 dstL, Img8 dstR, const int2 *map, const RectPlan &pl, int W, int H, int n, cudaStream_t s)
{
    if (pl.tma && (src_pitch % 16 == 0) && (src_frame % 16 == 0) && (((uintptr_t)srcL | (uintptr_t)srcR) % 16 == 0)) {
        CUtensorMap tmL, tmR;
        if (encode_src_map(&tmL, srcL, src_pitch, src_frame, W, H, n, pl.BW, pl.BH) &&
            encode_src_map(&tmR, srcR, src_pitch, src_frame, W, H, n, pl.BW, pl.BH)) {
            const int nt = pl.tiles_x * pl.tiles_y;
            static const int fpb_env = getenv("U96_RECT_FPB") ? atoi(getenv("U96_RECT_FPB")) : 0;
            int fpb = fpb_env;
            if (fpb <= 0) {
                // every CTA pays a set-up worth ~6 frames (map entries -> weights); the grid drains through `slots` resident CTAs and
                // ends with about one CTA duration of partially filled machine: time ~ work / slots + one CTA  (fitted on C2 and C4:
                // a ceil()-per-wave model picks far too few, too long CTAs)
                int dev = 0, sms = 148;
                cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                const double slots = 3.0 * sms, per = 2.0 * nt;
                double best = -1.0;
                for (int nz = 1; nz <= 64 && nz <= n; nz++) {
                    const int len = (n + nz - 1) / nz;
                    if (len < 8 && nz > 1) break;
                    const double cost = per * (n + 6.0 * ((n + len - 1) / len)) / slots + (len + 6.0);
                    if (best < 0 || cost < best) { best = cost; fpb = len; }
                }
            }
            static const int per_lane = getenv("U96_RECT_PER_LANE_BASE") ? 1 : 0;      // developer switch: the r01 behaviour
            static const int st_env = getenv("U96_RECT_STAGES") ? atoi(getenv("U96_RECT_STAGES")) : 0;
            // small boxes (gentle maps) want a deeper ring to cover the TMA latency; the wide boxes of steep maps carry enough bytes per stage
            const int stages = st_env > 0 ? st_env : (pl.stage_bytes <= 4096 ? 8 : 4);
            dim3 grid(nt, 2, (n + fpb - 1) / fpb);
            auto go = [&](auto kern, int st) {
                const int smem = st * pl.stage_bytes + 2 * st * (int)sizeof(uint64_t);
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, st * 16384 + 256);
                kern<<<grid, RT_THREADS, smem, s>>>(tmL, tmR, dstL.p, dstR.p, dstL.pitch, dstL.frame, map, pl.d_tiles, W, H, n,
                                                    pl.BW, pl.BH, pl.stage_bytes, pl.tiles_x, nt, fpb, per_lane);
            };
            if (stages >= 12) go(k_rect_remap_tma<12>, 12);
            else if (stages >= 8) go(k_rect_remap_tma<8>, 8);
            else if (stages >= 6) go(k_rect_remap_tma<6>, 6);
            else go(k_rect_remap_tma<4>, 4);
            return 1;
        }
    }
    const int tx = 128;
    dim3 grid((((W + 3) / 4) * H + tx - 1) / tx, 2 * ((n + RECT_FPB - 1) / RECT_FPB));
    k_rect_remap<<<grid, tx, 0, s>>>(srcL, srcR, src_pitch, src_frame, dstL.p, dstR.p, dstL.pitch, dstL.frame, map, W, H, n);
    return 1;
}

}  // namespace u96
