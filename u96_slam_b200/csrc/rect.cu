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
//   64*s + 2^15 = sum_rows (64*wy_row) * dp4a({left,right},{32-xf,xf}) + 2^15
// which equals the RTL's sum of four u1.10-weighted taps (rect_intp.v:337-378) by distributivity; the
// rounding ((s>>9)+1)>>1 == (s+512)>>10 (the 0xFF clamp of rect_intp.v:399-405 is unreachable: s <= 255*1024),
// so the result is byte 2 of the accumulator and four pixels are packed with three PRMTs.
// ROWS = 2 when no lane of the warp straddles a source-row step, else 3 (warp-uniform choice, no divergence).
constexpr int RECT_FPB = 8;

template <int ROWS>
__device__ __forceinline__ void remap_words(const uint32_t *__restrict__ w, uint8_t *__restrict__ dst, int spw, size_t sfw, size_t df,
                                            int nf, int mis, const uint32_t (&selw)[4], const uint32_t (&wx)[4],
                                            const uint32_t (&wy)[4][3])
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
            for (int r = 0; r < ROWS; r++) acc[k] += __dp4a(__byte_perm(lo[r], hi[r], selw[k]), wx[k], 0u) * wy[k][r];
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
        uint32_t selw[4], wx[4], wy[4][3];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t dk = (uint32_t)(xi[k] - xi[0]);
            selw[k] = dk | ((dk + 1) << 4);                           // bytes dk, dk+1 ; the upper two bytes meet zero weights
            wx[k] = (uint32_t)(32 - xf[k]) | ((uint32_t)xf[k] << 8);  // u8 weights for dp4a
            const uint32_t a = 64u * (uint32_t)(32 - yf[k]), b = 64u * (uint32_t)yf[k];
            const bool up = (yi[k] == ymin);
            wy[k][0] = up ? a : 0u; wy[k][1] = up ? b : a; wy[k][2] = up ? 0u : b;
        }
        const uint32_t *w = reinterpret_cast<const uint32_t *>(src) + (o0 >> 2);
        // the third row is only touched when some lane of the warp needs it; lanes at the bottom edge clamp it
        if (any3) {
            if (ymin + 2 < H) remap_words<3>(w, dst, sp >> 2, sf >> 2, df, nf, mis, selw, wx, wy);
            else              remap_words<2>(w, dst, sp >> 2, sf >> 2, df, nf, mis, selw, wx, wy);
        } else remap_words<2>(w, dst, sp >> 2, sf >> 2, df, nf, mis, selw, wx, wy);
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

int launch_rect_remap(const uint8_t *srcL, const uint8_t *srcR, int src_pitch, size_t src_frame,
                      Img8 dstL, Img8 dstR, const int2 *map, int W, int H, int n, cudaStream_t s)
{
    const int tx = 128;
    dim3 grid((((W + 3) / 4) * H + tx - 1) / tx, 2 * ((n + RECT_FPB - 1) / RECT_FPB));
    k_rect_remap<<<grid, tx, 0, s>>>(srcL, srcR, src_pitch, src_frame, dstL.p, dstR.p, dstL.pitch, dstL.frame, map, W, H, n);
    return 1;
}

}  // namespace u96
