// xsobel.cu -- x-Sobel prefilter with clipping, sm_100a.
//
//   RTL profile    : dvp/rtl/xsbl2.v:185-198 (limit), :682-698 (horizontal difference),
//                    :826-857 (vertical 1-2-1), :861-874 (column borders).
//                    out = clip(s,-32,31)+32 ; cols 0,W-1 = 32 ; rows 0,H-1 = 0 (never written,
//                    firmware memsets the bank to 0: StereoBM/src/fpga.c:113-114).
//   OPENCV profile : cv::StereoBM prefilterXSobel as called from slam/src/core/main.cpp:197-217
//                    out = clip(s,-cap,cap)+cap ; cols = cap ; rows reflect-101 ; odd H: last row = cap.
//
// HBM-bound: 1 B/px read + 1 B/px written.  One thread produces 16 output pixels with one
// 16-byte store; the three input rows are fetched as 16-byte vectors (L1/L2 absorb the 3x reuse).
// The arithmetic is 2x16-bit packed: vertical 1-2-1 on widened bytes, horizontal difference with a
// bias that keeps both halves non-negative, clamp with VIMNMX.U16x2.
#include "common.cuh"

namespace u96 {

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }

// 16 pixels of a row widened to eight u16x2 registers (px0,px1) ... (px14,px15): every input row is widened once and
// then serves three output rows (the vertical 1-2-1 sums are plain 32-bit adds: at most 4*255 per half, no carry)
struct WideRow { uint32_t v[8]; };
__device__ __forceinline__ WideRow widen16(const uint4 r)
{
    WideRow w;
    w.v[0] = prmt(r.x, 0, 0x4140); w.v[1] = prmt(r.x, 0, 0x4342);
    w.v[2] = prmt(r.y, 0, 0x4140); w.v[3] = prmt(r.y, 0, 0x4342);
    w.v[4] = prmt(r.z, 0, 0x4140); w.v[5] = prmt(r.z, 0, 0x4342);
    w.v[6] = prmt(r.w, 0, 0x4140); w.v[7] = prmt(r.w, 0, 0x4342);
    return w;
}

// One thread = 16 pixels x XS_R consecutive rows: the XS_R+2 input rows are all requested before any
// arithmetic (enough bytes in flight per SM to cover the HBM latency, each row fetched 1.5x instead of 3x),
// then a rolling window of three widened rows produces the outputs.
constexpr int XS_R = 4;

template <int PROFILE>
__global__ void __launch_bounds__(256) k_xsobel(const uint8_t *__restrict__ srcL, const uint8_t *__restrict__ srcR, int sp, size_t sf,
                                                uint8_t *__restrict__ dL, uint8_t *__restrict__ dR, int dp, size_t df,
                                                int W, int H, int cap)
{
    const int w16 = (W + 15) >> 4;
    const int nrg = (H + XS_R - 1) / XS_R;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;       // flattened (row group, 16-pixel group)
    if (item >= w16 * nrg) return;
    const int rg = item / w16;
    const int x0 = (item - rg * w16) * 16;
    const int y0 = rg * XS_R;
    const int lr = blockIdx.y & 1, f = blockIdx.y >> 1;
    const uint8_t *src = (lr ? srcR : srcL) + (size_t)f * sf + x0;
    uint8_t *dst = (lr ? dR : dL) + (size_t)f * df + (size_t)y0 * dp + x0;

    const uint32_t border = (PROFILE == U96_PROFILE_RTL) ? 32u : (uint32_t)cap;
    // s + BIAS, BIAS = 1024 + off keeps both halves positive (|s| <= 1020); clamp to [lo_off, hi_lim] afterwards
    const uint32_t bias = (1024u + border) * 0x00010001u;
    const uint32_t lo_off = 1024u * 0x00010001u;                                       // value 0 after clamp
    const uint32_t hi_lim = (1024u + ((PROFILE == U96_PROFILE_RTL) ? 63u : 2u * (uint32_t)cap)) * 0x00010001u;
    // neighbours left of x0 and right of x0+15 (clamped addresses; the border columns are overwritten below)
    const int xl = (x0 > 0) ? -1 : 0, xr = (x0 + 16 < W) ? 16 : 15;

    uint4 row[XS_R + 2];
    uint32_t nl[XS_R + 2], nr[XS_R + 2];
#pragma unroll
    for (int j = 0; j < XS_R + 2; j++) {
        const int yy = y0 - 1 + j;
        int r;
        if (PROFILE == U96_PROFILE_RTL) r = min(max(yy, 0), H - 1);                    // rows 0 / H-1 are not outputs
        else r = (yy < 0) ? 1 : (yy == H) ? H - 2 : min(yy, H - 1);                    // reflect-101
        const uint8_t *p = src + (size_t)r * sp;
        row[j] = *reinterpret_cast<const uint4 *>(p);
        nl[j] = p[xl]; nr[j] = p[xr];
    }
    WideRow wa = widen16(row[0]), wb = widen16(row[1]);
#pragma unroll
    for (int i = 0; i < XS_R; i++) {
        const int y = y0 + i;
        if (y >= H) break;
        const WideRow wc = widen16(row[i + 2]);
        uint4 o4;
        if (PROFILE == U96_PROFILE_RTL && (y == 0 || y == H - 1)) o4 = make_uint4(0, 0, 0, 0);       // invalid lines
        else if (PROFILE == U96_PROFILE_OPENCV && (H & 1) && y == H - 1) {
            const uint32_t c4 = (uint32_t)cap * 0x01010101u;
            o4 = make_uint4(c4, c4, c4, c4);
        } else {
            uint32_t v[10];      // v[0] = (-, px-1) ; v[1..8] = pairs (px0,px1)...(px14,px15) ; v[9] = (px16, -)
            v[0] = (nl[i] + 2u * nl[i + 1] + nl[i + 2]) << 16;
#pragma unroll
            for (int k = 0; k < 8; k++) v[1 + k] = wa.v[k] + 2u * wb.v[k] + wc.v[k];
            v[9] = nr[i] + 2u * nr[i + 1] + nr[i + 2];
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t r2[2];
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const int ii = 1 + 2 * q + k;                          // pair (x, x+1)
                    const uint32_t nxt = prmt(v[ii], v[ii + 1], 0x5432);   // (x+1, x+2)
                    const uint32_t prv = prmt(v[ii - 1], v[ii], 0x5432);   // (x-1, x)
                    uint32_t t = nxt + bias - prv;                         // s + off + 1024 per half, no borrow
                    t = __vmaxu2(t, lo_off);
                    t = __vminu2(t, hi_lim);
                    r2[k] = t;
                }
                o[q] = prmt(r2[0], r2[1], 0x6420);                        // low bytes: 1024 = 0x400 drops out
            }
            // column borders                                              xsbl2.v:869-872
            if (x0 == 0) o[0] = (o[0] & 0xFFFFFF00u) | border;
            if (x0 + 16 >= W) {
                const int k = W - 1 - x0;                                  // 0..15
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (q == (k >> 2)) o[q] = (o[q] & ~(0xFFu << (8 * (k & 3)))) | (border << (8 * (k & 3)));
            }
            o4 = make_uint4(o[0], o[1], o[2], o[3]);
        }
        *reinterpret_cast<uint4 *>(dst + (size_t)i * dp) = o4;
        wa = wb; wb = wc;
    }
}

int launch_xsobel(const uint8_t *srcL, const uint8_t *srcR, int src_pitch, size_t src_frame,
                  Img8 dstL, Img8 dstR, int W, int H, int n, int profile, int cap, cudaStream_t s)
{
    const int tx = 256;
    dim3 grid((align_up(W, 16) / 16 * ((H + XS_R - 1) / XS_R) + tx - 1) / tx, 2 * n);
    if (profile == U96_PROFILE_RTL)
        k_xsobel<U96_PROFILE_RTL><<<grid, tx, 0, s>>>(srcL, srcR, src_pitch, src_frame, dstL.p, dstR.p, dstL.pitch, dstL.frame, W, H, cap);
    else
        k_xsobel<U96_PROFILE_OPENCV><<<grid, tx, 0, s>>>(srcL, srcR, src_pitch, src_frame, dstL.p, dstR.p, dstL.pitch, dstL.frame, W, H, cap);
    return 1;
}

}  // namespace u96
