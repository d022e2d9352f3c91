// gftt.cu -- min-eigenvalue (Shi-Tomasi) map of the FPGA's GFTT accelerator, sm_100a.
//
// Reference: dvp/rtl/gftt.v and its stages; consumer Fpga::receiveEigen
// (slam/src/core/FPGA.cpp:281-296) -> generateKeypoints2 (slam/src/core/GFTT.cpp:41-170).
//   gftt_sbl.v:118-205      dx, dy = 3x3 Sobel of the rectified LEFT image, 0 at columns 0 and W-1
//   gftt_eig.v:104-124      dx2 = |dx|^2 >> 6, dy2 = |dy|^2 >> 6, dxdy = |dx||dy| >> 6   (sign of dx*dy dropped)
//   gftt_box.v:185,232-262  3x3 box sums a, c, b (horizontal sums forced to 0 at columns 0 and W-1, limit 0xFFFF)
//   gftt_eig.v:196-310      eig = (a+c) - floor(sqrt((|a-c|^2 >> 10) + (b^2 >> 8) limited to 22 bit, << 10)), clamped to u16
//   gftt_obuf.v:90-118      per-frame maximum; rows 2..H-3 are written, the rest of the bank stays 0 (fpga.c:107-108)
//
// A pixel depends on a 5x5 neighbourhood.  One thread owns 4 adjacent columns and marches down a strip of rows
// with everything rolling in registers: three input rows (8 bytes each, incl. the 2-pixel halo), three rows of the
// horizontal 3-sums of the three products.  HBM traffic is the algorithmic 1 B/px in + 2 B/px out (the halo
// re-reads are L1/L2 hits); the integer pipes bind first (three 32-bit products, a square root and three limiters
// per pixel), not HBM.
#include <type_traits>

#include "common.cuh"

namespace u96 {

constexpr int GF_PX = 4;            // columns per thread
constexpr int GF_RS = 28;           // output rows per strip (640x480: 476 = 17 x 28)

// floor(sqrt(x)), x = s << 10 with s < 2^22 (so sqrt(x) = 32 sqrt(s) and s is exact in fp32): approximate root, round to
// nearest through the 2^23 mantissa trick (no F2I), then one exact downward correction on integers.
__device__ __forceinline__ uint32_t gf_isqrt_s10(uint32_t s)
{
    float f;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(__uint2float_rn(s)));
    const float m = __fmaf_rn(f, 32.0f, 8388608.0f);                 // round(32 f) in the low mantissa bits
    uint32_t q = min(__float_as_uint(m) & 0x7FFFFFu, 65535u);        // |32 f - root| < 0.05  ->  q in {floor(root), floor(root)+1}
    if (q * q > (s << 10)) q--;
    return q;
}

__device__ __forceinline__ uint32_t gf_pabsdiff(uint32_t a, uint32_t b) { return __vmaxu2(a, b) - __vminu2(a, b); }   // |a-b| per u16 half

// Packed 2x16-bit formulation.  Columns x0-2 .. x0+5 of a row are four u16x2 registers W[0..3] = (x0-2,x0-1) (x0,x0+1) ...;
// "odd" pairs (x0-1,x0) (x0+1,x0+2) (x0+3,x0+4) come from one PRMT each.  With S = vertical 1-2-1 sum and G = horizontal
// 1-2-1 sum, |dx|(x) = |S(x+1) - S(x-1)| and |dy|(x) = |G(y+1) - G(y-1)| are lane-wise max-min of unsigned halves; the three
// products need 32 bits and are formed per pixel.
__global__ void __launch_bounds__(128) k_gftt_eig(const uint8_t *__restrict__ src, int sp, size_t sf,
                                                  uint16_t *__restrict__ eig, int ep, size_t ef,
                                                  uint32_t *__restrict__ fmax, int W, int H)
{
    const int cols = (W + GF_PX - 1) / GF_PX;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;   // flattened (strip, 4-column group)
    const int strip = item / cols;
    const int x0 = (item - strip * cols) * GF_PX;
    const int f = blockIdx.z;
    const int ys = 2 + strip * GF_RS;                      // first output row of the strip
    const int ye = min(ys + GF_RS, H - 2);                 // one past the last output row
    uint32_t tmax = 0;
    if (x0 < W && ys < ye) {
        const uint8_t *img = src + (size_t)f * sf;
        uint16_t *out = eig + (size_t)f * ef;
        const int pw = sp >> 2;
        const int wc = x0 >> 2;                            // x0 is a multiple of 4: aligned word of columns x0..x0+3
        // live-column masks: Sobel taps are 0 at columns 0, W-1 and beyond (gftt_sbl.v:151,198), so are the horizontal
        // box sums (gftt_box.v:185).  Only the threads at the two image edges carry a non-trivial mask.
        uint32_t mk[3];                                    // odd pairs (x0-1,x0) (x0+1,x0+2) (x0+3,x0+4)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int c = x0 - 1 + 2 * j;
            mk[j] = ((c >= 1 && c <= W - 2) ? 0xFFFFu : 0u) | ((c + 1 >= 1 && c + 1 <= W - 2) ? 0xFFFF0000u : 0u);
        }
        const bool edge = (mk[0] & mk[1] & mk[2]) != 0xFFFFFFFFu;
        uint32_t hm[GF_PX];
#pragma unroll
        for (int j = 0; j < GF_PX; j++) hm[j] = (x0 + j >= 1 && x0 + j <= W - 2) ? 0xFFFFFFFFu : 0u;

        // the three aligned words around columns x0..x0+3 of an input row (zero outside the row; those columns only feed
        // masked taps); rows past the image are never consumed
        auto fetch_row = [&](int y, uint32_t (&q)[3]) {
            const uint32_t *row = reinterpret_cast<const uint32_t *>(img + (size_t)min(y, H - 1) * sp);
            q[0] = (wc > 0) ? __ldg(row + wc - 1) : 0u;
            q[1] = __ldg(row + wc);
            q[2] = (wc + 1 < pw) ? __ldg(row + wc + 1) : 0u;
        };
        // widened row (columns x0-2..x0+5) and its horizontal 1-2-1 sums at columns x0-1..x0+4
        auto widen_row = [&](const uint32_t (&q)[3], uint32_t (&w)[4], uint32_t (&g)[3]) {
            w[0] = __byte_perm(q[0], 0, 0x4342); w[1] = __byte_perm(q[1], 0, 0x4140);
            w[2] = __byte_perm(q[1], 0, 0x4342); w[3] = __byte_perm(q[2], 0, 0x4140);
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const uint32_t o = __byte_perm(w[j], w[j + 1], 0x5432);      // odd pair
                g[j] = w[j] + 2u * o + w[j + 1];                             // <= 1020 per half
            }
        };
        auto load_row = [&](int y, uint32_t (&w)[4], uint32_t (&g)[3]) { uint32_t q[3]; fetch_row(y, q); widen_row(q, w, g); };
        uint32_t wr[3][4], gr[3][3];                       // ring of the last three input rows (slot = row mod 3, compile time)
        load_row(ys - 2, wr[0], gr[0]);
        load_row(ys - 1, wr[1], gr[1]);
        uint32_t nx[3];                                    // input row y+1 of the coming step, requested one step ahead
        fetch_row(ys, nx);
        uint32_t hs[3][3][GF_PX];                          // [row slot][dx2, dy2, dxdy][column]: horizontal 3-sums (<= 48768)
#pragma unroll
        for (int s = 0; s < 3; s++)
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int j = 0; j < GF_PX; j++) hs[s][k][j] = 0;

        // Sobel rows ys-1 .. ye (each needs input rows y-1..y+1); output row y-1 once three Sobel rows are in.
        // slot = (y - (ys-1)) mod 3: input rows y-1, y, y+1 sit in ring entries slot, slot+1, slot+2 (mod 3).
        auto step = [&](const int y, auto slot_c, auto edge_c) {
            constexpr int slot = decltype(slot_c)::value, i0 = slot, i1 = (slot + 1) % 3, i2 = (slot + 2) % 3;
            constexpr bool EDGE = decltype(edge_c)::value;
            widen_row(nx, wr[i2], gr[i2]);
            fetch_row(y + 2, nx);                          // in flight during this step's arithmetic
            uint32_t sv[4];                                // vertical 1-2-1 sums, columns x0-2..x0+5
#pragma unroll
            for (int j = 0; j < 4; j++) sv[j] = wr[i0][j] + 2u * wr[i1][j] + wr[i2][j];
            uint32_t v[3][6];                              // products at columns x0-1 .. x0+4
#pragma unroll
            for (int j = 0; j < 3; j++) {
                uint32_t ax2 = gf_pabsdiff(sv[j + 1], sv[j]);              // |dx| at the odd pair j   gftt_sbl.v:118-160
                uint32_t ay2 = gf_pabsdiff(gr[i2][j], gr[i0][j]);          // |dy|                     gftt_sbl.v:166-205
                if (EDGE) { ax2 &= mk[j]; ay2 &= mk[j]; }
                const uint32_t ax0 = ax2 & 0xFFFFu, ax1 = ax2 >> 16, ay0 = ay2 & 0xFFFFu, ay1 = ay2 >> 16;
                v[0][2 * j] = (ax0 * ax0) >> 6; v[1][2 * j] = (ay0 * ay0) >> 6; v[2][2 * j] = (ax0 * ay0) >> 6;      // gftt_eig.v:104-124
                v[0][2 * j + 1] = (ax1 * ax1) >> 6; v[1][2 * j + 1] = (ay1 * ay1) >> 6; v[2][2 * j + 1] = (ax1 * ay1) >> 6;
            }
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int j = 0; j < GF_PX; j++) {
                    hs[slot][k][j] = v[k][j] + v[k][j + 1] + v[k][j + 2];                                              // gftt_box.v:232-247
                    if (EDGE) hs[slot][k][j] &= hm[j];                                                                 // gftt_box.v:185
                }
            if (y >= ys + 1) {
                const int yo = y - 1;
                uint32_t o[GF_PX];
#pragma unroll
                for (int j = 0; j < GF_PX; j++) {
                    const uint32_t a = min(hs[0][0][j] + hs[1][0][j] + hs[2][0][j], 0xFFFFu);                          // gftt_box.v:249-262
                    const uint32_t c = min(hs[0][1][j] + hs[1][1][j] + hs[2][1][j], 0xFFFFu);
                    const uint32_t b = min(hs[0][2][j] + hs[1][2][j] + hs[2][2][j], 0xFFFFu);
                    const uint32_t amc = __usad(a, c, 0u);                                                             // gftt_eig.v:213-226
                    const uint32_t s = __viaddmin_u32((amc * amc) >> 10, (b * b) >> 8, 0x3FFFFFu);                      // gftt_eig.v:238-262
                    const int e = (int)(a + c) - (int)gf_isqrt_s10(s);                                                 // gftt_eig.v:293
                    o[j] = (uint32_t)min(max(e, 0), 0xFFFF);                                                           // gftt_eig.v:296-308
                    tmax = max(tmax, o[j]);
                }
                uint16_t *dst = out + (size_t)yo * ep + x0;
                if (!EDGE || x0 + GF_PX <= W) *reinterpret_cast<uint2 *>(dst) = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
                else
                    for (int j = 0; j < GF_PX; j++) if (x0 + j < W) dst[j] = (uint16_t)o[j];
            }
        };
        auto run = [&](auto edge_c) {
            for (int y = ys - 1; y <= ye; y += 3) {        // the ring slots rotate at compile time
                step(y, std::integral_constant<int, 0>(), edge_c);
                if (y + 1 <= ye) step(y + 1, std::integral_constant<int, 1>(), edge_c);
                if (y + 2 <= ye) step(y + 2, std::integral_constant<int, 2>(), edge_c);
            }
        };
        // warp-uniform choice (a warp that holds an edge thread runs the masked variant as a whole: no divergence)
        if (__any_sync(__activemask(), edge)) run(std::true_type()); else run(std::false_type());
    }
    tmax = __reduce_max_sync(0xFFFFFFFFu, tmax);
    if ((threadIdx.x & 31) == 0 && tmax) atomicMax(fmax + f, tmax);
}

// rows 0,1,H-2,H-1 of every frame (never written by the FPGA, zero in the bank) and the per-frame maxima
__global__ void k_gftt_clear(uint16_t *eig, int ep, size_t ef, uint32_t *fmax, int W, int H)
{
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) fmax[f] = 0;
    if (i < 4 * W) {
        const int r = i / W, x = i - r * W;
        const int y = (r < 2) ? r : H - 4 + r;
        if (y >= 0 && y < H) eig[(size_t)f * ef + (size_t)y * ep + x] = 0;
    }
}

int launch_gftt(const uint8_t *src, int sp, size_t sf, uint16_t *eig, int ep, size_t ef, uint32_t *fmax,
                int W, int H, int n, cudaStream_t s)
{
    k_gftt_clear<<<dim3((4 * W + 255) / 256, n), 256, 0, s>>>(eig, ep, ef, fmax, W, H);
    if (H < 5) return 1;
    const int tx = 128;
    const int cols = (W + GF_PX - 1) / GF_PX;
    const int strips = (H - 4 + GF_RS - 1) / GF_RS;
    dim3 grid((cols * strips + tx - 1) / tx, 1, n);
    k_gftt_eig<<<grid, tx, 0, s>>>(src, sp, sf, eig, ep, ef, fmax, W, H);
    return 2;
}

}  // namespace u96
