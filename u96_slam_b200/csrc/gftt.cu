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
// re-reads are L1/L2 hits); the integer pipes bind first (~70 instructions per pixel).
#include <type_traits>

#include "common.cuh"

namespace u96 {

constexpr int GF_PX = 4;            // columns per thread
constexpr int GF_RS = 28;           // output rows per strip (640x480: 476 = 17 x 28)

// floor(sqrt(x)), x < 2^32: approximate float root, then an exact +-1 correction on integers
__device__ __forceinline__ uint32_t gf_isqrt(uint32_t x)
{
    float f;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(__uint2float_rn(x)));
    uint32_t q = min((uint32_t)f, 65535u);
    if (q * q > x) q--;
    else if (q < 65535u && (q + 1u) * (q + 1u) <= x) q++;
    return q;
}

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

        // bytes x0-2 .. x0+5 of an input row (zero outside the row; those columns only feed zeroed Sobel taps)
        auto load_row = [&](int y, uint32_t &lo, uint32_t &hi) {
            const uint32_t *row = reinterpret_cast<const uint32_t *>(img + (size_t)y * sp);
            const uint32_t wl = (wc > 0) ? __ldg(row + wc - 1) : 0u;
            const uint32_t wm = __ldg(row + wc);
            const uint32_t wr = (wc + 1 < pw) ? __ldg(row + wc + 1) : 0u;
            lo = __byte_perm(wl, wm, 0x5432);              // x0-2, x0-1, x0, x0+1
            hi = __byte_perm(wm, wr, 0x5432);              // x0+2 .. x0+5
        };
        uint32_t r0l, r0h, r1l, r1h, r2l, r2h;             // rows y-1, y, y+1 of the Sobel row y
        load_row(ys - 2, r0l, r0h);
        load_row(ys - 1, r1l, r1h);
        uint32_t hs[3][3][GF_PX];                          // [row slot][dx2, dy2, dxdy][column]
#pragma unroll
        for (int s = 0; s < 3; s++)
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int j = 0; j < GF_PX; j++) hs[s][k][j] = 0;

        // Sobel rows ys-1 .. ye (each needs input rows y-1..y+1); output row y-1 once three Sobel rows are in
        auto step = [&](const int y, auto slot_c) {
            constexpr int slot = decltype(slot_c)::value;
            load_row(y + 1, r2l, r2h);
            int p0[8], p1[8], p2[8];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                p0[j] = (r0l >> (8 * j)) & 0xFF; p0[4 + j] = (r0h >> (8 * j)) & 0xFF;
                p1[j] = (r1l >> (8 * j)) & 0xFF; p1[4 + j] = (r1h >> (8 * j)) & 0xFF;
                p2[j] = (r2l >> (8 * j)) & 0xFF; p2[4 + j] = (r2h >> (8 * j)) & 0xFF;
            }
            int sv[8], dv[8];                              // vertical 1-2-1 sums and vertical differences, columns x0-2..x0+5
#pragma unroll
            for (int j = 0; j < 8; j++) { sv[j] = p0[j] + 2 * p1[j] + p2[j]; dv[j] = p2[j] - p0[j]; }
            uint32_t v[3][6];                              // products at columns x0-1 .. x0+4
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const int c = x0 - 1 + j;
                const int dx = sv[j + 2] - sv[j];                          // gftt_sbl.v:118-160
                const int dy = dv[j] + 2 * dv[j + 1] + dv[j + 2];          // gftt_sbl.v:166-205
                const bool live = (c >= 1) && (c <= W - 2);                // first/last sample and beyond: 0
                const uint32_t ax = live ? (uint32_t)abs(dx) : 0u, ay = live ? (uint32_t)abs(dy) : 0u;
                v[0][j] = (ax * ax) >> 6; v[1][j] = (ay * ay) >> 6; v[2][j] = (ax * ay) >> 6;
            }
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int j = 0; j < GF_PX; j++) {
                    const int c = x0 + j;
                    hs[slot][k][j] = (c >= 1 && c <= W - 2) ? v[k][j] + v[k][j + 1] + v[k][j + 2] : 0u;   // gftt_box.v:185
                }
            if (y >= ys + 1) {
                const int yo = y - 1;
                uint32_t o[GF_PX];
#pragma unroll
                for (int j = 0; j < GF_PX; j++) {
                    const uint32_t a = min(hs[0][0][j] + hs[1][0][j] + hs[2][0][j], 0xFFFFu);
                    const uint32_t c = min(hs[0][1][j] + hs[1][1][j] + hs[2][1][j], 0xFFFFu);
                    const uint32_t b = min(hs[0][2][j] + hs[1][2][j] + hs[2][2][j], 0xFFFFu);
                    const uint32_t amc = (a > c) ? a - c : c - a;
                    const uint32_t s = min(((amc * amc) >> 10) + ((b * b) >> 8), 0x3FFFFFu);
                    const int e = (int)(a + c) - (int)gf_isqrt(s << 10);
                    o[j] = (e < 0) ? 0u : (e & 0x10000) ? 0xFFFFu : (uint32_t)e;
                    tmax = max(tmax, o[j]);
                }
                uint16_t *dst = out + (size_t)yo * ep + x0;
                if (x0 + GF_PX <= W) *reinterpret_cast<uint2 *>(dst) = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
                else
                    for (int j = 0; j < GF_PX; j++) if (x0 + j < W) dst[j] = (uint16_t)o[j];
            }
            r0l = r1l; r0h = r1h; r1l = r2l; r1h = r2h;
        };
        for (int y = ys - 1; y <= ye; y += 3) {            // the row slots rotate at compile time
            step(y, std::integral_constant<int, 0>());
            if (y + 1 <= ye) step(y + 1, std::integral_constant<int, 1>());
            if (y + 2 <= ye) step(y + 2, std::integral_constant<int, 2>());
        }
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
