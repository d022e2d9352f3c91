// uvc.cu -- the UVC payload formatter of the R5 firmware as a kernel (StereoBM/src/xusb_main.c:293-376).
//
// Per stereo pair one YUYV frame of 2W x H pixels: Y byte + chroma byte 0x80 per pixel, left half = left image
// (or the disparity >> 4, truncated to 8 bit), right half = right image (or zero).  Pure streaming: 2-4 B/px read,
// 4 B/px written; one thread converts 8 pixels of both halves with 16-byte loads/stores.
#include "common.cuh"

namespace u96 {

__device__ __forceinline__ uint4 yuyv8(uint2 y)      // 8 luma bytes -> 8 (Y, 0x80) pairs
{
    uint4 o;
    o.x = __byte_perm(y.x, 0x80808080u, 0x4140);
    o.y = __byte_perm(y.x, 0x80808080u, 0x4342);
    o.z = __byte_perm(y.y, 0x80808080u, 0x4140);
    o.w = __byte_perm(y.y, 0x80808080u, 0x4342);
    return o;
}

// MODE 0: two u8 images (rect / xsbl);  MODE 1: s16 disparity on the left, zero on the right
template <int MODE>
__global__ void __launch_bounds__(256) k_pack_uvc(const uint8_t *__restrict__ srcL, const uint8_t *__restrict__ srcR, int sp, size_t sf,
                                                  const int16_t *__restrict__ disp, int dp, size_t df,
                                                  uint8_t *__restrict__ out, int W, int H)
{
    const int w8 = (W + 7) >> 3;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= w8 * H) return;
    const int y = item / w8, x0 = (item - y * w8) * 8, f = blockIdx.y;
    uint2 yl, yr;
    if (MODE == 0) {
        yl = *reinterpret_cast<const uint2 *>(srcL + (size_t)f * sf + (size_t)y * sp + x0);
        yr = *reinterpret_cast<const uint2 *>(srcR + (size_t)f * sf + (size_t)y * sp + x0);
    } else {
        const uint4 d = *reinterpret_cast<const uint4 *>(disp + (size_t)f * df + (size_t)y * dp + x0);   // 8 x s16
        // (u8)(s16 >> 4): bits 4..11 of every half
        const uint32_t a = (d.x >> 4) & 0x00FF00FFu, b = (d.y >> 4) & 0x00FF00FFu, c = (d.z >> 4) & 0x00FF00FFu, e = (d.w >> 4) & 0x00FF00FFu;
        yl.x = __byte_perm(a, b, 0x6420);
        yl.y = __byte_perm(c, e, 0x6420);
        yr = make_uint2(0u, 0u);
    }
    uint8_t *row = out + ((size_t)f * H + y) * (size_t)W * 4;          // 2W pixels x 2 bytes
    const uint4 ol = yuyv8(yl), orr = yuyv8(yr);
    if (x0 + 8 <= W && (W & 7) == 0) {                                  // 16-byte aligned stores need W % 8 == 0
        *reinterpret_cast<uint4 *>(row + (size_t)x0 * 2) = ol;
        *reinterpret_cast<uint4 *>(row + (size_t)(W + x0) * 2) = orr;
    } else {                                                           // ragged widths
        const uint32_t wl[4] = {ol.x, ol.y, ol.z, ol.w}, wr[4] = {orr.x, orr.y, orr.z, orr.w};
        for (int k = 0; k < 8 && x0 + k < W; k++) {
            const uint16_t vl = (uint16_t)(wl[k >> 1] >> (16 * (k & 1))), vr = (uint16_t)(wr[k >> 1] >> (16 * (k & 1)));
            *reinterpret_cast<uint16_t *>(row + (size_t)(x0 + k) * 2) = vl;
            *reinterpret_cast<uint16_t *>(row + (size_t)(W + x0 + k) * 2) = vr;
        }
    }
}

int launch_pack_uvc(const uint8_t *srcL, const uint8_t *srcR, int sp, size_t sf, const int16_t *disp, int dp, size_t df,
                    uint8_t *out, int W, int H, int n, cudaStream_t s)
{
    dim3 grid((((W + 7) / 8) * H + 255) / 256, n);
    if (disp) k_pack_uvc<1><<<grid, 256, 0, s>>>(nullptr, nullptr, 0, 0, disp, dp, df, out, W, H);
    else k_pack_uvc<0><<<grid, 256, 0, s>>>(srcL, srcR, sp, sf, nullptr, 0, 0, out, W, H);
    return 1;
}

}  // namespace u96
