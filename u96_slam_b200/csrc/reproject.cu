// reproject.cu -- 16x fixed-point disparity -> XYZ point cloud, sm_100a.
//
// Restates projectDisparityTo3D (slam/src/core/Stereo.cpp:157-182) over the (decimated) map as the
// reference's dense consumer does (slam/src/core/main.cpp:522-551, decimation SensorData.cpp:50-58).
// The reference mixes float and double; every rounding step is reproduced with explicit _rn
// intrinsics so the compiler cannot contract anything into an FMA.
#include <math_constants.h>

#include "common.cuh"

namespace u96 {

struct ReprojConst {
    double cx_l, cy_l, fx_l;
    double nx, ny;        // Tx_l/fx_l - Tx_r/fx_r ; Tx_l/fy_l - Tx_r/fy_r   (IEEE double, host computed)
    float c;              // (float)(cx_r - cx_l)
};

__global__ void __launch_bounds__(256) k_reproject(const int16_t *__restrict__ disp, int dpitch, size_t dframe, int ow, int oh,
                                                   int n, ReprojConst k, int decim, int flags, float *__restrict__ xyz)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per = (size_t)ow * oh;
    if (idx >= per * n) return;
    const int f = (int)(idx / per);
    const int rem = (int)(idx - (size_t)f * per);
    const int row = rem / ow, col = rem - row * ow;
    const int16_t s = disp[(size_t)f * dframe + (size_t)(row * decim) * dpitch + col * decim];
    const float d = __fdiv_rn((float)s, 16.0f);                   // main.cpp:529
    float X = CUDART_NAN_F, Y = CUDART_NAN_F, Z = CUDART_NAN_F;
    if (d > 0.0f) {
        const float u = (float)(col * decim), v = (float)(row * decim);
        const float dc = __fadd_rn(d, k.c);                       // float + float
        const float Wx = __double2float_rn(__ddiv_rn(k.nx, (double)dc));
        const float Wy = __double2float_rn(__ddiv_rn(k.ny, (double)dc));
        X = __double2float_rn(__dmul_rn(__dsub_rn((double)u, k.cx_l), (double)Wx));
        Y = __double2float_rn(__dmul_rn(__dsub_rn((double)v, k.cy_l), (double)Wy));
        Z = __double2float_rn(__dmul_rn(k.fx_l, (double)Wx));
        if ((flags & 1) && isfinite(X) && isfinite(Y) && isfinite(Z)) {
            // localTransform (StereoCameraModel.cpp:9-14): z-forward camera -> x-forward body
            const float tx = Z, ty = -X, tz = -Y;
            X = tx; Y = ty; Z = tz;
        }
    }
    float *o = xyz + idx * 3;
    o[0] = X; o[1] = Y; o[2] = Z;
}

int launch_reproject(const int16_t *disp, int dpitch, size_t dframe, int W, int H, int n,
                     const double *P_l, const double *P_r, int decim, int flags, float *xyz, cudaStream_t s)
{
    ReprojConst k;
    const double fx_l = P_l[0], fy_l = P_l[5], Tx_l = P_l[3];
    const double fx_r = P_r[0], fy_r = P_r[5], Tx_r = P_r[3];
    k.cx_l = P_l[2]; k.cy_l = P_l[6]; k.fx_l = fx_l;
    volatile double a = Tx_l / fx_l, b = Tx_r / fx_r, c = Tx_l / fy_l, d = Tx_r / fy_r;
    k.nx = a - b; k.ny = c - d;
    k.c = (float)(P_r[2] - P_l[2]);
    const int ow = W / decim, oh = H / decim;
    const size_t total = (size_t)ow * oh * n;
    k_reproject<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(disp, dpitch, dframe, ow, oh, n, k, decim, flags, xyz);
    return 1;
}

}  // namespace u96
