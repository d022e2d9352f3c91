// reproject.cu -- 16x fixed-point disparity -> XYZ, sm_100a.
//
//   k_reproject        : projectDisparityTo3D (slam/src/core/Stereo.cpp:157-182) over the (decimated) map as the
//                        reference's dense consumer does (slam/src/core/main.cpp:522-551, decimation
//                        SensorData.cpp:50-58), followed by transformPoint (Stereo.cpp:189-198) with the camera
//                        model's localTransform and the frame's pose when given.
//   k_reproject_points : generateKeypoints3DStereo (Stereo.cpp:53-117) for the dense-map depth methods: gather
//                        at (int)y,(int)x, d<0 -> 0, skip 0, projection with the FLOAT keypoint coordinates,
//                        min/max-depth gates, localTransform unless null.
// The reference mixes float and double; every rounding step is reproduced with explicit _rn intrinsics so
// the compiler cannot contract anything into an FMA.
#include <math_constants.h>

#include "common.cuh"

namespace u96 {

// std::numeric_limits<float>::quiet_NaN() of the reference's hosts (x86-64, AArch64): 0x7FC00000.  CUDA's CUDART_NAN_F is 0x7FFFFFFF;
// the bad-point marker is kept bit-identical to the reference's.
#define U96_QNAN __int_as_float(0x7FC00000)

struct ReprojConst {
    double cx_l, cy_l, fx_l;
    double nx, ny;        // Tx_l/fx_l - Tx_r/fx_r ; Tx_l/fy_l - Tx_r/fy_r   (IEEE double, host computed)
    float c;              // (float)(cx_r - cx_l)
    float T[12];          // localTransform, 3x4 row-major
    int has_T;
};

// Stereo.cpp:157-182 for disp > 0
__device__ __forceinline__ void project_one(const ReprojConst &k, float u, float v, float d, float &X, float &Y, float &Z)
{
    const float dc = __fadd_rn(d, k.c);                           // float + float
    const float Wx = __double2float_rn(__ddiv_rn(k.nx, (double)dc));
    const float Wy = __double2float_rn(__ddiv_rn(k.ny, (double)dc));
    X = __double2float_rn(__dmul_rn(__dsub_rn((double)u, k.cx_l), (double)Wx));
    Y = __double2float_rn(__dmul_rn(__dsub_rn((double)v, k.cy_l), (double)Wy));
    Z = __double2float_rn(__dmul_rn(k.fx_l, (double)Wx));
}

// Stereo.cpp:189-198: float products summed left to right
__device__ __forceinline__ void transform_point(const float *t, float &X, float &Y, float &Z)
{
    const float x = X, y = Y, z = Z;
    X = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t[0], x), __fmul_rn(t[1], y)), __fmul_rn(t[2], z)), t[3]);
    Y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t[4], x), __fmul_rn(t[5], y)), __fmul_rn(t[6], z)), t[7]);
    Z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t[8], x), __fmul_rn(t[9], y)), __fmul_rn(t[10], z)), t[11]);
}

// One thread per output sample; the three floats of 256 consecutive samples leave through shared memory as
// 16-byte vector stores (the sample stride of 12 B would otherwise cost three partial-sector stores per warp).
__global__ void __launch_bounds__(256) k_reproject(const int16_t *__restrict__ disp, int dpitch, size_t dframe, int ow, int oh,
                                                   int n, ReprojConst k, int decim, const float *__restrict__ poses,
                                                   float *__restrict__ xyz)
{
    __shared__ __align__(16) float s_out[256 * 3];
    const size_t base = (size_t)blockIdx.x * 256;
    const size_t idx = base + threadIdx.x;
    const size_t per = (size_t)ow * oh, total = per * n;
    float X = U96_QNAN, Y = U96_QNAN, Z = U96_QNAN;
    if (idx < total) {
        const int f = (int)(idx / per);
        const int rem = (int)(idx - (size_t)f * per);
        const int row = rem / ow, col = rem - row * ow;
        const int16_t s = disp[(size_t)f * dframe + (size_t)(row * decim) * dpitch + col * decim];
        const float d = __fdiv_rn((float)s, 16.0f);               // main.cpp:529
        if (d > 0.0f) {
            project_one(k, (float)(col * decim), (float)(row * decim), d, X, Y, Z);
            if (k.has_T || poses) {
                if (isfinite(X) && isfinite(Y) && isfinite(Z)) {  // main.cpp:535-538
                    if (k.has_T) transform_point(k.T, X, Y, Z);
                    if (poses) {
                        float p[12];
#pragma unroll
                        for (int i = 0; i < 12; i++) p[i] = __ldg(poses + (size_t)f * 12 + i);
                        transform_point(p, X, Y, Z);
                    }
                } else { X = Y = Z = U96_QNAN; }              // the consumer drops non-finite points
            }
        }
    }
    s_out[threadIdx.x * 3 + 0] = X; s_out[threadIdx.x * 3 + 1] = Y; s_out[threadIdx.x * 3 + 2] = Z;
    __syncthreads();
    const size_t left = (total > base) ? (total - base) : 0;
    const int nfl = (int)min((size_t)768, left * 3);              // floats this block owns
    float *o = xyz + base * 3;                                    // 256*3*4 B per block: 16-byte aligned when xyz is
    if (threadIdx.x < 192) {
        const int i4 = threadIdx.x * 4;
        if (i4 + 3 < nfl) *reinterpret_cast<float4 *>(o + i4) = *reinterpret_cast<const float4 *>(s_out + i4);
        else for (int j = i4; j < nfl; j++) o[j] = s_out[j];
    }
}

__global__ void __launch_bounds__(128) k_reproject_points(const int16_t *__restrict__ disp, int dpitch, int W, int H, ReprojConst k,
                                                          const float2 *__restrict__ uv, const uint8_t *__restrict__ mask, int n,
                                                          float min_depth, float max_depth, float *__restrict__ xyz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float X = U96_QNAN, Y = U96_QNAN, Z = U96_QNAN;
    const float2 p = uv[i];
    // a keypoint outside the map is undefined behaviour in the reference (cv::Mat::at without a check): bad point here
    if ((!mask || mask[i]) && p.x > -1.0f && p.x < (float)W && p.y > -1.0f && p.y < (float)H) {
        const int xi = (int)p.x, yi = (int)p.y;                   // Stereo.cpp:79 truncates toward zero
        const int16_t s = disp[(size_t)yi * dpitch + xi];
        float d = __fdiv_rn((float)s, 16.0f);
        if (d < 0.0f) d = 0.0f;                                   // :81-83
        if (d != 0.0f) {
            float x3, y3, z3;
            project_one(k, p.x, p.y, d, x3, y3, z3);
            if (isfinite(x3) && isfinite(y3) && isfinite(z3) && (min_depth < 0.0f || z3 > min_depth) &&
                (max_depth <= 0.0f || z3 <= max_depth)) {         // :100-104
                if (k.has_T) transform_point(k.T, x3, y3, z3);
                X = x3; Y = y3; Z = z3;
            }
        }
    }
    xyz[3 * (size_t)i] = X; xyz[3 * (size_t)i + 1] = Y; xyz[3 * (size_t)i + 2] = Z;
}

static ReprojConst make_const(const double *P_l, const double *P_r, const float *local_T)
{
    ReprojConst k;
    const double fx_l = P_l[0], fy_l = P_l[5], Tx_l = P_l[3];
    const double fx_r = P_r[0], fy_r = P_r[5], Tx_r = P_r[3];
    k.cx_l = P_l[2]; k.cy_l = P_l[6]; k.fx_l = fx_l;
    volatile double a = Tx_l / fx_l, b = Tx_r / fx_r, c = Tx_l / fy_l, d = Tx_r / fy_r;
    k.nx = a - b; k.ny = c - d;
    k.c = (float)(P_r[2] - P_l[2]);
    k.has_T = 0;
    for (int i = 0; i < 12; i++) {
        k.T[i] = local_T ? local_T[i] : 0.0f;
        if (k.T[i] != 0.0f) k.has_T = 1;                          // Transform::isNull (Transform.cpp:88-95): all zero = no transform
    }
    return k;
}

int launch_reproject(const int16_t *disp, int dpitch, size_t dframe, int W, int H, int n,
                     const double *P_l, const double *P_r, int decim, const float *local_T, const float *d_poses,
                     float *xyz, cudaStream_t s)
{
    const ReprojConst k = make_const(P_l, P_r, local_T);
    const int ow = W / decim, oh = H / decim;
    const size_t total = (size_t)ow * oh * n;
    k_reproject<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(disp, dpitch, dframe, ow, oh, n, k, decim, d_poses, xyz);
    return 1;
}

int launch_reproject_points(const int16_t *disp, int dpitch, int W, int H, const double *P_l, const double *P_r,
                            const float *d_uv, const uint8_t *d_mask, int n, float min_depth, float max_depth,
                            const float *local_T, float *xyz, cudaStream_t s)
{
    const ReprojConst k = make_const(P_l, P_r, local_T);
    k_reproject_points<<<(n + 127) / 128, 128, 0, s>>>(disp, dpitch, W, H, k, reinterpret_cast<const float2 *>(d_uv), d_mask, n,
                                                       min_depth, max_depth, xyz);
    return 1;
}

}  // namespace u96
