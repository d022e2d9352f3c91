// postfilter.cu -- the post filters cv::StereoBM::compute applies as configured by the reference's CPU mode
// (slam/src/core/main.cpp:210-212: disp12MaxDiff 1, speckleWindowSize 50, speckleRange 32), sm_100a.
//
//   k_validate  : cv::validateDisparity.  Per image row: the right-image disparity is the left disparity of the
//                 cheapest pixel that maps onto it (first pixel wins ties) -- a shared-memory atomicMin on the
//                 key (cost as signed short, biased)<<32 | x<<16 | d -- then a left pixel is dropped when BOTH roundings of its disparity
//                 hit a right pixel whose disparity differs by more than disp12MaxDiff.
//   k_cc_*      : cv::filterSpeckles.  4-connected components of the relation |d(p)-d(q)| <= maxDiff over valid
//                 pixels by lock-free union-find (roots = smallest pixel index, so labels are deterministic),
//                 component sizes by atomicAdd, components of at most maxSpeckleSize pixels become invalid.
//                 The result of the CPU flood fill does not depend on its traversal order, so the two agree.
#include "common.cuh"

namespace u96 {

constexpr int INVALID16 = -16;          // (minDisparity - 1) * 16, minDisparity == 0

__global__ void __launch_bounds__(256) k_validate(int16_t *__restrict__ disp, const int16_t *__restrict__ cost, int dpitch, size_t dframe,
                                                  int W, int ndisp, int maxdiff16)
{
    extern __shared__ unsigned long long s_key[];          // [W]
    const int y = blockIdx.x, f = blockIdx.y;
    int16_t *drow = disp + (size_t)f * dframe + (size_t)y * dpitch;
    const int16_t *crow = cost + (size_t)f * dframe + (size_t)y * dpitch;
    for (int x = threadIdx.x; x < W; x += blockDim.x) s_key[x] = ~0ull;
    __syncthreads();
    const int minX1 = ndisp;                                // max(minD + ndisp, 0), minD == 0 ; maxX1 = W
    for (int x = minX1 + threadIdx.x; x < W; x += blockDim.x) {
        const int d = drow[x];
        if (d == INVALID16) continue;
        const int x2 = x - ((d + 8) >> 4);
        if (x2 < 0 || x2 >= W) continue;                    // cannot happen for 0 <= d < 16*ndisp
        const unsigned long long k = ((unsigned long long)(unsigned)((int)crow[x] + 32768) << 32) | ((unsigned long long)x << 16) | (unsigned short)d;
        atomicMin(&s_key[x2], k);
    }
    __syncthreads();
    for (int x = minX1 + threadIdx.x; x < W; x += blockDim.x) {
        const int d = drow[x];
        if (d == INVALID16) continue;
        const int x0 = x - (d >> 4), x1 = x - ((d + 15) >> 4);
        bool bad0 = false, bad1 = false;
        if (x0 >= 0 && x0 < W) { const unsigned long long k = s_key[x0]; bad0 = (k != ~0ull) && (abs((int)(short)(k & 0xFFFF) - d) > maxdiff16); }
        if (x1 >= 0 && x1 < W) { const unsigned long long k = s_key[x1]; bad1 = (k != ~0ull) && (abs((int)(short)(k & 0xFFFF) - d) > maxdiff16); }
        if (bad0 && bad1) drow[x] = (int16_t)INVALID16;
    }
}

// ---- union-find over pixel indices of one frame (label[] is per frame, index = y*W + x) ----
__device__ __forceinline__ int uf_find(const int *label, int i)
{
    int r = label[i];
    while (r != label[r]) r = label[r];
    return r;
}
__device__ __forceinline__ void uf_union(int *label, int a, int b)
{
    while (true) {
        a = uf_find(label, a);
        b = uf_find(label, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }       // hook the larger root under the smaller one
        const int old = atomicMin(&label[b], a);
        if (old == b) return;
        b = old;
    }
}

__global__ void __launch_bounds__(256) k_cc_init(const int16_t *__restrict__ disp, int dpitch, size_t dframe, int W, int H, int *label, int *size)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (i >= W * H) return;
    const int y = i / W, x = i - y * W;
    const int d = disp[(size_t)f * dframe + (size_t)y * dpitch + x];
    label[(size_t)f * W * H + i] = (d == INVALID16) ? -1 : i;
    size[(size_t)f * W * H + i] = 0;
}

__global__ void __launch_bounds__(256) k_cc_merge(const int16_t *__restrict__ disp, int dpitch, size_t dframe, int W, int H, int maxdiff, int *label)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (i >= W * H) return;
    const int y = i / W, x = i - y * W;
    const int16_t *img = disp + (size_t)f * dframe;
    int *lab = label + (size_t)f * W * H;
    const int d = img[(size_t)y * dpitch + x];
    if (d == INVALID16) return;
    if (x + 1 < W) { const int e = img[(size_t)y * dpitch + x + 1]; if (e != INVALID16 && abs(e - d) <= maxdiff) uf_union(lab, i, i + 1); }
    if (y + 1 < H) { const int e = img[(size_t)(y + 1) * dpitch + x]; if (e != INVALID16 && abs(e - d) <= maxdiff) uf_union(lab, i, i + W); }
}

__global__ void __launch_bounds__(256) k_cc_count(int W, int H, int *label, int *size)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (i >= W * H) return;
    int *lab = label + (size_t)f * W * H;
    if (lab[i] < 0) return;
    const int r = uf_find(lab, i);
    lab[i] = r;                                             // path compression (roots never change any more)
    atomicAdd(&size[(size_t)f * W * H + r], 1);
}

__global__ void __launch_bounds__(256) k_cc_apply(int16_t *__restrict__ disp, int dpitch, size_t dframe, int W, int H, int max_size,
                                                  const int *label, const int *size)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (i >= W * H) return;
    const int r = label[(size_t)f * W * H + i];
    if (r < 0) return;
    if (size[(size_t)f * W * H + uf_find(label + (size_t)f * W * H, r)] <= max_size) {
        const int y = i / W, x = i - y * W;
        disp[(size_t)f * dframe + (size_t)y * dpitch + x] = (int16_t)INVALID16;
    }
}

int launch_postfilter(Img16 disp, const int16_t *cost, int W, int H, int n, int ndisp, int disp12_max_diff,
                      int speckle_window, int speckle_range, int *scratch, cudaStream_t s)
{
    int launches = 0;
    if (disp12_max_diff >= 0 && cost) {
        if (W * sizeof(unsigned long long) > 48 * 1024)          // rows wider than 6144 px need the opt-in shared-memory carve-out
            cudaFuncSetAttribute(k_validate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(W * sizeof(unsigned long long)));
        k_validate<<<dim3(H, n), 256, W * sizeof(unsigned long long), s>>>(disp.p, cost, disp.pitch, disp.frame, W, ndisp, disp12_max_diff * 16);
        launches++;
    }
    if (speckle_window > 0 && speckle_range >= 0 && scratch) {
        int *label = scratch, *size = scratch + (size_t)n * W * H;
        const dim3 grid((W * H + 255) / 256, n);
        k_cc_init<<<grid, 256, 0, s>>>(disp.p, disp.pitch, disp.frame, W, H, label, size);
        k_cc_merge<<<grid, 256, 0, s>>>(disp.p, disp.pitch, disp.frame, W, H, speckle_range, label);
        k_cc_count<<<grid, 256, 0, s>>>(W, H, label, size);
        k_cc_apply<<<grid, 256, 0, s>>>(disp.p, disp.pitch, disp.frame, W, H, speckle_window, label, size);
        launches += 4;
    }
    return launches;
}

}  // namespace u96
