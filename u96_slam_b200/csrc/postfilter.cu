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

// connected(p, q) of filterSpeckles: both valid and their 16x disparities differ by at most maxDiff
__device__ __forceinline__ bool cc_conn(int a, int b, int maxdiff) { return a != INVALID16 && b != INVALID16 && abs(a - b) <= maxdiff; }

// Pass 1, one CTA per image row: horizontal runs.  label[i] = pixel index of the first pixel of the run i belongs to (-1 for an
// invalid pixel).  Every thread owns K consecutive pixels; "the last cut (run start, or an invalid pixel = no open run) at or before
// x" is a prefix scan with the operator combine(a, b) = b has a cut ? b : a -- no atomics, one write per pixel.  size[] is cleared at
// run starts only (the only indices that can ever become roots).
__global__ void __launch_bounds__(256) k_cc_rows(const int16_t *__restrict__ disp, int dpitch, size_t dframe, int W, int H, int maxdiff,
                                                 int *__restrict__ label, int *__restrict__ size)
{
    __shared__ int s_cut[8], s_val[8];
    const int y = blockIdx.x, f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int16_t *row = disp + (size_t)f * dframe + (size_t)y * dpitch;
    int *lab = label + ((size_t)f * H + y) * W, *sz = size + ((size_t)f * H + y) * W;
    const int K = (W + 255) / 256;
    const int x0 = tid * K;
    const int before = (x0 > 0 && x0 - 1 < W) ? row[x0 - 1] : INVALID16;
    int hc = 0, v = -1, prev = before;
    for (int k = 0; k < K && x0 + k < W; k++) {
        const int d = row[x0 + k];
        if (d == INVALID16) { hc = 1; v = -1; }
        else if (!cc_conn(prev, d, maxdiff)) { hc = 1; v = x0 + k; }
        prev = d;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {                          // inclusive scan inside the warp
        const int phc = __shfl_up_sync(0xFFFFFFFFu, hc, o), pv = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= o && !hc) { hc = phc; v = pv; }
    }
    if (lane == 31) { s_cut[warp] = hc; s_val[warp] = v; }
    __syncthreads();
    int ehc = __shfl_up_sync(0xFFFFFFFFu, hc, 1), cur = __shfl_up_sync(0xFFFFFFFFu, v, 1);      // exclusive: threads before this one
    if (lane == 0) ehc = 0;
    if (!ehc) {
        cur = -1;
        for (int w = warp - 1; w >= 0; w--)
            if (s_cut[w]) { cur = s_val[w]; break; }
    }
    prev = before;
    for (int k = 0; k < K && x0 + k < W; k++) {
        const int x = x0 + k, d = row[x];
        if (d == INVALID16) cur = -1;
        else if (!cc_conn(prev, d, maxdiff)) { cur = x; sz[x] = 0; }
        lab[x] = (d == INVALID16) ? -1 : y * W + cur;
        prev = d;
    }
}

// Pass 2: vertical links.  Two runs of adjacent rows are united once, by the first pixel of their overlap that is vertically
// connected (a pixel whose left neighbour already linked the same two runs skips) -- unions per run pair, not per pixel.
__global__ void __launch_bounds__(256) k_cc_vmerge(const int16_t *__restrict__ disp, int dpitch, size_t dframe, int W, int H, int maxdiff, int *label)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (i >= W * H || i < W) return;
    const int y = i / W, x = i - y * W;
    const int16_t *r1 = disp + (size_t)f * dframe + (size_t)y * dpitch, *r0 = r1 - dpitch;
    const int d = r1[x], u = r0[x];
    if (!cc_conn(d, u, maxdiff)) return;
    if (x > 0) {
        const int dl = r1[x - 1], ul = r0[x - 1];
        if (cc_conn(dl, d, maxdiff) && cc_conn(ul, u, maxdiff) && cc_conn(dl, ul, maxdiff)) return;     // the same two runs, already linked
    }
    int *lab = label + (size_t)f * W * H;
    uf_union(lab, lab[i], lab[i - W]);
}

// Pass 3: component sizes, one atomicAdd per run (by its last pixel); run starts are compressed to their root on the way.
__global__ void __launch_bounds__(256) k_cc_count(const int16_t *__restrict__ disp, int dpitch, size_t dframe, int W, int H, int maxdiff,
                                                  int max_size, int *label, int *size)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (i >= W * H) return;
    const int y = i / W, x = i - y * W;
    const int16_t *row = disp + (size_t)f * dframe + (size_t)y * dpitch;
    const int d = row[x];
    if (d == INVALID16) return;
    if (x + 1 < W && cc_conn(d, row[x + 1], maxdiff)) return;   // not the last pixel of its run
    int *lab = label + (size_t)f * W * H;
    // first pixel of the run: every pixel but the first still holds it (only run starts are ever hooked under another root)
    const bool is_start = (x == 0) || !cc_conn(row[x - 1], d, maxdiff);
    const int s = is_start ? i : lab[i];
    const int r = uf_find(lab, s);
    if (r != s) lab[s] = r;                                     // path compression for pass 4 (roots never change any more)
    // only "at most max_size or more" matters: a component already known to be large takes no further atomics (the counter is
    // monotonic, a stale read merely adds once more) -- the big background components would otherwise serialise thousands of runs
    int *cnt = &size[(size_t)f * W * H + r];
    if (*reinterpret_cast<volatile int *>(cnt) <= max_size) atomicAdd(cnt, i - s + 1);
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
        k_cc_rows<<<dim3(H, n), 256, 0, s>>>(disp.p, disp.pitch, disp.frame, W, H, speckle_range, label, size);
        k_cc_vmerge<<<grid, 256, 0, s>>>(disp.p, disp.pitch, disp.frame, W, H, speckle_range, label);
        k_cc_count<<<grid, 256, 0, s>>>(disp.p, disp.pitch, disp.frame, W, H, speckle_range, speckle_window, label, size);
        k_cc_apply<<<grid, 256, 0, s>>>(disp.p, disp.pitch, disp.frame, W, H, speckle_window, label, size);
        launches += 4;
    }
    return launches;
}

}  // namespace u96
