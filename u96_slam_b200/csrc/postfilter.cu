// postfilter.cu -- the post filters cv::StereoBM::compute applies as configured by the reference's CPU mode
// (slam/src/core/main.cpp:210-212: disp12MaxDiff 1, speckleWindowSize 50, speckleRange 32), sm_100a.
//
//   k_validate  : cv::validateDisparity.  Per image row: the right-image disparity is the left disparity of the
//                 cheapest pixel that maps onto it (first pixel wins ties) -- a shared-memory atomicMin on the
//                 key (cost as signed short, biased)<<16 | d -- then a left pixel is dropped when BOTH roundings of its disparity
//                 hit a right pixel whose disparity differs by more than disp12MaxDiff.
//   k_cc_*      : cv::filterSpeckles.  4-connected components of the relation |d(p)-d(q)| <= maxDiff over valid
//                 pixels by lock-free union-find (roots = smallest pixel index, so labels are deterministic),
//                 component sizes by atomicAdd, components of at most maxSpeckleSize pixels become invalid.
//                 The result of the CPU flood fill does not depend on its traversal order, so the two agree.
#include <algorithm>

#include "common.cuh"

namespace u96 {

constexpr int INVALID16 = -16;          // (minDisparity - 1) * 16, minDisparity == 0

// 8 disparities / costs of one row as 16 bytes
__device__ __forceinline__ void ld8(const int16_t *p, int (&v)[8])
{
    const uint4 t = *reinterpret_cast<const uint4 *>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int k = 0; k < 4; k++) { v[2 * k] = (int16_t)(w[k] & 0xFFFF); v[2 * k + 1] = (int16_t)(w[k] >> 16); }
}

// Key of the right-image slot x2: (cost as signed short, biased) << 16 | d.  Among the pixels that map onto one slot the cheapest wins and
// the first one wins ties (cv::validateDisparity scans x upwards and replaces on strictly smaller cost): x = x2 + ((d + 8) >> 4), so the
// smaller x is the smaller d -- the minimum of the 32-bit key is exactly that pixel, and its low half is the disparity the check needs.
__global__ void __launch_bounds__(256) k_validate(int16_t *__restrict__ disp, const int16_t *__restrict__ cost, int dpitch, size_t dframe,
                                                  int W, int ndisp, int maxdiff16)
{
    extern __shared__ unsigned int s_key[];                // [W]
    const int y = blockIdx.x, f = blockIdx.y;
    int16_t *drow = disp + (size_t)f * dframe + (size_t)y * dpitch;
    const int16_t *crow = cost + (size_t)f * dframe + (size_t)y * dpitch;
    for (int x = threadIdx.x; x < W; x += blockDim.x) s_key[x] = ~0u;
    __syncthreads();
    const int minX1 = ndisp;                                // max(minD + ndisp, 0), minD == 0 ; maxX1 = W
    // 8 pixels per thread and trip (the pitch is a multiple of 64 elements: groups past W read padding and are masked by x < W)
    for (int x8 = 8 * threadIdx.x; x8 < W; x8 += 8 * blockDim.x) {
        if (x8 + 8 <= minX1) continue;
        int d[8], c[8];
        ld8(drow + x8, d); ld8(crow + x8, c);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int x = x8 + k;
            if (x < minX1 || x >= W || d[k] == INVALID16) continue;
            const int x2 = x - ((d[k] + 8) >> 4);
            if (x2 < 0 || x2 >= W) continue;                // cannot happen for 0 <= d < 16*ndisp
            atomicMin(&s_key[x2], ((unsigned)(c[k] + 32768) << 16) | (unsigned)(unsigned short)d[k]);
        }
    }
    __syncthreads();
    for (int x8 = 8 * threadIdx.x; x8 < W; x8 += 8 * blockDim.x) {
        if (x8 + 8 <= minX1) continue;
        int d[8];
        ld8(drow + x8, d);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int x = x8 + k;
            if (x < minX1 || x >= W || d[k] == INVALID16) continue;
            const int x0 = x - (d[k] >> 4), x1 = x - ((d[k] + 15) >> 4);
            bool bad0 = false, bad1 = false;
            if (x0 >= 0 && x0 < W) { const unsigned key = s_key[x0]; bad0 = (key != ~0u) && (abs((int)(short)(key & 0xFFFF) - d[k]) > maxdiff16); }
            if (x1 >= 0 && x1 < W) { const unsigned key = s_key[x1]; bad1 = (key != ~0u) && (abs((int)(short)(key & 0xFFFF) - d[k]) > maxdiff16); }
            if (bad0 && bad1) drow[x] = (int16_t)INVALID16;
        }
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
// x" is a prefix scan with the operator combine(a, b) = b has a cut ? b : a -- no atomics, one write per pixel.  size[] = the LENGTH
// of the run at its first pixel (written by whoever sees the run end) and 0 everywhere else: passes 3 and 4 never look at the
// disparities again, a run start is "size > 0", and a root's own length is already in its counter.
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
        const bool inv = (d == INVALID16);
        const bool cut = inv || !cc_conn(prev, d, maxdiff);
        if (cut) {
            if (cur >= 0) sz[cur] = x - cur;                    // the open run ends here (it may have started in another thread's pixels)
            cur = inv ? -1 : x;
        }
        if (inv || !cut) sz[x] = 0;                             // not a run start
        lab[x] = inv ? -1 : y * W + cur;
        prev = d;
        if (x == W - 1 && cur >= 0) sz[cur] = W - cur;          // the row ends inside a run
    }
}

// Pass 1 for rows of up to 2048 pixels whose width is a multiple of 8: thread = 8 consecutive pixels, read once as 16 bytes and kept in
// registers for both sweeps; labels and the zeroed sizes leave as 16-byte stores, the run lengths are written behind a barrier
// (a run start's slot is zeroed by the thread that owns the pixel and filled by the thread that sees the run end).
__global__ void __launch_bounds__(256) k_cc_rows8(const int16_t *__restrict__ disp, int dpitch, size_t dframe, int W, int H, int maxdiff,
                                                  int *__restrict__ label, int *__restrict__ size)
{
    __shared__ int s_cut[8], s_val[8];
    const int y = blockIdx.x, f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int16_t *row = disp + (size_t)f * dframe + (size_t)y * dpitch;
    int *lab = label + ((size_t)f * H + y) * W, *sz = size + ((size_t)f * H + y) * W;
    const int x0 = 8 * tid;
    const bool active = x0 < W;
    int d[8];
    if (active) ld8(row + x0, d);
    else {
#pragma unroll
        for (int k = 0; k < 8; k++) d[k] = INVALID16;
    }
    const int before = (active && x0 > 0) ? row[x0 - 1] : INVALID16;
    int hc = 0, v = -1, prev = before;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (d[k] == INVALID16) { hc = 1; v = -1; }
        else if (!cc_conn(prev, d[k], maxdiff)) { hc = 1; v = x0 + k; }
        prev = d[k];
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {                          // inclusive scan inside the warp
        const int phc = __shfl_up_sync(0xFFFFFFFFu, hc, o), pv = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= o && !hc) { hc = phc; v = pv; }
    }
    if (lane == 31) { s_cut[warp] = hc; s_val[warp] = v; }
    __syncthreads();
    int ehc = __shfl_up_sync(0xFFFFFFFFu, hc, 1), cur0 = __shfl_up_sync(0xFFFFFFFFu, v, 1);     // exclusive: threads before this one
    if (lane == 0) ehc = 0;
    if (!ehc) {
        cur0 = -1;
        for (int w = warp - 1; w >= 0; w--)
            if (s_cut[w]) { cur0 = s_val[w]; break; }
    }
    if (active) {
        int l8[8], cur = cur0;
        prev = before;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const bool inv = (d[k] == INVALID16);
            if (inv) cur = -1;
            else if (!cc_conn(prev, d[k], maxdiff)) cur = x0 + k;
            l8[k] = inv ? -1 : y * W + cur;
            prev = d[k];
        }
        *reinterpret_cast<int4 *>(lab + x0) = make_int4(l8[0], l8[1], l8[2], l8[3]);
        *reinterpret_cast<int4 *>(lab + x0 + 4) = make_int4(l8[4], l8[5], l8[6], l8[7]);
        *reinterpret_cast<int4 *>(sz + x0) = make_int4(0, 0, 0, 0);
        *reinterpret_cast<int4 *>(sz + x0 + 4) = make_int4(0, 0, 0, 0);
    }
    __syncthreads();
    if (active) {
        int cur = cur0;
        prev = before;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int x = x0 + k;
            const bool inv = (d[k] == INVALID16);
            if (inv || !cc_conn(prev, d[k], maxdiff)) {
                if (cur >= 0) sz[cur] = x - cur;                // the open run ends here
                cur = inv ? -1 : x;
            }
            prev = d[k];
        }
        if (x0 + 8 == W && cur >= 0) sz[cur] = W - cur;         // the row ends inside a run
    }
}

// G consecutive elements of a per-frame array (G = 4: one 16-byte / 8-byte access, the row width is a multiple of 4; G = 1 otherwise)
template <int G> struct Vec;
template <> struct Vec<4> {
    static __device__ __forceinline__ void ld(const int *p, int (&v)[4]) { const int4 t = *reinterpret_cast<const int4 *>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void ld(const int16_t *p, int (&v)[4])
    { const uint2 t = *reinterpret_cast<const uint2 *>(p); v[0] = (int16_t)(t.x & 0xFFFF); v[1] = (int16_t)(t.x >> 16); v[2] = (int16_t)(t.y & 0xFFFF); v[3] = (int16_t)(t.y >> 16); }
};
template <> struct Vec<1> {
    static __device__ __forceinline__ void ld(const int *p, int (&v)[1]) { v[0] = *p; }
    static __device__ __forceinline__ void ld(const int16_t *p, int (&v)[1]) { v[0] = *p; }
};

// Pass 2: vertical links.  Two runs of adjacent rows are united once, by the first pixel of their overlap that is vertically
// connected (a pixel whose left neighbour already linked the same two runs skips) -- unions per run pair, not per pixel.
template <int G>
__global__ void __launch_bounds__(256) k_cc_vmerge(const int16_t *__restrict__ disp, int dpitch, size_t dframe, int W, int H, int maxdiff, int *label)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) * G, f = blockIdx.y;
    if (i >= W * H || i < W) return;
    const int y = i / W, x = i - y * W;
    const int16_t *r1 = disp + (size_t)f * dframe + (size_t)y * dpitch, *r0 = r1 - dpitch;
    int d[G], u[G];
    Vec<G>::ld(r1 + x, d); Vec<G>::ld(r0 + x, u);
    int dl = INVALID16, ul = INVALID16;
    bool any = false;
#pragma unroll
    for (int k = 0; k < G; k++) any = any || cc_conn(d[k], u[k], maxdiff);
    if (!any) return;
    if (x > 0) { dl = r1[x - 1]; ul = r0[x - 1]; }
    int *lab = label + (size_t)f * W * H;
#pragma unroll
    for (int k = 0; k < G; k++) {
        if (cc_conn(d[k], u[k], maxdiff) &&
            !(cc_conn(dl, d[k], maxdiff) && cc_conn(ul, u[k], maxdiff) && cc_conn(dl, ul, maxdiff)))        // (not: the same two runs, already linked)
            uf_union(lab, lab[i + k], lab[i + k - W]);
        dl = d[k]; ul = u[k];
    }
}

// Pass 3: component sizes.  A run start (size > 0) that is not its component's root adds its length to the root's counter, which
// already holds the root run's own length; run starts are compressed to their root on the way.
template <int G>
__global__ void __launch_bounds__(256) k_cc_count(int W, int H, int max_size, int *label, int *size)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) * G, f = blockIdx.y;
    if (i >= W * H) return;
    int *lab = label + (size_t)f * W * H, *sz = size + (size_t)f * W * H;
    int len[G];
    Vec<G>::ld(sz + i, len);
#pragma unroll
    for (int k = 0; k < G; k++) {
        if (len[k] <= 0) continue;
        const int s = i + k;
        const int r = uf_find(lab, s);
        if (r == s) continue;                                   // (a root's counter may already have grown: its value is not a length any more, and not needed)
        lab[s] = r;                                             // path compression for pass 4 (roots never change any more)
        // only "at most max_size or more" matters: a component already known to be large takes no further atomics (the counter is
        // monotonic, a stale read merely adds once more) -- the big background components would otherwise serialise thousands of runs
        int *cnt = &sz[r];
        if (*reinterpret_cast<volatile int *>(cnt) <= max_size) atomicAdd(cnt, len[k]);
    }
}

// Pass 4: pixels of components of at most max_size pixels become invalid; neighbouring pixels of one run share the look-up
template <int G>
__global__ void __launch_bounds__(256) k_cc_apply(int16_t *__restrict__ disp, int dpitch, size_t dframe, int W, int H, int max_size,
                                                  const int *label, const int *size)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) * G, f = blockIdx.y;
    if (i >= W * H) return;
    const int *lab = label + (size_t)f * W * H, *sz = size + (size_t)f * W * H;
    int r[G];
    Vec<G>::ld(lab + i, r);
    const int y = i / W, x = i - y * W;
    int16_t *out = disp + (size_t)f * dframe + (size_t)y * dpitch + x;
    int last = -2;
    bool small = false;
#pragma unroll
    for (int k = 0; k < G; k++) {
        if (r[k] < 0) continue;
        if (r[k] != last) { last = r[k]; small = sz[uf_find(lab, r[k])] <= max_size; }
        if (small) out[k] = (int16_t)INVALID16;
    }
}

int launch_postfilter(Img16 disp, const int16_t *cost, int W, int H, int n, int ndisp, int disp12_max_diff,
                      int speckle_window, int speckle_range, int *scratch, cudaStream_t s)
{
    int launches = 0;
    if (disp12_max_diff >= 0 && cost) {
        if (W * sizeof(unsigned int) > 48 * 1024)                // rows wider than 12288 px need the opt-in shared-memory carve-out
            cudaFuncSetAttribute(k_validate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(W * sizeof(unsigned int)));
        const int vthreads = std::min(256, ((W + 7) / 8 + 31) / 32 * 32);
        k_validate<<<dim3(H, n), vthreads, W * sizeof(unsigned int), s>>>(disp.p, cost, disp.pitch, disp.frame, W, ndisp, disp12_max_diff * 16);
        launches++;
    }
    if (speckle_window > 0 && speckle_range >= 0 && scratch) {
        int *label = scratch, *size = scratch + (size_t)n * W * H;
        if (W % 8 == 0 && W <= 2048) k_cc_rows8<<<dim3(H, n), (W / 8 + 31) / 32 * 32, 0, s>>>(disp.p, disp.pitch, disp.frame, W, H, speckle_range, label, size);
        else k_cc_rows<<<dim3(H, n), 256, 0, s>>>(disp.p, disp.pitch, disp.frame, W, H, speckle_range, label, size);
        if (W % 4 == 0) {                                       // four pixels per thread: 16-byte label / size accesses, 8-byte disparity loads
            const dim3 grid((W * H / 4 + 255) / 256, n);
            k_cc_vmerge<4><<<grid, 256, 0, s>>>(disp.p, disp.pitch, disp.frame, W, H, speckle_range, label);
            k_cc_count<4><<<grid, 256, 0, s>>>(W, H, speckle_window, label, size);
            k_cc_apply<4><<<grid, 256, 0, s>>>(disp.p, disp.pitch, disp.frame, W, H, speckle_window, label, size);
        } else {
            const dim3 grid((W * H + 255) / 256, n);
            k_cc_vmerge<1><<<grid, 256, 0, s>>>(disp.p, disp.pitch, disp.frame, W, H, speckle_range, label);
            k_cc_count<1><<<grid, 256, 0, s>>>(W, H, speckle_window, label, size);
            k_cc_apply<1><<<grid, 256, 0, s>>>(disp.p, disp.pitch, disp.frame, W, H, speckle_window, label, size);
        }
        launches += 4;
    }
    return launches;
}

}  // namespace u96
