// microbench.cu -- issue-rate micro-kernels that give the INT / shared-memory roofline
// denominators for the BM kernel (MEASURED_PEAKS.json only has HBM and bf16 numbers).
// Every thread runs ILP independent dependency chains of one instruction kind; the result is
// lane-ops per second over the whole chip (or bytes/s for LDS).
#include "common.cuh"

namespace u96 {

constexpr int MB_ILP = 8;
constexpr int MB_ITERS = 4096;
constexpr int MB_THREADS = 512;

template <int WHICH>
__global__ void __launch_bounds__(MB_THREADS) k_mb(uint32_t *out, uint32_t seed)
{
    __shared__ uint4 sm[MB_THREADS];
    uint32_t v[MB_ILP];
#pragma unroll
    for (int i = 0; i < MB_ILP; i++) v[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u;
    uint32_t b = seed | 0x00010001u, c = (seed >> 3) | 1u;
    sm[threadIdx.x] = make_uint4(v[0], v[1], v[2], v[3]);
    __syncthreads();
    int lane_src = (threadIdx.x + 1) & 31;
    for (int it = 0; it < MB_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < MB_ILP; i++) {
            if (WHICH == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(b));
            if (WHICH == 1) v[i] = __vabsdiffu4(v[i], b);
            if (WHICH == 2) v[i] = __viaddmin_u16x2(v[i], b, 0x03FF03FFu);
            if (WHICH == 3) v[i] = __vimin3_u32(v[i], b, c + i);
            if (WHICH == 4) v[i] = __byte_perm(v[i], b, 0x4140 + i);
            if (WHICH == 5) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(b), "r"(c));
            if (WHICH == 7) v[i] = __shfl_sync(0xFFFFFFFFu, v[i], lane_src);
            if (WHICH == 8) v[i] = __vminu2(v[i], b + i);
            if (WHICH == 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(b), "r"(c));
        }
        if (WHICH == 6) {
#pragma unroll
            for (int i = 0; i < MB_ILP / 4; i++) {
                const uint4 t = sm[(threadIdx.x + it + 64 * i) & (MB_THREADS - 1)];
                v[4 * i] ^= t.x; v[4 * i + 1] ^= t.y; v[4 * i + 2] ^= t.z; v[4 * i + 3] ^= t.w;
            }
        }
        if (WHICH == 1 || WHICH == 2 || WHICH == 3 || WHICH == 8) b += 0x00010001u;   // keep operands changing
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < MB_ILP; i++) acc ^= v[i];
    if (acc == 0x12345678u) out[blockIdx.x] = acc;    // practically never; keeps the chains live
}

template <int WHICH>
static float mb_run(uint32_t *dout, int blocks)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mb<WHICH><<<blocks, MB_THREADS>>>(dout, 12345u);     // warm-up
    cudaEventRecord(e0);
    k_mb<WHICH><<<blocks, MB_THREADS>>>(dout, 6789u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
}

int run_microbench(int which, double *gops)
{
    cudaDeviceProp prop;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return U96_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 4;     // 2048 threads/SM resident
    uint32_t *dout = nullptr;
    if (cudaMalloc(&dout, blocks * sizeof(uint32_t)) != cudaSuccess) return U96_ERR_NOMEM;
    float ms = 0;
    switch (which) {
    case 0: ms = mb_run<0>(dout, blocks); break;
    case 1: ms = mb_run<1>(dout, blocks); break;
    case 2: ms = mb_run<2>(dout, blocks); break;
    case 3: ms = mb_run<3>(dout, blocks); break;
    case 4: ms = mb_run<4>(dout, blocks); break;
    case 5: ms = mb_run<5>(dout, blocks); break;
    case 6: ms = mb_run<6>(dout, blocks); break;
    case 7: ms = mb_run<7>(dout, blocks); break;
    case 8: ms = mb_run<8>(dout, blocks); break;
    case 9: ms = mb_run<9>(dout, blocks); break;
    default: cudaFree(dout); return U96_ERR_INVALID;
    }
    cudaFree(dout);
    if (cudaGetLastError() != cudaSuccess || ms <= 0) return U96_ERR_CUDA;
    double per_thread = (double)MB_ITERS * MB_ILP;                   // lane-ops
    if (which == 6) per_thread = (double)MB_ITERS * (MB_ILP / 4) * 16.0;   // bytes
    *gops = per_thread * MB_THREADS * blocks / (ms * 1e-3) / 1e9;
    return U96_OK;
}

}  // namespace u96
