// Which SM sub-partition does a warp land on?  Launches CTAs shaped like k_bm_fused<.,.,8> (224 threads, ~55 KB dynamic shared memory,
// 4 per SM) and records %smid, %warpid and the CTA-local warp index of every warp (developer probe, run under gpurun).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(224, 4) k(int *out, int spin)
{
    extern __shared__ int sm[];
    unsigned smid, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    sm[threadIdx.x] = threadIdx.x;
    long long t0 = clock64();
    while (clock64() - t0 < spin) { }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        int *o = out + (blockIdx.x * 8 + (threadIdx.x >> 5)) * 2;
        o[0] = (int)smid; o[1] = (int)wid;
    }
}
int main()
{
    const int nb = 148 * 8, smem = 55 * 1024;
    int *d; cudaMalloc(&d, nb * 8 * 2 * sizeof(int)); cudaMemset(d, 0xFF, nb * 8 * 2 * sizeof(int));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<<<nb, 224, smem>>>(d, 200000);
    cudaDeviceSynchronize();
    static int h[148 * 8 * 8 * 2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    for (int b = 0; b < nb; b++) {
        if (h[b * 16] > 1) continue;                       // SMs 0 and 1 only
        printf("block %4d sm %d warpids:", b, h[b * 16]);
        for (int w = 0; w < 7; w++) printf(" %2d", h[(b * 8 + w) * 2 + 1]);
        printf("   (mod 4:");
        for (int w = 0; w < 7; w++) printf(" %d", h[(b * 8 + w) * 2 + 1] & 3);
        printf(")\n");
    }
    return 0;
}
