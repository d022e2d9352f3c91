// TMA probe (developer tool): which cp.async.bulk.tensor configuration is accepted on this box.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void k_probe(const __grid_constant__ CUtensorMap tm, uint8_t *out, int bytes, int c0, int c1, int c2, int exit_others)
{
    extern __shared__ __align__(128) uint8_t sm[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 32768);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= 32) {
        if (exit_others && threadIdx.x != 32) return;
        if (threadIdx.x == 32) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
            if (RANK == 3)
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(smem_u32(sm)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(sm)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
        }
        return;
    }
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bytes; i += 32) out[i] = sm[i];
}

typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const int rank = atoi(argv[1]), BW = atoi(argv[2]), BH = atoi(argv[3]), exit_others = atoi(argv[4]), l2 = atoi(argv[5]);
    const int W = 640, H = 480, n = 4, pitch = 640;
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    PFN enc = (PFN)p;
    uint8_t *d, *o;
    cudaMalloc(&d, (size_t)pitch * H * n); cudaMalloc(&o, 65536);
    std::vector<uint8_t> h((size_t)pitch * H * n);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 9));
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * H};
    cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, 1u};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("rank=%d box=%dx%d exit_others=%d l2=%d encode=%d ", rank, BW, BH, exit_others, l2, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 0; }
    const int c0 = argc > 6 ? atoi(argv[6]) : 29, c1 = 8, c2 = 1;
    cudaFuncSetAttribute(k_probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 64);
    cudaFuncSetAttribute(k_probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 64);
    if (rank == 3) k_probe<3><<<1, 64, 32768 + 64>>>(tm, o, BW * BH, c0, c1, c2, exit_others);
    else k_probe<2><<<1, 64, 32768 + 64>>>(tm, o, BW * BH, c0, c1, c2, exit_others);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<uint8_t> ho(BW * BH);
        cudaMemcpy(ho.data(), o, BW * BH, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < BH; y++)
            for (int x = 0; x < BW; x++) {
                const int sx = c0 + x, sy = c1 + y;
                const uint8_t want = (sx >= 0 && sx < W && sy < H) ? h[(size_t)(rank == 3 ? c2 : 0) * pitch * H + (size_t)sy * pitch + sx] : 0;
                bad += ho[y * BW + x] != want;
            }
        printf("mismatch=%d", bad);
    }
    printf("\n");
    return 0;
}
