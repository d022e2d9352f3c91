// TMA probe 2 (developer tool): descriptor location / data type / 1-D bulk copy variants.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: descriptor = __grid_constant__ param; 1: descriptor in global memory; 2: 1-D cp.async.bulk (no descriptor)
__global__ void k_probe(const __grid_constant__ CUtensorMap tm, const CUtensorMap *gtm, const uint8_t *src, uint8_t *out, int bytes,
                        int c0, int c1, int mode)
{
    extern __shared__ __align__(128) uint8_t sm[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 32768);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 32) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        if (mode == 2) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sm)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
        } else {
            const CUtensorMap *p = (mode == 1) ? gtm : &tm;
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(sm)), "l"(reinterpret_cast<uint64_t>(p)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
        }
    }
    if (threadIdx.x >= 32) return;
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bytes; i += 32) out[i] = sm[i];
}

typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const int mode = atoi(argv[1]), dtype = atoi(argv[2]);      // dtype 0 = u8, 1 = f32 (box in elements)
    const int es_bytes = dtype ? 4 : 1;
    const int W = 640 / es_bytes, H = 480, pitch = 640, BW = 128 / es_bytes, BH = 16;
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    PFN enc = (PFN)p;
    int drv = 0, rtv = 0; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rtv);
    uint8_t *d, *o; CUtensorMap *gtm;
    cudaMalloc(&d, (size_t)pitch * H); cudaMalloc(&o, 65536); cudaMalloc(&gtm, sizeof(CUtensorMap));
    std::vector<uint8_t> h((size_t)pitch * H);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 9));
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    cuuint64_t strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)BH};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, dtype ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaMemcpy(gtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    printf("mode=%d dtype=%d drv=%d rt=%d getentry=%d q=%d encode=%d ", mode, dtype, drv, rtv, (int)ge, (int)q, (int)r);
    const int c0 = 32 / es_bytes, c1 = 8;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 64);
    const int bytes = (mode == 2) ? 2048 : 128 * BH;
    k_probe<<<1, 64, 32768 + 64>>>(tm, gtm, d + 1280, o, bytes, c0, c1, mode);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<uint8_t> ho(bytes);
        cudaMemcpy(ho.data(), o, bytes, cudaMemcpyDeviceToHost);
        int bad = 0;
        if (mode == 2) { for (int i = 0; i < bytes; i++) bad += ho[i] != h[1280 + i]; }
        else for (int y = 0; y < BH; y++) for (int x = 0; x < 128; x++) bad += ho[y * 128 + x] != h[(size_t)(c1 + y) * pitch + 32 + x];
        printf("mismatch=%d", bad);
    }
    printf("\n");
    return 0;
}
