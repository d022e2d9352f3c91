// bm_fused.cu -- instantiation unit of the fused-role BM kernel (bm_fused.cuh).
#include "bm_fused.cuh"

namespace u96 {

bool bm_fused_ok(const BmConfig &c) { return bm_fused_supported(c); }
int launch_bm_fused_rtl64(const uint8_t *xl, const uint8_t *xr, int pitch, size_t frame, Img16 disp, const BmConfig &c, int n, cudaStream_t s)
{ return launch_bm_fused(xl, xr, pitch, frame, disp, c, n, s); }

}  // namespace u96
