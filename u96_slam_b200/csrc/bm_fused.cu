// bm_fused.cu -- instantiation unit of the fused-role BM kernel (bm_fused.cuh).
#include "bm_fused.cuh"

namespace u96 {

bool bm_fused_ok(const BmConfig &c) { return bm_fused_supported(c); }
// where the fused kernel is the faster one (profiles/r02_summary.md)
bool bm_fused_preferred(const BmConfig &c, bool sat) { return c.profile == U96_PROFILE_OPENCV || !sat || c.D > 64; }
int launch_bm_fused_rtl(const uint8_t *xl, const uint8_t *xr, int pitch, size_t frame, Img16 disp, const BmConfig &c, int n, cudaStream_t s)
{ return launch_bm_fused(xl, xr, pitch, frame, disp, c, n, s); }

}  // namespace u96
