// bm_fused.cu -- instantiation unit of the fused-role BM kernel (bm_fused.cuh).
#include "bm_fused.cuh"

namespace u96 {

bool bm_fused_ok(const BmConfig &c) { return bm_fused_supported(c); }
// where the fused kernel is the faster one (profiles/r02_summary.md)
// measured (profiles/r02_summary.md): the fused-role kernel is the faster one in every configuration it supports since the saturating
// variants stage 16-bit rows (C2: 2.20 ms against 2.33 ms of k_bm_fast)
bool bm_fused_preferred(const BmConfig &c, bool sat) { (void)c; (void)sat; return true; }
size_t bm_sat_scratch_bytes(const BmConfig &c, int n) { return bm_fused_supported(c) ? fused_sat_scratch_bytes(c, n) : 0; }
int launch_bm_fused_rtl(const uint8_t *xl, const uint8_t *xr, int pitch, size_t frame, Img16 disp, const BmConfig &c, int n, cudaStream_t s)
{ return launch_bm_fused(xl, xr, pitch, frame, disp, c, n, s); }

}  // namespace u96
