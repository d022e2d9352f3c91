// bm_fast_cs4.cu -- instantiations of the fast BM kernel for numDisparities = 256 (cluster of 4 CTAs); see bm_fast.cuh.
#include "bm_fast.cuh"

namespace u96 {

int launch_bm_fast_cs4(const uint8_t *xl, const uint8_t *xr, int pitch, size_t frame, Img16 disp,
                        const BmConfig &c, int n, cudaStream_t s)
{
    return launch_bm_fast_cs<4>(xl, xr, pitch, frame, disp, c, n, s);
}

}  // namespace u96
