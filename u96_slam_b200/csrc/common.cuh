// common.cuh -- shared declarations for the libu96stereo kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/u96_stereo.h"

namespace u96 {

// Device image: n frames back to back, row pitch in bytes (multiple of 128).
struct Img8 {
    uint8_t *p;
    int pitch;           // bytes per row
    size_t frame;        // bytes per frame (pitch * H)
};
struct Img16 {
    int16_t *p;
    int pitch;           // elements per row
    size_t frame;        // elements per frame
};

static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// ---- stage launchers (each returns the number of kernels it launched) ----
struct RectMapParams { u96_rect_params p; int W, H; int wrap16; };

int launch_rect_build_map(const RectMapParams &rp, int2 *map /*[2][H][W]*/, cudaStream_t s);
// per-parameter-set plan of the TMA remap path: source bounding boxes of the destination tiles
struct RectPlan { bool tma = false; int BW = 0, BH = 0, stage_bytes = 0, tiles_x = 0, tiles_y = 0; int4 *d_tiles = nullptr; };
int rect_plan_build(RectPlan &pl, const int2 *map, int W, int H, cudaStream_t s);
void rect_plan_free(RectPlan &pl);
int launch_rect_remap(const uint8_t *srcL, const uint8_t *srcR, int src_pitch, size_t src_frame,
                      Img8 dstL, Img8 dstR, const int2 *map, const RectPlan &pl, int W, int H, int n, cudaStream_t s);
int launch_xsobel(const uint8_t *srcL, const uint8_t *srcR, int src_pitch, size_t src_frame,
                  Img8 dstL, Img8 dstR, int W, int H, int n, int profile, int cap, cudaStream_t s);

struct BmConfig {
    int W, H, D, wsz, profile;
    int uni_enable, uni_mode, uni_thr, x_store_offset, rtl_extended;   // RTL
    int cap, tex_thr, uniq;                                            // OPENCV
    int16_t *cost = nullptr;                                           // OPENCV: winning SAD per valid pixel (same pitch as disp) or null
    void *sat_scratch = nullptr; size_t sat_scratch_bytes = 0;        // RTL, a handful of pairs: band functions / states of the saturating chain
};
size_t bm_sat_scratch_bytes(const BmConfig &c, int n);                 // scratch launch_bm wants for n pairs of this configuration (0: none)
int  bm_smem_bytes(const BmConfig &c);
int  bm_wave_frames(const BmConfig &c);
int launch_bm(const uint8_t *xl, const uint8_t *xr, int pitch, size_t frame, Img16 disp,
              const BmConfig &c, int n, cudaStream_t s);
// invalid code into everything outside the valid rectangle of n frames (the BM kernels never write there)
int launch_bm_border(Img16 disp, const BmConfig &c, int n, cudaStream_t s);

bool bm_fast_supported(const BmConfig &c);
int launch_bm_fast(const uint8_t *xl, const uint8_t *xr, int pitch, size_t frame, Img16 disp,
                   const BmConfig &c, int n, cudaStream_t s);
int launch_bm_fast_cs1(const uint8_t *xl, const uint8_t *xr, int pitch, size_t frame, Img16 disp, const BmConfig &c, int n, cudaStream_t s);
int launch_bm_fast_cs2(const uint8_t *xl, const uint8_t *xr, int pitch, size_t frame, Img16 disp, const BmConfig &c, int n, cudaStream_t s);
// fused-role kernel (bm_fused.cuh): RTL profile, 64 / 128 / 256 disparities, uniqueness off
bool bm_fused_ok(const BmConfig &c);
bool bm_fused_preferred(const BmConfig &c, bool sat);
int launch_bm_fused_rtl(const uint8_t *xl, const uint8_t *xr, int pitch, size_t frame, Img16 disp, const BmConfig &c, int n, cudaStream_t s);
int launch_bm_fast_cs4(const uint8_t *xl, const uint8_t *xr, int pitch, size_t frame, Img16 disp, const BmConfig &c, int n, cudaStream_t s);

// cv::StereoBM post filters (OPENCV profile): validateDisparity then filterSpeckles; scratch = 2 int32 per pixel of the batch
int launch_postfilter(Img16 disp, const int16_t *cost, int W, int H, int n, int ndisp, int disp12_max_diff,
                      int speckle_window, int speckle_range, int *scratch, cudaStream_t s);

// local_T: 3x4 row-major float transform (host; null or all zero = none); d_poses: n x 12 floats in device memory or null
int launch_reproject(const int16_t *disp, int dpitch, size_t dframe, int W, int H, int n,
                     const double *P_l, const double *P_r, int decim, const float *local_T, const float *d_poses,
                     float *xyz, cudaStream_t s);
// generateKeypoints3DStereo on one frame's disparity map; d_uv = n float pairs (x, y), d_mask = n bytes or null (device)
int launch_reproject_points(const int16_t *disp, int dpitch, int W, int H, const double *P_l, const double *P_r,
                            const float *d_uv, const uint8_t *d_mask, int n, float min_depth, float max_depth,
                            const float *local_T, float *xyz, cudaStream_t s);

// UVC payload (xusb_main.c:293-376): YUYV frame of 2W x H; disp != null selects the disparity mode
int launch_pack_uvc(const uint8_t *srcL, const uint8_t *srcR, int sp, size_t sf, const int16_t *disp, int dp, size_t df,
                    uint8_t *out, int W, int H, int n, cudaStream_t s);

// GFTT min-eigenvalue map (dvp/rtl/gftt*.v) of one image per frame: u16 map + per-frame maximum; returns kernels launched
int launch_gftt(const uint8_t *src, int sp, size_t sf, uint16_t *eig, int ep, size_t ef, uint32_t *fmax,
                int W, int H, int n, cudaStream_t s);

int run_microbench(int which, double *gops);

}  // namespace u96
