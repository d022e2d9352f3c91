/*
 * u96_oracle.h -- CPU oracle for the U96-SLAM dense-stereo front end.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * arithmetic (FPGA RTL profile and cv::StereoBM profile).  It is the checker
 * for the CUDA path, never the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity pinning (see DESIGN.md "Oracle"):
 *   - x-Sobel (RTL):   pinned bit-exact by the reference's own golden vectors
 *                      data/ref_rect_{l,r} -> data/ref_xsbl_{l,r}.
 *   - rect_remap:      pinned against the reference's own C function compiled
 *                      from where it lies (oracle/_ref, fpga.c:303-366); that
 *                      build also regenerates the shipped src/dvp/sim/cmd.dat.
 *   - GFTT map:        PARITY UNPINNED (no eigen dump shipped; CORDIC sqrt IP
 *                      modelled as exact truncation); two readings agree.
 *   - diven closed forms: pinned against a bit-serial emulation of diven.v.
 *   - cv::StereoBM profile: pinned against cv2.StereoBM 4.13 (the third-party
 *                      library the reference calls; version unpinned upstream).
 *   - reprojection:    pinned against the reference's own Stereo.cpp /
 *                      StereoCameraModel.cpp / Transform.cpp compiled from where
 *                      they lie (oracle/_ref/libstereo_ref.so; OpenCV containers
 *                      stubbed, arithmetic untouched) + tests/golden fixture.
 *   - BM (RTL):        PARITY UNPINNED by reference assets (the reference ships
 *                      no disparity dump); cross-checked against the survey's
 *                      CRCs and an independent numpy reading (tests/rtl_bm_numpy.py).
 *
 * All citations are relative to /root/reference/.
 */
#ifndef U96_ORACLE_H
#define U96_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* RECT_PARAM (src/StereoBM/src/fpga.h:250-260); ch 0 = left, 1 = right. */
typedef struct {
    int32_t f[2][2];      /* [cam][x,y]  u10.16 focal length of the source camera        */
    int32_t c[2];         /* integer principal point of the source camera (shared)        */
    int32_t f2inv[2];     /* 2^32 / f'  (24 LSBs) of the rectified camera (shared)        */
    int32_t c2_f2[2];     /* c' / f'  u0.24 of the rectified camera (shared)              */
    int32_t rot[2][3][3]; /* [cam] inverse rectifying rotation, s0.24                     */
} orc_rect_params;

/* ---- diven.v:26-177, bit-serial emulation of the non-restoring divider ---- */
/* Returns the QW-bit quotient (unsigned container, two's complement inside).  */
uint64_t orc_diven(int DW, int VW, int QW, int MSB_INV, uint64_t dividend, uint64_t divisor);

/* ---- rectification: fpga.c:303-366 == rect_rmp.v:366-585 ------------------ */
/* Source coordinates (u10.5 / u9.5 fixed point) for every destination pixel. */
void orc_rect_remap(const orc_rect_params *p, int lr, int W, int H, int16_t *xs, int16_t *ys);
/* RTL-extended (no 16-bit wrap) variants for frames wider than 1023 or taller than 511 */
void orc_rect_remap32(const orc_rect_params *p, int lr, int W, int H, int32_t *xs, int32_t *ys);
void orc_rect_interp32(const uint8_t *src, int W, int H, int src_stride,
                       const int32_t *xs, const int32_t *ys, uint8_t *dst);
/* Bilinear interpolation with 5-bit fractions: rect_intp.v:288-412.
 * Out-of-image taps read 0 (the RTL leaves them undefined; build rule).       */
void orc_rect_interp(const uint8_t *src, int W, int H, int src_stride,
                     const int16_t *xs, const int16_t *ys, uint8_t *dst);

/* ---- x-Sobel prefilter ---------------------------------------------------- */
/* RTL: xsbl2.v:185-198, 682-698, 826-874.  clip [-32,31]+32; cols 0,W-1 = 32;
 * rows 0 and H-1 = 0 (never written).                                         */
void orc_xsobel_rtl(const uint8_t *src, int W, int H, uint8_t *dst);
/* cv::StereoBM prefilterXSobel. clip [-cap,cap]+cap.     */
void orc_xsobel_cv(const uint8_t *src, int W, int H, int cap, uint8_t *dst);

/* ---- block matching, RTL profile: bm*.v ---- */
typedef struct {
    int32_t wsz;            /* odd, 5 bit (bm.v:175)                             */
    int32_t ndisp;          /* multiple of 32, <= 256 (bm.v:176)                 */
    int32_t uni_enb, uni_mode, uni_thr;   /* bm.v:183-187                        */
    int32_t x_store_offset; /* 1 = what the DISP bank holds (bm_obuf2.v:125)     */
    int32_t rtl_extended;   /* 0 = sign-extend bit 15 like bm_obuf2.v:153        */
    int32_t bitserial_div;  /* 1 = use orc_diven for frac/uni, 0 = closed forms  */
} orc_bm_rtl_params;
/* xl/xr: 6-bit x-Sobel images (row stride W).  disp: W*H int16 (16x fixed).    */
int orc_bm_rtl(const uint8_t *xl, const uint8_t *xr, int W, int H,
               const orc_bm_rtl_params *p, int16_t *disp);
/* statistics side channel: number of 10-bit column-sum saturation events of
 * the last orc_bm_rtl call.                                    */
int64_t orc_bm_rtl_last_sat_events(void);

/* ---- block matching, cv::StereoBM profile --------- */
typedef struct {
    int32_t wsz, ndisp, prefilter_cap, texture_threshold, uniqueness_ratio;
} orc_bm_cv_params;
int orc_bm_cv(const uint8_t *pl, const uint8_t *pr, int W, int H,
              const orc_bm_cv_params *p, int16_t *disp);

int orc_bm_cv_cost(const uint8_t *pl, const uint8_t *pr, int W, int H,
                   const orc_bm_cv_params *p, int16_t *disp, int16_t *cost);
/* cv::StereoBM post filters (main.cpp:210-212): validateDisparity(disp12MaxDiff) then filterSpeckles */
void orc_validate_disparity(int16_t *disp, const int16_t *cost, int W, int H, int min_d, int ndisp, int disp12_max_diff);
void orc_filter_speckles(int16_t *img, int W, int H, int new_val, int max_size, int max_diff);

/* ---- disparity -> 3-D: slam/src/core/Stereo.cpp:157-182, main.cpp:522-551 - */
/* P_l, P_r: 3x4 row-major projection matrices.  decim = 1 or 4
 * (SensorData.cpp:50-58).  xyz: (W/decim)*(H/decim)*3 floats, NaN if d<=0.
 * apply_local: apply StereoCameraModel.cpp:9-14 localTransform.               */
void orc_reproject(const int16_t *disp, int W, int H, const double *P_l, const double *P_r,
                   int decim, int apply_local, float *xyz);
/* same with explicit 3x4 row-major float transforms (NULL = skip): localTransform then the optimised pose, each through
 * transformPoint (Stereo.cpp:189-198), exactly the dense consumer of main.cpp:522-551 */
void orc_reproject_ex(const int16_t *disp, int W, int H, const double *P_l, const double *P_r,
                      int decim, const float *local_T, const float *pose, float *xyz);
/* generateKeypoints3DStereo (Stereo.cpp:53-117), dense-map depth methods: gather at (int)y,(int)x, d<0 -> 0, skip 0,
 * projection with the float keypoint coordinates, min/max-depth gates, localTransform unless null.  NaN = bad point. */
void orc_reproject_points(const int16_t *disp, int W, int H, const double *P_l, const double *P_r,
                          const float *uv, int n, const uint8_t *mask, float min_depth, float max_depth,
                          const float *local_T, float *xyz);

/* UVC payload of the firmware (StereoBM/src/xusb_main.c:293-376): YUYV frame of 2W x H pixels.
 * mode 1 = USB_OUTPUT_STEREO_RECT / 2 = USB_OUTPUT_STEREO_XSBL (planar u8 L and R) / 3 = USB_OUTPUT_STEREO_BM (s16 disparity). */
void orc_pack_uvc(int mode, const uint8_t *L, const uint8_t *R, const int16_t *disp, int W, int H, uint8_t *frame);

/* ---- GFTT min-eigenvalue map: dvp/rtl/gftt{,_ibuf,_sbl,_eig,_box,_obuf}.v ----------------
 * src: rectified LEFT image (gftt.Address_In = BUF_RECT, fpga.c:166-167).  eig: W*H u16 (rows 0,1,H-2,H-1 = 0),
 * *max_out = per-frame maximum (gftt.Max register).  The CORDIC sqrt IP is modelled as an exact truncating
 * square root: PARITY UNPINNED (no eigen dump in the reference, IP model not runnable here).                */
void orc_gftt_eig(const uint8_t *src, int W, int H, int src_stride, uint16_t *eig, uint16_t *max_out);

#ifdef __cplusplus
}
#endif
#endif
