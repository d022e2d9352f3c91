/*
 * u96_stereo.h -- C ABI of libu96stereo, the B200 drop-in for the U96-SLAM
 * dense-stereo front end (rect -> x-Sobel -> SAD block matching -> disparity
 * -> 3-D).  It stands where the FPGA bitstream + DDR banks + mailbox handshake
 * stand in the reference; every entry point cites the reference interface it
 * replaces (paths relative to the reference repository root).
 *
 * Conventions (mirroring slam/include/core/FPGA.h:347-397):
 *   - one handle per GPU, used from one thread at a time (like class Fpga);
 *   - the library owns two banks (A=0, B=1) per buffer; submit_* fills a bank
 *     asynchronously, wait() returns the bank that became ready, receive_*
 *     COPIES OUT (the caller owns the copy, like cv::Mat::clone());
 *   - every function returns 0 or a negative U96_ERR_* code; the library
 *     never exits and never spins.
 *   - there is no CPU fallback: without a CUDA device u96_create fails.
 *   - like the FPGA's register file, the configuration (u96_set_bm_params, u96_set_bm_registers,
 *     u96_set_rect_params, u96_set_stream) may only change while no bank is in flight: the setters return
 *     U96_ERR_STATE between a submit and the u96_wait() that reports its bank.  A bank remembers the geometry
 *     it was filled with; receive_* / reproject / uvc use that, not the current configuration.
 *
 * A "batch" is n stereo pairs stored back to back (frame i of an image starts
 * at base + i*stride*height).
 */
#ifndef U96_STEREO_H
#define U96_STEREO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define U96_ABI_VERSION 5

enum {
    U96_OK = 0,
    U96_ERR_INVALID = -1,      /* bad argument / parameter combination          */
    U96_ERR_CUDA = -2,         /* CUDA runtime error (see u96_last_cuda_error)  */
    U96_ERR_NOMEM = -3,
    U96_ERR_STATE = -4,        /* e.g. receive from a bank that was never filled */
    U96_ERR_UNSUPPORTED = -5,
    U96_ERR_NODEVICE = -6
};

enum { U96_PROFILE_RTL = 0, U96_PROFILE_OPENCV = 1 };

/* which image of a bank (u96_bank_device_ptr) */
enum { U96_BUF_RAW_L = 0, U96_BUF_RAW_R, U96_BUF_RECT_L, U96_BUF_RECT_R,
       U96_BUF_XSBL_L, U96_BUF_XSBL_R, U96_BUF_DISP, U96_BUF_EIG, U96_BUF_COUNT };

typedef struct u96_handle u96_handle;

/* The "StereoBM entry point + parameter struct".
 * RTL profile      <-> struct FPGA_REG_BM (StereoBM/src/fpga.h:154-169):
 *   ImageSize=(H<<16)+W, BmSetting=(wsz<<16)+ndisp, UniFiltCtrl={enb[31],mode[16],thr[9:0]}
 *   (dvp/rtl/bm.v:148-229), as programmed by Fpga_Init (StereoBM/src/fpga.c:150-160).
 * OPENCV profile   <-> the cv::StereoBM setters at slam/src/core/main.cpp:198-212. */
typedef struct {
    int32_t width, height;
    int32_t block_size;         /* wsz / blockSize, odd                              */
    int32_t num_disparities;    /* RTL: multiple of 32, <=256; OPENCV: multiple of 16 */
    int32_t min_disparity;      /* must be 0 (both reference paths use 0)            */
    int32_t prefilter_cap;      /* OPENCV only (RTL clip is fixed [-32,31]+32)       */
    int32_t uniqueness_ratio;   /* OPENCV only, percent                              */
    int32_t texture_threshold;  /* OPENCV only                                       */
    int32_t profile;            /* U96_PROFILE_*                                     */
    int32_t uni_enable, uni_mode, uni_thr;   /* RTL only: UniFiltCtrl fields          */
    int32_t x_store_offset;     /* RTL only: 1 = DISP bank layout (bm_obuf2.v:125)   */
    int32_t rtl_extended;       /* RTL only: 1 = zero-extend >>4 (needed for D>128)  */
    /* OPENCV only: the post filters cv::StereoBM::compute applies (main.cpp:210-212) */
    int32_t disp12_max_diff;    /* validateDisparity threshold in pixels, < 0 = off  */
    int32_t speckle_window_size;/* filterSpeckles maxSpeckleSize, 0 = off            */
    int32_t speckle_range;      /* filterSpeckles maxDiff (on the 16x values, as cv::StereoBM passes it) */
} u96_bm_params;

/* = struct RECT_PARAM (StereoBM/src/fpga.h:250-260) / FPGA_REG_RECT (fpga.h:178-214);
 * index 0 = left camera, 1 = right camera. */
typedef struct {
    int32_t f[2][2];
    int32_t c[2];
    int32_t f2inv[2];
    int32_t c2_f2[2];
    int32_t rot[2][3][3];
} u96_rect_params;

/* ---- lifetime:  Fpga::registerOpen/memoryOpen/.../Close (slam/src/core/FPGA.cpp:27-139) ---- */
int  u96_create(u96_handle **out, int device, int max_w, int max_h, int max_batch);
void u96_destroy(u96_handle *h);

/* ---- configuration -------------------------------------------------------------------- */
/* writes to FPGA_REG_BM (fpga.c:150-160) / cv::StereoBM setters (main.cpp:198-212) */
int  u96_set_bm_params(u96_handle *h, const u96_bm_params *p);
/* same, register-level: the three words Fpga_Init writes (RTL profile) */
int  u96_set_bm_registers(u96_handle *h, uint32_t image_size, uint32_t bm_setting, uint32_t uni_filt_ctrl);
int  u96_get_bm_params(u96_handle *h, u96_bm_params *p);
/* set_rect_param() (fpga.c:267-301); also rebuilds the on-device map (rect_remap, fpga.c:303-366) */
int  u96_set_rect_params(u96_handle *h, const u96_rect_params *p);
/* fpga->gftt.Control = FPGA_GFTT_CTRL_ENABLE (StereoBM/src/fpga.c:162-172; register file dvp/rtl/gftt.v:117-160): when
 * enabled every submit_raw / submit_rect also produces the min-eigenvalue map of the rectified LEFT image
 * (gftt.Address_In = BUF_RECT).  Off by default, like the firmware without RETURN_DATA_GFTT. */
int  u96_set_gftt(u96_handle *h, int enable);
/* run kernels of this handle on a caller-owned cudaStream_t (NULL = internal per-bank streams) */
int  u96_set_stream(u96_handle *h, void *cuda_stream);

/* ---- per-frame data plane -------------------------------------------------------------- */
/* sensor path (CSI -> rect -> xsbl -> bm, dvp/rtl/rect.v:323-340): raw u8 pairs from HOST memory */
int  u96_submit_raw (u96_handle *h, int bank, const uint8_t *L, const uint8_t *R, int stride, int n);
/* Fpga::setRectImage + FPGA_XSBL_SW_START (FPGA.cpp:236-249, main.cpp:165-175): rectified pairs */
int  u96_submit_rect(u96_handle *h, int bank, const uint8_t *L, const uint8_t *R, int stride, int n);
/* sim_dvp.v SimMode 5 (LOAD_XSBL -> bm only): prefiltered pairs */
int  u96_submit_xsbl(u96_handle *h, int bank, const uint8_t *L, const uint8_t *R, int stride, int n);
/* pipelined variants: the 16x disparity of the batch is also copied to `disp_out` (host, n*H*W int16, ideally
 * pinned) behind the kernels; batches of >= 64 pairs are cut into chunks whose H2D copy, kernels and D2H copy
 * overlap on internal streams.  u96_wait() for the bank covers the copies.  disp_out may be NULL. */
int  u96_submit_raw_async (u96_handle *h, int bank, const uint8_t *L, const uint8_t *R, int stride, int n, int16_t *disp_out);
int  u96_submit_rect_async(u96_handle *h, int bank, const uint8_t *L, const uint8_t *R, int stride, int n, int16_t *disp_out);
/* same three entry points for inputs already resident in device memory.  16-byte aligned pointers with a stride that is
 * a multiple of 16 and >= width rounded up to 16 are used IN PLACE (no copy; the kernels' vector loads read the row padding
 * up to that rounded width, never write it): the caller's buffers must then stay valid and unchanged until the bank is
 * submitted again -- u96_receive_rect / u96_receive_xsbl / u96_receive_uvc of that bank read them after u96_wait().
 * Other pointers / strides are copied into the bank (device to device). */
int  u96_submit_raw_device (u96_handle *h, int bank, const void *dL, const void *dR, int stride, int n);
int  u96_submit_rect_device(u96_handle *h, int bank, const void *dL, const void *dR, int stride, int n);
int  u96_submit_xsbl_device(u96_handle *h, int bank, const void *dL, const void *dR, int stride, int n);

/* Fpga::setRectImage alone (FPGA.cpp:236-249): the n rectified pairs are written into the RECT bank and nothing runs;
 * returns when the caller's buffers may be reused (the reference's memcpy).  u96_receive_rect of the bank then reads
 * them back like Fpga::receiveRectImages would. */
int  u96_set_rect_image(u96_handle *h, int bank, const uint8_t *L, const uint8_t *R, int stride, int n);
/* reg->xsbl.Control |= FPGA_XSBL_SW_START (main.cpp:172-174, FPGA.h:289): x-Sobel -> BM on what the
 * RECT bank holds (filled by u96_set_rect_image); completion is reported by u96_wait like for a submit. */
int  u96_start_xsbl(u96_handle *h, int bank);

/* Fpga::waitIpcMessage(IPC_MSG2_DATA_READY) + IpcParameter2 (FPGA.cpp:217-220, 310-314):
 * blocks until the oldest submitted bank is complete and returns its index. */
int  u96_wait(u96_handle *h, int *active_bank);

/* Fpga::receiveRectImages (FPGA.cpp:251-268): n*H*W bytes each, row stride W */
int  u96_receive_rect(u96_handle *h, int bank, uint8_t *L, uint8_t *R);
/* USB_OUTPUT_STEREO_XSBL debug stream (fpga.c:57-66): planar 6-bit-in-byte images */
int  u96_receive_xsbl(u96_handle *h, int bank, uint8_t *L, uint8_t *R);
/* Fpga::receiveDepthMap (FPGA.cpp:270-279): CV_16SC1, 16x fixed-point disparity */
int  u96_receive_disp(u96_handle *h, int bank, int16_t *disp);
/* asynchronous variant for pipelined callers: enqueues the device->host copy of the disparity behind the bank's
 * kernels and returns at once; the NEXT u96_wait() that reports this bank also covers the copy.  Must be called
 * after the submit and before the wait of that bank; `disp` should be pinned (u96_host_alloc). */
int  u96_enqueue_receive_disp(u96_handle *h, int bank, int16_t *disp);
/* projectDisparityTo3D over the (decimated) map: Stereo.cpp:157-182, main.cpp:522-551,
 * SensorData.cpp:50-58.  P_l/P_r 3x4 row-major; decim in {1,2,4,8}; xyz = n*(H/decim)*(W/decim)*3 floats.
 * flags bit0: apply StereoCameraModel localTransform (StereoCameraModel.cpp:9-14). */
int  u96_reproject(u96_handle *h, int bank, const double P_l[12], const double P_r[12],
                   int decim, int flags, float *xyz);
/* The same dense consumer with the transforms of main.cpp:535-541 spelled out: every finite point goes through
 * transformPoint (Stereo.cpp:189-198) with local_T (3x4 row-major floats; NULL or all zero = none, Transform::isNull)
 * and then with the frame's pose (poses = n x 12 floats, host; NULL = none).  Points the reference skips are NaN. */
int  u96_reproject_ex(u96_handle *h, int bank, const double P_l[12], const double P_r[12], int decim,
                      const float *local_T, const float *poses, float *xyz);
/* generateKeypoints3DStereo (Stereo.cpp:53-117) as the real-time loop calls it every frame through generateKeypoints3D
 * (Stereo.cpp:119-154, main.cpp:250-252) with a dense-map depth method: for each of the n keypoints uv[i] = (x, y) of frame
 * `frame` of the bank: disparity = map[(int)y][(int)x] / 16.0f, negative -> 0, 0 -> bad point; projectDisparityTo3D with the
 * FLOAT keypoint coordinates; kept iff finite and (min_depth < 0 or z > min_depth) and (max_depth <= 0 or z <= max_depth);
 * then local_T as above.  mask: n bytes (0 = skip) or NULL.  xyz: n x 3 floats, NaN = bad point.  Keypoints outside the
 * map (undefined behaviour in the reference) are bad points.  generateKeypoints3D passes min_depth = max_depth = 0. */
int  u96_reproject_points(u96_handle *h, int bank, int frame, const double P_l[12], const double P_r[12],
                          const float *uv, int n, const uint8_t *mask, float min_depth, float max_depth,
                          const float *local_T, float *xyz);

/* Fpga::receiveEigen (FPGA.cpp:281-296): CV_16UC1 min-eigenvalue map (n*H*W u16, rows 0,1,H-2,H-1 zero like the
 * firmware-cleared GFTT bank) and the per-frame maximum the FPGA latches in gftt.Max (gftt_obuf.v:90-118); max_eig
 * receives n values (may be NULL).  Needs u96_set_gftt(h, 1) before the submit and a bank filled from raw or rect. */
int  u96_receive_eigen(u96_handle *h, int bank, uint16_t *eig, uint16_t *max_eig);

/* Xusb_ReceiveData (StereoBM/src/xusb_main.c:293-376): the UVC payload the R5 firmware streams for a bank -- per pair one
 * YUYV frame of 2W x H pixels (2 bytes per pixel, chroma byte 0x80): left half = left image / disparity, right half =
 * right image / zero.  which: U96_UVC_RECT (USB_OUTPUT_STEREO_RECT), U96_UVC_XSBL (USB_OUTPUT_STEREO_XSBL),
 * U96_UVC_BM (USB_OUTPUT_STEREO_BM: Y = (u8)(s16 disparity >> 4)).  frame = n * H * 2W * 2 bytes (host). */
enum { U96_UVC_RECT = 1, U96_UVC_XSBL = 2, U96_UVC_BM = 3 };
int  u96_receive_uvc(u96_handle *h, int bank, int which, uint8_t *frame);

/* device address + row pitch of a bank image (results stay resident for a GPU consumer) */
int  u96_bank_device_ptr(u96_handle *h, int bank, int which, void **dptr, int *pitch_bytes, size_t *frame_bytes);

/* ---- auxiliaries ----------------------------------------------------------------------- */
/* pinned host staging memory for full-rate H2D/D2H */
int  u96_host_alloc(void **p, size_t bytes);
/* write-combined pinned memory: for INPUT staging buffers the host only writes sequentially (reads are very slow) */
int  u96_host_alloc_wc(void **p, size_t bytes);
int  u96_host_free(void *p);
/* Perf-style stage timers (slam/include/core/Perf.h): ms of the last completed submit
 * [0]=h2d [1]=rect(+gftt) [2]=xsbl [3]=bm(+post filters) ; needs u96_set_profiling(h,1) */
int  u96_set_profiling(u96_handle *h, int on);
int  u96_last_stage_ms(u96_handle *h, int bank, float ms[4]);
/* every stage of the last completed (non-pipelined) submit of a bank on its own: ms[U96_STAGE_*], count <= U96_STAGE_COUNT */
enum { U96_STAGE_H2D = 0, U96_STAGE_RECT, U96_STAGE_GFTT, U96_STAGE_XSBL, U96_STAGE_BM, U96_STAGE_POST, U96_STAGE_COUNT };
int  u96_last_stage_ms_ex(u96_handle *h, int bank, float *ms, int count);
/* kernel time of the last u96_reproject(_ex) / u96_reproject_points / u96_receive_uvc call (profiling on) */
enum { U96_AUX_REPROJECT = 0, U96_AUX_REPROJECT_POINTS, U96_AUX_UVC, U96_AUX_COUNT };
int  u96_last_aux_ms(u96_handle *h, int which, float *ms);
/* number of kernels this handle has launched so far */
int64_t u96_kernel_launches(u96_handle *h);
/* issue-rate micro-benchmark used for the INT roofline denominator:
 * which: 0=IADD3 1=VABSDIFF4 2=VIADDMNMX.U16x2 3=VIMNMX3.U32 4=PRMT 5=IMAD 6=LDS.128 7=SHFL
 * returns giga lane-ops (or GB for LDS) per second over all SMs */
int  u96_microbench(int device, int which, double *gops);
const char *u96_strerror(int code);
const char *u96_last_cuda_error(void);
int  u96_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif
