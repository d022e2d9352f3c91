// u96_stereo.cu -- the C ABI of libu96stereo (include/u96_stereo.h): handle, banks, streams.
//
// The handle plays the role of the reference's `class Fpga` (slam/include/core/FPGA.h:347-397,
// slam/src/core/FPGA.cpp): it owns two banks (A/B) of every buffer the FPGA keeps in DDR
// (RECT, XSBL, DISP; StereoBM/src/fpga.h:50-68), fills a bank asynchronously on submit and copies
// results out on receive.  There is no CPU code path: every stage is a CUDA kernel.
#include <algorithm>
#include <cstdlib>
#include <deque>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"

using namespace u96;

static thread_local std::string g_cuda_err;

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            g_cuda_err = std::string(#call) + ": " + cudaGetErrorString(e__);             \
            return U96_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

namespace {

// 2-D copy that degenerates to one contiguous transfer when both pitches equal the row width (640-wide images in
// 640-byte rows): the DMA engines move one large block far faster than hundreds of thousands of row descriptors.
// A tight host buffer meeting a pitched device buffer (widths that are not a multiple of 128, e.g. KITTI's 1242) goes through
// a tight device staging area when one is given: ONE contiguous transfer over PCIe plus an on-device pitch conversion
// (~1 TB/s), instead of a row-descriptor DMA that reaches a third of the PCIe rate (measured 9.3 k -> 26 k frames/s on C3).
cudaError_t copy2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind, cudaStream_t s,
                   void *stage = nullptr)
{
    if (dpitch == width && spitch == width) return cudaMemcpyAsync(dst, src, width * height, kind, s);
    if (stage && kind == cudaMemcpyHostToDevice && spitch == width) {
        const cudaError_t e = cudaMemcpyAsync(stage, src, width * height, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return e;
        return cudaMemcpy2DAsync(dst, dpitch, stage, width, width, height, cudaMemcpyDeviceToDevice, s);
    }
    if (stage && kind == cudaMemcpyDeviceToHost && dpitch == width) {
        const cudaError_t e = cudaMemcpy2DAsync(stage, width, src, spitch, width, height, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync(dst, stage, width * height, cudaMemcpyDeviceToHost, s);
    }
    return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, kind, s);
}

enum { FROM_RAW = 0, FROM_RECT = 1, FROM_XSBL = 2 };
// per-stage events of a submit: start | h2d | rect | gftt | xsbl | bm | post filters
constexpr int NEV = U96_STAGE_COUNT + 1;

struct Bank {
    uint8_t *raw[2] = {nullptr, nullptr}, *rect[2] = {nullptr, nullptr}, *xsbl[2] = {nullptr, nullptr};
    int16_t *disp = nullptr;
    int16_t *cost = nullptr;            // OPENCV post filters: winning SAD per pixel (lazy)
    int *cc = nullptr;                  // filterSpeckles: label + size, 2 int32 per pixel of a batch (lazy, per bank)
    size_t cc_cap = 0;
    void *sat = nullptr; size_t sat_cap = 0;               // band functions / states of the saturating chain (a handful of pairs, bm_fused.cuh)
    uint8_t *stage = nullptr;           // tight staging area for host transfers of non-pitch-aligned widths (lazy): L | R | 16-bit out
    size_t stage_cap = 0;
    uint16_t *eig = nullptr;            // GFTT min-eigenvalue map (lazy), same pitch (in elements) as the u8 images
    uint32_t *eig_max = nullptr;        // per-frame maximum (gftt.Max)
    bool has_eig = false;
    // where the current contents of each stage live (internal buffer or a caller's device pointer)
    const uint8_t *cur_raw[2] = {nullptr, nullptr}, *cur_rect[2] = {nullptr, nullptr}, *cur_xsbl[2] = {nullptr, nullptr};
    int raw_pitch = 0, rect_pitch = 0, xsbl_pitch = 0;
    size_t raw_frame = 0, rect_frame = 0, xsbl_frame = 0;
    int n = 0, from = -1;
    int W = 0, H = 0;                   // geometry the bank was filled with (snapshot at submit: receive / reproject / uvc use these)
    bool filled = false, pending = false;
    bool has_disp = false;              // the kernels ran (false after u96_set_rect_image until u96_start_xsbl)
    bool staged = false;                // pipelined submit: per-stage events were not recorded
    uint64_t border_sig = 0;            // configuration the border of the DISP bank was filled for (0 = never)
    cudaStream_t stream = nullptr;
    cudaStream_t sub[3] = {nullptr, nullptr, nullptr};      // chunk pipeline: H2D / kernels / D2H of different chunks overlap
    cudaEvent_t sub_ev[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t done = nullptr, ev[NEV] = {};
};

}  // namespace

struct u96_handle {
    int device = 0, maxW = 0, maxH = 0, maxB = 0;
    int pitch = 0;                 // internal u8 row pitch (bytes)
    u96_bm_params bm{};
    u96_rect_params rect{};
    bool rect_set = false, map_valid = false;
    int2 *map = nullptr;
    RectPlan plan;
    Bank bank[2];
    std::deque<int> fifo;
    cudaStream_t user_stream = nullptr;
    bool use_user_stream = false, profiling = false, gftt = false;
    int64_t launches = 0;
    float *xyz = nullptr;
    size_t xyz_cap = 0;
    uint8_t *uvc = nullptr;
    size_t uvc_cap = 0;
    float *kp = nullptr;           // keypoint scratch: uv (2n floats) | xyz (3n floats) | mask (n bytes)
    size_t kp_cap = 0;
    float *poses = nullptr;        // dense reprojection: one 3x4 pose per frame
    cudaEvent_t aux_ev[2] = {nullptr, nullptr};
    float aux_ms[U96_AUX_COUNT] = {};
    bool aux_valid[U96_AUX_COUNT] = {};
};

static bool any_pending(const u96_handle *h) { return h->bank[0].pending || h->bank[1].pending; }

static int validate_bm(const u96_handle *h, const u96_bm_params &p)
{
    if (p.width <= 0 || p.height <= 0 || p.width > h->maxW || p.height > h->maxH) return U96_ERR_INVALID;
    if (p.min_disparity != 0) return U96_ERR_UNSUPPORTED;
    if (!(p.block_size & 1) || p.block_size < 3 || p.block_size > 31) return U96_ERR_INVALID;
    const int hw = p.block_size >> 1;
    int max_ad;
    if (p.profile == U96_PROFILE_RTL) {
        // bm.v:174-177: wsz 5 bit, ndisp 9 bit; processed in 32-disparity dphases (bm_ibuf.v:143-189)
        if (p.num_disparities < 32 || p.num_disparities > 256 || (p.num_disparities & 31)) return U96_ERR_INVALID;
        if (p.uni_thr < 0 || p.uni_thr > 1023) return U96_ERR_INVALID;
        if (p.x_store_offset != 0 && p.x_store_offset != 1) return U96_ERR_INVALID;
        if (p.width - p.num_disparities - 1 - 2 * hw <= 0) return U96_ERR_INVALID;
        max_ad = 63;
    } else if (p.profile == U96_PROFILE_OPENCV) {
        if (p.block_size < 5) return U96_ERR_INVALID;
        if (p.num_disparities < 16 || p.num_disparities > 256 || (p.num_disparities & 15)) return U96_ERR_INVALID;
        if (p.prefilter_cap < 1 || p.prefilter_cap > 63) return U96_ERR_INVALID;
        if (p.uniqueness_ratio < 0 || p.texture_threshold < 0) return U96_ERR_INVALID;
        if (p.speckle_window_size < 0 || p.speckle_window_size > 1000000) return U96_ERR_INVALID;
        if (p.width - p.num_disparities + 1 - 2 * hw <= 0) return U96_ERR_INVALID;
        if (p.disp12_max_diff >= 0 && (size_t)p.width * 4 > 227 * 1024) return U96_ERR_UNSUPPORTED;   // k_validate keeps one row of 32-bit keys in shared memory
        max_ad = 2 * p.prefilter_cap;
    } else return U96_ERR_INVALID;
    if (p.height - 2 * hw <= 0) return U96_ERR_INVALID;
    if (p.block_size * p.block_size * max_ad > 65535) return U96_ERR_UNSUPPORTED;     // window sums are u16
    BmConfig c{p.width, p.height, p.num_disparities, p.block_size, p.profile, 0, 0, 0, 0, 0, 0, 0, 0};
    if (bm_smem_bytes(c) > 227 * 1024) return U96_ERR_UNSUPPORTED;
    return U96_OK;
}

static BmConfig bm_config(const u96_bm_params &p)
{
    BmConfig c;
    c.W = p.width; c.H = p.height; c.D = p.num_disparities; c.wsz = p.block_size; c.profile = p.profile;
    c.uni_enable = p.uni_enable; c.uni_mode = p.uni_mode; c.uni_thr = p.uni_thr;
    c.x_store_offset = p.x_store_offset; c.rtl_extended = p.rtl_extended;
    c.cap = p.prefilter_cap; c.tex_thr = p.texture_threshold; c.uniq = p.uniqueness_ratio;
    return c;
}

extern "C" {

int u96_abi_version(void) { return U96_ABI_VERSION; }

const char *u96_strerror(int code)
{
    switch (code) {
    case U96_OK: return "ok";
    case U96_ERR_INVALID: return "invalid argument";
    case U96_ERR_CUDA: return "CUDA error";
    case U96_ERR_NOMEM: return "out of memory";
    case U96_ERR_STATE: return "bank not ready / wrong state";
    case U96_ERR_UNSUPPORTED: return "unsupported parameter combination";
    case U96_ERR_NODEVICE: return "no CUDA device (there is no CPU fallback)";
    default: return "unknown error";
    }
}

const char *u96_last_cuda_error(void) { return g_cuda_err.c_str(); }

int u96_create(u96_handle **out, int device, int max_w, int max_h, int max_batch)
{
    if (!out || max_w <= 0 || max_h <= 0 || max_batch <= 0) return U96_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { g_cuda_err = "no CUDA device"; return U96_ERR_NODEVICE; }
    if (device < 0 || device >= ndev) return U96_ERR_INVALID;
    CK(cudaSetDevice(device));
    u96_handle *h = new (std::nothrow) u96_handle();
    if (!h) return U96_ERR_NOMEM;
    h->device = device; h->maxW = max_w; h->maxH = max_h; h->maxB = max_batch;
    h->pitch = align_up(max_w, 128);
    const size_t img = (size_t)h->pitch * max_h * max_batch;
    auto fail = [&](int rc) { u96_destroy(h); return rc; };
    for (int b = 0; b < 2; b++) {
        Bank &k = h->bank[b];
        for (int i = 0; i < 2; i++) {
            if (cudaMalloc(&k.raw[i], img) != cudaSuccess || cudaMalloc(&k.rect[i], img) != cudaSuccess ||
                cudaMalloc(&k.xsbl[i], img) != cudaSuccess) return fail(U96_ERR_NOMEM);
        }
        if (cudaMalloc(&k.disp, img * sizeof(int16_t)) != cudaSuccess) return fail(U96_ERR_NOMEM);
        // the banks start zeroed (like the firmware's memset of the DDR banks, fpga.c:105-114): the vectorised loads of the
        // kernels touch the pitch padding right of the image, which no kernel ever writes
        for (int i = 0; i < 2; i++)
            if (cudaMemset(k.raw[i], 0, img) != cudaSuccess || cudaMemset(k.rect[i], 0, img) != cudaSuccess ||
                cudaMemset(k.xsbl[i], 0, img) != cudaSuccess) return fail(U96_ERR_CUDA);
        if (cudaMemset(k.disp, 0, img * sizeof(int16_t)) != cudaSuccess) return fail(U96_ERR_CUDA);
        if (cudaStreamCreateWithFlags(&k.stream, cudaStreamNonBlocking) != cudaSuccess) return fail(U96_ERR_CUDA);
        if (cudaEventCreateWithFlags(&k.done, cudaEventDisableTiming) != cudaSuccess) return fail(U96_ERR_CUDA);
        for (int i = 0; i < 3; i++)
            if (cudaStreamCreateWithFlags(&k.sub[i], cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&k.sub_ev[i], cudaEventDisableTiming) != cudaSuccess) return fail(U96_ERR_CUDA);
        for (int i = 0; i < NEV; i++)
            if (cudaEventCreate(&k.ev[i]) != cudaSuccess) return fail(U96_ERR_CUDA);
    }
    for (int i = 0; i < 2; i++)
        if (cudaEventCreate(&h->aux_ev[i]) != cudaSuccess) return fail(U96_ERR_CUDA);
    if (cudaMalloc(&h->map, sizeof(int2) * 2 * (size_t)max_w * max_h) != cudaSuccess) return fail(U96_ERR_NOMEM);
    // defaults = what the firmware programs (fpga.c:150-160): 640x480, wsz 21, 64 disparities, uniqueness off
    u96_bm_params d{};
    d.width = max_w; d.height = max_h; d.block_size = 21; d.num_disparities = 64; d.prefilter_cap = 31;
    d.uniqueness_ratio = 10; d.texture_threshold = 10; d.profile = U96_PROFILE_RTL; d.x_store_offset = 1;
    d.disp12_max_diff = -1; d.speckle_window_size = 0; d.speckle_range = 0;
    h->bm = d;
    *out = h;
    return U96_OK;
}

void u96_destroy(u96_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    for (int b = 0; b < 2; b++) {
        Bank &k = h->bank[b];
        if (k.stream) cudaStreamSynchronize(k.stream);
        for (int i = 0; i < 2; i++) { cudaFree(k.raw[i]); cudaFree(k.rect[i]); cudaFree(k.xsbl[i]); }
        cudaFree(k.disp);
        cudaFree(k.stage);
        cudaFree(k.cost);
        cudaFree(k.cc);
        cudaFree(k.sat);
        cudaFree(k.eig);
        cudaFree(k.eig_max);
        if (k.done) cudaEventDestroy(k.done);
        for (int i = 0; i < NEV; i++) if (k.ev[i]) cudaEventDestroy(k.ev[i]);
        for (int i = 0; i < 3; i++) { if (k.sub_ev[i]) cudaEventDestroy(k.sub_ev[i]); if (k.sub[i]) { cudaStreamSynchronize(k.sub[i]); cudaStreamDestroy(k.sub[i]); } }
        if (k.stream) cudaStreamDestroy(k.stream);
    }
    cudaFree(h->map);
    rect_plan_free(h->plan);
    cudaFree(h->xyz);
    cudaFree(h->uvc);
    cudaFree(h->kp);
    cudaFree(h->poses);
    for (int i = 0; i < 2; i++) if (h->aux_ev[i]) cudaEventDestroy(h->aux_ev[i]);
    delete h;
}

int u96_set_bm_params(u96_handle *h, const u96_bm_params *p)
{
    if (!h || !p) return U96_ERR_INVALID;
    const int rc = validate_bm(h, *p);
    if (rc != U96_OK) return rc;
    // a bank in flight still reads the current configuration (and, through the map and the tile plan, the geometry):
    // like the FPGA's registers, the parameters may only change while the pipeline is idle
    if (any_pending(h)) return U96_ERR_STATE;
    if (p->width != h->bm.width || p->height != h->bm.height) h->map_valid = false;
    h->bm = *p;
    return U96_OK;
}

int u96_set_bm_registers(u96_handle *h, uint32_t image_size, uint32_t bm_setting, uint32_t uni_filt_ctrl)
{
    if (!h) return U96_ERR_INVALID;
    u96_bm_params p = h->bm;
    p.profile = U96_PROFILE_RTL;
    p.height = (image_size >> 16) & 0x1FF;            // bm.v:170-171
    p.width = image_size & 0x3FF;
    p.block_size = (bm_setting >> 16) & 0x1F;         // bm.v:174-177
    p.num_disparities = bm_setting & 0x1FF;
    p.uni_enable = (uni_filt_ctrl >> 31) & 1;         // bm.v:183-187
    p.uni_mode = (uni_filt_ctrl >> 16) & 1;
    p.uni_thr = uni_filt_ctrl & 0x3FF;
    p.min_disparity = 0;
    return u96_set_bm_params(h, &p);
}

int u96_get_bm_params(u96_handle *h, u96_bm_params *p)
{
    if (!h || !p) return U96_ERR_INVALID;
    *p = h->bm;
    return U96_OK;
}

int u96_set_rect_params(u96_handle *h, const u96_rect_params *p)
{
    if (!h || !p) return U96_ERR_INVALID;
    if (any_pending(h)) return U96_ERR_STATE;                 // the pending bank's remap kernel reads the map this call invalidates
    h->rect = *p;
    h->rect_set = true;
    h->map_valid = false;
    return U96_OK;
}

int u96_set_stream(u96_handle *h, void *cuda_stream)
{
    if (!h) return U96_ERR_INVALID;
    if (any_pending(h)) return U96_ERR_STATE;
    h->user_stream = (cudaStream_t)cuda_stream;
    h->use_user_stream = true;
    return U96_OK;
}

int u96_set_gftt(u96_handle *h, int enable)
{
    if (!h) return U96_ERR_INVALID;
    h->gftt = enable != 0;
    return U96_OK;
}

int u96_set_profiling(u96_handle *h, int on)
{
    if (!h) return U96_ERR_INVALID;
    h->profiling = on != 0;
    return U96_OK;
}

int64_t u96_kernel_launches(u96_handle *h) { return h ? h->launches : 0; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------
static cudaStream_t bank_stream(u96_handle *h, int bank) { return h->use_user_stream ? h->user_stream : h->bank[bank].stream; }

// Tight staging area of a bank (see copy2d): L | R | 16-bit output, each for maxB frames of the current geometry.  Only
// needed when the image width differs from the internal pitch; the bank must be idle when it grows.
static int ensure_stage(u96_handle *h, Bank &k, int W, int H)
{
    const size_t px = (size_t)W * H * h->maxB;
    if (W == h->pitch) return U96_OK;                         // contiguous transfers anyway
    if (k.stage_cap >= 4 * px) return U96_OK;
    cudaFree(k.stage); k.stage = nullptr; k.stage_cap = 0;
    if (cudaMalloc(&k.stage, 4 * px) != cudaSuccess) return U96_ERR_NOMEM;
    k.stage_cap = 4 * px;
    return U96_OK;
}
static uint8_t *stage_in(u96_handle *h, Bank &k, int i) { return k.stage ? k.stage + (size_t)i * k.W * k.H * h->maxB : nullptr; }
static uint8_t *stage_out(u96_handle *h, Bank &k) { return k.stage ? k.stage + (size_t)2 * k.W * k.H * h->maxB : nullptr; }

static int ensure_map(u96_handle *h, cudaStream_t s)
{
    if (h->map_valid) return U96_OK;
    if (!h->rect_set) return U96_ERR_STATE;
    RectMapParams rp;
    rp.p = h->rect; rp.W = h->bm.width; rp.H = h->bm.height;
    rp.wrap16 = (rp.W <= 1023 && rp.H <= 511) ? 1 : 0;      // RTL counter widths; beyond = RTL-extended
    h->launches += launch_rect_build_map(rp, h->map, s);
    rect_plan_free(h->plan);                                  // no bank is in flight here: the setters refuse while one is pending
    h->launches += rect_plan_build(h->plan, h->map, rp.W, rp.H, s);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));                             // other bank's stream may use the map next
    h->map_valid = true;
    return U96_OK;
}

// lazily allocated per-bank buffers the current configuration needs; called before anything is enqueued
static int ensure_bank_buffers(u96_handle *h, Bank &k, int from, int n, bool host_src)
{
    const int W = h->bm.width, H = h->bm.height, pitch = h->pitch;
    const bool cv = (h->bm.profile == U96_PROFILE_OPENCV);
    if (cv && h->bm.disp12_max_diff >= 0 && !k.cost)
        if (cudaMalloc(&k.cost, (size_t)pitch * h->maxH * h->maxB * sizeof(int16_t)) != cudaSuccess) return U96_ERR_NOMEM;
    if (cv && h->bm.speckle_window_size > 0) {
        const size_t need = (size_t)2 * n * W * H;
        if (need > k.cc_cap) {                                // the bank is idle here (not pending)
            cudaFree(k.cc); k.cc = nullptr; k.cc_cap = 0;
            if (cudaMalloc(&k.cc, need * sizeof(int)) != cudaSuccess) return U96_ERR_NOMEM;
            k.cc_cap = need;
        }
    }
    {
        const size_t need = bm_sat_scratch_bytes(bm_config(h->bm), n);
        if (need > k.sat_cap) {                               // the bank is idle here (not pending)
            cudaFree(k.sat); k.sat = nullptr; k.sat_cap = 0;
            if (cudaMalloc(&k.sat, need) == cudaSuccess) k.sat_cap = need;
            else { k.sat = nullptr; cudaGetLastError(); }       // no scratch: the chain of this batch simply stays sequential (launch_bm checks)
        }
    }
    if (h->gftt && from <= FROM_RECT && !k.eig)
        if (cudaMalloc(&k.eig, (size_t)pitch * h->maxH * h->maxB * sizeof(uint16_t)) != cudaSuccess ||
            cudaMalloc(&k.eig_max, (size_t)h->maxB * sizeof(uint32_t)) != cudaSuccess) return U96_ERR_NOMEM;
    if (host_src) return ensure_stage(h, k, W, H);
    return U96_OK;
}

// The DISP bank's border (outside the valid rectangle) is written once per bank and configuration: the BM kernels never touch it
// (the firmware memsets the DDR banks once, fpga.c:105-106).  All maxB frames are filled, so later submits of any batch size hold.
static int ensure_border(u96_handle *h, Bank &k, cudaStream_t s)
{
    const u96_bm_params &p = h->bm;
    const uint64_t sig = 1ull | ((uint64_t)p.profile << 1) | ((uint64_t)p.width << 4) | ((uint64_t)p.height << 20) |
                         ((uint64_t)p.num_disparities << 36) | ((uint64_t)p.block_size << 48) | ((uint64_t)(p.x_store_offset & 1) << 56);
    if (k.border_sig == sig) return U96_OK;
    const Img16 disp{k.disp, h->pitch, (size_t)h->pitch * p.height};
    h->launches += launch_bm_border(disp, bm_config(p), h->maxB, s);
    CK(cudaGetLastError());
    k.border_sig = sig;
    return U96_OK;
}

// kernels of frames [f0, f0+nf) of a bank on stream s; `prof` records the Perf-style stage events
static int run_range(u96_handle *h, Bank &k, int from, int f0, int nf, cudaStream_t s, bool prof)
{
    const int W = k.W, H = k.H, pitch = h->pitch;
    const size_t frame = (size_t)pitch * H, o = (size_t)f0 * frame;
    const Img8 rectL{k.rect[0] + o, pitch, frame}, rectR{k.rect[1] + o, pitch, frame};
    const Img8 xsblL{k.xsbl[0] + o, pitch, frame}, xsblR{k.xsbl[1] + o, pitch, frame};
    const Img16 disp{k.disp + o, pitch, frame};
    if (from == FROM_RAW)
        h->launches += launch_rect_remap(k.cur_raw[0] + (size_t)f0 * k.raw_frame, k.cur_raw[1] + (size_t)f0 * k.raw_frame, k.raw_pitch,
                                         k.raw_frame, rectL, rectR, h->map, h->plan, W, H, nf, s);
    if (prof) CK(cudaEventRecord(k.ev[2], s));
    if (h->gftt && from <= FROM_RECT)                        // the GFTT block reads the RECT bank (fpga.c:166-167)
        h->launches += launch_gftt(k.cur_rect[0] + (size_t)f0 * k.rect_frame, k.rect_pitch, k.rect_frame, k.eig + o, pitch, frame,
                                   k.eig_max + f0, W, H, nf, s);
    if (prof) CK(cudaEventRecord(k.ev[3], s));
    if (from <= FROM_RECT)
        h->launches += launch_xsobel(k.cur_rect[0] + (size_t)f0 * k.rect_frame, k.cur_rect[1] + (size_t)f0 * k.rect_frame, k.rect_pitch,
                                     k.rect_frame, xsblL, xsblR, W, H, nf, h->bm.profile, h->bm.prefilter_cap, s);
    if (prof) CK(cudaEventRecord(k.ev[4], s));
    BmConfig cfg = bm_config(h->bm);
    const bool cv = (h->bm.profile == U96_PROFILE_OPENCV);
    const bool want_validate = cv && h->bm.disp12_max_diff >= 0;
    const bool want_speckle = cv && h->bm.speckle_window_size > 0 && h->bm.speckle_range >= 0;
    if (want_validate) cfg.cost = k.cost + o;
    cfg.sat_scratch = k.sat; cfg.sat_scratch_bytes = k.sat_cap;
    h->launches += launch_bm(k.cur_xsbl[0] + (size_t)f0 * k.xsbl_frame, k.cur_xsbl[1] + (size_t)f0 * k.xsbl_frame, k.xsbl_pitch,
                             k.xsbl_frame, disp, cfg, nf, s);
    if (prof) CK(cudaEventRecord(k.ev[5], s));
    if (want_validate || want_speckle)                       // cv::StereoBM::compute post filters (main.cpp:210-212)
        h->launches += launch_postfilter(disp, want_validate ? k.cost + o : nullptr, W, H, nf, h->bm.num_disparities,
                                         want_validate ? h->bm.disp12_max_diff : -1, want_speckle ? h->bm.speckle_window_size : 0,
                                         h->bm.speckle_range, want_speckle ? k.cc + (size_t)2 * f0 * W * H : nullptr, s);
    if (prof) CK(cudaEventRecord(k.ev[6], s));
    return U96_OK;
}

// A submit failed after work was already enqueued: drain everything the bank has in flight so that a retry cannot overwrite
// buffers that are still being read, then hand the error back.  The bank stays idle (not pending, not filled).
static int abort_submit(u96_handle *h, Bank &k, cudaStream_t s, int rc)
{
    const std::string first = g_cuda_err;
    for (int i = 0; i < 3; i++) cudaStreamSynchronize(k.sub[i]);
    cudaStreamSynchronize(s);
    cudaGetLastError();
    g_cuda_err = first;
    k.filled = false; k.has_disp = false;
    (void)h;
    return rc;
}
#define CKA(call)                                                                         \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            g_cuda_err = std::string(#call) + ": " + cudaGetErrorString(e__);             \
            return abort_submit(h, k, s, U96_ERR_CUDA);                                   \
        }                                                                                 \
    } while (0)

// device_src: pointers are device memory (zero-copy when aligned).  disp_out (host, optional): the disparity is
// copied out behind the kernels; large host batches are cut into chunks that flow through three sub-streams so
// the H2D copy, the kernels and the D2H copy of different chunks overlap (both copy engines + SMs busy).
// upload_only: Fpga::setRectImage -- the pair lands in the RECT bank and nothing runs until u96_start_xsbl.
static int submit_common(u96_handle *h, int bank, int from, const uint8_t *L, const uint8_t *R, int stride, int n, bool device_src,
                         int16_t *disp_out = nullptr, bool upload_only = false)
{
    if (!h || !L || !R || bank < 0 || bank > 1 || n <= 0 || n > h->maxB) return U96_ERR_INVALID;
    const int W = h->bm.width, H = h->bm.height;
    if (stride < W) return U96_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    Bank &k = h->bank[bank];
    if (k.pending) return U96_ERR_STATE;                      // bank still in flight: wait() first
    cudaStream_t s = bank_stream(h, bank);
    const int pitch = h->pitch;
    const size_t frame = (size_t)pitch * H;
    // ---- everything that can fail without touching the GPU queues comes first ----
    if (from == FROM_RAW) { const int rc = ensure_map(h, s); if (rc != U96_OK) return rc; }
    k.W = W; k.H = H;
    { const int rc = ensure_bank_buffers(h, k, from, n, !device_src); if (rc != U96_OK) return rc; }

    uint8_t *dstbuf[2];
    for (int i = 0; i < 2; i++) dstbuf[i] = (from == FROM_RAW) ? k.raw[i] : (from == FROM_RECT) ? k.rect[i] : k.xsbl[i];
    const uint8_t *src[2] = {L, R};
    const bool zero_copy = device_src && !upload_only && (stride % 16 == 0) && (stride >= align_up(W, 16)) &&
                           (((uintptr_t)L | (uintptr_t)R) % 16 == 0);
    const uint8_t *cur[2] = {zero_copy ? L : dstbuf[0], zero_copy ? R : dstbuf[1]};
    const int cur_pitch = zero_copy ? stride : pitch;
    const size_t cur_frame = zero_copy ? (size_t)stride * H : frame;
    if (from == FROM_RAW) {
        k.cur_raw[0] = cur[0]; k.cur_raw[1] = cur[1]; k.raw_pitch = cur_pitch; k.raw_frame = cur_frame;
        k.cur_rect[0] = k.rect[0]; k.cur_rect[1] = k.rect[1]; k.rect_pitch = pitch; k.rect_frame = frame;
    } else if (from == FROM_RECT) {
        k.cur_rect[0] = cur[0]; k.cur_rect[1] = cur[1]; k.rect_pitch = cur_pitch; k.rect_frame = cur_frame;
    }
    if (from <= FROM_RECT) { k.cur_xsbl[0] = k.xsbl[0]; k.cur_xsbl[1] = k.xsbl[1]; k.xsbl_pitch = pitch; k.xsbl_frame = frame; }
    else { k.cur_xsbl[0] = cur[0]; k.cur_xsbl[1] = cur[1]; k.xsbl_pitch = cur_pitch; k.xsbl_frame = cur_frame; }
    k.has_eig = h->gftt && from <= FROM_RECT && !upload_only;
    k.n = n; k.from = from;

    // ---- from here on work is enqueued: errors drain the bank before they are returned (abort_submit) ----
    if (!upload_only) { const int rc = ensure_border(h, k, s); if (rc != U96_OK) return abort_submit(h, k, s, rc); }
    const bool prof = h->profiling && !upload_only;
    if (prof) CKA(cudaEventRecord(k.ev[0], s));
    static const int pipe_min_env = getenv("U96_PIPE_MIN") ? atoi(getenv("U96_PIPE_MIN")) : 0;        // developer override (frames)
    const int pipe_min = pipe_min_env > 0 ? pipe_min_env : 64;
    const bool pipelined = !device_src && !h->use_user_stream && !h->profiling && n >= pipe_min && !upload_only;
    k.staged = pipelined;
    if (!pipelined) {
        if (!zero_copy)
            for (int i = 0; i < 2; i++)
                CKA(copy2d(dstbuf[i], pitch, src[i], stride, W, (size_t)H * n,
                           device_src ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s, device_src ? nullptr : stage_in(h, k, i)));
        if (upload_only) {
            CKA(cudaStreamSynchronize(s));                    // like the memcpy of FPGA.cpp:236-249: the caller may reuse its buffers
            k.filled = true; k.has_disp = false;
            return U96_OK;
        }
        if (prof) CKA(cudaEventRecord(k.ev[1], s));
        const int rc = run_range(h, k, from, 0, n, s, prof);
        if (rc != U96_OK) return abort_submit(h, k, s, rc);
        if (disp_out)
            CKA(copy2d(disp_out, (size_t)W * 2, k.disp, (size_t)pitch * 2, (size_t)W * 2, (size_t)H * n, cudaMemcpyDeviceToHost, s,
                       device_src ? nullptr : stage_out(h, k)));
    } else {
        // chunks of whole BM waves (a chunk that fills half the SMs would make the kernels, not PCIe, the bottleneck)
        const int wave = bm_wave_frames(bm_config(h->bm));
        // two waves per chunk, one when two would exceed ~128 MB of input (C2 with 148-frame waves: one 296-frame chunk would leave
        // nothing to overlap inside a step; measured best against the PCIe ceiling: ~148 frames of 640x480)
        const size_t two_waves_in = (size_t)2 * wave * 2 * W * H;
        int csz = std::max((two_waves_in <= ((size_t)128 << 20)) ? 2 * wave : wave, (32 + wave - 1) / wave * wave);
        static const int csz_env = getenv("U96_CHUNK") ? atoi(getenv("U96_CHUNK")) : 0;     // developer override (frames per chunk)
        if (csz_env > 0) csz = csz_env;
        if (csz > n) csz = n;
        CKA(cudaEventRecord(k.done, s));                      // the sub-streams start behind whatever the bank stream holds
        for (int i = 0; i < 3; i++) CKA(cudaStreamWaitEvent(k.sub[i], k.done, 0));
        for (int c = 0, f0 = 0; f0 < n; c++, f0 += csz) {
            const int nf = std::min(csz, n - f0);
            cudaStream_t cs = k.sub[c % 3];
            for (int i = 0; i < 2; i++)
                CKA(copy2d(dstbuf[i] + (size_t)f0 * frame, pitch, src[i] + (size_t)f0 * stride * H, stride, W, (size_t)H * nf,
                           cudaMemcpyHostToDevice, cs, k.stage ? stage_in(h, k, i) + (size_t)f0 * W * H : nullptr));
            const int rc = run_range(h, k, from, f0, nf, cs, false);
            if (rc != U96_OK) return abort_submit(h, k, s, rc);
            if (disp_out)
                CKA(copy2d(disp_out + (size_t)f0 * W * H, (size_t)W * 2, k.disp + (size_t)f0 * frame, (size_t)pitch * 2,
                           (size_t)W * 2, (size_t)H * nf, cudaMemcpyDeviceToHost, cs,
                           k.stage ? stage_out(h, k) + (size_t)f0 * W * H * 2 : nullptr));
        }
        for (int i = 0; i < 3; i++) {                         // join: the bank stream (and `done`) follows all chunks
            CKA(cudaEventRecord(k.sub_ev[i], k.sub[i]));
            CKA(cudaStreamWaitEvent(s, k.sub_ev[i], 0));
        }
    }
    CKA(cudaGetLastError());
    CKA(cudaEventRecord(k.done, s));
    k.filled = true; k.has_disp = true; k.pending = true;
    h->fifo.push_back(bank);
    return U96_OK;
}

static int receive_u8(u96_handle *h, int bank, const uint8_t *const cur[2], int pitch, int min_from, uint8_t *L, uint8_t *R)
{
    if (!h || bank < 0 || bank > 1 || !L || !R) return U96_ERR_INVALID;
    Bank &k = h->bank[bank];
    if (!k.filled || k.from > min_from) return U96_ERR_STATE;
    if (min_from == FROM_XSBL && !k.has_disp) return U96_ERR_STATE;      // set_rect_image without start: no x-Sobel images yet
    CK(cudaSetDevice(h->device));
    cudaStream_t s = bank_stream(h, bank);
    uint8_t *dst[2] = {L, R};
    if (!k.pending) { const int rc = ensure_stage(h, k, k.W, k.H); if (rc != U96_OK) return rc; }
    for (int i = 0; i < 2; i++)
        CK(copy2d(dst[i], k.W, cur[i], pitch, k.W, (size_t)k.H * k.n, cudaMemcpyDeviceToHost, s, k.pending ? nullptr : stage_in(h, k, i)));
    CK(cudaStreamSynchronize(s));
    return U96_OK;
}

// kernel time of an auxiliary call (reproject / keypoints / uvc) when profiling is on
struct AuxTimer {
    u96_handle *h; int which; cudaStream_t s; bool on;
    AuxTimer(u96_handle *h_, int which_, cudaStream_t s_) : h(h_), which(which_), s(s_), on(h_->profiling) { if (on) cudaEventRecord(h->aux_ev[0], s); }
    void stop() { if (on) cudaEventRecord(h->aux_ev[1], s); }
    void read()
    {
        h->aux_valid[which] = false;
        if (on && cudaEventElapsedTime(&h->aux_ms[which], h->aux_ev[0], h->aux_ev[1]) == cudaSuccess) h->aux_valid[which] = true;
    }
};

extern "C" {

int u96_submit_raw(u96_handle *h, int bank, const uint8_t *L, const uint8_t *R, int stride, int n)
{ return submit_common(h, bank, FROM_RAW, L, R, stride, n, false); }
int u96_submit_rect(u96_handle *h, int bank, const uint8_t *L, const uint8_t *R, int stride, int n)
{ return submit_common(h, bank, FROM_RECT, L, R, stride, n, false); }
int u96_submit_xsbl(u96_handle *h, int bank, const uint8_t *L, const uint8_t *R, int stride, int n)
{ return submit_common(h, bank, FROM_XSBL, L, R, stride, n, false); }
int u96_submit_raw_async(u96_handle *h, int bank, const uint8_t *L, const uint8_t *R, int stride, int n, int16_t *disp_out)
{ return submit_common(h, bank, FROM_RAW, L, R, stride, n, false, disp_out); }
int u96_submit_rect_async(u96_handle *h, int bank, const uint8_t *L, const uint8_t *R, int stride, int n, int16_t *disp_out)
{ return submit_common(h, bank, FROM_RECT, L, R, stride, n, false, disp_out); }
int u96_submit_raw_device(u96_handle *h, int bank, const void *dL, const void *dR, int stride, int n)
{ return submit_common(h, bank, FROM_RAW, (const uint8_t *)dL, (const uint8_t *)dR, stride, n, true); }
int u96_submit_rect_device(u96_handle *h, int bank, const void *dL, const void *dR, int stride, int n)
{ return submit_common(h, bank, FROM_RECT, (const uint8_t *)dL, (const uint8_t *)dR, stride, n, true); }
int u96_submit_xsbl_device(u96_handle *h, int bank, const void *dL, const void *dR, int stride, int n)
{ return submit_common(h, bank, FROM_XSBL, (const uint8_t *)dL, (const uint8_t *)dR, stride, n, true); }

int u96_set_rect_image(u96_handle *h, int bank, const uint8_t *L, const uint8_t *R, int stride, int n)
{ return submit_common(h, bank, FROM_RECT, L, R, stride, n, false, nullptr, true); }

int u96_start_xsbl(u96_handle *h, int bank)
{
    if (!h || bank < 0 || bank > 1) return U96_ERR_INVALID;
    Bank &k = h->bank[bank];
    if (k.pending || !k.filled || k.from != FROM_RECT) return U96_ERR_STATE;
    if (k.W != h->bm.width || k.H != h->bm.height) return U96_ERR_STATE;      // the geometry changed since set_rect_image
    CK(cudaSetDevice(h->device));
    cudaStream_t s = bank_stream(h, bank);
    { const int rc = ensure_bank_buffers(h, k, FROM_RECT, k.n, false); if (rc != U96_OK) return rc; }
    k.has_eig = h->gftt;
    k.staged = false;
    { const int rc = ensure_border(h, k, s); if (rc != U96_OK) return abort_submit(h, k, s, rc); }
    const bool prof = h->profiling;
    if (prof) { CKA(cudaEventRecord(k.ev[0], s)); CKA(cudaEventRecord(k.ev[1], s)); }
    const int rc = run_range(h, k, FROM_RECT, 0, k.n, s, prof);
    if (rc != U96_OK) return abort_submit(h, k, s, rc);
    CKA(cudaGetLastError());
    CKA(cudaEventRecord(k.done, s));
    k.has_disp = true; k.pending = true;
    h->fifo.push_back(bank);
    return U96_OK;
}

int u96_wait(u96_handle *h, int *active_bank)
{
    if (!h) return U96_ERR_INVALID;
    if (h->fifo.empty()) return U96_ERR_STATE;
    const int b = h->fifo.front();
    CK(cudaSetDevice(h->device));
    CK(cudaEventSynchronize(h->bank[b].done));
    h->fifo.pop_front();
    h->bank[b].pending = false;
    if (active_bank) *active_bank = b;
    return U96_OK;
}

int u96_receive_rect(u96_handle *h, int bank, uint8_t *L, uint8_t *R)
{
    if (!h || bank < 0 || bank > 1) return U96_ERR_INVALID;
    Bank &k = h->bank[bank];
    return receive_u8(h, bank, k.cur_rect, k.rect_pitch, FROM_RECT, L, R);
}

int u96_receive_xsbl(u96_handle *h, int bank, uint8_t *L, uint8_t *R)
{
    if (!h || bank < 0 || bank > 1) return U96_ERR_INVALID;
    Bank &k = h->bank[bank];
    return receive_u8(h, bank, k.cur_xsbl, k.xsbl_pitch, FROM_XSBL, L, R);
}

int u96_receive_disp(u96_handle *h, int bank, int16_t *disp)
{
    if (!h || bank < 0 || bank > 1 || !disp) return U96_ERR_INVALID;
    Bank &k = h->bank[bank];
    if (!k.filled || !k.has_disp) return U96_ERR_STATE;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = bank_stream(h, bank);
    if (!k.pending) { const int rc = ensure_stage(h, k, k.W, k.H); if (rc != U96_OK) return rc; }
    CK(copy2d(disp, (size_t)k.W * 2, k.disp, (size_t)h->pitch * 2, (size_t)k.W * 2, (size_t)k.H * k.n,
                         cudaMemcpyDeviceToHost, s, k.pending ? nullptr : stage_out(h, k)));
    CK(cudaStreamSynchronize(s));
    return U96_OK;
}

int u96_receive_eigen(u96_handle *h, int bank, uint16_t *eig, uint16_t *max_eig)
{
    if (!h || bank < 0 || bank > 1 || !eig) return U96_ERR_INVALID;
    Bank &k = h->bank[bank];
    if (!k.filled || !k.has_eig || !k.has_disp) return U96_ERR_STATE;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = bank_stream(h, bank);
    CK(copy2d(eig, (size_t)k.W * 2, k.eig, (size_t)h->pitch * 2, (size_t)k.W * 2, (size_t)k.H * k.n, cudaMemcpyDeviceToHost, s));
    std::vector<uint32_t> mx(max_eig ? k.n : 0);
    if (max_eig) CK(cudaMemcpyAsync(mx.data(), k.eig_max, (size_t)k.n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int i = 0; i < (int)mx.size(); i++) max_eig[i] = (uint16_t)mx[i];
    return U96_OK;
}

int u96_enqueue_receive_disp(u96_handle *h, int bank, int16_t *disp)
{
    if (!h || bank < 0 || bank > 1 || !disp) return U96_ERR_INVALID;
    Bank &k = h->bank[bank];
    if (!k.filled || !k.pending) return U96_ERR_STATE;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = bank_stream(h, bank);
    CK(copy2d(disp, (size_t)k.W * 2, k.disp, (size_t)h->pitch * 2, (size_t)k.W * 2, (size_t)k.H * k.n,
                         cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(k.done, s));                          // wait() now covers the copy as well
    return U96_OK;
}

// StereoCameraModel.cpp:9-14
static const float kLocalTransform[12] = {0.0f, 0.0f, 1.0f, 0.0f, -1.0f, 0.0f, 0.0f, 0.0f, 0.0f, -1.0f, 0.0f, 0.0f};

int u96_reproject_ex(u96_handle *h, int bank, const double P_l[12], const double P_r[12], int decim,
                     const float *local_T, const float *poses, float *xyz)
{
    if (!h || bank < 0 || bank > 1 || !P_l || !P_r || !xyz) return U96_ERR_INVALID;
    if (decim != 1 && decim != 2 && decim != 4 && decim != 8) return U96_ERR_INVALID;
    Bank &k = h->bank[bank];
    if (!k.filled || !k.has_disp) return U96_ERR_STATE;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = bank_stream(h, bank);
    const size_t count = (size_t)(k.W / decim) * (k.H / decim) * k.n * 3;
    if (count > h->xyz_cap) {
        cudaFree(h->xyz); h->xyz = nullptr; h->xyz_cap = 0;
        if (cudaMalloc(&h->xyz, count * sizeof(float)) != cudaSuccess) return U96_ERR_NOMEM;
        h->xyz_cap = count;
    }
    if (poses) {
        if (!h->poses && cudaMalloc(&h->poses, (size_t)h->maxB * 12 * sizeof(float)) != cudaSuccess) return U96_ERR_NOMEM;
        CK(cudaMemcpyAsync(h->poses, poses, (size_t)k.n * 12 * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    AuxTimer t(h, U96_AUX_REPROJECT, s);
    h->launches += launch_reproject(k.disp, h->pitch, (size_t)h->pitch * k.H, k.W, k.H, k.n, P_l, P_r, decim, local_T,
                                    poses ? h->poses : nullptr, h->xyz, s);
    t.stop();
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(xyz, h->xyz, count * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    t.read();
    return U96_OK;
}

int u96_reproject(u96_handle *h, int bank, const double P_l[12], const double P_r[12], int decim, int flags, float *xyz)
{ return u96_reproject_ex(h, bank, P_l, P_r, decim, (flags & 1) ? kLocalTransform : nullptr, nullptr, xyz); }

int u96_reproject_points(u96_handle *h, int bank, int frame, const double P_l[12], const double P_r[12], const float *uv, int n,
                         const uint8_t *mask, float min_depth, float max_depth, const float *local_T, float *xyz)
{
    if (!h || bank < 0 || bank > 1 || !P_l || !P_r || n < 0 || (n > 0 && (!uv || !xyz))) return U96_ERR_INVALID;
    Bank &k = h->bank[bank];
    if (!k.filled || !k.has_disp) return U96_ERR_STATE;
    if (frame < 0 || frame >= k.n) return U96_ERR_INVALID;
    if (n == 0) return U96_OK;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = bank_stream(h, bank);
    const size_t need = (size_t)n * 6;                       // floats: uv | xyz | mask bytes (rounded up to n/4 floats + 1)
    const size_t words = need + (size_t)n / 4 + 1;
    if (words > h->kp_cap) {
        cudaFree(h->kp); h->kp = nullptr; h->kp_cap = 0;
        const size_t cap = std::max(words, (size_t)8192);
        if (cudaMalloc(&h->kp, cap * sizeof(float)) != cudaSuccess) return U96_ERR_NOMEM;
        h->kp_cap = cap;
    }
    float *d_uv = h->kp, *d_xyz = h->kp + (size_t)2 * n;
    uint8_t *d_mask = reinterpret_cast<uint8_t *>(h->kp + (size_t)5 * n);
    CK(cudaMemcpyAsync(d_uv, uv, (size_t)n * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    if (mask) CK(cudaMemcpyAsync(d_mask, mask, (size_t)n, cudaMemcpyHostToDevice, s));
    AuxTimer t(h, U96_AUX_REPROJECT_POINTS, s);
    h->launches += launch_reproject_points(k.disp + (size_t)frame * h->pitch * k.H, h->pitch, k.W, k.H, P_l, P_r, d_uv,
                                           mask ? d_mask : nullptr, n, min_depth, max_depth, local_T, d_xyz, s);
    t.stop();
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(xyz, d_xyz, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    t.read();
    return U96_OK;
}

int u96_receive_uvc(u96_handle *h, int bank, int which, uint8_t *frame)
{
    if (!h || bank < 0 || bank > 1 || !frame) return U96_ERR_INVALID;
    if (which != U96_UVC_RECT && which != U96_UVC_XSBL && which != U96_UVC_BM) return U96_ERR_INVALID;
    Bank &k = h->bank[bank];
    if (!k.filled) return U96_ERR_STATE;
    if ((which == U96_UVC_RECT && k.from > FROM_RECT)) return U96_ERR_STATE;      // the bank holds no rectified images
    if (which != U96_UVC_RECT && !k.has_disp) return U96_ERR_STATE;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = bank_stream(h, bank);
    const int W = k.W, H = k.H;
    const size_t bytes = (size_t)k.n * H * W * 4;
    if (bytes > h->uvc_cap) {
        cudaFree(h->uvc); h->uvc = nullptr; h->uvc_cap = 0;
        if (cudaMalloc(&h->uvc, bytes) != cudaSuccess) return U96_ERR_NOMEM;
        h->uvc_cap = bytes;
    }
    AuxTimer t(h, U96_AUX_UVC, s);
    if (which == U96_UVC_BM)
        h->launches += launch_pack_uvc(nullptr, nullptr, 0, 0, k.disp, h->pitch, (size_t)h->pitch * H, h->uvc, W, H, k.n, s);
    else if (which == U96_UVC_RECT)
        h->launches += launch_pack_uvc(k.cur_rect[0], k.cur_rect[1], k.rect_pitch, k.rect_frame, nullptr, 0, 0, h->uvc, W, H, k.n, s);
    else
        h->launches += launch_pack_uvc(k.cur_xsbl[0], k.cur_xsbl[1], k.xsbl_pitch, k.xsbl_frame, nullptr, 0, 0, h->uvc, W, H, k.n, s);
    t.stop();
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(frame, h->uvc, bytes, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    t.read();
    return U96_OK;
}

int u96_bank_device_ptr(u96_handle *h, int bank, int which, void **dptr, int *pitch_bytes, size_t *frame_bytes)
{
    if (!h || bank < 0 || bank > 1 || !dptr || which < 0 || which >= U96_BUF_COUNT) return U96_ERR_INVALID;
    Bank &k = h->bank[bank];
    const int H = k.filled ? k.H : h->bm.height;
    const void *p = nullptr; int pitch = h->pitch; size_t frame = (size_t)h->pitch * H;
    switch (which) {
    case U96_BUF_RAW_L: case U96_BUF_RAW_R: p = k.raw[which - U96_BUF_RAW_L]; break;
    case U96_BUF_RECT_L: case U96_BUF_RECT_R: p = k.rect[which - U96_BUF_RECT_L]; break;
    case U96_BUF_XSBL_L: case U96_BUF_XSBL_R: p = k.xsbl[which - U96_BUF_XSBL_L]; break;
    case U96_BUF_DISP: p = k.disp; pitch *= 2; frame *= 2; break;
    case U96_BUF_EIG: p = k.eig; pitch *= 2; frame *= 2; break;
    }
    *dptr = const_cast<void *>(p);
    if (pitch_bytes) *pitch_bytes = pitch;
    if (frame_bytes) *frame_bytes = frame;
    return U96_OK;
}

int u96_host_alloc(void **p, size_t bytes)
{
    if (!p) return U96_ERR_INVALID;
    CK(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    return U96_OK;
}

int u96_host_alloc_wc(void **p, size_t bytes)
{
    if (!p) return U96_ERR_INVALID;
    CK(cudaHostAlloc(p, bytes, cudaHostAllocWriteCombined));
    return U96_OK;
}

int u96_host_free(void *p)
{
    CK(cudaFreeHost(p));
    return U96_OK;
}

int u96_last_stage_ms_ex(u96_handle *h, int bank, float *ms, int count)
{
    if (!h || bank < 0 || bank > 1 || !ms || count < 1 || count > U96_STAGE_COUNT) return U96_ERR_INVALID;
    Bank &k = h->bank[bank];
    if (!h->profiling || !k.filled || !k.has_disp || k.pending || k.staged) return U96_ERR_STATE;
    for (int i = 0; i < count; i++) CK(cudaEventElapsedTime(&ms[i], k.ev[i], k.ev[i + 1]));
    return U96_OK;
}

int u96_last_stage_ms(u96_handle *h, int bank, float ms[4])
{
    if (!ms) return U96_ERR_INVALID;
    float v[U96_STAGE_COUNT];
    const int rc = u96_last_stage_ms_ex(h, bank, v, U96_STAGE_COUNT);
    if (rc != U96_OK) return rc;
    ms[0] = v[U96_STAGE_H2D];
    ms[1] = v[U96_STAGE_RECT] + v[U96_STAGE_GFTT];
    ms[2] = v[U96_STAGE_XSBL];
    ms[3] = v[U96_STAGE_BM] + v[U96_STAGE_POST];
    return U96_OK;
}

int u96_last_aux_ms(u96_handle *h, int which, float *ms)
{
    if (!h || !ms || which < 0 || which >= U96_AUX_COUNT) return U96_ERR_INVALID;
    if (!h->aux_valid[which]) return U96_ERR_STATE;
    *ms = h->aux_ms[which];
    return U96_OK;
}

int u96_microbench(int device, int which, double *gops)
{
    if (!gops) return U96_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return U96_ERR_NODEVICE;
    CK(cudaSetDevice(device));
    return run_microbench(which, gops);
}

}  // extern "C"
