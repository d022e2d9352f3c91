// Fpga.hpp -- C++ host shim with the data-plane interface of the reference's `class Fpga`
// (slam/include/core/FPGA.h:347-397, slam/src/core/FPGA.cpp) over the C ABI in include/u96_stereo.h.
//
// A src/slam-style caller keeps its code: registerOpen()/memoryOpen(), setRectImage(bank, L, R) followed by
// the software start (reference: `reg->xsbl.Control |= FPGA_XSBL_SW_START`, main.cpp:172-174 -> here
// startXsbl(bank)), receiveData()/receiveRectImages()/receiveDepthMap() which COPY OUT of the bank like
// cv::Mat::clone() does in the reference.  Images are passed as u96::Mat8/Mat16 (rows, cols, data) so the shim
// has no OpenCV dependency; with OpenCV present `cv::Mat(rows, cols, CV_8UC1, m.data.data())` wraps them.
// Unlike the reference (exit(1) on mmap failure, busy-wait on the mailbox, LOG_ERROR spins) every method
// returns 0 / -1 and never blocks forever.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/u96_stereo.h"

namespace u96 {

struct Mat8 {  int rows = 0, cols = 0; std::vector<uint8_t> data;
               Mat8() = default; Mat8(int r, int c) : rows(r), cols(c), data((size_t)r * c) {}
               bool empty() const { return data.empty(); } };
struct Mat16 { int rows = 0, cols = 0; std::vector<int16_t> data;
               Mat16() = default; Mat16(int r, int c) : rows(r), cols(c), data((size_t)r * c) {}
               short at(int y, int x) const { return data[(size_t)y * cols + x]; } };

// SensorData.cpp:50-58 -- keep every `scale`-th sample of the disparity map
inline Mat16 decimateDisparity(const Mat16 &d, int scale = 4)
{
    Mat16 o(d.rows / scale, d.cols / scale);
    for (int r = 0; r < o.rows; r++)
        for (int c = 0; c < o.cols; c++) o.data[(size_t)r * o.cols + c] = d.at(r * scale, c * scale);
    return o;
}

class Fpga {
public:
    static constexpr int IMAGE_WIDTH = 640, IMAGE_HEIGHT = 480;      // FPGA.h:55-56

    explicit Fpga(int device = 0) : device_(device) {}
    ~Fpga() { registerClose(); }
    Fpga(const Fpga &) = delete;
    Fpga &operator=(const Fpga &) = delete;

    // FPGA.cpp:27-60 -- returns 0 / -1
    int registerOpen()
    {
        if (h_) return 0;
        if (u96_create(&h_, device_, IMAGE_WIDTH, IMAGE_HEIGHT, 1) != U96_OK) { h_ = nullptr; return -1; }
        // Fpga_Init BM block (StereoBM/src/fpga.c:150-160): 640x480, window 21, 64 disparities, uniqueness off
        if (u96_set_bm_registers(h_, (IMAGE_HEIGHT << 16) + IMAGE_WIDTH, 0x00150040u, 0u) != U96_OK) return -1;
        return 0;
    }
    int registerClose() { if (h_) { u96_destroy(h_); h_ = nullptr; } return 0; }
    int memoryOpen() { return h_ ? 0 : -1; }        // the banks live in HBM inside the handle
    int memoryClose() { return 0; }

    // set_rect_param (StereoBM/src/fpga.c:267-301)
    int setRectParam(const u96_rect_params &p) { return u96_set_rect_params(h_, &p) == U96_OK ? 0 : -1; }
    // UniFiltCtrl register (bm.v:183-187)
    int setUniquenessFilter(bool enable, int mode, int thr)
    {
        return u96_set_bm_registers(h_, (IMAGE_HEIGHT << 16) + IMAGE_WIDTH, 0x00150040u,
                                    ((uint32_t)enable << 31) | ((uint32_t)(mode & 1) << 16) | (uint32_t)(thr & 0x3FF)) == U96_OK ? 0 : -1;
    }

    // FPGA.cpp:236-249 -- copies the rectified pair into the RECT bank, at call time like the reference's memcpy:
    // receiveRectImages(bank) right after it reads the pair back, and the caller's images may be reused at once.
    // (void in the reference; the status is returned here because nothing in this library exits or spins.)
    int setRectImage(int bank, const Mat8 &imageLeft, const Mat8 &imageRight)
    {
        if (imageLeft.empty() || imageRight.empty() || imageLeft.cols != IMAGE_WIDTH || imageLeft.rows != IMAGE_HEIGHT ||
            imageRight.cols != IMAGE_WIDTH || imageRight.rows != IMAGE_HEIGHT) return -1;
        return u96_set_rect_image(h_, bank, imageLeft.data.data(), imageRight.data.data(), IMAGE_WIDTH, 1) == U96_OK ? 0 : -1;
    }
    // main.cpp:172-174 `reg->xsbl.Control |= FPGA_XSBL_SW_START`: xsbl -> bm on what the RECT bank holds
    int startXsbl(int bank) { return u96_start_xsbl(h_, bank) == U96_OK ? 0 : -1; }
    // sensor path (CameraStereoImages.cpp:134-149): raw pair -> rect -> xsbl -> bm
    int captureRaw(int bank, const Mat8 &rawLeft, const Mat8 &rawRight)
    {
        return u96_submit_raw(h_, bank, rawLeft.data.data(), rawRight.data.data(), rawLeft.cols, 1) == U96_OK ? 0 : -1;
    }
    // waitIpcMessage(IPC_MSG2_DATA_READY) + IpcParameter2 (FPGA.cpp:217-220, 314); returns the active bank or -1
    int waitDataReady() { int b = -1; return u96_wait(h_, &b) == U96_OK ? b : -1; }

    // FPGA.cpp:251-268
    int receiveRectImages(int bank, Mat8 &matLeft, Mat8 &matRight)
    {
        Mat8 l(IMAGE_HEIGHT, IMAGE_WIDTH), r(IMAGE_HEIGHT, IMAGE_WIDTH);
        if (u96_receive_rect(h_, bank, l.data.data(), r.data.data()) != U96_OK) return -1;
        matLeft = std::move(l); matRight = std::move(r);
        return 0;
    }
    // FPGA.cpp:270-279 -- CV_16SC1, 16x fixed-point disparity
    int receiveDepthMap(int bank, Mat16 &matDepth)
    {
        Mat16 d(IMAGE_HEIGHT, IMAGE_WIDTH);
        if (u96_receive_disp(h_, bank, d.data.data()) != U96_OK) return -1;
        matDepth = std::move(d);
        return 0;
    }
    // fpga->gftt.Control = FPGA_GFTT_CTRL_ENABLE (StereoBM/src/fpga.c:162-172)
    int enableGftt(bool on) { return u96_set_gftt(h_, on ? 1 : 0) == U96_OK ? 0 : -1; }
    // FPGA.cpp:281-296 -- CV_16UC1 min-eigenvalue map + the bank's half of reg->gftt.Max
    int receiveEigen(int bank, std::vector<uint16_t> &matEigen, unsigned short *maxEigen)
    {
        std::vector<uint16_t> e((size_t)IMAGE_HEIGHT * IMAGE_WIDTH);
        uint16_t mx = 0;
        if (u96_receive_eigen(h_, bank, e.data(), &mx) != U96_OK) return -1;
        matEigen = std::move(e);
        if (maxEigen) *maxEigen = mx;
        return 0;
    }
    // FPGA.cpp:310-347 (stereo part): wait, then copy the active bank out
    int receiveData(Mat8 &rectLeft, Mat8 &rectRight, Mat16 &depth)
    {
        const int bank = waitDataReady();
        if (bank < 0) return -1;
        if (receiveRectImages(bank, rectLeft, rectRight) != 0 || receiveDepthMap(bank, depth) != 0) return -1;
        return bank;
    }
    // Stereo.cpp:157-182 over the decimated map (main.cpp:522-551); xyz = (H/decim)*(W/decim)*3 floats
    int projectDisparityTo3D(int bank, const double P_l[12], const double P_r[12], int decim, bool localTransform, std::vector<float> &xyz)
    {
        xyz.resize((size_t)(IMAGE_HEIGHT / decim) * (IMAGE_WIDTH / decim) * 3);
        return u96_reproject(h_, bank, P_l, P_r, decim, localTransform ? 1 : 0, xyz.data()) == U96_OK ? 0 : -1;
    }
    // generateKeypoints3D (Stereo.cpp:119-154, called per frame at main.cpp:250-252) with DEPTH_METHOD_FPGA_BM: keypoints =
    // n (x, y) float pairs in the left rectified image -> n (X, Y, Z) in the body frame (localTransform applied), NaN = bad
    // point; minDepth = maxDepth = 0 like the reference's call
    int generateKeypoints3D(int bank, const double P_l[12], const double P_r[12], const std::vector<float> &keypoints,
                            std::vector<float> &kpts3d, float minDepth = 0.0f, float maxDepth = 0.0f)
    {
        static const float local[12] = {0.f, 0.f, 1.f, 0.f, -1.f, 0.f, 0.f, 0.f, 0.f, -1.f, 0.f, 0.f};   // StereoCameraModel.cpp:9-14
        const int n = (int)(keypoints.size() / 2);
        kpts3d.resize((size_t)n * 3);
        return u96_reproject_points(h_, bank, 0, P_l, P_r, keypoints.data(), n, nullptr, minDepth, maxDepth, local, kpts3d.data()) == U96_OK ? 0 : -1;
    }
    u96_handle *handle() { return h_; }

private:
    int device_;
    u96_handle *h_ = nullptr;
};

}  // namespace u96
