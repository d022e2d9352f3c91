// stereo_ref_shim.cpp -- C entry points around the reference's OWN 3-D code, compiled unmodified from /root/reference:
//   src/slam/src/core/Stereo.cpp            (projectDisparityTo3D :157-182, generateKeypoints3DStereo :53-117,
//                                            isFinite :184-187, transformPoint :189-198)
//   src/slam/src/core/StereoCameraModel.cpp (constructor = localTransform :8-14, KITTI calib loader + 640x480 rescale :68-119)
//   src/slam/src/core/Transform.cpp         (storage of the 3x4 float transform, isNull :88-95)
// TEST INFRASTRUCTURE: loaded by tests/ only, to pin oracle/u96_oracle.c's restatement of SURVEY 8(a) row a16.
// Nothing here computes: the loops below only feed the reference's functions, in the order its callers do
// (dense: main.cpp:522-551 on the SensorData.cpp:50-58 decimated map; keypoints: Stereo.cpp:119-154).
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <limits>

#include "core/Stereo.h"
#include "core/Logger.h"
#include "opencv/CvLKStereo.h"

// Logger.cpp spins forever on LOG_ERROR and needs the application's global settings; warnings are counted instead
static int g_warnings = 0;
void log_write(LOG_LEVEL, const char *, int, const char *, const char *, ...) { g_warnings++; }
// sparse LK depth method (DEPTH_METHOD_CV_LK) is not part of the dense path; referenced by Stereo.cpp:26-38 only
void calcOpticalFlowPyrLKStereo(cv::InputArray, cv::InputArray, cv::InputArray, cv::InputOutputArray, cv::OutputArray, cv::OutputArray,
                                cv::Size, int, cv::TermCriteria, int, double) {}

extern "C" {

int ref_warnings(void) { return g_warnings; }

// StereoCameraModel::load on a KITTI-style calib.txt -> {fx,fy,cx,cy,Tx} of the left then the right camera
int ref_model_load(const char *calib, int doResize, double out[10])
{
    StereoCameraModel m;
    if (!m.load(calib, "", doResize)) return -1;
    const double v[10] = {m.fx_l(), m.fy_l(), m.cx_l(), m.cy_l(), m.Tx_l(), m.fx_r(), m.fy_r(), m.cx_r(), m.cy_r(), m.Tx_r()};
    for (int i = 0; i < 10; i++) out[i] = v[i];
    return 0;
}

// the model's localTransform as 12 floats (row-major 3x4)
int ref_local_transform(float out[12])
{
    StereoCameraModel m;
    const Transform &t = m.localTransform();
    const float v[12] = {t.r11(), t.r12(), t.r13(), t.o14(), t.r21(), t.r22(), t.r23(), t.o24(), t.r31(), t.r32(), t.r33(), t.o34()};
    for (int i = 0; i < 12; i++) out[i] = v[i];
    return t.isNull() ? 1 : 0;
}

// generateKeypoints3DStereo (Stereo.cpp:53-117) with the dense-map depth methods, as generateKeypoints3D calls it
int ref_keypoints3d(const char *calib, int doResize, const float *uv, int n, const int16_t *disp, int W, int H,
                    float minDepth, float maxDepth, int depthMethod, float *xyz)
{
    StereoCameraModel m;
    if (!m.load(calib, "", doResize)) return -1;
    std::vector<cv::Point2f> left((size_t)n), right;
    for (int i = 0; i < n; i++) left[i] = cv::Point2f(uv[2 * i], uv[2 * i + 1]);
    std::vector<unsigned char> mask;
    cv::Mat d(H, W, CV_16SC1, const_cast<int16_t *>(disp));
    const std::vector<cv::Point3f> p = generateKeypoints3DStereo(left, right, m, mask, minDepth, maxDepth, d, depthMethod);
    for (int i = 0; i < n; i++) { xyz[3 * i] = p[i].x; xyz[3 * i + 1] = p[i].y; xyz[3 * i + 2] = p[i].z; }
    return 0;
}

// the dense consumer's inner loop (main.cpp:522-551): every sample of the decimated map -> projectDisparityTo3D ->
// isFinite -> localTransform -> pose; samples the reference skips come back as NaN
int ref_dense_cloud(const char *calib, int doResize, const int16_t *depth, int rows, int cols, int scale, int apply_local,
                    const float *pose12, float *xyz)
{
    StereoCameraModel m;
    if (!m.load(calib, "", doResize)) return -1;
    const float nan = std::numeric_limits<float>::quiet_NaN();
    Transform pose;
    if (pose12) pose = Transform(pose12[0], pose12[1], pose12[2], pose12[3], pose12[4], pose12[5], pose12[6], pose12[7],
                                 pose12[8], pose12[9], pose12[10], pose12[11]);
    cv::Mat dm(rows, cols, CV_16SC1, const_cast<int16_t *>(depth));
    for (int row = 0; row < rows; row++)
        for (int col = 0; col < cols; col++) {
            float *o = xyz + ((size_t)row * cols + col) * 3;
            o[0] = o[1] = o[2] = nan;
            float disparity = (float)(dm.at<short>(row, col) / 16.0f);
            if (disparity > 0) {
                cv::Point2f pt2d = cv::Point2f((float)(col * scale), (float)(row * scale));
                cv::Point3f pt3d = projectDisparityTo3D(pt2d, disparity, m);
                if (isFinite(pt3d)) {
                    if (apply_local) pt3d = transformPoint(pt3d, m.localTransform());
                    if (pose12) pt3d = transformPoint(pt3d, pose);
                    o[0] = pt3d.x; o[1] = pt3d.y; o[2] = pt3d.z;
                }
            }
        }
    return 0;
}

}  // extern "C"
