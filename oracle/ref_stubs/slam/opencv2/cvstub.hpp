// cvstub.hpp -- the few OpenCV core types the reference's Stereo.cpp / StereoCameraModel.cpp / Transform.cpp touch,
// so that those three files compile UNMODIFIED from /root/reference without the OpenCV C++ SDK (absent in this image).
// TEST INFRASTRUCTURE (oracle/_ref build only).  Only storage and element access are provided -- no arithmetic of the
// reference is restated here: every floating-point operation of the 3-D path runs in the reference's own source.
#pragma once
#include <cstddef>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8UC1 0
#define CV_16SC1 3
#define CV_32F 5
#define CV_32FC1 5
#define CV_64FC1 6

namespace cv {

template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f;
template <typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<float> Point3f;
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };

class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char *data = nullptr;
    size_t step = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void *ext) : rows(r), cols(c), data((unsigned char *)ext), step((size_t)c * esz(type)), type_(type) {}
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); std::memset(m.data, 0, m.step * r); return m; }
    int type() const { return type_; }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    size_t elemSize() const { return esz(type_); }
    template <typename T> T &at(int r, int c) { return *reinterpret_cast<T *>(data + step * r + sizeof(T) * c); }
    template <typename T> const T &at(int r, int c) const { return *reinterpret_cast<const T *>(data + step * r + sizeof(T) * c); }
    Mat clone() const
    {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++) std::memcpy(m.data + m.step * r, data + step * r, (size_t)cols * esz(type_));
        return m;
    }
    Mat colRange(int a, int b) const { Mat m = *this; m.cols = b - a; m.data = data + esz(type_) * a; return m; }
    void convertTo(Mat &dst, int type) const
    {
        dst.create(rows, cols, type);
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++) {
                const double v = (type_ == CV_64FC1) ? at<double>(r, c) : (type_ == CV_32FC1) ? (double)at<float>(r, c) : 0.0;
                if (type == CV_32FC1) dst.at<float>(r, c) = (float)v; else dst.at<double>(r, c) = v;
            }
    }
private:
    static size_t esz(int t) { return t == CV_8UC1 ? 1 : t == CV_16SC1 ? 2 : t == CV_32FC1 ? 4 : 8; }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type; step = (size_t)c * esz(type);
        buf_ = std::shared_ptr<unsigned char>(new unsigned char[step * r + 8], std::default_delete<unsigned char[]>());
        data = buf_.get();
    }
    int type_ = 0;
    std::shared_ptr<unsigned char> buf_;       // copies share the pixels, like cv::Mat
};

inline int countNonZero(const Mat &m)
{
    int n = 0;
    for (int r = 0; r < m.rows; r++)
        for (int c = 0; c < m.cols; c++) n += (m.type() == CV_32FC1 ? m.at<float>(r, c) != 0.0f : m.at<double>(r, c) != 0.0);
    return n;
}

struct KeyPoint {
    Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
    static void convert(const std::vector<KeyPoint> &k, std::vector<Point2f> &p)
    { p.resize(k.size()); for (size_t i = 0; i < k.size(); i++) p[i] = k[i].pt; }
};
struct TermCriteria { enum { COUNT = 1, MAX_ITER = 1, EPS = 2 }; int type, maxCount; double epsilon;
                      TermCriteria(int t = 0, int m = 0, double e = 0) : type(t), maxCount(m), epsilon(e) {} };
enum { OPTFLOW_LK_GET_MIN_EIGENVALS = 8 };
// array proxies: only passed through to calcOpticalFlowPyrLKStereo, which the dense-disparity depth methods never call
struct _AnyArray { _AnyArray() {} template <typename T> _AnyArray(const T &) {} };
typedef const _AnyArray &InputArray;
typedef const _AnyArray &OutputArray;
typedef const _AnyArray &InputOutputArray;

// the OpenCV-yml calibration route needs a YAML parser; not provided: open() fails and StereoCameraModel::load returns false.
// The KITTI route of the same function is plain fscanf and works.
struct FileNode {
    enum { NONE = 0 };
    int type() const { return NONE; }
    FileNode operator[](const char *) const { return FileNode(); }
    operator int() const { return 0; }
    template <typename T> void operator>>(T &) const {}
};
struct FileStorage {
    enum { READ = 0 };
    bool open(const std::string &, int) { return false; }
    FileNode operator[](const char *) const { return FileNode(); }
    void release() {}
};

}  // namespace cv
