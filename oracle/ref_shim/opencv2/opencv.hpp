// TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the few OpenCV 2.4.9 types and functions
// that the reference's path files use, so that those files (CDecodeGray.cpp, CDecodePhase.cpp,
// CCalculation.cpp, CSensorV.cpp, GlobalFunction.cpp, compiled IN PLACE from /root/reference by
// oracle/Makefile, never copied) build with plain g++ where no OpenCV C++ exists.
//
// What is the reference's own code in the resulting oracle/_ref binary: every loop, type,
// conversion and operation order of the path.  What is restated here (third-party, not under
// /root/reference): the cv::Mat container, Mat - Mat / Mat + Mat / Mat * Mat, FileStorage's
// matrix reader, imread's BMP decoder, blur's 3x3 box filter and cvFastArctan -- each pinned
// against the container's cv2 4.13 in tests/ (see oracle/sl_oracle.h, oracle/bmp_oracle.py).
// One deliberate difference: Mat::create zero-fills (OpenCV leaves new memory uninitialised and
// the reference reads such memory for pixels without a decoded U, CCalculation.cpp:678-682).
#ifndef REF_SHIM_OPENCV_HPP_
#define REF_SHIM_OPENCV_HPP_

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8UC1 0
#define CV_16UC1 2
#define CV_32FC1 5
#define CV_64FC1 6
#define CV_8UC3 16
#define CV_PI 3.1415926535897932384626433832795
#define CV_LOAD_IMAGE_GRAYSCALE 0

namespace cv {

struct Size {
    int width = 0, height = 0;
    Size() {}
    Size(int w, int h) : width(w), height(h) {}
};

struct Range {
    int start = 0, end = 0;
    Range() {}
    Range(int s, int e) : start(s), end(e) {}
};

inline size_t shim_elem_size(int type)
{
    switch (type) {
    case CV_8UC1: return 1;
    case CV_16UC1: return 2;
    case CV_32FC1: return 4;
    case CV_64FC1: return 8;
    case CV_8UC3: return 3;
    default: std::fprintf(stderr, "ref_shim: unsupported Mat type %d\n", type); std::abort();
    }
}

class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    size_t step = 0;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }

    void create(int r, int c, int type)
    {
        if (data && r == rows && c == cols && type == type_) return;      // same shape: keep (as cv::Mat::create)
        rows = r; cols = c; type_ = type;
        step = (size_t)c * shim_elem_size(type);
        const size_t bytes = step * (size_t)r;
        owner_.reset(static_cast<uchar*>(std::calloc(bytes ? bytes : 1, 1)), std::free);
        data = owner_.get();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }

    int type() const { return type_; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    size_t elemSize() const { return shim_elem_size(type_); }

    template <typename T> T& at(int i, int j) { return *reinterpret_cast<T*>(data + (size_t)i * step + (size_t)j * sizeof(T)); }
    template <typename T> const T& at(int i, int j) const { return *reinterpret_cast<const T*>(data + (size_t)i * step + (size_t)j * sizeof(T)); }
    uchar* ptr(int i = 0) { return data + (size_t)i * step; }
    const uchar* ptr(int i = 0) const { return data + (size_t)i * step; }

    Mat colRange(int a, int b) const { Mat m(*this); m.cols = b - a; m.data = data + (size_t)a * elemSize(); return m; }
    Mat colRange(const Range& r) const { return colRange(r.start, r.end); }
    Mat rowRange(int a, int b) const { Mat m(*this); m.rows = b - a; m.data = data + (size_t)a * step; return m; }
    Mat rowRange(const Range& r) const { return rowRange(r.start, r.end); }

    void copyTo(Mat& dst) const
    {
        if (empty()) { dst = Mat(); return; }
        dst.create(rows, cols, type_);
        copy_rows(dst);
    }
    void copyTo(Mat&& view) const      // a temporary header onto existing memory (m.colRange(...))
    {
        if (view.rows != rows || view.cols != cols || view.type_ != type_) {
            std::fprintf(stderr, "ref_shim: copyTo into a view of another shape\n");
            std::abort();
        }
        copy_rows(view);
    }
    Mat clone() const { Mat m; copyTo(m); return m; }

    Mat& setTo(double v)
    {
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) {
                switch (type_) {
                case CV_8UC1: at<uchar>(i, j) = (uchar)v; break;
                case CV_16UC1: at<ushort>(i, j) = (ushort)v; break;
                case CV_32FC1: at<float>(i, j) = (float)v; break;
                case CV_64FC1: at<double>(i, j) = v; break;
                default: std::abort();
                }
            }
        return *this;
    }
    void release() { *this = Mat(); }

private:
    void copy_rows(Mat& dst) const
    {
        const size_t row_bytes = (size_t)cols * elemSize();
        for (int i = 0; i < rows; i++) std::memcpy(dst.data + (size_t)i * dst.step, data + (size_t)i * step, row_bytes);
    }
    int type_ = CV_8UC1;
    std::shared_ptr<uchar> owner_;
};

inline void shim_check_same(const Mat& a, const Mat& b)
{
    if (a.rows != b.rows || a.cols != b.cols || a.type() != b.type()) {
        std::fprintf(stderr, "ref_shim: operands of different shape / type\n");
        std::abort();
    }
}

// cv::subtract / cv::add semantics per depth: u8 saturates, floating types are plain IEEE
inline Mat operator-(const Mat& a, const Mat& b)
{
    shim_check_same(a, b);
    Mat d(a.rows, a.cols, a.type());
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) {
            switch (a.type()) {
            case CV_8UC1: { const int v = (int)a.at<uchar>(i, j) - (int)b.at<uchar>(i, j); d.at<uchar>(i, j) = (uchar)(v < 0 ? 0 : v); break; }
            case CV_32FC1: d.at<float>(i, j) = a.at<float>(i, j) - b.at<float>(i, j); break;
            case CV_64FC1: d.at<double>(i, j) = a.at<double>(i, j) - b.at<double>(i, j); break;
            default: std::abort();
            }
        }
    return d;
}

inline Mat operator+(const Mat& a, const Mat& b)
{
    shim_check_same(a, b);
    Mat d(a.rows, a.cols, a.type());
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) {
            switch (a.type()) {
            case CV_8UC1: { const int v = (int)a.at<uchar>(i, j) + (int)b.at<uchar>(i, j); d.at<uchar>(i, j) = (uchar)(v > 255 ? 255 : v); break; }
            case CV_32FC1: d.at<float>(i, j) = a.at<float>(i, j) + b.at<float>(i, j); break;
            case CV_64FC1: d.at<double>(i, j) = a.at<double>(i, j) + b.at<double>(i, j); break;
            default: std::abort();
            }
        }
    return d;
}

// matrix product (cv::gemm, CV_64F): every element accumulates a(i,k)*b(k,j) over k in order
inline Mat operator*(const Mat& a, const Mat& b)
{
    if (a.type() != CV_64FC1 || b.type() != CV_64FC1 || a.cols != b.rows) {
        std::fprintf(stderr, "ref_shim: Mat * Mat needs conforming CV_64FC1 operands\n");
        std::abort();
    }
    Mat d(a.rows, b.cols, CV_64FC1);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < b.cols; j++) {
            double s = 0;
            for (int k = 0; k < a.cols; k++) s += a.at<double>(i, k) * b.at<double>(k, j);
            d.at<double>(i, j) = s;
        }
    return d;
}

// ---- FileStorage: the OpenCV YAML matrix reader (CCalculation.cpp:124-132) -------------------
class FileNode {
public:
    FileNode() {}
    FileNode(int r, int c, std::vector<double> v) : rows_(r), cols_(c), vals_(std::move(v)) {}
    void read(Mat& m) const
    {
        if ((size_t)rows_ * cols_ != vals_.size() || vals_.empty()) { m = Mat(); return; }
        m.create(rows_, cols_, CV_64FC1);
        for (int i = 0; i < rows_; i++)
            for (int j = 0; j < cols_; j++) m.at<double>(i, j) = vals_[(size_t)i * cols_ + j];
    }
private:
    int rows_ = 0, cols_ = 0;
    std::vector<double> vals_;
};
inline void operator>>(const FileNode& n, Mat& m) { n.read(m); }

class FileStorage {
public:
    enum { READ = 0 };
    FileStorage(const std::string& path, int)
    {
        std::ifstream f(path.c_str());
        if (!f) return;
        std::stringstream ss;
        ss << f.rdbuf();
        text_ = ss.str();
    }
    bool isOpened() const { return !text_.empty(); }
    void release() { text_.clear(); }
    FileNode operator[](const char* key) const
    {
        const std::string tag = std::string(key) + ":";
        size_t p = 0;
        while ((p = text_.find(tag, p)) != std::string::npos) {
            if (p == 0 || text_[p - 1] == '\n') break;
            p += tag.size();
        }
        if (p == std::string::npos) return FileNode();
        const int r = (int)field(p, "rows:"), c = (int)field(p, "cols:");
        const size_t d = text_.find("data:", p);
        const size_t lb = text_.find('[', d), rb = text_.find(']', lb);
        if (d == std::string::npos || lb == std::string::npos || rb == std::string::npos) return FileNode();
        std::string body = text_.substr(lb + 1, rb - lb - 1);
        for (char& ch : body) if (ch == ',') ch = ' ';
        std::stringstream ss(body);
        std::vector<double> v;
        std::string tok;
        while (ss >> tok) v.push_back(std::strtod(tok.c_str(), nullptr));
        return FileNode(r, c, v);
    }
private:
    double field(size_t from, const char* name) const
    {
        const size_t p = text_.find(name, from);
        return p == std::string::npos ? 0.0 : std::strtod(text_.c_str() + p + std::strlen(name), nullptr);
    }
    std::string text_;
};

// ---- cv::fastAtan2 == cvFastArctan (CDecodePhase.cpp:67): degrees, f32, unfused --------------
inline float fastAtan2(float y, float x)
{
    const float p1 = 0.9997878412794807f * (float)(180 / CV_PI), p3 = -0.3258083974640975f * (float)(180 / CV_PI),
                p5 = 0.1555786518463281f * (float)(180 / CV_PI), p7 = -0.04432655554792128f * (float)(180 / CV_PI);
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---- cv::blur on CV_32F, default border BORDER_REFLECT_101: sums in double, * (1./area) ------
inline int shim_reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}
inline void blur(const Mat& src, Mat& dst, Size k)
{
    if (src.type() != CV_32FC1) { std::fprintf(stderr, "ref_shim: blur supports CV_32FC1 only\n"); std::abort(); }
    Mat in = src.clone();                       // dst may alias src
    dst.create(in.rows, in.cols, CV_32FC1);
    const int ax = k.width / 2, ay = k.height / 2;
    const double scale = 1. / (k.width * k.height);
    for (int i = 0; i < in.rows; i++)
        for (int j = 0; j < in.cols; j++) {
            double s = 0;
            for (int di = -ay; di < k.height - ay; di++) {
                double rs = 0;
                for (int dj = -ax; dj < k.width - ax; dj++)
                    rs += (double)in.at<float>(shim_reflect101(i + di, in.rows), shim_reflect101(j + dj, in.cols));
                s += rs;
            }
            dst.at<float>(i, j) = (float)(s * scale);
        }
}

// ---- imread(path, CV_LOAD_IMAGE_GRAYSCALE) for .bmp (CSensorV.cpp:111-114) -------------------
// Uncompressed 8 / 24 / 32 bpp; gray = (B*1868 + G*9617 + R*4899 + 8192) >> 14 (highgui utils.cpp).
// '\\' in the path is treated as '/', and decoded files are cached per path: the reference's
// StripRegression re-reads the whole dyna group for every frame (CCalculation.cpp:791).
Mat imread(const std::string& path, int flags = 1);

}  // namespace cv

inline float cvFastArctan(float y, float x) { return cv::fastAtan2(y, x); }

#endif  // REF_SHIM_OPENCV_HPP_
