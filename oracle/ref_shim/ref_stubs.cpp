// TEST INFRASTRUCTURE ONLY -- the pieces of the stand-in that need a translation unit:
// imread's BMP decoder and a no-op CVisualization (the reference's own CVisualization.cpp is a
// highgui window wrapper, disabled by VISUAL_DEBUG = false, StaticParameters.cpp:22).
#include <opencv2/opencv.hpp>

#include "CVisualization.h"

namespace cv {

namespace {
uint32_t rd32(const uchar* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint32_t rd16(const uchar* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
uchar bgr2gray(uint32_t b, uint32_t g, uint32_t r) { return (uchar)((b * 1868u + g * 9617u + r * 4899u + 8192u) >> 14); }
std::map<std::string, Mat>& cache()
{
    static std::map<std::string, Mat> c;
    return c;
}
}  // namespace

Mat imread(const std::string& path_in, int)
{
    std::string path = path_in;
    for (char& ch : path) if (ch == '\\') ch = '/';
    auto it = cache().find(path);
    if (it != cache().end()) return it->second;          // shared header; callers copyTo() before writing
    Mat out;
    std::ifstream f(path.c_str(), std::ios::binary | std::ios::ate);
    if (!f) return out;
    const std::streamsize n = f.tellg();
    if (n < 54) return out;
    std::vector<uchar> d((size_t)n);
    f.seekg(0);
    f.read(reinterpret_cast<char*>(d.data()), n);
    if (d[0] != 'B' || d[1] != 'M') return out;
    const uint32_t off = rd32(&d[10]), hdr = rd32(&d[14]);
    const int w = (int)rd32(&d[18]), hs = (int)rd32(&d[22]);
    const uint32_t bpp = rd16(&d[28]), comp = rd32(&d[30]);
    uint32_t used = rd32(&d[46]);
    if (hdr < 40 || w <= 0 || hs == 0 || comp != 0 || (bpp != 8 && bpp != 24 && bpp != 32)) return out;
    const int h = hs < 0 ? -hs : hs;
    const size_t stride = (((size_t)w * (bpp / 8)) + 3) & ~(size_t)3;
    if ((size_t)off + stride * (size_t)h > (size_t)n) return out;
    uchar lut[256] = {0};
    if (bpp == 8) {
        if (used == 0 || used > 256) used = 256;
        const uchar* pal = &d[14 + hdr];
        for (uint32_t i = 0; i < used; i++) lut[i] = bgr2gray(pal[4 * i], pal[4 * i + 1], pal[4 * i + 2]);
    }
    out.create(h, w, CV_8UC1);
    for (int y = 0; y < h; y++) {
        const uchar* src = &d[off + stride * (size_t)(hs < 0 ? y : h - 1 - y)];
        uchar* dst = out.ptr(y);
        if (bpp == 8) for (int x = 0; x < w; x++) dst[x] = lut[src[x]];
        else for (int x = 0; x < w; x++) dst[x] = bgr2gray(src[x * (bpp / 8)], src[x * (bpp / 8) + 1], src[x * (bpp / 8) + 2]);
    }
    cache()[path] = out;
    return out;
}

}  // namespace cv

CVisualization::CVisualization(string winName) : m_winName(winName) {}
CVisualization::~CVisualization() {}
int CVisualization::Show(Mat, int, bool, double, bool, string) { return 0; }
