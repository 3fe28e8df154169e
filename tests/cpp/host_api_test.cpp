// host_api_test.cpp -- reference-style host code on the re-hosted classes.
//
// Mirrors what main.cpp:42-45 and CCalculation::FillFirstProjectorU
// (CCalculation.cpp:536-559) do: feed Gray and phase images to the decoders,
// Decode(), GetResult(); Init(), CalculateFirst(), Result().  Driven by
// tests/test_cpp_host_api.py, which supplies the planes and checks the outputs
// against the CPU oracle.
//
//   host_api_test <dir> <W> <H> <PW> <G> <N>
// reads  <dir>/planes.u8 (2G+N planes), <dir>/parameters.yml, <dir>/vGrayCode.txt
// writes <dir>/gray.f64 <dir>/phase.f64 <dir>/xyzw.f32 <dir>/mask.u8 <dir>/projU.f64 <dir>/cloud.txt
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "dynaframe_b200.hpp"

using namespace dynaframe;

static bool write_file(const std::string& path, const void* data, size_t bytes)
{
    std::ofstream f(path, std::ios::binary);
    f.write(static_cast<const char*>(data), (std::streamsize)bytes);
    return (bool)f;
}

#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::fprintf(stderr, "CHECK failed line %d: %s\n", __LINE__, #cond); \
            return 1;                                                      \
        }                                                                  \
    } while (0)

int main(int argc, char** argv)
{
    if (argc < 7) { std::fprintf(stderr, "usage\n"); return 2; }
    const std::string dir = argv[1];
    StaticParameters sp;
    sp.CAMERA_RESLINE = std::atoi(argv[2]);
    sp.CAMERA_RESROW = std::atoi(argv[3]);
    sp.PROJECTOR_RESLINE = std::atoi(argv[4]);
    sp.GRAY_V_NUMDIGIT = std::atoi(argv[5]);
    sp.PHASE_NUMDIGIT = std::atoi(argv[6]);
    sp.DATA_PATH = dir + "/";
    const int W = sp.CAMERA_RESLINE, H = sp.CAMERA_RESROW, G = sp.GRAY_V_NUMDIGIT, N = sp.PHASE_NUMDIGIT;
    const size_t npx = (size_t)W * H;

    std::vector<uint8_t> planes((size_t)(2 * G + N) * npx);
    {
        std::ifstream f(dir + "/planes.u8", std::ios::binary);
        f.read(reinterpret_cast<char*>(planes.data()), (std::streamsize)planes.size());
        CHECK((size_t)f.gcount() == planes.size());
    }
    auto plane = [&](int i) { return Mat(H, W, CV_8UC1, planes.data() + (size_t)i * npx); };

    // ---- error behaviour of the decoder objects (CDecodeGray.cpp:26-40, CDecodePhase.cpp:109-123)
    {
        CDecodeGray g(sp);
        CHECK(!g.SetMat(0, plane(0)));                 // before SetNumDigit
        CHECK(!g.SetNumDigit(0, true));
        CHECK(!g.SetNumDigit(17, true));
        CDecodePhase ph(sp);
        CHECK(!ph.SetMat(0, plane(0)));                // before SetNumMat
        CHECK(!ph.SetNumMat(0, 16));
    }

    // ---- CDecodeGray: CCalculation.cpp:537-546
    CDecodeGray gray(sp);
    CHECK(gray.SetNumDigit(G, true));
    CHECK(gray.SetMatFileName(dir + "/", "missing.txt"));
    for (int i = 0; i < 2 * G; i++) CHECK(gray.SetMat(i, plane(i)));
    CHECK(!gray.Decode());                             // "Gray Decode->Open file error."
    CHECK(LastErrorMessage() == "Gray Decode->Open file error.");
    CHECK(gray.SetMatFileName(dir + "/", "vGrayCode.txt"));
    CHECK(gray.Decode());
    Mat vGrayMat = gray.GetResult();
    CHECK(vGrayMat.type() == CV_64FC1 && vGrayMat.rows == H && vGrayMat.cols == W);
    CHECK(write_file(dir + "/gray.f64", vGrayMat.ptr(), npx * 8));

    // ---- CDecodePhase: CCalculation.cpp:550-559
    const int v_pixPeriod = sp.PROJECTOR_RESLINE / (1 << (G - 1));
    CDecodePhase phase(sp);
    CHECK(phase.SetNumMat(N, v_pixPeriod));
    for (int i = 0; i < N; i++) CHECK(phase.SetMat(i, plane(2 * G + i)));
    CHECK(phase.Decode());
    Mat vPhaseMat = phase.GetResult();
    CHECK(write_file(dir + "/phase.f64", vPhaseMat.ptr(), npx * 8));
    // GetResult hands out deep copies (CDecodePhase.cpp:99-104)
    vPhaseMat.at<double>(0, 0) = -1.0;
    CHECK(phase.GetResult().at<double>(0, 0) != -1.0);

    // ---- CCalculation: main.cpp:42-44
    CCalculation calc(sp);
    CHECK(!calc.CalculateFirst());                     // before Init (CCalculation.cpp:176-181)
    calc.SetParameterFile("parameters.yml");
    calc.SetGrayCodeFile(dir + "/", "vGrayCode.txt");
    CHECK(calc.Init());
    CHECK(!calc.Init());                               // twice (CCalculation.cpp:80-83)
    for (int i = 0; i < 2 * G; i++) CHECK(calc.Sensor()->StoreDatas(0, i, plane(i)));
    for (int i = 0; i < N; i++) CHECK(calc.Sensor()->StoreDatas(1, i, plane(2 * G + i)));
    CHECK(calc.CalculateFirst());
    CHECK(!calc.CalculateOther());                     // no dynamic frames stored yet
    {   // dynamic frames (optional input <dir>/dyna.u8 = n images)
        std::ifstream f(dir + "/dyna.u8", std::ios::binary | std::ios::ate);
        if (f) {
            const size_t bytes = (size_t)f.tellg();
            const int n = (int)(bytes / npx);
            std::vector<uint8_t> dyn(bytes);
            f.seekg(0);
            f.read(reinterpret_cast<char*>(dyn.data()), (std::streamsize)bytes);
            for (int i = 0; i < n; i++) CHECK(calc.Sensor()->StoreDatas(2, i, Mat(H, W, CV_8UC1, dyn.data() + (size_t)i * npx)));
            CHECK(calc.CalculateOther());
            CHECK(calc.FrameCount() == n);
            std::ofstream o(dir + "/dyna_xyzw.f32", std::ios::binary), m(dir + "/dyna_mask.u8", std::ios::binary),
                dz(dir + "/dyna_dz.f32", std::ios::binary);
            for (int i = 1; i < n; i++) {
                o.write(reinterpret_cast<const char*>(calc.PointMap(i).ptr()), (std::streamsize)(npx * 16));
                m.write(reinterpret_cast<const char*>(calc.ValidMask(i).ptr()), (std::streamsize)npx);
                dz.write(reinterpret_cast<const char*>(calc.DeltaZ(i).ptr()), (std::streamsize)(npx * 4));
            }
            CHECK(calc.Result(dir + "/cloud_dyn1.txt", 1));
        }
    }
    CHECK(calc.PointMap().type() == CV_32FC4 && calc.ValidMask().type() == CV_8UC1);
    CHECK(write_file(dir + "/xyzw.f32", calc.PointMap().ptr(), npx * 16));
    CHECK(write_file(dir + "/mask.u8", calc.ValidMask().ptr(), npx));
    Mat U = calc.GetProjectorU();
    CHECK(write_file(dir + "/projU.f64", U.ptr(), npx * 8));
    CHECK(calc.Result(dir + "/cloud.txt", 0));
    CHECK(calc.ResultPly(dir + "/cloud.ply", 0));
    CHECK(!calc.Result(dir + "/no/such/dir/cloud.txt", 0));   // open failure (CCalculation.cpp:327-331)
    Mat z = calc.GetZ();
    CHECK(z.type() == CV_64FC1 && z.at<double>(H / 2, W / 2) == (double)calc.PointMap().at<float>(H / 2, 4 * (W / 2) + 2));
    {   // file-backed sensor (CSensorV.cpp:31-133): <dir>/group/iFrame/vGrayCam{i}.bmp ... through a second CCalculation
        std::ifstream probe(dir + "/group/iFrame/vGrayCam0.bmp", std::ios::binary);
        if (probe) {
            CCalculation fcalc(sp);
            fcalc.SetParameterFile("parameters.yml");
            fcalc.SetGrayCodeFile(dir + "/", "vGrayCode.txt");
            fcalc.SetPointCloudFile("");
            fcalc.SetGroupDataPath(dir + "/group");
            CHECK(fcalc.Init());
            CHECK(fcalc.Sensor()->FileName(0, 3) == dir + "/group/iFrame/vGrayCam3.bmp");
            CHECK(fcalc.Sensor()->FileName(1, 0) == dir + "/group/iFrame/vPhaseCam0.bmp");
            CHECK(fcalc.Sensor()->FileName(2, 7) == dir + "/group/cFrame/dynaCam7.bmp");
            CHECK(fcalc.CalculateFirst());
            CHECK(std::memcmp(fcalc.PointMap().ptr(), calc.PointMap().ptr(), npx * 16) == 0);
            CHECK(std::memcmp(fcalc.ValidMask().ptr(), calc.ValidMask().ptr(), npx) == 0);
            CHECK(fcalc.CalculateOther());
            CHECK(fcalc.FrameCount() == calc.FrameCount());
            for (int i = 1; i < fcalc.FrameCount(); i++)
                CHECK(std::memcmp(fcalc.PointMap(i).ptr(), calc.PointMap(i).ptr(), npx * 16) == 0);
            std::printf("file-backed sensor ok\n");
        }
    }
    {   // several GPUs behind the same protocol (here: two members on device 0): n frame sets in one call,
        // contiguous shards, results identical to CCalculation's
        CCalculationPool pool(sp);
        CHECK(!pool.CalculateFirstBatch(planes.data(), 1, nullptr, nullptr));      // before Init
        pool.SetParameterFile("parameters.yml");
        pool.SetGrayCodeFile(dir + "/", "vGrayCode.txt");
        pool.SetPipeline(1, 2);
        CHECK(pool.Init({0, 0}));
        CHECK(!pool.Init({0}));                                                    // twice
        CHECK(pool.Devices() == 2);
        int lo = -1, hi = -1;
        CHECK(pool.ShardRange(3, 0, lo, hi) && lo == 0 && hi == 2);
        CHECK(pool.ShardRange(3, 1, lo, hi) && lo == 2 && hi == 3);
        const int n = 3;
        const size_t sb = planes.size();
        uint8_t* in = static_cast<uint8_t*>(slc_host_alloc(n * sb));
        float* xyzw = static_cast<float*>(slc_host_alloc(n * npx * 16));
        uint8_t* mask = static_cast<uint8_t*>(slc_host_alloc(n * npx));
        CHECK(in && xyzw && mask);
        for (int i = 0; i < n; i++) std::memcpy(in + i * sb, planes.data(), sb);
        CHECK(pool.CalculateFirstBatch(in, n, xyzw, mask));
        for (int i = 0; i < n; i++) {
            CHECK(std::memcmp(xyzw + (size_t)i * npx * 4, calc.PointMap().ptr(), npx * 16) == 0);
            CHECK(std::memcmp(mask + (size_t)i * npx, calc.ValidMask().ptr(), npx) == 0);
        }
        // depth-only result: z plane + one validity bit per pixel
        float* depth = static_cast<float*>(slc_host_alloc(n * npx * 4));
        uint8_t* bits = static_cast<uint8_t*>(slc_host_alloc(n * ((npx + 7) / 8) + 4));
        slc_result r;
        std::memset(&r, 0, sizeof(r));
        r.format = SLC_RESULT_DEPTH;
        r.depth = depth;
        r.mask_bits = bits;
        CHECK(pool.CalculateFirstBatch(in, n, r));
        for (size_t q = 0; q < npx; q++) {
            CHECK(depth[2 * npx + q] == calc.PointMap().at<float>((int)(q / W), 4 * (int)(q % W) + 2));
            CHECK(((bits[2 * ((npx + 7) / 8) + (q >> 3)] >> (q & 7)) & 1) == calc.ValidMask().at<uint8_t>((int)(q / W), (int)(q % W)));
        }
        slc_host_free(in); slc_host_free(xyzw); slc_host_free(mask); slc_host_free(depth); slc_host_free(bits);
        std::printf("pool ok\n");
    }
    std::printf("host_api_test ok\n");
    return 0;
}
