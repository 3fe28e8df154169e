// TEST INFRASTRUCTURE ONLY -- drives the reference's own classes (compiled from /root/reference,
// see oracle/Makefile) exactly as main.cpp:42-45 does, and dumps the planes they compute so that
// tests can compare the CPU restatement (oracle/sl_oracle.c) and the CUDA path against the
// reference's real loops.  Configuration comes from DYNAFRAME_* environment variables
// (oracle/ref_shim/ref_params.cpp).
//
//   dynaframe_ref first <outdir>     Init, FillFirstProjectorU, FillCoordinate(0); dumps gray / phase /
//                                    ProjectorU / x / y / z (f64) and A, B, P, cC, cD
//   dynaframe_ref full <outdir>      Init, CalculateFirst, CalculateOther (the reference writes its own
//                                    text clouds under DATA_PATH); dumps per-frame stripB/W, deltaP,
//                                    ProjectorU, x, y, z, deltaZ
//   dynaframe_ref app <outdir>       Init, CalculateFirst, CalculateOther and nothing else: the reference program
//                                    as main.cpp:42-45 runs it (its text clouds land under DATA_PATH)
//   dynaframe_ref time <reps>        Init once, then reps x (FillFirstProjectorU + FillCoordinate(0)),
//                                    prints seconds per repetition (hot loops; images come from the
//                                    stand-in imread's in-memory cache after the first repetition)
#include <opencv2/opencv.hpp>

#include <chrono>
#include <iostream>
#include <strstream>
#include <unistd.h>

// the planes are private members of CCalculation; the class layout does not depend on access
#define private public
#include "CCalculation.h"
#undef private

static bool dump(const std::string& path, const cv::Mat& m)
{
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const size_t row = (size_t)m.cols * m.elemSize();
    for (int i = 0; i < m.rows; i++) std::fwrite(m.ptr(i), 1, row, f);
    std::fclose(f);
    return true;
}

static bool dump_scalars(const std::string& path, const std::vector<double>& v)
{
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fwrite(v.data(), sizeof(double), v.size(), f);
    std::fclose(f);
    return true;
}

int main(int argc, char** argv)
{
    if (argc < 3) {
        std::fprintf(stderr, "usage: dynaframe_ref first|full <outdir> | time <reps>\n");
        return 2;
    }
    const std::string mode = argv[1];
    if (const char* cwd = std::getenv("DYNAFRAME_CWD")) {
        if (chdir(cwd) != 0) { std::perror("chdir"); return 3; }       // where Patterns/vGrayCode.txt lives (CCalculation.cpp:538)
    }
    CCalculation calc;
    if (!calc.Init()) { std::fprintf(stderr, "Init failed\n"); return 4; }
    if (calc.m_C.empty() || calc.m_P.empty()) { std::fprintf(stderr, "calibration not read\n"); return 5; }

    if (mode == "time") {
        const int reps = std::atoi(argv[2]);
        std::printf("{\"seconds\": [");
        for (int r = 0; r < reps; r++) {
            const auto t0 = std::chrono::steady_clock::now();
            calc.FillFirstProjectorU();
            calc.FillCoordinate(0);
            const auto t1 = std::chrono::steady_clock::now();
            std::printf("%s%.6f", r ? ", " : "", std::chrono::duration<double>(t1 - t0).count());
        }
        std::printf("], \"z_centre\": %.17g}\n", calc.m_zMat[0].at<double>(CAMERA_RESROW / 2, CAMERA_RESLINE / 2));
        return 0;
    }

    if (mode == "app") {
        if (!calc.CalculateFirst() || !calc.CalculateOther()) { std::fprintf(stderr, "Calculate* failed\n"); return 6; }
        std::printf("dynaframe_ref ok\n");
        return 0;
    }

    const std::string out = std::string(argv[2]) + "/";
    std::vector<double> sc = {calc.m_cA, calc.m_cB};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) sc.push_back(calc.m_P.at<double>(i, j));
    bool ok = dump_scalars(out + "A_B_P.f64", sc) && dump(out + "cC.f64", calc.m_cC) && dump(out + "cD.f64", calc.m_cD);

    if (mode == "first") {
        calc.FillFirstProjectorU();
        calc.FillCoordinate(0);
        ok = ok && dump(out + "gray.f64", calc.m_decodeGrayv->GetResult()) &&
             dump(out + "phase.f64", calc.m_decodePhasev->GetResult());
    } else if (mode == "full") {
        if (!calc.CalculateFirst() || !calc.CalculateOther()) { std::fprintf(stderr, "Calculate* failed\n"); return 6; }
        for (int f = 0; f < DYNAFRAME_MAXNUM; f++) {
            std::ostringstream n;
            n << f;
            ok = ok && dump(out + "stripB" + n.str() + ".f32", calc.m_stripB[f]) &&
                 dump(out + "stripW" + n.str() + ".f32", calc.m_stripW[f]) &&
                 dump(out + "projU" + n.str() + ".f64", calc.m_ProjectorU[f]) &&
                 dump(out + "x" + n.str() + ".f64", calc.m_xMat[f]) && dump(out + "y" + n.str() + ".f64", calc.m_yMat[f]) &&
                 dump(out + "z" + n.str() + ".f64", calc.m_zMat[f]);
            if (f > 0)
                ok = ok && dump(out + "deltaP" + n.str() + ".f32", calc.m_deltaP[f]) &&
                     dump(out + "deltaZ" + n.str() + ".f64", calc.m_deltaZ[f]);
        }
    } else {
        std::fprintf(stderr, "unknown mode %s\n", mode.c_str());
        return 2;
    }
    ok = ok && dump(out + "projU.f64", calc.m_ProjectorU[0]) && dump(out + "x.f64", calc.m_xMat[0]) &&
         dump(out + "y.f64", calc.m_yMat[0]) && dump(out + "z.f64", calc.m_zMat[0]);
    if (!ok) { std::fprintf(stderr, "dump failed\n"); return 7; }
    std::printf("dynaframe_ref ok\n");
    return 0;
}
