// dynaframe_main.cpp -- the reference's main() (main.cpp:42-45: Init, CalculateFirst, CalculateOther)
// on the B200 library: the same input files in (parameters.yml, iFrame/vGrayCam{i}.bmp,
// iFrame/vPhaseCam{i}.bmp, cFrame/dynaCam{i}.bmp, Patterns/vGrayCode.txt), the same text clouds
// out (iFrame.txt, cFrame{f}.txt).  Only include/dynaframe_b200.hpp is used.
//
//   dynaframe_main <data_path> <cam_w> <cam_h> <projector_w> <gray_digits> <phase_steps> <n_dyna_frames> <out_dir>
//
// <data_path> holds parameters.yml and 20161103/MoveBoard1103/{iFrame,cFrame}/ like the
// reference's DATA_PATH (StaticParameters.cpp:30, CSensorV.cpp:35-41); Patterns/vGrayCode.txt is
// looked up relative to the working directory (CCalculation.cpp:538).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "dynaframe_b200.hpp"

using namespace dynaframe;

int main(int argc, char** argv)
{
    if (argc != 9) {
        std::fprintf(stderr, "usage: %s data_path cam_w cam_h projector_w gray_digits phase_steps n_dyna_frames out_dir\n", argv[0]);
        return 2;
    }
    StaticParameters sp;
    sp.DATA_PATH = std::string(argv[1]) + "/";
    sp.CAMERA_RESLINE = std::atoi(argv[2]);
    sp.CAMERA_RESROW = std::atoi(argv[3]);
    sp.PROJECTOR_RESLINE = std::atoi(argv[4]);
    sp.GRAY_V_NUMDIGIT = std::atoi(argv[5]);
    sp.PHASE_NUMDIGIT = std::atoi(argv[6]);
    sp.DYNAFRAME_MAXNUM = std::atoi(argv[7]);
    const std::string out = std::string(argv[8]) + "/";

    const auto t0 = std::chrono::steady_clock::now();
    CCalculation myCalculation(sp);
    myCalculation.SetParameterFile("parameters.yml");                          // CCalculation.cpp:86-88
    myCalculation.SetGrayCodeFile("Patterns/", "vGrayCode.txt");               // :538
    myCalculation.SetGroupDataPath(sp.DATA_PATH + "20161103/MoveBoard1103");   // CSensorV.cpp:35
    myCalculation.SetPointCloudFile("");                                       // clouds are written below, outside DATA_PATH
    if (!myCalculation.Init()) return 3;                                       // main.cpp:43
    const auto t1 = std::chrono::steady_clock::now();
    if (!myCalculation.CalculateFirst()) return 4;                             // main.cpp:44
    if (!myCalculation.Result(out + "iFrame.txt", 0)) return 5;                // CCalculation.cpp:192-197
    const auto t2 = std::chrono::steady_clock::now();
    if (sp.DYNAFRAME_MAXNUM > 1) {
        if (!myCalculation.CalculateOther()) return 6;                         // main.cpp:45
        for (int f = 1; f < myCalculation.FrameCount(); f++)                   // CCalculation.cpp:310-315
            if (!myCalculation.Result(out + "cFrame" + std::to_string(f) + ".txt", f)) return 7;
    }
    const auto t3 = std::chrono::steady_clock::now();
    auto s = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double>(b - a).count();
    };
    std::printf("{\"init_s\": %.6f, \"first_s\": %.6f, \"other_s\": %.6f, \"total_s\": %.6f, \"frames\": %d}\n", s(t0, t1),
                s(t1, t2), s(t2, t3), s(t0, t3), myCalculation.FrameCount());
    return 0;
}
