// TEST INFRASTRUCTURE ONLY -- definitions of the reference's configuration globals.
// The reference declares them `extern const` in StaticParameters.h and defines them as compile-time
// literals in StaticParameters.cpp (1280x1024 camera, 1280 projector, 6 Gray digits, 4 steps,
// "D:\\Structured_Light_Data\\").  This file replaces that one translation unit -- and only that
// one -- so that the SAME reference loops can be run at the geometries of BASELINE.json: every value
// is taken from the environment when the binary starts, defaulting to the reference's literal.
#include <cstdlib>
#include <string>

#include "StaticParameters.h"

namespace {
int env_int(const char* name, int dflt)
{
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}
std::string env_str(const char* name, const char* dflt)
{
    const char* e = std::getenv(name);
    return e ? std::string(e) : std::string(dflt);
}
}  // namespace

const int PROJECTOR_RESLINE = env_int("DYNAFRAME_PROJECTOR_RESLINE", 1280);   // StaticParameters.cpp:4
const int PROJECTOR_RESROW = env_int("DYNAFRAME_PROJECTOR_RESROW", 800);      // :5
const int CAMERA_RESLINE = env_int("DYNAFRAME_CAMERA_RESLINE", 1280);         // :8
const int CAMERA_RESROW = env_int("DYNAFRAME_CAMERA_RESROW", 1024);           // :9
const int PC_BIASLINE = 1366;
const int PC_BIASROW = 0;
const int GRAY_V_NUMDIGIT = env_int("DYNAFRAME_GRAY_V_NUMDIGIT", 6);          // :16
const int GRAY_H_NUMDIGIT = 5;
const int PHASE_NUMDIGIT = env_int("DYNAFRAME_PHASE_NUMDIGIT", 4);            // :18
const int SHOW_PICTURE_TIME = 500;
const bool VISUAL_DEBUG = false;                                              // :22
const string DATA_PATH = env_str("DYNAFRAME_DATA_PATH", "D:\\Structured_Light_Data\\");   // :30
const int DYNAFRAME_MAXNUM = env_int("DYNAFRAME_MAXNUM", 100);                // :31
const int FOV_MIN_DISTANCE = env_int("DYNAFRAME_FOV_MIN_DISTANCE", 10);       // :34
const int FOV_MAX_DISTANCE = env_int("DYNAFRAME_FOV_MAX_DISTANCE", 100);      // :35
const int RECO_WINDOW_SIZE = env_int("DYNAFRAME_RECO_WINDOW_SIZE", 21);       // :38
