// dynaframe_b200.hpp -- the reference's own C++ entry points, re-hosted on the
// C ABI of slcalc_b200.h.
//
// elevenface/Structured-Light-Calculation has no plugin layer; host code talks
// to three classes (paths relative to DynaFrame/DynaFrame/):
//     CDecodeGray   CDecodeGray.h:18-53
//     CDecodePhase  CDecodePhase.h:12-39
//     CCalculation  CCalculation.h:10-95
// with a setter -> Decode()/Calculate*() -> getter protocol, bool returns and
// ErrorHandling(string) for messages (GlobalFunction.h:9).  The classes below
// keep those names, argument meanings and error behaviour, so reference-style
// host code ports by changing an #include and a namespace.  What differs:
//   * cv::Mat is replaced by dynaframe::Mat, a small ref-counted dense 2-D array
//     with the members this path uses (rows, cols, type(), at<T>(), copyTo,
//     clone, create, empty) -- OpenCV is not a dependency;
//   * the compile-time globals of StaticParameters.{h,cpp} become the runtime
//     struct StaticParameters handed to the constructors;
//   * the file-backed CSensor (CSensorV.cpp) is replaced by an in-memory
//     CSensor with the same LoadDatas / SetProPicture / GetCamPicture calls;
//   * all arithmetic runs on the GPU through slcalc_b200.h.  No CPU fallback:
//     without a CUDA device every Decode()/Calculate*() returns false.
#ifndef DYNAFRAME_B200_HPP_
#define DYNAFRAME_B200_HPP_

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "slcalc_b200.h"

namespace dynaframe {

// OpenCV type codes used on the path
enum { CV_8UC1 = 0, CV_16SC1 = 3, CV_32FC1 = 5, CV_64FC1 = 6, CV_32FC4 = 29 };

class Mat {
public:
    int rows = 0, cols = 0;
    Mat() = default;
    Mat(int rows_, int cols_, int type_) { create(rows_, cols_, type_); }
    // wrap caller-owned memory (no copy, like cv::Mat(rows, cols, type, data, step))
    Mat(int rows_, int cols_, int type_, void* data, size_t step_bytes = 0);
    void create(int rows_, int cols_, int type_);
    bool empty() const { return data_ == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    size_t elemSize() const;
    size_t step() const { return step_; }
    bool isContinuous() const { return step_ == (size_t)cols * elemSize(); }
    uint8_t* ptr(int r = 0) { return data_ + (size_t)r * step_; }
    const uint8_t* ptr(int r = 0) const { return data_ + (size_t)r * step_; }
    template <typename T> T& at(int r, int c) { return reinterpret_cast<T*>(data_ + (size_t)r * step_)[c]; }
    template <typename T> const T& at(int r, int c) const {
        return reinterpret_cast<const T*>(data_ + (size_t)r * step_)[c];
    }
    void copyTo(Mat& dst) const;   // deep copy
    Mat clone() const { Mat m; copyTo(m); return m; }

private:
    int type_ = CV_8UC1;
    size_t step_ = 0;
    uint8_t* data_ = nullptr;
    std::shared_ptr<uint8_t> owner_;
};

// StaticParameters.cpp:4-38 as a runtime value; defaults are the reference's.
struct StaticParameters {
    int PROJECTOR_RESLINE = 1280;
    int PROJECTOR_RESROW = 800;
    int CAMERA_RESLINE = 1280;
    int CAMERA_RESROW = 1024;
    int GRAY_V_NUMDIGIT = 6;
    int GRAY_H_NUMDIGIT = 5;
    int PHASE_NUMDIGIT = 4;
    bool VISUAL_DEBUG = false;
    std::string DATA_PATH = "";
    int DYNAFRAME_MAXNUM = 100;
    int FOV_MIN_DISTANCE = 10;
    int FOV_MAX_DISTANCE = 100;
    int RECO_WINDOW_SIZE = 21;
    // additions of this implementation
    float MODULATION_MIN = 0.0f;   // [EXT] 0 = disabled (reference behaviour)
    int CUDA_DEVICE = 0;
};

// GlobalFunction.cpp:3-8 without the system("PAUSE"): prints the message,
// remembers it, returns 0.
int ErrorHandling(std::string message);
const std::string& LastErrorMessage();

// The sensor (CSensorV.h:15-61): group 0 = vGray images, 1 = vPhase images, 2 = dyna images.
// Either fed from memory (StoreDatas) or file-backed like the reference: after AttachFiles,
// LoadDatas(group) reads <groupDataPath>/iFrame/vGrayCam{i}.bmp, .../iFrame/vPhaseCam{i}.bmp or
// .../cFrame/dynaCam{i}.bmp (CSensorV.cpp:35-41,74-121); the BMP pixel work (row flip, padding,
// palette / BGR -> gray as imread(..., CV_LOAD_IMAGE_GRAYSCALE)) runs on the device.
class CSensor {
public:
    bool InitSensor();
    bool CloseSensor();
    bool StoreDatas(int groupNum, int idx, const Mat& picture);   // feeds what LoadDatas would read from disk
    void AttachFiles(slc_context* ctx, const StaticParameters& sp, const std::string& groupDataPath);
    std::string FileName(int groupNum, int idx) const;
    bool LoadDatas(int groupNum);
    bool UnloadDatas();
    bool SetProPicture(int nowNum);
    Mat GetCamPicture();        // deep copy, like the reference

private:
    std::vector<Mat> groups_[3];
    int group_ = -1;
    int now_ = 0;
    // file-backed mode
    slc_context* ctx_ = nullptr;
    int counts_[3] = {0, 0, 0};
    int rows_ = 0, cols_ = 0;
    std::string m_groupDataPath, m_iFramePath = "iFrame/", m_cFramePath = "cFrame/";
    std::string m_vGrayName = "vGrayCam", m_vPhaseName = "vPhaseCam", m_dynaName = "dynaCam", m_dataFileSuffix = ".bmp";
};

// CDecodeGray.h:18-53
class CDecodeGray {
public:
    explicit CDecodeGray(const StaticParameters& sp = StaticParameters());
    ~CDecodeGray();
    CDecodeGray(const CDecodeGray&) = delete;
    CDecodeGray& operator=(const CDecodeGray&) = delete;

    bool Decode();
    Mat GetResult();                                         // CV_64FC1, deep copy
    bool SetMat(int num, Mat pic);                           // deep-copies the input
    bool SetNumDigit(int numDigit, bool ver);
    bool SetMatFileName(std::string codeFilePath, std::string codeFileName);

private:
    bool AllocateSpace();
    bool ReleaseSpace();
    StaticParameters sp_;
    int m_numDigit = 0;
    int m_grayCodeSize = 0;
    std::vector<int16_t> m_gray2bin;
    std::string m_codeFilePath, m_codeFileName;
    bool m_vertical = true;
    bool allocated_ = false;
    uint8_t* pinned_ = nullptr;          // [2G][H][W] staging, pinned
    std::vector<uint8_t> have_;
    Mat m_result;
    slc_context* ctx_ = nullptr;
};

// CDecodePhase.h:12-39
class CDecodePhase {
public:
    explicit CDecodePhase(const StaticParameters& sp = StaticParameters());
    ~CDecodePhase();
    CDecodePhase(const CDecodePhase&) = delete;
    CDecodePhase& operator=(const CDecodePhase&) = delete;

    bool Decode();
    Mat GetResult();                                         // CV_64FC1, deep copy
    bool SetMat(int num, Mat pic);
    bool SetNumMat(int numMat, int pixperiod);

private:
    bool DeleteSpace();
    StaticParameters sp_;
    int m_numMat = 0;
    int m_pixPeroid = 16;
    bool allocated_ = false;
    uint8_t* pinned_ = nullptr;
    Mat m_result;
    slc_context* ctx_ = nullptr;
};

// CCalculation.h:10-95 (first-frame path).
class CCalculation {
public:
    explicit CCalculation(const StaticParameters& sp = StaticParameters());
    ~CCalculation();
    CCalculation(const CCalculation&) = delete;
    CCalculation& operator=(const CCalculation&) = delete;

    // Where Init() finds what the reference hard-codes (CCalculation.cpp:86-93,124-127):
    void SetParameterFile(const std::string& ymlPath) { m_paraFile = ymlPath; }
    void SetGrayCodeFile(const std::string& path, const std::string& name) { m_codePath = path; m_codeName = name; }
    void SetPointCloudFile(const std::string& file) { m_pcFile = file; }
    // file-backed sensor: the directory that holds iFrame/ and cFrame/ (CSensorV.cpp:35); empty = in-memory sensor
    void SetGroupDataPath(const std::string& path) { m_groupDataPath = path; }
    CSensor* Sensor() { return m_sensor; }                  // valid after Init()

    bool Init();
    bool CalculateFirst();
    bool CalculateOther();      // dynamic frames 1 .. DYNAFRAME_MAXNUM-1 (CCalculation.cpp:208-320)
    bool Result(std::string fileName, int i);       // the reference's text cloud, formatted on the device
    bool ResultPly(std::string fileName, int i);    // binary little-endian PLY of the same points
    // line ends / exponent digits of Result(): crlf + exp3 reproduce a Windows MSVC 2013 run byte for byte
    void SetResultTextStyle(bool crlf, bool exp3) { m_textFlags = (crlf ? SLC_TEXT_CRLF : 0u) | (exp3 ? SLC_TEXT_EXP3 : 0u); }
    void SetDynamicPointCloudPrefix(const std::string& prefix) { m_pcDynaPrefix = prefix; }   // "cFrame" + idx + ".txt"

    // results of frame 0 (valid after CalculateFirst): CV_32FC4 (x,y,z,U) and CV_8UC1 mask,
    // plus the reference's separate planes on request
    const Mat& PointMap() const { return m_xyzw; }
    const Mat& ValidMask() const { return m_mask; }
    // dynamic frames (valid after CalculateOther); i in 1 .. FrameCount()-1
    int FrameCount() const { return 1 + (int)m_dynXyzw.size(); }
    const Mat& PointMap(int i) const { return i == 0 ? m_xyzw : m_dynXyzw.at((size_t)i - 1); }
    const Mat& ValidMask(int i) const { return i == 0 ? m_mask : m_dynMask.at((size_t)i - 1); }
    const Mat& DeltaZ(int i) const { return m_dynDeltaZ.at((size_t)i - 1); }                     // CV_32FC1, m_deltaZ[i]
    Mat GetX() const;   // CV_64FC1 views of m_xMat[0] / m_yMat[0] / m_zMat[0] / m_ProjectorU[0]
    Mat GetY() const;
    Mat GetZ() const;
    Mat GetProjectorU() const;

private:
    bool ReleaseSpace();
    bool FillFirstProjectorUAndCoordinate();
    StaticParameters sp_;
    CSensor* m_sensor = nullptr;
    slc_context* ctx_ = nullptr;
    std::string m_paraFile = "parameters.yml";
    std::string m_codePath = "Patterns/", m_codeName = "vGrayCode.txt";
    std::string m_pcFile = "iFrame.txt";
    std::string m_groupDataPath;
    uint8_t* pinned_stack_ = nullptr;
    Mat m_xyzw, m_mask, m_projU;
    std::vector<Mat> m_dynXyzw, m_dynMask, m_dynDeltaZ, m_dynProjU;   // headers onto m_dynBlock
    uint8_t* m_dynBlock = nullptr;      // pinned: every map of the dynamic sequence
    char* m_textBuf = nullptr;          // pinned: Result()'s text of one frame
    uint32_t m_textFlags = 0;
    std::string m_pcDynaPrefix;
    bool calibrated_ = false;
};

// Several GPUs behind the reference's call protocol.  The reference's main() runs one
// CCalculation over one sequence (main.cpp:42-45); its frame loop (CCalculation.cpp:221) is what a
// multi-GPU user would shard.  CCalculationPool keeps Init() -> Calculate...() -> bool + ErrorHandling:
// Init() reads the same parameters.yml / vGrayCode.txt as CCalculation::Init and replicates them to one
// context per listed device; CalculateFirstBatch() runs n independent frame sets (each the 2G+N images
// CalculateFirst would pull from the sensor, plane-major, contiguous) through slc_pool_reconstruct_host:
// contiguous shards, one feeder thread + pinned multi-slot pipeline per GPU, no collective.
class CCalculationPool {
public:
    explicit CCalculationPool(const StaticParameters& sp = StaticParameters());
    ~CCalculationPool();
    CCalculationPool(const CCalculationPool&) = delete;
    CCalculationPool& operator=(const CCalculationPool&) = delete;

    void SetParameterFile(const std::string& ymlPath) { m_paraFile = ymlPath; }
    void SetGrayCodeFile(const std::string& path, const std::string& name) { m_codePath = path; m_codeName = name; }
    // frame sets per upload / launch / download chunk and stream slots per GPU (before Init)
    void SetPipeline(int framesPerChunk, int slots) { m_chunk = framesPerChunk; m_slots = slots; }

    bool Init(const std::vector<int>& devices);     // twice => false, like CCalculation::Init (CCalculation.cpp:80-83)
    int Devices() const;
    // [lo, hi) of the n frame sets GPU `member` computes (slc_shard_range)
    bool ShardRange(int n, int member, int& lo, int& hi) const;
    // stacks: u8 [n][2G+N][H][W]; xyzw: f32 [n][H][W][4]; mask: u8 [n][H][W].  Host memory, ideally pinned.
    bool CalculateFirstBatch(const uint8_t* stacks, int n, float* xyzw, uint8_t* mask);
    // the same with a result format (SLC_RESULT_DEPTH / SLC_RESULT_POINTS: fewer bytes back over PCIe)
    bool CalculateFirstBatch(const uint8_t* stacks, int n, const slc_result& out);
    slc_pool* Handle() { return pool_; }

private:
    StaticParameters sp_;
    slc_pool* pool_ = nullptr;
    std::string m_paraFile = "parameters.yml";
    std::string m_codePath = "Patterns/", m_codeName = "vGrayCode.txt";
    int m_chunk = 2, m_slots = 4;
};

// Helpers shared by the classes and usable on their own.
bool ReadCalibrationYaml(const std::string& path, double cam[9], double pro[9], double R[9], double T[3]);
bool ReadGrayCodeFile(const std::string& file, int grayCodeSize, std::vector<int16_t>& gray2bin);

}  // namespace dynaframe

#endif  // DYNAFRAME_B200_HPP_
