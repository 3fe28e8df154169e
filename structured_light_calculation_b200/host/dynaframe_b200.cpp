// dynaframe_b200.cpp -- host-side mirror of the reference's class API
// (CDecodeGray.h:18-53, CDecodePhase.h:12-39, CCalculation.h:10-95) on top of
// the C ABI.  Plain C++17, no CUDA headers, no OpenCV.
#include "dynaframe_b200.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

namespace dynaframe {

// ---------------------------------------------------------------- Mat -----
static size_t elem_size_of(int type)
{
    switch (type) {
    case CV_8UC1: return 1;
    case CV_16SC1: return 2;
    case CV_32FC1: return 4;
    case CV_64FC1: return 8;
    case CV_32FC4: return 16;
    default: return 1;
    }
}

size_t Mat::elemSize() const { return elem_size_of(type_); }

Mat::Mat(int rows_, int cols_, int type__, void* data, size_t step_bytes)
{
    rows = rows_; cols = cols_; type_ = type__;
    step_ = step_bytes ? step_bytes : (size_t)cols_ * elem_size_of(type__);
    data_ = static_cast<uint8_t*>(data);
}

void Mat::create(int rows_, int cols_, int type__)
{
    if (rows == rows_ && cols == cols_ && type_ == type__ && owner_ && isContinuous()) return;
    rows = rows_; cols = cols_; type_ = type__;
    step_ = (size_t)cols_ * elem_size_of(type__);
    const size_t bytes = step_ * (size_t)rows_;
    void* p = nullptr;
    if (bytes && posix_memalign(&p, 64, bytes) != 0) p = nullptr;
    owner_.reset(static_cast<uint8_t*>(p), [](uint8_t* q) { free(q); });
    data_ = owner_.get();
    if (bytes && !data_) { rows = cols = 0; step_ = 0; }     // allocation failed: the Mat is empty(), not a null map of size rows x cols
}

void Mat::copyTo(Mat& dst) const
{
    if (empty()) { dst = Mat(); return; }
    dst.create(rows, cols, type_);
    const size_t row_bytes = (size_t)cols * elemSize();
    for (int r = 0; r < rows; r++) std::memcpy(dst.ptr(r), ptr(r), row_bytes);
}

// ------------------------------------------------------- ErrorHandling ----
static thread_local std::string g_last_message;

int ErrorHandling(std::string message)
{
    // GlobalFunction.cpp:5 prints the same line; the system("PAUSE") at :6 is dropped
    std::cout << "An Error Occurs:" << message << std::endl;
    g_last_message = message;
    return 0;
}

const std::string& LastErrorMessage() { return g_last_message; }

// ------------------------------------------------------------ helpers -----
bool ReadGrayCodeFile(const std::string& file, int grayCodeSize, std::vector<int16_t>& gray2bin)
{
    // CDecodeGray.cpp:113-125
    std::ifstream codeFile(file.c_str(), std::ios::in);
    if (!codeFile) return false;
    gray2bin.assign((size_t)grayCodeSize, 0);
    for (int i = 0; i < grayCodeSize; i++) {
        int binCode = 0, grayCode = 0;
        if (!(codeFile >> binCode >> grayCode)) return false;
        if (grayCode < 0 || grayCode >= grayCodeSize) return false;
        gray2bin[(size_t)grayCode] = (int16_t)binCode;
    }
    return true;
}

static bool parse_matrix(const std::string& text, const std::string& key, int want, double* out)
{
    // the subset of YAML 1.0 that cv::FileStorage writes for !!opencv-matrix
    const size_t k = text.find(key + ":");
    if (k == std::string::npos) return false;
    const size_t d = text.find("data:", k);
    if (d == std::string::npos) return false;
    const size_t lb = text.find('[', d), rb = text.find(']', d);
    if (lb == std::string::npos || rb == std::string::npos || rb < lb) return false;
    std::string body = text.substr(lb + 1, rb - lb - 1);
    for (char& c : body) if (c == ',' || c == '\n' || c == '\r') c = ' ';
    std::istringstream ss(body);
    int n = 0;
    std::string tok;
    while (ss >> tok) {
        if (n >= want) return false;
        out[n++] = std::strtod(tok.c_str(), nullptr);   // accepts "0." and "1.2e+003"
    }
    return n == want;
}

bool ReadCalibrationYaml(const std::string& path, double cam[9], double pro[9], double R[9], double T[3])
{
    // CCalculation.cpp:124-132: keys CamMat, ProMat, R, T
    std::ifstream f(path.c_str(), std::ios::in | std::ios::binary);
    if (!f) return false;
    std::stringstream buf;
    buf << f.rdbuf();
    const std::string text = buf.str();
    // "R:" and "T:" must match at line start so they are not found inside other keys
    auto find_line_key = [&](const std::string& key, int want, double* out) {
        size_t pos = 0;
        while ((pos = text.find(key + ":", pos)) != std::string::npos) {
            if (pos == 0 || text[pos - 1] == '\n') return parse_matrix(text.substr(pos), key, want, out);
            pos += key.size();
        }
        return false;
    };
    return find_line_key("CamMat", 9, cam) && find_line_key("ProMat", 9, pro) && find_line_key("R", 9, R) &&
           find_line_key("T", 3, T);
}

static slc_config make_cfg(const StaticParameters& sp, int projector_width, int gray_digits, int phase_steps)
{
    slc_config c;
    std::memset(&c, 0, sizeof(c));
    c.width = sp.CAMERA_RESLINE;
    c.height = sp.CAMERA_RESROW;
    c.projector_width = projector_width;
    c.gray_digits = gray_digits;
    c.phase_steps = phase_steps;
    c.fov_min = sp.FOV_MIN_DISTANCE;
    c.fov_max = sp.FOV_MAX_DISTANCE;
    c.modulation_min = sp.MODULATION_MIN;
    c.flags = 0;
    c.device = sp.CUDA_DEVICE;
    c.max_batch = 1;
    c.num_slots = 1;
    return c;
}

static bool copy_plane(const Mat& pic, const StaticParameters& sp, uint8_t* dst, const char* who)
{
    if (pic.empty() || pic.type() != CV_8UC1 || pic.rows != sp.CAMERA_RESROW || pic.cols != sp.CAMERA_RESLINE) {
        ErrorHandling(std::string(who) + "->picture must be CV_8UC1 of the camera resolution.");
        return false;
    }
    for (int r = 0; r < pic.rows; r++)
        std::memcpy(dst + (size_t)r * pic.cols, pic.ptr(r), (size_t)pic.cols);
    return true;
}

// ------------------------------------------------------------ CSensor -----
bool CSensor::InitSensor() { group_ = -1; now_ = 0; return true; }
bool CSensor::CloseSensor() { for (auto& g : groups_) g.clear(); group_ = -1; return true; }

bool CSensor::StoreDatas(int groupNum, int idx, const Mat& picture)
{
    if (groupNum < 0 || groupNum > 2 || idx < 0) return false;
    if ((size_t)idx >= groups_[groupNum].size()) groups_[groupNum].resize((size_t)idx + 1);
    picture.copyTo(groups_[groupNum][(size_t)idx]);
    return true;
}

void CSensor::AttachFiles(slc_context* ctx, const StaticParameters& sp, const std::string& groupDataPath)
{
    ctx_ = ctx;
    m_groupDataPath = groupDataPath;
    if (!m_groupDataPath.empty() && m_groupDataPath.back() != '/' && m_groupDataPath.back() != '\\') m_groupDataPath += '/';
    counts_[0] = sp.GRAY_V_NUMDIGIT * 2;       // CSensorV.cpp:74
    counts_[1] = sp.PHASE_NUMDIGIT;            // :82
    counts_[2] = sp.DYNAFRAME_MAXNUM;          // :90
    rows_ = sp.CAMERA_RESROW;
    cols_ = sp.CAMERA_RESLINE;
}

std::string CSensor::FileName(int groupNum, int idx) const
{
    // CSensorV.cpp:111-114: m_filePath + m_fileName + idx + m_fileSuffix
    std::ostringstream ss;
    ss << m_groupDataPath << (groupNum == 2 ? m_cFramePath : m_iFramePath)
       << (groupNum == 0 ? m_vGrayName : groupNum == 1 ? m_vPhaseName : m_dynaName) << idx << m_dataFileSuffix;
    return ss.str();
}

bool CSensor::LoadDatas(int groupNum)
{
    // CSensorV.cpp:60-133
    if (groupNum < 0 || groupNum > 2) { ErrorHandling("CSensor::LoadDatas->invalid groupNum."); return false; }
    group_ = groupNum;
    now_ = 0;
    if (ctx_ == nullptr || m_groupDataPath.empty() || !groups_[groupNum].empty()) return true;   // fed from memory
    std::vector<Mat>& g = groups_[groupNum];
    for (int i = 0; i < counts_[groupNum]; i++) {
        const std::string path = FileName(groupNum, i);
        std::ifstream f(path.c_str(), std::ios::binary | std::ios::ate);
        Mat pic;
        bool ok = (bool)f;
        if (ok) {
            const std::streamsize n = f.tellg();
            std::vector<char> bytes((size_t)(n > 0 ? n : 0));
            f.seekg(0);
            ok = n > 0 && (bool)f.read(bytes.data(), n);
            if (ok) {
                pic.create(rows_, cols_, CV_8UC1);
                ok = slc_bmp_decode_host(ctx_, bytes.data(), (int64_t)n, pic.ptr(), cols_, rows_) == SLC_OK;
            }
        }
        if (!ok) {
            if (groupNum == 2 && i >= 2) break;          // a shorter dynamic sequence than DYNAFRAME_MAXNUM
            ErrorHandling("CSensor::LoadPatterns::<Read>, imread error: " + path);   // :122-129 (reports, continues)
            pic = Mat();
        }
        g.push_back(pic);
    }
    return true;
}

bool CSensor::UnloadDatas() { group_ = -1; return true; }

bool CSensor::SetProPicture(int nowNum)
{
    if (group_ < 0 || nowNum < 0 || (size_t)nowNum >= groups_[group_].size()) return false;
    now_ = nowNum;
    return true;
}

Mat CSensor::GetCamPicture()
{
    Mat out;
    if (group_ >= 0 && (size_t)now_ < groups_[group_].size()) groups_[group_][(size_t)now_].copyTo(out);
    return out;
}

// -------------------------------------------------------- CDecodeGray -----
CDecodeGray::CDecodeGray(const StaticParameters& sp) : sp_(sp) {}
CDecodeGray::~CDecodeGray() { ReleaseSpace(); }

bool CDecodeGray::SetNumDigit(int numDigit, bool ver)
{
    if ((numDigit <= 0) || (numDigit > 16)) return false;          // CDecodeGray.cpp:39-40
    m_numDigit = numDigit;
    m_grayCodeSize = 1 << m_numDigit;
    m_vertical = ver;
    if (allocated_) ReleaseSpace();                                 // :48-49
    return AllocateSpace();
}

bool CDecodeGray::SetMatFileName(std::string codeFilePath, std::string codeFileName)
{
    m_codeFilePath = codeFilePath;
    m_codeFileName = codeFileName;
    return true;
}

bool CDecodeGray::AllocateSpace()
{
    if ((m_numDigit <= 0) || (m_numDigit > 16)) return false;
    const int pw = m_vertical ? sp_.PROJECTOR_RESLINE : sp_.PROJECTOR_RESROW;   // CDecodeGray.cpp:182-185
    slc_config cfg = make_cfg(sp_, pw, m_numDigit, 4);
    if (slc_create(&cfg, &ctx_) != SLC_OK) {
        ErrorHandling(std::string("CDecodeGray.AllocateSpace->") + slc_last_error(nullptr));
        ctx_ = nullptr;
        return false;
    }
    const size_t bytes = (size_t)2 * m_numDigit * sp_.CAMERA_RESROW * sp_.CAMERA_RESLINE;
    pinned_ = static_cast<uint8_t*>(slc_host_alloc(bytes));
    if (!pinned_) { ErrorHandling("CDecodeGray.AllocateSpace->pinned allocation failed."); ReleaseSpace(); return false; }
    std::memset(pinned_, 0, bytes);
    m_gray2bin.assign((size_t)m_grayCodeSize, 0);
    allocated_ = true;
    return true;
}

bool CDecodeGray::ReleaseSpace()
{
    if (pinned_) { slc_host_free(pinned_); pinned_ = nullptr; }
    if (ctx_) { slc_destroy(ctx_); ctx_ = nullptr; }
    m_gray2bin.clear();
    allocated_ = false;
    return true;
}

bool CDecodeGray::SetMat(int num, Mat pic)
{
    if (!allocated_) {                                              // CDecodeGray.cpp:26-30
        ErrorHandling("CDecodeGray.SetMat->grePicture Space is not allocated.");
        return false;
    }
    if (num < 0 || num >= 2 * m_numDigit) {
        ErrorHandling("CDecodeGray.SetMat->num out of range.");
        return false;
    }
    const size_t plane = (size_t)sp_.CAMERA_RESROW * sp_.CAMERA_RESLINE;
    return copy_plane(pic, sp_, pinned_ + (size_t)num * plane, "CDecodeGray.SetMat");   // :31 deep copy
}

bool CDecodeGray::Decode()
{
    if (!allocated_) { ErrorHandling("Gray Decode->Space is not allocated."); return false; }
    if (!ReadGrayCodeFile(m_codeFilePath + m_codeFileName, m_grayCodeSize, m_gray2bin)) {
        ErrorHandling("Gray Decode->Open file error.");             // CDecodeGray.cpp:115-119
        return false;
    }
    if (slc_set_gray_lut(ctx_, m_gray2bin.data(), m_grayCodeSize) != SLC_OK) {
        ErrorHandling(std::string("Gray Decode->") + slc_last_error(ctx_));
        return false;
    }
    m_result.create(sp_.CAMERA_RESROW, sp_.CAMERA_RESLINE, CV_64FC1);   // :187
    if (slc_decode_gray_host(ctx_, pinned_, reinterpret_cast<double*>(m_result.ptr()), nullptr) != SLC_OK) {
        ErrorHandling(std::string("Gray Decode->") + slc_last_error(ctx_));
        return false;
    }
    return true;
}

Mat CDecodeGray::GetResult()
{
    Mat result;                                                     // CDecodeGray.cpp:142-147
    m_result.copyTo(result);
    return result;
}

// ------------------------------------------------------- CDecodePhase -----
CDecodePhase::CDecodePhase(const StaticParameters& sp) : sp_(sp) {}
CDecodePhase::~CDecodePhase() { DeleteSpace(); }

bool CDecodePhase::DeleteSpace()
{
    if (pinned_) { slc_host_free(pinned_); pinned_ = nullptr; }
    if (ctx_) { slc_destroy(ctx_); ctx_ = nullptr; }
    allocated_ = false;
    return true;
}

bool CDecodePhase::SetNumMat(int numMat, int pixperiod)
{
    if ((numMat <= 0)) return false;                                // CDecodePhase.cpp:122-123
    m_numMat = numMat;
    m_pixPeroid = pixperiod;
    if (allocated_) DeleteSpace();                                  // :130-133
    // a 1-digit Gray geometry whose phase period PW / 2^0 is exactly pixperiod
    slc_config cfg = make_cfg(sp_, pixperiod, 1, numMat);
    if (slc_create(&cfg, &ctx_) != SLC_OK) {
        ErrorHandling(std::string("CDecodePhase.SetNumMat->") + slc_last_error(nullptr));
        ctx_ = nullptr;
        return false;
    }
    const size_t bytes = (size_t)numMat * sp_.CAMERA_RESROW * sp_.CAMERA_RESLINE;
    pinned_ = static_cast<uint8_t*>(slc_host_alloc(bytes));
    if (!pinned_) { ErrorHandling("CDecodePhase.SetNumMat->pinned allocation failed."); DeleteSpace(); return false; }
    std::memset(pinned_, 0, bytes);
    allocated_ = true;
    return true;
}

bool CDecodePhase::SetMat(int num, Mat pic)
{
    if (!allocated_) {                                              // CDecodePhase.cpp:109-113
        ErrorHandling("CDecodePhase.SetMat->grePicture Space is not allocated.");
        return false;
    }
    if (num < 0 || num >= m_numMat) {
        ErrorHandling("CDecodePhase.SetMat->num out of range.");
        return false;
    }
    const size_t plane = (size_t)sp_.CAMERA_RESROW * sp_.CAMERA_RESLINE;
    return copy_plane(pic, sp_, pinned_ + (size_t)num * plane, "CDecodePhase.SetMat");
}

bool CDecodePhase::Decode()
{
    if (!allocated_) { ErrorHandling("CDecodePhase.Decode()->CountResult fault"); return false; }
    m_result.create(sp_.CAMERA_RESROW, sp_.CAMERA_RESLINE, CV_64FC1);   // CDecodePhase.cpp:50
    if (slc_decode_phase_host(ctx_, pinned_, reinterpret_cast<double*>(m_result.ptr()), nullptr) != SLC_OK) {
        ErrorHandling("CDecodePhase.Decode()->CountResult fault");      // :88
        return false;
    }
    return true;
}

Mat CDecodePhase::GetResult()
{
    Mat result;                                                     // CDecodePhase.cpp:99-104
    m_result.copyTo(result);
    return result;
}

// ------------------------------------------------------- CCalculation -----
CCalculation::CCalculation(const StaticParameters& sp) : sp_(sp) {}
CCalculation::~CCalculation() { ReleaseSpace(); }

bool CCalculation::ReleaseSpace()
{
    if (m_sensor) { delete m_sensor; m_sensor = nullptr; }
    if (pinned_stack_) { slc_host_free(pinned_stack_); pinned_stack_ = nullptr; }
    m_dynXyzw.clear(); m_dynMask.clear(); m_dynDeltaZ.clear(); m_dynProjU.clear();
    if (m_dynBlock) { slc_host_free(m_dynBlock); m_dynBlock = nullptr; }
    if (m_textBuf) { slc_host_free(m_textBuf); m_textBuf = nullptr; }
    if (ctx_) { slc_destroy(ctx_); ctx_ = nullptr; }
    calibrated_ = false;
    return true;
}

bool CCalculation::Init()
{
    if (m_sensor != nullptr) return false;                          // CCalculation.cpp:80-81
    if (ctx_ != nullptr) return false;                              // :82-83
    m_sensor = new CSensor;                                         // :96-97
    m_sensor->InitSensor();
    slc_config cfg = make_cfg(sp_, sp_.PROJECTOR_RESLINE, sp_.GRAY_V_NUMDIGIT, sp_.PHASE_NUMDIGIT);
    if (slc_create(&cfg, &ctx_) != SLC_OK) {
        ErrorHandling(std::string("CCalculation::Init()->") + slc_last_error(nullptr));
        ctx_ = nullptr;
        ReleaseSpace();
        return false;
    }
    if (!m_groupDataPath.empty()) m_sensor->AttachFiles(ctx_, sp_, m_groupDataPath);
    const size_t npx = (size_t)sp_.CAMERA_RESROW * sp_.CAMERA_RESLINE;
    const size_t planes = (size_t)2 * sp_.GRAY_V_NUMDIGIT + sp_.PHASE_NUMDIGIT;
    pinned_stack_ = static_cast<uint8_t*>(slc_host_alloc(planes * npx));
    if (!pinned_stack_) { ErrorHandling("CCalculation::Init()->pinned allocation failed."); ReleaseSpace(); return false; }
    // :124-132 calibration
    double cam[9], pro[9], R[9], T[3];
    if (!ReadCalibrationYaml(sp_.DATA_PATH + m_paraFile, cam, pro, R, T)) {
        ErrorHandling("CCalculation::Init() OpenFile Error:" + sp_.DATA_PATH + m_paraFile);
        ReleaseSpace();
        return false;
    }
    if (slc_set_calibration(ctx_, cam, pro, R, T) != SLC_OK) {      // :135-166
        ErrorHandling(std::string("CCalculation::Init()->") + slc_last_error(ctx_));
        ReleaseSpace();
        return false;
    }
    calibrated_ = true;
    return true;
}

bool CCalculation::FillFirstProjectorUAndCoordinate()
{
    const size_t npx = (size_t)sp_.CAMERA_RESROW * sp_.CAMERA_RESLINE;
    const int G = sp_.GRAY_V_NUMDIGIT, N = sp_.PHASE_NUMDIGIT;
    // CCalculation.cpp:536-544: Gray images through the sensor, in SetMat order
    m_sensor->LoadDatas(0);
    for (int i = 0; i < G * 2; i++) {
        m_sensor->SetProPicture(i);
        Mat sensorMat = m_sensor->GetCamPicture();
        if (!copy_plane(sensorMat, sp_, pinned_stack_ + (size_t)i * npx, "CCalculation::FillFirstProjectorU")) return false;
    }
    // :549-557 phase images
    m_sensor->LoadDatas(1);
    for (int i = 0; i < N; i++) {
        m_sensor->SetProPicture(i);
        Mat sensorMat = m_sensor->GetCamPicture();
        if (!copy_plane(sensorMat, sp_, pinned_stack_ + (size_t)(2 * G + i) * npx, "CCalculation::FillFirstProjectorU")) return false;
    }
    // :538 + CDecodeGray.cpp:113-125 Gray code table
    if (!m_codeName.empty()) {
        std::vector<int16_t> lut;
        if (!ReadGrayCodeFile(m_codePath + m_codeName, 1 << G, lut)) {
            ErrorHandling("Gray Decode->Open file error.");
            return false;
        }
        if (slc_set_gray_lut(ctx_, lut.data(), 1 << G) != SLC_OK) {
            ErrorHandling(std::string("Gray Decode->") + slc_last_error(ctx_));
            return false;
        }
    }
    m_xyzw.create(sp_.CAMERA_RESROW, sp_.CAMERA_RESLINE, CV_32FC4);
    m_mask.create(sp_.CAMERA_RESROW, sp_.CAMERA_RESLINE, CV_8UC1);
    m_projU.create(sp_.CAMERA_RESROW, sp_.CAMERA_RESLINE, CV_64FC1);
    slc_parity_planes par;
    std::memset(&par, 0, sizeof(par));
    par.proj_u = reinterpret_cast<double*>(m_projU.ptr());
    // :545-589 decode + combine and :666-771 FillCoordinate(0): one fused launch
    if (slc_reconstruct_host(ctx_, pinned_stack_, 1, reinterpret_cast<float*>(m_xyzw.ptr()), m_mask.ptr(), &par) != SLC_OK) {
        ErrorHandling(std::string("CCalculation::CalculateFirst()->") + slc_last_error(ctx_));
        return false;
    }
    return true;
}

bool CCalculation::CalculateFirst()
{
    if (m_sensor == nullptr) return false;                          // CCalculation.cpp:176-177
    if (ctx_ == nullptr) return false;                              // :178-179
    if (!calibrated_) return false;                                 // :180-181
    std::cout << "Begin calculate first frame." << std::endl;       // :183
    if (!FillFirstProjectorUAndCoordinate()) return false;
    if (!m_pcFile.empty()) {                                        // :192-198
        printf("Begin writing...");
        Result(sp_.DATA_PATH + m_pcFile, 0);
        printf("Finished FirstFrame PointCloud.\n");
    }
    std::cout << "First frame finished." << std::endl;              // :203
    return true;
}

bool CCalculation::CalculateOther()
{
    if (m_sensor == nullptr) return false;                          // CCalculation.cpp:213-214
    if (ctx_ == nullptr) return false;                              // :215-216
    if (!calibrated_ || m_projU.empty()) return false;              // :217-218 (+ needs ProjectorU[0])
    const size_t npx = (size_t)sp_.CAMERA_RESROW * sp_.CAMERA_RESLINE;
    // :791-795 the dyna images through the sensor (group 2), frame 0 included (StripRegression(0), :201)
    m_sensor->LoadDatas(2);
    int n = 0;
    while (n < sp_.DYNAFRAME_MAXNUM && m_sensor->SetProPicture(n)) n++;
    if (n < 2) { ErrorHandling("CCalculation::CalculateOther()->fewer than two dynamic frames loaded."); return false; }
    uint8_t* frames = static_cast<uint8_t*>(slc_host_alloc((size_t)n * npx));
    if (!frames) { ErrorHandling("CCalculation::CalculateOther()->pinned allocation failed."); return false; }
    for (int i = 0; i < n; i++) {
        m_sensor->SetProPicture(i);
        Mat img = m_sensor->GetCamPicture();
        if (!copy_plane(img, sp_, frames + (size_t)i * npx, "CCalculation::StripRegression")) { slc_host_free(frames); return false; }
    }
    const size_t no = (size_t)n - 1;
    // one pinned block holds every map of the sequence, planes in descending alignment
    // (xyzw 16 B | ProjectorU 8 B | deltaZ 4 B | mask 1 B) so that every plane starts aligned for its
    // element type whatever the parity of the pixel count; the per-frame Mats are headers onto it, so
    // the download is the only copy
    const size_t per = npx * (16 + 8 + 4 + 1);
    if (m_dynBlock) { slc_host_free(m_dynBlock); m_dynBlock = nullptr; }
    m_dynXyzw.clear(); m_dynMask.clear(); m_dynDeltaZ.clear(); m_dynProjU.clear();
    m_dynBlock = static_cast<uint8_t*>(slc_host_alloc(no * per));
    if (!m_dynBlock) { slc_host_free(frames); ErrorHandling("CCalculation::CalculateOther()->pinned allocation failed."); return false; }
    float* xyzw = reinterpret_cast<float*>(m_dynBlock);
    double* pu = reinterpret_cast<double*>(m_dynBlock + no * npx * 16);
    float* dz = reinterpret_cast<float*>(m_dynBlock + no * npx * 24);
    uint8_t* mask = m_dynBlock + no * npx * 28;
    // StripRegression + FillOtherDeltaProU + FillCoordinate for every frame: two launches
    slc_dyna_parity par;
    std::memset(&par, 0, sizeof(par));
    par.proj_u = pu;                                // m_ProjectorU[f]: Result(f) prints from the f64 plane
    const int rc = slc_dyna_track_host(ctx_, frames, n, sp_.RECO_WINDOW_SIZE, reinterpret_cast<const double*>(m_projU.ptr()),
                                       xyzw, mask, dz, &par);
    slc_host_free(frames);
    if (rc != SLC_OK) {
        ErrorHandling(std::string("CCalculation::CalculateOther()->") + slc_last_error(ctx_));
        return false;
    }
    for (size_t f = 0; f < no; f++) {
        std::cout << "Frame: " << (f + 1) << " begin." << std::endl;       // :228
        m_dynXyzw.emplace_back(sp_.CAMERA_RESROW, sp_.CAMERA_RESLINE, CV_32FC4, xyzw + f * npx * 4);
        m_dynMask.emplace_back(sp_.CAMERA_RESROW, sp_.CAMERA_RESLINE, CV_8UC1, mask + f * npx);
        m_dynDeltaZ.emplace_back(sp_.CAMERA_RESROW, sp_.CAMERA_RESLINE, CV_32FC1, dz + f * npx);
        m_dynProjU.emplace_back(sp_.CAMERA_RESROW, sp_.CAMERA_RESLINE, CV_64FC1, pu + f * npx);
    }
    for (size_t f = 0; f < no; f++) {
        if (!m_pcDynaPrefix.empty()) {                                      // :309-315
            std::ostringstream name;
            name << sp_.DATA_PATH << m_pcDynaPrefix << (f + 1) << ".txt";
            Result(name.str(), (int)f + 1);
        }
    }
    return true;
}

bool CCalculation::Result(std::string fileName, int i)
{
    // CCalculation.cpp:323-357: "x y z\n" per in-FOV pixel, u outer / v inner, `file << double`.
    // The text is produced on the device from the f64 ProjectorU plane of frame i (x, y, z
    // recomputed in f64 in the reference's operation order, formatted like printf "%g") and
    // written with one fwrite.
    if (i < 0 || i >= FrameCount() || m_xyzw.empty()) return false;
    const Mat& U = (i == 0) ? m_projU : m_dynProjU.at((size_t)i - 1);
    FILE* file = std::fopen(fileName.c_str(), "wb");
    if (!file) {
        ErrorHandling("CCalculation::Result() OpenFile Error:" + fileName);
        return false;
    }
    const size_t npx = (size_t)sp_.CAMERA_RESROW * sp_.CAMERA_RESLINE;
    const int64_t cap = (int64_t)(npx * 43 + 16);
    if (!m_textBuf) m_textBuf = static_cast<char*>(slc_host_alloc((size_t)cap));      // pinned, kept for the next frame
    char* text = m_textBuf;
    bool ok = (text != nullptr);
    int64_t bytes = 0, points = 0;
    if (ok && slc_pointcloud_text_host(ctx_, reinterpret_cast<const double*>(U.ptr()), m_textFlags, text, cap, &bytes,
                                       &points) != SLC_OK) {
        ErrorHandling(std::string("CCalculation::Result()->") + slc_last_error(ctx_));
        ok = false;
    }
    if (ok && bytes > 0) ok = (std::fwrite(text, 1, (size_t)bytes, file) == (size_t)bytes);
    std::fclose(file);
    return ok;
}

bool CCalculation::ResultPly(std::string fileName, int i)
{
    // binary companion of Result(): the valid points of frame i, same u-outer / v-inner order,
    // compacted on the device, as a little-endian PLY
    if (i < 0 || i >= FrameCount() || m_xyzw.empty()) return false;
    const Mat& xyzw = PointMap(i);
    const Mat& maskm = ValidMask(i);
    const size_t npx = (size_t)sp_.CAMERA_RESROW * sp_.CAMERA_RESLINE;
    std::vector<float> xyz(npx * 3);
    int64_t points = 0;
    if (slc_pointcloud_compact_host(ctx_, reinterpret_cast<const float*>(xyzw.ptr()), maskm.ptr(), SLC_ORDER_REFERENCE,
                                    xyz.data(), (int64_t)npx, &points) != SLC_OK) {
        ErrorHandling(std::string("CCalculation::ResultPly()->") + slc_last_error(ctx_));
        return false;
    }
    FILE* file = std::fopen(fileName.c_str(), "wb");
    if (!file) {
        ErrorHandling("CCalculation::ResultPly() OpenFile Error:" + fileName);
        return false;
    }
    std::fprintf(file, "ply\nformat binary_little_endian 1.0\nelement vertex %lld\nproperty float x\nproperty float y\n"
                       "property float z\nend_header\n", (long long)points);
    const bool ok = std::fwrite(xyz.data(), 12, (size_t)points, file) == (size_t)points;
    std::fclose(file);
    return ok;
}

static Mat channel_as_f64(const Mat& xyzw, int ch)
{
    Mat out;
    if (xyzw.empty()) return out;
    out.create(xyzw.rows, xyzw.cols, CV_64FC1);
    for (int r = 0; r < xyzw.rows; r++) {
        const float* src = reinterpret_cast<const float*>(xyzw.ptr(r));
        double* dst = reinterpret_cast<double*>(out.ptr(r));
        for (int c = 0; c < xyzw.cols; c++) dst[c] = (double)src[4 * c + ch];
    }
    return out;
}

Mat CCalculation::GetX() const { return channel_as_f64(m_xyzw, 0); }
Mat CCalculation::GetY() const { return channel_as_f64(m_xyzw, 1); }
Mat CCalculation::GetZ() const { return channel_as_f64(m_xyzw, 2); }
Mat CCalculation::GetProjectorU() const { return m_projU.clone(); }

// ---- CCalculationPool -------------------------------------------------------------------
CCalculationPool::CCalculationPool(const StaticParameters& sp) : sp_(sp) {}

CCalculationPool::~CCalculationPool()
{
    if (pool_) slc_pool_destroy(pool_);
}

bool CCalculationPool::Init(const std::vector<int>& devices)
{
    if (pool_ != nullptr) return false;                             // CCalculation.cpp:80-83
    if (devices.empty()) { ErrorHandling("CCalculationPool::Init()->no device listed."); return false; }
    slc_config cfg = make_cfg(sp_, sp_.PROJECTOR_RESLINE, sp_.GRAY_V_NUMDIGIT, sp_.PHASE_NUMDIGIT);
    cfg.max_batch = m_chunk;
    cfg.num_slots = m_slots;
    std::vector<int32_t> devs(devices.begin(), devices.end());
    if (slc_pool_create(&cfg, devs.data(), (int32_t)devs.size(), &pool_) != SLC_OK) {
        ErrorHandling(std::string("CCalculationPool::Init()->") + slc_pool_last_error(nullptr));
        pool_ = nullptr;
        return false;
    }
    auto fail = [&](const std::string& msg) {
        ErrorHandling(msg);
        slc_pool_destroy(pool_);
        pool_ = nullptr;
        return false;
    };
    // :124-132 calibration, replicated to every GPU
    double cam[9], pro[9], R[9], T[3];
    if (!ReadCalibrationYaml(sp_.DATA_PATH + m_paraFile, cam, pro, R, T))
        return fail("CCalculationPool::Init() OpenFile Error:" + sp_.DATA_PATH + m_paraFile);
    if (slc_pool_set_calibration(pool_, cam, pro, R, T) != SLC_OK)
        return fail(std::string("CCalculationPool::Init()->") + slc_pool_last_error(pool_));
    // CDecodeGray.cpp:113-125 Gray code table
    if (!m_codeName.empty()) {
        const int n = 1 << sp_.GRAY_V_NUMDIGIT;
        std::vector<int16_t> lut;
        if (!ReadGrayCodeFile(m_codePath + m_codeName, n, lut)) return fail("Gray Decode->Open file error.");
        if (slc_pool_set_gray_lut(pool_, lut.data(), n) != SLC_OK)
            return fail(std::string("Gray Decode->") + slc_pool_last_error(pool_));
    }
    return true;
}

int CCalculationPool::Devices() const { return slc_pool_size(pool_); }

bool CCalculationPool::ShardRange(int n, int member, int& lo, int& hi) const
{
    int64_t a = 0, b = 0;
    if (!pool_ || slc_shard_range(n, member, slc_pool_size(pool_), &a, &b) != SLC_OK) return false;
    lo = (int)a;
    hi = (int)b;
    return true;
}

bool CCalculationPool::CalculateFirstBatch(const uint8_t* stacks, int n, const slc_result& out)
{
    if (pool_ == nullptr) { ErrorHandling("CCalculationPool::CalculateFirstBatch()->Init first."); return false; }   // :176-181
    if (slc_pool_reconstruct_host(pool_, stacks, n, &out) != SLC_OK) {
        ErrorHandling(std::string("CCalculationPool::CalculateFirstBatch()->") + slc_pool_last_error(pool_));
        return false;
    }
    return true;
}

bool CCalculationPool::CalculateFirstBatch(const uint8_t* stacks, int n, float* xyzw, uint8_t* mask)
{
    slc_result r;
    std::memset(&r, 0, sizeof(r));
    r.format = SLC_RESULT_XYZW;
    r.xyzw = xyzw;
    r.mask = mask;
    return CalculateFirstBatch(stacks, n, r);
}

}  // namespace dynaframe
