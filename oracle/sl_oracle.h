/*
 * sl_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, no OpenCV) of the DynaFrame first-frame
 * reconstruction path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product library
 * (libslcalc_b200.so) never links or calls it.
 *
 * PARITY PINNING: the reference ships no tests or golden vectors (SURVEY.md
 * section 4).  This restatement is pinned against
 *   - the reference ITSELF: oracle/_ref/dynaframe_ref, the reference's own path
 *     sources compiled in place (oracle/Makefile) against a minimal OpenCV
 *     stand-in (oracle/ref_shim/); bit for bit on every f64 plane and byte for
 *     byte on the text clouds (tests/test_reference_pinning.py, and the
 *     committed run tests/golden/reference_run_g6n4.npz);
 *   - the container's cv2 4.13 scalar primitives (cv2.fastAtan2,
 *     cv2.subtract, cv2.gemm, cv2.blur) through committed fixtures
 *     tests/golden/ (.npz files) made by tests/golden/make_golden.py, and
 *   - the hand-derived known-answer tables of SURVEY.md section 8(c)
 *     (KAT-E, KAT-T) which follow the cited reference formulas.
 * The third-party arithmetic on the path (cvFastArctan, OpenCV 2.4.9
 * opencv_core, CDecodePhase.cpp:67) is restated from the published OpenCV
 * algorithm; that 2.4.9 used this exact polynomial is unpinned by any
 * reference-owned vector => "parity unpinned" for that one primitive.
 *
 * All paths cited are relative to /root/reference/DynaFrame/DynaFrame/.
 */
#ifndef SL_ORACLE_H_
#define SL_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t width;            /* CAMERA_RESLINE     StaticParameters.cpp:8  */
    int32_t height;           /* CAMERA_RESROW      StaticParameters.cpp:9  */
    int32_t projector_width;  /* PROJECTOR_RESLINE  StaticParameters.cpp:4  */
    int32_t gray_digits;      /* GRAY_V_NUMDIGIT    StaticParameters.cpp:16 */
    int32_t phase_steps;      /* PHASE_NUMDIGIT     StaticParameters.cpp:18 */
    double  fov_min;          /* FOV_MIN_DISTANCE   StaticParameters.cpp:34 */
    double  fov_max;          /* FOV_MAX_DISTANCE   StaticParameters.cpp:35 */
    float   modulation_min;   /* [EXT] b_min; 0 => disabled (reference-faithful) */
    int32_t threads;          /* 1 = the reference's execution model; >1 = OpenMP rows */
} slo_config;

/* Calibration as read from the OpenCV YAML (CCalculation.cpp:124-132),
 * row-major f64. */
typedef struct {
    double cam[9];
    double pro[9];
    double R[9];
    double T[3];
} slo_calib;

/* Every output pointer may be NULL (then that plane is computed internally
 * and dropped).  All planes are row-major [height][width]. */
typedef struct {
    double  *gray_val;   /* a4  CDecodeGray::CountResult result            */
    double  *phase_pix;  /* a6  CDecodePhase::CountResult result, (0,T]     */
    double  *proj_u;     /* a7  ProjectorU                                  */
    double  *x, *y, *z;  /* a9/a10; invalid => 0 (SURVEY hard part 4)       */
    int16_t *kbin;       /* [EXT] gray2bin[code]                            */
    int8_t  *corr;       /* [EXT] wrap correction taken, -1/0/+1            */
    uint8_t *mask;       /* [EXT] modulation_ok && U!=0 && fov_min<=z<=fov_max */
    uint8_t *mod_ok;     /* [EXT] modulation test alone                     */
} slo_outputs;

/* OpenCV cv::fastAtan2(y, x) (== cvFastArctan) restated; degrees in [0,360). */
float slo_fast_atan2(float y, float x);
void slo_phase_pix_array(const float *deg, float *pix, size_t n, int m_pixPeroid);
unsigned long long slo_check_div360(float lo, float hi, unsigned long long *checked);
void slo_fast_atan2_array(const float *y, const float *x, float *out, size_t n);

/* gray2bin table for a reflected binary Gray code with n_digits bits, in the
 * form CDecodeGray::Decode builds it from Patterns/vGrayCode.txt
 * (CDecodeGray.cpp:120-125): lut[bin ^ (bin>>1)] = bin. */
void slo_default_gray_lut(int n_digits, int16_t *lut);

/* Derived integer geometry (CDecodeGray.cpp:183, CCalculation.cpp:550,563). */
int slo_gray_period(const slo_config *cfg);   /* gp = PW / 2^G       */
int slo_phase_period(const slo_config *cfg);  /* T  = PW / 2^(G-1)   */

/* CCalculation::Init calibration block (CCalculation.cpp:135-166).
 * cC, cD: [height][width] f64 (may be NULL); A, B scalars; P is 3x4. */
void slo_calibration(const slo_config *cfg, const slo_calib *cal,
                     double *A, double *B, double *cC, double *cD, double P[12]);

/* The whole first-frame path for one stack, staged exactly as the reference
 * stages it (six full-image passes with f64 intermediates).  planes is
 * [2G+N][height][width] u8: Gray pairs (2b, 2b+1) LSB first, then N phase
 * images.  gray_lut may be NULL (default reflected code).  Returns 0, or -1 on
 * invalid configuration. */
int slo_reconstruct(const slo_config *cfg, const slo_calib *cal,
                    const int16_t *gray_lut, const uint8_t *planes,
                    const slo_outputs *out);

/* Standalone decoders (the CDecodeGray / CDecodePhase objects). */
int slo_decode_gray(const slo_config *cfg, const int16_t *gray_lut,
                    const uint8_t *gray_planes, double *gray_val, int16_t *kbin);
int slo_decode_phase(const slo_config *cfg, const uint8_t *phase_planes,
                     double *phase_pix, uint8_t *mod_ok);

/* Wall-clock seconds for `reps` back-to-back runs of slo_reconstruct on the
 * same stack (hot loops only, no I/O); writes per-rep seconds to secs[reps]. */
int slo_time_reconstruct(const slo_config *cfg, const slo_calib *cal,
                         const uint8_t *planes, int reps, double *secs);

/* Dynamic frames: StripRegression (CCalculation.cpp:789-892),
 * FillOtherDeltaProU (:595-663), FillCoordinate(i>0) + deltaZ (:666-775). */
void slo_strip_regression(const slo_config *cfg, int window, const uint8_t *CamMat,
                          float *stripB, float *stripW);
void slo_delta_p(const slo_config *cfg, const float *B0, const float *W0, const float *B1, const float *W1,
                 float *deltaP);
void slo_dyna_frame(const slo_config *cfg, const slo_calib *cal, const double *U0, const float *deltaP,
                    const double *z0, double *U1, double *x, double *y, double *z, double *deltaZ,
                    uint8_t *mask);

/* [EXT] over-determined triangulation from both projector coordinates; see the
 * definition above the function.  U, V, x, y, z: f64 [height][width]. */
void slo_triangulate_uv(const slo_config *cfg, const slo_calib *cal, const double *U, const double *V,
                        double *x, double *y, double *z, uint8_t *mask);

/* Point-cloud text: CCalculation::Result (CCalculation.cpp:323-357) --
 * u outer / v inner, pixels with z outside [fov_min, fov_max] skipped, "x y z" + line end with
 * each number as `ostream << double` prints it, i.e. printf("%g") (precision 6).  flags bit 0:
 * "\r\n" line ends (text-mode fstream on the reference's platform); bit 1: three exponent
 * digits (MSVC 2013 CRT).  Writes at most cap bytes; returns the size of the whole text;
 * *n_points = number of lines. */
long long slo_result_text(const slo_config *cfg, const double *x, const double *y, const double *z,
                          unsigned flags, char *out, long long cap, long long *n_points);
/* One number as above into out (>= 16 chars); returns its length. */
int slo_format_g6(double v, unsigned flags, char *out);

int slo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif /* SL_ORACLE_H_ */
