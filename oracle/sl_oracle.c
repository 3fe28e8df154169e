/*
 * sl_oracle.c -- TEST INFRASTRUCTURE ONLY (see sl_oracle.h for the parity
 * pinning statement).  Plain-C restatement of the DynaFrame first-frame path:
 *
 *   a3  CDecodeGray::Grey2Bin        CDecodeGray.cpp:150-176
 *   a4  CDecodeGray::CountResult     CDecodeGray.cpp:179-204
 *   a6  CDecodePhase::CountResult    CDecodePhase.cpp:48-80
 *   a7  FillFirstProjectorU combine  CCalculation.cpp:550,562-589
 *   a8  Init calibration / LUTs      CCalculation.cpp:135-166
 *   a9  FillCoordinate z             CCalculation.cpp:672-708
 *   a10 FillCoordinate x,y           CCalculation.cpp:756-771
 *
 * Types (u8 / f32 / f64) and evaluation order follow the reference line by
 * line; build with -O2 -ffp-contract=off (the reference was built MSVC
 * /fp:precise x64: SSE2 scalar, no contraction; x64/Release/DynaFrame.log:4).
 * Staging is the reference's too: six full-image passes over f64
 * intermediates, loops a8/a9/a10 column-outer -- so timing this file is timing
 * "the reference's CPU path" (threads=1), and with threads>1 the same loops
 * split across cores with OpenMP.
 */
#include "sl_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------ */
/* cv::fastAtan2 (OpenCV core mathfuncs.cpp), the body of cvFastArctan called
 * at CDecodePhase.cpp:67.  Degree-7 odd minimax polynomial on [0,1], f32
 * throughout, no FMA. */
static const float k_p1 = 0.9997878412794807f * (float)(180 / M_PI);
static const float k_p3 = -0.3258083974640975f * (float)(180 / M_PI);
static const float k_p5 = 0.1555786518463281f * (float)(180 / M_PI);
static const float k_p7 = -0.04432655554792128f * (float)(180 / M_PI);

float slo_fast_atan2(float y, float x)
{
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((k_p7 * c2 + k_p5) * c2 + k_p3) * c2 + k_p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((k_p7 * c2 + k_p5) * c2 + k_p3) * c2 + k_p1) * c;
    }
    if (x < 0)
        a = 180.f - a;
    if (y < 0)
        a = 360.f - a;
    return a;
}

void slo_fast_atan2_array(const float *y, const float *x, float *out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = slo_fast_atan2(y[i], x[i]);
}

/* CDecodePhase.cpp:69-75 for an array of angles (degrees) */
void slo_phase_pix_array(const float *deg, float *pix, size_t n, int m_pixPeroid)
{
    for (size_t i = 0; i < n; i++) {
        float p = (deg[i]) / (360) * (double)(m_pixPeroid);
        p += 0.5;
        if (p > m_pixPeroid) p -= m_pixPeroid;
        pix[i] = p;
    }
}

/* Property the CUDA kernel relies on (csrc/slc_device.cuh div360_rn): for every
 * f32 x in [lo, hi] the FMA sequence q0 = x*y, r = fma(-q0, 360, x),
 * q = fma(r, y, q0) with y = RN(1/360) equals the IEEE quotient x / 360.
 * Returns the number of mismatches. */
unsigned long long slo_check_div360(float lo, float hi, unsigned long long *checked)
{
    const float y = 1.0f / 360.0f;
    uint32_t ulo, uhi;
    unsigned long long bad = 0, n = 0;
    memcpy(&ulo, &lo, 4);
    memcpy(&uhi, &hi, 4);
    for (uint32_t u = ulo; u <= uhi; u++) {
        float x;
        memcpy(&x, &u, 4);
        const float q0 = x * y;
        const float r = fmaf(-q0, 360.f, x);
        const float q = fmaf(r, y, q0);
        if (q != x / 360.f) bad++;
        n++;
    }
    if (checked) *checked = n;
    return bad;
}

void slo_default_gray_lut(int n_digits, int16_t *lut)
{
    /* Patterns/vGrayCode.txt rows are "bin gray" with gray = bin^(bin>>1);
     * CDecodeGray.cpp:120-125 stores m_gray2bin[gray] = bin. */
    int n = 1 << n_digits;
    for (int bin = 0; bin < n; bin++)
        lut[bin ^ (bin >> 1)] = (int16_t)bin;
}

int slo_gray_period(const slo_config *cfg)
{
    return cfg->projector_width / (1 << cfg->gray_digits); /* CDecodeGray.cpp:183 */
}

int slo_phase_period(const slo_config *cfg)
{
    /* CCalculation.cpp:550: PROJECTOR_RESLINE / (1 << GRAY_V_NUMDIGIT - 1);
     * '-' binds tighter than '<<' (the C4554 warning in the build log). */
    return cfg->projector_width / (1 << (cfg->gray_digits - 1));
}

static int cfg_ok(const slo_config *cfg)
{
    if (cfg->width <= 0 || cfg->height <= 0) return 0;
    if (cfg->gray_digits <= 0 || cfg->gray_digits > 16) return 0; /* CDecodeGray.cpp:39 */
    if (cfg->phase_steps < 3) return 0;
    if (slo_gray_period(cfg) < 1) return 0; /* else :570 divides by zero */
    return 1;
}

/* ---- [EXT] over-determined (u, v) triangulation ------------------------- */
/* The reference decodes projector columns only and never touches row 1 of P.  Definition: each
 * decoded projector coordinate gives one linear equation in z, built from its row of P exactly
 * the way CCalculation.cpp:159-164,686-687 builds the column one,
 *     (c0 - c2*U) z = B*U - A,      (c1 - c2*V) z = B*V - E,
 * c_r(u,v) = (u-cu)*fv*P_r0 + (v-cv)*fu*P_r1 + fu*fv*P_r2,  A = fu*fv*P03, E = fu*fv*P13, B = fu*fv*P23,
 * and z is their least-squares solution.  Valid iff U != 0, V != 0 and fov_min <= z <= fov_max;
 * x, y as CCalculation.cpp:766-767. */
void slo_triangulate_uv(const slo_config *cfg, const slo_calib *cal, const double *U, const double *V,
                        double *xMat, double *yMat, double *zMat, uint8_t *mask)
{
    const int W = cfg->width, H = cfg->height;
    double P[12], A, B;
    slo_calibration(cfg, cal, &A, &B, NULL, NULL, P);
    const double fu = cal->cam[0], fv = cal->cam[4], cu = cal->cam[2], cv = cal->cam[5];
    const double E = fu * fv * P[7];
    const double k02 = fu * fv * P[2], k12 = fu * fv * P[6], k22 = fu * fv * P[10];
    for (int v = 0; v < H; v++) {
        for (int u = 0; u < W; u++) {
            const size_t p = (size_t)v * W + u;
            double x = 0, y = 0, z = 0;
            uint8_t ok = 0;
            if (U[p] != 0 && V[p] != 0) {
                const double du = (u - cu) * fv, dv = (v - cv) * fu;
                const double c0 = du * P[0] + dv * P[1] + k02;
                const double c1 = du * P[4] + dv * P[5] + k12;
                const double c2 = du * P[8] + dv * P[9] + k22;
                const double a1 = c0 - c2 * U[p], b1 = B * U[p] - A;
                const double a2 = c1 - c2 * V[p], b2 = B * V[p] - E;
                const double zd = (a1 * b1 + a2 * b2) / (a1 * a1 + a2 * a2);
                if (!((zd < cfg->fov_min) || (zd > cfg->fov_max))) {
                    ok = 1;
                    z = zd;
                    x = zd * (u - cu) / fu;
                    y = zd * (v - cv) / fv;
                }
            }
            xMat[p] = x; yMat[p] = y; zMat[p] = z;
            if (mask) mask[p] = ok;
        }
    }
}

/* ---- point-cloud text: CCalculation::Result (CCalculation.cpp:323-357) ------------------ */
/* `file << double` with default stream flags is printf("%g") with precision 6 (C++ [ostream.inserters.arithmetic]
 * -> num_put -> printf conversion %g).  glibc's printf converts exactly (round-half-even on the
 * binary value).  The MSVC 2013 CRT the reference was built with prints three exponent digits. */
int slo_format_g6(double v, unsigned flags, char *out)
{
    char buf[40];
    int n = snprintf(buf, sizeof buf, "%g", v);
    if (flags & 2u) {
        char *e = strchr(buf, 'e');
        if (e && strlen(e + 2) == 2) {          /* e+XX -> e+0XX */
            memmove(e + 3, e + 2, 3);
            e[2] = '0';
            n++;
        }
    }
    memcpy(out, buf, (size_t)n);
    return n;
}

long long slo_result_text(const slo_config *cfg, const double *x, const double *y, const double *z,
                          unsigned flags, char *out, long long cap, long long *n_points)
{
    const int W = cfg->width, H = cfg->height;
    long long n = 0, pts = 0;
    for (int u = 0; u < W; u++) {                     /* :336 */
        for (int v = 0; v < H; v++) {                 /* :338 */
            const size_t i = (size_t)v * W + u;
            const double valZ = z[i];
            if ((valZ < cfg->fov_min) || (valZ > cfg->fov_max)) continue;   /* :341-345 */
            char line[64];
            int l = slo_format_g6(x[i], flags, line);                       /* :348-350 */
            line[l++] = ' ';
            l += slo_format_g6(y[i], flags, line + l);
            line[l++] = ' ';
            l += slo_format_g6(z[i], flags, line + l);
            if (flags & 1u) line[l++] = '\r';
            line[l++] = '\n';
            if (n + l <= cap) memcpy(out + n, line, (size_t)l);
            n += l;
            pts++;
        }
    }
    if (n_points) *n_points = pts;
    return n;
}

int slo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static int n_threads(const slo_config *cfg)
{
    return cfg->threads > 1 ? cfg->threads : 1;
}

/* ------------------------------------------------------------------------ */
/* a8: CCalculation.cpp:135-166 */
void slo_calibration(const slo_config *cfg, const slo_calib *cal,
                     double *A, double *B, double *cC, double *cD, double Pout[12])
{
    const int W = cfg->width, H = cfg->height;
    double C[12], RT[12], P[12];
    /* :135-137  C = [CamMat | 0] */
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) C[r * 4 + c] = cal->cam[r * 3 + c];
        C[r * 4 + 3] = 0.0;
    }
    /* :141-145  P = ProMat * [R | T]  (cv::Mat product: sum over k in order) */
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) RT[r * 4 + c] = cal->R[r * 3 + c];
        RT[r * 4 + 3] = cal->T[r];
    }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += cal->pro[r * 3 + k] * RT[k * 4 + c];
            P[r * 4 + c] = s;
        }
    if (Pout) memcpy(Pout, P, sizeof(P));
    const double fu = C[0], fv = C[5], cu = C[2], cv = C[6];
    /* :151-152 */
    *A = fu * fv * P[3];
    *B = fu * fv * P[11];
    if (!cC && !cD) return;
    /* :155-166, u outer / v inner as in the reference */
    const int nt = n_threads(cfg);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int u = 0; u < W; u++) {
        for (int v = 0; v < H; v++) {
            if (cC)
                cC[(size_t)v * W + u] = (u - cu) * fv * P[0] + (v - cv) * fu * P[1] + fu * fv * P[2];
            if (cD)
                cD[(size_t)v * W + u] = (u - cu) * fv * P[8] + (v - cv) * fu * P[9] + fu * fv * P[10];
        }
    }
}

/* ------------------------------------------------------------------------ */
/* a3 + a4 */
static void gray_stage(const slo_config *cfg, const int16_t *lut,
                       const uint8_t *gray_planes, double *gray_val,
                       int16_t *kbin, uint8_t *bin_planes, uint8_t *tmp)
{
    const int W = cfg->width, H = cfg->height, G = cfg->gray_digits;
    const size_t npx = (size_t)W * H;
    const int nt = n_threads(cfg);
    (void)nt;
    /* a3 CDecodeGray.cpp:155-174: tempMat = pattern - inverse (cv::Mat
     * operator- on CV_8U saturates), then threshold > 0 -> 0xFF. */
    for (int b = 0; b < G; b++) {
        const uint8_t *pa = gray_planes + (size_t)(2 * b) * npx;
        const uint8_t *pb = gray_planes + (size_t)(2 * b + 1) * npx;
        uint8_t *bin = bin_planes + (size_t)b * npx;
#pragma omp parallel for num_threads(nt) schedule(static)
        for (int i = 0; i < H; i++)
            for (int j = 0; j < W; j++) {
                int d = (int)pa[(size_t)i * W + j] - (int)pb[(size_t)i * W + j];
                tmp[(size_t)i * W + j] = (uint8_t)(d < 0 ? 0 : d);
            }
#pragma omp parallel for num_threads(nt) schedule(static)
        for (int i = 0; i < H; i++)
            for (int j = 0; j < W; j++)
                bin[(size_t)i * W + j] = tmp[(size_t)i * W + j] > 0 ? 0xFF : 0;
    }
    /* a4 CDecodeGray.cpp:181-202 */
    double pixPeriod = 0;
    pixPeriod = cfg->projector_width / (1 << G);
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++) {
            unsigned grayCode = 0;
            for (int b = 0; b < G; b++)
                if (bin_planes[(size_t)b * npx + (size_t)i * W + j] == 255)
                    grayCode += 1u << b;
            int16_t bin = lut[grayCode & 0xFFFFu];
            if (kbin) kbin[(size_t)i * W + j] = bin;
            gray_val[(size_t)i * W + j] = (double)bin * pixPeriod;
        }
}

/* [EXT] N-step tables: I_k = a + b sin(theta + 2 pi k / N).
 * N == 4 uses the reference formula verbatim; even N != 4 uses the N/2
 * differences d_k = I_k - I_{k+N/2}; odd N the plain sums.  Accumulation is
 * sequential fmaf in f32 so CPU and GPU agree bit for bit. */
static void phase_tables(int N, float *ck, float *sk)
{
    for (int k = 0; k < N; k++) {
        double c = cos(2.0 * M_PI * (double)k / (double)N);
        double s = sin(2.0 * M_PI * (double)k / (double)N);
        if (fabs(c) < 1e-9) c = 0.0;
        if (fabs(s) < 1e-9) s = 0.0;
        ck[k] = (float)c;
        sk[k] = (float)s;
    }
}

static float modulation_thr2(const slo_config *cfg)
{
    /* amplitude of (S, Cc) is b for N == 4 (reference /2 kept), (N/2) b else */
    double scale = cfg->phase_steps == 4 ? 1.0 : 0.5 * (double)cfg->phase_steps;
    double t = (double)cfg->modulation_min * scale;
    return (float)(t * t);
}

/* a6 CDecodePhase.cpp:48-80 */
static void phase_stage(const slo_config *cfg, const uint8_t *phase_planes,
                        double *phase_pix, uint8_t *mod_ok)
{
    const int W = cfg->width, H = cfg->height, N = cfg->phase_steps;
    const size_t npx = (size_t)W * H;
    const int m_pixPeroid = slo_phase_period(cfg);
    const int nt = n_threads(cfg);
    (void)nt;
    float ck[64], sk[64];
    phase_tables(N > 64 ? 64 : N, ck, sk);
    const float thr2 = modulation_thr2(cfg);
    const int use_mod = cfg->modulation_min > 0.0f;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int i = 0; i < H; i++) {
        for (int j = 0; j < W; j++) {
            const size_t p = (size_t)i * W + j;
            float sinValue, cosValue;
            if (N == 4) {
                float greyValue0 = phase_planes[0 * npx + p];
                float greyValue1 = phase_planes[1 * npx + p];
                float greyValue2 = phase_planes[2 * npx + p];
                float greyValue3 = phase_planes[3 * npx + p];
                sinValue = (greyValue0 - greyValue2) / 2;
                cosValue = (greyValue1 - greyValue3) / 2;
            } else if ((N & 1) == 0) {
                sinValue = 0.f;
                cosValue = 0.f;
                for (int k = 0; k < N / 2; k++) {
                    float d = (float)phase_planes[(size_t)k * npx + p] -
                              (float)phase_planes[(size_t)(k + N / 2) * npx + p];
                    sinValue = fmaf(d, ck[k], sinValue);
                    cosValue = fmaf(d, sk[k], cosValue);
                }
            } else {
                sinValue = 0.f;
                cosValue = 0.f;
                for (int k = 0; k < N; k++) {
                    float g = (float)phase_planes[(size_t)k * npx + p];
                    sinValue = fmaf(g, ck[k], sinValue);
                    cosValue = fmaf(g, sk[k], cosValue);
                }
            }
            /* :67 */
            float x = slo_fast_atan2(sinValue, cosValue);
            /* :69-75, usual arithmetic conversions as written */
            float pix = (x) / (360) * (double)(m_pixPeroid);
            pix += 0.5;
            if (pix > m_pixPeroid) {
                pix -= m_pixPeroid;
            }
            phase_pix[p] = (double)pix;
            if (mod_ok) {
                float m2 = sinValue * sinValue + cosValue * cosValue;
                mod_ok[p] = (!use_mod || m2 >= thr2) ? 1 : 0;
            }
        }
    }
}

/* a7 CCalculation.cpp:562-589 */
static void combine_stage(const slo_config *cfg, const double *vGrayMat,
                          double *vPhaseMat, double *ProjectorU, int8_t *corr,
                          double *vProjectorMat)
{
    const int W = cfg->width, H = cfg->height, G = cfg->gray_digits;
    const int v_pixPeriod = slo_phase_period(cfg);
    const int vGrayNum = 1 << G;
    const int vGrayPeriod = cfg->projector_width / vGrayNum;
    const int nt = n_threads(cfg);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int h = 0; h < H; h++) {
        for (int w = 0; w < W; w++) {
            const size_t p = (size_t)h * W + w;
            double grayVal = vGrayMat[p];
            double phaseVal = vPhaseMat[p];
            int8_t c = 0;
            if ((int)(grayVal / vGrayPeriod) % 2 == 0) {
                if (phaseVal > (double)v_pixPeriod * 0.75) {
                    vPhaseMat[p] = phaseVal - v_pixPeriod;
                    c = -1;
                }
            } else {
                if (phaseVal < (double)v_pixPeriod * 0.25) {
                    vPhaseMat[p] = phaseVal + v_pixPeriod;
                    c = 1;
                }
                vPhaseMat[p] = vPhaseMat[p] - 0.5 * v_pixPeriod;
            }
            if (corr) corr[p] = c;
        }
    }
    /* :587 vProjectorMat = vGrayMat + vPhaseMat;  :589 copyTo(m_ProjectorU[0]) */
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int h = 0; h < H; h++)
        for (int w = 0; w < W; w++)
            vProjectorMat[(size_t)h * W + w] = vGrayMat[(size_t)h * W + w] + vPhaseMat[(size_t)h * W + w];
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int h = 0; h < H; h++)
        memcpy(ProjectorU + (size_t)h * W, vProjectorMat + (size_t)h * W, sizeof(double) * (size_t)W);
}

/* a9 + a10 CCalculation.cpp:672-708, 756-771 (u outer / v inner) */
static void coordinate_stage(const slo_config *cfg, const slo_calib *cal,
                             double cA, double cB, const double *cC, const double *cD,
                             const double *ProjectorU, const uint8_t *mod_ok,
                             double *xMat, double *yMat, double *zMat, uint8_t *mask)
{
    const int W = cfg->width, H = cfg->height;
    const double FOV_MIN_DISTANCE = cfg->fov_min, FOV_MAX_DISTANCE = cfg->fov_max;
    const int nt = n_threads(cfg);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int u = 0; u < W; u++) {
        for (int v = 0; v < H; v++) {
            const size_t p = (size_t)v * W + u;
            double z = 0;
            uint8_t ok = 0;
            /* [EXT] a pixel failing the modulation test is treated like the
             * reference's "no value" case (:678-682) */
            if (ProjectorU[p] == 0 || (mod_ok && !mod_ok[p])) {
                z = 0;
            } else {
                z = -(cA - cB * ProjectorU[p]) / (cC[p] - cD[p] * ProjectorU[p]);
                ok = 1;
                if ((z < FOV_MIN_DISTANCE) || (z > FOV_MAX_DISTANCE)) {
                    z = 0;
                    ok = 0;
                }
            }
            /* the reference leaves skipped pixels unwritten in an
             * uninitialised Mat; the oracle defines them as 0 */
            zMat[p] = z;
            if (mask) mask[p] = ok;
        }
    }
    const double cu = cal->cam[2], cv = cal->cam[5];
    const double fu = cal->cam[0], fv = cal->cam[4];
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int u = 0; u < W; u++) {
        for (int v = 0; v < H; v++) {
            const size_t p = (size_t)v * W + u;
            double z = zMat[p];
            double uc = u - cu;
            double vc = v - cv;
            xMat[p] = z * uc / fu;
            yMat[p] = z * vc / fv;
        }
    }
}

/* ------------------------------------------------------------------------ */
typedef struct {
    uint8_t *bin_planes, *tmp, *mod_ok;
    int16_t *lut;
    double *gray, *phase, *proj, *U, *cC, *cD, *x, *y, *z;
    double A, B;
} workspace;

static void ws_free(workspace *w)
{
    free(w->bin_planes); free(w->tmp); free(w->mod_ok); free(w->lut);
    free(w->gray); free(w->phase); free(w->proj); free(w->U);
    free(w->cC); free(w->cD); free(w->x); free(w->y); free(w->z);
    memset(w, 0, sizeof(*w));
}

static int ws_alloc(const slo_config *cfg, workspace *w, int with_calib)
{
    const size_t npx = (size_t)cfg->width * cfg->height;
    memset(w, 0, sizeof(*w));
    w->bin_planes = (uint8_t *)malloc(npx * (size_t)cfg->gray_digits);
    w->tmp = (uint8_t *)malloc(npx);
    w->mod_ok = (uint8_t *)malloc(npx);
    w->lut = (int16_t *)calloc(65536, sizeof(int16_t));
    w->gray = (double *)malloc(npx * sizeof(double));
    w->phase = (double *)malloc(npx * sizeof(double));
    w->proj = (double *)malloc(npx * sizeof(double));
    w->U = (double *)malloc(npx * sizeof(double));
    int ok = w->bin_planes && w->tmp && w->mod_ok && w->lut && w->gray && w->phase && w->proj && w->U;
    if (with_calib) {
        w->cC = (double *)malloc(npx * sizeof(double));
        w->cD = (double *)malloc(npx * sizeof(double));
        w->x = (double *)malloc(npx * sizeof(double));
        w->y = (double *)malloc(npx * sizeof(double));
        w->z = (double *)malloc(npx * sizeof(double));
        ok = ok && w->cC && w->cD && w->x && w->y && w->z;
    }
    if (!ok) { ws_free(w); return -1; }
    return 0;
}

static void hot_loops(const slo_config *cfg, const slo_calib *cal,
                      const uint8_t *planes, workspace *w, int16_t *kbin,
                      int8_t *corr, uint8_t *mask, double *phase_raw)
{
    const size_t npx = (size_t)cfg->width * cfg->height;
    const int use_mod = cfg->modulation_min > 0.0f;
    gray_stage(cfg, w->lut, planes, w->gray, kbin, w->bin_planes, w->tmp);
    phase_stage(cfg, planes + (size_t)(2 * cfg->gray_digits) * npx, w->phase, w->mod_ok);
    if (phase_raw) memcpy(phase_raw, w->phase, npx * sizeof(double)); /* parity output only */
    /* the combine stage edits vPhaseMat in place (CCalculation.cpp:574-583) */
    combine_stage(cfg, w->gray, w->phase, w->U, corr, w->proj);
    /* [EXT] a pixel the modulation test rejects has no projector column: ProjectorU = 0, the
     * reference's own "no value" sentinel (CCalculation.cpp:678), so that Result(), FillCoordinate(i)
     * and the dynamic frames -- which only test U == 0 -- skip it as well */
    if (use_mod)
        for (size_t q = 0; q < npx; q++)
            if (!w->mod_ok[q]) w->U[q] = 0.0;
    coordinate_stage(cfg, cal, w->A, w->B, w->cC, w->cD, w->U,
                     use_mod ? w->mod_ok : NULL, w->x, w->y, w->z, mask);
}

int slo_reconstruct(const slo_config *cfg, const slo_calib *cal,
                    const int16_t *gray_lut, const uint8_t *planes,
                    const slo_outputs *out)
{
    if (!cfg_ok(cfg)) return -1;
    const size_t npx = (size_t)cfg->width * cfg->height;
    workspace w;
    if (ws_alloc(cfg, &w, 1)) return -2;
    if (gray_lut) memcpy(w.lut, gray_lut, sizeof(int16_t) * ((size_t)1 << cfg->gray_digits));
    else slo_default_gray_lut(cfg->gray_digits, w.lut);
    slo_calibration(cfg, cal, &w.A, &w.B, w.cC, w.cD, NULL);

    hot_loops(cfg, cal, planes, &w, out ? out->kbin : NULL, out ? out->corr : NULL,
              out ? out->mask : NULL, out ? out->phase_pix : NULL);
    if (out) {
        if (out->gray_val) memcpy(out->gray_val, w.gray, npx * sizeof(double));
        if (out->proj_u) memcpy(out->proj_u, w.U, npx * sizeof(double));
        if (out->x) memcpy(out->x, w.x, npx * sizeof(double));
        if (out->y) memcpy(out->y, w.y, npx * sizeof(double));
        if (out->z) memcpy(out->z, w.z, npx * sizeof(double));
        if (out->mod_ok) memcpy(out->mod_ok, w.mod_ok, npx);
    }
    ws_free(&w);
    return 0;
}

int slo_decode_gray(const slo_config *cfg, const int16_t *gray_lut,
                    const uint8_t *gray_planes, double *gray_val, int16_t *kbin)
{
    if (!cfg_ok(cfg)) return -1;
    workspace w;
    if (ws_alloc(cfg, &w, 0)) return -2;
    if (gray_lut) memcpy(w.lut, gray_lut, sizeof(int16_t) * ((size_t)1 << cfg->gray_digits));
    else slo_default_gray_lut(cfg->gray_digits, w.lut);
    gray_stage(cfg, w.lut, gray_planes, gray_val, kbin, w.bin_planes, w.tmp);
    ws_free(&w);
    return 0;
}

int slo_decode_phase(const slo_config *cfg, const uint8_t *phase_planes,
                     double *phase_pix, uint8_t *mod_ok)
{
    if (!cfg_ok(cfg)) return -1;
    phase_stage(cfg, phase_planes, phase_pix, mod_ok);
    return 0;
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int slo_time_reconstruct(const slo_config *cfg, const slo_calib *cal,
                         const uint8_t *planes, int reps, double *secs)
{
    if (!cfg_ok(cfg)) return -1;
    workspace w;
    if (ws_alloc(cfg, &w, 1)) return -2;
    slo_default_gray_lut(cfg->gray_digits, w.lut);
    /* calibration LUT set-up is a once-per-session step (Init), not timed */
    slo_calibration(cfg, cal, &w.A, &w.B, w.cC, w.cD, NULL);
    for (int r = 0; r < reps; r++) {
        double t0 = now_s();
        hot_loops(cfg, cal, planes, &w, NULL, NULL, NULL, NULL);
        secs[r] = now_s() - t0;
    }
    ws_free(&w);
    return 0;
}


/* ======================================================================== */
/* Dynamic frames ("next" row):
 *   StripRegression       CCalculation.cpp:789-892
 *   FillOtherDeltaProU    CCalculation.cpp:595-663   (cv::blur 3x3, BORDER_REFLECT_101)
 *   FillCoordinate(i>0)   CCalculation.cpp:666-775   (incl. deltaZ)
 * cv::blur on CV_32F sums in double and stores (float)(sum * (1./9)) -- pinned
 * against the container's cv2.blur in tests/golden/make_golden.py. */

void slo_strip_regression(const slo_config *cfg, int window, const uint8_t *CamMat,
                          float *stripB, float *stripW)
{
    const int W = cfg->width, H = cfg->height;
    const int RECO_WINDOW_SIZE = window;
    const int half = RECO_WINDOW_SIZE / 2;
    const size_t npx = (size_t)W * H;
    const int nt = n_threads(cfg);
    (void)nt;
    float *valSum = (float *)calloc(npx, sizeof(float));            /* :798-800 */
    /* :801-812 first row of sums */
    for (int w = half; w < W - half; w++) {
        float sum = 0;
        for (int hc = 0; hc < RECO_WINDOW_SIZE; hc++) sum += (float)CamMat[(size_t)hc * W + w];
        valSum[(size_t)half * W + w] = sum;
    }
    /* :814-823 running update down the rows */
    for (int h = half + 1; h < H - half; h++)
        for (int w = half; w < W - half; w++)
            valSum[(size_t)h * W + w] = valSum[(size_t)(h - 1) * W + w]
                - (float)CamMat[(size_t)(h - half - 1) * W + w]
                + (float)CamMat[(size_t)(h + half) * W + w];
    /* :826-889 */
    memset(stripB, 0, npx * sizeof(float));
    memset(stripW, 0, npx * sizeof(float));
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int h = half; h < H - half; h++) {
        for (int w = half; w < W - half; w++) {
            float max = valSum[(size_t)h * W + w];
            float maxIdx = 0;
            float min = valSum[(size_t)h * W + w];
            float minIdx = 0;
            for (int i = -half; i < half; i++) {
                float value = valSum[(size_t)h * W + w + i];
                if (value > max) { max = value; maxIdx = i; }
                if (value < min) { min = value; minIdx = i; }
            }
            stripB[(size_t)h * W + w] = minIdx;
            stripW[(size_t)h * W + w] = maxIdx;
        }
    }
    free(valSum);
}

static int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}

/* :599-650: nearer-of-two delta, then blur(temp, deltaP, Size(3,3)) */
void slo_delta_p(const slo_config *cfg, const float *B0, const float *W0, const float *B1, const float *W1,
                 float *deltaP)
{
    const int W = cfg->width, H = cfg->height;
    const size_t npx = (size_t)W * H;
    const int nt = n_threads(cfg);
    (void)nt;
    float *temp = (float *)malloc(npx * sizeof(float));
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int h = 0; h < H; h++)
        for (int w = 0; w < W; w++) {
            const size_t p = (size_t)h * W + w;
            float f0W = W0[p], f0B = B0[p], f1W = W1[p], f1B = B1[p];
            float fBbias = fabsf(f0B - f1B);
            float fWbias = fabsf(f0W - f1W);
            temp[p] = (fBbias < fWbias) ? (f0B - f1B) : (f0W - f1W);
        }
    const double scale = 1. / 9;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int h = 0; h < H; h++)
        for (int w = 0; w < W; w++) {
            double sum = 0;
            for (int dy = -1; dy <= 1; dy++) {
                const int hh = reflect101(h + dy, H);
                double rs = 0;
                for (int dx = -1; dx <= 1; dx++) rs += (double)temp[(size_t)hh * W + reflect101(w + dx, W)];
                sum += rs;
            }
            deltaP[(size_t)h * W + w] = (float)(sum * scale);
        }
    free(temp);
}

/* One dynamic frame: U1 = U0 + deltaP (:652-660), FillCoordinate (:672-771), deltaZ (:772-775).
 * z0 is the previous frame's z plane.  Outputs may alias nothing. */
void slo_dyna_frame(const slo_config *cfg, const slo_calib *cal, const double *U0, const float *deltaP,
                    const double *z0, double *U1, double *x, double *y, double *z, double *deltaZ,
                    uint8_t *mask)
{
    const int W = cfg->width, H = cfg->height;
    const size_t npx = (size_t)W * H;
    double A, B;
    double *cC = (double *)malloc(npx * sizeof(double));
    double *cD = (double *)malloc(npx * sizeof(double));
    slo_calibration(cfg, cal, &A, &B, cC, cD, NULL);
    for (size_t p = 0; p < npx; p++) U1[p] = U0[p] + deltaP[p];
    slo_config c1 = *cfg;
    c1.modulation_min = 0.f;
    coordinate_stage(&c1, cal, A, B, cC, cD, U1, NULL, x, y, z, mask);
    if (deltaZ)
        for (size_t p = 0; p < npx; p++) deltaZ[p] = z[p] - z0[p];
    free(cC);
    free(cD);
}
