// slc_dyna.cu -- dynamic frames: the reference's CalculateOther path
// (CCalculation.cpp:208-320).
//
//   strip_regression21_kernel  StripRegression (CCalculation.cpp:789-892) for the reference's window
//                            of 21: 21-row box sum per column, then the offset of the minimum /
//                            maximum of that sum over the 20 columns [w-10, w+9] -> (stripB, stripW)
//                            per pixel.  Independent per frame: one launch covers every frame
//                            (grid.z).  strip_regression_kernel: any other odd window / alignment.
//   dyna_fused_kernel        FillOtherDeltaProU (CCalculation.cpp:595-663) + FillCoordinate(i>0)
//                            (CCalculation.cpp:666-775): nearer-of-two delta, 3x3 cv::blur
//                            (BORDER_REFLECT_101, double sum * (1./9) narrowed to f32),
//                            U[f] = U[f-1] + deltaP, triangulation, deltaZ.  The frame-to-frame
//                            recurrence is per pixel, so a block walks all frames of its tile:
//                            a single launch for the whole sequence, no intermediate plane.
//   delta_sum_generic_kernel + dyna_track_kernel   the same in two kernels through a u16 plane of
//                            3x3 sums, for widths that are not a multiple of 8.
//
// Everything up to U is integer / exactly-representable arithmetic (sums of <= 21 u8 values,
// index differences, multiples of 2^-27 well inside a double), so strips, deltaP and U are
// bit-exact against the CPU path whatever the summation order.
#include "slc_kernels.h"

#include <cstdlib>

namespace slc {

namespace {

constexpr int kSrTileW = 128;
constexpr int kSrTileH = 32;
constexpr int kSrMaxHalf = 16;                      // window <= 33
constexpr int kSrStride = kSrTileW + 2 * kSrMaxHalf;

__device__ __forceinline__ int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
    return p;
}

__global__ void __launch_bounds__(256)
strip_regression_kernel(const uint8_t* __restrict__ frames, char2* __restrict__ strips, int W, int H, int half)
{
    __shared__ unsigned short s_sum[kSrTileH][kSrStride];
    const long long npx = (long long)W * H;
    const uint8_t* img = frames + (long long)blockIdx.z * npx;
    char2* out = strips + (long long)blockIdx.z * npx;
    const int x0 = blockIdx.x * kSrTileW, y0 = blockIdx.y * kSrTileH;
    const int ncols = kSrTileW + 2 * half;

    // phase 1: valSum(h, c) for the tile rows and tile columns +- half (CCalculation.cpp:801-823);
    // 0 outside [half, H-half) x [half, W-half) exactly like the zero-initialised valSum Mat
    for (int t = threadIdx.x; t < ncols; t += blockDim.x) {
        const int c = x0 - half + t;
        const bool col_ok = (c >= half) && (c < W - half);
        int sum = 0;
        bool have = false;
        for (int r = 0; r < kSrTileH; r++) {
            const int h = y0 + r;
            const bool ok = col_ok && (h >= half) && (h < H - half);
            if (ok) {
                if (!have) {
                    sum = 0;
                    for (int k = h - half; k <= h + half; k++) sum += img[(long long)k * W + c];
                    have = true;
                } else {
                    sum += (int)img[(long long)(h + half) * W + c] - (int)img[(long long)(h - half - 1) * W + c];
                }
            }
            s_sum[r][t] = ok ? (unsigned short)sum : (unsigned short)0;
        }
    }
    __syncthreads();

    // phase 2: offsets of the window minimum / maximum (CCalculation.cpp:828-889).  The scan
    // starts from the centre value with index 0 and replaces on strict </>, so a tie with the
    // centre keeps 0 and otherwise the first extremum wins.
    for (int idx = threadIdx.x; idx < kSrTileW * kSrTileH; idx += blockDim.x) {
        const int r = idx / kSrTileW, x = idx % kSrTileW;
        const int h = y0 + r, w = x0 + x;
        if (h >= H || w >= W) continue;
        char2 res = make_char2(0, 0);
        if (h >= half && h < H - half && w >= half && w < W - half) {
            const unsigned short* row = &s_sum[r][x];      // row[half + i] == valSum(h, w + i)
            int mx = row[half], mn = row[half], mxi = 0, mni = 0;
            for (int i = -half; i < half; i++) {
                const int v = row[half + i];
                if (v > mx) { mx = v; mxi = i; }
                if (v < mn) { mn = v; mni = i; }
            }
            res = make_char2((signed char)mni, (signed char)mxi);   // (stripB, stripW)
        }
        out[(long long)h * W + w] = res;
    }
}

// ---- StripRegression, window 21 (the reference's RECO_WINDOW_SIZE), fast path ----------
// Tile: 160 output columns x 64 rows per block.
//  Phase 1 (184 threads): a thread owns 4 adjacent columns (one 32-bit load per image row) and
//    16 consecutive rows; one running 21-row sum per column, built once from 21 rows and then
//    slid down 15 times with byte dot products.  Each sum goes to shared memory as a key
//    sum << 9 | 0x100 | tile_column.
//  Phase 2 (256 threads): a thread owns 20 consecutive outputs of one row.  Their 20-column
//    windows [w-10, w+9] all straddle one boundary between two 20-element blocks, so the window
//    minimum is min(suffix-minimum of block 1, prefix-minimum of block 2) (van Herk / Gil-Werman):
//    3 min ops per output instead of 19.  The maximum is the minimum of the complemented key
//    (8191 - sum) << 9 | 0x100 | tile_column, so for both the smaller column wins a tie -- the
//    first occurrence in the reference's left-to-right scan (CCalculation.cpp:833-849).  The scan
//    starts from the centre with offset 0 and replaces on strict </> only, i.e. the centre wins
//    every tie: a third candidate, the centre's key with bit 8 cleared, does that in the same
//    3-input min, and its column field decodes to offset 0.
constexpr int kVhW = 160, kVhHalf = 10, kVhPad = 12;
constexpr int kVhCols = kVhW + 2 * kVhPad;        // 184 key columns: tile column tc <-> image column x0 - 12 + tc
constexpr int kVhQuads = kVhCols / 4;             // 46
constexpr int kVhStride = kVhCols + 2;            // 186 words: conflict-free 8-byte reads in phase 2
constexpr int kVhOut = 20;                        // outputs per phase-2 thread (= window - 1)
constexpr unsigned kVhFlag = 0x100u;
constexpr unsigned kVhInv = 8191u << 9;

// d = c + sum_i (unsigned byte i of a) * (signed byte i of b)
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// kRowIn / kColIn: every image row / column the tile touches exists and every sum row / column of
// the tile lies inside the reference's [10, H-10) / [10, W-10) region, so nothing along that axis
// needs a predicate.  Both: no predicates anywhere (3/4 of the tiles of a 1280x1024 image).
template <bool kRowIn, bool kColIn, int kVhH, int kVhSeg>       // + tile rows, rows per phase-1 thread
__device__ __forceinline__ void strip21_tile(const uint8_t* __restrict__ img, char2* __restrict__ out, int W, int H,
                                             int x0, int y0, uint32_t (*s_key)[kVhStride])
{
    constexpr bool kInterior = kRowIn && kColIn;
    const int t = threadIdx.x;
    if (t < kVhQuads * (kVhH / kVhSeg)) {
        const int q = t % kVhQuads, seg = t / kVhQuads;
        const int c0 = x0 - kVhPad + 4 * q;               // W % 4 == 0: the quad is inside or outside as a whole
        const int hb = y0 + seg * kVhSeg;
        // a quad outside the image is read from a clamped column instead: its sums are never used
        // (cok below is false for every column of it), so only the address has to be legal
        const uint8_t* colp = img + (kColIn ? c0 : min(max(c0, 0), W - 4));
        auto ld = [&](int row) -> uint32_t {
            if (kRowIn) return __ldg(reinterpret_cast<const uint32_t*>(colp + (long long)row * W));
            return (row >= 0 && row < H) ? __ldg(reinterpret_cast<const uint32_t*>(colp + (long long)row * W)) : 0u;
        };
        // one 32-bit running sum per column, maintained with byte dot products (IDP4A: selector byte 1
        // adds a row's pixel, selector byte -1 removes one) -- they issue on the FMA pipe, beside the
        // integer min / max work of phase 2 that saturates the ALU pipe
        int sum[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = -kVhHalf; k <= kVhHalf; k++) {
            const uint32_t w = ld(hb + k);
#pragma unroll
            for (int j = 0; j < 4; j++) sum[j] = dp4a_us(w, 1 << (8 * j), sum[j]);
        }
        // valSum is 0 outside [10, H-10) x [10, W-10) (the zero-initialised Mat, CCalculation.cpp:801-823)
        bool cok[4];
#pragma unroll
        for (int j = 0; j < 4; j++) cok[j] = kColIn || ((unsigned)(c0 + j - kVhHalf) < (unsigned)(W - 2 * kVhHalf));
        const uint32_t kbase = kVhFlag | (unsigned)(4 * q);
#pragma unroll
        for (int r = 0; r < kVhSeg; r++) {
            const int h = hb + r;
            if (r > 0) {
                const uint32_t wa = ld(h + kVhHalf), ws = ld(h - kVhHalf - 1);
#pragma unroll
                for (int j = 0; j < 4; j++) sum[j] = dp4a_us(ws, 0xFF << (8 * j), dp4a_us(wa, 1 << (8 * j), sum[j]));
            }
            const bool row_ok = kRowIn || ((h >= kVhHalf) && (h < H - kVhHalf));
            uint32_t key[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                key[j] = ((row_ok && cok[j]) ? (uint32_t)sum[j] : 0u) * 512u + (kbase + (unsigned)j);
            uint2* dst = reinterpret_cast<uint2*>(&s_key[seg * kVhSeg + r][4 * q]);
            dst[0] = make_uint2(key[0], key[1]);
            dst[1] = make_uint2(key[2], key[3]);
        }
    }
    __syncthreads();

    const int g = t & 7;
    const int wbase = x0 + kVhOut * g;
    const int cb = kVhOut * g + 2;                        // tile column of window element j = 0  (= wbase - 10)
    const uint32_t K2 = kVhInv + 2u * (kVhFlag + (unsigned)cb);   // complemented key of element j:  K2 + 2j - key_j
    const uint32_t negc = (uint32_t)(-(cb + kVhHalf));    // -(tile column of output 0)
#pragma unroll 1
    for (int r = t >> 3; r < kVhH; r += 32) {
        const int h = y0 + r;
        if (!kInterior && (h >= H || wbase >= W)) break;
        const uint2* rowp = reinterpret_cast<const uint2*>(&s_key[r][cb]);
        uint32_t smn[kVhOut], smx[kVhOut];                // suffix minima of block 1 (j = 0..19)
        uint32_t cmn[kVhOut], cmx[kVhOut];                // centre candidates: element j = i + 10, bit 8 cleared
#pragma unroll
        for (int j2 = 0; j2 < kVhOut / 2; j2++) {
            const uint2 e = rowp[j2];
            smn[2 * j2] = e.x;
            smn[2 * j2 + 1] = e.y;
        }
#pragma unroll
        for (int j = 0; j < kVhOut; j++) smx[j] = K2 + 2u * (unsigned)j - smn[j];
#pragma unroll
        for (int j = kVhHalf; j < kVhOut; j++) {
            cmn[j - kVhHalf] = smn[j] - kVhFlag;
            cmx[j - kVhHalf] = smx[j] - kVhFlag;
        }
#pragma unroll
        for (int j = kVhOut - 2; j >= 0; j--) {
            smn[j] = min(smn[j], smn[j + 1]);
            smx[j] = min(smx[j], smx[j + 1]);
        }
        uint32_t e2[kVhOut];                              // block 2 (j = 20..39; j = 39 is never used)
#pragma unroll
        for (int j2 = 0; j2 < kVhOut / 2; j2++) {
            const uint2 e = rowp[kVhOut / 2 + j2];
            e2[2 * j2] = e.x;
            e2[2 * j2 + 1] = e.y;
        }
        uint32_t pmn = 0xFFFFFFFFu, pmx = 0xFFFFFFFFu;
        uint32_t bn[kVhOut], bx[kVhOut];                  // low byte = offset of the minimum / maximum
#pragma unroll
        for (int i = 0; i < kVhOut; i++) {
            if (i >= 1) {
                const int j = kVhOut - 1 + i;             // newest element of the window
                const uint32_t ej = e2[j - kVhOut];
                const uint32_t xj = K2 + 2u * (unsigned)j - ej;
                pmn = min(pmn, ej);
                pmx = min(pmx, xj);
                if (j < kVhOut + kVhHalf) {
                    cmn[j - kVhHalf] = ej - kVhFlag;
                    cmx[j - kVhHalf] = xj - kVhFlag;
                }
            }
            const uint32_t kn = min(min(smn[i], pmn), cmn[i]);
            const uint32_t kx = min(min(smx[i], pmx), cmx[i]);
            bn[i] = kn + negc - (unsigned)i;              // column field - column of output i
            bx[i] = kx + negc - (unsigned)i;
        }
        uint32_t packed[kVhOut / 2];                      // two pixels per word: (B, W, B, W)
#pragma unroll
        for (int m = 0; m < kVhOut / 2; m++) {
            const uint32_t nn = __byte_perm(bn[2 * m], bn[2 * m + 1], 0x0040);   // (B0, B1, ., .)
            const uint32_t xx = __byte_perm(bx[2 * m], bx[2 * m + 1], 0x0040);   // (W0, W1, ., .)
            packed[m] = __byte_perm(nn, xx, 0x5140);
        }
        char2* dst = out + (long long)h * W + wbase;
        const bool row_in = (h >= kVhHalf) && (h < H - kVhHalf);
        if (kInterior || (row_in && (wbase >= kVhHalf) && (wbase + kVhOut <= W - kVhHalf))) {
#pragma unroll
            for (int m = 0; m < kVhOut / 4; m++)
                reinterpret_cast<uint2*>(dst)[m] = make_uint2(packed[2 * m], packed[2 * m + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < kVhOut; i++) {
                const int w = wbase + i;
                if (w < W) {
                    const bool in = row_in && (w >= kVhHalf) && (w < W - kVhHalf);
                    const uint32_t pr = in ? ((packed[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu) : 0u;
                    dst[i] = make_char2((signed char)(pr & 0xFFu), (signed char)(pr >> 8));
                }
            }
        }
    }
}

template <int kVhH, int kVhSeg, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks)
strip_regression21_kernel(const uint8_t* __restrict__ frames, char2* __restrict__ strips, int W, int H)
{
    static_assert(kVhQuads * (kVhH / kVhSeg) <= 256 && kVhH % 32 == 0 && kVhH % kVhSeg == 0, "tile shape");
    __shared__ __align__(16) uint32_t s_key[kVhH][kVhStride];
    const long long npx = (long long)W * H;
    const uint8_t* img = frames + (long long)blockIdx.z * npx;
    char2* out = strips + (long long)blockIdx.z * npx;
    const int x0 = blockIdx.x * kVhW, y0 = blockIdx.y * kVhH;
    const bool col_in = (x0 >= kVhPad + kVhHalf) && (x0 + kVhW + kVhPad + kVhHalf <= W);
    const bool row_in = (y0 >= kVhHalf) && (y0 + kVhH + kVhHalf <= H);
    if (row_in && col_in) strip21_tile<true, true, kVhH, kVhSeg>(img, out, W, H, x0, y0, s_key);
    else if (row_in) strip21_tile<true, false, kVhH, kVhSeg>(img, out, W, H, x0, y0, s_key);
    else strip21_tile<false, false, kVhH, kVhSeg>(img, out, W, H, x0, y0, s_key);
}

// ---- FillOtherDeltaProU up to the blur's 3x3 sum (CCalculation.cpp:603-650) ----------------
// sums[f-1](h,w) = 576 + the sum over the 3x3 neighbourhood (BORDER_REFLECT_101) of the
// nearer-of-two delta between frames f-1 and f (each in [-19, 19]), stored as u16.  grid.z = f - 1.
constexpr int kDsBias = 64, kDsBias9 = 9 * kDsBias;

// generic path: any width / alignment
constexpr int kDgW = 128, kDgH = 8;

__global__ void __launch_bounds__(256)
delta_sum_generic_kernel(const char2* __restrict__ strips, unsigned short* __restrict__ sums, int W, int H)
{
    __shared__ short s_t[kDgH + 2][kDgW + 2];
    const long long npx = (long long)W * H;
    const char2* s0 = strips + (long long)blockIdx.z * npx;
    const char2* s1 = s0 + npx;
    const int x0 = blockIdx.x * kDgW, y0 = blockIdx.y * kDgH;
    for (int e = threadIdx.x; e < (kDgH + 2) * (kDgW + 2); e += 256) {
        const int ry = e / (kDgW + 2), rx = e % (kDgW + 2);
        const int y = reflect101(min(y0 + ry - 1, H), H), x = reflect101(min(x0 + rx - 1, W), W);
        const char2 a = s0[(long long)y * W + x], b = s1[(long long)y * W + x];
        const int dB = (int)a.x - (int)b.x;       // f0B - f1B   (CCalculation.cpp:603-617)
        const int dW = (int)a.y - (int)b.y;       // f0W - f1W
        s_t[ry][rx] = (short)((abs(dB) < abs(dW)) ? dB : dW);
    }
    __syncthreads();
    unsigned short* dst = sums + (long long)blockIdx.z * npx;
    for (int e = threadIdx.x; e < kDgH * kDgW; e += 256) {
        const int ry = e / kDgW, rx = e % kDgW;
        const int y = y0 + ry, x = x0 + rx;
        if (y >= H || x >= W) continue;
        int s = kDsBias9;
#pragma unroll
        for (int dy = 0; dy < 3; dy++)
#pragma unroll
            for (int dx = 0; dx < 3; dx++) s += s_t[ry + dy][rx + dx];
        dst[(long long)y * W + x] = (unsigned short)s;
    }
}

// byte / 16-bit-lane SIMD pieces of the fused kernel below (W % 8 == 0):
//  per 32-bit word (2 pixels x (B, W)):  D = a - b + 64 per byte, |d| by VABSDIFF4, the nearer of
//  (dB, dW) picked per 16-bit lane -> biased delta in [45, 83]; the 3x3 sum then works on those
//  lanes (<= 747, no carries).

__device__ __forceinline__ uint32_t nearer_delta_lanes(uint32_t a, uint32_t b)
{
    // a, b: (B0, W0, B1, W1) signed bytes of frame f-1 and frame f, each in [-10, 9]
    const uint32_t D = (a ^ 0x80808080u) + 0x40404040u - (b ^ 0x80808080u);   // d + 64 per byte, no borrows
    const uint32_t ab = __vabsdiffu4(D, 0x40404040u);                          // |d| per byte
    const uint32_t absB = ab & 0x00FF00FFu, absW = __byte_perm(ab, 0u, 0x4341);
    const uint32_t Z = absB + 0x00800080u - absW;        // bit 7 of a lane set <=> |dB| >= |dW| -> take dW
    uint32_t sel;
    asm("prmt.b32 %0, %1, %2, 0x4a48;" : "=r"(sel) : "r"(Z), "r"(0u));   // 0x00FF per lane whose bit 7 is set
    const uint32_t DB = D & 0x00FF00FFu, DW = __byte_perm(D, 0u, 0x4341);
    return DB ^ ((DB ^ DW) & sel);                       // (abs(dB) < abs(dW)) ? dB : dW   (CCalculation.cpp:603-617)
}

// z_exact + FOV test for a projector column held in f64
static __device__ __noinline__ float4 resolve_f64_u(const KParams& p, double U, int u, int v, int* valid_out)
{
    const double zd = z_exact(p, U, u, v);
    const bool valid = !((zd < p.fov_min) || (zd > p.fov_max));
    const float z = valid ? (float)zd : 0.f;
    *valid_out = valid ? 1 : 0;
    return make_float4(z * fmaf(p.rx1, (float)u, p.rx0), z * fmaf(p.ry1, (float)v, p.ry0), z, (float)U);
}

struct DynaOut {
    float4* xyzw;       // [n_frames-1][H][W]
    uint8_t* mask;      // [n_frames-1][H][W]
    float* delta_z;     // optional
    float* delta_p;     // optional parity: blurred deltaP
    double* proj_u;     // optional parity: U per frame
    double* u_final;    // optional: U after the last frame (state for a following call)
};

// The frame-to-frame recurrence U[f] = U[f-1] + deltaP[f] (CCalculation.cpp:656-658) is per
// pixel: one thread walks all frames of its pixel, then FillCoordinate (:672-775) per frame.
//
// The kernel is instruction-issue bound (with every store removed it ran only 14 % faster), so the
// frame loop is written for a short instruction stream: everything that depends on the pixel only
// (C, D, the x / y factors, the output pointers) is computed once, pointers advance by a constant
// stride per frame, deltaP comes from a shared-memory table addressed with one 32-bit add, the
// optional planes are template parameters, and the rare f64 re-solve is out of line.
constexpr int kLutN = 2 * 9 * 31 + 1;               // 3x3 sums of deltas; window <= 33: offsets in [-16, 15], deltas in [-31, 31]

__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

template <bool kDz, bool kParity>
__global__ void __launch_bounds__(256)
dyna_track_kernel(const __grid_constant__ KParams p, const unsigned short* __restrict__ sums, int n_frames,
                  const double* __restrict__ u0, const DynaOut o)
{
    // deltaP for every possible 3x3 sum: cv::blur on CV_32F is a double sum * (1./9) narrowed to
    // float (:650); kept as the double it is added to U as (:656-658)
    __shared__ double s_dp[kLutN];
    for (int i = threadIdx.x; i < kLutN; i += blockDim.x)
        s_dp[i] = (double)(float)__dmul_rn((double)(i - kLutN / 2), 1.0 / 9.0);
    __syncthreads();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.npx) return;
    int v, u;
    split_row_col(p, (unsigned)idx, v, u);
    const RowConst rc = make_row_const(p, v);
    const float uf = (float)u;
    // per-pixel constants of the f32 solve (triangulate_split evaluates the same expressions)
    const float C = fmaf(p.cu1, uf, rc.rowC), nD = -fmaf(p.du1, uf, rc.rowD);
    const float xr = fmaf(p.rx1, uf, p.rx0), yr = rc.ry;
    const float B32 = p.B32, nA32 = -p.A32, mid = p.fov_mid32, half = p.fov_half32, guard = p.guard_band,
                ng = p.num_guard, dg = p.den_guard;
    // table address of sum value s:  lut0 + 8*s  (s is stored with a bias of 576)
    const uint32_t lut0 = (uint32_t)__cvta_generic_to_shared(s_dp) - 8u * (uint32_t)(kDsBias9 - kLutN / 2);

    double U = u0[idx];
    // z of the frame before the first dynamic one: FillCoordinate(0) on U0 (CCalculation.cpp:189)
    float z_prev = 0.f;
    if (kDz && U != 0.0) {
        int ok;
        z_prev = resolve_f64_u(p, U, u, v, &ok).z;
    }
    const long long npx = p.npx;
    float4* px = o.xyzw + idx;
    uint8_t* pm = o.mask + idx;
    float* pz = kDz ? o.delta_z + idx : nullptr;
    float* pdp = (kParity && o.delta_p) ? o.delta_p + idx : nullptr;
    double* ppu = (kParity && o.proj_u) ? o.proj_u + idx : nullptr;
    const unsigned short* ps = sums + idx;

    // the 3x3 sums of the next group of frames are in flight while this group is processed
    constexpr int kAhead = 4;
    const int n_dyn = n_frames - 1;
    unsigned cur[kAhead], nxt[kAhead];
#pragma unroll
    for (int k = 0; k < kAhead; k++) cur[k] = (k < n_dyn) ? (unsigned)__ldcs(ps + (long long)k * npx) : (unsigned)kDsBias9;
    ps += (long long)kAhead * npx;
    for (int g = 0; g < n_dyn; g += kAhead) {
#pragma unroll
        for (int k = 0; k < kAhead; k++)
            nxt[k] = (g + kAhead + k < n_dyn) ? (unsigned)__ldcs(ps + (long long)k * npx) : (unsigned)kDsBias9;
        ps += (long long)kAhead * npx;
#pragma unroll
        for (int k = 0; k < kAhead; k++) {
            if (g + k >= n_dyn) break;
            const double dP = lds_f64(lut0 + 8u * cur[k]);
            U = __dadd_rn(U, dP);                                // :656-658
            // FillCoordinate (:672-771): f32 solve on U split exactly into two floats
            const float a = (float)U;
            const float b = (float)(U - (double)a);
            const float num = fmaf(B32, b, fmaf(B32, a, nA32));  // B*U - A
            const float den = fmaf(nD, b, fmaf(nD, a, C));       // C - D*U
            float rden;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rden) : "f"(den));
            float z = num * rden;
            const float dist = fabsf(z - mid) - half;            // <= 0 <=> inside the FOV
            const bool has_u = (U != 0.0);
            const bool need64 = (!(fabsf(dist) >= guard) || (fabsf(num) < ng) || (fabsf(den) < dg)) && has_u;
            bool ok = (dist <= 0.f) && has_u;
            z = ok ? z : 0.f;
            float4 outv = make_float4(z * xr, z * yr, z, a + b);
            if (need64) {                                        // undecidable in f32: the reference's f64 solve
                int ok64;
                outv = resolve_f64_u(p, U, u, v, &ok64);
                ok = ok64 != 0;
            }
            // write-once maps: streaming (evict-first) stores
            st_stream_f4(px, outv);
            asm volatile("st.global.cs.u8 [%0], %1;" :: "l"(pm), "r"(ok ? 1u : 0u) : "memory");
            px += npx;
            pm += npx;
            if (kDz) {                                           // :772-775
                asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(pz), "f"(outv.z - z_prev) : "memory");
                pz += npx;
                z_prev = outv.z;
            }
            if (kParity) {
                if (pdp) { *pdp = (float)dP; pdp += npx; }
                if (ppu) { *ppu = U; ppu += npx; }
            }
        }
#pragma unroll
        for (int k = 0; k < kAhead; k++) cur[k] = nxt[k];
    }
    if (o.u_final) o.u_final[idx] = U;
}


// ---- FillOtherDeltaProU + FillCoordinate fused (W % 8 == 0) --------------------------------
// One block owns a 128-column x (2*kPxT)-row tile and walks every frame of the sequence, so the
// 3x3 sums never leave the SM and the strips are read once per frame:
//   P  each thread keeps the strips of "its" four-pixel items of the tile + 1-pixel halo in
//      registers, has the next frame's in flight (one 8-byte load per item, issued a frame ahead)
//      and writes the nearer-of-two delta (+64, 16-bit lanes) to shared memory.  cv::blur's
//      BORDER_REFLECT_101 is folded into the item's source address (column -1 <- 1, W <- W-2, rows
//      alike).  Every cell of the region is written (dead ones once, with the bias): S adds whole
//      32-bit words, so even lanes it does not use must not be able to carry;
//   S  3x3 sums on 16-bit lanes: a thread owns 8 adjacent pixels; vertical = 3-input adds,
//      horizontal = two funnel shifts + one 3-input add per pixel pair;
//   T  the lean per-pixel loop of dyna_track_kernel; a thread owns kPxT pixels of one column, two
//      rows apart, so every store instruction of a warp covers 32 adjacent pixels.
// The tile height is picked on the host so that the grid fills whole waves of resident blocks
// (every block runs for the whole sequence, so a partly filled last wave costs a full one).
constexpr int kTfW = 128;
constexpr int kTfCols = kTfW + 8;                   // image columns x0-4 .. x0+131; smem index = x - x0 + 4
constexpr int kTfQuads = kTfCols / 4;               // 34

// (a single barrier per frame with both staging arrays double buffered was measured 16 % slower)
template <int kPxT, bool kDz, bool kParity>
__global__ void __launch_bounds__(256, kPxT <= 4 ? 3 : 2)
dyna_fused_kernel(const __grid_constant__ KParams p, const char2* __restrict__ strips, int n_frames,
                  const double* __restrict__ u0, const DynaOut o)
{
    constexpr int kTH = 2 * kPxT;                   // tile rows
    constexpr int kRows = kTH + 2;                  // + halo rows y0-1, y0+kTH
    constexpr int kItems = (kRows * kTfQuads + 255) / 256;
    __shared__ __align__(16) unsigned short s_d[kRows][kTfCols];
    __shared__ __align__(16) unsigned short s_sum[kTH][kTfW];
    __shared__ double s_dp[kLutN];
    const int W = p.W, H = p.H;
    const long long npx = p.npx;
    const int t = threadIdx.x;
    const int x0 = blockIdx.x * kTfW, y0 = blockIdx.y * kTH;

    for (int i = t; i < kLutN; i += 256)
        s_dp[i] = (double)(float)__dmul_rn((double)(i - kLutN / 2), 1.0 / 9.0);

    // ---- P items
    int src[kItems];              // pixel offset of the item's source quad inside a frame; -1 = dead
    int mode[kItems];             // 0 normal, 1 left mirror, 2 right mirror
    uint32_t dsts[kItems];        // shared-memory address of the item's four cells
    uint2 prev[kItems], cur[kItems];
#pragma unroll
    for (int it = 0; it < kItems; it++) {
        const int e = t + 256 * it;
        const bool cell = (e < kRows * kTfQuads);
        const int ry = cell ? e / kTfQuads : 0, qx = cell ? e - ry * kTfQuads : 0;
        int y = y0 + ry - 1, x = x0 - 4 + 4 * qx;
        bool live = cell;
        mode[it] = 0;
        if (y == -1) y = 1; else if (y == H) y = H - 2; else if (y > H) live = false;
        if (x == -4) { x = 0; mode[it] = 1; } else if (x == W) { x = W - 4; mode[it] = 2; } else if (x > W) live = false;
        src[it] = live ? y * W + x : -1;
        dsts[it] = (uint32_t)__cvta_generic_to_shared(&s_d[ry][4 * qx]);
        prev[it] = cur[it] = make_uint2(0u, 0u);
        if (live) {
            prev[it] = __ldg(reinterpret_cast<const uint2*>(strips + src[it]));
            if (n_frames > 1) cur[it] = __ldg(reinterpret_cast<const uint2*>(strips + npx + src[it]));
        } else if (cell) {
            *reinterpret_cast<uint2*>(&s_d[ry][4 * qx]) = make_uint2(0x00400040u, 0x00400040u);
        }
    }

    // ---- T pixels: column t & 127, rows (t >> 7) + 2j
    const int col = t & 127, row0 = t >> 7;
    const int u = x0 + col;
    const float uf = (float)u;
    const float xr = fmaf(p.rx1, uf, p.rx0);
    const float B32 = p.B32, nA32 = -p.A32, mid = p.fov_mid32, half = p.fov_half32, guard = p.guard_band,
                ng = p.num_guard, dg = p.den_guard;
    double U[kPxT];
    float zprev[kPxT], Cc[kPxT], nD[kPxT], yr[kPxT];
    bool in[kPxT];
#pragma unroll
    for (int j = 0; j < kPxT; j++) {
        const int v = y0 + row0 + 2 * j;
        in[j] = (v < H) && (u < W);
        const RowConst rc = make_row_const(p, v);
        Cc[j] = fmaf(p.cu1, uf, rc.rowC);
        nD[j] = -fmaf(p.du1, uf, rc.rowD);
        yr[j] = rc.ry;
        U[j] = in[j] ? u0[(long long)v * W + u] : 0.0;
        // z of the frame before the first dynamic one: FillCoordinate(0) on U0 (CCalculation.cpp:189)
        zprev[j] = 0.f;
        if (kDz && in[j] && U[j] != 0.0) {
            int ok;
            zprev[j] = resolve_f64_u(p, U[j], u, v, &ok).z;
        }
    }
    const long long pbase = (long long)(y0 + row0) * W + u;          // pixel j sits 2*j*W further
    float4* px = o.xyzw + pbase;
    uint8_t* pm = o.mask + pbase;
    float* pz = kDz ? o.delta_z + pbase : nullptr;
    float* pdp = (kParity && o.delta_p) ? o.delta_p + pbase : nullptr;
    double* ppu = (kParity && o.proj_u) ? o.proj_u + pbase : nullptr;
    const uint32_t lut0 = (uint32_t)__cvta_generic_to_shared(s_dp) - 8u * (uint32_t)(kDsBias9 - kLutN / 2);
    const uint32_t sum0 = (uint32_t)__cvta_generic_to_shared(&s_sum[row0][col]);
    const int sk = t & 15, sry = t >> 4;   // S role: 8 pixels from column 8*sk of tile row sry
    const char2* sp = strips + 2 * npx;    // frame f + 1 for the loop below

    for (int f = 1; f < n_frames; f++) {
        // ---- P
#pragma unroll
        for (int it = 0; it < kItems; it++) {
            if (src[it] >= 0) {
                const uint32_t rx = nearer_delta_lanes(prev[it].x, cur[it].x);
                const uint32_t ry = nearer_delta_lanes(prev[it].y, cur[it].y);
                const uint32_t w0 = (mode[it] == 2) ? ry : rx, w1 = (mode[it] == 1) ? rx : ry;
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(dsts[it]), "r"(w0), "r"(w1) : "memory");
                prev[it] = cur[it];
                if (f + 1 < n_frames) cur[it] = __ldg(reinterpret_cast<const uint2*>(sp + src[it]));
            }
        }
        sp += npx;
        __syncthreads();
        // ---- S
        if (sry < kTH) {
            uint32_t V[8];
#pragma unroll
            for (int rr = 0; rr < 3; rr++) {
                const uint4 lo = *reinterpret_cast<const uint4*>(&s_d[sry + rr][8 * sk]);        // columns c-4 .. c+3
                const uint4 hi = *reinterpret_cast<const uint4*>(&s_d[sry + rr][8 * sk + 8]);    // columns c+4 .. c+11
                const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                for (int i = 1; i < 7; i++) V[i] = (rr == 0) ? w[i] : V[i] + w[i];
            }
            uint32_t S[4];
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const uint32_t L = __funnelshift_l(V[m + 1], V[m + 2], 16);     // columns (c+2m-1, c+2m)
                const uint32_t R = __funnelshift_l(V[m + 2], V[m + 3], 16);     // columns (c+2m+1, c+2m+2)
                S[m] = L + V[m + 2] + R;
            }
            *reinterpret_cast<uint4*>(&s_sum[sry][8 * sk]) = make_uint4(S[0], S[1], S[2], S[3]);
        }
        __syncthreads();
        // ---- T
        // the shared-memory reads of all kPxT pixels first: independent chains for the scheduler
        uint32_t sb[kPxT];                                                // 576 + 3x3 sum
        double dPs[kPxT];
#pragma unroll
        for (int j = 0; j < kPxT; j++)
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(sb[j]) : "r"(sum0 + (uint32_t)(2 * j * kTfW * 2)));
#pragma unroll
        for (int j = 0; j < kPxT; j++) dPs[j] = lds_f64(lut0 + 8u * sb[j]);
#pragma unroll
        for (int j = 0; j < kPxT; j++) {
            if (!in[j]) continue;
            const double dP = dPs[j];
            U[j] = __dadd_rn(U[j], dP);                                   // :656-658
            const float a = (float)U[j];
            const float b = (float)(U[j] - (double)a);
            const float num = fmaf(B32, b, fmaf(B32, a, nA32));           // B*U - A
            const float den = fmaf(nD[j], b, fmaf(nD[j], a, Cc[j]));      // C - D*U
            float rden;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rden) : "f"(den));
            float z = num * rden;
            const float dist = fabsf(z - mid) - half;                     // <= 0 <=> inside the FOV
            const bool has_u = (U[j] != 0.0);
            const bool need64 = (!(fabsf(dist) >= guard) || (fabsf(num) < ng) || (fabsf(den) < dg)) && has_u;
            bool ok = (dist <= 0.f) && has_u;
            z = ok ? z : 0.f;
            float4 outv = make_float4(z * xr, z * yr[j], z, a + b);
            if (need64) {                                                 // undecidable in f32: the reference's f64 solve
                int ok64;
                outv = resolve_f64_u(p, U[j], u, y0 + row0 + 2 * j, &ok64);
                ok = ok64 != 0;
            }
            const long long oj = (long long)(2 * j) * W;
            st_stream_f4(px + oj, outv);
            asm volatile("st.global.cs.u8 [%0], %1;" :: "l"(pm + oj), "r"(ok ? 1u : 0u) : "memory");
            if (kDz) {                                                    // :772-775
                asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(pz + oj), "f"(outv.z - zprev[j]) : "memory");
                zprev[j] = outv.z;
            }
            if (kParity) {
                if (pdp) pdp[oj] = (float)dP;
                if (ppu) ppu[oj] = U[j];
            }
        }
        px += npx;
        pm += npx;
        if (kDz) pz += npx;
        if (kParity) { if (pdp) pdp += npx; if (ppu) ppu += npx; }
        // the next P writes s_d only after every thread has passed the second barrier (S is done),
        // the next S writes s_sum only after the next first barrier (T is done)
    }
    if (o.u_final) {
#pragma unroll
        for (int j = 0; j < kPxT; j++)
            if (in[j]) o.u_final[(long long)(y0 + row0 + 2 * j) * W + u] = U[j];
    }
}

}  // namespace

cudaError_t launch_strip_regression(const uint8_t* d_frames, int n_frames, int W, int H, int window,
                                    signed char* d_strips, cudaStream_t stream)
{
    const int half = window / 2;
    if (half < 1 || half > kSrMaxHalf || n_frames < 1 || n_frames > 65535) return cudaErrorInvalidValue;
    const bool aligned = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_frames) & 3) == 0) &&
                         ((reinterpret_cast<uintptr_t>(d_strips) & 7) == 0);
    if (half == kVhHalf && aligned) {
        // 64-row tiles, 16 rows per phase-1 thread, 3 blocks / SM: the fastest of the shapes tried
        // (32x8 rows: +5 %, 64x32: +17 %, 2 blocks / SM at 128 registers: +9 %)
        constexpr int kTileH = 64;
        dim3 grid((W + kVhW - 1) / kVhW, (H + kTileH - 1) / kTileH, n_frames);
        strip_regression21_kernel<kTileH, 16, 3><<<grid, 256, 0, stream>>>(d_frames, reinterpret_cast<char2*>(d_strips), W, H);
    } else {
        dim3 grid((W + kSrTileW - 1) / kSrTileW, (H + kSrTileH - 1) / kSrTileH, n_frames);
        strip_regression_kernel<<<grid, 256, 0, stream>>>(d_frames, reinterpret_cast<char2*>(d_strips), W, H, half);
    }
    return cudaGetLastError();
}

cudaError_t launch_delta_sum(const signed char* d_strips, int n_frames, int W, int H, unsigned short* d_sums,
                             cudaStream_t stream)
{
    if (n_frames < 2) return cudaSuccess;
    dim3 grid((W + kDgW - 1) / kDgW, (H + kDgH - 1) / kDgH, n_frames - 1);
    delta_sum_generic_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const char2*>(d_strips), d_sums, W, H);
    return cudaGetLastError();
}

cudaError_t launch_dyna_track(KParams p, const unsigned short* d_sums, int n_frames, const double* d_u0,
                              float* d_xyzw, uint8_t* d_mask, float* d_delta_z, float* d_delta_p,
                              double* d_proj_u, double* d_u_final, cudaStream_t stream)
{
    p.row_magic = ((unsigned long long)p.npx * (unsigned long long)p.W < (1ull << 40))
                      ? ((1ull << 40) / (unsigned long long)p.W + 1ull) : 0ull;
    DynaOut o{reinterpret_cast<float4*>(d_xyzw), d_mask, d_delta_z, d_delta_p, d_proj_u, d_u_final};
    // one pixel per thread (2 and 4 per thread measured slower)
    const unsigned blocks = (unsigned)((p.npx + 255) / 256);
    const bool parity = d_delta_p || d_proj_u;
    if (parity) {
        if (d_delta_z) dyna_track_kernel<true, true><<<blocks, 256, 0, stream>>>(p, d_sums, n_frames, d_u0, o);
        else dyna_track_kernel<false, true><<<blocks, 256, 0, stream>>>(p, d_sums, n_frames, d_u0, o);
    } else {
        if (d_delta_z) dyna_track_kernel<true, false><<<blocks, 256, 0, stream>>>(p, d_sums, n_frames, d_u0, o);
        else dyna_track_kernel<false, false><<<blocks, 256, 0, stream>>>(p, d_sums, n_frames, d_u0, o);
    }
    return cudaGetLastError();
}

bool dyna_fused_supported(int W, const signed char* d_strips)
{
    return (W % 8 == 0) && ((reinterpret_cast<uintptr_t>(d_strips) & 7) == 0);
}

namespace {
template <int kPxT>
cudaError_t launch_fused_rows(const KParams& p, const char2* st, int n_frames, const double* d_u0, const DynaOut& o,
                              bool dz, bool parity, cudaStream_t stream)
{
    dim3 grid((p.W + kTfW - 1) / kTfW, (p.H + 2 * kPxT - 1) / (2 * kPxT), 1);
    if (parity) {
        if (dz) dyna_fused_kernel<kPxT, true, true><<<grid, 256, 0, stream>>>(p, st, n_frames, d_u0, o);
        else dyna_fused_kernel<kPxT, false, true><<<grid, 256, 0, stream>>>(p, st, n_frames, d_u0, o);
    } else {
        if (dz) dyna_fused_kernel<kPxT, true, false><<<grid, 256, 0, stream>>>(p, st, n_frames, d_u0, o);
        else dyna_fused_kernel<kPxT, false, false><<<grid, 256, 0, stream>>>(p, st, n_frames, d_u0, o);
    }
    return cudaGetLastError();
}

// how full the last wave of resident blocks is for a grid of long-running blocks
template <int kPxT>
double fused_wave_fill(const KParams& p, int sm_count)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dyna_fused_kernel<kPxT, true, false>, 256, 0) != cudaSuccess ||
        per_sm < 1)
        return 0.0;
    const long long blocks = (long long)((p.W + kTfW - 1) / kTfW) * ((p.H + 2 * kPxT - 1) / (2 * kPxT));
    const long long slots = (long long)per_sm * sm_count;
    const long long waves = (blocks + slots - 1) / slots;
    return (double)blocks / (double)(waves * slots);
}
}  // namespace

cudaError_t launch_dyna_fused(KParams p, const signed char* d_strips, int n_frames, const double* d_u0,
                              float* d_xyzw, uint8_t* d_mask, float* d_delta_z, float* d_delta_p,
                              double* d_proj_u, double* d_u_final, int sm_count, cudaStream_t stream)
{
    DynaOut o{reinterpret_cast<float4*>(d_xyzw), d_mask, d_delta_z, d_delta_p, d_proj_u, d_u_final};
    const char2* st = reinterpret_cast<const char2*>(d_strips);
    const bool dz = d_delta_z != nullptr, parity = d_delta_p || d_proj_u;
    // 8-row tiles (4 pixels per thread) unless 4-row tiles fill the waves of resident blocks clearly
    // better.  Measured at 1280x1024 x 100 frames: 8 rows 570 us, 4 rows 660 us, 16 rows 1090 us
    // (2 blocks / SM at 118 registers and a last wave that is 16 % full).
    const bool small = fused_wave_fill<4>(p, sm_count) < 0.8 && fused_wave_fill<2>(p, sm_count) > fused_wave_fill<4>(p, sm_count) + 0.1;
    return small ? launch_fused_rows<2>(p, st, n_frames, d_u0, o, dz, parity, stream)
                 : launch_fused_rows<4>(p, st, n_frames, d_u0, o, dz, parity, stream);
}

}  // namespace slc
