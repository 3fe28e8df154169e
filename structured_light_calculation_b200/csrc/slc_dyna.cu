// slc_dyna.cu -- dynamic frames: the reference's CalculateOther path
// (CCalculation.cpp:208-320) as two kernels.
//
//   strip_regression_kernel  StripRegression (CCalculation.cpp:789-892): 21-row box sum per
//                            column, then the offset of the minimum / maximum of that sum
//                            over the 20 columns [w-10, w+9] -> (stripB, stripW) per pixel.
//                            Independent per frame: one launch covers every frame (grid.z).
//   dyna_track_kernel        FillOtherDeltaProU (CCalculation.cpp:595-663) + FillCoordinate(i>0)
//                            (CCalculation.cpp:666-775): nearer-of-two delta, 3x3 cv::blur
//                            (BORDER_REFLECT_101, double sum * (1./9) narrowed to f32),
//                            U[f] = U[f-1] + deltaP, triangulation, deltaZ.  The frame-to-frame
//                            recurrence is per pixel, so one thread walks all frames of its
//                            pixel: a single launch for the whole sequence.
//
// Everything up to U is integer / exactly-representable arithmetic (sums of <= 21 u8 values,
// index differences, multiples of 2^-27 well inside a double), so strips, deltaP and U are
// bit-exact against the CPU path whatever the summation order.
#include "slc_kernels.h"

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

__global__ void __launch_bounds__(256)
dyna_track_kernel(const __grid_constant__ KParams p, const char2* __restrict__ strips, int n_frames,
                  const double* __restrict__ u0, const DynaOut o)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.npx) return;
    int v, u;
    split_row_col(p, (unsigned)idx, v, u);
    const RowConst rc = make_row_const(p, v);
    const float uf = (float)u;

    // the 3x3 neighbourhood with cv::blur's default border (BORDER_REFLECT_101)
    int nb[9];
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++)
            nb[(dy + 1) * 3 + (dx + 1)] = reflect101(v + dy, p.H) * p.W + reflect101(u + dx, p.W);

    char2 prev[9];
#pragma unroll
    for (int k = 0; k < 9; k++) prev[k] = strips[nb[k]];

    double U = u0[idx];
    // z of the frame before the first dynamic one: FillCoordinate(0) on U0 (CCalculation.cpp:189)
    float z_prev;
    {
        int ok = 0;
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (U != 0.0) r0 = resolve_f64_u(p, U, u, v, &ok);
        z_prev = r0.z;
    }

    for (int f = 1; f < n_frames; f++) {
        const char2* cur_plane = strips + (long long)f * p.npx;
        int s = 0;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const char2 c = cur_plane[nb[k]];
            const int dB = (int)prev[k].x - (int)c.x;       // f0B - f1B   (CCalculation.cpp:603-617)
            const int dW = (int)prev[k].y - (int)c.y;       // f0W - f1W
            s += (abs(dB) < abs(dW)) ? dB : dW;
            prev[k] = c;
        }
        // cv::blur on CV_32F: double sum, * (1./9), narrowed to float (:650)
        const float dP = (float)__dmul_rn((double)s, 1.0 / 9.0);
        U = __dadd_rn(U, (double)dP);                        // :656-658
        // FillCoordinate (:672-771): f32 solve on U split exactly into two floats
        const float a = (float)U;
        const float b = (float)(U - (double)a);
        PixelResult r;
        triangulate_split<false>(p, rc, a, b, U != 0.0, uf, r);
        float4 outv = make_float4(r.x, r.y, r.z, r.w);
        int ok = r.valid ? 1 : 0;
        if (r.need64) outv = resolve_f64_u(p, U, u, v, &ok);
        const long long q = (long long)(f - 1) * p.npx + idx;
        o.xyzw[q] = outv;
        o.mask[q] = (uint8_t)ok;
        if (o.delta_z) o.delta_z[q] = outv.z - z_prev;       // :772-775
        if (o.delta_p) o.delta_p[q] = dP;
        if (o.proj_u) o.proj_u[q] = U;
        z_prev = outv.z;
    }
    if (o.u_final) o.u_final[idx] = U;
}

}  // namespace

cudaError_t launch_strip_regression(const uint8_t* d_frames, int n_frames, int W, int H, int window,
                                    signed char* d_strips, cudaStream_t stream)
{
    const int half = window / 2;
    if (half < 1 || half > kSrMaxHalf || n_frames < 1 || n_frames > 65535) return cudaErrorInvalidValue;
    dim3 grid((W + kSrTileW - 1) / kSrTileW, (H + kSrTileH - 1) / kSrTileH, n_frames);
    strip_regression_kernel<<<grid, 256, 0, stream>>>(d_frames, reinterpret_cast<char2*>(d_strips), W, H, half);
    return cudaGetLastError();
}

cudaError_t launch_dyna_track(KParams p, const signed char* d_strips, int n_frames, const double* d_u0,
                              float* d_xyzw, uint8_t* d_mask, float* d_delta_z, float* d_delta_p,
                              double* d_proj_u, double* d_u_final, cudaStream_t stream)
{
    p.row_magic = ((unsigned long long)p.npx * (unsigned long long)p.W < (1ull << 40))
                      ? ((1ull << 40) / (unsigned long long)p.W + 1ull) : 0ull;
    DynaOut o{reinterpret_cast<float4*>(d_xyzw), d_mask, d_delta_z, d_delta_p, d_proj_u, d_u_final};
    const long long blocks = (p.npx + 255) / 256;
    dyna_track_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p, reinterpret_cast<const char2*>(d_strips), n_frames,
                                                           d_u0, o);
    return cudaGetLastError();
}

}  // namespace slc
