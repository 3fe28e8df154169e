// slc_ingest.cu -- input ingest: what CSensor::LoadDatas obtains from
// imread(path, CV_LOAD_IMAGE_GRAYSCALE) for a .bmp (CSensorV.cpp:111-114), done on the device.
// The host only parses the 54-byte header + palette (slc_bmp_parse in slc_capi.cu) and uploads
// the raw file; this kernel undoes the bottom-up row order and the 4-byte row padding and maps
// palette indices / BGR(A) pixels to gray with OpenCV's fixed-point weights
//   gray = (B*1868 + G*9617 + R*4899 + 8192) >> 14        (highgui utils.cpp, SCALE = 14)
// writing straight into a plane of the plane-major u8 stack the fused kernel reads.
#include "slc_kernels.h"

namespace slc {

namespace {

struct BmpArgs {
    int width, height, bpp, top_down, row_stride, identity;
    uint8_t gray[256];
};

__device__ __forceinline__ uint32_t bgr_gray(uint32_t b, uint32_t g, uint32_t r)
{
    return (b * 1868u + g * 9617u + r * 4899u + 8192u) >> 14;
}

// one thread = 4 consecutive output pixels of a row
__global__ void __launch_bounds__(256)
bmp_unpack_kernel(const uint8_t* __restrict__ px, const __grid_constant__ BmpArgs a, uint8_t* __restrict__ plane)
{
    __shared__ uint8_t s_gray[256];
    if (a.bpp == 8 && !a.identity) {
        s_gray[threadIdx.x] = a.gray[threadIdx.x];
        __syncthreads();
    }
    const int quads = (a.width + 3) / 4;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)quads * a.height) return;
    const int y = (int)((unsigned)i / (unsigned)quads), x0 = (int)((unsigned)i - (unsigned)y * (unsigned)quads) * 4;   // i < W * H < 2^31
    const uint8_t* src = px + (long long)(a.top_down ? y : a.height - 1 - y) * a.row_stride;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = min(x0 + k, a.width - 1);
        if (a.bpp == 8) {
            const uint32_t idx = src[x];
            v[k] = a.identity ? idx : s_gray[idx];
        } else {
            const uint8_t* q = src + x * (a.bpp >> 3);
            v[k] = bgr_gray(q[0], q[1], q[2]);
        }
    }
    uint8_t* dst = plane + (long long)y * a.width + x0;
    if (x0 + 4 <= a.width && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0)) {
        *reinterpret_cast<uint32_t*>(dst) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
    } else {
        for (int k = 0; k < 4 && x0 + k < a.width; k++) dst[k] = (uint8_t)v[k];
    }
}


// ---- a whole set of files in one launch (grid.y = file) ------------------------------------
// 2.3 MB planes are a few microseconds of HBM time each, so one launch per plane is bound by launch
// latency; a frame set (2G+N files) unpacked by one grid runs at memory speed.  8-bit files whose
// rows can be written as aligned uint4 take the wide path: a thread owns 16 consecutive pixels, reads
// them as aligned vectors (the pixel array of a .bmp starts 1078 bytes into the file, so it is
// seldom better aligned than that), realigns with funnel shifts, maps through the palette if it is
// not the identity, and writes one uint4.
constexpr int kWideItems = 4;

struct BmpBatchArgs {
    BmpPlane plane[kBmpBatchMax];
};
static_assert(sizeof(BmpBatchArgs) <= 32764, "kernel parameters are limited to 32764 bytes (CUDA 12.1+, sm_70+)");

__device__ __forceinline__ uint32_t lut4(const uint8_t* g, uint32_t w)
{
    return (uint32_t)g[w & 0xFFu] | ((uint32_t)g[(w >> 8) & 0xFFu] << 8) | ((uint32_t)g[(w >> 16) & 0xFFu] << 16) |
           ((uint32_t)g[w >> 24] << 24);
}

__global__ void __launch_bounds__(256)
bmp_unpack_batch_kernel(const __grid_constant__ BmpBatchArgs b)
{
    __shared__ uint8_t s_gray[256];
    const BmpPlane& a = b.plane[blockIdx.y];
    const bool lut = (a.bpp == 8) && !a.identity;
    if (lut) {
        s_gray[threadIdx.x] = a.gray[threadIdx.x];
        __syncthreads();
    }
    if (a.wide) {
        // kWideItems 16-pixel items per thread, all loads first: one item per thread leaves too few
        // bytes in flight to cover the memory latency (measured 0.60 of the copy bandwidth)
        const int per_row = a.width >> 4;
        const long long n_items = (long long)per_row * a.height;
        const long long i0 = (long long)blockIdx.x * (256 * kWideItems) + threadIdx.x;
        const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(a.px) & 15u);
        // (row_stride and x0 are multiples of 16 here, so every item of a file has the misalignment of px)
        uint4 A[kWideItems], B[kWideItems];
        long long dsto[kWideItems];
#pragma unroll
        for (int k = 0; k < kWideItems; k++) {
            const long long i = i0 + 256 * k;
            dsto[k] = -1;
            A[k] = B[k] = make_uint4(0u, 0u, 0u, 0u);
            if (i < n_items) {
                // fewer than 2^27 items (W * H < 2^31): 32-bit division, not the 64-bit emulation
                const unsigned yu = (unsigned)i / (unsigned)per_row;
                const int y = (int)yu, x0 = (int)((unsigned)i - yu * (unsigned)per_row) << 4;
                const int srow = a.top_down ? y : a.height - 1 - y;
                const uint8_t* src = a.px + (long long)srow * a.row_stride + x0;
                dsto[k] = (long long)y * a.width + x0;
                const uint4* q = reinterpret_cast<const uint4*>(src - mis);
                A[k] = __ldg(q);
                if (mis != 0u) {
                    // the second vector reaches up to 15 bytes past the 16 this item needs: the very
                    // last item of the pixel array gathers its tail bytewise instead
                    const bool at_end = (srow == a.height - 1) && (x0 + 16 == a.width);
                    if (!at_end) {
                        B[k] = __ldg(q + 1);
                    } else {
                        uint32_t t4[4] = {0u, 0u, 0u, 0u};
                        for (unsigned m = 0; m < mis; m++) t4[m >> 2] |= (uint32_t)src[16 - mis + m] << (8 * (m & 3));
                        B[k] = make_uint4(t4[0], t4[1], t4[2], t4[3]);
                    }
                }
            }
        }
        const unsigned bs = 8u * (mis & 3u);
#pragma unroll
        for (int k = 0; k < kWideItems; k++) {
            if (dsto[k] < 0) continue;
            const uint32_t r[8] = {A[k].x, A[k].y, A[k].z, A[k].w, B[k].x, B[k].y, B[k].z, B[k].w};
            uint32_t w[4];
            switch (mis >> 2) {               // uniform per file
            case 0:
#pragma unroll
                for (int j = 0; j < 4; j++) w[j] = __funnelshift_r(r[j], r[j + 1], bs);
                break;
            case 1:
#pragma unroll
                for (int j = 0; j < 4; j++) w[j] = __funnelshift_r(r[j + 1], r[j + 2], bs);
                break;
            case 2:
#pragma unroll
                for (int j = 0; j < 4; j++) w[j] = __funnelshift_r(r[j + 2], r[j + 3], bs);
                break;
            default:
#pragma unroll
                for (int j = 0; j < 4; j++) w[j] = __funnelshift_r(r[j + 3], r[j + 4], bs);
                break;
            }
            if (lut) {
#pragma unroll
                for (int j = 0; j < 4; j++) w[j] = lut4(s_gray, w[j]);
            }
            st_stream_u4(a.out + dsto[k], make_uint4(w[0], w[1], w[2], w[3]));
        }
        return;
    }
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // any other flavour: 4 pixels per thread, as bmp_unpack_kernel
    const int quads = (a.width + 3) / 4;
    if (i >= (long long)quads * a.height) return;
    const int y = (int)((unsigned)i / (unsigned)quads), x0 = (int)((unsigned)i - (unsigned)y * (unsigned)quads) * 4;   // i < W * H < 2^31
    const uint8_t* src = a.px + (long long)(a.top_down ? y : a.height - 1 - y) * a.row_stride;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = min(x0 + k, a.width - 1);
        if (a.bpp == 8) {
            const uint32_t idx = src[x];
            v[k] = a.identity ? idx : s_gray[idx];
        } else {
            const uint8_t* q = src + x * (a.bpp >> 3);
            v[k] = bgr_gray(q[0], q[1], q[2]);
        }
    }
    uint8_t* dst = a.out + (long long)y * a.width + x0;
    if (x0 + 4 <= a.width && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0)) {
        *reinterpret_cast<uint32_t*>(dst) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
    } else {
        for (int k = 0; k < 4 && x0 + k < a.width; k++) dst[k] = (uint8_t)v[k];
    }
}

}  // namespace

cudaError_t launch_bmp_unpack(const uint8_t* d_pixels, int width, int height, int bpp, int top_down, int row_stride,
                              int identity, const uint8_t* gray256, uint8_t* d_plane, cudaStream_t stream)
{
    BmpArgs a;
    a.width = width; a.height = height; a.bpp = bpp; a.top_down = top_down; a.row_stride = row_stride;
    a.identity = identity;
    for (int i = 0; i < 256; i++) a.gray[i] = gray256[i];
    const long long n = (long long)((width + 3) / 4) * height;
    bmp_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_pixels, a, d_plane);
    return cudaGetLastError();
}

// planes[i].wide is set here; every plane of one launch shares the grid, sized for the largest.
cudaError_t launch_bmp_unpack_batch(const BmpPlane* planes, int n, cudaStream_t stream)
{
    for (int done = 0; done < n; done += kBmpBatchMax) {
        const int m = (n - done) < kBmpBatchMax ? (n - done) : kBmpBatchMax;
        BmpBatchArgs b;
        long long items = 1;
        for (int k = 0; k < m; k++) {
            BmpPlane& a = b.plane[k];
            a = planes[done + k];
            a.wide = (a.bpp == 8) && (a.width % 16 == 0) && (a.row_stride % 16 == 0) &&
                     ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0) ? 1 : 0;
            // blocks needed, in units of 256 threads: a wide thread takes kWideItems 16-pixel items
            const long long it = a.wide ? ((long long)(a.width / 16) * a.height + kWideItems - 1) / kWideItems
                                        : (long long)((a.width + 3) / 4) * a.height;
            if (it > items) items = it;
        }
        dim3 grid((unsigned)((items + 255) / 256), (unsigned)m, 1);
        bmp_unpack_batch_kernel<<<grid, 256, 0, stream>>>(b);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace slc
