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
    const int y = (int)(i / quads), x0 = (int)(i - (long long)y * quads) * 4;
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

}  // namespace slc
