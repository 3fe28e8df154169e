// slc_kernels.h -- internal launch interface between the C ABI (slc_capi.cu)
// and the kernels (slc_kernels.cu).  Not installed; the public surface is
// include/slcalc_b200.h.
#pragma once

#include "slc_device.cuh"

namespace slc {

struct LaunchInfo {
    int variant = -1;   // 0 specialised vector, 1 generic vector, 2 scalar
    int regs = 0;
    int block = 0;
    int smem = 0;
    int pxt = 0;        // pixels per thread
    bool query_only = false;
};

// Which kernel runs a (geometry, mode, output layout): chosen once per context (plan_reconstruct pays the
// table scan, the shared-memory attribute and the register query), used by every launch.
struct LaunchPlan {
    bool valid = false;
    int mode = 0;                 // 0 production, 1 + modulation test, 2 parity planes / custom table / Z_FP64
    int out = 0;                  // 0 xyzw + u8 mask, 1 depth + bit mask (SLC_RESULT_DEPTH)
    const void* vec = nullptr;    // the vector kernel, or nullptr when the geometry has none
    int pxt = 0, smem = 0, vec_regs = 0, scalar_regs = 0;
    bool specialised = false;
};
int plan_mode(const KParams& p);   // from the buffers / flags set in p
cudaError_t plan_reconstruct(const KParams& geom, int mode, int out, int pxt_override, LaunchPlan* plan);

// The fused decode -> unwrap -> triangulate kernel over p.n_stacks stacks (any count: split over
// launches of 65535).  p.depth != nullptr selects the SLC_RESULT_DEPTH layout.
cudaError_t launch_reconstruct(KParams p, const LaunchPlan& plan, bool force_scalar, cudaStream_t stream,
                               LaunchInfo* info);

cudaError_t launch_decode_gray(const KParams& p, const uint8_t* d_planes, double* d_gray_val,
                               int16_t* d_kbin, cudaStream_t stream);
cudaError_t launch_decode_phase(const KParams& p, const uint8_t* d_planes, double* d_phase_pix,
                                uint8_t* d_mod_ok, cudaStream_t stream);
cudaError_t launch_triangulate(const KParams& p, const double* d_proj_u, float* d_xyzw, uint8_t* d_mask,
                               cudaStream_t stream);

cudaError_t launch_triangulate_uv(const KParams& p, const double* d_proj_u, const double* d_proj_v, float* d_xyzw,
                                  uint8_t* d_mask, cudaStream_t stream);

cudaError_t launch_eval_phase(const float* d_s, const float* d_c, long long n, float Tf, float* d_deg, float* d_pix,
                              cudaStream_t stream);

// dynamic frames (slc_dyna.cu)
cudaError_t launch_strip_regression(const uint8_t* d_frames, int n_frames, int W, int H, int window,
                                    signed char* d_strips, cudaStream_t stream);
cudaError_t launch_delta_sum(const signed char* d_strips, int n_frames, int W, int H, unsigned short* d_sums,
                             cudaStream_t stream);
cudaError_t launch_dyna_track(KParams p, const unsigned short* d_sums, int n_frames, const double* d_u0,
                              float* d_xyzw, uint8_t* d_mask, float* d_delta_z, float* d_delta_p,
                              double* d_proj_u, double* d_u_final, cudaStream_t stream);

// W % 8 == 0: delta + 3x3 sum + track in one kernel straight from the strips (no sums plane)
bool dyna_fused_supported(int W, const signed char* d_strips);
cudaError_t launch_dyna_fused(KParams p, const signed char* d_strips, int n_frames, const double* d_u0,
                              float* d_xyzw, uint8_t* d_mask, float* d_delta_z, float* d_delta_p,
                              double* d_proj_u, double* d_u_final, int sm_count, cudaStream_t stream);

// point-cloud output (slc_pointcloud.cu)
size_t pointcloud_scratch_bytes(long long npx);
// d_scratch: pointcloud_scratch_bytes() bytes, ZEROED when allocated; text_epoch: 1, 2, 3, 1, ... advancing with every
// mode-0 launch on this scratch (the look-back words of the text kernel carry it)
cudaError_t launch_pointcloud(const KParams& p, int mode, int order, unsigned flags, const double* d_proj_u,
                              const float* d_xyzw, const uint8_t* d_mask, void* d_out, unsigned long long capacity,
                              void* d_scratch, unsigned text_epoch, const unsigned long long** d_totals, cudaStream_t stream);
cudaError_t launch_format_g6(const double* d_values, long long n, unsigned flags, char* d_text, uint8_t* d_len,
                             cudaStream_t stream);

// valid-only point lists, one launch per batch (slc_compact.cu)
bool compact_supported(int W, int H, const void* d_mask, const void* d_xyzw);
int compact_tiles(int W, int H, int order);
size_t compact_state_bytes(int W, int H, int n_stacks);
cudaError_t launch_compact(int W, int H, int n_stacks, int order, const float* d_xyzw, const uint8_t* d_mask,
                           float* d_points, long long point_stride, uint8_t* d_mask_bits, long long bits_stride,
                           unsigned long long* d_counts, unsigned long long* d_state, unsigned epoch,
                           cudaStream_t stream);

// z plane + validity bits of finished (xyzw, mask) maps: the SLC_RESULT_DEPTH layout for the dynamic frames
cudaError_t launch_pack_depth(const float* d_xyzw, const uint8_t* d_mask, long long npx, int n_maps, float* d_depth,
                              uint8_t* d_bits, long long bits_stride, cudaStream_t stream);

// input ingest (slc_ingest.cu)
cudaError_t launch_bmp_unpack(const uint8_t* d_pixels, int width, int height, int bpp, int top_down, int row_stride,
                              int identity, const uint8_t* gray256, uint8_t* d_plane, cudaStream_t stream);
// one file of a batched unpack: pixel array in, one [height][width] u8 plane out
struct BmpPlane {
    const uint8_t* px;
    uint8_t* out;
    int width, height, bpp, top_down, row_stride, identity;
    int wide, pad_;               // set by launch_bmp_unpack_batch
    uint8_t gray[256];
};
constexpr int kBmpBatchMax = 64;  // files per launch (the descriptors travel as kernel parameters: 64 x 304 B < 32 KB)
cudaError_t launch_bmp_unpack_batch(const BmpPlane* planes, int n, cudaStream_t stream);

}  // namespace slc
