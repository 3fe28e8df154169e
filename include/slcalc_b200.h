/*
 * slcalc_b200.h -- C ABI of libslcalc_b200.so, the B200 (sm_100a) drop-in for
 * DynaFrame's first-frame structured-light reconstruction path.
 *
 * The reference (elevenface/Structured-Light-Calculation) has no FFI layer:
 * its seam is the C++ class API CDecodeGray / CDecodePhase / CCalculation.
 * Each entry point below names the reference interface it replaces; paths are
 * relative to DynaFrame/DynaFrame/ in the reference tree.  The C++ classes with
 * the reference's own names live in include/dynaframe_b200.hpp and are built
 * on nothing but this header.
 *
 * Conventions
 *   - plain pointers and sizes only; no CUDA, torch or OpenCV types;
 *   - every function returns an slc_status (0 = OK); the message for the last
 *     failure on a context is slc_last_error(ctx) (slc_last_error(NULL) for a
 *     failed slc_create).  The reference's bool + ErrorHandling(string)
 *     convention (GlobalFunction.cpp:3-8) maps to status != SLC_OK + message;
 *   - a context is bound to one CUDA device and is single-caller; distinct
 *     contexts may be driven from distinct host threads (one per GPU) -- there is
 *     no process-wide mutable state.  Calls on ONE context must be stream-ordered:
 *     the context owns scratch buffers (look-back state, strips, staging) that an
 *     asynchronous _device call is still using until the work it queued has
 *     finished, so a second _device call on the same context must go to the same
 *     stream or wait for the first.  slc_pool runs several contexts (one feeder
 *     thread each) behind one call;
 *   - there is NO CPU fallback: without a CUDA device slc_create fails with
 *     SLC_ERR_NO_DEVICE.
 *
 * Data layout
 *   stack   : uint8 [n_stacks][P][H][W], P = 2*G + N, plane-major.  Planes
 *             2b, 2b+1 are the Gray pattern / inverse pair of bit b, b = 0 the
 *             LSB (CDecodeGray.cpp:159,193-198; the order CCalculation.cpp
 *             :539-544 feeds SetMat), followed by the N phase images
 *             (CDecodePhase.cpp:59-62; CCalculation.cpp:552-557).
 *   xyzw    : float  [n_stacks][H][W][4] = (x, y, z, U) -- m_xMat/m_yMat/m_zMat
 *             (CCalculation.cpp:118-120) packed, w = decoded projector column
 *             (m_ProjectorU, :115) rounded ONCE to f32 (0 where the [EXT] modulation
 *             test rejects the pixel).  Invalid pixel => x=y=z=0.
 *             Precision of w as an unwrapped phase 2*pi*U/T: half an f32 ulp of U
 *             times 2*pi/T, i.e. <= 1e-4 rad (the north-star bar) for every column
 *             inside the projector raster iff  ulp_f32(PW - 1)/2 * 2*pi/T <= 1e-4:
 *             PW <= 4096 with T >= 8, PW <= 2048 with T >= 4, PW <= 1024 with
 *             T >= 2 -- all BASELINE geometries (worst: 4096 / T = 8, 9.6e-5 rad).
 *             Outside that range (e.g. PW 4096 with T = 4: 1.9e-4 rad; PW 65536 with
 *             T = 2: 6.1e-3 rad) f32 cannot hold the bar; the exact value is the
 *             proj_u parity plane (f64, bit-exact for every geometry).
 *   mask    : uint8  [n_stacks][H][W], 1 = modulation_ok && U != 0 &&
 *             fov_min <= z <= fov_max (CCalculation.cpp:678-682,701-704).
 *   parity planes (optional): kbin int16, corr int8, phase_pix float,
 *             proj_u double, each [n_stacks][H][W].
 */
#ifndef SLCALC_B200_H_
#define SLCALC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLC_ABI_VERSION 2

typedef enum {
    SLC_OK = 0,
    SLC_ERR_INVALID_ARG = 1,     /* bad geometry, NULL pointer, digit count outside 1..16 (CDecodeGray.cpp:39) */
    SLC_ERR_NOT_INITIALISED = 2, /* calibration missing (CCalculation.cpp:176-181), SetMat before SetNum* */
    SLC_ERR_CUDA = 3,            /* a CUDA runtime call or kernel launch failed */
    SLC_ERR_NO_DEVICE = 4,       /* no usable CUDA device: there is no CPU path */
    SLC_ERR_OUT_OF_MEMORY = 5,
    SLC_ERR_STATE = 6            /* e.g. slot busy, Init twice (CCalculation.cpp:80-83) */
} slc_status;

/* slc_config.flags */
#define SLC_FLAG_Z_FP64        (1u << 0) /* solve every z in f64 (validation mode); default is f32 with an
                                            f64 re-solve of pixels near the FOV limits, which already makes
                                            the mask bit-exact */
#define SLC_FLAG_SCALAR_KERNEL (1u << 1) /* force the one-pixel-per-thread kernel (used for widths that are
                                            not a multiple of 16, and as a cross-check) */

typedef struct slc_context slc_context;

/* The reference's compile-time constants (StaticParameters.cpp:4-9,16-18,
 * 34-35) as a runtime struct. */
typedef struct {
    int32_t  width;            /* CAMERA_RESLINE */
    int32_t  height;           /* CAMERA_RESROW */
    int32_t  projector_width;  /* PROJECTOR_RESLINE */
    int32_t  gray_digits;      /* GRAY_V_NUMDIGIT: pattern/inverse pairs, 1..16 */
    int32_t  phase_steps;      /* PHASE_NUMDIGIT: N >= 3 (reference: 4) */
    double   fov_min;          /* FOV_MIN_DISTANCE */
    double   fov_max;          /* FOV_MAX_DISTANCE */
    float    modulation_min;   /* [EXT] minimum fringe amplitude b in grey levels; 0 = off (reference) */
    uint32_t flags;            /* SLC_FLAG_* */
    int32_t  device;           /* CUDA device ordinal */
    int32_t  max_batch;        /* stacks per slot the host path can stage on the device (>= 1) */
    int32_t  num_slots;        /* stream slots for the pipelined host path, 1..8 */
} slc_config;

typedef struct {
    int32_t  planes;           /* P = 2G + N */
    int32_t  gray_period;      /* gp = PW / 2^G      (CDecodeGray.cpp:183) */
    int32_t  phase_period;     /* T  = PW / 2^(G-1)  (CCalculation.cpp:550) */
    int32_t  sm_count;
    int64_t  pixels;           /* W*H */
    int64_t  stack_bytes;      /* P*W*H */
    int64_t  xyzw_bytes;       /* 16*W*H */
    int64_t  mask_bytes;       /* W*H */
    int32_t  kernel_variant;   /* 0 = specialised <G,N> vector kernel, 1 = generic vector, 2 = scalar */
    int32_t  kernel_regs;      /* registers per thread of the kernel that will run */
    int32_t  kernel_block;     /* threads per block */
    int32_t  kernel_smem;      /* dynamic shared memory per block, bytes */
} slc_info;

/* Optional parity / debug planes.  NULL members are skipped. */
typedef struct {
    int16_t *kbin;       /* gray2bin[code]: half-period index (CDecodeGray.cpp:200 before the multiply) */
    int8_t  *corr;       /* wrap correction taken at CCalculation.cpp:572-581: -1, 0, +1 */
    float   *phase_pix;  /* CDecodePhase::CountResult value, (0, T] (CDecodePhase.cpp:69-75) */
    double  *proj_u;     /* ProjectorU, exact f64 (CCalculation.cpp:587-589) */
} slc_parity_planes;

/* ---- life cycle ------------------------------------------------------- */
/* replaces: CCalculation::Init allocation block (CCalculation.cpp:95-121) +
 * CDecodeGray::SetNumDigit (CDecodeGray.cpp:36-53) + CDecodePhase::SetNumMat
 * (CDecodePhase.cpp:119-137) */
int slc_create(const slc_config *cfg, slc_context **out);
/* replaces: CCalculation::ReleaseSpace (CCalculation.cpp:26-75) */
void slc_destroy(slc_context *ctx);
const char *slc_last_error(const slc_context *ctx);
const char *slc_status_string(int status);
int slc_abi_version(void);
int slc_get_info(const slc_context *ctx, slc_info *out);

/* replaces: the FileStorage read + C/P/A/B/cC/cD set-up of CCalculation::Init
 * (CCalculation.cpp:124-166).  Row-major f64: CamMat 3x3, ProMat 3x3, R 3x3,
 * T 3x1.  The per-pixel cC/cD LUT planes of the reference are not
 * materialised: the kernel evaluates them from (u, v). */
int slc_set_calibration(slc_context *ctx, const double cam[9], const double pro[9],
                        const double R[9], const double T[3]);

/* replaces: the gray2bin table CDecodeGray::Decode loads from
 * Patterns/vGrayCode.txt (CDecodeGray.cpp:113-125): lut[grayCode] = binCode,
 * n = 2^G entries.  Optional -- the default is the reflected binary code the
 * shipped file holds, decoded arithmetically. */
int slc_set_gray_lut(slc_context *ctx, const int16_t *gray2bin, int32_t n);

/* ---- memory helpers --------------------------------------------------- */
void *slc_host_alloc(size_t bytes);            /* pinned host memory */
#define SLC_HOST_WRITE_COMBINED (1u << 0)      /* upload-only staging: fast for the CPU to fill, slow to read back */
#define SLC_HOST_HUGE_PAGES     (1u << 1)      /* 2 MB pages (hugetlb pool if the box has one, else transparent) */
void *slc_host_alloc_ex(size_t bytes, uint32_t flags);   /* free with slc_host_free */
void slc_host_free(void *p);
int slc_host_register(void *p, size_t bytes);  /* pin caller-owned memory */
int slc_host_unregister(void *p);
void *slc_device_alloc(slc_context *ctx, size_t bytes);
void slc_device_free(slc_context *ctx, void *p);
int slc_copy_to_device(slc_context *ctx, void *dst, const void *src, size_t bytes);
int slc_copy_to_host(slc_context *ctx, void *dst, const void *src, size_t bytes);
int slc_synchronize(slc_context *ctx);

/* ---- the hot path ----------------------------------------------------- */
/* replaces: CCalculation::CalculateFirst's compute, i.e. FillFirstProjectorU
 * (CCalculation.cpp:525-592, with CDecodeGray::Decode CDecodeGray.cpp:108-204
 * and CDecodePhase::Decode CDecodePhase.cpp:48-96 inside) + FillCoordinate(0)
 * (CCalculation.cpp:666-771), for n_stacks independent frame sets, as ONE
 * fused kernel launch.  All pointers are DEVICE pointers; cuda_stream is a
 * cudaStream_t passed as void* (NULL = the context's own stream; the legacy default stream, whose handle is 0, cannot be named).
 * Asynchronous with respect to the host.  n_stacks == 0 is a no-op (SLC_OK, no launch, buffers
 * may be NULL), here and in slc_reconstruct_host; n_stacks < 0 is SLC_ERR_INVALID_ARG. */
int slc_reconstruct_device(slc_context *ctx, const uint8_t *d_stack, int32_t n_stacks,
                           float *d_xyzw, uint8_t *d_mask,
                           const slc_parity_planes *d_parity, void *cuda_stream);

/* Same computation with HOST buffers: upload, kernel, download, synchronous.
 * n_stacks may exceed max_batch; the call then pipelines chunks over the
 * context's stream slots (copy-in / compute / copy-out overlapped).  Host
 * buffers should be pinned (slc_host_alloc / slc_host_register) for the
 * copies to overlap. */
int slc_reconstruct_host(slc_context *ctx, const uint8_t *h_stack, int32_t n_stacks,
                         float *h_xyzw, uint8_t *h_mask,
                         const slc_parity_planes *h_parity);

/* Explicit double-buffering: start upload+kernel+download of up to max_batch
 * stacks on a slot, return immediately; slc_wait blocks until that slot's
 * results are in the host buffers. */
int slc_submit_host(slc_context *ctx, int32_t slot, const uint8_t *h_stack, int32_t n_stacks,
                    float *h_xyzw, uint8_t *h_mask);
int slc_wait(slc_context *ctx, int32_t slot);

/* ---- result formats ---------------------------------------------------- */
/* What a reconstruct call hands back.  The reference's only consumer of the maps,
 * CCalculation::Result (CCalculation.cpp:336-346), keeps the pixels with FOV_MIN <= z <= FOV_MAX and
 * drops the rest, so the full float4 map + byte mask (17 B/px) is more than anything downstream reads;
 * on the host path the result bytes cross PCIe, which is what bounds it.  Every format is defined as
 * a selection of the SLC_RESULT_XYZW output, bit for bit:
 *   SLC_RESULT_XYZW    xyzw [n][H][W][4] + mask [n][H][W]                         17     B/px
 *   SLC_RESULT_DEPTH   depth [n][H][W] = z (0 where invalid) + mask_bits            4.125 B/px
 *                      (x = z*(u-cu)/fu, y = z*(v-cv)/fv: CCalculation.cpp:766-767)
 *   SLC_RESULT_POINTS  points [n][point_stride][3] = (x, y, z) of the valid pixels only, in `order`
 *                      (SLC_ORDER_ROW_MAJOR, or SLC_ORDER_REFERENCE = the order Result() walks),
 *                      n_points [n], + mask_bits                                    12 B/valid px + 0.125 B/px
 * mask_bits: u8 [n][(H*W+7)/8], bit (i & 7) of byte (i >> 3) = pixel i valid (numpy packbits,
 * bitorder "little"); 4-byte aligned, total size padded to a multiple of 4 bytes.
 * SLC_RESULT_POINTS needs a camera width that is a multiple of 8 and a height of at most 12000 (the
 * chained-scan kernels of slc_compact.cu); the other formats take any geometry. */
#define SLC_RESULT_XYZW   0
#define SLC_RESULT_DEPTH  1
#define SLC_RESULT_POINTS 2
#define SLC_ORDER_ROW_MAJOR 0   /* v outer, u inner: the memory order of the maps */
#define SLC_ORDER_REFERENCE 1   /* u outer, v inner: the order Result() walks (CCalculation.cpp:336-338) */

typedef struct {
    int32_t  format;        /* SLC_RESULT_* */
    int32_t  order;         /* POINTS: SLC_ORDER_* */
    float   *xyzw;          /* XYZW: required.  POINTS: optional on the host path, required on the device path
                               (the maps the points are taken from are written there) */
    uint8_t *mask;          /* as xyzw */
    float   *depth;         /* DEPTH */
    uint8_t *mask_bits;     /* DEPTH: required.  POINTS: optional */
    float   *points;        /* POINTS */
    int64_t  point_stride;  /* POINTS: points reserved per frame set; a frame set with more valid pixels keeps the
                               first point_stride of them, n_points still reports them all, and the call returns
                               SLC_ERR_INVALID_ARG after writing everything that fits */
    int64_t *n_points;      /* POINTS: [n] */
} slc_result;

/* slc_reconstruct_device / _host with a result format.  XYZW is exactly the plain call.  DEPTH is
 * written by the fused kernel itself (26.1 instead of 39 B/px of HBM traffic at 1920x1200 G9 N4);
 * POINTS adds ONE chained-scan launch per call (slc_compact.cu).  On the host path only the selected
 * bytes are downloaded; POINTS downloads exactly 12 * n_points[i] bytes per frame set. */
int slc_reconstruct_device_ex(slc_context *ctx, const uint8_t *d_stack, int32_t n_stacks,
                              const slc_result *d_out, void *cuda_stream);
int slc_reconstruct_host_ex(slc_context *ctx, const uint8_t *h_stack, int32_t n_stacks,
                            const slc_result *h_out);
/* slc_submit_host with a result format (up to max_batch frame sets on a free slot; slc_wait completes it,
 * and for SLC_RESULT_POINTS is also where the point lists are fetched, once their counts are known). */
int slc_submit_host_ex(slc_context *ctx, int32_t slot, const uint8_t *h_stack, int32_t n_stacks,
                       const slc_result *h_out);

/* ---- several GPUs behind one call --------------------------- */
/* replaces: the frame loop of CCalculation::CalculateOther (CCalculation.cpp:221) seen as independent
 * frame sets, i.e. what main.cpp:42-45 would shard.  A pool owns one context per listed device (a device
 * may be listed more than once) and one feeder thread per context; calibration and Gray table are
 * replicated.  slc_pool_reconstruct_host hands frame sets out ON DEMAND, max_batch at a time: each feeder
 * keeps the stream slots of its context full (slc_submit_host_ex / slc_wait) and takes the next chunk as
 * soon as a slot frees up, so a GPU behind a faster host link takes more of the batch (on the pool's
 * 8-GPU hosts four of the GPUs share an uplink and move 2/3 of what the other four do,
 * profiles/r02_hostlink_probe.txt) and the pipeline never drains between chunks.  Results land at the index
 * of their frame set whoever computed it.  slc_pool_reconstruct_device takes caller-made shards
 * (slc_shard_range gives the contiguous equal split).  There is no data-path collective.  cfg->device is
 * ignored. */
typedef struct slc_pool slc_pool;
int slc_pool_create(const slc_config *cfg, const int32_t *devices, int32_t n_devices, slc_pool **out);
void slc_pool_destroy(slc_pool *pool);
const char *slc_pool_last_error(const slc_pool *pool);   /* slc_pool_last_error(NULL): a failed slc_pool_create */
int slc_pool_size(const slc_pool *pool);
slc_context *slc_pool_context(slc_pool *pool, int32_t member);   /* for per-GPU calls (device_alloc, info ...) */
int slc_pool_set_calibration(slc_pool *pool, const double cam[9], const double pro[9],
                             const double R[9], const double T[3]);
int slc_pool_set_gray_lut(slc_pool *pool, const int16_t *gray2bin, int32_t n);
/* contiguous block [*lo, *hi) of n_items for shard `index` of `n_shards`; sizes differ by at most one */
int slc_shard_range(int64_t n_items, int32_t index, int32_t n_shards, int64_t *lo, int64_t *hi);
/* n_stacks frame sets in host memory -> results in host memory, all members at once; blocks.
 * A failure on one member (including a SLC_RESULT_POINTS frame set that does not fit point_stride) ends the
 * call early: the other members finish the chunks they hold and take no more, the first failing member's
 * status and message are returned.  slc_pool_last_shares(): how many frame sets each member took in the
 * last call. */
int slc_pool_reconstruct_host(slc_pool *pool, const uint8_t *h_stack, int32_t n_stacks,
                              const slc_result *h_out);
int slc_pool_last_shares(const slc_pool *pool, int32_t *frame_sets_per_member, int32_t n_members);
/* Device-resident shards: member i runs n_stacks[i] frame sets from d_stack[i] into d_out[i] (device
 * pointers on that member's GPU) and the call returns when every GPU has finished. */
int slc_pool_reconstruct_device(slc_pool *pool, const uint8_t *const *d_stack, const int32_t *n_stacks,
                                const slc_result *d_out);

/* ---- the decoder objects on their own --------------------------------- */
/* replaces: CDecodeGray::Decode + GetResult (CDecodeGray.cpp:108-147): 2G
 * planes -> CV_64FC1 plane of left-edge projector columns; kbin optional. */
int slc_decode_gray_host(slc_context *ctx, const uint8_t *h_gray_planes,
                         double *h_gray_val, int16_t *h_kbin);
/* replaces: CDecodePhase::Decode + GetResult (CDecodePhase.cpp:83-104): N
 * planes -> CV_64FC1 plane of in-period offsets in (0, T]; mod_ok optional. */
int slc_decode_phase_host(slc_context *ctx, const uint8_t *h_phase_planes,
                          double *h_phase_pix, uint8_t *h_mod_ok);
/* replaces: CCalculation::FillCoordinate(i) for an arbitrary ProjectorU plane
 * (CCalculation.cpp:666-771), the step the dynamic-frame mode re-uses. */
int slc_triangulate_host(slc_context *ctx, const double *h_proj_u,
                         float *h_xyzw, uint8_t *h_mask);

/* [EXT] both projector coordinates: ProjectorU decoded from the vertical patterns and
 * ProjectorV from horizontal ones (decode the second stack with a context whose projector_width is the
 * projector HEIGHT and whose gray_digits is GRAY_H_NUMDIGIT -- what CDecodeGray::SetNumDigit(n, false)
 * selects, CDecodeGray.cpp:182-185 -- and ask for its proj_u parity plane).  Each coordinate gives one
 * linear equation in z built from its row of P the way CCalculation.cpp:159-164,686-687 builds the
 * column one; z is their least-squares solution, in f64.  mask = U != 0 && V != 0 && fov_min <= z <=
 * fov_max.  The reference itself never uses row 1 of P: there is no reference output to match, the
 * oracle's restatement of this definition is matched bit for bit. */
int slc_triangulate_uv_device(slc_context *ctx, const double *d_proj_u, const double *d_proj_v,
                              float *d_xyzw, uint8_t *d_mask, void *cuda_stream);
int slc_triangulate_uv_host(slc_context *ctx, const double *h_proj_u, const double *h_proj_v,
                            float *h_xyzw, uint8_t *h_mask);

/* ---- dynamic frames ("next" row: CCalculation::CalculateOther) ----------- */
/* Optional parity planes of the dynamic path.  NULL members are skipped. */
typedef struct {
    int8_t *strips;      /* [n_frames][H][W][2] = (stripB, stripW) offsets, StripRegression (CCalculation.cpp:886-887) */
    float  *delta_p;     /* [n_frames-1][H][W] blurred deltaP (CCalculation.cpp:650) */
    double *proj_u;      /* [n_frames-1][H][W] ProjectorU[f] (CCalculation.cpp:656-658) */
} slc_dyna_parity;

/* replaces: the per-frame body of CCalculation::CalculateOther (CCalculation.cpp:221-243):
 * StripRegression(f) (:789-892), FillOtherDeltaProU(f) (:595-663) and FillCoordinate(f)
 * (:666-775) for f = 1 .. n_frames-1.  frames is uint8 [n_frames][H][W] (the dynaCam images;
 * frames[0] is the one StripRegression(0) sees at CCalculation.cpp:201), window is
 * RECO_WINDOW_SIZE (StaticParameters.cpp:38, odd, <= 33), u0 is ProjectorU[0] (f64 [H][W], e.g.
 * the proj_u parity plane of the first-frame call).  Outputs hold n_frames-1 maps: xyzw, mask,
 * and optionally deltaZ (m_deltaZ, :772-775).  Two kernel launches for the whole sequence. */
int slc_dyna_track_device(slc_context *ctx, const uint8_t *d_frames, int32_t n_frames, int32_t window,
                          const double *d_u0, float *d_xyzw, uint8_t *d_mask, float *d_delta_z,
                          const slc_dyna_parity *d_parity, void *cuda_stream);
int slc_dyna_track_host(slc_context *ctx, const uint8_t *h_frames, int32_t n_frames, int32_t window,
                        const double *h_u0, float *h_xyzw, uint8_t *h_mask, float *h_delta_z,
                        const slc_dyna_parity *h_parity);

/* slc_dyna_track_host with a result format for the n_frames-1 maps (slc_result, above): the host path of the
 * dynamic frames is bound by the 21 B/px it sends back, and SLC_RESULT_DEPTH (z + one bit per pixel; deltaZ is
 * z[f] - z[f-1]) or SLC_RESULT_POINTS (the valid points of every frame, as Result() would print them) are a
 * fifth / about two thirds of that.  Each format is a bit-for-bit selection of the xyzw + mask maps. */
int slc_dyna_track_host_ex(slc_context *ctx, const uint8_t *h_frames, int32_t n_frames, int32_t window,
                           const double *h_u0, const slc_result *h_out);

/* ---- input ingest ("next" row: CSensor::LoadDatas) ------------------------- */
/* What the reference's imread(file, CV_LOAD_IMAGE_GRAYSCALE) (CSensorV.cpp:111-114) yields for a
 * .bmp: uncompressed 8-bit paletted, 24-bit BGR or 32-bit BGRA, bottom-up or top-down. */
typedef struct {
    int32_t width, height;        /* pixels */
    int32_t bits_per_pixel;       /* 8, 24 or 32 */
    int32_t top_down;             /* 1: first stored row is the top row (biHeight < 0) */
    int32_t row_stride;           /* bytes per stored row (padded to 4) */
    int32_t palette_is_identity;  /* 8 bpp: gray[i] == i for every i */
    int64_t pixel_offset;         /* bfOffBits: start of the pixel array inside the file */
    uint8_t gray[256];            /* 8 bpp: palette index -> gray level, OpenCV's fixed-point BGR weights */
} slc_bmp_info;

/* Header + palette only (host, no context, no pixel work).  SLC_ERR_INVALID_ARG for anything that
 * is not one of the flavours above (RLE, 1/4/16 bpp, truncated file). */
int slc_bmp_parse(const void *file_bytes, int64_t n_bytes, slc_bmp_info *info);
/* Pixel array (device pointer to file_bytes + pixel_offset) -> one [height][width] u8 plane on the
 * device: row flip, padding removal, palette / BGR -> gray.  Asynchronous. */
int slc_bmp_unpack_device(slc_context *ctx, const uint8_t *d_pixels, const slc_bmp_info *info,
                          uint8_t *d_plane, void *cuda_stream);
/* The same for a whole set of files in ONE launch: pixel array i (d_pixels is a HOST array of n_files
 * device pointers, infos a host array) -> plane i of d_stack (u8 [n_files][H][W]; every file must have
 * the context's camera size).  This is the loop of CCalculation::FillFirstProjectorU
 * (CCalculation.cpp:536-557) over images that are already in device memory.  Asynchronous. */
int slc_bmp_unpack_batch_device(slc_context *ctx, const uint8_t *const *d_pixels, const slc_bmp_info *infos,
                                int32_t n_files, uint8_t *d_stack, void *cuda_stream);
/* The whole file from host memory to a host plane (upload, unpack, download; synchronous): what
 * the file-backed CSensor hands to SetMat. */
int slc_bmp_decode_host(slc_context *ctx, const void *h_file_bytes, int64_t n_bytes, uint8_t *h_plane,
                        int32_t expect_width, int32_t expect_height);
/* replaces: CSensor::LoadDatas + the SetProPicture/GetCamPicture/SetMat loop of
 * CCalculation::FillFirstProjectorU (CCalculation.cpp:536-557).  Reads n_files .bmp files (each must
 * be width x height of the context) through pinned double-buffered staging and unpacks file i
 * into plane i of d_stack (u8 [n_files][H][W], e.g. a whole 2G+N stack or a dynaCam sequence).
 * Synchronous; on failure the message names the file. */
int slc_load_bmp_planes(slc_context *ctx, const char *const *paths, int32_t n_files, uint8_t *d_stack);

/* ---- point-cloud output ("next" row: CCalculation::Result) ----------------- */
#define SLC_TEXT_CRLF (1u << 0) /* "\r\n" line ends: what the reference's text-mode fstream writes on its platform */
#define SLC_TEXT_EXP3 (1u << 1) /* three exponent digits (the MSVC 2013 CRT of the reference build) instead of two */

/* replaces: CCalculation::Result (CCalculation.cpp:323-357).  The text the reference writes for one
 * frame -- "x y z\n" per pixel with FOV_MIN <= z <= FOV_MAX, u outer / v inner, each number as
 * `ostream << double` prints it (printf "%g", precision 6) -- produced on the device: x, y, z are
 * recomputed in f64 from the f64 ProjectorU plane in the reference's operation order
 * (:686-687, :761-767), formatted with exact round-half-even decimal conversion, and compacted
 * into one contiguous buffer.  text must be 16-byte aligned; *bytes receives the size of the
 * whole text even when it exceeds capacity (then nothing past capacity is written and the call
 * returns SLC_ERR_INVALID_ARG); *points the number of lines.  Worst case 43 bytes per pixel.
 * The _device call synchronises the stream to return the two totals. */
int slc_pointcloud_text_device(slc_context *ctx, const double *d_proj_u, uint32_t flags, char *d_text,
                               int64_t capacity, int64_t *bytes, int64_t *points, void *cuda_stream);
int slc_pointcloud_text_host(slc_context *ctx, const double *h_proj_u, uint32_t flags, char *h_text,
                             int64_t capacity, int64_t *bytes, int64_t *points);
/* Binary cloud: float[points][3] = (x, y, z) of the pixels with mask != 0 of one (xyzw, mask) map,
 * in SLC_ORDER_ROW_MAJOR or SLC_ORDER_REFERENCE order.  capacity_points bounds d_xyz. */
int slc_pointcloud_compact_device(slc_context *ctx, const float *d_xyzw, const uint8_t *d_mask, int32_t order,
                                  float *d_xyz, int64_t capacity_points, int64_t *points, void *cuda_stream);
int slc_pointcloud_compact_host(slc_context *ctx, const float *h_xyzw, const uint8_t *h_mask, int32_t order,
                                float *h_xyz, int64_t capacity_points, int64_t *points);
/* The same for n_maps maps in ONE asynchronous launch (e.g. every frame of a dynamic sequence):
 * d_points [n_maps][point_stride][3], d_n_points [n_maps] and the optional d_mask_bits [n_maps][(H*W+7)/8]
 * stay on the device; nothing is read back, the call does not synchronise.  A map with more valid pixels
 * than point_stride keeps the first point_stride of them (d_n_points still counts them all).  Needs a
 * camera width that is a multiple of 8. */
int slc_compact_points_device(slc_context *ctx, const float *d_xyzw, const uint8_t *d_mask, int32_t n_maps, int32_t order,
                              float *d_points, int64_t point_stride, uint8_t *d_mask_bits, int64_t *d_n_points,
                              void *cuda_stream);
/* Parity hook for the number formatting alone: n doubles -> n slots of 16 chars (zero padded) and
 * their lengths, exactly as the text kernel formats them. */
int slc_format_g6_host(slc_context *ctx, const double *h_values, int64_t n, uint32_t flags, char *h_text16,
                       uint8_t *h_len);

/* Parity hook for the arctangent of CDecodePhase.cpp:67-75 alone: evaluates
 * cvFastArctan(sin, cos) in degrees and the in-period offset on the device, with
 * exactly the arithmetic of the fused kernel, for n caller-supplied pairs. */
int slc_eval_phase_host(slc_context *ctx, const float *h_sin, const float *h_cos, int64_t n,
                        float *h_deg, float *h_pix);

/* ---- measurement ------------------------------------------------------ */
/* Launch the fused kernel `iters` times back to back on the context stream and
 * report the average duration per launch in milliseconds, measured with CUDA
 * events recorded on that same stream. */
int slc_time_reconstruct_device(slc_context *ctx, const uint8_t *d_stack, int32_t n_stacks,
                                float *d_xyzw, uint8_t *d_mask, int32_t iters,
                                float *ms_per_launch);
/* Number of kernels this library has launched on this context so far. */
int64_t slc_launch_count(const slc_context *ctx);
/* Tuning hook for bench/tests: pixels owned by one thread of this context's vector kernel
 * (4, 8 or 16; 0 = chosen per geometry, the default).  Results do not depend on it. */
int slc_set_pixels_per_thread(slc_context *ctx, int32_t pxt);

#ifdef __cplusplus
}
#endif
#endif /* SLCALC_B200_H_ */
