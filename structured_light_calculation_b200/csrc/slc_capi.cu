// slc_capi.cu -- the C ABI of include/slcalc_b200.h: context, calibration
// set-up, device/pinned memory, stream slots, and the launches.
//
// Host-side f64 arithmetic that must reproduce CCalculation::Init
// (CCalculation.cpp:135-166) is written one operation per statement and the
// file is compiled with -ffp-contract=off, so it matches an SSE2 /fp:precise
// build of the reference operation for operation.
#include "../../include/slcalc_b200.h"
#include "slc_kernels.h"

#include <sys/mman.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

using slc::KParams;

namespace {

thread_local std::string g_create_error;

// huge-page pinned blocks handed out by slc_host_alloc_ex (they are unmapped, not cudaFreeHost'ed)
std::mutex g_huge_mu;
std::vector<std::pair<void*, size_t>> g_huge;

struct Slot {
    cudaStream_t stream = nullptr;
    uint8_t* d_stack = nullptr;
    float* d_xyzw = nullptr;
    uint8_t* d_mask = nullptr;
    // parity planes, allocated on first use
    int16_t* d_kbin = nullptr;
    int8_t* d_corr = nullptr;
    float* d_pix = nullptr;
    double* d_proj_u = nullptr;
    // SLC_RESULT_POINTS, allocated on first use: packed points, bit mask, counts, look-back state
    float* d_points = nullptr;
    uint8_t* d_bits = nullptr;
    unsigned long long* d_counts = nullptr;
    unsigned long long* d_cstate = nullptr;
    unsigned long long* h_counts = nullptr;     // pinned
    cudaEvent_t counts_ready = nullptr;
    unsigned epoch = 0;
    // a POINTS chunk whose counts are on their way: its point download is issued once they are known
    struct { bool active = false; int n = 0; float* h_points = nullptr; int64_t* h_n_points = nullptr; int64_t stride = 0; } pending;
    bool busy = false;
};

}  // namespace

constexpr int kBmpSlots = 8;   // staging slots of slc_load_bmp_planes

struct slc_context {
    slc_config cfg{};
    int gp = 0, T = 0;
    int sm_count = 0;
    bool calibrated = false;
    KParams kp{};            // geometry + calibration constants, buffers unset
    int16_t* d_lut = nullptr;
    cudaStream_t stream = nullptr;
    std::vector<Slot> slots;
    // scratch for the stand-alone decoder entry points
    void* d_scratch_in = nullptr;  size_t scratch_in_bytes = 0;
    void* d_scratch_out = nullptr; size_t scratch_out_bytes = 0;
    void* d_scratch_aux = nullptr; size_t scratch_aux_bytes = 0;
    void* d_strips = nullptr;      size_t strips_bytes = 0;      // dynamic frames: (stripB, stripW) per frame
    void* d_dsums = nullptr;       size_t dsums_bytes = 0;       // dynamic frames: 3x3 sums of the nearer delta
    void* d_pc_scratch = nullptr;  size_t pc_scratch_bytes = 0;  // point cloud: block sums, totals, look-back words
    unsigned pc_epoch = 0;                                       // launch epoch of the text kernel's look-back words
    void* d_pc_in = nullptr;       size_t pc_in_bytes = 0;       // point cloud: staging for the host entry points
    void* d_pc_out = nullptr;      size_t pc_out_bytes = 0;
    void* h_pc_totals = nullptr;                                   // point cloud: pinned (bytes, records) read-back
    void* d_bmp[kBmpSlots] = {};   size_t bmp_bytes[kBmpSlots] = {};     // ingest: raw file staging (device)
    void* h_bmp[kBmpSlots] = {};   size_t h_bmp_bytes[kBmpSlots] = {};   // ingest: raw file staging (pinned)
    cudaEvent_t bmp_done[kBmpSlots] = {};
    void* d_dyna = nullptr;        size_t dyna_bytes = 0;        // dynamic frames: staging for the host entry point
    slc::LaunchPlan plans[3][2][2];                              // [mode][output layout][small launch], chosen on first use
    int pxt_override = 0;                                        // slc_set_pixels_per_thread
    void* d_cstate = nullptr;      size_t cstate_bytes = 0;      // POINTS on the device path: look-back state
    unsigned cstate_epoch = 0;
    long long launches = 0;
    std::string err;
};

namespace {

int fail(slc_context* ctx, int status, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return status;
}

#define SLC_CUDA(ctx, call)                                                                     \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? SLC_ERR_OUT_OF_MEMORY : SLC_ERR_CUDA, \
                        "%s failed: %s", #call, cudaGetErrorString(e__));                       \
    } while (0)

size_t stack_bytes(const slc_context* c) { return (size_t)c->kp.P * (size_t)c->kp.npx; }
size_t xyzw_bytes(const slc_context* c) { return 16 * (size_t)c->kp.npx; }
size_t mask_bytes(const slc_context* c) { return (size_t)c->kp.npx; }

int ensure_scratch(slc_context* ctx, void** p, size_t* have, size_t want)
{
    if (*have >= want) return SLC_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *have = 0;
    SLC_CUDA(ctx, cudaMalloc(p, want));
    *have = want;
    return SLC_OK;
}

int ensure_parity(slc_context* ctx, Slot& s, const slc_parity_planes* want)
{
    const size_t n = (size_t)ctx->cfg.max_batch * (size_t)ctx->kp.npx;
    if (want->kbin && !s.d_kbin) SLC_CUDA(ctx, cudaMalloc(&s.d_kbin, n * sizeof(int16_t)));
    if (want->corr && !s.d_corr) SLC_CUDA(ctx, cudaMalloc(&s.d_corr, n * sizeof(int8_t)));
    if (want->phase_pix && !s.d_pix) SLC_CUDA(ctx, cudaMalloc(&s.d_pix, n * sizeof(float)));
    if (want->proj_u && !s.d_proj_u) SLC_CUDA(ctx, cudaMalloc(&s.d_proj_u, n * sizeof(double)));
    return SLC_OK;
}

size_t bits_bytes(const slc_context* c) { return ((size_t)c->kp.npx + 7) / 8; }

// A launch below this many pixels does not fill the GPU for long enough to hide its ramp-up and tail:
// the 4-pixels-per-thread shape (twice the threads, finer tail) is faster there and slower above
// (profiles/r02_launch_size_curve*.txt: 1 / 4 / 8 / 256 frame sets of 1920x1200 run at 0.72 / 0.92 / 0.95 / 0.95
// of the HBM peak with 4 pixels per thread and at 0.67 / 0.88 / 0.95 / 1.01 with 8).
constexpr long long kSmallLaunchPixels = 16ll << 20;

// the kernel for (mode, output layout, launch size class) of this context: chosen once, then only launched
int plan_for(slc_context* ctx, int mode, int out, int n_stacks, const slc::LaunchPlan** plan)
{
    const int small = (ctx->pxt_override == 0 && (long long)n_stacks * ctx->kp.npx < kSmallLaunchPixels) ? 1 : 0;
    slc::LaunchPlan& pl = ctx->plans[mode][out][small];
    if (!pl.valid) {
        SLC_CUDA(ctx, slc::plan_reconstruct(ctx->kp, mode, out, ctx->pxt_override, &pl));
        if (small && pl.vec != nullptr && pl.pxt == 8) {
            // only where the geometry has its own 4-pixel instance: a generic one would lose more than it gains
            slc::LaunchPlan alt;
            SLC_CUDA(ctx, slc::plan_reconstruct(ctx->kp, mode, out, 4, &alt));
            if (alt.vec != nullptr && alt.pxt == 4 && alt.specialised == pl.specialised) pl = alt;
        }
    }
    *plan = &pl;
    return SLC_OK;
}

// One fused launch.  d_depth != nullptr selects the SLC_RESULT_DEPTH layout (d_bits beside it),
// otherwise xyzw + mask.
int launch(slc_context* ctx, const uint8_t* d_stack, int n_stacks, float* d_xyzw, uint8_t* d_mask,
           const slc_parity_planes* par, cudaStream_t stream, float* d_depth = nullptr, uint8_t* d_bits = nullptr)
{
    KParams p = ctx->kp;
    p.n_stacks = n_stacks;
    p.stack = d_stack;
    p.xyzw = reinterpret_cast<float4*>(d_xyzw);
    p.mask = d_mask;
    p.depth = d_depth;
    p.mask_bits = d_bits;
    p.bits_stride = (long long)bits_bytes(ctx);
    p.kbin = par ? par->kbin : nullptr;
    p.corr = par ? par->corr : nullptr;
    p.phase_pix = par ? par->phase_pix : nullptr;
    p.proj_u = par ? par->proj_u : nullptr;
    p.lut = ctx->d_lut;
    const bool scalar = (ctx->cfg.flags & SLC_FLAG_SCALAR_KERNEL) != 0;
    const slc::LaunchPlan* plan = nullptr;
    int rc = plan_for(ctx, slc::plan_mode(p), d_depth ? 1 : 0, n_stacks, &plan);
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, slc::launch_reconstruct(p, *plan, scalar, stream, nullptr));
    ctx->launches += (n_stacks + 65534) / 65535;
    return SLC_OK;
}

int check_ready(slc_context* ctx, const void* a, const void* b, const void* c, int n_stacks)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (!ctx->calibrated)
        return fail(ctx, SLC_ERR_NOT_INITIALISED, "calibration not set: call slc_set_calibration first");
    if (n_stacks < 0) return fail(ctx, SLC_ERR_INVALID_ARG, "n_stacks < 0");
    if (n_stacks == 0) return SLC_OK;             // an empty batch is a no-op: buffers may be NULL
    if (!a || !b || !c) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL buffer");
    return SLC_OK;
}

// device staging of a slot (max_batch frame sets), on first use
int ensure_slot(slc_context* ctx, Slot& s)
{
    const size_t nb = (size_t)ctx->cfg.max_batch;
    if (!s.d_stack) SLC_CUDA(ctx, cudaMalloc(&s.d_stack, stack_bytes(ctx) * nb));
    if (!s.d_xyzw) SLC_CUDA(ctx, cudaMalloc(&s.d_xyzw, xyzw_bytes(ctx) * nb));
    if (!s.d_mask) SLC_CUDA(ctx, cudaMalloc(&s.d_mask, mask_bytes(ctx) * nb));
    return SLC_OK;
}

// Enqueue upload + kernel + download of one chunk on a slot.
int enqueue_chunk(slc_context* ctx, Slot& s, const uint8_t* h_stack, int n, float* h_xyzw, uint8_t* h_mask,
                  const slc_parity_planes* h_par, size_t par_offset_px)
{
    const size_t npx = (size_t)ctx->kp.npx;
    {
        const int rc = ensure_slot(ctx, s);
        if (rc != SLC_OK) return rc;
    }
    SLC_CUDA(ctx, cudaMemcpyAsync(s.d_stack, h_stack, stack_bytes(ctx) * n, cudaMemcpyHostToDevice, s.stream));
    slc_parity_planes dpar{};
    const slc_parity_planes* dparp = nullptr;
    if (h_par && (h_par->kbin || h_par->corr || h_par->phase_pix || h_par->proj_u)) {
        int rc = ensure_parity(ctx, s, h_par);
        if (rc != SLC_OK) return rc;
        dpar.kbin = h_par->kbin ? s.d_kbin : nullptr;
        dpar.corr = h_par->corr ? s.d_corr : nullptr;
        dpar.phase_pix = h_par->phase_pix ? s.d_pix : nullptr;
        dpar.proj_u = h_par->proj_u ? s.d_proj_u : nullptr;
        dparp = &dpar;
    }
    int rc = launch(ctx, s.d_stack, n, s.d_xyzw, s.d_mask, dparp, s.stream);
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, cudaMemcpyAsync(h_xyzw, s.d_xyzw, xyzw_bytes(ctx) * n, cudaMemcpyDeviceToHost, s.stream));
    SLC_CUDA(ctx, cudaMemcpyAsync(h_mask, s.d_mask, mask_bytes(ctx) * n, cudaMemcpyDeviceToHost, s.stream));
    if (dparp) {
        const size_t cnt = npx * n;
        if (dpar.kbin)
            SLC_CUDA(ctx, cudaMemcpyAsync(h_par->kbin + par_offset_px, dpar.kbin, cnt * sizeof(int16_t),
                                          cudaMemcpyDeviceToHost, s.stream));
        if (dpar.corr)
            SLC_CUDA(ctx, cudaMemcpyAsync(h_par->corr + par_offset_px, dpar.corr, cnt * sizeof(int8_t),
                                          cudaMemcpyDeviceToHost, s.stream));
        if (dpar.phase_pix)
            SLC_CUDA(ctx, cudaMemcpyAsync(h_par->phase_pix + par_offset_px, dpar.phase_pix, cnt * sizeof(float),
                                          cudaMemcpyDeviceToHost, s.stream));
        if (dpar.proj_u)
            SLC_CUDA(ctx, cudaMemcpyAsync(h_par->proj_u + par_offset_px, dpar.proj_u, cnt * sizeof(double),
                                          cudaMemcpyDeviceToHost, s.stream));
    }
    s.busy = true;
    return SLC_OK;
}

}  // namespace

extern "C" {

namespace { int finish_points(slc_context* ctx, Slot& s, bool* overflow); }   // result formats, below


int slc_abi_version(void) { return SLC_ABI_VERSION; }

const char* slc_status_string(int status)
{
    switch (status) {
    case SLC_OK: return "ok";
    case SLC_ERR_INVALID_ARG: return "invalid argument";
    case SLC_ERR_NOT_INITIALISED: return "not initialised";
    case SLC_ERR_CUDA: return "CUDA error";
    case SLC_ERR_NO_DEVICE: return "no CUDA device (there is no CPU path)";
    case SLC_ERR_OUT_OF_MEMORY: return "out of memory";
    case SLC_ERR_STATE: return "invalid state";
    default: return "unknown status";
    }
}

const char* slc_last_error(const slc_context* ctx)
{
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int slc_create(const slc_config* cfg, slc_context** out)
{
    if (!cfg || !out) return fail(nullptr, SLC_ERR_INVALID_ARG, "slc_create: NULL argument");
    *out = nullptr;
    if (cfg->width <= 0 || cfg->height <= 0)
        return fail(nullptr, SLC_ERR_INVALID_ARG, "camera size %dx%d is not positive", cfg->width, cfg->height);
    if ((long long)cfg->width * cfg->height > 0x7fffffffLL)
        return fail(nullptr, SLC_ERR_INVALID_ARG, "camera size %dx%d exceeds 2^31 pixels", cfg->width, cfg->height);
    // CDecodeGray::SetNumDigit accepts 1..16 (CDecodeGray.cpp:39)
    if (cfg->gray_digits <= 0 || cfg->gray_digits > 16)
        return fail(nullptr, SLC_ERR_INVALID_ARG, "gray_digits %d outside 1..16", cfg->gray_digits);
    // CDecodePhase::SetNumMat rejects numMat <= 0 (CDecodePhase.cpp:122); phase shifting needs >= 3
    if (cfg->phase_steps < 3)
        return fail(nullptr, SLC_ERR_INVALID_ARG, "phase_steps %d < 3", cfg->phase_steps);
    const bool even = (cfg->phase_steps & 1) == 0;
    if ((even && cfg->phase_steps > 2 * slc::kMaxPhaseTable) || (!even && cfg->phase_steps > slc::kMaxPhaseTable))
        return fail(nullptr, SLC_ERR_INVALID_ARG, "phase_steps %d too large (max %d even / %d odd)",
                    cfg->phase_steps, 2 * slc::kMaxPhaseTable, slc::kMaxPhaseTable);
    const int gp = cfg->projector_width / (1 << cfg->gray_digits);          // CDecodeGray.cpp:183
    const int T = cfg->projector_width / (1 << (cfg->gray_digits - 1));     // CCalculation.cpp:550
    if (cfg->projector_width <= 0 || gp < 1)
        return fail(nullptr, SLC_ERR_INVALID_ARG,
                    "projector_width %d gives a Gray stripe of %d px for %d digits (needs >= 1; the reference "
                    "would divide by zero at CCalculation.cpp:570)", cfg->projector_width, gp, cfg->gray_digits);
    if (!(cfg->fov_min <= cfg->fov_max))
        return fail(nullptr, SLC_ERR_INVALID_ARG, "fov_min > fov_max");
    if (cfg->max_batch < 1 || cfg->num_slots < 1 || cfg->num_slots > 8)
        return fail(nullptr, SLC_ERR_INVALID_ARG, "max_batch must be >= 1 and num_slots in 1..8");
    if (cfg->modulation_min < 0.f || !(cfg->modulation_min == cfg->modulation_min))
        return fail(nullptr, SLC_ERR_INVALID_ARG, "modulation_min must be >= 0");

    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev <= 0)
        return fail(nullptr, SLC_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= n_dev)
        return fail(nullptr, SLC_ERR_INVALID_ARG, "device %d out of range (0..%d)", cfg->device, n_dev - 1);

    slc_context* ctx = new (std::nothrow) slc_context();
    if (!ctx) return fail(nullptr, SLC_ERR_OUT_OF_MEMORY, "host allocation failed");
    ctx->cfg = *cfg;
    ctx->gp = gp;
    ctx->T = T;
    KParams& p = ctx->kp;
    p.W = cfg->width; p.H = cfg->height;
    p.G = cfg->gray_digits; p.N = cfg->phase_steps; p.P = 2 * p.G + p.N;
    p.npx = (long long)p.W * p.H;
    p.gp = gp; p.T = T;
    p.Tf = (float)T; p.T075 = (float)(0.75 * T); p.T025 = (float)(0.25 * T); p.halfT = (float)(0.5 * T);
    p.gpf = (float)gp;
    // [EXT] phase tables and modulation threshold; same expressions as the oracle
    const int nk = (p.N == 4) ? 0 : (even ? p.N / 2 : p.N);
    for (int k = 0; k < slc::kMaxPhaseTable; k++) { p.ck[k] = 0.f; p.sk[k] = 0.f; }
    for (int k = 0; k < nk; k++) {
        double c = std::cos(2.0 * M_PI * (double)k / (double)p.N);
        double s = std::sin(2.0 * M_PI * (double)k / (double)p.N);
        if (std::fabs(c) < 1e-9) c = 0.0;
        if (std::fabs(s) < 1e-9) s = 0.0;
        p.ck[k] = (float)c;
        p.sk[k] = (float)s;
    }
    {
        const double scale = p.N == 4 ? 1.0 : 0.5 * (double)p.N;
        const double t = (double)cfg->modulation_min * scale;
        p.thr2 = (float)(t * t);
        p.use_mod = cfg->modulation_min > 0.f ? 1 : 0;
    }
    p.z_fp64 = (cfg->flags & SLC_FLAG_Z_FP64) ? 1 : 0;
    p.magic_one = 0x4B000000u; p.magic_half = 0x4A800000u;
    p.fov_min = cfg->fov_min; p.fov_max = cfg->fov_max;
    {
        // valid <=> |z - mid| <= half; pixels within `band` of either limit are re-solved in f64
        const double mid = 0.5 * (cfg->fov_min + cfg->fov_max), half = 0.5 * (cfg->fov_max - cfg->fov_min);
        const double band = 1e-3 * std::fmax(std::fmax(std::fabs(cfg->fov_min), std::fabs(cfg->fov_max)), 1e-30);
        p.fov_mid32 = (float)mid; p.fov_half32 = (float)half; p.guard_band = (float)band;
    }

#define SLC_CREATE_CUDA(call)                                                                      \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            int st__ = fail(nullptr, e__ == cudaErrorMemoryAllocation ? SLC_ERR_OUT_OF_MEMORY : SLC_ERR_CUDA, \
                            "%s failed: %s", #call, cudaGetErrorString(e__));                      \
            slc_destroy(ctx);                                                                      \
            return st__;                                                                           \
        }                                                                                          \
    } while (0)

    SLC_CREATE_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    SLC_CREATE_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    ctx->sm_count = prop.multiProcessorCount;
    SLC_CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->slots.resize(cfg->num_slots);
    // the slots' device buffers (stack + xyzw + mask, max_batch frame sets each) are allocated by the
    // first host-path call that needs them: a context that only serves the stand-alone decoder objects,
    // device-resident callers or the point-cloud / ingest entry points stays a few streams large
    for (Slot& s : ctx->slots) SLC_CREATE_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
#undef SLC_CREATE_CUDA
    *out = ctx;
    return SLC_OK;
}

void slc_destroy(slc_context* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    for (Slot& s : ctx->slots) {
        if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
        cudaFree(s.d_stack); cudaFree(s.d_xyzw); cudaFree(s.d_mask);
        cudaFree(s.d_kbin); cudaFree(s.d_corr); cudaFree(s.d_pix); cudaFree(s.d_proj_u);
        cudaFree(s.d_points); cudaFree(s.d_bits); cudaFree(s.d_counts); cudaFree(s.d_cstate);
        if (s.h_counts) cudaFreeHost(s.h_counts);
        if (s.counts_ready) cudaEventDestroy(s.counts_ready);
    }
    if (ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    cudaFree(ctx->d_lut);
    cudaFree(ctx->d_scratch_in); cudaFree(ctx->d_scratch_out); cudaFree(ctx->d_scratch_aux);
    cudaFree(ctx->d_strips); cudaFree(ctx->d_dsums); cudaFree(ctx->d_dyna); cudaFree(ctx->d_cstate);
    cudaFree(ctx->d_pc_scratch); cudaFree(ctx->d_pc_in); cudaFree(ctx->d_pc_out);
    if (ctx->h_pc_totals) cudaFreeHost(ctx->h_pc_totals);
    for (int k = 0; k < kBmpSlots; k++) {
        cudaFree(ctx->d_bmp[k]);
        if (ctx->h_bmp[k]) cudaFreeHost(ctx->h_bmp[k]);
        if (ctx->bmp_done[k]) cudaEventDestroy(ctx->bmp_done[k]);
    }
    delete ctx;
}

int slc_get_info(const slc_context* cctx, slc_info* out)
{
    slc_context* ctx = const_cast<slc_context*>(cctx);
    if (!ctx || !out) return SLC_ERR_INVALID_ARG;
    std::memset(out, 0, sizeof(*out));
    out->planes = ctx->kp.P;
    out->gray_period = ctx->gp;
    out->phase_period = ctx->T;
    out->sm_count = ctx->sm_count;
    out->pixels = ctx->kp.npx;
    out->stack_bytes = (int64_t)stack_bytes(ctx);
    out->xyzw_bytes = (int64_t)xyzw_bytes(ctx);
    out->mask_bytes = (int64_t)mask_bytes(ctx);
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    slc::LaunchInfo li;
    li.query_only = true;
    KParams p = ctx->kp;
    p.n_stacks = 1;
    // the shape the kernel takes for buffers aligned the way cudaMalloc aligns them
    p.stack = reinterpret_cast<const uint8_t*>(uintptr_t{256});
    p.xyzw = reinterpret_cast<float4*>(uintptr_t{256});
    p.mask = reinterpret_cast<uint8_t*>(uintptr_t{256});
    p.lut = ctx->d_lut;
    const slc::LaunchPlan* plan = nullptr;
    int rc = plan_for(ctx, slc::plan_mode(p), 0, 1 << 20, &plan);   // the shape of a large launch
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, slc::launch_reconstruct(p, *plan, (ctx->cfg.flags & SLC_FLAG_SCALAR_KERNEL) != 0, nullptr, &li));
    out->kernel_variant = li.variant;
    out->kernel_regs = li.regs;
    out->kernel_block = li.block;
    out->kernel_smem = li.smem;
    return SLC_OK;
}

int slc_set_calibration(slc_context* ctx, const double cam[9], const double pro[9], const double R[9],
                        const double T[3])
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (!cam || !pro || !R || !T) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL calibration matrix");
    for (int i = 0; i < 9; i++)
        if (!std::isfinite(cam[i]) || !std::isfinite(pro[i]) || !std::isfinite(R[i]))
            return fail(ctx, SLC_ERR_INVALID_ARG, "calibration contains a non-finite value");
    for (int i = 0; i < 3; i++)
        if (!std::isfinite(T[i])) return fail(ctx, SLC_ERR_INVALID_ARG, "calibration contains a non-finite value");
    if (cam[0] == 0.0 || cam[4] == 0.0) return fail(ctx, SLC_ERR_INVALID_ARG, "camera focal length is zero");

    // P = ProMat * [R | T]  (CCalculation.cpp:141-145), accumulated k = 0,1,2
    double RT[12], P[12];
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) RT[r * 4 + c] = R[r * 3 + c];
        RT[r * 4 + 3] = T[r];
    }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) {
                const double prod = pro[r * 3 + k] * RT[k * 4 + c];
                s = s + prod;
            }
            P[r * 4 + c] = s;
        }
    KParams& p = ctx->kp;
    const double fu = cam[0], fv = cam[4], cu = cam[2], cv = cam[5];
    const double fufv = fu * fv;
    p.fu = fu; p.fv = fv; p.cu = cu; p.cv = cv;
    p.A = fufv * P[3];        // :151
    p.B = fufv * P[11];       // :152
    p.P00 = P[0]; p.P01 = P[1]; p.fufvP02 = fufv * P[2];     // :159-161
    p.P20 = P[8]; p.P21 = P[9]; p.fufvP22 = fufv * P[10];    // :162-164
    p.E = fufv * P[7];                                        // [EXT] row 1 of P, same construction
    p.P10 = P[4]; p.P11 = P[5]; p.fufvP12 = fufv * P[6];

    // f32 coefficients of the same rational map, divided through by fu*fv:
    // C(u,v) = fv*P00*(u-cu) + fu*P01*(v-cv) + fu*fv*P02, D likewise with row 2.
    const double s = 1.0 / fufv;
    const double au = fv * P[0], av = fu * P[1], a0 = fufv * P[2];
    const double bu = fv * P[8], bv = fu * P[9], b0 = fufv * P[10];
    p.A32 = (float)(p.A * s); p.B32 = (float)(p.B * s);
    p.c0 = (float)((a0 - au * cu - av * cv) * s); p.cu1 = (float)(au * s); p.cv1 = (float)(av * s);
    p.d0 = (float)((b0 - bu * cu - bv * cv) * s); p.du1 = (float)(bu * s); p.dv1 = (float)(bv * s);
    // Cancellation guards: 2^-8 of the largest magnitude that can be summed into num / den
    // anywhere in the image for any decodable U (|U| <= PW + 2T).
    {
        const double as = std::fabs(s);
        const double umax = (double)(p.W - 1), vmax = (double)(p.H - 1);
        const double Umax = (double)ctx->cfg.projector_width + 2.0 * (double)ctx->T;
        const double Cm = (std::fabs(a0) + std::fabs(au * cu) + std::fabs(av * cv) + std::fabs(au) * umax +
                           std::fabs(av) * vmax) * as;
        const double Dm = (std::fabs(b0) + std::fabs(bu * cu) + std::fabs(bv * cv) + std::fabs(bu) * umax +
                           std::fabs(bv) * vmax) * as;
        p.num_guard = (float)((std::fabs(p.B * s) * Umax + std::fabs(p.A * s)) / 256.0);
        p.den_guard = (float)((Cm + Dm * Umax) / 256.0);
    }
    p.rx1 = (float)(1.0 / fu); p.rx0 = (float)(-cu / fu);    // x = z*(u-cu)/fu  (:766)
    p.ry1 = (float)(1.0 / fv); p.ry0 = (float)(-cv / fv);    // y = z*(v-cv)/fv  (:767)
    ctx->calibrated = true;
    return SLC_OK;
}

int slc_set_gray_lut(slc_context* ctx, const int16_t* lut, int32_t n)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    const int want = 1 << ctx->cfg.gray_digits;
    if (!lut || n != want) return fail(ctx, SLC_ERR_INVALID_ARG, "gray LUT must have 2^G = %d entries", want);
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    bool standard = true;
    for (int g = 0; g < want && standard; g++) {
        int b = g;
        for (int sft = 1; sft < 16; sft <<= 1) b ^= b >> sft;   // inverse of bin ^ (bin >> 1)
        standard = (lut[g] == (int16_t)b);
    }
    for (Slot& s : ctx->slots) SLC_CUDA(ctx, cudaStreamSynchronize(s.stream));
    SLC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (standard) {   // the shipped Patterns/vGrayCode.txt: decode arithmetically
        cudaFree(ctx->d_lut);
        ctx->d_lut = nullptr;
        return SLC_OK;
    }
    if (!ctx->d_lut) SLC_CUDA(ctx, cudaMalloc(&ctx->d_lut, sizeof(int16_t) * 65536));
    SLC_CUDA(ctx, cudaMemset(ctx->d_lut, 0, sizeof(int16_t) * 65536));
    SLC_CUDA(ctx, cudaMemcpy(ctx->d_lut, lut, sizeof(int16_t) * want, cudaMemcpyHostToDevice));
    return SLC_OK;
}

/* ---- memory ---------------------------------------------------------- */
void* slc_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) return nullptr;
    return p;
}
void slc_host_free(void* p)
{
    if (!p) return;
    size_t huge = 0;
    {
        std::lock_guard<std::mutex> lk(g_huge_mu);
        for (size_t i = 0; i < g_huge.size(); i++)
            if (g_huge[i].first == p) { huge = g_huge[i].second; g_huge.erase(g_huge.begin() + (long)i); break; }
    }
    if (huge) { cudaHostUnregister(p); munmap(p, huge); }
    else cudaFreeHost(p);
}
void* slc_host_alloc_ex(size_t bytes, uint32_t flags)
{
    if (bytes == 0) bytes = 1;
    if (flags & SLC_HOST_HUGE_PAGES) {
        const size_t two_mb = (size_t)2 << 20, len = (bytes + two_mb - 1) & ~(two_mb - 1);
        void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_HUGETLB, -1, 0);
        if (p == MAP_FAILED) {
            // no hugetlb pool: an aligned anonymous mapping the kernel may back with transparent huge pages
            void* raw = mmap(nullptr, len + two_mb, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (raw == MAP_FAILED) return nullptr;
            const uintptr_t a = (reinterpret_cast<uintptr_t>(raw) + two_mb - 1) & ~(uintptr_t)(two_mb - 1);
            if (a > reinterpret_cast<uintptr_t>(raw)) munmap(raw, a - reinterpret_cast<uintptr_t>(raw));
            const uintptr_t end = reinterpret_cast<uintptr_t>(raw) + len + two_mb;
            if (end > a + len) munmap(reinterpret_cast<void*>(a + len), end - (a + len));
            p = reinterpret_cast<void*>(a);
            madvise(p, len, MADV_HUGEPAGE);
        }
        std::memset(p, 0, len);   // fault the pages in before pinning them
        if (cudaHostRegister(p, len, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); munmap(p, len); return nullptr; }
        std::lock_guard<std::mutex> lk(g_huge_mu);
        g_huge.emplace_back(p, len);
        return p;
    }
    void* p = nullptr;
    const unsigned f = cudaHostAllocPortable | ((flags & SLC_HOST_WRITE_COMBINED) ? cudaHostAllocWriteCombined : 0u);
    if (cudaHostAlloc(&p, bytes, f) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
int slc_host_register(void* p, size_t bytes)
{
    return cudaHostRegister(p, bytes, cudaHostRegisterPortable) == cudaSuccess ? SLC_OK : SLC_ERR_CUDA;
}
int slc_host_unregister(void* p) { return cudaHostUnregister(p) == cudaSuccess ? SLC_OK : SLC_ERR_CUDA; }

void* slc_device_alloc(slc_context* ctx, size_t bytes)
{
    if (!ctx) return nullptr;
    void* p = nullptr;
    if (cudaSetDevice(ctx->cfg.device) != cudaSuccess) return nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { fail(ctx, SLC_ERR_OUT_OF_MEMORY, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return nullptr; }
    return p;
}
void slc_device_free(slc_context* ctx, void* p)
{
    if (!ctx || !p) return;
    cudaSetDevice(ctx->cfg.device);
    cudaFree(p);
}
int slc_copy_to_device(slc_context* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    SLC_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    SLC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SLC_OK;
}
int slc_copy_to_host(slc_context* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    SLC_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SLC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SLC_OK;
}
int slc_synchronize(slc_context* ctx)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    for (Slot& s : ctx->slots) { SLC_CUDA(ctx, cudaStreamSynchronize(s.stream)); s.busy = false; }
    SLC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SLC_OK;
}

/* ---- hot path --------------------------------------------------------- */
int slc_reconstruct_device(slc_context* ctx, const uint8_t* d_stack, int32_t n_stacks, float* d_xyzw,
                           uint8_t* d_mask, const slc_parity_planes* d_parity, void* cuda_stream)
{
    int rc = check_ready(ctx, d_stack, d_xyzw, d_mask, n_stacks);
    if (rc != SLC_OK) return rc;
    if (n_stacks == 0) return SLC_OK;
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    // any number of frame sets: launch_reconstruct splits beyond the 65535 of grid.y
    return launch(ctx, d_stack, n_stacks, d_xyzw, d_mask, d_parity, st);
}

int slc_reconstruct_host(slc_context* ctx, const uint8_t* h_stack, int32_t n_stacks, float* h_xyzw,
                         uint8_t* h_mask, const slc_parity_planes* h_parity)
{
    int rc = check_ready(ctx, h_stack, h_xyzw, h_mask, n_stacks);
    if (rc != SLC_OK) return rc;
    if (n_stacks == 0) return SLC_OK;
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t npx = (size_t)ctx->kp.npx;
    const int chunk = ctx->cfg.max_batch;
    int slot = 0;
    // Stream order on a slot serialises "download chunk i" before "upload chunk
    // i + num_slots" into the same device buffers; different slots overlap.
    for (int done = 0; done < n_stacks; done += chunk) {
        const int n = (n_stacks - done) < chunk ? (n_stacks - done) : chunk;
        Slot& s = ctx->slots[slot];
        rc = enqueue_chunk(ctx, s, h_stack + (size_t)done * stack_bytes(ctx), n, h_xyzw + (size_t)done * npx * 4,
                           h_mask + (size_t)done * npx, h_parity, (size_t)done * npx);
        if (rc != SLC_OK) {
            // earlier chunks are still copying into the caller's buffers: drain them before returning
            for (Slot& q : ctx->slots) { cudaStreamSynchronize(q.stream); q.busy = false; }
            return rc;
        }
        slot = (slot + 1) % (int)ctx->slots.size();
    }
    for (Slot& s : ctx->slots) { SLC_CUDA(ctx, cudaStreamSynchronize(s.stream)); s.busy = false; }
    return SLC_OK;
}

int slc_submit_host(slc_context* ctx, int32_t slot, const uint8_t* h_stack, int32_t n_stacks, float* h_xyzw,
                    uint8_t* h_mask)
{
    int rc = check_ready(ctx, h_stack, h_xyzw, h_mask, n_stacks);
    if (rc != SLC_OK) return rc;
    if (slot < 0 || slot >= (int)ctx->slots.size()) return fail(ctx, SLC_ERR_INVALID_ARG, "slot %d out of range", slot);
    if (n_stacks < 1 || n_stacks > ctx->cfg.max_batch)
        return fail(ctx, SLC_ERR_INVALID_ARG, "n_stacks %d outside 1..max_batch (%d)", n_stacks, ctx->cfg.max_batch);
    Slot& s = ctx->slots[slot];
    if (s.busy) return fail(ctx, SLC_ERR_STATE, "slot %d still in flight: call slc_wait first", slot);
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    return enqueue_chunk(ctx, s, h_stack, n_stacks, h_xyzw, h_mask, nullptr, 0);
}

int slc_wait(slc_context* ctx, int32_t slot)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (slot < 0 || slot >= (int)ctx->slots.size()) return fail(ctx, SLC_ERR_INVALID_ARG, "slot %d out of range", slot);
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    Slot& s = ctx->slots[slot];
    bool overflow = false;
    const int64_t stride = s.pending.stride;
    const int rc = finish_points(ctx, s, &overflow);     // SLC_RESULT_POINTS: the counts are known now, fetch the lists
    const cudaError_t e = cudaStreamSynchronize(s.stream);
    s.busy = false;
    if (rc != SLC_OK) return rc;
    if (e != cudaSuccess) return fail(ctx, SLC_ERR_CUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(e));
    if (overflow)
        return fail(ctx, SLC_ERR_INVALID_ARG, "a frame set has more valid pixels than point_stride = %lld: its list was cut (n_points holds the full counts)",
                    (long long)stride);
    return SLC_OK;
}

/* ---- result formats ---------------------------------------------------- */
namespace {

int check_result(slc_context* ctx, const slc_result* r, bool device_path)
{
    if (!r) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL slc_result");
    switch (r->format) {
    case SLC_RESULT_XYZW:
        if (!r->xyzw || !r->mask) return fail(ctx, SLC_ERR_INVALID_ARG, "SLC_RESULT_XYZW needs xyzw and mask");
        break;
    case SLC_RESULT_DEPTH:
        if (!r->depth || !r->mask_bits) return fail(ctx, SLC_ERR_INVALID_ARG, "SLC_RESULT_DEPTH needs depth and mask_bits");
        if (device_path && (reinterpret_cast<uintptr_t>(r->mask_bits) & 3))
            return fail(ctx, SLC_ERR_INVALID_ARG, "mask_bits must be 4-byte aligned");
        break;
    case SLC_RESULT_POINTS:
        if (!r->points || !r->n_points || r->point_stride < 0)
            return fail(ctx, SLC_ERR_INVALID_ARG, "SLC_RESULT_POINTS needs points, n_points and point_stride >= 0");
        if (r->order != SLC_ORDER_ROW_MAJOR && r->order != SLC_ORDER_REFERENCE)
            return fail(ctx, SLC_ERR_INVALID_ARG, "order %d is neither SLC_ORDER_ROW_MAJOR nor SLC_ORDER_REFERENCE", r->order);
        if (device_path && (!r->xyzw || !r->mask))
            return fail(ctx, SLC_ERR_INVALID_ARG, "SLC_RESULT_POINTS on the device path needs xyzw and mask (the maps the points are taken from)");
        if (ctx->kp.W % 8 != 0 || ctx->kp.H > 12000)
            return fail(ctx, SLC_ERR_INVALID_ARG, "SLC_RESULT_POINTS needs a camera width that is a multiple of 8 and a height of at most 12000");
        break;
    default:
        return fail(ctx, SLC_ERR_INVALID_ARG, "unknown result format %d", r->format);
    }
    return SLC_OK;
}

// per-slot buffers of the POINTS pipeline, on first use
int ensure_points(slc_context* ctx, Slot& s)
{
    const size_t nb = (size_t)ctx->cfg.max_batch, npx = (size_t)ctx->kp.npx;
    if (!s.d_points) SLC_CUDA(ctx, cudaMalloc(&s.d_points, nb * npx * 12));
    if (!s.d_bits) SLC_CUDA(ctx, cudaMalloc(&s.d_bits, (nb * bits_bytes(ctx) + 3) & ~(size_t)3));
    if (!s.d_counts) SLC_CUDA(ctx, cudaMalloc(&s.d_counts, nb * sizeof(unsigned long long)));
    if (!s.d_cstate) {
        const size_t bytes = slc::compact_state_bytes(ctx->kp.W, ctx->kp.H, (int)nb);
        SLC_CUDA(ctx, cudaMalloc(&s.d_cstate, bytes));
        SLC_CUDA(ctx, cudaMemsetAsync(s.d_cstate, 0, bytes, s.stream));
        s.epoch = 0;
    }
    if (!s.h_counts) SLC_CUDA(ctx, cudaHostAlloc(&s.h_counts, nb * sizeof(unsigned long long), cudaHostAllocDefault));
    if (!s.counts_ready) SLC_CUDA(ctx, cudaEventCreateWithFlags(&s.counts_ready, cudaEventDisableTiming));
    return SLC_OK;
}

// The counts of the slot's POINTS chunk are on their way (or there): once known, download exactly
// 12 * count bytes per frame set.
int finish_points(slc_context* ctx, Slot& s, bool* overflow)
{
    if (!s.pending.active) return SLC_OK;
    s.pending.active = false;
    const int64_t point_stride = s.pending.stride;
    SLC_CUDA(ctx, cudaEventSynchronize(s.counts_ready));
    const size_t npx = (size_t)ctx->kp.npx;
    for (int i = 0; i < s.pending.n; i++) {
        const int64_t cnt = (int64_t)s.h_counts[i];
        s.pending.h_n_points[i] = cnt;
        if (cnt > point_stride) *overflow = true;
        const int64_t m = cnt < point_stride ? cnt : point_stride;
        if (m > 0)
            SLC_CUDA(ctx, cudaMemcpyAsync(s.pending.h_points + (size_t)i * (size_t)point_stride * 3,
                                          s.d_points + (size_t)i * npx * 3, (size_t)m * 12, cudaMemcpyDeviceToHost, s.stream));
    }
    return SLC_OK;
}

// Upload + kernel(s) + download of one chunk in DEPTH or POINTS format; `o` already points at this chunk.
int enqueue_chunk_fmt(slc_context* ctx, Slot& s, const uint8_t* h_stack, int n, const slc_result& o)
{
    const size_t npx = (size_t)ctx->kp.npx, bb = bits_bytes(ctx);
    {
        const int rc = ensure_slot(ctx, s);
        if (rc != SLC_OK) return rc;
    }
    SLC_CUDA(ctx, cudaMemcpyAsync(s.d_stack, h_stack, stack_bytes(ctx) * n, cudaMemcpyHostToDevice, s.stream));
    if (o.format == SLC_RESULT_DEPTH) {
        // the slot's xyzw / mask buffers hold the (smaller) depth plane / bit mask
        int rc = launch(ctx, s.d_stack, n, nullptr, nullptr, nullptr, s.stream, s.d_xyzw, s.d_mask);
        if (rc != SLC_OK) return rc;
        SLC_CUDA(ctx, cudaMemcpyAsync(o.depth, s.d_xyzw, npx * 4 * n, cudaMemcpyDeviceToHost, s.stream));
        SLC_CUDA(ctx, cudaMemcpyAsync(o.mask_bits, s.d_mask, bb * n, cudaMemcpyDeviceToHost, s.stream));
    } else {
        int rc = ensure_points(ctx, s);
        if (rc == SLC_OK) rc = launch(ctx, s.d_stack, n, s.d_xyzw, s.d_mask, nullptr, s.stream);
        if (rc != SLC_OK) return rc;
        SLC_CUDA(ctx, slc::launch_compact(ctx->kp.W, ctx->kp.H, n, o.order, s.d_xyzw, s.d_mask, s.d_points, (long long)npx,
                                          o.mask_bits ? s.d_bits : nullptr, (long long)bb, s.d_counts, s.d_cstate,
                                          ++s.epoch, s.stream));
        ctx->launches++;
        SLC_CUDA(ctx, cudaMemcpyAsync(s.h_counts, s.d_counts, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost, s.stream));
        SLC_CUDA(ctx, cudaEventRecord(s.counts_ready, s.stream));
        if (o.xyzw) SLC_CUDA(ctx, cudaMemcpyAsync(o.xyzw, s.d_xyzw, npx * 16 * n, cudaMemcpyDeviceToHost, s.stream));
        if (o.mask) SLC_CUDA(ctx, cudaMemcpyAsync(o.mask, s.d_mask, npx * n, cudaMemcpyDeviceToHost, s.stream));
        if (o.mask_bits) SLC_CUDA(ctx, cudaMemcpyAsync(o.mask_bits, s.d_bits, bb * n, cudaMemcpyDeviceToHost, s.stream));
        s.pending.active = true;
        s.pending.n = n;
        s.pending.h_points = o.points;
        s.pending.h_n_points = o.n_points;
        s.pending.stride = o.point_stride;
    }
    s.busy = true;
    return SLC_OK;
}

// look-back state of the chained scan for n_stacks maps (context-owned: calls are stream-ordered)
int ensure_cstate(slc_context* ctx, int n_stacks, cudaStream_t st)
{
    const size_t want = slc::compact_state_bytes(ctx->kp.W, ctx->kp.H, n_stacks);
    if (ctx->cstate_bytes < want) {
        int rc = ensure_scratch(ctx, &ctx->d_cstate, &ctx->cstate_bytes, want);
        if (rc != SLC_OK) return rc;
        SLC_CUDA(ctx, cudaMemsetAsync(ctx->d_cstate, 0, want, st));
        ctx->cstate_epoch = 0;
    }
    return SLC_OK;
}

// `r` advanced by `done` frame sets
slc_result result_at(const slc_context* ctx, const slc_result& r, size_t done)
{
    const size_t npx = (size_t)ctx->kp.npx;
    slc_result o = r;
    if (r.xyzw) o.xyzw = r.xyzw + done * npx * 4;
    if (r.mask) o.mask = r.mask + done * npx;
    if (r.depth) o.depth = r.depth + done * npx;
    if (r.mask_bits) o.mask_bits = r.mask_bits + done * bits_bytes(ctx);
    if (r.points) o.points = r.points + done * (size_t)r.point_stride * 3;
    if (r.n_points) o.n_points = r.n_points + done;
    return o;
}

}  // namespace

int slc_reconstruct_device_ex(slc_context* ctx, const uint8_t* d_stack, int32_t n_stacks, const slc_result* d_out,
                              void* cuda_stream)
{
    int rc = check_ready(ctx, d_stack, d_out, d_out, n_stacks);
    if (rc != SLC_OK) return rc;
    if (n_stacks == 0) return SLC_OK;
    rc = check_result(ctx, d_out, true);
    if (rc != SLC_OK) return rc;
    if (d_out->format == SLC_RESULT_XYZW)
        return slc_reconstruct_device(ctx, d_stack, n_stacks, d_out->xyzw, d_out->mask, nullptr, cuda_stream);
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    if (d_out->format == SLC_RESULT_DEPTH)
        return launch(ctx, d_stack, n_stacks, nullptr, nullptr, nullptr, st, d_out->depth, d_out->mask_bits);
    // POINTS: the maps, then one chained-scan launch over the whole batch
    rc = launch(ctx, d_stack, n_stacks, d_out->xyzw, d_out->mask, nullptr, st);
    if (rc != SLC_OK) return rc;
    if (!slc::compact_supported(ctx->kp.W, ctx->kp.H, d_out->mask, d_out->xyzw))
        return fail(ctx, SLC_ERR_INVALID_ARG, "SLC_RESULT_POINTS needs 16-byte aligned xyzw and 8-byte aligned mask");
    rc = ensure_cstate(ctx, n_stacks, st);
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, slc::launch_compact(ctx->kp.W, ctx->kp.H, n_stacks, d_out->order, d_out->xyzw, d_out->mask, d_out->points,
                                      (long long)d_out->point_stride, d_out->mask_bits, (long long)bits_bytes(ctx),
                                      reinterpret_cast<unsigned long long*>(d_out->n_points),
                                      static_cast<unsigned long long*>(ctx->d_cstate), ++ctx->cstate_epoch, st));
    ctx->launches += (n_stacks + 65534) / 65535;
    return SLC_OK;
}

int slc_submit_host_ex(slc_context* ctx, int32_t slot, const uint8_t* h_stack, int32_t n_stacks, const slc_result* h_out)
{
    int rc = check_ready(ctx, h_stack, h_out, h_out, n_stacks);
    if (rc != SLC_OK) return rc;
    rc = check_result(ctx, h_out, false);
    if (rc != SLC_OK) return rc;
    if (h_out->format == SLC_RESULT_XYZW) return slc_submit_host(ctx, slot, h_stack, n_stacks, h_out->xyzw, h_out->mask);
    if (slot < 0 || slot >= (int)ctx->slots.size()) return fail(ctx, SLC_ERR_INVALID_ARG, "slot %d out of range", slot);
    if (n_stacks < 1 || n_stacks > ctx->cfg.max_batch)
        return fail(ctx, SLC_ERR_INVALID_ARG, "n_stacks %d outside 1..max_batch (%d)", n_stacks, ctx->cfg.max_batch);
    Slot& s = ctx->slots[slot];
    if (s.busy) return fail(ctx, SLC_ERR_STATE, "slot %d still in flight: call slc_wait first", slot);
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    return enqueue_chunk_fmt(ctx, s, h_stack, n_stacks, *h_out);
}

int slc_reconstruct_host_ex(slc_context* ctx, const uint8_t* h_stack, int32_t n_stacks, const slc_result* h_out)
{
    int rc = check_ready(ctx, h_stack, h_out, h_out, n_stacks);
    if (rc != SLC_OK) return rc;
    if (n_stacks == 0) return SLC_OK;
    rc = check_result(ctx, h_out, false);
    if (rc != SLC_OK) return rc;
    if (h_out->format == SLC_RESULT_XYZW)
        return slc_reconstruct_host(ctx, h_stack, n_stacks, h_out->xyzw, h_out->mask, nullptr);
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const int chunk = ctx->cfg.max_batch, S = (int)ctx->slots.size();
    const bool points = h_out->format == SLC_RESULT_POINTS;
    bool overflow = false;
    int slot = 0;
    rc = SLC_OK;
    for (int done = 0; done < n_stacks && rc == SLC_OK; done += chunk) {
        const int n = (n_stacks - done) < chunk ? (n_stacks - done) : chunk;
        Slot& s = ctx->slots[slot];
        // the chunk this slot ran S chunks ago: its counts are known by now, queue its point download
        if (points) rc = finish_points(ctx, s, &overflow);
        if (rc == SLC_OK)
            rc = enqueue_chunk_fmt(ctx, s, h_stack + (size_t)done * stack_bytes(ctx), n, result_at(ctx, *h_out, (size_t)done));
        slot = (slot + 1) % S;
    }
    for (int k = 0; k < S && points; k++) {       // oldest first
        const int rc2 = finish_points(ctx, ctx->slots[(slot + k) % S], &overflow);
        if (rc == SLC_OK) rc = rc2;
    }
    for (Slot& s : ctx->slots) {
        const cudaError_t e = cudaStreamSynchronize(s.stream);
        s.busy = false;
        s.pending.active = false;
        if (e != cudaSuccess && rc == SLC_OK) rc = fail(ctx, SLC_ERR_CUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(e));
    }
    if (rc == SLC_OK && overflow)
        return fail(ctx, SLC_ERR_INVALID_ARG, "a frame set has more valid pixels than point_stride = %lld: its list was cut (n_points holds the full counts)",
                    (long long)h_out->point_stride);
    return rc;
}

/* ---- decoder objects -------------------------------------------------- */
int slc_decode_gray_host(slc_context* ctx, const uint8_t* h_planes, double* h_gray_val, int16_t* h_kbin)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (!h_planes || !h_gray_val) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL buffer");
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t npx = (size_t)ctx->kp.npx;
    const size_t in_bytes = npx * 2 * ctx->kp.G;
    int rc = ensure_scratch(ctx, &ctx->d_scratch_in, &ctx->scratch_in_bytes, in_bytes);
    if (rc == SLC_OK) rc = ensure_scratch(ctx, &ctx->d_scratch_out, &ctx->scratch_out_bytes, npx * sizeof(double));
    if (rc == SLC_OK) rc = ensure_scratch(ctx, &ctx->d_scratch_aux, &ctx->scratch_aux_bytes, npx * sizeof(int16_t));
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch_in, h_planes, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    KParams p = ctx->kp;
    p.lut = ctx->d_lut;
    SLC_CUDA(ctx, slc::launch_decode_gray(p, (const uint8_t*)ctx->d_scratch_in, (double*)ctx->d_scratch_out,
                                          h_kbin ? (int16_t*)ctx->d_scratch_aux : nullptr, ctx->stream));
    ctx->launches++;
    SLC_CUDA(ctx, cudaMemcpyAsync(h_gray_val, ctx->d_scratch_out, npx * sizeof(double), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    if (h_kbin)
        SLC_CUDA(ctx, cudaMemcpyAsync(h_kbin, ctx->d_scratch_aux, npx * sizeof(int16_t), cudaMemcpyDeviceToHost,
                                      ctx->stream));
    SLC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SLC_OK;
}

int slc_decode_phase_host(slc_context* ctx, const uint8_t* h_planes, double* h_phase_pix, uint8_t* h_mod_ok)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (!h_planes || !h_phase_pix) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL buffer");
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t npx = (size_t)ctx->kp.npx;
    const size_t in_bytes = npx * ctx->kp.N;
    int rc = ensure_scratch(ctx, &ctx->d_scratch_in, &ctx->scratch_in_bytes, in_bytes);
    if (rc == SLC_OK) rc = ensure_scratch(ctx, &ctx->d_scratch_out, &ctx->scratch_out_bytes, npx * sizeof(double));
    if (rc == SLC_OK) rc = ensure_scratch(ctx, &ctx->d_scratch_aux, &ctx->scratch_aux_bytes, npx * sizeof(int16_t));
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch_in, h_planes, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    SLC_CUDA(ctx, slc::launch_decode_phase(ctx->kp, (const uint8_t*)ctx->d_scratch_in, (double*)ctx->d_scratch_out,
                                           h_mod_ok ? (uint8_t*)ctx->d_scratch_aux : nullptr, ctx->stream));
    ctx->launches++;
    SLC_CUDA(ctx, cudaMemcpyAsync(h_phase_pix, ctx->d_scratch_out, npx * sizeof(double), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    if (h_mod_ok)
        SLC_CUDA(ctx, cudaMemcpyAsync(h_mod_ok, ctx->d_scratch_aux, npx, cudaMemcpyDeviceToHost, ctx->stream));
    SLC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SLC_OK;
}

int slc_triangulate_host(slc_context* ctx, const double* h_proj_u, float* h_xyzw, uint8_t* h_mask)
{
    int rc = check_ready(ctx, h_proj_u, h_xyzw, h_mask, 1);
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t npx = (size_t)ctx->kp.npx;
    rc = ensure_scratch(ctx, &ctx->d_scratch_in, &ctx->scratch_in_bytes, npx * sizeof(double));
    if (rc == SLC_OK) rc = ensure_scratch(ctx, &ctx->d_scratch_out, &ctx->scratch_out_bytes, npx * 16);
    if (rc == SLC_OK) rc = ensure_scratch(ctx, &ctx->d_scratch_aux, &ctx->scratch_aux_bytes, npx * sizeof(int16_t));
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch_in, h_proj_u, npx * sizeof(double), cudaMemcpyHostToDevice,
                                  ctx->stream));
    SLC_CUDA(ctx, slc::launch_triangulate(ctx->kp, (const double*)ctx->d_scratch_in, (float*)ctx->d_scratch_out,
                                          (uint8_t*)ctx->d_scratch_aux, ctx->stream));
    ctx->launches++;
    SLC_CUDA(ctx, cudaMemcpyAsync(h_xyzw, ctx->d_scratch_out, npx * 16, cudaMemcpyDeviceToHost, ctx->stream));
    SLC_CUDA(ctx, cudaMemcpyAsync(h_mask, ctx->d_scratch_aux, npx, cudaMemcpyDeviceToHost, ctx->stream));
    SLC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SLC_OK;
}

/* ---- [EXT] over-determined (u, v) triangulation ---------------------------- */
int slc_triangulate_uv_device(slc_context* ctx, const double* d_proj_u, const double* d_proj_v, float* d_xyzw,
                              uint8_t* d_mask, void* cuda_stream)
{
    int rc = check_ready(ctx, d_proj_u, d_xyzw, d_mask, 1);
    if (rc != SLC_OK) return rc;
    if (!d_proj_v) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL ProjectorV plane");
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    SLC_CUDA(ctx, slc::launch_triangulate_uv(ctx->kp, d_proj_u, d_proj_v, d_xyzw, d_mask, st));
    ctx->launches++;
    return SLC_OK;
}

int slc_triangulate_uv_host(slc_context* ctx, const double* h_proj_u, const double* h_proj_v, float* h_xyzw,
                            uint8_t* h_mask)
{
    int rc = check_ready(ctx, h_proj_u, h_xyzw, h_mask, 1);
    if (rc != SLC_OK) return rc;
    if (!h_proj_v) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL ProjectorV plane");
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t npx = (size_t)ctx->kp.npx;
    rc = ensure_scratch(ctx, &ctx->d_scratch_in, &ctx->scratch_in_bytes, 2 * npx * sizeof(double));
    if (rc == SLC_OK) rc = ensure_scratch(ctx, &ctx->d_scratch_out, &ctx->scratch_out_bytes, npx * 16);
    if (rc == SLC_OK) rc = ensure_scratch(ctx, &ctx->d_scratch_aux, &ctx->scratch_aux_bytes, npx * sizeof(int16_t));
    if (rc != SLC_OK) return rc;
    double* d_u = static_cast<double*>(ctx->d_scratch_in);
    double* d_v = d_u + npx;
    SLC_CUDA(ctx, cudaMemcpyAsync(d_u, h_proj_u, npx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SLC_CUDA(ctx, cudaMemcpyAsync(d_v, h_proj_v, npx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    rc = slc_triangulate_uv_device(ctx, d_u, d_v, (float*)ctx->d_scratch_out, (uint8_t*)ctx->d_scratch_aux, ctx->stream);
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, cudaMemcpyAsync(h_xyzw, ctx->d_scratch_out, npx * 16, cudaMemcpyDeviceToHost, ctx->stream));
    SLC_CUDA(ctx, cudaMemcpyAsync(h_mask, ctx->d_scratch_aux, npx, cudaMemcpyDeviceToHost, ctx->stream));
    SLC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SLC_OK;
}

/* ---- dynamic frames ---------------------------------------------------- */
int slc_dyna_track_device(slc_context* ctx, const uint8_t* d_frames, int32_t n_frames, int32_t window,
                          const double* d_u0, float* d_xyzw, uint8_t* d_mask, float* d_delta_z,
                          const slc_dyna_parity* d_parity, void* cuda_stream)
{
    // a single frame writes no map (only its strips): the outputs may be NULL then
    int rc = check_ready(ctx, d_frames, n_frames > 1 ? static_cast<const void*>(d_xyzw) : d_frames,
                         n_frames > 1 ? static_cast<const void*>(d_mask) : d_frames, n_frames);
    if (rc != SLC_OK) return rc;
    if (!d_u0) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL ProjectorU[0]");
    if (n_frames < 1 || n_frames > 65535) return fail(ctx, SLC_ERR_INVALID_ARG, "n_frames %d outside 1..65535", n_frames);
    if (window < 3 || window > 33 || (window & 1) == 0)
        return fail(ctx, SLC_ERR_INVALID_ARG, "window %d must be odd and in 3..33 (RECO_WINDOW_SIZE)", window);
    if (ctx->kp.W <= window || ctx->kp.H <= window)
        return fail(ctx, SLC_ERR_INVALID_ARG, "camera %dx%d smaller than the %d-pixel window", ctx->kp.W, ctx->kp.H, window);
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    const size_t npx = (size_t)ctx->kp.npx;
    signed char* strips = d_parity && d_parity->strips ? reinterpret_cast<signed char*>(d_parity->strips) : nullptr;
    if (!strips) {
        rc = ensure_scratch(ctx, &ctx->d_strips, &ctx->strips_bytes, 2 * npx * (size_t)n_frames);
        if (rc != SLC_OK) return rc;
        strips = static_cast<signed char*>(ctx->d_strips);
    }
    SLC_CUDA(ctx, slc::launch_strip_regression(d_frames, n_frames, ctx->kp.W, ctx->kp.H, window, strips, st));
    ctx->launches++;
    if (n_frames > 1 && slc::dyna_fused_supported(ctx->kp.W, strips)) {
        SLC_CUDA(ctx, slc::launch_dyna_fused(ctx->kp, strips, n_frames, d_u0, d_xyzw, d_mask, d_delta_z,
                                             d_parity ? d_parity->delta_p : nullptr,
                                             d_parity ? d_parity->proj_u : nullptr, nullptr, ctx->sm_count, st));
        ctx->launches++;
    } else if (n_frames > 1) {
        rc = ensure_scratch(ctx, &ctx->d_dsums, &ctx->dsums_bytes, 2 * npx * (size_t)(n_frames - 1));
        if (rc != SLC_OK) return rc;
        unsigned short* sums = static_cast<unsigned short*>(ctx->d_dsums);
        SLC_CUDA(ctx, slc::launch_delta_sum(strips, n_frames, ctx->kp.W, ctx->kp.H, sums, st));
        ctx->launches++;
        SLC_CUDA(ctx, slc::launch_dyna_track(ctx->kp, sums, n_frames, d_u0, d_xyzw, d_mask, d_delta_z,
                                             d_parity ? d_parity->delta_p : nullptr,
                                             d_parity ? d_parity->proj_u : nullptr, nullptr, st));
        ctx->launches++;
    }
    return SLC_OK;
}

int slc_dyna_track_host(slc_context* ctx, const uint8_t* h_frames, int32_t n_frames, int32_t window,
                        const double* h_u0, float* h_xyzw, uint8_t* h_mask, float* h_delta_z,
                        const slc_dyna_parity* h_parity)
{
    int rc = check_ready(ctx, h_frames, n_frames > 1 ? static_cast<const void*>(h_xyzw) : h_frames,
                         n_frames > 1 ? static_cast<const void*>(h_mask) : h_frames, n_frames);
    if (rc != SLC_OK) return rc;
    if (!h_u0 || n_frames < 1) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL ProjectorU[0] or n_frames < 1");
    // everything slc_dyna_track_device would reject is rejected here, before a byte is uploaded
    if (n_frames > 65535) return fail(ctx, SLC_ERR_INVALID_ARG, "n_frames %d outside 1..65535", n_frames);
    if (window < 3 || window > 33 || (window & 1) == 0)
        return fail(ctx, SLC_ERR_INVALID_ARG, "window %d must be odd and in 3..33 (RECO_WINDOW_SIZE)", window);
    if (ctx->kp.W <= window || ctx->kp.H <= window)
        return fail(ctx, SLC_ERR_INVALID_ARG, "camera %dx%d smaller than the %d-pixel window", ctx->kp.W, ctx->kp.H, window);
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t npx = (size_t)ctx->kp.npx;
    const size_t nf = (size_t)n_frames, no = nf - 1;
    // one staging block: frames | u0 | xyzw | mask | deltaZ | strips | deltaP | projU
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_fr = take(nf * npx), o_u0 = take(npx * 8), o_xyzw = take(no * npx * 16), o_mask = take(no * npx),
                 o_dz = take(h_delta_z ? no * npx * 4 : 0),
                 o_st = take(nf * npx * 2),
                 o_dp = take(h_parity && h_parity->delta_p ? no * npx * 4 : 0),
                 o_pu = take(h_parity && h_parity->proj_u ? no * npx * 8 : 0);
    rc = ensure_scratch(ctx, &ctx->d_dyna, &ctx->dyna_bytes, off);
    if (rc != SLC_OK) return rc;
    char* base = static_cast<char*>(ctx->d_dyna);
    cudaStream_t st = ctx->stream;
    SLC_CUDA(ctx, cudaMemcpyAsync(base + o_fr, h_frames, nf * npx, cudaMemcpyHostToDevice, st));
    SLC_CUDA(ctx, cudaMemcpyAsync(base + o_u0, h_u0, npx * 8, cudaMemcpyHostToDevice, st));
    slc_dyna_parity dpar{};
    dpar.strips = reinterpret_cast<int8_t*>(base + o_st);
    dpar.delta_p = (h_parity && h_parity->delta_p) ? reinterpret_cast<float*>(base + o_dp) : nullptr;
    dpar.proj_u = (h_parity && h_parity->proj_u) ? reinterpret_cast<double*>(base + o_pu) : nullptr;
    rc = slc_dyna_track_device(ctx, reinterpret_cast<uint8_t*>(base + o_fr), n_frames, window,
                               reinterpret_cast<double*>(base + o_u0), reinterpret_cast<float*>(base + o_xyzw),
                               reinterpret_cast<uint8_t*>(base + o_mask),
                               h_delta_z ? reinterpret_cast<float*>(base + o_dz) : nullptr, &dpar, st);
    if (rc != SLC_OK) return rc;
    if (no > 0) {
        SLC_CUDA(ctx, cudaMemcpyAsync(h_xyzw, base + o_xyzw, no * npx * 16, cudaMemcpyDeviceToHost, st));
        SLC_CUDA(ctx, cudaMemcpyAsync(h_mask, base + o_mask, no * npx, cudaMemcpyDeviceToHost, st));
        if (h_delta_z) SLC_CUDA(ctx, cudaMemcpyAsync(h_delta_z, base + o_dz, no * npx * 4, cudaMemcpyDeviceToHost, st));
        if (dpar.delta_p)
            SLC_CUDA(ctx, cudaMemcpyAsync(h_parity->delta_p, dpar.delta_p, no * npx * 4, cudaMemcpyDeviceToHost, st));
        if (dpar.proj_u)
            SLC_CUDA(ctx, cudaMemcpyAsync(h_parity->proj_u, dpar.proj_u, no * npx * 8, cudaMemcpyDeviceToHost, st));
    }
    if (h_parity && h_parity->strips)
        SLC_CUDA(ctx, cudaMemcpyAsync(h_parity->strips, dpar.strips, nf * npx * 2, cudaMemcpyDeviceToHost, st));
    SLC_CUDA(ctx, cudaStreamSynchronize(st));
    return SLC_OK;
}

int slc_dyna_track_host_ex(slc_context* ctx, const uint8_t* h_frames, int32_t n_frames, int32_t window,
                           const double* h_u0, const slc_result* h_out)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (!h_out) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL slc_result");
    if (h_out->format == SLC_RESULT_XYZW)
        return slc_dyna_track_host(ctx, h_frames, n_frames, window, h_u0, h_out->xyzw, h_out->mask, nullptr, nullptr);
    int rc = check_ready(ctx, h_frames, h_frames, h_frames, n_frames);
    if (rc != SLC_OK) return rc;
    if (!h_u0 || n_frames < 1 || n_frames > 65535) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL ProjectorU[0] or n_frames outside 1..65535");
    if (window < 3 || window > 33 || (window & 1) == 0)
        return fail(ctx, SLC_ERR_INVALID_ARG, "window %d must be odd and in 3..33 (RECO_WINDOW_SIZE)", window);
    if (ctx->kp.W <= window || ctx->kp.H <= window)
        return fail(ctx, SLC_ERR_INVALID_ARG, "camera %dx%d smaller than the %d-pixel window", ctx->kp.W, ctx->kp.H, window);
    if (n_frames == 1) return SLC_OK;                         // a single frame has no dynamic map
    rc = check_result(ctx, h_out, false);
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t npx = (size_t)ctx->kp.npx, nf = (size_t)n_frames, no = nf - 1, bb = bits_bytes(ctx);
    const bool points = h_out->format == SLC_RESULT_POINTS;
    // one staging block: frames | u0 | xyzw | mask | strips | depth or points | bits | counts
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_fr = take(nf * npx), o_u0 = take(npx * 8), o_xyzw = take(no * npx * 16), o_mask = take(no * npx),
                 o_st = take(nf * npx * 2), o_val = take(points ? no * npx * 12 : no * npx * 4), o_bits = take(no * bb + 4),
                 o_cnt = take(no * sizeof(unsigned long long));
    rc = ensure_scratch(ctx, &ctx->d_dyna, &ctx->dyna_bytes, off);
    if (rc != SLC_OK) return rc;
    char* base = static_cast<char*>(ctx->d_dyna);
    cudaStream_t st = ctx->stream;
    SLC_CUDA(ctx, cudaMemcpyAsync(base + o_fr, h_frames, nf * npx, cudaMemcpyHostToDevice, st));
    SLC_CUDA(ctx, cudaMemcpyAsync(base + o_u0, h_u0, npx * 8, cudaMemcpyHostToDevice, st));
    slc_dyna_parity dpar{};
    dpar.strips = reinterpret_cast<int8_t*>(base + o_st);
    float* d_xyzw = reinterpret_cast<float*>(base + o_xyzw);
    uint8_t* d_mask = reinterpret_cast<uint8_t*>(base + o_mask);
    uint8_t* d_bits = reinterpret_cast<uint8_t*>(base + o_bits);
    rc = slc_dyna_track_device(ctx, reinterpret_cast<uint8_t*>(base + o_fr), n_frames, window,
                               reinterpret_cast<double*>(base + o_u0), d_xyzw, d_mask, nullptr, &dpar, st);
    if (rc != SLC_OK) return rc;
    if (!points) {
        float* d_depth = reinterpret_cast<float*>(base + o_val);
        SLC_CUDA(ctx, slc::launch_pack_depth(d_xyzw, d_mask, (long long)npx, (int)no, d_depth, d_bits, (long long)bb, st));
        ctx->launches++;
        SLC_CUDA(ctx, cudaMemcpyAsync(h_out->depth, d_depth, no * npx * 4, cudaMemcpyDeviceToHost, st));
        SLC_CUDA(ctx, cudaMemcpyAsync(h_out->mask_bits, d_bits, no * bb, cudaMemcpyDeviceToHost, st));
        SLC_CUDA(ctx, cudaStreamSynchronize(st));
        return SLC_OK;
    }
    // POINTS: one chained-scan launch over every map of the sequence, the counts first, then exactly 12 * count bytes per map
    float* d_points = reinterpret_cast<float*>(base + o_val);
    unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(base + o_cnt);
    rc = ensure_cstate(ctx, (int)no, st);
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, slc::launch_compact(ctx->kp.W, ctx->kp.H, (int)no, h_out->order, d_xyzw, d_mask, d_points, (long long)npx,
                                      h_out->mask_bits ? d_bits : nullptr, (long long)bb, d_counts,
                                      static_cast<unsigned long long*>(ctx->d_cstate), ++ctx->cstate_epoch, st));
    ctx->launches++;
    SLC_CUDA(ctx, cudaMemcpyAsync(h_out->n_points, d_counts, no * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    if (h_out->mask_bits) SLC_CUDA(ctx, cudaMemcpyAsync(h_out->mask_bits, d_bits, no * bb, cudaMemcpyDeviceToHost, st));
    if (h_out->xyzw) SLC_CUDA(ctx, cudaMemcpyAsync(h_out->xyzw, d_xyzw, no * npx * 16, cudaMemcpyDeviceToHost, st));
    if (h_out->mask) SLC_CUDA(ctx, cudaMemcpyAsync(h_out->mask, d_mask, no * npx, cudaMemcpyDeviceToHost, st));
    SLC_CUDA(ctx, cudaStreamSynchronize(st));                 // n_points is in host memory now
    bool overflow = false;
    for (size_t f = 0; f < no; f++) {
        const int64_t cnt = h_out->n_points[f];
        if (cnt > h_out->point_stride) overflow = true;
        const int64_t m = cnt < h_out->point_stride ? cnt : h_out->point_stride;
        if (m > 0)
            SLC_CUDA(ctx, cudaMemcpyAsync(h_out->points + f * (size_t)h_out->point_stride * 3, d_points + f * npx * 3,
                                          (size_t)m * 12, cudaMemcpyDeviceToHost, st));
    }
    SLC_CUDA(ctx, cudaStreamSynchronize(st));
    if (overflow)
        return fail(ctx, SLC_ERR_INVALID_ARG, "a frame has more valid pixels than point_stride = %lld: its list was cut (n_points holds the full counts)",
                    (long long)h_out->point_stride);
    return SLC_OK;
}

/* ---- input ingest ------------------------------------------------------ */
namespace {

uint32_t rd_u32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint32_t rd_u16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

int ensure_pinned(slc_context* ctx, void** p, size_t* have, size_t want)
{
    if (*have >= want) return SLC_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *have = 0;
    SLC_CUDA(ctx, cudaHostAlloc(p, want, cudaHostAllocDefault));
    *have = want;
    return SLC_OK;
}

int bmp_unpack(slc_context* ctx, const uint8_t* d_pixels, const slc_bmp_info* info, uint8_t* d_plane, cudaStream_t st)
{
    SLC_CUDA(ctx, slc::launch_bmp_unpack(d_pixels, info->width, info->height, info->bits_per_pixel, info->top_down,
                                         info->row_stride, info->palette_is_identity, info->gray, d_plane, st));
    ctx->launches++;
    return SLC_OK;
}

}  // namespace

int slc_bmp_parse(const void* file_bytes, int64_t n_bytes, slc_bmp_info* info)
{
    if (!file_bytes || !info || n_bytes < 54) return SLC_ERR_INVALID_ARG;
    const uint8_t* f = static_cast<const uint8_t*>(file_bytes);
    if (f[0] != 'B' || f[1] != 'M') return SLC_ERR_INVALID_ARG;
    const uint32_t off = rd_u32(f + 10), hdr = rd_u32(f + 14);
    if (hdr < 40) return SLC_ERR_INVALID_ARG;                      /* BITMAPCOREHEADER files are not camera output */
    const int32_t w = (int32_t)rd_u32(f + 18), hs = (int32_t)rd_u32(f + 22);
    const uint32_t bpp = rd_u16(f + 28), comp = rd_u32(f + 30);
    uint32_t used = rd_u32(f + 46);
    if (w <= 0 || hs == 0 || hs == INT32_MIN) return SLC_ERR_INVALID_ARG;
    if (comp != 0 || (bpp != 8 && bpp != 24 && bpp != 32)) return SLC_ERR_INVALID_ARG;
    std::memset(info, 0, sizeof(*info));
    info->width = w;
    info->height = hs < 0 ? -hs : hs;
    info->bits_per_pixel = (int32_t)bpp;
    info->top_down = hs < 0;
    const int64_t stride = (((int64_t)w * (bpp / 8)) + 3) & ~(int64_t)3;
    if (stride > INT32_MAX) return SLC_ERR_INVALID_ARG;
    info->row_stride = (int32_t)stride;
    info->pixel_offset = off;
    if ((int64_t)off + stride * info->height > n_bytes) return SLC_ERR_INVALID_ARG;   /* truncated */
    info->palette_is_identity = 1;
    if (bpp == 8) {
        if (used == 0 || used > 256) used = 256;
        if (14 + (int64_t)hdr + 4 * (int64_t)used > n_bytes) return SLC_ERR_INVALID_ARG;
        const uint8_t* pal = f + 14 + hdr;                          /* B G R 0 */
        for (uint32_t i = 0; i < used; i++) {
            const uint32_t g = (pal[4 * i] * 1868u + pal[4 * i + 1] * 9617u + pal[4 * i + 2] * 4899u + 8192u) >> 14;
            info->gray[i] = (uint8_t)g;
        }
        for (int i = 0; i < 256; i++)
            if (info->gray[i] != i) info->palette_is_identity = 0;
    }
    return SLC_OK;
}

int slc_bmp_unpack_device(slc_context* ctx, const uint8_t* d_pixels, const slc_bmp_info* info, uint8_t* d_plane,
                          void* cuda_stream)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (!d_pixels || !info || !d_plane) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL argument");
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    return bmp_unpack(ctx, d_pixels, info, d_plane, st);
}

int slc_bmp_unpack_batch_device(slc_context* ctx, const uint8_t* const* d_pixels, const slc_bmp_info* infos,
                                int32_t n_files, uint8_t* d_stack, void* cuda_stream)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (!d_pixels || !infos || !d_stack || n_files < 0) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL argument or n_files < 0");
    if (n_files == 0) return SLC_OK;
    std::vector<slc::BmpPlane> planes;
    try {
        planes.resize((size_t)n_files);
    } catch (const std::exception&) {             // no exception leaves the C ABI
        return fail(ctx, SLC_ERR_OUT_OF_MEMORY, "host allocation for %d file descriptors failed", n_files);
    }
    const size_t npx = (size_t)ctx->kp.npx;
    for (int i = 0; i < n_files; i++) {
        const slc_bmp_info& f = infos[i];
        if (!d_pixels[i]) return fail(ctx, SLC_ERR_INVALID_ARG, "file %d: NULL pixel array", i);
        if (f.width != ctx->kp.W || f.height != ctx->kp.H)
            return fail(ctx, SLC_ERR_INVALID_ARG, "file %d is %dx%d, the context's camera is %dx%d", i, f.width, f.height,
                        ctx->kp.W, ctx->kp.H);
        if ((f.bits_per_pixel != 8 && f.bits_per_pixel != 24 && f.bits_per_pixel != 32) ||
            (int64_t)f.row_stride < ((int64_t)f.width * f.bits_per_pixel + 7) / 8)
            return fail(ctx, SLC_ERR_INVALID_ARG, "file %d: not a parsed slc_bmp_info (bpp %d, stride %d)", i,
                        f.bits_per_pixel, f.row_stride);
        slc::BmpPlane& a = planes[(size_t)i];
        a.px = d_pixels[i];
        a.out = d_stack + (size_t)i * npx;
        a.width = f.width; a.height = f.height; a.bpp = f.bits_per_pixel; a.top_down = f.top_down;
        a.row_stride = f.row_stride; a.identity = f.palette_is_identity;
        a.wide = 0; a.pad_ = 0;
        std::memcpy(a.gray, f.gray, 256);
    }
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    SLC_CUDA(ctx, slc::launch_bmp_unpack_batch(planes.data(), n_files, st));
    ctx->launches += (n_files + slc::kBmpBatchMax - 1) / slc::kBmpBatchMax;
    return SLC_OK;
}

int slc_bmp_decode_host(slc_context* ctx, const void* h_file_bytes, int64_t n_bytes, uint8_t* h_plane,
                        int32_t expect_width, int32_t expect_height)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (!h_file_bytes || !h_plane) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL buffer");
    slc_bmp_info info;
    if (slc_bmp_parse(h_file_bytes, n_bytes, &info) != SLC_OK)
        return fail(ctx, SLC_ERR_INVALID_ARG, "not an uncompressed 8/24/32-bit BMP (or truncated)");
    if ((expect_width > 0 && info.width != expect_width) || (expect_height > 0 && info.height != expect_height))
        return fail(ctx, SLC_ERR_INVALID_ARG, "BMP is %dx%d, expected %dx%d", info.width, info.height, expect_width,
                    expect_height);
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t raw = (size_t)info.row_stride * info.height, plane = (size_t)info.width * info.height;
    int rc = ensure_scratch(ctx, &ctx->d_bmp[0], &ctx->bmp_bytes[0], raw + plane);
    if (rc != SLC_OK) return rc;
    uint8_t* d_raw = static_cast<uint8_t*>(ctx->d_bmp[0]);
    uint8_t* d_plane = d_raw + raw;
    SLC_CUDA(ctx, cudaMemcpyAsync(d_raw, static_cast<const uint8_t*>(h_file_bytes) + info.pixel_offset, raw,
                                  cudaMemcpyHostToDevice, ctx->stream));
    rc = bmp_unpack(ctx, d_raw, &info, d_plane, ctx->stream);
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, cudaMemcpyAsync(h_plane, d_plane, plane, cudaMemcpyDeviceToHost, ctx->stream));
    SLC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SLC_OK;
}

static int load_bmp_planes_impl(slc_context* ctx, const char* const* paths, int32_t n_files, uint8_t* d_stack);

int slc_load_bmp_planes(slc_context* ctx, const char* const* paths, int32_t n_files, uint8_t* d_stack)
{
    try {
        return load_bmp_planes_impl(ctx, paths, n_files, d_stack);
    } catch (const std::exception& ex) {          // no exception leaves the C ABI (host allocations before any thread starts)
        return fail(ctx, SLC_ERR_OUT_OF_MEMORY, "slc_load_bmp_planes: %s", ex.what());
    }
}

static int load_bmp_planes_impl(slc_context* ctx, const char* const* paths, int32_t n_files, uint8_t* d_stack)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (!paths || !d_stack || n_files < 0) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL argument or n_files < 0");
    if (n_files == 0) return SLC_OK;
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t npx = (size_t)ctx->kp.npx;
    // file sizes first: one staging size for every slot
    std::vector<size_t> sizes((size_t)n_files);
    size_t max_size = 0;
    for (int i = 0; i < n_files; i++) {
        if (!paths[i]) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL path for file %d", i);
        FILE* f = std::fopen(paths[i], "rb");
        if (!f) return fail(ctx, SLC_ERR_INVALID_ARG, "imread error: %s", paths[i]);     /* CSensorV.cpp:122-129 */
        std::fseek(f, 0, SEEK_END);
        const long sz = std::ftell(f);
        std::fclose(f);
        if (sz < 54) return fail(ctx, SLC_ERR_INVALID_ARG, "imread error: %s", paths[i]);
        sizes[(size_t)i] = (size_t)sz;
        if ((size_t)sz > max_size) max_size = (size_t)sz;
    }
    // kBmpSlots staging slots (pinned host + device + event); file i uses slot i % kBmpSlots.  Reader
    // threads fill slots ahead of the uploads; this thread issues upload + unpack in file order.
    const int K = kBmpSlots;
    for (int k = 0; k < K; k++) {
        int rc = ensure_pinned(ctx, &ctx->h_bmp[k], &ctx->h_bmp_bytes[k], max_size);
        if (rc == SLC_OK) rc = ensure_scratch(ctx, &ctx->d_bmp[k], &ctx->bmp_bytes[k], max_size + npx);
        if (rc != SLC_OK) return rc;
        if (!ctx->bmp_done[k]) SLC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->bmp_done[k], cudaEventDisableTiming));
    }
    const int R = std::max(1, std::min({4, n_files, (int)std::thread::hardware_concurrency()}));
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int> state((size_t)n_files, 0);      // 0 pending, 1 read ok, -1 read failed
    std::vector<int> slot_free_for((size_t)K);       // index of the file allowed to use slot k next
    for (int k = 0; k < K; k++) slot_free_for[(size_t)k] = k;
    bool abort_all = false;
    auto reader = [&](int r) {
        for (int i = r; i < n_files; i += R) {
            const int k = i % K;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return abort_all || slot_free_for[(size_t)k] == i; });
                if (abort_all) return;
            }
            bool ok = false;
            if (FILE* f = std::fopen(paths[i], "rb")) {
                ok = std::fread(ctx->h_bmp[k], 1, sizes[(size_t)i], f) == sizes[(size_t)i];
                std::fclose(f);
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                state[(size_t)i] = ok ? 1 : -1;
            }
            cv.notify_all();
        }
    };
    std::vector<std::thread> threads;
    auto stop = [&] {
        { std::lock_guard<std::mutex> lk(mu); abort_all = true; }
        cv.notify_all();
        for (auto& th : threads)
            if (th.joinable()) th.join();
    };
    try {
        threads.reserve((size_t)R);
        for (int r = 0; r < R; r++) threads.emplace_back(reader, r);
    } catch (const std::exception& ex) {          // no exception leaves the C ABI; readers already started are joined
        stop();
        return fail(ctx, SLC_ERR_STATE, "could not start the reader threads: %s", ex.what());
    }
    int rc = SLC_OK;
    for (int i = 0; i < n_files && rc == SLC_OK; i++) {
        const int k = i % K;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return state[(size_t)i] != 0; });
        }
        slc_bmp_info info;
        if (state[(size_t)i] < 0 || slc_bmp_parse(ctx->h_bmp[k], (int64_t)sizes[(size_t)i], &info) != SLC_OK) {
            rc = fail(ctx, SLC_ERR_INVALID_ARG, "imread error (not an uncompressed 8/24/32-bit BMP): %s", paths[i]);
            break;
        }
        if (info.width != ctx->kp.W || info.height != ctx->kp.H) {
            rc = fail(ctx, SLC_ERR_INVALID_ARG, "%s is %dx%d, the context is %dx%d", paths[i], info.width, info.height,
                      ctx->kp.W, ctx->kp.H);
            break;
        }
        const size_t raw = (size_t)info.row_stride * info.height;
        cudaStream_t fs = ctx->stream;
        cudaError_t e = cudaMemcpyAsync(ctx->d_bmp[k], static_cast<uint8_t*>(ctx->h_bmp[k]) + info.pixel_offset, raw,
                                        cudaMemcpyHostToDevice, fs);
        if (e == cudaSuccess)
            e = slc::launch_bmp_unpack(static_cast<const uint8_t*>(ctx->d_bmp[k]), info.width, info.height, info.bits_per_pixel,
                                       info.top_down, info.row_stride, info.palette_is_identity, info.gray,
                                       d_stack + (size_t)i * npx, fs);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->bmp_done[k], fs);
        if (e != cudaSuccess) { rc = fail(ctx, SLC_ERR_CUDA, "BMP upload / unpack failed: %s", cudaGetErrorString(e)); break; }
        ctx->launches++;
        // hand the slot of an older file back to the readers once its upload + unpack are done
        const int j = i - K / 2;
        if (j >= 0 && j + K < n_files) {
            e = cudaEventSynchronize(ctx->bmp_done[j % K]);
            if (e != cudaSuccess) { rc = fail(ctx, SLC_ERR_CUDA, "cudaEventSynchronize failed: %s", cudaGetErrorString(e)); break; }
            { std::lock_guard<std::mutex> lk(mu); slot_free_for[(size_t)(j % K)] = j + K; }
            cv.notify_all();
        }
    }
    stop();
    if (rc != SLC_OK) { cudaStreamSynchronize(ctx->stream); return rc; }
    SLC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SLC_OK;
}

/* ---- point-cloud output ------------------------------------------------ */
namespace {

// block sums / totals / look-back words of the point-cloud kernels; zeroed when (re)allocated
int ensure_pc_scratch(slc_context* ctx, cudaStream_t st)
{
    const size_t want = slc::pointcloud_scratch_bytes(ctx->kp.npx);
    if (ctx->pc_scratch_bytes >= want) return SLC_OK;
    int rc = ensure_scratch(ctx, &ctx->d_pc_scratch, &ctx->pc_scratch_bytes, want);
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, cudaMemsetAsync(ctx->d_pc_scratch, 0, want, st));
    ctx->pc_epoch = 0;
    return SLC_OK;
}

int pointcloud_run(slc_context* ctx, int mode, int order, uint32_t flags, const double* d_proj_u, const float* d_xyzw,
                   const uint8_t* d_mask, void* d_out, int64_t capacity_bytes, int64_t* bytes, int64_t* records,
                   cudaStream_t st)
{
    if (!ctx->calibrated) return fail(ctx, SLC_ERR_NOT_INITIALISED, "calibration not set (CCalculation.cpp:176-181)");
    if (capacity_bytes < 0 || !d_out) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL output or negative capacity");
    if (order != SLC_ORDER_ROW_MAJOR && order != SLC_ORDER_REFERENCE)
        return fail(ctx, SLC_ERR_INVALID_ARG, "order %d is neither SLC_ORDER_ROW_MAJOR nor SLC_ORDER_REFERENCE", order);
    if (reinterpret_cast<uintptr_t>(d_out) & 15) return fail(ctx, SLC_ERR_INVALID_ARG, "output buffer must be 16-byte aligned");
    int rc = ensure_pc_scratch(ctx, st);
    if (rc != SLC_OK) return rc;
    const unsigned long long* d_totals = nullptr;
    if (mode == 0) ctx->pc_epoch = ctx->pc_epoch % 3u + 1u;
    SLC_CUDA(ctx, slc::launch_pointcloud(ctx->kp, mode, order, flags, d_proj_u, d_xyzw, d_mask, d_out,
                                         (unsigned long long)capacity_bytes, ctx->d_pc_scratch, ctx->pc_epoch, &d_totals, st));
    ctx->launches += mode == 0 ? 1 : 2;
    // the counts come back through a pinned word pair (a pageable destination costs a staged copy per call)
    if (!ctx->h_pc_totals) SLC_CUDA(ctx, cudaHostAlloc(&ctx->h_pc_totals, 2 * sizeof(unsigned long long), cudaHostAllocDefault));
    unsigned long long* totals = static_cast<unsigned long long*>(ctx->h_pc_totals);
    SLC_CUDA(ctx, cudaMemcpyAsync(totals, d_totals, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    SLC_CUDA(ctx, cudaStreamSynchronize(st));
    if (bytes) *bytes = (int64_t)totals[0];
    if (records) *records = (int64_t)totals[1];
    if ((int64_t)totals[0] > capacity_bytes)
        return fail(ctx, SLC_ERR_INVALID_ARG, "point cloud needs %lld bytes, buffer holds %lld", (long long)totals[0],
                    (long long)capacity_bytes);
    return SLC_OK;
}

}  // namespace

namespace {

// float3 of the valid pixels of ONE map: the chained-scan kernel of slc_compact.cu (one launch) where the
// geometry allows it, else the two-pass record emitter.
int compact_run(slc_context* ctx, int order, const float* d_xyzw, const uint8_t* d_mask, float* d_xyz,
                int64_t capacity_points, int64_t* points, cudaStream_t st)
{
    if (!slc::compact_supported(ctx->kp.W, ctx->kp.H, d_mask, d_xyzw))
        return pointcloud_run(ctx, 1, order, 0u, nullptr, d_xyzw, d_mask, d_xyz, capacity_points * 12, nullptr, points, st);
    if (capacity_points < 0 || !d_xyz) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL output or negative capacity");
    if (order != SLC_ORDER_ROW_MAJOR && order != SLC_ORDER_REFERENCE)
        return fail(ctx, SLC_ERR_INVALID_ARG, "order %d is neither SLC_ORDER_ROW_MAJOR nor SLC_ORDER_REFERENCE", order);
    int rc = ensure_cstate(ctx, 1, st);
    if (rc == SLC_OK) rc = ensure_pc_scratch(ctx, st);
    if (rc != SLC_OK) return rc;
    unsigned long long* d_count = static_cast<unsigned long long*>(ctx->d_pc_scratch);
    SLC_CUDA(ctx, slc::launch_compact(ctx->kp.W, ctx->kp.H, 1, order, d_xyzw, d_mask, d_xyz, (long long)capacity_points, nullptr, 0,
                                      d_count, static_cast<unsigned long long*>(ctx->d_cstate), ++ctx->cstate_epoch, st));
    ctx->launches++;
    if (!ctx->h_pc_totals) SLC_CUDA(ctx, cudaHostAlloc(&ctx->h_pc_totals, 2 * sizeof(unsigned long long), cudaHostAllocDefault));
    unsigned long long* totals = static_cast<unsigned long long*>(ctx->h_pc_totals);
    SLC_CUDA(ctx, cudaMemcpyAsync(totals, d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    SLC_CUDA(ctx, cudaStreamSynchronize(st));
    if (points) *points = (int64_t)totals[0];
    if ((int64_t)totals[0] > capacity_points)
        return fail(ctx, SLC_ERR_INVALID_ARG, "point cloud needs room for %lld points, buffer holds %lld", (long long)totals[0],
                    (long long)capacity_points);
    return SLC_OK;
}

}  // namespace

int slc_pointcloud_text_device(slc_context* ctx, const double* d_proj_u, uint32_t flags, char* d_text,
                               int64_t capacity, int64_t* bytes, int64_t* points, void* cuda_stream)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (!d_proj_u) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL ProjectorU plane");
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    return pointcloud_run(ctx, 0, SLC_ORDER_REFERENCE, flags, d_proj_u, nullptr, nullptr, d_text, capacity, bytes,
                          points, st);
}

int slc_pointcloud_text_host(slc_context* ctx, const double* h_proj_u, uint32_t flags, char* h_text,
                             int64_t capacity, int64_t* bytes, int64_t* points)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (!h_proj_u || !h_text || capacity < 0) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL buffer or negative capacity");
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t npx = (size_t)ctx->kp.npx;
    const size_t worst = npx * 43 + 16;
    const size_t cap = (size_t)capacity < worst ? (size_t)capacity : worst;
    int rc = ensure_scratch(ctx, &ctx->d_pc_in, &ctx->pc_in_bytes, npx * sizeof(double));
    if (rc == SLC_OK) rc = ensure_scratch(ctx, &ctx->d_pc_out, &ctx->pc_out_bytes, cap + 16);
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, cudaMemcpyAsync(ctx->d_pc_in, h_proj_u, npx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    int64_t nb = 0;
    rc = pointcloud_run(ctx, 0, SLC_ORDER_REFERENCE, flags, static_cast<const double*>(ctx->d_pc_in), nullptr, nullptr,
                        ctx->d_pc_out, (int64_t)cap, &nb, points, ctx->stream);
    if (bytes) *bytes = nb;
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, cudaMemcpyAsync(h_text, ctx->d_pc_out, (size_t)nb, cudaMemcpyDeviceToHost, ctx->stream));
    SLC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SLC_OK;
}

int slc_pointcloud_compact_device(slc_context* ctx, const float* d_xyzw, const uint8_t* d_mask, int32_t order,
                                  float* d_xyz, int64_t capacity_points, int64_t* points, void* cuda_stream)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (!d_xyzw || !d_mask) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL map");
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    return compact_run(ctx, order, d_xyzw, d_mask, d_xyz, capacity_points, points, st);
}

int slc_compact_points_device(slc_context* ctx, const float* d_xyzw, const uint8_t* d_mask, int32_t n_maps, int32_t order,
                              float* d_points, int64_t point_stride, uint8_t* d_mask_bits, int64_t* d_n_points,
                              void* cuda_stream)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (n_maps < 0) return fail(ctx, SLC_ERR_INVALID_ARG, "n_maps < 0");
    if (n_maps == 0) return SLC_OK;
    if (!d_xyzw || !d_mask || !d_points || !d_n_points || point_stride < 0)
        return fail(ctx, SLC_ERR_INVALID_ARG, "NULL buffer or negative point_stride");
    if (order != SLC_ORDER_ROW_MAJOR && order != SLC_ORDER_REFERENCE)
        return fail(ctx, SLC_ERR_INVALID_ARG, "order %d is neither SLC_ORDER_ROW_MAJOR nor SLC_ORDER_REFERENCE", order);
    if (!slc::compact_supported(ctx->kp.W, ctx->kp.H, d_mask, d_xyzw))
        return fail(ctx, SLC_ERR_INVALID_ARG, "batched compaction needs a camera width that is a multiple of 8, a height below 65536, "
                                              "16-byte aligned xyzw and 8-byte aligned mask");
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    int rc = ensure_cstate(ctx, n_maps, st);
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, slc::launch_compact(ctx->kp.W, ctx->kp.H, n_maps, order, d_xyzw, d_mask, d_points, (long long)point_stride,
                                      d_mask_bits, (long long)bits_bytes(ctx), reinterpret_cast<unsigned long long*>(d_n_points),
                                      static_cast<unsigned long long*>(ctx->d_cstate), ++ctx->cstate_epoch, st));
    ctx->launches += (n_maps + 65534) / 65535;
    return SLC_OK;
}

int slc_pointcloud_compact_host(slc_context* ctx, const float* h_xyzw, const uint8_t* h_mask, int32_t order,
                                float* h_xyz, int64_t capacity_points, int64_t* points)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (!h_xyzw || !h_mask || !h_xyz || capacity_points < 0)
        return fail(ctx, SLC_ERR_INVALID_ARG, "NULL buffer or negative capacity");
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t npx = (size_t)ctx->kp.npx;
    const size_t cap_pts = (size_t)capacity_points < npx ? (size_t)capacity_points : npx;
    int rc = ensure_scratch(ctx, &ctx->d_pc_in, &ctx->pc_in_bytes, npx * 17);
    if (rc == SLC_OK) rc = ensure_scratch(ctx, &ctx->d_pc_out, &ctx->pc_out_bytes, cap_pts * 12 + 16);
    if (rc != SLC_OK) return rc;
    float* d_xyzw = static_cast<float*>(ctx->d_pc_in);
    uint8_t* d_mask = static_cast<uint8_t*>(ctx->d_pc_in) + npx * 16;
    SLC_CUDA(ctx, cudaMemcpyAsync(d_xyzw, h_xyzw, npx * 16, cudaMemcpyHostToDevice, ctx->stream));
    SLC_CUDA(ctx, cudaMemcpyAsync(d_mask, h_mask, npx, cudaMemcpyHostToDevice, ctx->stream));
    int64_t n = 0;
    rc = compact_run(ctx, order, d_xyzw, d_mask, static_cast<float*>(ctx->d_pc_out), (int64_t)cap_pts, &n, ctx->stream);
    if (points) *points = n;
    if (rc != SLC_OK) return rc;
    SLC_CUDA(ctx, cudaMemcpyAsync(h_xyz, ctx->d_pc_out, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->stream));
    SLC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SLC_OK;
}

int slc_format_g6_host(slc_context* ctx, const double* h_values, int64_t n, uint32_t flags, char* h_text16,
                       uint8_t* h_len)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (!h_values || !h_text16 || !h_len || n < 0) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL buffer or n < 0");
    if (n == 0) return SLC_OK;
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    int rc = ensure_scratch(ctx, &ctx->d_pc_in, &ctx->pc_in_bytes, (size_t)n * 8);
    if (rc == SLC_OK) rc = ensure_scratch(ctx, &ctx->d_pc_out, &ctx->pc_out_bytes, (size_t)n * 17);
    if (rc != SLC_OK) return rc;
    char* d_text = static_cast<char*>(ctx->d_pc_out);
    uint8_t* d_len = static_cast<uint8_t*>(ctx->d_pc_out) + (size_t)n * 16;
    SLC_CUDA(ctx, cudaMemcpyAsync(ctx->d_pc_in, h_values, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    SLC_CUDA(ctx, slc::launch_format_g6(static_cast<const double*>(ctx->d_pc_in), n, flags, d_text, d_len, ctx->stream));
    ctx->launches++;
    SLC_CUDA(ctx, cudaMemcpyAsync(h_text16, d_text, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    SLC_CUDA(ctx, cudaMemcpyAsync(h_len, d_len, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    SLC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SLC_OK;
}

int slc_eval_phase_host(slc_context* ctx, const float* h_sin, const float* h_cos, int64_t n, float* h_deg,
                        float* h_pix)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (!h_sin || !h_cos || !h_deg || !h_pix || n < 0) return fail(ctx, SLC_ERR_INVALID_ARG, "NULL buffer or n < 0");
    if (n == 0) return SLC_OK;
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t bytes = (size_t)n * sizeof(float);
    int rc = ensure_scratch(ctx, &ctx->d_scratch_in, &ctx->scratch_in_bytes, 2 * bytes);
    if (rc == SLC_OK) rc = ensure_scratch(ctx, &ctx->d_scratch_out, &ctx->scratch_out_bytes, 2 * bytes);
    if (rc != SLC_OK) return rc;
    float* d_s = (float*)ctx->d_scratch_in;
    float* d_c = d_s + n;
    float* d_deg = (float*)ctx->d_scratch_out;
    float* d_pix = d_deg + n;
    SLC_CUDA(ctx, cudaMemcpyAsync(d_s, h_sin, bytes, cudaMemcpyHostToDevice, ctx->stream));
    SLC_CUDA(ctx, cudaMemcpyAsync(d_c, h_cos, bytes, cudaMemcpyHostToDevice, ctx->stream));
    SLC_CUDA(ctx, slc::launch_eval_phase(d_s, d_c, n, ctx->kp.Tf, d_deg, d_pix, ctx->stream));
    ctx->launches++;
    SLC_CUDA(ctx, cudaMemcpyAsync(h_deg, d_deg, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SLC_CUDA(ctx, cudaMemcpyAsync(h_pix, d_pix, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SLC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SLC_OK;
}

/* ---- measurement ------------------------------------------------------ */
int slc_time_reconstruct_device(slc_context* ctx, const uint8_t* d_stack, int32_t n_stacks, float* d_xyzw,
                                uint8_t* d_mask, int32_t iters, float* ms_per_launch)
{
    int rc = check_ready(ctx, d_stack, d_xyzw, d_mask, n_stacks);
    if (rc != SLC_OK) return rc;
    if (iters < 1 || !ms_per_launch || n_stacks < 1)
        return fail(ctx, SLC_ERR_INVALID_ARG, "bad iters / n_stacks / output pointer");
    SLC_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    cudaEvent_t e0, e1;
    SLC_CUDA(ctx, cudaEventCreate(&e0));
    SLC_CUDA(ctx, cudaEventCreate(&e1));
    SLC_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    for (int i = 0; i < iters; i++) {
        rc = launch(ctx, d_stack, n_stacks, d_xyzw, d_mask, nullptr, ctx->stream);
        if (rc != SLC_OK) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
    }
    SLC_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    SLC_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    SLC_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_per_launch = ms / (float)iters;
    return SLC_OK;
}

int64_t slc_launch_count(const slc_context* ctx) { return ctx ? ctx->launches : 0; }

int slc_set_pixels_per_thread(slc_context* ctx, int32_t pxt)
{
    if (!ctx) return SLC_ERR_INVALID_ARG;
    if (pxt != 0 && pxt != 4 && pxt != 8 && pxt != 16) return fail(ctx, SLC_ERR_INVALID_ARG, "pixels per thread must be 0, 4, 8 or 16");
    ctx->pxt_override = pxt;
    for (auto& a : ctx->plans)
        for (auto& b : a)
            for (auto& pl : b) pl = slc::LaunchPlan{};    // chosen again on the next launch
    return SLC_OK;
}

}  // extern "C"
