// slc_kernels.cu -- sm_100a kernels of the DynaFrame first-frame path.
//
// reconstruct_vec_kernel is the product: ONE fused pass per pixel that replaces
// the reference's six full-image loops (CDecodeGray.cpp:150-204,
// CDecodePhase.cpp:48-80, CCalculation.cpp:562-589, 672-708, 756-771) and
// writes no intermediates to HBM.
//
// Mapping.  A thread owns PXT consecutive pixels of one image row (PXT = 16, 8
// or 4): one PXT-byte load per u8 plane, so a warp reads 32*PXT contiguous
// bytes of every plane (whole 128 B lines).  Results are float4 per pixel; to
// keep the stores whole-line too, each warp transposes its 32*PXT float4
// through a private shared-memory tile padded to an odd row stride (PXT+1
// slots: bank-conflict-free both ways) and writes 512 contiguous bytes per
// store instruction.  Loads are
// ld.global.nc.L1::no_allocate, stores st.global.cs: every
// byte is touched exactly once.  Nothing here is a contraction, so tensor
// cores / TMEM are not used; the bound is HBM bandwidth.
#include "slc_kernels.h"

#include <cstdio>

namespace slc {

namespace {

constexpr int kBlock = 256;

template <int PXT> struct VecLoad;
template <> struct VecLoad<16> {
    static __device__ __forceinline__ void load(const uint8_t* p, uint32_t (&w)[4]) {
        const uint4 v = ld_stream_u4(p); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    }
};
template <> struct VecLoad<8> {
    static __device__ __forceinline__ void load(const uint8_t* p, uint32_t (&w)[2]) {
        const uint2 v = ld_stream_u2(p); w[0] = v.x; w[1] = v.y;
    }
};
template <> struct VecLoad<4> {
    static __device__ __forceinline__ void load(const uint8_t* p, uint32_t (&w)[1]) {
        uint32_t v;
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
        w[0] = v;
    }
};

// Transpose tile: each warp owns 32 rows of PXT float4 padded to PXT+1 slots.  With
// the odd stride both the owner-major write (8 lanes, same i) and the pixel-major
// read (8 consecutive pixels of one owner) touch 8 distinct 16-byte bank groups,
// and every address is "per-thread base + compile-time immediate".
template <int PXT> struct Tile { static constexpr int kStride = PXT + 1; static constexpr int kSlots = 32 * (PXT + 1); };

// What a pixel leaves in its tile slot: the whole float4 (OUT 0: xyzw + u8 mask), or z alone
// (OUT 1: SLC_RESULT_DEPTH, z plane + one validity bit per pixel).  A pixel awaiting the f64 re-solve
// parks (gint, pix) in the first two words either way.
template <int OUT> struct Slot;
template <> struct Slot<0> {
    using type = float4;
    static __device__ __forceinline__ float4 result(const float4& r) { return r; }
};
template <> struct Slot<1> {
    using type = float2;
    static __device__ __forceinline__ float2 result(const float4& r) { return make_float2(r.z, 0.f); }
};

// MODE 0: no parity planes, reflected Gray code, f32 z with f64 guard band, no modulation test
// MODE 1: as 0 with the [EXT] modulation test
// MODE 2: everything decided at run time (parity planes, custom LUT, SLC_FLAG_Z_FP64, modulation)
template <int MODE, bool Z64>
__device__ __forceinline__ float solve_pixel(const KParams& p, const RowConst& rc, int kbin, float s, float c,
                                             float uf, PixelResult& r)
{
    const float pix = phase_to_pix(fast_atan2_deg(s, c), p.Tf);
    bool mod_ok = true;
    if (MODE == 1 || (MODE == 2 && p.use_mod)) mod_ok = __fadd_rn(__fmul_rn(s, s), __fmul_rn(c, c)) >= p.thr2;
    // [EXT] a pixel the modulation test rejects has no projector column: w = 0 and proj_u = 0 (ZERO_W, r.has_u)
    unwrap_and_triangulate<MODE == 2, Z64, MODE != 0>(p, rc, kbin, pix, mod_ok, uf, r);
    return pix;
}

// valid bits b0..b3 -> bytes 0/1
__device__ __forceinline__ uint32_t spread_bits4(uint32_t b) { return ((b & 0xFu) * 0x00204081u) & 0x01010101u; }

// The fused kernel.  G_T / N_T > 0 bake the digit and step counts in (all plane
// loops unroll and every load is issued up front); 0 means "read it from p".
template <int PXT, int G_T, int N_T, int MODE, int OUT>
__global__ void __launch_bounds__(kBlock)
reconstruct_vec_kernel(const __grid_constant__ KParams p)
{
    constexpr int NW = PXT / 4;  // 32-bit words (4 pixels each) per thread
    constexpr int kStride = Tile<PXT>::kStride;
    using SlotT = typename Slot<OUT>::type;
    extern __shared__ float4 s_tile[];

    const int G = G_T > 0 ? G_T : p.G;
    const int N = N_T > 0 ? N_T : p.N;
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    SlotT* tile = reinterpret_cast<SlotT*>(s_tile) + warp_in_block * Tile<PXT>::kSlots;

    // grid.y = stack, grid.x covers the stack's pixel groups: a warp tile never straddles stacks
    const int stack = blockIdx.y;
    const unsigned n_groups = (unsigned)p.n_groups;
    const unsigned g = blockIdx.x * kBlock + threadIdx.x;       // group inside the stack
    const bool active = g < n_groups;
    const long long out0 = (long long)stack * p.npx;            // first output pixel of the stack
    uint32_t validbits = 0;

    if (active) {
        const unsigned off = g * PXT;             // first pixel of the group inside the stack
        int v, u0;
        split_row_col(p, off, v, u0);
        // block-uniform plane base + 32-bit per-thread offset
        const uint8_t* sbase = p.stack + (long long)stack * p.P * p.npx;
        auto plane = [&](int q) { return sbase + (long long)q * p.npx + off; };

        // ---- a3 + a4: Gray pairs -> per-pixel code bits (CDecodeGray.cpp:155-199) ----
        uint32_t lo[NW], hi[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) { lo[w] = 0; hi[w] = 0; }
        auto gray_bit = [&](int b) {
            uint32_t pa[NW], pb[NW];
            VecLoad<PXT>::load(plane(2 * b), pa);
            VecLoad<PXT>::load(plane(2 * b + 1), pb);
#pragma unroll
            for (int w = 0; w < NW; w++) {
                const uint32_t t = gt_u8x4_msb(pa[w], pb[w]);
                if (b < 8) lo[w] |= (t >> (7 - b)) & (0x01010101u << b);
                else       hi[w] |= (t >> (15 - b)) & (0x01010101u << (b - 8));
            }
        };
        if constexpr (G_T > 0) {
#pragma unroll
            for (int b = 0; b < G_T; b++) gray_bit(b);
        } else {
#pragma unroll 4
            for (int b = 0; b < G; b++) gray_bit(b);     // generic depth: four pairs of loads in flight
        }
        // gray2bin (CDecodeGray.cpp:120-125,200): arithmetic for the reflected code
        const bool use_lut = (MODE == 2) && (p.lut != nullptr);
        uint32_t bl[NW], bh[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) {
            if (use_lut) { bl[w] = lo[w]; bh[w] = hi[w]; }
            else {
                bh[w] = (G > 8) ? prefix_xor_u8x4(hi[w]) : 0u;
                bl[w] = prefix_xor_u8x4(lo[w]) ^ ((bh[w] & 0x01010101u) * 0xFFu);
            }
        }

        // ---- a6: phase images -> (sin, cos) sums (CDecodePhase.cpp:59-65) ----
        float sv[PXT], cv[PXT];
        if (N == 4) {
            uint32_t q0[NW], q1[NW], q2[NW], q3[NW];
            VecLoad<PXT>::load(plane(2 * G), q0);
            VecLoad<PXT>::load(plane(2 * G + 1), q1);
            VecLoad<PXT>::load(plane(2 * G + 2), q2);
            VecLoad<PXT>::load(plane(2 * G + 3), q3);
#pragma unroll
            for (int w = 0; w < NW; w++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    // (float(I0) - float(I2)) / 2, exact: both operands are 2^22 + I/2
                    sv[4 * w + j] = __fsub_rn(u8_magic_half(q0[w], j, p.magic_half), u8_magic_half(q2[w], j, p.magic_half));
                    cv[4 * w + j] = __fsub_rn(u8_magic_half(q1[w], j, p.magic_half), u8_magic_half(q3[w], j, p.magic_half));
                }
        } else if ((N & 1) == 0) {
            // [EXT] even N: d_k = I_k - I_{k+N/2}; S = sum d_k cos(2 pi k/N), Cc = sum d_k sin(2 pi k/N)
#pragma unroll
            for (int i = 0; i < PXT; i++) { sv[i] = 0.f; cv[i] = 0.f; }
            const int half = N >> 1;
            auto phase_pair = [&](int k) {
                uint32_t qa[NW], qb[NW];
                VecLoad<PXT>::load(plane(2 * G + k), qa);
                VecLoad<PXT>::load(plane(2 * G + k + half), qb);
                const float ck = p.ck[k], sk = p.sk[k];
#pragma unroll
                for (int w = 0; w < NW; w++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float d = __fsub_rn(u8_magic(qa[w], j, p.magic_one), u8_magic(qb[w], j, p.magic_one));
                        sv[4 * w + j] = __fmaf_rn(d, ck, sv[4 * w + j]);
                        cv[4 * w + j] = __fmaf_rn(d, sk, cv[4 * w + j]);
                    }
            };
            if constexpr (N_T > 0) {
#pragma unroll
                for (int k = 0; k < N_T / 2; k++) phase_pair(k);
            } else {
#pragma unroll 2
                for (int k = 0; k < half; k++) phase_pair(k);
            }
        } else {
            // [EXT] odd N: plain sums
#pragma unroll
            for (int i = 0; i < PXT; i++) { sv[i] = 0.f; cv[i] = 0.f; }
#pragma unroll 3
            for (int k = 0; k < N; k++) {
                uint32_t qa[NW];
                VecLoad<PXT>::load(plane(2 * G + k), qa);
                const float ck = p.ck[k], sk = p.sk[k];
#pragma unroll
                for (int w = 0; w < NW; w++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float gk = __fsub_rn(u8_magic(qa[w], j, p.magic_one), 8388608.f);
                        sv[4 * w + j] = __fmaf_rn(gk, ck, sv[4 * w + j]);
                        cv[4 * w + j] = __fmaf_rn(gk, sk, cv[4 * w + j]);
                    }
            }
        }

        // ---- per pixel: arctan, offset, unwrap, f32 triangulation (branch free) ----
        const RowConst rc = make_row_const(p, v);
        const float u0f = (float)u0;
        SlotT* trow = tile + lane * kStride;
        uint32_t slowbits = 0;
        const bool z64 = (MODE == 2) && (p.z_fp64 != 0);
#pragma unroll
        for (int w = 0; w < NW; w++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = 4 * w + j;
                int kbin;
                if (G > 8 || use_lut)
                    kbin = (int)__byte_perm(bl[w], bh[w], 0x4400u + (uint32_t)j + ((uint32_t)(4 + j) << 4)) & 0xFFFF;
                else
                    kbin = (int)((bl[w] >> (8 * j)) & 0xFFu);
                if (use_lut) kbin = (int)__ldg(p.lut + kbin);
                else if (G == 16) kbin = (int)(short)kbin;   // m_gray2bin is `short` (CDecodeGray.h:23)
                const float uf = u0f + (float)i;
                PixelResult r;
                float pix;
                if (z64) pix = solve_pixel<MODE, true>(p, rc, kbin, sv[i], cv[i], uf, r);
                else pix = solve_pixel<MODE, false>(p, rc, kbin, sv[i], cv[i], uf, r);
                // pixels awaiting the f64 re-solve park (gint, pix) in their tile slot
                if constexpr (OUT == 0) trow[i] = make_float4(r.need64 ? r.gint : r.x, r.need64 ? pix : r.y, r.z, r.w);
                else trow[i] = make_float2(r.need64 ? r.gint : r.z, pix);
                validbits |= (r.valid ? 1u : 0u) << i;
                slowbits |= (r.need64 ? 1u : 0u) << i;
                if (MODE == 2) {
                    const long long o = out0 + off + i;
                    if (p.kbin) p.kbin[o] = (int16_t)kbin;
                    if (p.corr) p.corr[o] = (int8_t)r.corr;
                    if (p.phase_pix) p.phase_pix[o] = pix;
                    if (p.proj_u) p.proj_u[o] = r.has_u ? __dadd_rn((double)r.gint, (double)pix) : 0.0;
                }
            }
        }
        // ---- rare: pixels whose validity f32 cannot decide are re-solved in f64 ----
        while (slowbits != 0u) {
            const int i = __ffs((int)slowbits) - 1;
            slowbits &= slowbits - 1u;
            const SlotT t = trow[i];
            int ok;
            trow[i] = Slot<OUT>::result(resolve_f64(p, t.x, t.y, u0 + i, v, &ok));
            validbits = (validbits & ~(1u << i)) | ((uint32_t)ok << i);
        }
        if constexpr (OUT == 0) {
            // validity mask: PXT contiguous bytes per thread, 32*PXT per warp
            uint8_t* mptr = p.mask + out0 + off;
            if constexpr (PXT == 16)
                st_stream_u4(mptr, make_uint4(spread_bits4(validbits), spread_bits4(validbits >> 4),
                                              spread_bits4(validbits >> 8), spread_bits4(validbits >> 12)));
            else if constexpr (PXT == 8)
                st_stream_u2(mptr, make_uint2(spread_bits4(validbits), spread_bits4(validbits >> 4)));
            else
                st_stream_u1(mptr, spread_bits4(validbits));
        }
    }

    if constexpr (OUT == 1) {
        // ---- SLC_RESULT_DEPTH: one validity bit per pixel (32 * PXT / 8 contiguous bytes per warp) and
        //      the z plane, 512 contiguous bytes per store instruction ----
        uint8_t* bptr = p.mask_bits + (long long)stack * p.bits_stride + (size_t)(g * (PXT / 4)) / 2;
        if constexpr (PXT == 4) {
            // two threads share a byte (n_groups is even: npx % 16 == 0)
            const uint32_t hi_nib = __shfl_down_sync(0xFFFFFFFFu, validbits, 1);
            if (active && (lane & 1) == 0) *bptr = (uint8_t)(validbits | (hi_nib << 4));
        } else if constexpr (PXT == 8) {
            if (active) *bptr = (uint8_t)validbits;
        } else {
            if (active) *reinterpret_cast<uint16_t*>(bptr) = (uint16_t)validbits;
        }
        __syncwarp();
        const unsigned px0 = (blockIdx.x * kBlock + (threadIdx.x & ~31u)) * PXT;
        const unsigned stack_px = n_groups * PXT;
        float* zout = p.depth + out0 + px0;
#pragma unroll
        for (int it = 0; it < PXT / 4; it++) {
            const int q = it * 32 + lane;                 // quad of pixels 4q .. 4q+3 of the warp's 32*PXT
            const SlotT* src = tile + ((4 * q) / PXT) * kStride + ((4 * q) % PXT);
            if (px0 + 4u * q < stack_px)
                st_stream_f4(reinterpret_cast<float4*>(zout + 4 * q), make_float4(src[0].x, src[1].x, src[2].x, src[3].x));
        }
        return;
    }

    if constexpr (OUT == 0) {
        // ---- transposed, fully coalesced float4 stores: 512 contiguous bytes per instruction ----
        __syncwarp();
        const unsigned px0 = (blockIdx.x * kBlock + (threadIdx.x & ~31u)) * PXT;  // warp's first pixel in the stack
        const unsigned stack_px = n_groups * PXT;
        float4* outp = p.xyzw + out0 + px0 + lane;
        // pixel it*32 + lane belongs to owner (it*32 + lane) / PXT, element (it*32 + lane) % PXT
        const float4* tsrc = tile + (lane / PXT) * kStride + (lane % PXT);
        if (px0 + 32u * PXT <= stack_px) {
#pragma unroll
            for (int it = 0; it < PXT; it++)
                st_stream_f4(outp + it * 32, tsrc[(it * 32 / PXT) * kStride]);
        } else {
#pragma unroll
            for (int it = 0; it < PXT; it++)
                if (px0 + it * 32 + lane < stack_px) st_stream_f4(outp + it * 32, tsrc[(it * 32 / PXT) * kStride]);
        }
    }
}

// One pixel per thread: any width, any alignment.  Same per-pixel math.
template <bool PARITY>
__global__ void __launch_bounds__(kBlock)
reconstruct_scalar_kernel(const __grid_constant__ KParams p)
{
    const long long total = p.npx * (long long)p.n_stacks;
    const long long idx = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (idx >= total) return;
    const int stack = (int)(idx / p.npx);
    const long long off = idx - (long long)stack * p.npx;
    const int v = (int)(off / p.W);
    const int u = (int)(off - (long long)v * p.W);
    const uint8_t* base = p.stack + (long long)stack * p.P * p.npx + off;
    unsigned code = 0;
    for (int b = 0; b < p.G; b++) {
        const unsigned a = base[(long long)(2 * b) * p.npx];
        const unsigned bb = base[(long long)(2 * b + 1) * p.npx];
        code |= (a > bb ? 1u : 0u) << b;   // sat_u8(a-b) > 0  (CDecodeGray.cpp:159,168)
    }
    int kbin;
    if (p.lut) {
        kbin = (int)__ldg(p.lut + code);
    } else {
        unsigned g = code;
        g ^= g >> 1; g ^= g >> 2; g ^= g >> 4; g ^= g >> 8;
        kbin = (int)(short)(g & 0xFFFFu);   // m_gray2bin is `short` (CDecodeGray.h:23)
    }
    const uint8_t* ph = base + (long long)(2 * p.G) * p.npx;
    float s, c;
    if (p.N == 4) {
        s = __fmul_rn(__fsub_rn((float)ph[0], (float)ph[2 * p.npx]), 0.5f);
        c = __fmul_rn(__fsub_rn((float)ph[p.npx], (float)ph[3 * p.npx]), 0.5f);
    } else if ((p.N & 1) == 0) {
        s = 0.f; c = 0.f;
        const int half = p.N >> 1;
        for (int k = 0; k < half; k++) {
            const float d = __fsub_rn((float)ph[(long long)k * p.npx], (float)ph[(long long)(k + half) * p.npx]);
            s = __fmaf_rn(d, p.ck[k], s);
            c = __fmaf_rn(d, p.sk[k], c);
        }
    } else {
        s = 0.f; c = 0.f;
        for (int k = 0; k < p.N; k++) {
            const float gk = (float)ph[(long long)k * p.npx];
            s = __fmaf_rn(gk, p.ck[k], s);
            c = __fmaf_rn(gk, p.sk[k], c);
        }
    }
    // the scalar kernel keeps the compiler's full-range div.rn (it doubles as a cross-check
    // of the vector kernel's safe-range divisions)
    const float pix = phase_to_pix<true>(fast_atan2_deg<true>(s, c), p.Tf);
    bool mod_ok = true;
    if (p.use_mod) mod_ok = __fadd_rn(__fmul_rn(s, s), __fmul_rn(c, c)) >= p.thr2;
    PixelResult r;
    const RowConst rc = make_row_const(p, v);
    if (p.z_fp64) unwrap_and_triangulate<true, true, true>(p, rc, kbin, pix, mod_ok, (float)u, r);
    else unwrap_and_triangulate<true, false, true>(p, rc, kbin, pix, mod_ok, (float)u, r);
    float4 outv = make_float4(r.x, r.y, r.z, r.w);                  // [EXT] rejected by the modulation test: w = 0
    int ok = r.valid ? 1 : 0;
    if (r.need64) outv = resolve_f64(p, r.gint, pix, u, v, &ok);
    if (p.depth) {
        // SLC_RESULT_DEPTH on the any-geometry path: the bit plane was zeroed by the launcher
        p.depth[idx] = outv.z;
        if (ok) {
            const long long bit = (long long)stack * p.bits_stride * 8 + off;
            atomicOr(reinterpret_cast<unsigned*>(p.mask_bits) + (bit >> 5), 1u << (bit & 31));
        }
    } else {
        p.xyzw[idx] = outv;
        p.mask[idx] = (uint8_t)ok;
    }
    if (PARITY) {
        if (p.kbin) p.kbin[idx] = (int16_t)kbin;
        if (p.corr) p.corr[idx] = (int8_t)r.corr;
        if (p.phase_pix) p.phase_pix[idx] = pix;
        if (p.proj_u) p.proj_u[idx] = r.has_u ? __dadd_rn((double)r.gint, (double)pix) : 0.0;
    }
}

// CDecodeGray::Decode on its own: 2G planes -> f64 plane (+ optional kbin).
__global__ void __launch_bounds__(kBlock)
decode_gray_kernel(const __grid_constant__ KParams p, const uint8_t* __restrict__ planes,
                   double* __restrict__ gray_val, int16_t* __restrict__ kbin_out)
{
    const long long idx = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (idx >= p.npx) return;
    unsigned code = 0;
    for (int b = 0; b < p.G; b++) {
        const unsigned a = planes[(long long)(2 * b) * p.npx + idx];
        const unsigned bb = planes[(long long)(2 * b + 1) * p.npx + idx];
        code |= (a > bb ? 1u : 0u) << b;
    }
    int kbin;
    if (p.lut) {
        kbin = (int)__ldg(p.lut + code);
    } else {
        unsigned g = code;
        g ^= g >> 1; g ^= g >> 2; g ^= g >> 4; g ^= g >> 8;
        kbin = (int)(short)(g & 0xFFFFu);   // m_gray2bin is `short` (CDecodeGray.h:23)
    }
    // (double)m_gray2bin[grayCode] * pixPeriod   (CDecodeGray.cpp:200)
    gray_val[idx] = __dmul_rn((double)kbin, (double)p.gp);
    if (kbin_out) kbin_out[idx] = (int16_t)kbin;
}

// CDecodePhase::Decode on its own: N planes -> f64 plane of offsets in (0, T].
__global__ void __launch_bounds__(kBlock)
decode_phase_kernel(const __grid_constant__ KParams p, const uint8_t* __restrict__ ph,
                    double* __restrict__ phase_pix, uint8_t* __restrict__ mod_out)
{
    const long long idx = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (idx >= p.npx) return;
    float s, c;
    if (p.N == 4) {
        s = __fmul_rn(__fsub_rn((float)ph[idx], (float)ph[2 * p.npx + idx]), 0.5f);
        c = __fmul_rn(__fsub_rn((float)ph[p.npx + idx], (float)ph[3 * p.npx + idx]), 0.5f);
    } else if ((p.N & 1) == 0) {
        s = 0.f; c = 0.f;
        const int half = p.N >> 1;
        for (int k = 0; k < half; k++) {
            const float d = __fsub_rn((float)ph[(long long)k * p.npx + idx],
                                      (float)ph[(long long)(k + half) * p.npx + idx]);
            s = __fmaf_rn(d, p.ck[k], s);
            c = __fmaf_rn(d, p.sk[k], c);
        }
    } else {
        s = 0.f; c = 0.f;
        for (int k = 0; k < p.N; k++) {
            const float gk = (float)ph[(long long)k * p.npx + idx];
            s = __fmaf_rn(gk, p.ck[k], s);
            c = __fmaf_rn(gk, p.sk[k], c);
        }
    }
    phase_pix[idx] = (double)phase_to_pix<true>(fast_atan2_deg<true>(s, c), p.Tf);
    if (mod_out) {
        const bool ok = !p.use_mod || (__fadd_rn(__fmul_rn(s, s), __fmul_rn(c, c)) >= p.thr2);
        mod_out[idx] = ok ? 1 : 0;
    }
}

// CCalculation::FillCoordinate(i) for an arbitrary f64 ProjectorU plane
// (CCalculation.cpp:666-771).  U is arbitrary here (the dynamic mode adds f32
// deltas to it), so z is always solved in f64.
__global__ void __launch_bounds__(kBlock)
triangulate_kernel(const __grid_constant__ KParams p, const double* __restrict__ proj_u,
                   float4* __restrict__ xyzw, uint8_t* __restrict__ mask)
{
    const long long idx = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (idx >= p.npx) return;
    const int v = (int)(idx / p.W);
    const int u = (int)(idx - (long long)v * p.W);
    const double U = proj_u[idx];
    float z = 0.f;
    int valid = 0;
    if (U != 0.0) {
        const double zd = z_exact(p, U, u, v);
        valid = !((zd < p.fov_min) || (zd > p.fov_max));
        z = valid ? (float)zd : 0.f;
    }
    const float x = z * fmaf(p.rx1, (float)u, p.rx0);
    const float y = z * fmaf(p.ry1, (float)v, p.ry0);
    xyzw[idx] = make_float4(x, y, z, (float)U);
    mask[idx] = (uint8_t)valid;
}

// [EXT] SURVEY 8(f) rank 4: both projector coordinates decoded (vertical and horizontal patterns;
// the reference only has the m_vertical switch of its Gray decoder, CDecodeGray.cpp:182-185, and never
// uses row 1 of P).  Each coordinate gives one linear equation in z, built exactly like the
// reference's (CCalculation.cpp:159-164,686-687) from its row of P:
//     (c0 - c2*U) z = B*U - A          (c2*V ... E for the projector row)
// and z is their least-squares solution (a1*b1 + a2*b2) / (a1^2 + a2^2).  All in f64, one operation per
// intrinsic, so the oracle's C restatement is matched bit for bit.
__global__ void __launch_bounds__(kBlock)
triangulate_uv_kernel(const __grid_constant__ KParams p, const double* __restrict__ proj_u,
                      const double* __restrict__ proj_v, float4* __restrict__ xyzw, uint8_t* __restrict__ mask)
{
    const long long idx = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (idx >= p.npx) return;
    const int v = (int)(idx / p.W);
    const int u = (int)(idx - (long long)v * p.W);
    const double U = proj_u[idx], V = proj_v[idx];
    float x = 0.f, y = 0.f, z = 0.f;
    int valid = 0;
    if (U != 0.0 && V != 0.0) {
        const double du = __dmul_rn(__dsub_rn((double)u, p.cu), p.fv);
        const double dv = __dmul_rn(__dsub_rn((double)v, p.cv), p.fu);
        const double c0 = __dadd_rn(__dadd_rn(__dmul_rn(du, p.P00), __dmul_rn(dv, p.P01)), p.fufvP02);
        const double c1 = __dadd_rn(__dadd_rn(__dmul_rn(du, p.P10), __dmul_rn(dv, p.P11)), p.fufvP12);
        const double c2 = __dadd_rn(__dadd_rn(__dmul_rn(du, p.P20), __dmul_rn(dv, p.P21)), p.fufvP22);
        const double a1 = __dsub_rn(c0, __dmul_rn(c2, U)), b1 = __dsub_rn(__dmul_rn(p.B, U), p.A);
        const double a2 = __dsub_rn(c1, __dmul_rn(c2, V)), b2 = __dsub_rn(__dmul_rn(p.B, V), p.E);
        const double num = __dadd_rn(__dmul_rn(a1, b1), __dmul_rn(a2, b2));
        const double den = __dadd_rn(__dmul_rn(a1, a1), __dmul_rn(a2, a2));
        const double zd = __ddiv_rn(num, den);
        valid = !((zd < p.fov_min) || (zd > p.fov_max));
        if (valid) {
            z = (float)zd;
            x = (float)__ddiv_rn(__dmul_rn(zd, __dsub_rn((double)u, p.cu)), p.fu);
            y = (float)__ddiv_rn(__dmul_rn(zd, __dsub_rn((double)v, p.cv)), p.fv);
        }
    }
    xyzw[idx] = make_float4(x, y, z, (float)U);
    mask[idx] = (uint8_t)valid;
}

// Parity hook: the vector kernel's arctangent + offset arithmetic (safe-range
// divisions) on caller-supplied (sin, cos) sums.
__global__ void __launch_bounds__(kBlock)
eval_phase_kernel(const float* __restrict__ s, const float* __restrict__ c, long long n, float Tf,
                  float* __restrict__ deg, float* __restrict__ pix)
{
    const long long idx = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (idx >= n) return;
    const float a = fast_atan2_deg<false>(s[idx], c[idx]);
    deg[idx] = a;
    pix[idx] = phase_to_pix<false>(a, Tf);
}

// ---------------------------------------------------------------------------
using VecKernel = void (*)(const KParams);

struct VecEntry { int G, N, pxt, mode, out; VecKernel fn; };

// MODE 0 / 1 (production) come in both output layouts, MODE 2 (parity planes, custom table, Z_FP64)
// in the xyzw + mask layout only.
#define SLC_VEC(PXT, G, N) \
    { G, N, PXT, 0, 0, reconstruct_vec_kernel<PXT, G, N, 0, 0> }, \
    { G, N, PXT, 1, 0, reconstruct_vec_kernel<PXT, G, N, 1, 0> }, \
    { G, N, PXT, 2, 0, reconstruct_vec_kernel<PXT, G, N, 2, 0> }, \
    { G, N, PXT, 0, 1, reconstruct_vec_kernel<PXT, G, N, 0, 1> }, \
    { G, N, PXT, 1, 1, reconstruct_vec_kernel<PXT, G, N, 1, 1> }
// tuning-only shapes (slc_set_pixels_per_thread): xyzw + mask layout
#define SLC_VEC_TUNE(PXT, G, N) \
    { G, N, PXT, 0, 0, reconstruct_vec_kernel<PXT, G, N, 0, 0> }, \
    { G, N, PXT, 1, 0, reconstruct_vec_kernel<PXT, G, N, 1, 0> }, \
    { G, N, PXT, 2, 0, reconstruct_vec_kernel<PXT, G, N, 2, 0> }

// Specialised <G, N> instances: the reference default, BASELINE.json's configurations and the
// other 4-step Gray depths (every plane loop unrolled, loads issued ahead of their use); anything
// else runs the generic (0, 0) instance, which is latency bound (0.58-0.65 of the HBM peak measured
// at G = 8, N = 4 before that pair got its own instance, profiles/r01_sweep_geometry.txt).
const VecEntry kVecTable[] = {
    SLC_VEC(8, 6, 4),   SLC_VEC(8, 9, 4),   SLC_VEC(8, 5, 4),   SLC_VEC(8, 10, 4),
    SLC_VEC(8, 8, 8),   SLC_VEC(8, 10, 12),
    SLC_VEC_TUNE(8, 7, 4), SLC_VEC_TUNE(8, 8, 4),
    // Gray depth fixed, any number of phase steps (the Gray planes are most of the loads)
    SLC_VEC_TUNE(8, 5, 0),   SLC_VEC_TUNE(8, 6, 0),   SLC_VEC_TUNE(8, 7, 0),   SLC_VEC_TUNE(8, 8, 0),   SLC_VEC_TUNE(8, 9, 0),
    SLC_VEC_TUNE(8, 10, 0),  SLC_VEC_TUNE(8, 0, 0),
    SLC_VEC_TUNE(16, 9, 4),  SLC_VEC_TUNE(16, 7, 4), SLC_VEC_TUNE(16, 8, 4), SLC_VEC_TUNE(16, 0, 0),
    SLC_VEC(4, 7, 4),   SLC_VEC(4, 8, 4),
    // three-step phase shifting, the usual alternative to four steps, at the usual Gray depths
    SLC_VEC(4, 7, 3),   SLC_VEC(4, 8, 3),   SLC_VEC(4, 9, 3),   SLC_VEC(4, 10, 3),
    // ... and five / six steps (0.75 / 0.79 of the HBM peak through the <G, 0> instances, profiles/r02_sweep_geometry.txt)
    SLC_VEC(4, 7, 5),   SLC_VEC(4, 8, 5),   SLC_VEC(4, 9, 5),   SLC_VEC(4, 10, 5),
    SLC_VEC(4, 7, 6),   SLC_VEC(4, 8, 6),   SLC_VEC(4, 9, 6),   SLC_VEC(4, 10, 6),
    // eight / twelve steps beside the two BASELINE pairs
    SLC_VEC(8, 7, 8),   SLC_VEC(8, 9, 8),   SLC_VEC(8, 10, 8),  SLC_VEC(8, 8, 12),  SLC_VEC(8, 9, 12),
    SLC_VEC_TUNE(4, 9, 4),   SLC_VEC_TUNE(4, 6, 4), SLC_VEC_TUNE(4, 8, 8), SLC_VEC_TUNE(4, 10, 12),
    SLC_VEC(4, 5, 0),   SLC_VEC(4, 6, 0),   SLC_VEC(4, 7, 0),   SLC_VEC(4, 8, 0),   SLC_VEC(4, 9, 0),   SLC_VEC(4, 10, 0),
    // the remaining legal depths (CDecodeGray.cpp:39: 1..16), so that no geometry falls back to run-time loops
    SLC_VEC(4, 1, 0),   SLC_VEC(4, 2, 0),   SLC_VEC(4, 3, 0),   SLC_VEC(4, 4, 0),   SLC_VEC(4, 11, 0),  SLC_VEC(4, 12, 0),
    SLC_VEC(4, 13, 0),  SLC_VEC(4, 14, 0),  SLC_VEC(4, 15, 0),  SLC_VEC(4, 16, 0),
    SLC_VEC(4, 0, 0),
};

const VecEntry* find_vec(int G, int N, int pxt, int mode, int out, bool* specialised)
{
    const VecEntry *generic = nullptr, *gray_only = nullptr;
    for (const VecEntry& e : kVecTable) {
        if (e.pxt != pxt || e.mode != mode || e.out != out) continue;
        if (e.G == G && e.N == N) { *specialised = true; return &e; }
        if (e.G == G && e.N == 0) gray_only = &e;
        if (e.G == 0 && e.N == 0) generic = &e;
    }
    *specialised = false;
    return gray_only ? gray_only : generic;
}

// Pixels per thread: 8 (one 8-byte load per plane, 64 registers, 4 blocks / SM) is the fastest shape
// for most instances; the 4-step G = 7 and G = 8 instances run at 0.87 / 0.83 of the HBM peak with 8 and
// at 0.94 / 0.98 with 4 (40-48 registers, 6 blocks / SM) -- profiles/r01_sweep_geometry.txt.
// The instances without a compile-time step count are also faster with 4 (0.75-0.87 against 0.70-0.76).
int preferred_pxt(int G, int N)
{
    if (N == 4 && (G == 7 || G == 8)) return 4;
    bool exact = false;
    const VecEntry* e = find_vec(G, N, 8, 0, 0, &exact);
    return (e != nullptr && exact) ? 8 : 4;
}

}  // namespace

bool vector_kernel_applicable(const KParams& p, int pxt)
{
    // groups must not straddle rows and every plane base must stay PXT-aligned
    auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool outs = p.depth ? (a16(p.depth) && (reinterpret_cast<uintptr_t>(p.mask_bits) & 1) == 0 && (p.bits_stride & 1) == 0)
                              : (a16(p.mask) && a16(p.xyzw));
    return (p.W % pxt == 0) && (p.npx % 16 == 0) && a16(p.stack) && outs;
}

// Chooses the kernel for (geometry, mode, output layout) ONCE: table scan, dynamic shared-memory
// attribute and register count are paid here, not per launch.
cudaError_t plan_reconstruct(const KParams& geom, int mode, int out, int pxt_override, LaunchPlan* plan)
{
    *plan = LaunchPlan{};
    plan->mode = mode;
    plan->out = out;
    int pxt = pxt_override ? pxt_override : preferred_pxt(geom.G, geom.N);
    while (pxt >= 4 && geom.W % pxt != 0) pxt >>= 1;   // groups must not straddle rows
    cudaFuncAttributes fa;
    if (pxt >= 4 && geom.npx % 16 == 0 && geom.npx <= 0x7fffffffLL && !(mode == 2 && out == 1)) {
        bool spec = false;
        const VecEntry* e = find_vec(geom.G, geom.N, pxt, mode, out, &spec);
        if (e == nullptr && pxt_override) {             // a forced shape that this (mode, layout) does not have
            pxt = preferred_pxt(geom.G, geom.N);
            while (pxt >= 4 && geom.W % pxt != 0) pxt >>= 1;
            e = pxt >= 4 ? find_vec(geom.G, geom.N, pxt, mode, out, &spec) : nullptr;
        }
        if (e != nullptr) {
            const int slot = out == 0 ? (int)sizeof(float4) : (int)sizeof(float2);
            const int smem = (kBlock / 32) * 32 * (pxt + 1) * slot;
            cudaError_t err = cudaFuncSetAttribute(e->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (err == cudaSuccess) err = cudaFuncGetAttributes(&fa, e->fn);
            if (err != cudaSuccess) return err;
            plan->vec = reinterpret_cast<const void*>(e->fn);
            plan->pxt = pxt;
            plan->smem = smem;
            plan->vec_regs = fa.numRegs;
            plan->specialised = spec;
        }
    }
    cudaError_t err = cudaFuncGetAttributes(&fa, mode == 2 ? reconstruct_scalar_kernel<true> : reconstruct_scalar_kernel<false>);
    if (err != cudaSuccess) return err;
    plan->scalar_regs = fa.numRegs;
    plan->valid = true;
    return cudaSuccess;
}

int plan_mode(const KParams& p)
{
    const bool parity = p.kbin || p.corr || p.phase_pix || p.proj_u;
    return (parity || p.lut != nullptr || p.z_fp64) ? 2 : (p.use_mod ? 1 : 0);
}

cudaError_t launch_reconstruct(KParams p, const LaunchPlan& plan, bool force_scalar, cudaStream_t stream,
                               LaunchInfo* info)
{
    // v = off / W by multiplication: exact while off * W < 2^40 (split_row_col)
    p.row_magic = ((unsigned long long)p.npx * (unsigned long long)p.W < (1ull << 40))
                      ? ((1ull << 40) / (unsigned long long)p.W + 1ull) : 0ull;
    const bool vec = !force_scalar && plan.vec != nullptr && vector_kernel_applicable(p, plan.pxt);
    if (info) {
        info->variant = vec ? (plan.specialised ? 0 : 1) : 2;
        info->regs = vec ? plan.vec_regs : plan.scalar_regs;
        info->block = kBlock;
        info->smem = vec ? plan.smem : 0;
        info->pxt = vec ? plan.pxt : 1;
        if (info->query_only) return cudaSuccess;
    }
    if (p.n_stacks <= 0) return cudaSuccess;
    if (vec) {
        VecKernel fn = reinterpret_cast<VecKernel>(const_cast<void*>(plan.vec));
        p.n_groups = p.npx / plan.pxt;
        const long long blocks = (p.n_groups + kBlock - 1) / kBlock;
        if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        // grid.y carries the frame set: at most 65535 per launch, any number per call
        const int total = p.n_stacks;
        for (int done = 0; done < total; done += 65535) {
            KParams q = p;
            q.n_stacks = (total - done) < 65535 ? (total - done) : 65535;
            q.stack = p.stack + (size_t)done * p.P * p.npx;
            if (p.depth) { q.depth = p.depth + (size_t)done * p.npx; q.mask_bits = p.mask_bits + (size_t)done * p.bits_stride; }
            else { q.xyzw = p.xyzw + (size_t)done * p.npx; q.mask = p.mask + (size_t)done * p.npx; }
            if (p.kbin) q.kbin = p.kbin + (size_t)done * p.npx;
            if (p.corr) q.corr = p.corr + (size_t)done * p.npx;
            if (p.phase_pix) q.phase_pix = p.phase_pix + (size_t)done * p.npx;
            if (p.proj_u) q.proj_u = p.proj_u + (size_t)done * p.npx;
            fn<<<dim3((unsigned)blocks, (unsigned)q.n_stacks), kBlock, plan.smem, stream>>>(q);
            const cudaError_t err = cudaGetLastError();
            if (err != cudaSuccess) return err;
        }
        return cudaSuccess;
    }
    const long long total = p.npx * (long long)p.n_stacks;
    const long long blocks = (total + kBlock - 1) / kBlock;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    if (p.depth) {
        // the any-geometry kernel sets validity bits with atomicOr
        const cudaError_t err = cudaMemsetAsync(p.mask_bits, 0, (size_t)p.bits_stride * p.n_stacks, stream);
        if (err != cudaSuccess) return err;
    }
    auto fn = plan.mode == 2 ? reconstruct_scalar_kernel<true> : reconstruct_scalar_kernel<false>;
    fn<<<(unsigned)blocks, kBlock, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_decode_gray(const KParams& p, const uint8_t* d_planes, double* d_gray_val,
                               int16_t* d_kbin, cudaStream_t stream)
{
    const long long blocks = (p.npx + kBlock - 1) / kBlock;
    decode_gray_kernel<<<(unsigned)blocks, kBlock, 0, stream>>>(p, d_planes, d_gray_val, d_kbin);
    return cudaGetLastError();
}

cudaError_t launch_decode_phase(const KParams& p, const uint8_t* d_planes, double* d_phase_pix,
                                uint8_t* d_mod_ok, cudaStream_t stream)
{
    const long long blocks = (p.npx + kBlock - 1) / kBlock;
    decode_phase_kernel<<<(unsigned)blocks, kBlock, 0, stream>>>(p, d_planes, d_phase_pix, d_mod_ok);
    return cudaGetLastError();
}

cudaError_t launch_eval_phase(const float* d_s, const float* d_c, long long n, float Tf, float* d_deg, float* d_pix,
                              cudaStream_t stream)
{
    const long long blocks = (n + kBlock - 1) / kBlock;
    if (blocks <= 0) return cudaSuccess;
    eval_phase_kernel<<<(unsigned)blocks, kBlock, 0, stream>>>(d_s, d_c, n, Tf, d_deg, d_pix);
    return cudaGetLastError();
}

cudaError_t launch_triangulate(const KParams& p, const double* d_proj_u, float* d_xyzw, uint8_t* d_mask,
                               cudaStream_t stream)
{
    const long long blocks = (p.npx + kBlock - 1) / kBlock;
    triangulate_kernel<<<(unsigned)blocks, kBlock, 0, stream>>>(p, d_proj_u, reinterpret_cast<float4*>(d_xyzw),
                                                                d_mask);
    return cudaGetLastError();
}

cudaError_t launch_triangulate_uv(const KParams& p, const double* d_proj_u, const double* d_proj_v, float* d_xyzw,
                                  uint8_t* d_mask, cudaStream_t stream)
{
    const long long blocks = (p.npx + kBlock - 1) / kBlock;
    triangulate_uv_kernel<<<(unsigned)blocks, kBlock, 0, stream>>>(p, d_proj_u, d_proj_v,
                                                                   reinterpret_cast<float4*>(d_xyzw), d_mask);
    return cudaGetLastError();
}

}  // namespace slc
