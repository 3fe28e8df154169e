// slc_pointcloud.cu -- point-cloud output: CCalculation::Result
// (CCalculation.cpp:323-357) on the device.
//
// The reference walks the image u outer / v inner, skips pixels whose z is outside
// [FOV_MIN_DISTANCE, FOV_MAX_DISTANCE] and writes "x y z\n" with ostream << double, i.e.
// printf("%g") with precision 6.  Here that is a variable-length record emission in ONE launch
// (pc_text_kernel), a chained scan over tiles of 2048 records in output order:
//   phase A  everything that needs f64, once per number: x, y, z from the f64 ProjectorU plane in the
//            reference's operation order (:686-687, :761-767) and their six exactly rounded digits +
//            decimal exponent (decode_g6), kept in SHARED memory as three 32-bit codes + the line length
//            per record; the tile's byte / line totals are published and the totals of the tiles before it
//            summed (decoupled look-back; counts, flag and launch epoch travel in one 64-bit word)
//   phase B  the characters of every record (emit_g6_fast: integer work only) go straight into their
//            place in a shared-memory chunk laid out at the output's own 16-byte phase, written as uint4.
// (Until round 2 this was two launches with a 16 B/px scratch round trip between them.)  The text is
// byte-identical to what the reference's doubles print.  pc_emit_kernel -- packed float3 xyz of the
// valid pixels of a float4 map in two passes -- remains for the geometries slc_compact.cu does not take.
#include "slc_kernels.h"

namespace slc {

namespace {

constexpr int kPcThreads = 256;
constexpr int kPcIters = 8;                      // records per thread
constexpr int kPcChunk = kPcThreads * kPcIters;  // records per block
constexpr int kPcMaxNum = 13;                    // "-1.23457e-005"
constexpr int kPcMaxLine = 3 * kPcMaxNum + 2 + 2;   // two blanks, "\r\n"

__constant__ double c_pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// |x| * 10^k as a correctly rounded double p plus what is needed to compare the EXACT product
// with a half-integer h:  sign(exact - h) = sign(cmp(p - h)).   10^|k| is exact for |k| <= 22.
struct Scaled {
    double p, m, aux;
    int kind;          // 0: p = ax*m, aux = fma(ax, m, -p);  1: p = ax/m, aux = fma(-p, m, ax);  2: approximate
};
__device__ __forceinline__ Scaled scale10(double ax, int k)
{
    Scaled s;
    if (k >= 0 && k <= 22) {
        s.m = c_pow10[k];
        s.p = __dmul_rn(ax, s.m);
        s.aux = __fma_rn(ax, s.m, -s.p);
        s.kind = 0;
    } else if (k < 0 && k >= -22) {
        s.m = c_pow10[-k];
        s.p = __ddiv_rn(ax, s.m);
        s.aux = __fma_rn(-s.p, s.m, ax);
        s.kind = 1;
    } else {            // |x| < 1e-17 or >= 1e28: not a coordinate; last digit not guaranteed
        double v = ax;
        int kk = k;
        while (kk > 22) { v *= 1e22; kk -= 22; }
        while (kk < -22) { v /= 1e22; kk += 22; }
        s.m = 1.0;
        s.p = (kk >= 0) ? v * c_pow10[kk] : v / c_pow10[-kk];
        s.aux = 0.0;
        s.kind = 2;
    }
    return s;
}

// printf("%g", x) (precision 6, the C locale) == what `ostream << double` writes with default
// flags, in two steps so that the expensive one runs once per number:
//   decode_g6  the exactly rounded six significant digits D and the decimal exponent E, packed in
//              32 bits:  [19:0] D (100000..999999; 0 = zero, 1 = inf, 2 = nan), [29:20] E + 512, [31] sign
//   len_g6 / emit_g6   the characters %g makes of (D, E): notation, stripped zeros, exponent field.
//              exp3: three exponent digits (the MSVC 2013 CRT the reference was built with), not two.
constexpr unsigned kG6Zero = 0u, kG6Inf = 1u, kG6Nan = 2u;

__device__ __forceinline__ unsigned decode_g6(double x)
{
    const unsigned long long bits = (unsigned long long)__double_as_longlong(x);
    const unsigned sign = (unsigned)(bits >> 63) << 31;
    const double ax = fabs(x);
    if (ax == 0.0) return sign | kG6Zero;
    const int bexp = (int)((bits >> 52) & 0x7FFu);
    if (bexp == 0x7FF) return sign | (((bits & 0xFFFFFFFFFFFFFull) != 0ull) ? kG6Nan : kG6Inf);
    // decimal exponent: 10^E <= |x| < 10^(E+1); the estimate is within one, then one fix-up
    int E = (int)floor((double)(bexp == 0 ? -1074 + (63 - __clzll((long long)(bits & 0xFFFFFFFFFFFFFull)))
                                          : bexp - 1023) * 0.30102999566398120);
    Scaled s = scale10(ax, 5 - E);
    if (s.p < 1e5) { E--; s = scale10(ax, 5 - E); }
    else if (s.p > 1e6) { E++; s = scale10(ax, 5 - E); }
    // D = round-half-even of the exact |x| * 10^(5-E)
    const double fl = floor(s.p);
    const double diff = s.p - (fl + 0.5);                       // exact (Sterbenz)
    const double t = (s.kind == 0) ? diff + s.aux : (s.kind == 1) ? __fma_rn(diff, s.m, s.aux) : diff;
    unsigned D = (unsigned)fl;
    if (t > 0.0 || (t == 0.0 && (D & 1u))) D++;
    if (D >= 1000000u) { D = 100000u; E++; }
    return sign | ((unsigned)(E + 512) << 20) | D;
}

// significant digits left after %g strips trailing zeros
__device__ __forceinline__ int g6_digits(unsigned D)
{
    // 6 - (trailing decimal zeros of D): five divisibility tests, no loop (a `while (q % 10 == 0)` loop
    // diverges inside a warp and was a fifth of pass 1's instructions).  D = 0 gives 1.
    const int tz = (D % 10u == 0u) + (D % 100u == 0u) + (D % 1000u == 0u) + (D % 10000u == 0u) + (D % 100000u == 0u);
    return 6 - tz;
}

// characters emit_g6 writes for (code, nd = g6_digits(D))
__device__ __forceinline__ int len_g6(unsigned code, int nd, bool exp3)
{
    const int neg = (int)(code >> 31);
    const unsigned D = code & 0xFFFFFu;
    if (D < 100000u) return neg + (D == kG6Zero ? 1 : 3);
    const int E = (int)((code >> 20) & 0x3FFu) - 512;
    if (E < -4 || E >= 6) {
        const int ae = E < 0 ? -E : E;
        return neg + 1 + (nd > 1 ? nd : 0) + 2 + ((exp3 || ae >= 100) ? 3 : 2);
    }
    if (E >= 0) return neg + (E + 1) + (nd > E + 1 ? nd - E : 0);
    return neg + 1 - E + nd;
}

__device__ __forceinline__ int emit_g6(unsigned code, char* out, bool exp3)
{
    int n = 0;
    if (code >> 31) out[n++] = '-';
    unsigned D = code & 0xFFFFFu;
    if (D < 100000u) {
        if (D == kG6Zero) { out[n++] = '0'; return n; }
        const bool is_nan = (D == kG6Nan);
        out[n++] = is_nan ? 'n' : 'i'; out[n++] = is_nan ? 'a' : 'n'; out[n++] = is_nan ? 'n' : 'f';
        return n;
    }
    const int E = (int)((code >> 20) & 0x3FFu) - 512;
    const int nd = g6_digits(D);
    auto next_digit = [&]() -> char { const unsigned d = D / 100000u; D = (D - d * 100000u) * 10u; return (char)('0' + d); };
    if (E < -4 || E >= 6) {
        out[n++] = next_digit();
        if (nd > 1) {
            out[n++] = '.';
            for (int i = 1; i < nd; i++) out[n++] = next_digit();
        }
        out[n++] = 'e';
        out[n++] = (E < 0) ? '-' : '+';
        const unsigned ae = (unsigned)(E < 0 ? -E : E);
        if (exp3 || ae >= 100u) out[n++] = (char)('0' + ae / 100u);
        out[n++] = (char)('0' + (ae / 10u) % 10u);
        out[n++] = (char)('0' + ae % 10u);
    } else if (E >= 0) {
        for (int i = 0; i <= E; i++) out[n++] = next_digit();   // digits past nd are zeros
        if (nd > E + 1) {
            out[n++] = '.';
            for (int i = E + 1; i < nd; i++) out[n++] = next_digit();
        }
    } else {
        out[n++] = '0';
        out[n++] = '.';
        for (int i = 0; i < -E - 1; i++) out[n++] = '0';
        for (int i = 0; i < nd; i++) out[n++] = next_digit();
    }
    return n;
}

// emit_g6 without loops for the numbers a point cloud is made of: 1 <= |x| < 10^6, i.e. fixed notation
// "ddd.ddd".  The six ASCII digits are made once (two divisions by 1000 / 100 / 10 as multiply-shift) and
// every character is one predicated byte store at a position known from (E, nd): digit k sits at
// k + (k > E), the point at E + 1, and digits at or beyond max(nd, E + 1) are not written.  Anything
// else (|x| < 1, exponent notation, zero, inf, nan) takes emit_g6.  nd = g6_digits(D).
__device__ __forceinline__ int emit_g6_fast(unsigned code, int nd, char* out, bool exp3)
{
    const unsigned D = code & 0xFFFFFu;
    const int E = (int)((code >> 20) & 0x3FFu) - 512;
    if (D < 100000u || (unsigned)E > 5u) return emit_g6(code, out, exp3);
    const unsigned hi = D / 1000u, lo = D - hi * 1000u;
    const unsigned a0 = hi / 100u, r0 = hi - a0 * 100u, b0 = r0 / 10u, c0 = r0 - b0 * 10u;
    const unsigned a1 = lo / 100u, r1 = lo - a1 * 100u, b1 = r1 / 10u, c1 = r1 - b1 * 10u;
    const unsigned dg[6] = {a0, b0, c0, a1, b1, c1};
    const int neg = (int)(code >> 31);
    if (neg) out[0] = '-';
    char* q = out + neg;
    const int last = max(nd, E + 1);                 // digits written (those past nd are the zeros of D)
#pragma unroll
    for (int k = 0; k < 6; k++)
        if (k < last) q[k + (k > E ? 1 : 0)] = (char)('0' + dg[k]);
    const int point = (nd > E + 1) ? 1 : 0;
    if (point) q[E + 1] = '.';
    return neg + last + point;
}

__device__ __forceinline__ int format_g6(double x, char* out, bool exp3)
{
    const unsigned code = decode_g6(x);
    return emit_g6_fast(code, g6_digits(code & 0xFFFFFu), out, exp3);
}

struct PcArgs {
    int W, H;
    long long npx;
    int order;                       // 0 row-major, 1 reference (u outer, v inner: CCalculation.cpp:336-338)
    const double* proj_u;            // text
    const float4* xyzw;              // float3 records
    const uint8_t* mask;             // float3 records
    unsigned flags;
    unsigned long long* block_sums;  // [2 * n_blocks + 2]: (bytes, records) per block, totals last
    unsigned char* out;
    unsigned long long capacity;     // bytes
};

// the line "x y z" + line end of pixel (v, u) as three decode_g6 codes and (length, digit counts); length 0
// if the reference skips the pixel
__device__ __forceinline__ uint4 text_line_codes(const KParams& p, double U, int u, int v, bool exp3, bool crlf)
{
    if (U == 0.0) return make_uint4(0u, 0u, 0u, 0u);                // :678-682
    const double z = z_exact(p, U, u, v);                          // :686-687
    if ((z < p.fov_min) || (z > p.fov_max)) return make_uint4(0u, 0u, 0u, 0u);   // :701-704 and :341-345
    const double x = __ddiv_rn(__dmul_rn(z, __dsub_rn((double)u, p.cu)), p.fu);   // :766
    const double y = __ddiv_rn(__dmul_rn(z, __dsub_rn((double)v, p.cv)), p.fv);   // :767
    uint4 r;
    r.x = decode_g6(x); r.y = decode_g6(y); r.z = decode_g6(z);
    // w: [5:0] line length, [8:6] / [11:9] / [14:12] significant digits of x / y / z after zero stripping
    const unsigned ndx = (unsigned)g6_digits(r.x & 0xFFFFFu), ndy = (unsigned)g6_digits(r.y & 0xFFFFFu),
                   ndz = (unsigned)g6_digits(r.z & 0xFFFFFu);
    r.w = (unsigned)(len_g6(r.x, (int)ndx, exp3) + len_g6(r.y, (int)ndy, exp3) + len_g6(r.z, (int)ndz, exp3) + 3 + (crlf ? 1 : 0)) |
          (ndx << 6) | (ndy << 9) | (ndz << 12);
    return r;
}

__device__ __forceinline__ void text_line_emit(const uint4& r, char* o, bool exp3, bool crlf)
{
    int len = emit_g6_fast(r.x, (int)((r.w >> 6) & 7u), o, exp3);
    o[len++] = ' ';
    len += emit_g6_fast(r.y, (int)((r.w >> 9) & 7u), o + len, exp3);
    o[len++] = ' ';
    len += emit_g6_fast(r.z, (int)((r.w >> 12) & 7u), o + len, exp3);
    if (crlf) o[len++] = '\r';
    o[len++] = '\n';
}

// inclusive scan of `v` over the block; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kPcThreads / 32; w++) {
        const int sw = s_warp[w];
        if (w < warp) base += sw;
        tot += sw;
    }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}

// Packed float3 xyz of the valid pixels of a float4 map in two passes (count, then scan + emit): the
// path of the geometries the chained-scan kernels of slc_compact.cu do not take (widths that are not a
// multiple of 8).
template <bool WRITE>
__global__ void __launch_bounds__(kPcThreads)
pc_emit_kernel(const PcArgs a)
{
    __shared__ __align__(16) unsigned char s_txt[WRITE ? kPcThreads * 12 + 32 : 16];
    __shared__ int s_warp[kPcThreads / 32];
    const int t = threadIdx.x;
    const long long base = (long long)blockIdx.x * kPcChunk;
    unsigned long long nbytes = 0ull, nrec = 0ull;
    unsigned long long gofs = 0ull;
    if (WRITE) {
        // this block's offset = sum of the byte counts of all earlier blocks (a few thousand values
        // sitting in L2); the last block also publishes the totals
        __shared__ unsigned long long s_part[3][kPcThreads / 32];
        const int nb = (int)gridDim.x, me = (int)blockIdx.x;
        const bool last = (me == nb - 1);
        unsigned long long before = 0ull, all_b = 0ull, all_r = 0ull;
        for (int i = t; i < (last ? nb : me); i += kPcThreads) {
            const unsigned long long vb = a.block_sums[2 * i];
            if (i < me) before += vb;
            all_b += vb;
            all_r += a.block_sums[2 * i + 1];
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            before += __shfl_down_sync(0xFFFFFFFFu, before, d);
            all_b += __shfl_down_sync(0xFFFFFFFFu, all_b, d);
            all_r += __shfl_down_sync(0xFFFFFFFFu, all_r, d);
        }
        if ((t & 31) == 0) { s_part[0][t >> 5] = before; s_part[1][t >> 5] = all_b; s_part[2][t >> 5] = all_r; }
        __syncthreads();
        unsigned long long tb = 0ull, ta = 0ull, tr = 0ull;
        for (int w = 0; w < kPcThreads / 32; w++) { tb += s_part[0][w]; ta += s_part[1][w]; tr += s_part[2][w]; }
        gofs = tb;
        if (last && t == 0) {
            a.block_sums[2 * nb] = ta;
            a.block_sums[2 * nb + 1] = tr;
        }
    }

    for (int it = 0; it < kPcIters; it++) {
        const long long i = base + (long long)it * kPcThreads + t;
        int len = 0;
        long long px = 0;
        if (i < a.npx) {
            // npx < 2^31 (slc_create): 32-bit division, not the 64-bit emulation
            const unsigned ii = (unsigned)i;
            unsigned u, v;
            if (a.order == 1) { u = ii / (unsigned)a.H; v = ii - u * (unsigned)a.H; }
            else { v = ii / (unsigned)a.W; u = ii - v * (unsigned)a.W; }
            px = (long long)v * a.W + u;
            len = (a.mask[px] != 0) ? 12 : 0;
        }
        int total;
        const int excl = block_exclusive_scan(len, s_warp, &total);
        nbytes += (unsigned long long)len;       // per thread; reduced below
        nrec += (len > 0) ? 1ull : 0ull;
        if (WRITE) {
            const unsigned align = (unsigned)(gofs & 15ull);
            if (gofs + (unsigned long long)total <= a.capacity) {
                if (len) {
                    // 12-byte records: align + excl is a multiple of 4 (gofs is a multiple of 12 from a 16-aligned base)
                    const float4 q = a.xyzw[px];
                    float* d = reinterpret_cast<float*>(&s_txt[align + excl]);
                    d[0] = q.x; d[1] = q.y; d[2] = q.z;
                }
                __syncthreads();
                unsigned char* g = a.out + (gofs - align);         // 16-byte aligned
                const int end = (int)align + total;
                const int body0 = align ? 16 : 0, body1 = end & ~15;
                for (int j = (int)align + t; j < min(16, end) && align; j += kPcThreads) g[j] = s_txt[j];
                for (int j = body0 + 16 * t; j < body1; j += 16 * kPcThreads)
                    *reinterpret_cast<uint4*>(g + j) = *reinterpret_cast<const uint4*>(&s_txt[j]);
                for (int j = max(body1, body0) + t; j < end; j += kPcThreads) g[j] = s_txt[j];
                __syncthreads();
            }
            gofs += (unsigned long long)total;
        }
    }
    if (!WRITE) {
        // block totals: reduce the per-thread counters
        __shared__ unsigned long long s_red[2][kPcThreads / 32];
        const int lane = t & 31, warp = t >> 5;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            nbytes += __shfl_down_sync(0xFFFFFFFFu, nbytes, d);
            nrec += __shfl_down_sync(0xFFFFFFFFu, nrec, d);
        }
        if (lane == 0) { s_red[0][warp] = nbytes; s_red[1][warp] = nrec; }
        __syncthreads();
        if (t == 0) {
            unsigned long long b = 0, r = 0;
            for (int w = 0; w < kPcThreads / 32; w++) { b += s_red[0][w]; r += s_red[1][w]; }
            a.block_sums[2 * blockIdx.x] = b;
            a.block_sums[2 * blockIdx.x + 1] = r;
        }
    }
}

// ---- the text cloud in one launch ---------------------------------------------------------------
// look-back word: [63:62] flag (1 = this tile's totals, 2 = inclusive prefix), [61:60] launch epoch (1..3;
// every launch rewrites every tile's word, so a stale word always carries the previous launch's epoch),
// [59:26] bytes, [25:0] lines -- the two counts add without meeting while a frame stays below 2^34 bytes
// and 2^26 lines (67 M pixels).
constexpr unsigned long long kTxAggregate = 1ull << 62, kTxInclusive = 2ull << 62, kTxCounts = (1ull << 60) - 1ull;

__device__ __forceinline__ unsigned long long tx_ld(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void tx_st(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// exclusive prefix (packed bytes | lines) of the tiles before `tile`; every thread of warp 0 calls it
__device__ __forceinline__ unsigned long long tx_lookback(unsigned long long* state, int tile, unsigned long long mine,
                                                          unsigned epoch)
{
    const int lane = threadIdx.x & 31;
    const unsigned long long tag = (unsigned long long)epoch << 60;
    if (tile == 0) {
        if (lane == 0) tx_st(state, kTxInclusive | tag | mine);
        return 0ull;
    }
    if (lane == 0) tx_st(state + tile, kTxAggregate | tag | mine);
    unsigned long long prefix = 0ull;
    for (int j = tile - 1;; j -= 32) {
        const int idx = j - lane;
        unsigned long long w;
        bool ready;
        do {
            w = (idx >= 0) ? tx_ld(state + idx) : (kTxInclusive | tag);
            ready = (((w >> 60) & 3ull) == (unsigned long long)epoch) && ((w >> 62) != 0ull);
        } while (!__all_sync(0xFFFFFFFFu, ready));
        const unsigned incl = __ballot_sync(0xFFFFFFFFu, (w >> 62) == 2ull);
        const int stop = incl ? (__ffs((int)incl) - 1) : 31;
        unsigned long long v = (lane <= stop) ? (w & kTxCounts) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
        prefix += v;
        if (incl) break;
    }
    if (lane == 0) tx_st(state + tile, kTxInclusive | tag | (prefix + mine));
    return prefix;
}

__global__ void __launch_bounds__(kPcThreads)
pc_text_kernel(const __grid_constant__ KParams p, const PcArgs a, unsigned long long* __restrict__ state, unsigned epoch)
{
    __shared__ uint4 s_rec[kPcChunk];
    __shared__ __align__(16) unsigned char s_txt[kPcThreads * kPcMaxLine + 32];
    __shared__ int s_warp[kPcThreads / 32];
    __shared__ unsigned long long s_red[kPcThreads / 32];
    __shared__ unsigned long long s_prefix;
    const int t = threadIdx.x, tile = (int)blockIdx.x;
    const long long base = (long long)tile * kPcChunk;
    const bool crlf = (a.flags & 1u) != 0u, exp3 = (a.flags & 2u) != 0u;

    // ---- phase A: the codes of the tile's records, in output order (u outer, v inner) ----
    unsigned long long mine = 0ull;                     // bytes << 26 | lines, this thread's records
#pragma unroll 2
    for (int it = 0; it < kPcIters; it++) {
        const long long i = base + (long long)it * kPcThreads + t;
        uint4 rec = make_uint4(0u, 0u, 0u, 0u);
        if (i < a.npx) {
            // npx < 2^31 (slc_create): 32-bit division, not the 64-bit emulation
            const unsigned ii = (unsigned)i;
            unsigned u, v;
            if (a.order == 1) { u = ii / (unsigned)a.H; v = ii - u * (unsigned)a.H; }
            else { v = ii / (unsigned)a.W; u = ii - v * (unsigned)a.W; }
            rec = text_line_codes(p, a.proj_u[(long long)v * a.W + u], (int)u, (int)v, exp3, crlf);
        }
        s_rec[it * kPcThreads + t] = rec;
        const unsigned len = rec.w & 63u;
        mine += ((unsigned long long)len << 26) + (len ? 1ull : 0ull);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_down_sync(0xFFFFFFFFu, mine, d);
    if ((t & 31) == 0) s_red[t >> 5] = mine;
    __syncthreads();
    if (t < 32) {
        unsigned long long tot = 0ull;
#pragma unroll
        for (int w = 0; w < kPcThreads / 32; w++) tot += s_red[w];
        const unsigned long long before = tx_lookback(state, tile, tot, epoch);
        if (t == 0) {
            s_prefix = before;
            if (tile == (int)gridDim.x - 1) {
                const unsigned long long all = before + tot;
                a.block_sums[0] = all >> 26;                     // bytes of the whole text
                a.block_sums[1] = all & ((1ull << 26) - 1ull);   // lines
            }
        }
    }
    __syncthreads();
    unsigned long long gofs = s_prefix >> 26;

    // ---- phase B: characters, 256 records at a time, staged at the output's own 16-byte phase ----
    for (int it = 0; it < kPcIters; it++) {
        const uint4 rec = s_rec[it * kPcThreads + t];
        const int len = (int)(rec.w & 63u);
        int total;
        const int excl = block_exclusive_scan(len, s_warp, &total);
        const unsigned align = (unsigned)(gofs & 15ull);
        if (gofs + (unsigned long long)total <= a.capacity) {
            if (len) text_line_emit(rec, reinterpret_cast<char*>(&s_txt[align + excl]), exp3, crlf);
            __syncthreads();
            unsigned char* g = a.out + (gofs - align);         // 16-byte aligned
            const int end = (int)align + total;
            const int body0 = align ? 16 : 0, body1 = end & ~15;
            for (int j = (int)align + t; j < min(16, end) && align; j += kPcThreads) g[j] = s_txt[j];
            for (int j = body0 + 16 * t; j < body1; j += 16 * kPcThreads)
                *reinterpret_cast<uint4*>(g + j) = *reinterpret_cast<const uint4*>(&s_txt[j]);
            for (int j = max(body1, body0) + t; j < end; j += kPcThreads) g[j] = s_txt[j];
            __syncthreads();
        }
        gofs += (unsigned long long)total;
    }
}

__global__ void format_g6_kernel(const double* __restrict__ v, long long n, unsigned flags, char* __restrict__ text,
                                 uint8_t* __restrict__ len)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    char buf[16];
    const int l = format_g6(v[i], buf, (flags & 2u) != 0u);
    for (int j = 0; j < 16; j++) text[i * 16 + j] = (j < l) ? buf[j] : 0;
    len[i] = (uint8_t)l;
}

}  // namespace

static size_t pointcloud_sums_bytes(long long npx)
{
    const long long blocks = (npx + kPcChunk - 1) / kPcChunk;
    return ((size_t)(2 * blocks + 2) * sizeof(unsigned long long) + 255) & ~(size_t)255;
}

size_t pointcloud_scratch_bytes(long long npx)
{
    return 2 * pointcloud_sums_bytes(npx);   // block sums + totals (two-pass float3 emitter) | look-back words (text)
}

// mode 0: text lines from d_proj_u; mode 1: float3 of the valid pixels of (d_xyzw, d_mask).
// d_totals (inside d_scratch, 2 x u64: bytes, records) is valid once the stream has drained.
cudaError_t launch_pointcloud(const KParams& p, int mode, int order, unsigned flags, const double* d_proj_u,
                              const float* d_xyzw, const uint8_t* d_mask, void* d_out, unsigned long long capacity,
                              void* d_scratch, unsigned text_epoch, const unsigned long long** d_totals, cudaStream_t stream)
{
    if ((reinterpret_cast<uintptr_t>(d_out) & 15) != 0) return cudaErrorMisalignedAddress;
    const int blocks = (int)((p.npx + kPcChunk - 1) / kPcChunk);
    PcArgs a{};
    a.W = p.W; a.H = p.H; a.npx = p.npx; a.order = order;
    a.proj_u = d_proj_u;
    a.xyzw = reinterpret_cast<const float4*>(d_xyzw);
    a.mask = d_mask;
    a.flags = flags;
    a.block_sums = static_cast<unsigned long long*>(d_scratch);
    a.out = static_cast<unsigned char*>(d_out);
    a.capacity = capacity;
    if (mode == 0) {
        // one launch: the totals land in the first two words of the scratch, the look-back words behind the sums
        if (p.npx >= (1ll << 26)) return cudaErrorInvalidValue;    // line count field of the look-back word
        unsigned long long* state = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(d_scratch) + pointcloud_sums_bytes(p.npx));
        pc_text_kernel<<<blocks, kPcThreads, 0, stream>>>(p, a, state, text_epoch);
        *d_totals = a.block_sums;
    } else {
        pc_emit_kernel<false><<<blocks, kPcThreads, 0, stream>>>(a);
        pc_emit_kernel<true><<<blocks, kPcThreads, 0, stream>>>(a);
        *d_totals = a.block_sums + 2 * blocks;
    }
    return cudaGetLastError();
}

cudaError_t launch_format_g6(const double* d_values, long long n, unsigned flags, char* d_text, uint8_t* d_len,
                             cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    format_g6_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_values, n, flags, d_text, d_len);
    return cudaGetLastError();
}

}  // namespace slc
