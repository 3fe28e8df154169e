// slc_compact.cu -- valid-only point lists straight from the (xyzw, mask) maps, in ONE pass.
//
// The reference's only consumer of the maps, CCalculation::Result (CCalculation.cpp:336-346),
// keeps the pixels inside the field of view and drops the rest; the full float4 map + byte mask
// (17 B/px) is therefore more than anything downstream reads.  These kernels turn a batch of maps
// into
//     points    float [n][point_stride][3]   (x, y, z) of the valid pixels of each frame set,
//                                            in the maps' memory order (SLC_ORDER_ROW_MAJOR) or in the
//                                            order Result() walks (SLC_ORDER_REFERENCE: u outer, v inner)
//     mask_bits u8    [n][npx/8]             one validity bit per pixel (bit i&7 of byte i>>3)
//     counts    u64   [n]                    points per frame set
// with a chained ("decoupled look-back") scan: a block counts the valid pixels of its tile,
// publishes the count, sums the counts / inclusive prefixes of the tiles before it and writes its
// points at the resulting offset -- one launch for the whole batch, no count pass, no scratch
// round trip.  The look-back words carry a launch epoch, so the state array never needs clearing.
//
// Reference order: a tile is 8 adjacent columns over the full image height (contiguous in the
// u-major output order).  The float4 map is read row-major, 128 contiguous bytes per row, into a
// padded shared tile and emitted column-major (the transpose of the first-frame kernel's store
// stage), so neither side touches a 32-byte sector twice.
#include "slc_kernels.h"

namespace slc {

namespace {

constexpr int kCThreads = 256;
constexpr int kRowTile = 2048;     // pixels per tile, row-major order (8 per thread)
constexpr int kColTileW = 8;       // columns per tile, reference order
constexpr int kColChunk = 128;     // rows staged per iteration, reference order

constexpr unsigned long long kFlagAggregate = 1ull << 32, kFlagInclusive = 2ull << 32;

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// Exclusive prefix of this tile's `total` over the tiles before it (same frame set).  Called by
// every thread of warp 0; count and flag travel in one 64-bit word, so no fence is needed.
__device__ __forceinline__ unsigned lookback_exclusive(unsigned long long* state, int tile, unsigned total, unsigned epoch)
{
    const int lane = threadIdx.x & 31;
    const unsigned long long tag = (unsigned long long)epoch << 34;
    if (tile == 0) {
        if (lane == 0) st_state(state, tag | kFlagInclusive | total);
        return 0u;
    }
    if (lane == 0) st_state(state + tile, tag | kFlagAggregate | total);
    unsigned prefix = 0u;
    for (int j = tile - 1;; j -= 32) {
        const int idx = j - lane;
        unsigned long long w;
        bool ready;
        do {
            w = (idx >= 0) ? ld_state(state + idx) : (tag | kFlagInclusive);
            ready = ((w >> 34) == (unsigned long long)epoch) && ((w & (kFlagAggregate | kFlagInclusive)) != 0ull);
        } while (!__all_sync(0xFFFFFFFFu, ready));
        const unsigned incl = __ballot_sync(0xFFFFFFFFu, (w & kFlagInclusive) != 0ull);
        const int stop = incl ? (__ffs((int)incl) - 1) : 31;   // nearest tile that already holds an inclusive prefix
        unsigned v = (lane <= stop) ? (unsigned)(w & 0xFFFFFFFFull) : 0u;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
        prefix += v;
        if (incl) break;
    }
    if (lane == 0) st_state(state + tile, tag | kFlagInclusive | (unsigned long long)(prefix + total));
    return prefix;
}

// block-wide exclusive scan of one unsigned per thread (kCThreads threads); *total = block sum
__device__ __forceinline__ unsigned block_scan_u32(unsigned v, unsigned* s_warp, unsigned* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kCThreads / 32; w++) {
        const unsigned sw = s_warp[w];
        if (w < warp) base += sw;
        tot += sw;
    }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}

// bytes != 0 -> bits, for 8 mask bytes held in a uint2: bit k of the result = byte k is non-zero
__device__ __forceinline__ unsigned nonzero_bits8(uint2 m)
{
    auto nz4 = [](unsigned w) {   // 0x01 in every byte that is non-zero
        return ((((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u) >> 7;
    };
    // bits 0, 8, 16, 24 -> bits 24..27 of the product (no two partial products meet there)
    const unsigned pa = (nz4(m.x) * 0x01020408u) >> 24, pb = (nz4(m.y) * 0x01020408u) >> 24;
    return (pa & 0xFu) | ((pb & 0xFu) << 4);
}

struct CompactArgs {
    int W, H;
    long long npx;
    const float4* xyzw;                 // [n][npx]
    const uint8_t* mask;                // [n][npx]
    float* points;                      // [n][point_stride][3]
    long long point_stride;
    uint8_t* mask_bits;                 // [n][bits_stride] or nullptr
    long long bits_stride;
    unsigned long long* counts;         // [n]
    unsigned long long* state;          // [n][tiles]
    int tiles;
    unsigned epoch;
};

// ---- SLC_ORDER_ROW_MAJOR: tile = 2048 consecutive pixels ------------------------------------
__global__ void __launch_bounds__(kCThreads)
compact_rows_kernel(const CompactArgs a)
{
    __shared__ __align__(16) float s_pts[kRowTile * 3 + 4];
    __shared__ short s_rank[kRowTile];
    __shared__ unsigned s_warp[kCThreads / 32];
    __shared__ unsigned s_base;
    const int t = threadIdx.x, tile = blockIdx.x, stack = blockIdx.y;
    const long long tile0 = (long long)tile * kRowTile;
    const uint8_t* mask = a.mask + (long long)stack * a.npx;
    const float4* xyzw = a.xyzw + (long long)stack * a.npx;

    // phase A: validity of this thread's 8 consecutive pixels
    const long long p0 = tile0 + 8 * t;
    unsigned bits = 0u;
    if (p0 + 8 <= a.npx) {
        bits = nonzero_bits8(*reinterpret_cast<const uint2*>(mask + p0));
    } else {
        for (int j = 0; j < 8; j++)
            if (p0 + j < a.npx && mask[p0 + j] != 0) bits |= 1u << j;
    }
    if (a.mask_bits != nullptr && p0 < a.npx) a.mask_bits[(long long)stack * a.bits_stride + (p0 >> 3)] = (uint8_t)bits;
    unsigned total;
    const unsigned excl = block_scan_u32((unsigned)__popc(bits), s_warp, &total);
#pragma unroll
    for (int j = 0; j < 8; j++)
        s_rank[8 * t + j] = ((bits >> j) & 1u) ? (short)(excl + (unsigned)__popc(bits & ((1u << j) - 1u))) : (short)-1;
    if (t < 32) {
        const unsigned b = lookback_exclusive(a.state + (long long)stack * a.tiles, tile, total, a.epoch);
        if (t == 0) {
            s_base = b;
            if (tile == a.tiles - 1) a.counts[stack] = (unsigned long long)b + total;
        }
    }
    __syncthreads();
    // phase B: float4 of the valid pixels, 512 contiguous bytes per warp load -> packed float3 in shared memory,
    // staged at the output's own 16-byte phase so that the body goes out as 16-byte stores
    const long long base = (long long)s_base;
    const long long g0 = ((long long)stack * a.point_stride + base) * 3;      // first float of this tile in `points`
    const unsigned ph = (unsigned)(g0 & 3);                                   // (points is 16-byte aligned: cudaMalloc / slc_host_alloc)
    float4 q[8];
    int r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r[i] = s_rank[i * kCThreads + t];
        if (r[i] >= 0) q[i] = __ldcs(xyzw + tile0 + i * kCThreads + t);
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
        if (r[i] >= 0) { float* d = s_pts + ph + 3 * r[i]; d[0] = q[i].x; d[1] = q[i].y; d[2] = q[i].z; }
    __syncthreads();
    long long room = a.point_stride - base;                 // points that still fit this frame set's slice
    if (room <= 0) return;
    const unsigned n_out = (unsigned)(room < (long long)total ? room : (long long)total) * 3u;
    float* dst = a.points + g0 - ph;                        // 16-byte aligned; float k of the tile sits at dst[ph + k]
    const unsigned end = ph + n_out, body0 = ph ? 4u : 0u, body1 = end & ~3u;
    if ((reinterpret_cast<uintptr_t>(a.points) & 15) == 0) {
        for (unsigned j = ph + t; j < min(4u, end) && ph; j += kCThreads) __stcs(dst + j, s_pts[j]);
        for (unsigned j = body0 + 4 * t; j < body1; j += 4 * kCThreads)
            __stcs(reinterpret_cast<float4*>(dst + j), *reinterpret_cast<const float4*>(s_pts + j));
        for (unsigned j = max(body1, body0) + t; j < end; j += kCThreads) __stcs(dst + j, s_pts[j]);
    } else {
        for (unsigned j = ph + t; j < end; j += kCThreads) __stcs(dst + j, s_pts[j]);
    }
}

// ---- SLC_ORDER_REFERENCE: tile = 8 adjacent columns x H rows --------------------------------
// dynamic shared memory: H validity bytes (one per row of the tile: bit c = column c valid), then
// H x 8 u16: the rank of pixel (row, column) inside its column
__host__ __device__ inline size_t cols_rank_offset(int H) { return ((size_t)H + 15) & ~(size_t)15; }
__host__ __device__ inline size_t cols_smem_bytes(int H) { return cols_rank_offset(H) + (size_t)H * 16; }

__global__ void __launch_bounds__(kCThreads)
compact_cols_kernel(const CompactArgs a)
{
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ float4 s_q[kColTileW * (kColChunk + 1)];
    __shared__ unsigned s_warp4[4][kCThreads / 32];
    __shared__ unsigned s_colbase[kColTileW];
    const int t = threadIdx.x, tile = blockIdx.x, stack = blockIdx.y;
    const int H = a.H, W = a.W, u0 = tile * kColTileW;
    unsigned char* s_rowbits = s_dyn;
    unsigned short* s_rowrank = reinterpret_cast<unsigned short*>(s_dyn + cols_rank_offset(H));
    const uint8_t* mask = a.mask + (long long)stack * a.npx;
    const float4* xyzw = a.xyzw + (long long)stack * a.npx;
    const int R = (H + kCThreads - 1) / kCThreads;            // rows per thread, contiguous

    // phase A: one validity byte per row (8 columns), per-thread column counts, block scan per column
    unsigned cnt[4] = {0u, 0u, 0u, 0u};                       // two 16-bit counters per word: columns (0,1) (2,3) (4,5) (6,7)
    auto spread = [](unsigned bits, unsigned (&d)[4]) {       // bit c -> +1 in the 16-bit lane of column c
        d[0] = (bits & 1u) | ((bits & 2u) << 15);
        d[1] = ((bits >> 2) & 1u) | ((bits & 8u) << 13);
        d[2] = ((bits >> 4) & 1u) | ((bits & 32u) << 11);
        d[3] = ((bits >> 6) & 1u) | ((bits & 128u) << 9);
    };
    const int vbeg = t * R, vend = min(H, vbeg + R);
    for (int v = vbeg; v < vend; v++) {
        const unsigned px = (unsigned)v * (unsigned)W + (unsigned)u0;
        const unsigned bits = nonzero_bits8(*reinterpret_cast<const uint2*>(mask + px));
        s_rowbits[v] = (unsigned char)bits;
        if (a.mask_bits != nullptr) a.mask_bits[(long long)stack * a.bits_stride + (px >> 3)] = (uint8_t)bits;
        unsigned d[4];
        spread(bits, d);
#pragma unroll
        for (int k = 0; k < 4; k++) cnt[k] += d[k];
    }
    // exclusive scan over threads of the four packed words (16-bit lanes cannot overflow: sums <= H < 65536)
    unsigned excl[4], tot[4];
    {
        const int lane = t & 31, warp = t >> 5;
        unsigned inc[4];
#pragma unroll
        for (int k = 0; k < 4; k++) inc[k] = cnt[k];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned o = __shfl_up_sync(0xFFFFFFFFu, inc[k], d);
                if (lane >= d) inc[k] += o;
            }
        }
        if (lane == 31) {
#pragma unroll
            for (int k = 0; k < 4; k++) s_warp4[k][warp] = inc[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned base = 0, sum = 0;
#pragma unroll
            for (int w = 0; w < kCThreads / 32; w++) {
                const unsigned sw = s_warp4[k][w];
                if (w < warp) base += sw;
                sum += sw;
            }
            excl[k] = base + inc[k] - cnt[k];
            tot[k] = sum;
        }
    }
    // the rank of every pixel of this thread's rows inside its column: one 16-byte store per row
    {
        unsigned run[4] = {excl[0], excl[1], excl[2], excl[3]};
        for (int v = vbeg; v < vend; v++) {
            *reinterpret_cast<uint4*>(s_rowrank + 8 * v) = make_uint4(run[0], run[1], run[2], run[3]);
            unsigned d[4];
            spread(s_rowbits[v], d);
#pragma unroll
            for (int k = 0; k < 4; k++) run[k] += d[k];
        }
    }
    // column totals -> exclusive prefix over the 8 columns, tile total
    unsigned colcnt[kColTileW];
#pragma unroll
    for (int k = 0; k < 4; k++) { colcnt[2 * k] = tot[k] & 0xFFFFu; colcnt[2 * k + 1] = tot[k] >> 16; }
    unsigned total = 0;
    unsigned colpre[kColTileW];
#pragma unroll
    for (int c = 0; c < kColTileW; c++) { colpre[c] = total; total += colcnt[c]; }
    if (t < 32) {
        const unsigned b = lookback_exclusive(a.state + (long long)stack * a.tiles, tile, total, a.epoch);
        if (t < kColTileW) s_colbase[t] = b + colpre[t];
        if (t == 0 && tile == a.tiles - 1) a.counts[stack] = (unsigned long long)b + total;
    }
    __syncthreads();

    // phase B: chunks of 128 rows.  Load: thread -> (row t>>3 (+32, +64, +96), column t&7): 128 contiguous
    // bytes per row.  Emit: thread -> (column t>>5, row t&31 (+32 ...)): a warp writes consecutive points.
    const int lr = t >> 3, lc = t & 7;
    const int ec = t >> 5, er = t & 31;
    float* out = a.points + (long long)stack * a.point_stride * 3;
    const unsigned colbase = s_colbase[ec];
    const unsigned room = (unsigned)min((long long)0xFFFFFFFFll, a.point_stride);   // npx < 2^31: 32-bit indices
    float4 q[4];
    bool ok[4];
    auto load_chunk = [&](int v0) {                     // issue the chunk's global loads (valid pixels only)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int v = v0 + lr + 32 * k;
            ok[k] = v < H && ((s_rowbits[v] >> lc) & 1u);
            if (ok[k]) q[k] = __ldcs(xyzw + ((unsigned)v * (unsigned)W + (unsigned)(u0 + lc)));
        }
    };
    load_chunk(0);
    for (int v0 = 0; v0 < H; v0 += kColChunk) {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (ok[k]) s_q[lc * (kColChunk + 1) + lr + 32 * k] = q[k];
        __syncthreads();
        if (v0 + kColChunk < H) load_chunk(v0 + kColChunk);   // in flight while this chunk is emitted
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int v = v0 + er + 32 * k;
            if (v < H && ((s_rowbits[v] >> ec) & 1u)) {
                const unsigned idx = colbase + s_rowrank[8 * v + ec];
                if (idx < room) {
                    const float4 p = s_q[ec * (kColChunk + 1) + er + 32 * k];
                    float* d = out + 3ull * idx;
                    __stcs(d, p.x); __stcs(d + 1, p.y); __stcs(d + 2, p.z);
                }
            }
        }
        __syncthreads();
    }
}

// ---- SLC_RESULT_DEPTH from finished maps: z plane + one validity bit per pixel ----------------
// (the first-frame kernel writes this layout itself; the dynamic frames, whose kernel walks a
// sequence, get it from their maps.)  A thread owns 8 consecutive pixels of one map.
__global__ void __launch_bounds__(kCThreads)
pack_depth_kernel(const float4* __restrict__ xyzw, const uint8_t* __restrict__ mask, long long npx,
                  float* __restrict__ depth, uint8_t* __restrict__ bits, long long bits_stride)
{
    const long long map = blockIdx.y;
    const long long p0 = ((long long)blockIdx.x * kCThreads + threadIdx.x) * 8;
    if (p0 >= npx) return;
    const float4* src = xyzw + map * npx + p0;
    const uint8_t* m = mask + map * npx + p0;
    float* dst = depth + map * npx + p0;
    unsigned b = 0u;
    const int n = (int)((npx - p0) < 8 ? (npx - p0) : 8);
    float z[8];
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (j < n) { z[j] = __ldcs(src + j).z; b |= (m[j] != 0 ? 1u : 0u) << j; }
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (j < n) __stcs(dst + j, z[j]);
    bits[map * bits_stride + (p0 >> 3)] = (uint8_t)b;
}

}  // namespace

cudaError_t launch_pack_depth(const float* d_xyzw, const uint8_t* d_mask, long long npx, int n_maps, float* d_depth,
                              uint8_t* d_bits, long long bits_stride, cudaStream_t stream)
{
    if (n_maps <= 0) return cudaSuccess;
    const long long groups = (npx + 7) / 8;
    for (int done = 0; done < n_maps; done += 65535) {
        const int n = (n_maps - done) < 65535 ? (n_maps - done) : 65535;
        pack_depth_kernel<<<dim3((unsigned)((groups + kCThreads - 1) / kCThreads), (unsigned)n), kCThreads, 0, stream>>>(
            reinterpret_cast<const float4*>(d_xyzw) + (size_t)done * npx, d_mask + (size_t)done * npx, npx,
            d_depth + (size_t)done * npx, d_bits + (size_t)done * bits_stride, bits_stride);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

bool compact_supported(int W, int H, const void* d_mask, const void* d_xyzw)
{
    // reference order keeps 17 bytes of shared memory per image row
    return (W % 8 == 0) && H < 65536 && cols_smem_bytes(H) <= 200 * 1024 && ((reinterpret_cast<uintptr_t>(d_mask) & 7) == 0) &&
           ((reinterpret_cast<uintptr_t>(d_xyzw) & 15) == 0);
}

int compact_tiles(int W, int H, int order)
{
    return order == 1 ? W / kColTileW : (int)(((long long)W * H + kRowTile - 1) / kRowTile);
}

size_t compact_state_bytes(int W, int H, int n_stacks)
{
    const int t0 = compact_tiles(W, H, 0), t1 = compact_tiles(W, H, 1);
    const int tiles = t0 > t1 ? t0 : t1;
    return (size_t)tiles * (size_t)n_stacks * sizeof(unsigned long long);
}

// One launch for n_stacks maps.  d_state: compact_state_bytes() bytes, zeroed once when allocated;
// epoch: a value never used before with this d_state (1, 2, 3, ...).
cudaError_t launch_compact(int W, int H, int n_stacks, int order, const float* d_xyzw, const uint8_t* d_mask,
                           float* d_points, long long point_stride, uint8_t* d_mask_bits, long long bits_stride,
                           unsigned long long* d_counts, unsigned long long* d_state, unsigned epoch,
                           cudaStream_t stream)
{
    if (n_stacks <= 0) return cudaSuccess;
    CompactArgs a{};
    a.W = W; a.H = H; a.npx = (long long)W * H;
    a.xyzw = reinterpret_cast<const float4*>(d_xyzw);
    a.mask = d_mask;
    a.points = d_points; a.point_stride = point_stride;
    a.mask_bits = d_mask_bits; a.bits_stride = bits_stride;
    a.counts = d_counts; a.state = d_state;
    a.tiles = compact_tiles(W, H, order);
    a.epoch = epoch & 0x3FFFFFFFu;
    for (int done = 0; done < n_stacks; done += 65535) {
        CompactArgs b = a;
        const int n = (n_stacks - done) < 65535 ? (n_stacks - done) : 65535;
        b.xyzw = a.xyzw + (size_t)done * a.npx;
        b.mask = a.mask + (size_t)done * a.npx;
        b.points = a.points + (size_t)done * point_stride * 3;
        if (a.mask_bits) b.mask_bits = a.mask_bits + (size_t)done * bits_stride;
        b.counts = a.counts + done;
        b.state = a.state + (size_t)done * a.tiles;
        if (order == 1) {
            const size_t smem = cols_smem_bytes(H);
            if (smem > 24 * 1024) {
                // static (17.7 KB) + dynamic shared memory beyond 48 KB needs the opt-in (images taller than ~1750 rows)
                const cudaError_t e = cudaFuncSetAttribute(compact_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
            }
            compact_cols_kernel<<<dim3((unsigned)a.tiles, (unsigned)n), kCThreads, smem, stream>>>(b);
        } else {
            compact_rows_kernel<<<dim3((unsigned)a.tiles, (unsigned)n), kCThreads, 0, stream>>>(b);
        }
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace slc
