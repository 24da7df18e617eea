// tsplat.cu -- hand-written sm_100a kernels + C ABI for topsy's SPH projection hot path.
// See include/tsplat.h for the boundary and DESIGN.md for the data layout / roofline of each kernel.
//
// Kernels
//   K1  k_project_stream<MODE>  (tsplat_project.cuh) one pass over the SoA particle arrays as a per-warp software pipeline:
//                               cp.async staging, rotate/project/cull, classify by projected footprint; footprints up to
//                               8 px are compacted and splatted with vector REDs (REDG.E.ADD.F32{,x2,x4}) into the
//                               L2-resident image, larger ones are appended to a 32-byte projected-record queue.
//   K2  k_bin_count / k_bin_scan / k_bin_fill      counting sort of the queue by 64x32-pixel tile
//   K3  k_tile_gather<MODE>     column strips in registers, LUT row addresses staged per (record, tile row),
//                               whole kernel LUT in shared memory, no atomics in the loop
//   K3b k_queue_atomic<MODE>    cooperative (warp / CTA per record) atomic splat: small calls, > 4096 px, pair overflow
//   K5  k_colormap              fused normalise + log/linear + LUT (1-D / 2-D) or tri-band gamma map -> RGBA8/16F/32F
//   K6  k_reduce_colormap       multi-GPU: image sum over NVLink peer memory fused with the colormap
//   K7  k_periodic_accumulate   PeriodicSPH replica sum;  K8 k_content_stats / k_content_select  device autorange
//   K9-K11 (tsplat_surface.cuh) surface mode: z-buffered splat, bilateral filter, lighting
//   K4  k_cell_* (tsplat_cells.cu)  CellLayout.from_positions: cell keys + stable counting sort
#include "tsplat_device.cuh"
#include "../../include/tsplat.h"

#include <cuda_fp16.h>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cmath>
#include <cstdlib>
#include <new>
#include <algorithm>

using namespace tsplat;

// ------------------------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int set_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return set_err(TSPLAT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                           __FILE__, __LINE__);                                                          \
    } while (0)

// ------------------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------------------
constexpr int TILE_W = 64, TILE_H = 32;     // tile of the gather path (pixels): 8 warps x (16 x 16) pixel blocks
constexpr int MAX_RANGES = 1 << 16;         // ranges per render call (reference: <= n_cells = 4096 per buffer)
constexpr int RANGE_SLOTS = 4;              // pinned staging ring for range tables
constexpr float DIRECT_MAX_WPX = 8.0f;      // footprints up to this width are splatted by the projecting thread
constexpr float HUGE_MIN_WPX = 4096.0f;     // footprints above this go to the cooperative atomic kernel (as does pair overflow)

constexpr int STAT_SLOTS = 256;             // power of two
constexpr int PAIR_STRIPES = 1024;          // power of two: a same-address ATOMG costs ~23 ns in the L2, serialised per address
struct StatSlot { unsigned long long culled_direct, reds; };   // culled in the low 32 bits, direct in the high 32

struct Counters {
    unsigned long long reds, culled, direct, tiled, huge, pairs;
    StatSlot slots[STAT_SLOTS];  // K1's per-warp statistics land here (spread to avoid same-address atomics)
    unsigned int q_count;        // deferred records in the queue (this call)
    unsigned int huge_count;     // records routed to the cooperative atomic kernel (this call)
    unsigned int pair_total;     // nonzero once any (particle, tile) pair was reserved (this call)
    unsigned int work_counter;   // persistent-kernel work ticket
    unsigned int n_segments;     // gather work units (this call)
    unsigned int call_particles; // particles of this call (written by K1; read back with q_count as the next call's hint)
    unsigned int pair_sub[PAIR_STRIPES];   // pair-capacity reservations, striped: same-address atomics serialise in the L2
};

struct RangeTable {              // device layout of a multi-range call
    const int64_t *start;        // [n]
    const int64_t *end;          // [n]
    const int64_t *gprefix;      // [n+1] exclusive prefix of 4-particle group counts
    int n;
};

struct tsplat_ctx {
    int device;
    int R;
    float *d_lut;                // LUT_TOTAL floats
    float *d_slut;               // local-sphere LUT of the surface mode (LUT_TOTAL floats)
    float density_cut;
    bool surface_set;
    bool lut_set, camera_set;
    Camera cam;
    const float *x, *y, *z, *h;
    int64_t n;
    const float *w0, *w1, *w2;
    float *image;
    int channels;
    void *scratch;
    int64_t scratch_bytes;
    Counters *d_counters;
    void *d_select;              // histograms / statistics of the device autorange (32 KB)
    unsigned *h_select;          // pinned staging of the same size (per context: contexts may be used from different threads)
    // range staging
    int64_t *h_ranges[RANGE_SLOTS];
    int64_t *d_ranges[RANGE_SLOTS];
    cudaEvent_t range_evt[RANGE_SLOTS];
    int range_slot;
    cudaStream_t last_stream;
    int64_t launches;
    int sm_count;
    bool bin_attr_set;
    bool gather_attr_set[4];
    size_t bilateral_smem_set;   // largest dynamic shared-memory size k_bilateral_filter has been opted into
    // length of the deferred queue after the previous render call of this context (read back asynchronously into pinned
    // memory): a call whose predecessor deferred only a few records skips the four binning / gather launches and sends its
    // own deferred records straight to the cooperative atomic kernel, which is correct for any number of records
    unsigned *h_qhint;           // [0] = q_count, [5] = call_particles of the most recent call whose read-back has landed
    // optional live timing of K1 (tsplat_enable_kernel_timing): ring of event pairs, [timing_read, timing_write) pending
    bool timing;
    cudaEvent_t t_begin[TSPLAT_TIMING_SLOTS], t_end[TSPLAT_TIMING_SLOTS];
    bool timing_events_created;
    int64_t timing_read, timing_write;
};

extern "C" const char *tsplat_last_error(void) { return g_err; }
extern "C" int tsplat_abi_version(void) { return TSPLAT_ABI_VERSION; }
extern "C" int tsplat_mode_channels(int mode)
{
    switch (mode) {
    case TSPLAT_MODE_DENSITY: return 1;
    case TSPLAT_MODE_WEIGHTED: return 2;
    case TSPLAT_MODE_RGB: return 4;
    case TSPLAT_MODE_DEPTH: return 2;
    case TSPLAT_MODE_SURFACE: return 2;
    default: return -1;
    }
}

// ------------------------------------------------------------------------------------------------------------
// image accumulation primitives: one RED per pixel (vector REDs need sm_90+; this file is sm_100a only)
// ------------------------------------------------------------------------------------------------------------
template <int MODE> struct ModeTraits;
template <> struct ModeTraits<TSPLAT_MODE_DENSITY> { static constexpr int C = 1; };
template <> struct ModeTraits<TSPLAT_MODE_WEIGHTED> { static constexpr int C = 2; };
template <> struct ModeTraits<TSPLAT_MODE_RGB> { static constexpr int C = 4; };
template <> struct ModeTraits<TSPLAT_MODE_DEPTH> { static constexpr int C = 2; };

// v0..v2 follow the Deferred convention; K is the kernel value.
template <int MODE>
__device__ __forceinline__ void red_pixel(float *__restrict__ image, size_t pix, float K, float v0, float v1, float v2)
{
    if (MODE == TSPLAT_MODE_DENSITY) {
        atomicAdd(image + pix, K * v0);
    } else if (MODE == TSPLAT_MODE_RGB) {
        atomicAdd(reinterpret_cast<float4 *>(image) + pix, make_float4(v0 * K, v1 * K, v2 * K, 1.0f));
    } else {  // WEIGHTED / DEPTH: (val, val * q|cz)
        const float val = K * v0;
        atomicAdd(reinterpret_cast<float2 *>(image) + pix, make_float2(val, val * v1));
    }
}

// ------------------------------------------------------------------------------------------------------------
// K1: project / cull / classify / direct splat
// ------------------------------------------------------------------------------------------------------------
struct ProjectArgs {
    const float *x, *y, *z, *h, *w0, *w1, *w2;
    Camera cam;
    int R;
    float *image;
    const float *lut;            // all levels (device)
    Deferred *queue;
    unsigned int queue_cap;
    Counters *counters;
    // single range (n_ranges == 1): particles [start, end); groups [g0, g0 + n_groups)
    int64_t start, end, g0, n_groups;
    int64_t n_total;             // particles in the buffers (the last 4-group may be partial)
    unsigned call_particles;     // particles submitted by this call
    int small_call;              // the call is too small for the tile binning: deferred records all go to K3b
    RangeTable table;            // used when table.n > 0
};

// L2 residency control: the particle stream is read exactly once (evict_first), the accumulation image is re-hit by
// every particle (evict_last), so the streaming reads must not push image lines out of the 126 MB L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// streaming 128-bit load of 4 consecutive particles of one SoA array (read once: no L1 allocation, L2 evict-first)
__device__ __forceinline__ float4 ld4(const float *__restrict__ p, int64_t group, uint64_t pol)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(reinterpret_cast<const float4 *>(p) + group), "l"(pol));
    return r;
}

__device__ __forceinline__ void red_v4(float *addr, float a, float b, float c, float d, uint64_t pol)
{
    asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                 :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d), "l"(pol) : "memory");
}

__device__ __forceinline__ void red_v2(float *addr, float a, float b, uint64_t pol)
{
    asm volatile("red.global.add.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" :: "l"(addr), "f"(a), "f"(b), "l"(pol) : "memory");
}

__device__ __forceinline__ void red_v1(float *addr, float a, uint64_t pol)
{
    asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(addr), "f"(a), "l"(pol) : "memory");
}

__device__ __forceinline__ float4 ld4_tail(const float *__restrict__ p, int64_t base, int64_t left)
{
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (left > 0) r.x = p[base];
    if (left > 1) r.y = p[base + 1];
    if (left > 2) r.z = p[base + 2];
    return r;
}

#include "tsplat_project.cuh"

// ------------------------------------------------------------------------------------------------------------
// K3b: cooperative atomic splat of queue records (one warp per record, lanes along pixel rows)
// ------------------------------------------------------------------------------------------------------------
struct QueueArgs {
    const Deferred *queue;
    const unsigned int *indices;     // optional indirection (huge list); nullptr -> identity
    const unsigned int *count;       // device pointer to the number of entries
    unsigned int cap;
    const float *lut;
    float *image;
    int R;
};

constexpr float COOP_MIN_WPX = 48.0f;        // above this a record is splatted by a whole CTA, below by one warp

// Splat one queue record with atomics.  The (row, column) grid of covered pixels is flattened over the `nthreads`
// cooperating threads (a warp or a whole CTA), so narrow footprints still fill all lanes.  The kernel LUT is read
// through the read-only L1 path (22 KB, hot): no shared-memory staging, no barrier before the first RED.
template <int MODE>
__device__ __forceinline__ void atomic_splat_record(const QueueArgs &a, const float4 q0, const float4 q1, unsigned tid,
                                                    unsigned nthreads)
{
    const float px0 = q0.x, px1 = q0.y, py0 = q0.z, py1 = q0.w, wpx = q1.x;
    const float inv = 1.0f / wpx;
    int j0, j1, k0, k1;
    pixel_range(px0, px1, a.R, j0, j1);
    pixel_range(py0, py1, a.R, k0, k1);
    if (j1 < j0 || k1 < k0) return;
    const unsigned ncols = (unsigned)(j1 - j0 + 1), total = ncols * (unsigned)(k1 - k0 + 1);
    const float rcols = 1.0f / (float)ncols;
    for (unsigned t = tid; t < total; t += nthreads) {
        unsigned dk = (unsigned)((float)t * rcols);               // t / ncols up to rounding; fixed below
        unsigned dj = t - dk * ncols;
        if ((int)dj < 0) { --dk; dj += ncols; } else if (dj >= ncols) { ++dk; dj -= ncols; }
        const int k = k0 + (int)dk, j = j0 + (int)dj;
        const float K = sample_lut(a.lut, wpx, inv, px0, py1, (float)j + 0.5f, (float)k + 0.5f);
        if (MODE != TSPLAT_MODE_RGB && K == 0.0f) continue;
        red_pixel<MODE>(a.image, (size_t)k * a.R + j, K, q1.y, q1.z, q1.w);
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) k_queue_atomic(const QueueArgs a)
{
    const unsigned count = min(*a.count, a.cap);
    if (blockIdx.x >= count) return;      // neither pass has work for this CTA
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // pass 1: one warp per record (modest footprints)
    for (unsigned w = blockIdx.x * 8u + warp; w < count; w += gridDim.x * 8u) {
        const unsigned idx = a.indices ? a.indices[w] : w;
        const float4 q1 = __ldg(reinterpret_cast<const float4 *>(a.queue + idx) + 1);
        if (q1.x > COOP_MIN_WPX) continue;
        atomic_splat_record<MODE>(a, __ldg(reinterpret_cast<const float4 *>(a.queue + idx)), q1, lane, 32u);
    }
    // pass 2: one CTA per record (big footprints)
    for (unsigned w = blockIdx.x; w < count; w += gridDim.x) {
        const unsigned idx = a.indices ? a.indices[w] : w;
        const float4 q1 = __ldg(reinterpret_cast<const float4 *>(a.queue + idx) + 1);
        if (!(q1.x > COOP_MIN_WPX)) continue;
        atomic_splat_record<MODE>(a, __ldg(reinterpret_cast<const float4 *>(a.queue + idx)), q1, threadIdx.x, 256u);
    }
}

// ------------------------------------------------------------------------------------------------------------
// K2: tile binning of the deferred queue (counting sort by 16x16-pixel tile, all on the device)
// ------------------------------------------------------------------------------------------------------------
constexpr int SEG = 1024;                     // pairs per gather work unit
constexpr unsigned ROUTE_TILED = 1, ROUTE_HUGE = 2;

struct BinArgs {
    const Deferred *queue;
    unsigned queue_cap;
    Counters *counters;
    unsigned char *route;            // [queue_cap]
    unsigned int *huge_idx;          // [queue_cap]
    unsigned int *tile_count;        // [nt]
    unsigned int *tile_offset;       // [nt + 1]
    unsigned int *tile_cursor;       // [nt]
    unsigned int *seg_prefix;        // [nt + 1]
    unsigned int *seg_tile;          // [seg_cap] tile of each gather work unit
    unsigned int seg_cap;
    unsigned int *pairs;             // [pairs_cap]
    unsigned int pairs_cap;
    int R, ntx, nt;
};

__device__ __forceinline__ void tile_range(const Deferred &d, int R, int &tx0, int &tx1, int &ty0, int &ty1)
{
    int j0, j1, k0, k1;
    pixel_range(d.px0, d.px1, R, j0, j1);
    pixel_range(d.py0, d.py1, R, k0, k1);
    if (j1 < j0 || k1 < k0) { tx0 = ty0 = 1; tx1 = ty1 = 0; return; }
    tx0 = j0 / TILE_W; tx1 = j1 / TILE_W; ty0 = k0 / TILE_H; ty1 = k1 / TILE_H;
}

constexpr unsigned SMALL_QUEUE = 131072;      // below this many deferred records the binning machinery is not worth it

// Pass 1: route every deferred record (tile gather / cooperative atomics) and histogram the (record, tile) pairs.
// The histogram is privatised in shared memory (native integer ATOMS) and flushed with one global RED per
// non-empty (CTA, tile): global atomics on a few thousand hot tile counters would serialise in the L2.
// USE_SMEM == false is the fallback for resolutions whose tile table does not fit in shared memory.
template <bool USE_SMEM>
__global__ void __launch_bounds__(1024) k_bin_count(const BinArgs a)
{
    extern __shared__ unsigned s_hist[];
    const unsigned Q = min(a.counters->q_count, a.queue_cap);
    if (USE_SMEM) {
        for (int i = threadIdx.x; i < a.nt; i += blockDim.x) s_hist[i] = 0u;
        __syncthreads();
    }
    const bool all_atomic = Q < SMALL_QUEUE;
    // contiguous chunk per CTA (k_bin_fill uses the same partition)
    const unsigned per = (Q + gridDim.x - 1) / gridDim.x;
    const unsigned r0 = blockIdx.x * per, r1 = min(Q, r0 + per);
    unsigned n_tiled = 0, n_huge = 0, n_pairs = 0;
    const int lane = threadIdx.x & 31;
    for (unsigned rb = r0 + (threadIdx.x & ~31u); rb < r1; rb += blockDim.x) {      // warp-uniform trip count
        const unsigned r = rb + lane;
        const bool valid = r < r1;
        Deferred d;
        d.wpx = 0.0f;
        int tx0 = 1, tx1 = 0, ty0 = 1, ty1 = 0;
        if (valid) { d = a.queue[r]; tile_range(d, a.R, tx0, tx1, ty0, ty1); }
        const bool covers = valid && tx1 >= tx0 && ty1 >= ty0;
        const unsigned np = covers ? (unsigned)(tx1 - tx0 + 1) * (unsigned)(ty1 - ty0 + 1) : 0u;
        bool tiled = covers && !all_atomic && d.wpx <= HUGE_MIN_WPX;
        // one pair-capacity reservation per warp (a same-address ATOMG per record would serialise in the L2)
        unsigned want = tiled ? np : 0u, incl = want;
        for (int s = 1; s < 32; s <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= s) incl += t;
        }
        const unsigned warp_total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned base = 0;
        // capacity check only (positions come from the scan): each warp reserves from one of PAIR_STRIPES equal shares
        if (lane == 0 && warp_total) {
            base = atomicAdd(&a.counters->pair_sub[(blockIdx.x * 32u + (threadIdx.x >> 5)) & (PAIR_STRIPES - 1)], warp_total);
            if (base == 0u) a.counters->pair_total = 1u;
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (tiled) {
            const unsigned end = base + incl;
            tiled = end <= a.pairs_cap / PAIR_STRIPES && end >= base;
        }
        if (valid) {
            unsigned route = 0;
            if (tiled) {
                route = ROUTE_TILED;
                ++n_tiled; n_pairs += np;
                for (int ty = ty0; ty <= ty1; ++ty)
                    for (int tx = tx0; tx <= tx1; ++tx) {
                        if (USE_SMEM) atomicAdd(&s_hist[ty * a.ntx + tx], 1u);
                        else atomicAdd(&a.tile_count[ty * a.ntx + tx], 1u);
                    }
            } else if (covers) {
                route = ROUTE_HUGE;
                ++n_huge;
            } else {
                ++n_tiled;      // covers no pixel centre inside the image
            }
            a.route[r] = (unsigned char)route;
        }
        // warp-aggregated append to the cooperative-atomic list
        const unsigned hm = __ballot_sync(0xffffffffu, valid && covers && !tiled);
        if (hm) {
            unsigned hb = 0;
            if (lane == 0) hb = atomicAdd(&a.counters->huge_count, (unsigned)__popc(hm));
            hb = __shfl_sync(0xffffffffu, hb, 0);
            if (valid && covers && !tiled) a.huge_idx[hb + __popc(hm & ((1u << lane) - 1u))] = r;
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        n_tiled += __shfl_down_sync(0xffffffffu, n_tiled, d);
        n_huge += __shfl_down_sync(0xffffffffu, n_huge, d);
        n_pairs += __shfl_down_sync(0xffffffffu, n_pairs, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_tiled) atomicAdd(&a.counters->tiled, (unsigned long long)n_tiled);
        if (n_huge) atomicAdd(&a.counters->huge, (unsigned long long)n_huge);
        if (n_pairs) atomicAdd(&a.counters->pairs, (unsigned long long)n_pairs);
    }
    if (USE_SMEM) {
        __syncthreads();
        for (int i = threadIdx.x; i < a.nt; i += blockDim.x) {
            const unsigned v = s_hist[i];
            if (v) atomicAdd(&a.tile_count[i], v);
        }
    }
}

// single CTA: exclusive scans of the tile counts (pair offsets) and of the per-tile segment counts (work units),
// plus the work-unit -> tile table the gather kernel indexes with its ticket
__global__ void __launch_bounds__(1024) k_bin_scan(const BinArgs a)
{
    constexpr int PER = 8, CHUNK = 1024 * PER;
    __shared__ unsigned s_val[CHUNK];            // tile counts of the current chunk (coalesced staging)
    __shared__ unsigned s_cnt[1024], s_seg[1024];
    unsigned carry_c = 0, carry_g = 0;           // running totals (identical in every thread)
    if (a.counters->pair_total == 0u) {          // nothing was routed to the tiles (e.g. small queue -> atomic path)
        if (threadIdx.x == 0) { a.counters->n_segments = 0u; a.counters->work_counter = 0u; }
        return;
    }
    for (int c0 = 0; c0 < a.nt; c0 += CHUNK) {
        const int n = min(CHUNK, a.nt - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < CHUNK; i += 1024) s_val[i] = i < n ? a.tile_count[c0 + i] : 0u;
        __syncthreads();
        unsigned c = 0, g = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const unsigned t = s_val[threadIdx.x * PER + i];
            c += t; g += (t + SEG - 1) / SEG;
        }
        s_cnt[threadIdx.x] = c; s_seg[threadIdx.x] = g;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {     // Hillis-Steele inclusive scan over the 1024 partials
            unsigned tc = 0, tg = 0;
            if ((int)threadIdx.x >= d) { tc = s_cnt[threadIdx.x - d]; tg = s_seg[threadIdx.x - d]; }
            __syncthreads();
            s_cnt[threadIdx.x] += tc; s_seg[threadIdx.x] += tg;
            __syncthreads();
        }
        unsigned ec = carry_c + s_cnt[threadIdx.x] - c, eg = carry_g + s_seg[threadIdx.x] - g;     // exclusive
        carry_c += s_cnt[1023]; carry_g += s_seg[1023];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int idx = threadIdx.x * PER + i;
            const unsigned t = s_val[idx];
            const unsigned ns = (t + SEG - 1) / SEG;
            if (idx < n) {
                a.tile_offset[c0 + idx] = ec; a.seg_prefix[c0 + idx] = eg;
                for (unsigned q = 0; q < ns; ++q)
                    if (eg + q < a.seg_cap) a.seg_tile[eg + q] = (unsigned)(c0 + idx);
            }
            ec += t; eg += ns;
        }
    }
    if (threadIdx.x == 0) {
        a.tile_offset[a.nt] = carry_c;
        a.seg_prefix[a.nt] = carry_g;
        a.counters->n_segments = min(carry_g, a.seg_cap);
        a.counters->work_counter = 0u;
    }
}

// Pass 2: scatter record indices into their tiles.  Each CTA re-histograms its chunk in shared memory, reserves one
// contiguous run per non-empty tile with a single global atomic, then hands out slots with shared-memory atomics.
template <bool USE_SMEM>
__global__ void __launch_bounds__(1024) k_bin_fill(const BinArgs a)
{
    extern __shared__ unsigned s_mem[];
    unsigned *s_base = s_mem, *s_cur = s_mem + a.nt;
    const unsigned Q = min(a.counters->q_count, a.queue_cap);
    const unsigned per = (Q + gridDim.x - 1) / gridDim.x;
    const unsigned r0 = blockIdx.x * per, r1 = min(Q, r0 + per);
    // both sweeps are latency-bound streaming reads of (route byte, first half of the record): four records per thread
    // are loaded before any is processed
    constexpr int UNROLL = 4;
    auto sweep = [&](auto &&per_tile) {
        for (unsigned rb = r0 + threadIdx.x; rb < r1; rb += UNROLL * blockDim.x) {
            unsigned char route[UNROLL];
            float4 q0[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const unsigned r = rb + u * blockDim.x;
                route[u] = 0;
                if (r < r1) { route[u] = a.route[r]; q0[u] = *reinterpret_cast<const float4 *>(a.queue + r); }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (route[u] != ROUTE_TILED) continue;
                Deferred d; d.px0 = q0[u].x; d.px1 = q0[u].y; d.py0 = q0[u].z; d.py1 = q0[u].w;
                int tx0, tx1, ty0, ty1;
                tile_range(d, a.R, tx0, tx1, ty0, ty1);
                for (int ty = ty0; ty <= ty1; ++ty)
                    for (int tx = tx0; tx <= tx1; ++tx) per_tile(ty * a.ntx + tx, rb + u * blockDim.x);
            }
        }
    };
    if (USE_SMEM) {
        for (int i = threadIdx.x; i < 2 * a.nt; i += blockDim.x) s_mem[i] = 0u;
        __syncthreads();
        sweep([&](int t, unsigned) { atomicAdd(&s_cur[t], 1u); });
        __syncthreads();
        for (int i = threadIdx.x; i < a.nt; i += blockDim.x) {
            const unsigned v = s_cur[i];
            if (v) s_base[i] = a.tile_offset[i] + atomicAdd(&a.tile_cursor[i], v);
            s_cur[i] = 0u;
        }
        __syncthreads();
    }
    sweep([&](int t, unsigned r) {
        if (USE_SMEM) a.pairs[s_base[t] + atomicAdd(&s_cur[t], 1u)] = r;
        else a.pairs[a.tile_offset[t] + atomicAdd(&a.tile_cursor[t], 1u)] = r;
    });
}

// ------------------------------------------------------------------------------------------------------------
// K3: tile gather -- pixel-column strips in registers, no atomics inside the loop
// ------------------------------------------------------------------------------------------------------------
// A tile is TILE_W x TILE_H = 64 x 32 pixels; persistent 256-thread CTAs pull (tile, <= SEG-pair segment) tickets.
// Each warp owns a 16 x 16 block of the tile and each lane ONE pixel column of 8 rows in it (column = lane & 15,
// rows (lane >> 4) * 8 .. + 7), with the 8 x C partial sums in registers.
//
// Nearest-texel sampling (every footprint below 64 px) is separable in the index: K[k][j] = T[iv(k)][iu(j)].
//   * iv(k) is the same for every pixel of an image row, so the CTA computes it ONCE per (record, tile row) while
//     staging a batch, as a ready-made 16-bit shared-memory address of the LUT row (rows outside the footprint point
//     at a row of zeros);
//   * iu(j) is computed by the lane once per record for its column (columns outside the footprint select a zero
//     column: every LUT level is stored with one extra zero row and column);
//   * a pixel update is then  ld.shared [row_addr[k] + col_term]  +  one FFMA per channel: no per-pixel coverage
//     tests, no per-pixel float->int conversions; two row addresses are advanced by one packed 2 x 16-bit add.
// Levels 1..3 are stored twice, interleaved (texel = 8 bytes): the lower half-warp reads the even words, the upper
// half-warp (8 rows further down) the odd ones, so the two halves can never collide on a bank.
// Bilinear records (footprint >= 64 px) take a warp-uniform slow branch with the full sampler.
// Per-warp lists (built with shared-memory atomics while staging) keep warps off records that miss their block.
struct GatherArgs {
    const Deferred *queue;
    const unsigned int *pairs;
    const unsigned int *tile_offset;
    const unsigned int *seg_prefix;
    const unsigned int *seg_tile;
    Counters *counters;
    const float *lut;
    float *image;
    int R, ntx, nt;
};

// shared-memory copy of the kernel LUT (float offsets): level 0 = 65 x 65 floats (row 64 / column 64 zero);
// level l >= 1 (n = 64 >> l) = (n + 1) x (n + 1) texels of 2 identical floats (row n / column n zero).
constexpr int PLUT_FLOATS = 7144;             // 4226 + 2*33*33 + 2*17*17 + 2*9*9
__host__ __device__ constexpr int plut_base(int level) { return level == 0 ? 0 : level == 1 ? 4226 : level == 2 ? 6404 : 6982; }
__host__ __device__ constexpr int plut_row_bytes(int level) { return level == 0 ? 260 : 8 * ((64 >> level) + 1); }
constexpr int G_BATCH = 128;                  // records staged per round
constexpr int GATHER_CTAS_PER_SM = 3;         // sizeof(GatherSmem) = 70 KB

struct GatherSmem {
    float lut[PLUT_FLOATS];
    float4 a[G_BATCH];                        // px0 px1 scale v0      (scale = inv * n, or inv when bilinear)
    float4 y[G_BATCH];                        // py0 py1 scale packed(level-base address | n << 16 | row bytes << 23)
    float2 v[G_BATCH];                        // v1 v2
    // per (record, tile row) LUT row data, 256 bytes per record:
    //   nearest-texel records use the first 64 bytes: 32 x u16 shared-memory address of the LUT row
    //   bilinear records use all of it: 32 x { u32 address of texel row b0 | address of b1 << 16 | reuse flags, f32 fv }
    uint2 row[G_BATCH][TILE_H];
    unsigned list[8][G_BATCH];                // per-warp record lists: record * 16 | n << 16   (n == 0: bilinear)
    unsigned nlist[16];                       // [0..7] nearest-texel entries (from the front), [8..15] bilinear (from the back)
    unsigned work[4];                         // tile, first pair, pair count
};
static_assert(sizeof(GatherSmem) * GATHER_CTAS_PER_SM <= 227 * 1024, "gather shared memory");

__device__ __forceinline__ float lds_f32(unsigned addr)
{
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));     // the LUT is read-only after the first barrier
    return v;
}

// Eight LUT gathers of one lane, predicated on `p` (the lane's column is inside the footprint).  Lanes that are off do
// not take part in the access, so they cannot cause bank conflicts; their K registers keep an earlier (finite) LUT
// value that the caller multiplies by zero.  The LUT is read-only after the first barrier (no memory clobber needed).
__device__ __forceinline__ void lds8_pred(float (&K)[8], const unsigned (&ad)[8], bool p)
{
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %16, 0;\n\t"
        "@p ld.shared.f32 %0, [%8];\n\t@p ld.shared.f32 %1, [%9];\n\t@p ld.shared.f32 %2, [%10];\n\t@p ld.shared.f32 %3, [%11];\n\t"
        "@p ld.shared.f32 %4, [%12];\n\t@p ld.shared.f32 %5, [%13];\n\t@p ld.shared.f32 %6, [%14];\n\t@p ld.shared.f32 %7, [%15];\n\t}"
        : "+f"(K[0]), "+f"(K[1]), "+f"(K[2]), "+f"(K[3]), "+f"(K[4]), "+f"(K[5]), "+f"(K[6]), "+f"(K[7])
        : "r"(ad[0]), "r"(ad[1]), "r"(ad[2]), "r"(ad[3]), "r"(ad[4]), "r"(ad[5]), "r"(ad[6]), "r"(ad[7]), "r"((unsigned)p));
}

template <int MODE>
__device__ __forceinline__ void gather_accumulate(float (&acc)[8][ModeTraits<MODE>::C], int k, float K, float v0, float v1,
                                                  float v2, float cnt)
{
    constexpr int C = ModeTraits<MODE>::C;
    acc[k][0] = fmaf(K, v0, acc[k][0]);
    if (C >= 2) acc[k][1 % C] = fmaf(K, v1, acc[k][1 % C]);          // WEIGHTED / DEPTH: v1 = v0 * (q | cz)
    if (C == 4) { acc[k][2 % C] = fmaf(K, v2, acc[k][2 % C]); acc[k][3 % C] += cnt; }
}

template <int MODE>
__global__ void __launch_bounds__(256, GATHER_CTAS_PER_SM) k_tile_gather(const GatherArgs a)
{
    constexpr int C = ModeTraits<MODE>::C;
    extern __shared__ __align__(16) unsigned char g_smem_raw[];
    GatherSmem &S = *reinterpret_cast<GatherSmem *>(g_smem_raw);
    const unsigned n_seg = a.counters->n_segments;
    if (n_seg == 0u) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < PLUT_FLOATS; i += 256) S.lut[i] = 0.0f;
    __syncthreads();
    for (int i = tid; i < 64 * 64; i += 256) S.lut[(i >> 6) * 65 + (i & 63)] = a.lut[i];
#pragma unroll
    for (int level = 1; level < 4; ++level) {
        const int n = 64 >> level, sh = 6 - level;
        for (int i = tid; i < n * n; i += 256) {
            const float t = a.lut[lut_offset(level) + i];
            reinterpret_cast<float2 *>(S.lut + plut_base(level))[(i >> sh) * (n + 1) + (i & (n - 1))] = make_float2(t, t);
        }
    }
    const unsigned lut_sa = (unsigned)__cvta_generic_to_shared(S.lut);      // < 64 KB: row addresses fit 16 bits
    if (lut_sa + 260u * 64u >= 0x8000u) __trap();     // level-0 row addresses carry flags in bits 15 / 31 (never fires: lut is first)
    const int lx = (warp & 3) * 16 + (lane & 15);         // pixel column inside the tile
    const int ly0 = (warp >> 2) * 16 + (lane >> 4) * 8;   // first of the lane's 8 rows inside the tile
    const unsigned half4 = (unsigned)(lane >> 4) * 4u;    // which copy of an interleaved texel this half-warp reads
    const unsigned *my_list = S.list[warp];
    float Kreg[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // LUT values of the current record (see lds8_pred)
    float Kreg2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};    // ... and of the second record in flight

    for (;;) {
        __syncthreads();                                  // work / batch buffers free (and LUT loaded on round 0)
        if (tid == 0) {
            const unsigned ticket = atomicAdd(&a.counters->work_counter, 1u);
            if (ticket < n_seg) {
                const unsigned l = a.seg_tile[ticket];
                const unsigned first = a.tile_offset[l] + (ticket - a.seg_prefix[l]) * SEG;
                const unsigned end = a.tile_offset[l + 1];
                S.work[0] = l; S.work[1] = first; S.work[2] = min((unsigned)SEG, end - first);
            } else {
                S.work[0] = 0xffffffffu;
            }
        }
        if (tid < 16) S.nlist[tid] = 0u;
        __syncthreads();
        const unsigned tile = S.work[0];
        if (tile == 0xffffffffu) break;
        const unsigned first = S.work[1], count = S.work[2];
        const int tx = (int)(tile % (unsigned)a.ntx), ty = (int)(tile / (unsigned)a.ntx);
        const int px = tx * TILE_W + lx, py0 = ty * TILE_H + ly0;
        const float fx = (float)px + 0.5f;
        const float tcx = (float)(tx * TILE_W) + 0.5f, tcy = (float)(ty * TILE_H) + 0.5f;   // first pixel centre of the tile
        float acc[8][C];
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[k][c] = 0.0f;

        for (unsigned b0 = 0; b0 < count; b0 += G_BATCH) {
            const unsigned nb = min((unsigned)G_BATCH, count - b0);
            if (b0) {
                __syncthreads();
                if (tid < 16) S.nlist[tid] = 0u;
                __syncthreads();
            }
            // ---- stage: one record per lane 0..15; warp w owns records 16w .. 16w+15 of the batch, so that the row
            // staging below only needs a warp-level barrier ---------------------------------------------------------
            const unsigned rec = (unsigned)warp * 16u + (unsigned)lane;
            if (lane < 16 && rec < nb) {
                const unsigned ridx = a.pairs[first + b0 + rec];
                const float4 q0 = reinterpret_cast<const float4 *>(a.queue + ridx)[0];      // px0 px1 py0 py1
                const float4 q1 = reinterpret_cast<const float4 *>(a.queue + ridx)[1];      // wpx v0 v1 v2
                const float wpx = q1.x, inv = 1.0f / wpx;
                unsigned n, packed; float scale;
                if (wpx >= 64.0f) { n = 0u; packed = 0u; scale = inv; }
                else {
                    const int level = wpx > LEVEL_T0 ? 0 : wpx > LEVEL_T1 ? 1 : wpx > LEVEL_T2 ? 2 : 3;
                    n = 64u >> level;
                    scale = inv * (float)n;
                    packed = (lut_sa + 4u * (unsigned)plut_base(level)) | (n << 16) | ((unsigned)plut_row_bytes(level) << 23);
                }
                S.a[rec] = make_float4(q0.x, q0.y, scale, q1.y);
                S.y[rec] = make_float4(q0.z, q0.w, scale, __uint_as_float(packed));
                if (C == 2) S.v[rec] = make_float2(q1.y * q1.z, 0.0f);       // (m/h^2) * (q | cz)
                if (C == 4) S.v[rec] = make_float2(q1.z, q1.w);
                const unsigned entry = (rec << 4) | (n << 16);
                // which of the 8 warp blocks (16 x 16 pixel centres each) does the record's span touch?
#pragma unroll
                for (int w = 0; w < 8; ++w) {
                    const float x0 = tcx + (float)((w & 3) * 16), y0 = tcy + (float)((w >> 2) * 16);
                    if (x0 < q0.y && x0 + 15.0f >= q0.x && y0 < q0.w && y0 + 15.0f >= q0.z) {
                        if (n) S.list[w][atomicAdd(&S.nlist[w], 1u)] = entry;
                        else S.list[w][G_BATCH - 1 - atomicAdd(&S.nlist[8 + w], 1u)] = entry;
                    }
                }
            }
            __syncwarp();
            // ---- stage: LUT row data of every (record, tile row) of the warp's own records; lane = tile row ----
            {
                const float fy = tcy + (float)lane;
                const unsigned r_end = min(nb, (unsigned)warp * 16u + 16u);
                for (unsigned r = (unsigned)warp * 16u; r < r_end; ++r) {
                    const float4 Y = S.y[r];
                    const unsigned packed = __float_as_uint(Y.w);
                    const int n = (int)((packed >> 16) & 127u);
                    const bool ok = fy >= Y.x && fy < Y.y;
                    if (n) {
                        const int iv = min(__float2int_rd((Y.y - fy) * Y.z), n - 1);
                        reinterpret_cast<unsigned short *>(S.row[r])[lane] =
                            (unsigned short)((packed & 0xffffu) + (unsigned)(ok ? iv : n) * (packed >> 23));
                    } else {
                        // bilinear on level 0 (row stride 260 bytes, row 64 is zero): the row half of sample_lut()
                        const float tv = fmaf((Y.y - fy) * Y.z, 64.0f, -0.5f);
                        const float ivf = floorf(tv);
                        const int ib = (int)ivf;
                        const unsigned a0 = lut_sa + 260u * (unsigned)(ok ? min(max(ib, 0), 63) : 64);
                        const unsigned a1 = lut_sa + 260u * (unsigned)(ok ? min(max(ib + 1, 0), 63) : 64);
                        // reuse flags for the lane that walks rows k-1, k in the same group of 8 (see the accumulate loop):
                        // bit 15: both texel rows are those of the previous pixel row; bit 31: b1 is the previous row's b0
                        const unsigned p0 = __shfl_up_sync(0xffffffffu, a0, 1), p1 = __shfl_up_sync(0xffffffffu, a1, 1);
                        unsigned w = a0 | (a1 << 16);
                        if ((lane & 7) != 0) {
                            if (a0 == p0 && a1 == p1) w |= 0x8000u;
                            else if (a1 == p0) w |= 0x80000000u;
                        }
                        S.row[r][lane] = make_uint2(w, __float_as_uint(tv - ivf));
                    }
                }
            }
            __syncthreads();
            // ---- accumulate: every warp walks its own lists ------------------------------------------------
            // nearest-texel records (front of the list): branch-free body, two records in flight to cover LDS latency
            auto nearest = [&](const unsigned e, float (&K)[8]) {
                const unsigned r16 = e & 0xfff0u;
                const int n = (int)(e >> 16);
                const float4 A = *reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(S.a) + r16);
                const uint4 rw = *reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(S.row) + r16 * 16u + ly0 * 2);
                const bool colok = fx >= A.x && fx < A.y;
                float v1 = 0.0f, v2 = 0.0f;
                if (C >= 2) { const float2 V = *reinterpret_cast<const float2 *>(reinterpret_cast<const char *>(S.v) + (r16 >> 1)); v1 = V.x; v2 = V.y; }
                const int iu = min(__float2int_rd((fx - A.x) * A.z), n - 1);      // garbage (unused) where !colok
                const unsigned col = n == 64 ? (unsigned)iu * 4u : (unsigned)iu * 8u + half4;
                const unsigned col2 = col * 0x10001u;                 // the same column term for both packed rows
                const unsigned w4[4] = {rw.x, rw.y, rw.z, rw.w};
                unsigned ad[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const unsigned X = w4[q] + col2;                  // no carry: both halves stay below 64 K
                    ad[2 * q] = X & 0xffffu; ad[2 * q + 1] = X >> 16;
                }
                lds8_pred(K, ad, colok);
                const float m0 = colok ? A.w : 0.0f, m1 = colok ? v1 : 0.0f, m2 = colok ? v2 : 0.0f;
                float2 Yr = make_float2(0.f, 0.f);
                if (C == 4) { const float4 Y = *reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(S.y) + r16); Yr = make_float2(Y.x, Y.y); }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float cnt = 0.0f;
                    if (C == 4) {                                     // RGB counts every covered pixel, even where K == 0
                        const float fy = (float)(py0 + k) + 0.5f;
                        cnt = (colok && fy >= Yr.x && fy < Yr.y) ? 1.0f : 0.0f;
                    }
                    gather_accumulate<MODE>(acc, k, K[k], m0, m1, m2, cnt);
                }
            };
            const unsigned nl = S.nlist[warp];
            unsigned i = 0;
            for (; i + 2 <= nl; i += 2) {
                const uint2 ee = *reinterpret_cast<const uint2 *>(my_list + i);
                nearest(ee.x, Kreg);
                nearest(ee.y, Kreg2);
            }
            if (i < nl) nearest(my_list[i], Kreg);
            // bilinear records (back of the list): footprints of 64 px and more magnify level 0.  The bilinear weights are
            // separable: the column part (two texel columns + fraction fu) is computed once per record by the lane, the
            // row part (two texel rows + fraction fv) once per (record, tile row) while staging.  H(b) = lerp of texel row
            // b at the lane's column; consecutive pixel rows mostly share texel rows (texel pitch >= 1 px), so H values
            // are carried from row to row and only reloaded when the staged flags say so.  Same operations and FMA
            // placement as sample_lut(): bit-identical samples.
            const unsigned nlb = S.nlist[8 + warp];
            for (unsigned ib = 0; ib < nlb; ++ib) {
                const unsigned r16 = my_list[G_BATCH - 1 - ib] & 0xfff0u;
                const float4 A = *reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(S.a) + r16);
                const bool colok = fx >= A.x && fx < A.y;
                float v1 = 0.0f, v2 = 0.0f;
                if (C >= 2) { const float2 V = *reinterpret_cast<const float2 *>(reinterpret_cast<const char *>(S.v) + (r16 >> 1)); v1 = V.x; v2 = V.y; }
                const float tu = fmaf((fx - A.x) * A.z, 64.0f, -0.5f);
                const float iuf = floorf(tu);
                const float fu = tu - iuf;
                const int ia = (int)iuf;
                const unsigned c0 = 4u * (unsigned)min(max(ia, 0), 63), c1 = 4u * (unsigned)min(max(ia + 1, 0), 63);
                const float m0 = colok ? A.w : 0.0f, m1 = colok ? v1 : 0.0f, m2 = colok ? v2 : 0.0f;
                const uint4 *rows = reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(S.row) + r16 * 16u + ly0 * 8);
                float Htop = 0.0f, Hbot = 0.0f;
                const uint4 rr4[4] = {rows[0], rows[1], rows[2], rows[3]};     // all 8 rows up front: one LDS latency, not four
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) {
                    const uint4 rr = rr4[k2];                           // two pixel rows: {w, fv} {w, fv}
                    const unsigned ws[2] = {rr.x, rr.z};
                    const float fvs[2] = {__uint_as_float(rr.y), __uint_as_float(rr.w)};
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const unsigned w = ws[h];
                        if (!(w & 0x8000u)) {                           // the pixel row needs (at least) a new top texel row
                            const unsigned a0 = w & 0x7fffu, a1 = (w >> 16) & 0x7fffu;
                            if (w & 0x80000000u) Hbot = Htop;
                            else { const float t0 = lds_f32(a1 + c0), t1 = lds_f32(a1 + c1); Hbot = fmaf(fu, t1 - t0, t0); }
                            const float t0 = lds_f32(a0 + c0), t1 = lds_f32(a0 + c1);
                            Htop = fmaf(fu, t1 - t0, t0);
                        }
                        const float K = fmaf(fvs[h], Hbot - Htop, Htop);
                        float cnt = 0.0f;
                        if (C == 4) cnt = (colok && (w & 0x7fffu) != lut_sa + 260u * 64u) ? 1.0f : 0.0f;
                        gather_accumulate<MODE>(acc, 2 * k2 + h, K, m0, m1, m2, cnt);
                    }
                }
            }
        }
        if (px < a.R) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int py = py0 + k;
                if (py >= a.R) break;
                const size_t pix = (size_t)py * a.R + px;
                if (C == 1) { if (acc[k][0] != 0.0f) atomicAdd(a.image + pix, acc[k][0]); }
                else if (C == 2) {
                    if (acc[k][0] != 0.0f || acc[k][1 % C] != 0.0f)
                        atomicAdd(reinterpret_cast<float2 *>(a.image) + pix, make_float2(acc[k][0], acc[k][1 % C]));
                } else {
                    if (acc[k][3 % C] != 0.0f)
                        atomicAdd(reinterpret_cast<float4 *>(a.image) + pix,
                                  make_float4(acc[k][0], acc[k][1 % C], acc[k][2 % C], acc[k][3 % C]));
                }
            }
        }
    }
}

#include "tsplat_surface.cuh"

// ------------------------------------------------------------------------------------------------------------
// K5: fused normalise + log/linear + colormap LUT  (colormap.wgsl:75-159)
// ------------------------------------------------------------------------------------------------------------
struct CmapArgs {
    const float *image;
    int res, channels;
    tsplat_colormap_params p;
    const float *lut;
    int lut_w, lut_h;
    void *out;
    int out_w, out_h, out_fmt;
};

__device__ __forceinline__ float wgsl_log10(float v) { return logf(v) / 2.30258509f; }

__device__ __forceinline__ float clamp01(float t) { return fminf(fmaxf(t, 0.0f), 1.0f); }   // NaN -> 0

__device__ __forceinline__ float4 lut_fetch(const float *__restrict__ lut, int idx)
{
    return __ldg(reinterpret_cast<const float4 *>(lut) + idx);
}

__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float t)
{
    return make_float4(a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t, a.w + (b.w - a.w) * t);
}

// linear-filtered, clamp-to-edge fetch at normalised coordinate t (texel centres at (i + 0.5)/n)
__device__ __forceinline__ float4 lut_sample_1d(const float *__restrict__ lut, int n, float t)
{
    const float p = t * (float)n - 0.5f;
    const float i0 = floorf(p);
    const float f = p - i0;
    const int a = min(max((int)i0, 0), n - 1), b = min(max((int)i0 + 1, 0), n - 1);
    return lerp4(lut_fetch(lut, a), lut_fetch(lut, b), f);
}

__device__ __forceinline__ float4 lut_sample_2d(const float *__restrict__ lut, int nx, int ny, float tx, float ty)
{
    const float px = tx * (float)nx - 0.5f, py = ty * (float)ny - 0.5f;
    const float ix = floorf(px), iy = floorf(py);
    const float fx = px - ix, fy = py - iy;
    const int x0 = min(max((int)ix, 0), nx - 1), x1 = min(max((int)ix + 1, 0), nx - 1);
    const int y0 = min(max((int)iy, 0), ny - 1), y1 = min(max((int)iy + 1, 0), ny - 1);
    const float4 top = lerp4(lut_fetch(lut, y0 * nx + x0), lut_fetch(lut, y0 * nx + x1), fx);
    const float4 bot = lerp4(lut_fetch(lut, y1 * nx + x0), lut_fetch(lut, y1 * nx + x1), fx);
    return lerp4(top, bot, fy);
}

__device__ __forceinline__ void load_pixel(const float *__restrict__ img, int res, int C, int col, int row, float v[4])
{
    const size_t pix = (size_t)row * res + col;
    if (C == 1) { v[0] = img[pix]; v[1] = 0.f; v[2] = 0.f; v[3] = 0.f; }
    else if (C == 2) { const float2 t = reinterpret_cast<const float2 *>(img)[pix]; v[0] = t.x; v[1] = t.y; v[2] = 0.f; v[3] = 0.f; }
    else { const float4 t = reinterpret_cast<const float4 *>(img)[pix]; v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
}

// value -> colour: the body of fragment_main / fragment_main_tri (colormap.wgsl:79-159)
__device__ __forceinline__ float4 colormap_value(const float v[4], const tsplat_colormap_params &p,
                                                 const float *__restrict__ lut, int lut_w, int lut_h)
{
    if (p.kind == TSPLAT_CMAP_RGB) {
        float c3[3] = {v[0], v[1], v[2]};
        for (int c = 0; c < 3; ++c) {
            float val = c3[c];
            if (p.log_scale) val = wgsl_log10(val);
            const float t = fmaxf((val - p.vmin) / (p.vmax - p.vmin), 0.0f);
            c3[c] = p.gamma == 1.0f ? t : powf(t, p.gamma);      // pow(t, 1) == t: skip ~100 instructions per band in the default case
        }
        return make_float4(c3[0], c3[1], c3[2], 1.0f);
    }
    const bool weighted = (p.kind == TSPLAT_CMAP_WEIGHTED || p.kind == TSPLAT_CMAP_BIVARIATE_WEIGHTED);
    float val = weighted ? v[1] / v[0] : v[0];
    if (p.log_scale) val = wgsl_log10(val);
    const float t = clamp01((val - p.vmin) / (p.vmax - p.vmin));
    if (p.kind == TSPLAT_CMAP_BIVARIATE || p.kind == TSPLAT_CMAP_BIVARIATE_WEIGHTED) {
        const float d = clamp01((wgsl_log10(v[0]) - p.density_vmin) / (p.density_vmax - p.density_vmin));
        return lut_sample_2d(lut, lut_w, lut_h, d, t);
    }
    return lut_sample_1d(lut, lut_w, t);
}

__device__ __forceinline__ void store_rgba(void *out, size_t o, int out_fmt, float4 rgba)
{
    if (out_fmt == TSPLAT_FMT_RGBA8) {
        uchar4 q;
        q.x = (unsigned char)__float2int_rn(__saturatef(rgba.x) * 255.0f);
        q.y = (unsigned char)__float2int_rn(__saturatef(rgba.y) * 255.0f);
        q.z = (unsigned char)__float2int_rn(__saturatef(rgba.z) * 255.0f);
        q.w = (unsigned char)__float2int_rn(__saturatef(rgba.w) * 255.0f);
        reinterpret_cast<uchar4 *>(out)[o] = q;
    } else if (out_fmt == TSPLAT_FMT_RGBA16F) {
        __half2 lo = __floats2half2_rn(rgba.x, rgba.y), hi = __floats2half2_rn(rgba.z, rgba.w);
        uint2 pk;
        pk.x = *reinterpret_cast<unsigned int *>(&lo);
        pk.y = *reinterpret_cast<unsigned int *>(&hi);
        reinterpret_cast<uint2 *>(out)[o] = pk;
    } else {
        reinterpret_cast<float4 *>(out)[o] = rgba;
    }
}

__global__ void __launch_bounds__(256) k_colormap(const CmapArgs a)
{
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y;
    if (ox >= a.out_w || oy >= a.out_h) return;
    float v[4];
    if (a.out_w == a.res && a.out_h == a.res) {
        load_pixel(a.image, a.res, a.channels, ox, oy, v);       // texel centres coincide with pixel centres
    } else {
        // vertex_main (colormap.wgsl:41-73): the square image covers the larger window dimension
        const float asp = a.p.window_aspect_ratio;
        const float sx = asp > 1.0f ? 1.0f : 1.0f / asp, sy = asp > 1.0f ? asp : 1.0f;
        const float X = -1.0f + (2.0f * ox + 1.0f) / (float)a.out_w;
        const float Y = 1.0f - (2.0f * oy + 1.0f) / (float)a.out_h;
        const float u = clamp01((X / sx + 1.0f) * 0.5f), w = clamp01((1.0f - Y / sy) * 0.5f);
        const float texels_per_px = (float)a.res / fmaxf((float)a.out_w, (float)a.out_h);
        if (texels_per_px <= 1.0f) {        // magnification: linear filter (implementation.py:278-279)
            const float px = u * a.res - 0.5f, py = w * a.res - 0.5f;
            const float ix = floorf(px), iy = floorf(py);
            const float fx = px - ix, fy = py - iy;
            const int x0 = min(max((int)ix, 0), a.res - 1), x1 = min(max((int)ix + 1, 0), a.res - 1);
            const int y0 = min(max((int)iy, 0), a.res - 1), y1 = min(max((int)iy + 1, 0), a.res - 1);
            float t00[4], t01[4], t10[4], t11[4];
            load_pixel(a.image, a.res, a.channels, x0, y0, t00);
            load_pixel(a.image, a.res, a.channels, x1, y0, t01);
            load_pixel(a.image, a.res, a.channels, x0, y1, t10);
            load_pixel(a.image, a.res, a.channels, x1, y1, t11);
            for (int c = 0; c < 4; ++c) {
                const float top = t00[c] + (t01[c] - t00[c]) * fx, bot = t10[c] + (t11[c] - t10[c]) * fx;
                v[c] = top + (bot - top) * fy;
            }
        } else {                            // minification: nearest (WebGPU default min_filter)
            const int x = min(max((int)floorf(u * a.res), 0), a.res - 1);
            const int y = min(max((int)floorf(w * a.res), 0), a.res - 1);
            load_pixel(a.image, a.res, a.channels, x, y, v);
        }
    }

    const float4 rgba = colormap_value(v, a.p, a.lut, a.lut_w, a.lut_h);
    store_rgba(a.out, (size_t)oy * a.out_w + ox, a.out_fmt, rgba);
}

// K11: surface lighting (surface.wgsl:24-123); the sampling helpers live in tsplat_surface.cuh
__global__ void __launch_bounds__(256) k_surface_shade(const tsplat_surface::ShadeArgs a)
{
    using tsplat_surface::sample_rg;
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y;
    if (ox >= a.out_w || oy >= a.out_h) return;
    // vertex_main: the square image covers the larger window dimension (same mapping as colormap.wgsl)
    const float asp = a.p.window_aspect_ratio;
    const float sx = asp > 1.0f ? 1.0f : 1.0f / asp, sy = asp > 1.0f ? asp : 1.0f;
    const float X = -1.0f + (2.0f * ox + 1.0f) / (float)a.out_w;
    const float Y = 1.0f - (2.0f * oy + 1.0f) / (float)a.out_h;
    const float u = (X / sx + 1.0f) * 0.5f, v = (1.0f - Y / sy) * 0.5f;
    const bool linear = (float)a.res / fmaxf((float)a.out_w, (float)a.out_h) <= 1.0f;
    const float tx = 1.0f / (float)a.out_w, ty = 1.0f / (float)a.out_h;       // uniforms.texelSize
    const float ds = a.p.depth_scale;
    const float2 centre = sample_rg(a.image, a.res, u, v, linear);
    const float d_l = sample_rg(a.image, a.res, u - tx, v, linear).y * ds, d_r = sample_rg(a.image, a.res, u + tx, v, linear).y * ds;
    const float d_u = sample_rg(a.image, a.res, u, v - ty, linear).y * ds, d_d = sample_rg(a.image, a.res, u, v + ty, linear).y * ds;
    const float dX = (d_r - d_l) * 0.5f, dY = (d_d - d_u) * 0.5f;
    const float nlen = sqrtf(dX * dX + dY * dY + tx * tx);
    const float nx = -dX / nlen, ny = -dY / nlen, nz = tx / nlen;
    const float ndotl = fmaxf(nx * a.p.light_direction[0] + ny * a.p.light_direction[1] + nz * a.p.light_direction[2], 0.0f);
    float mat[3] = {1.0f, 1.0f, 1.0f};
    if (a.p.material_colormap) {
        float val = centre.x;
        if (a.p.log_scale) val = wgsl_log10(val);
        const float4 c = lut_sample_1d(a.lut, a.lut_w, clamp01((val - a.p.vmin) / (a.p.vmax - a.p.vmin)));
        mat[0] = c.x; mat[1] = c.y; mat[2] = c.z;
    }
    const float dim = fminf(fmaxf(centre.y * ds, 0.0f), 0.5f) * 2.0f;
    float rgb[3];
    for (int c = 0; c < 3; ++c) rgb[c] = (a.p.light_color[c] * ndotl * mat[c] + a.p.ambient_color[c] * mat[c]) * dim;
    store_rgba(a.out, (size_t)oy * a.out_w + ox, a.out_fmt, make_float4(rgb[0], rgb[1], rgb[2], 1.0f));
}

// ------------------------------------------------------------------------------------------------------------
// K6: fused image sum-reduce over NVLink peers + colormap (multi-GPU).  Every rank owns a slab of rows; it loads that
// slab from every peer's accumulation image (peer pointers from PyTorch symmetric memory), adds the partial images in
// rank order (deterministic), optionally stores the fp32 sum, applies the colormap and stores RGBA -- possibly into
// another rank's buffer.  One kernel replaces reduce-scatter + colormap + gather.
// ------------------------------------------------------------------------------------------------------------
constexpr int MAX_PEERS = 16;

struct ReduceArgs {
    const float *peer[MAX_PEERS];
    int n_peers;
    int res, channels;
    int row0, nrows;
    tsplat_colormap_params p;
    const float *lut;
    int lut_w, lut_h;
    void *out;              // full R x R x 4 output image (row-major) or nullptr
    int out_fmt;
    float *sum_out;         // full R x R x C fp32 image receiving the reduced slab, or nullptr
};

// Each thread owns 4 consecutive floats of the slab (4 / 2 / 1 pixels for 1 / 2 / 4 channels): one 128-bit load per peer,
// all of them issued before the first add so that the NVLink round trips overlap (a 4-byte load per thread and peer, one
// after the other, reached 318 GB/s of a 770 GB/s link: profiles/r02/k6_peer_timing_v1.json).
template <int NP>      // peers known at compile time (2..8; 16 = any number, predicated): all NP loads are issued back to back
__global__ void __launch_bounds__(256) k_reduce_colormap(const ReduceArgs a)
{
    const int64_t first = (int64_t)a.row0 * a.res * a.channels, count = (int64_t)a.nrows * a.res * a.channels;
    const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= count) return;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (i4 + 4 <= count && ((first + i4) & 3) == 0) {
        float4 t[NP];
#pragma unroll
        for (int r = 0; r < NP; ++r)
            if (NP < MAX_PEERS || r < a.n_peers) t[r] = __ldg(reinterpret_cast<const float4 *>(a.peer[r] + first + i4));
#pragma unroll
        for (int r = 0; r < NP; ++r)
            if (NP < MAX_PEERS || r < a.n_peers) { v[0] += t[r].x; v[1] += t[r].y; v[2] += t[r].z; v[3] += t[r].w; }
    } else {
        for (int e = 0; e < 4 && i4 + e < count; ++e)
            for (int r = 0; r < a.n_peers; ++r) v[e] += a.peer[r][first + i4 + e];
    }
    const int n_valid = (int)min((int64_t)4, count - i4);
    if (a.sum_out)
        for (int e = 0; e < n_valid; ++e) a.sum_out[first + i4 + e] = v[e];
    if (a.out) {
        const int C = a.channels;
        if (a.out_fmt == TSPLAT_FMT_RGBA8 && n_valid == 4 && C < 4) {
            // 4 (or 2) pixels of this thread leave as ONE 16 (8) byte store: the output usually lives on another GPU, and
            // 4-byte stores at a 16-byte stride cross NVLink as partial sectors (c5 on 8 GPUs: 0.48 ms instead of 0.14)
            unsigned packed[4];
            for (int e = 0; e < 4; e += C) {
                float px[4] = {v[e], C > 1 ? v[e + 1] : 0.f, 0.f, 0.f};
                const float4 rgba = colormap_value(px, a.p, a.lut, a.lut_w, a.lut_h);
                packed[e / C] = (unsigned)__float2int_rn(__saturatef(rgba.x) * 255.0f) | ((unsigned)__float2int_rn(__saturatef(rgba.y) * 255.0f) << 8) |
                                ((unsigned)__float2int_rn(__saturatef(rgba.z) * 255.0f) << 16) | ((unsigned)__float2int_rn(__saturatef(rgba.w) * 255.0f) << 24);
            }
            unsigned *dst = reinterpret_cast<unsigned *>(a.out) + (first + i4) / C;
            if (C == 1) *reinterpret_cast<uint4 *>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            else *reinterpret_cast<uint2 *>(dst) = make_uint2(packed[0], packed[1]);
        } else {
            for (int e = 0; e + C <= n_valid; e += C) {
                float px[4] = {v[e], C > 1 ? v[e + 1] : 0.f, C > 2 ? v[e + 2] : 0.f, C > 2 ? v[e + 3] : 0.f};
                store_rgba(a.out, (size_t)((first + i4 + e) / C), a.out_fmt, colormap_value(px, a.p, a.lut, a.lut_w, a.lut_h));
            }
        }
    }
}

// K6b: image all-reduce over NVLink peer memory for the drop-in classes: every rank reduces its slab of rows over all
// partial images and stores the result into every peer's reduced image (reduce-scatter + all-gather in one kernel).
struct AllReduceArgs {
    const float *peer[MAX_PEERS];
    float *out[MAX_PEERS];
    const float *scale[MAX_PEERS];
    int n_peers, op;
    int64_t first, count;      // slab as a range of floats of the flattened (R, R, C) image
};

template <int NP>
__global__ void __launch_bounds__(256) k_allreduce_image(const AllReduceArgs a)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (a.op == TSPLAT_REDUCE_ZMAX) {           // (quantity, depth) pixels as 64-bit keys, depth in the high word
        const int64_t n2 = a.count >> 1;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
            unsigned long long best = 0ull;
            for (int r = 0; r < a.n_peers; ++r)
                best = max(best, reinterpret_cast<const unsigned long long *>(a.peer[r] + a.first)[i]);
            for (int r = 0; r < a.n_peers; ++r) reinterpret_cast<unsigned long long *>(a.out[r] + a.first)[i] = best;
        }
        return;
    }
    float w[MAX_PEERS];
#pragma unroll
    for (int r = 0; r < MAX_PEERS; ++r) w[r] = r < a.n_peers ? *a.scale[r] : 0.0f;
    if (((a.first | a.count) & 3) == 0) {       // 128-bit path (rows of R * C floats: a multiple of 4 for every usual R)
        const int64_t n4 = a.count >> 2;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 t[NP];
#pragma unroll
            for (int r = 0; r < NP; ++r)
                if (NP < MAX_PEERS || r < a.n_peers) t[r] = __ldg(reinterpret_cast<const float4 *>(a.peer[r] + a.first) + i);
#pragma unroll
            for (int r = 0; r < NP; ++r)
                if (NP < MAX_PEERS || r < a.n_peers) { v.x += w[r] * t[r].x; v.y += w[r] * t[r].y; v.z += w[r] * t[r].z; v.w += w[r] * t[r].w; }
#pragma unroll
            for (int r = 0; r < NP; ++r)
                if (NP < MAX_PEERS || r < a.n_peers) reinterpret_cast<float4 *>(a.out[r] + a.first)[i] = v;
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.count; i += stride) {
            float v = 0.0f;
            for (int r = 0; r < a.n_peers; ++r) v += w[r] * a.peer[r][a.first + i];
            for (int r = 0; r < a.n_peers; ++r) a.out[r][a.first + i] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// K7: periodic tiling -- out = sum_i w_i * image shifted by offset_i (reference: periodic_sph.py:16-88 draws the SPH
// texture as <= 125 instanced quads through overlay.wgsl with a linear sampler and ONE/ONE blending)
// ------------------------------------------------------------------------------------------------------------
constexpr int MAX_REPLICAS = 128;             // Overlay.MAX_INSTANCES (overlay.py:21)

struct PeriodicArgs {
    const float *src;
    float *dst;
    int res, channels, n;
    float ox[MAX_REPLICAS], oy[MAX_REPLICAS], w[MAX_REPLICAS];
};

__global__ void __launch_bounds__(256) k_periodic_accumulate(const PeriodicArgs a)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (j >= a.res) return;
    const float Rf = (float)a.res;
    const float X = (2.0f * (float)j + 1.0f) / Rf - 1.0f;      // pixel centre in clip space, row 0 = +y
    const float Y = 1.0f - (2.0f * (float)k + 1.0f) / Rf;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < a.n; ++i) {
        const float u = (X - (a.ox[i] - 1.0f)) * 0.5f;
        const float v = ((a.oy[i] + 1.0f) - Y) * 0.5f;
        if (!(u >= 0.0f && u < 1.0f && v > 0.0f && v <= 1.0f)) continue;       // outside this replica's quad
        const float px = u * Rf - 0.5f, py = v * Rf - 0.5f;
        const float ix = floorf(px), iy = floorf(py);
        const float fx = px - ix, fy = py - iy;
        const int x0 = min(max((int)ix, 0), a.res - 1), x1 = min(max((int)ix + 1, 0), a.res - 1);
        const int y0 = min(max((int)iy, 0), a.res - 1), y1 = min(max((int)iy + 1, 0), a.res - 1);
        float t00[4], t01[4], t10[4], t11[4];
        load_pixel(a.src, a.res, a.channels, x0, y0, t00);
        load_pixel(a.src, a.res, a.channels, x1, y0, t01);
        load_pixel(a.src, a.res, a.channels, x0, y1, t10);
        load_pixel(a.src, a.res, a.channels, x1, y1, t11);
        const float wi = a.w[i];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float top = t00[c] + (t01[c] - t00[c]) * fx, bot = t10[c] + (t11[c] - t10[c]) * fx;
            acc[c] += wi * (top + (bot - top) * fy);
        }
    }
    const size_t pix = (size_t)k * a.res + j;
    if (a.channels == 1) a.dst[pix] = acc[0];
    else if (a.channels == 2) reinterpret_cast<float2 *>(a.dst)[pix] = make_float2(acc[0], acc[1]);
    else reinterpret_cast<float4 *>(a.dst)[pix] = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// ------------------------------------------------------------------------------------------------------------
// K8: device-side autorange.  Replaces the host np.percentile of
// Colormap._autorange_using_values / RGBColormap.autorange_vmin_vmax / BivariateColormap.autorange_vmin_vmax
// (colormap/implementation.py:381-425, :512-531, :576-588): min/max/sign statistics in one pass, then exact order
// statistics by a 3-pass (11 + 11 + 10 bit) radix select over the monotone integer image of the float values.
// ------------------------------------------------------------------------------------------------------------
enum { CONTENT_CH0 = 0, CONTENT_RATIO = 1, CONTENT_ALL = 2, CONTENT_CH0_WHERE_CH1 = 3 };

struct ContentArgs {
    const float *image;
    int64_t n_values;            // res*res (CH0, RATIO) or res*res*channels (ALL)
    int channels, content;
    float scale;
    int use_log;
};

__device__ __forceinline__ float content_value(const ContentArgs &a, int64_t i)
{
    if (a.content == CONTENT_ALL) return a.image[i] * a.scale;
    const float c0 = a.image[i * a.channels] * a.scale;
    if (a.content == CONTENT_CH0) return c0;
    if (a.content == CONTENT_CH0_WHERE_CH1)        // surface maps: the material value of the pixels that received a fragment
        return a.image[i * a.channels + 1] > 0.0f ? c0 : __int_as_float(0x7fc00000);
    return (a.image[i * a.channels + 1] * a.scale) / c0;
}

// order-preserving map float -> uint32 (finite values only)
__device__ __forceinline__ unsigned float_key(float v)
{
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__host__ __device__ inline float key_to_float(unsigned k)
{
    const unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}

struct ContentStats {           // device + host layout
    unsigned lin_min_key, lin_max_key, log_min_key, log_max_key;
    unsigned any_negative, pad;
    unsigned long long n_finite_lin, n_finite_log;
};

__global__ void __launch_bounds__(256) k_content_stats(const ContentArgs a, ContentStats *out)
{
    unsigned lmin = 0xffffffffu, lmax = 0u, gmin = 0xffffffffu, gmax = 0u, neg = 0u;
    unsigned long long nl = 0, ng = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_values; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = content_value(a, i);
        if (v < 0.0f) neg = 1u;
        if (isfinite(v)) { const unsigned k = float_key(v); lmin = min(lmin, k); lmax = max(lmax, k); ++nl; }
        const float lv = log10f(v);
        if (isfinite(lv)) { const unsigned k = float_key(lv); gmin = min(gmin, k); gmax = max(gmax, k); ++ng; }
    }
    for (int d = 16; d > 0; d >>= 1) {
        lmin = min(lmin, __shfl_down_sync(0xffffffffu, lmin, d)); lmax = max(lmax, __shfl_down_sync(0xffffffffu, lmax, d));
        gmin = min(gmin, __shfl_down_sync(0xffffffffu, gmin, d)); gmax = max(gmax, __shfl_down_sync(0xffffffffu, gmax, d));
        neg |= __shfl_down_sync(0xffffffffu, neg, d);
        nl += __shfl_down_sync(0xffffffffu, nl, d); ng += __shfl_down_sync(0xffffffffu, ng, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&out->lin_min_key, lmin); atomicMax(&out->lin_max_key, lmax);
        atomicMin(&out->log_min_key, gmin); atomicMax(&out->log_max_key, gmax);
        if (neg) atomicOr(&out->any_negative, 1u);
        if (nl) atomicAdd(&out->n_finite_lin, nl);
        if (ng) atomicAdd(&out->n_finite_log, ng);
    }
}

constexpr int SELECT_MAX_RANKS = 4;
struct SelectArgs {
    ContentArgs c;
    int n_ranks, shift, bits;              // digit = (key >> shift) & ((1 << bits) - 1)
    unsigned prefix[SELECT_MAX_RANKS];     // already-resolved high bits of each rank's key
    unsigned prefix_mask;                  // mask of the resolved bits
    unsigned *hist;                        // [n_ranks][2048]
};

__global__ void __launch_bounds__(256) k_content_select(const SelectArgs a)
{
    __shared__ unsigned s_hist[SELECT_MAX_RANKS][2048];
    for (int i = threadIdx.x; i < a.n_ranks * 2048; i += blockDim.x) (&s_hist[0][0])[i] = 0u;
    __syncthreads();
    const unsigned dmask = (1u << a.bits) - 1u;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.c.n_values; i += (int64_t)gridDim.x * blockDim.x) {
        float v = content_value(a.c, i);
        if (a.c.use_log) v = log10f(v);
        if (!isfinite(v)) continue;
        const unsigned k = float_key(v);
        for (int r = 0; r < a.n_ranks; ++r)
            if ((k & a.prefix_mask) == a.prefix[r]) atomicAdd(&s_hist[r][(k >> a.shift) & dmask], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.n_ranks * 2048; i += blockDim.x) {
        const unsigned v = (&s_hist[0][0])[i];
        if (v) atomicAdd(&a.hist[i], v);
    }
}

__global__ void k_axpy(float *__restrict__ dst, const float *__restrict__ src, float scale, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = dst[i] + src[i] * scale;
}

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
extern "C" int tsplat_create(int device_ordinal, int resolution, tsplat_ctx **out)
{
    if (!out) return set_err(TSPLAT_ERR_INVALID, "out is NULL");
    if (resolution <= 0 || resolution > 32768) return set_err(TSPLAT_ERR_INVALID, "bad resolution %d", resolution);
    CUDA_TRY(cudaSetDevice(device_ordinal));
    tsplat_ctx *c = new (std::nothrow) tsplat_ctx();
    if (!c) return set_err(TSPLAT_ERR_NOMEM, "out of host memory");
    memset(c, 0, sizeof(*c));
    c->device = device_ordinal;
    c->R = resolution;
    c->cam.R = (float)resolution;
    c->cam.halfR = 0.5f * (float)resolution;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device_ordinal));
    c->sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaMalloc(&c->d_lut, LUT_TOTAL * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c->d_slut, LUT_TOTAL * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c->d_counters, sizeof(Counters)));
    CUDA_TRY(cudaMemset(c->d_counters, 0, sizeof(Counters)));
    CUDA_TRY(cudaMalloc(&c->d_select, sizeof(unsigned) * 4 * 2048));
    CUDA_TRY(cudaMallocHost(&c->h_select, sizeof(unsigned) * 4 * 2048));
    CUDA_TRY(cudaMallocHost(&c->h_qhint, 6 * sizeof(unsigned)));
    memset(c->h_qhint, 0, 6 * sizeof(unsigned));
    for (int s = 0; s < RANGE_SLOTS; ++s) {
        CUDA_TRY(cudaMallocHost(&c->h_ranges[s], sizeof(int64_t) * (3 * (size_t)MAX_RANGES + 1)));
        CUDA_TRY(cudaMalloc(&c->d_ranges[s], sizeof(int64_t) * (3 * (size_t)MAX_RANGES + 1)));
        CUDA_TRY(cudaEventCreateWithFlags(&c->range_evt[s], cudaEventDisableTiming));
    }
    *out = c;
    return TSPLAT_OK;
}

extern "C" int tsplat_destroy(tsplat_ctx *c)
{
    if (!c) return TSPLAT_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    cudaFree(c->d_lut);
    cudaFree(c->d_slut);
    cudaFree(c->d_counters);
    cudaFree(c->d_select);
    cudaFreeHost(c->h_select);
    cudaFreeHost(c->h_qhint);
    if (c->timing_events_created)
        for (int i = 0; i < TSPLAT_TIMING_SLOTS; ++i) { cudaEventDestroy(c->t_begin[i]); cudaEventDestroy(c->t_end[i]); }
    for (int s = 0; s < RANGE_SLOTS; ++s) {
        cudaFreeHost(c->h_ranges[s]);
        cudaFree(c->d_ranges[s]);
        cudaEventDestroy(c->range_evt[s]);
    }
    delete c;
    return TSPLAT_OK;
}

extern "C" int tsplat_set_kernel_lut(tsplat_ctx *c, const float *host_lut, int n_floats)
{
    if (!c || !host_lut) return set_err(TSPLAT_ERR_INVALID, "NULL argument");
    if (n_floats != LUT_TOTAL) return set_err(TSPLAT_ERR_INVALID, "kernel LUT must have %d floats, got %d", LUT_TOTAL, n_floats);
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpy(c->d_lut, host_lut, LUT_TOTAL * sizeof(float), cudaMemcpyHostToDevice));
    c->lut_set = true;
    return TSPLAT_OK;
}

extern "C" int tsplat_set_camera(tsplat_ctx *c, const float *M16, float scale_factor)
{
    if (!c || !M16) return set_err(TSPLAT_ERR_INVALID, "NULL argument");
    for (int i = 0; i < 12; ++i) c->cam.m[i] = M16[i];
    c->cam.sf = scale_factor;
    c->camera_set = true;
    return TSPLAT_OK;
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int tsplat_set_particles(tsplat_ctx *c, const float *x, const float *y, const float *z, const float *h, int64_t n)
{
    if (!c) return set_err(TSPLAT_ERR_INVALID, "NULL context");
    if (n < 0 || n >= ((int64_t)1 << 33)) return set_err(TSPLAT_ERR_INVALID, "bad particle count %lld", (long long)n);
    if (n > 0 && (!x || !y || !z || !h)) return set_err(TSPLAT_ERR_INVALID, "NULL particle array");
    if (!aligned16(x) || !aligned16(y) || !aligned16(z) || !aligned16(h))
        return set_err(TSPLAT_ERR_INVALID, "particle arrays must be 16-byte aligned (128-bit loads)");
    c->x = x; c->y = y; c->z = z; c->h = h; c->n = n;
    c->w0 = c->w1 = c->w2 = nullptr;
    return TSPLAT_OK;
}

extern "C" int tsplat_set_weights(tsplat_ctx *c, const float *w0, const float *w1, const float *w2)
{
    if (!c) return set_err(TSPLAT_ERR_INVALID, "NULL context");
    if (!aligned16(w0) || !aligned16(w1) || !aligned16(w2))
        return set_err(TSPLAT_ERR_INVALID, "weight arrays must be 16-byte aligned (128-bit loads)");
    c->w0 = w0; c->w1 = w1; c->w2 = w2;
    return TSPLAT_OK;
}

extern "C" int tsplat_set_image(tsplat_ctx *c, float *image, int channels)
{
    if (!c || !image) return set_err(TSPLAT_ERR_INVALID, "NULL argument");
    if (channels != 1 && channels != 2 && channels != 4) return set_err(TSPLAT_ERR_INVALID, "channels must be 1, 2 or 4");
    if (!aligned16(image)) return set_err(TSPLAT_ERR_INVALID, "image must be 16-byte aligned");
    c->image = image; c->channels = channels;
    return TSPLAT_OK;
}

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct ScratchLayout {
    int64_t queue_cap, pairs_cap;
    int ntx, nt;
    int64_t seg_cap;
    int64_t queue_off, route_off, huge_off, tcount_off, toffset_off, tcursor_off, segpref_off, segtile_off, pairs_off, total;
};

constexpr int64_t PAIRS_PER_PARTICLE = 6;     // pair capacity relative to the queue capacity (overflow -> atomic path)

static ScratchLayout scratch_layout(int R, int64_t cap)
{
    ScratchLayout L;
    if (cap < 0) cap = 0;
    if (cap > 0x7fffff00ll) cap = 0x7fffff00ll;
    L.queue_cap = cap;
    L.pairs_cap = cap * PAIRS_PER_PARTICLE;
    if (L.pairs_cap > 0xfffffff0ll) L.pairs_cap = 0xfffffff0ll;
    L.ntx = (R + TILE_W - 1) / TILE_W;
    L.nt = L.ntx * ((R + TILE_H - 1) / TILE_H);
    int64_t o = 0;
    L.queue_off = o;   o += align_up(cap * (int64_t)sizeof(Deferred), 256);
    L.route_off = o;   o += align_up(cap, 256);
    L.huge_off = o;    o += align_up(cap * 4, 256);
    L.tcount_off = o;  o += align_up(((int64_t)L.nt + 1) * 4, 256);
    L.tcursor_off = o; o += align_up(((int64_t)L.nt + 1) * 4, 256);      // contiguous with tcount: one memset clears both
    L.toffset_off = o; o += align_up(((int64_t)L.nt + 1) * 4, 256);
    L.segpref_off = o; o += align_up(((int64_t)L.nt + 1) * 4, 256);
    L.seg_cap = (int64_t)L.nt + L.pairs_cap / SEG + 1;
    L.segtile_off = o; o += align_up(L.seg_cap * 4, 256);
    L.pairs_off = o;   o += align_up(L.pairs_cap * 4, 256);
    L.total = o;
    return L;
}

// largest queue capacity whose layout fits in `bytes`
static ScratchLayout scratch_layout_from_bytes(int R, int64_t bytes)
{
    int64_t lo = 0, hi = bytes / (int64_t)sizeof(Deferred) + 1;      // layout(lo) fits (or nothing does), layout(hi) does not
    if (scratch_layout(R, 0).total > bytes) return scratch_layout(R, 0);
    while (hi - lo > 1) {
        const int64_t mid = lo + (hi - lo) / 2;
        if (scratch_layout(R, mid).total <= bytes) lo = mid; else hi = mid;
    }
    return scratch_layout(R, lo);
}

extern "C" int64_t tsplat_scratch_bytes(int resolution, int64_t max_particles_per_call)
{
    if (max_particles_per_call < 0) return -1;
    return scratch_layout(resolution, max_particles_per_call).total;
}

extern "C" int tsplat_set_scratch(tsplat_ctx *c, void *scratch, int64_t bytes)
{
    if (!c) return set_err(TSPLAT_ERR_INVALID, "NULL context");
    if (bytes < 0 || (bytes > 0 && !scratch)) return set_err(TSPLAT_ERR_INVALID, "bad scratch");
    if (!aligned16(scratch)) return set_err(TSPLAT_ERR_INVALID, "scratch must be 16-byte aligned");
    c->scratch = scratch; c->scratch_bytes = bytes;
    return TSPLAT_OK;
}

template <int MODE>
static int launch_render(tsplat_ctx *c, const ProjectArgs &pa, int64_t n_groups, const ScratchLayout &L, cudaStream_t st)
{
    const int64_t blocks = (n_groups + KP_THREADS - 1) / KP_THREADS;
    if (n_groups > 0x7fffffffll) return set_err(TSPLAT_ERR_INVALID, "too many particles in one call");
    char *sc = static_cast<char *>(c->scratch);
    // per-call state: q_count .. pad, tile counters and cursors
    CUDA_TRY(cudaMemsetAsync(&c->d_counters->q_count, 0, (6 + PAIR_STRIPES) * sizeof(unsigned int), st));
    if (!pa.small_call) CUDA_TRY(cudaMemsetAsync(sc + L.tcount_off, 0, (size_t)(L.tcursor_off - L.tcount_off) * 2, st));
    if (blocks > 0) {
        // cell width of the vector REDs: as many pixels as fit 128 bits, if rows keep the cells aligned
        constexpr int CW = ModeTraits<MODE>::C == 1 ? 4 : ModeTraits<MODE>::C == 2 ? 2 : 1;
        const bool timed = c->timing && c->timing_write - c->timing_read < TSPLAT_TIMING_SLOTS;
        if (timed) CUDA_TRY(cudaEventRecord(c->t_begin[c->timing_write % TSPLAT_TIMING_SLOTS], st));
        {
            // persistent warps: one batch of 128 particles per warp and iteration, as many CTAs as the SMs hold at once
            const int64_t n_batches = (n_groups + 31) >> 5;
            int64_t grid = (n_batches + KP_WARPS - 1) / KP_WARPS;
            if (CW > 1 && (c->R % CW) == 0) {
                const int64_t max_grid = (int64_t)c->sm_count * kp_ctas_per_sm<MODE, CW>();
                k_project_stream<MODE, CW><<<(unsigned)(grid < max_grid ? grid : max_grid), KP_THREADS, 0, st>>>(pa);
            } else {
                const int64_t max_grid = (int64_t)c->sm_count * kp_ctas_per_sm<MODE, 1>();
                k_project_stream<MODE, 1><<<(unsigned)(grid < max_grid ? grid : max_grid), KP_THREADS, 0, st>>>(pa);
            }
        }
        if (timed) { CUDA_TRY(cudaEventRecord(c->t_end[c->timing_write % TSPLAT_TIMING_SLOTS], st)); c->timing_write++; }
        c->launches++;
    }
    if (pa.queue_cap > 0 && pa.small_call) {
        // an interactive block of a few thousand particles: the binning machinery (4 more launches) costs more than it
        // saves, k_bin_count would route every record to the cooperative atomic kernel anyway (Q < SMALL_QUEUE)
        QueueArgs qa;
        qa.queue = pa.queue; qa.indices = nullptr; qa.count = &c->d_counters->q_count; qa.cap = pa.queue_cap;
        qa.lut = c->d_lut; qa.image = c->image; qa.R = c->R;
        k_queue_atomic<MODE><<<c->sm_count * 8, 256, 0, st>>>(qa);
        c->launches++;
    } else if (pa.queue_cap > 0) {
        BinArgs ba;
        ba.queue = pa.queue; ba.queue_cap = pa.queue_cap; ba.counters = c->d_counters;
        ba.route = reinterpret_cast<unsigned char *>(sc + L.route_off);
        ba.huge_idx = reinterpret_cast<unsigned int *>(sc + L.huge_off);
        ba.tile_count = reinterpret_cast<unsigned int *>(sc + L.tcount_off);
        ba.tile_cursor = reinterpret_cast<unsigned int *>(sc + L.tcursor_off);
        ba.tile_offset = reinterpret_cast<unsigned int *>(sc + L.toffset_off);
        ba.seg_prefix = reinterpret_cast<unsigned int *>(sc + L.segpref_off);
        ba.pairs = reinterpret_cast<unsigned int *>(sc + L.pairs_off);
        ba.pairs_cap = (unsigned)L.pairs_cap;
        ba.R = c->R; ba.ntx = L.ntx; ba.nt = L.nt;
        ba.seg_tile = reinterpret_cast<unsigned int *>(sc + L.segtile_off);
        ba.seg_cap = (unsigned)L.seg_cap;
        const size_t hist_bytes = (size_t)L.nt * sizeof(unsigned);
        const bool use_smem = 2 * hist_bytes <= 200 * 1024;
        const int bin_grid = c->sm_count * 2;      // 32 registers x 1024 threads: two CTAs per SM hide the load latency
        if (use_smem) {
            if (!c->bin_attr_set) {
                CUDA_TRY(cudaFuncSetAttribute(k_bin_count<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                CUDA_TRY(cudaFuncSetAttribute(k_bin_fill<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                c->bin_attr_set = true;
            }
            k_bin_count<true><<<bin_grid, 1024, hist_bytes, st>>>(ba);
            k_bin_scan<<<1, 1024, 0, st>>>(ba);
            k_bin_fill<true><<<c->sm_count, 1024, 2 * hist_bytes, st>>>(ba);
        } else {
            k_bin_count<false><<<bin_grid, 1024, 0, st>>>(ba);
            k_bin_scan<<<1, 1024, 0, st>>>(ba);
            k_bin_fill<false><<<c->sm_count, 1024, 0, st>>>(ba);
        }
        GatherArgs ga;
        ga.queue = pa.queue; ga.pairs = ba.pairs; ga.tile_offset = ba.tile_offset; ga.seg_prefix = ba.seg_prefix;
        ga.seg_tile = ba.seg_tile;
        ga.counters = c->d_counters; ga.lut = c->d_lut; ga.image = c->image; ga.R = c->R; ga.ntx = L.ntx; ga.nt = L.nt;
        if (!c->gather_attr_set[MODE]) {
            CUDA_TRY(cudaFuncSetAttribute(k_tile_gather<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GatherSmem)));
            c->gather_attr_set[MODE] = true;
        }
        k_tile_gather<MODE><<<c->sm_count * GATHER_CTAS_PER_SM, 256, sizeof(GatherSmem), st>>>(ga);
        QueueArgs qa;
        qa.queue = pa.queue; qa.indices = ba.huge_idx; qa.count = &c->d_counters->huge_count; qa.cap = pa.queue_cap;
        qa.lut = c->d_lut; qa.image = c->image; qa.R = c->R;
        k_queue_atomic<MODE><<<c->sm_count * 8, 256, 0, st>>>(qa);
        c->launches += 5;
    }
    if (pa.queue_cap > 0)       // q_count .. call_particles of this call -> pinned memory, for the next calls' routing hint
        CUDA_TRY(cudaMemcpyAsync(c->h_qhint, &c->d_counters->q_count, 6 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaGetLastError());
    return TSPLAT_OK;
}

// surface mode: project + inline z-buffer fragments, then the deferred (large) footprints; no binning needed
static int launch_render_surface(tsplat_ctx *c, const ProjectArgs &pa, int64_t n_groups, cudaStream_t st)
{
    const int64_t blocks = (n_groups + 255) / 256;
    if (blocks > 0x7fffffffll) return set_err(TSPLAT_ERR_INVALID, "too many particles in one call");
    CUDA_TRY(cudaMemsetAsync(&c->d_counters->q_count, 0, (6 + PAIR_STRIPES) * sizeof(unsigned int), st));
    if (blocks > 0) {
        ProjectArgs a2 = pa;
        a2.lut = c->d_slut;
        tsplat_surface::k_project_surface<<<(unsigned)blocks, 256, 0, st>>>(a2, c->density_cut);
        QueueArgs qa;
        qa.queue = pa.queue; qa.indices = nullptr; qa.count = &c->d_counters->q_count; qa.cap = pa.queue_cap;
        qa.lut = c->d_slut; qa.image = c->image; qa.R = c->R;
        tsplat_surface::k_queue_surface<<<c->sm_count * 8, 256, 0, st>>>(qa);
        c->launches += 2;
    }
    CUDA_TRY(cudaGetLastError());
    return TSPLAT_OK;
}

extern "C" int tsplat_render(tsplat_ctx *c, const int64_t *starts, const int64_t *lens, int n_ranges, int mode,
                             int clear, void *stream)
{
    if (!c) return set_err(TSPLAT_ERR_INVALID, "NULL context");
    const int C = tsplat_mode_channels(mode);
    if (C < 0) return set_err(TSPLAT_ERR_INVALID, "unknown mode %d", mode);
    if (mode == TSPLAT_MODE_SURFACE ? !c->surface_set : !c->lut_set)
        return set_err(TSPLAT_ERR_STATE, mode == TSPLAT_MODE_SURFACE ? "surface LUT / density cut not set" : "kernel LUT not set");
    if (!c->camera_set) return set_err(TSPLAT_ERR_STATE, "camera not set");
    if (!c->image) return set_err(TSPLAT_ERR_STATE, "image not set");
    if (c->channels != C) return set_err(TSPLAT_ERR_INVALID, "mode %d needs a %d-channel image, have %d", mode, C, c->channels);
    if (c->n > 0 && !c->x) return set_err(TSPLAT_ERR_STATE, "particles not set");
    if (c->n > 0 && !c->w0) return set_err(TSPLAT_ERR_STATE, "weights not set");
    if ((mode == TSPLAT_MODE_WEIGHTED || mode == TSPLAT_MODE_RGB || mode == TSPLAT_MODE_SURFACE) && c->n > 0 && !c->w1)
        return set_err(TSPLAT_ERR_STATE, "second weight array not set");
    if (mode == TSPLAT_MODE_RGB && c->n > 0 && !c->w2) return set_err(TSPLAT_ERR_STATE, "third weight array not set");
    if (n_ranges < 0) return set_err(TSPLAT_ERR_INVALID, "n_ranges %d is negative", n_ranges);
    if (n_ranges > 0 && (!starts || !lens)) return set_err(TSPLAT_ERR_INVALID, "NULL range arrays");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(c->device));
    c->last_stream = st;

    if (clear) {
        CUDA_TRY(cudaMemsetAsync(c->image, 0, sizeof(float) * (size_t)c->R * c->R * C, st));
        CUDA_TRY(cudaMemsetAsync(c->d_counters, 0, sizeof(Counters), st));
    }
    const ScratchLayout L = scratch_layout_from_bytes(c->R, c->scratch_bytes);

    // validate + count
    int64_t one_start = 0, one_len = c->n;
    if (n_ranges == 0) { starts = &one_start; lens = &one_len; n_ranges = 1; }
    int64_t total = 0, total_groups = 0;
    for (int r = 0; r < n_ranges; ++r) {
        if (starts[r] < 0 || lens[r] < 0 || starts[r] + lens[r] > c->n)
            return set_err(TSPLAT_ERR_INVALID, "range %d = [%lld, +%lld) outside the %lld particles of the buffer", r,
                           (long long)starts[r], (long long)lens[r], (long long)c->n);
        total += lens[r];
        if (lens[r] > 0) total_groups += ((starts[r] + lens[r] + 3) >> 2) - (starts[r] >> 2);
    }
    if (total == 0) return TSPLAT_OK;
    if (L.queue_cap < total && L.queue_cap < (int64_t)1 << 20)
        return set_err(TSPLAT_ERR_STATE, "scratch too small: need tsplat_scratch_bytes(R, >= min(particles per call, 2^20))");

    ProjectArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.x = c->x; pa.y = c->y; pa.z = c->z; pa.h = c->h; pa.w0 = c->w0; pa.w1 = c->w1; pa.w2 = c->w2;
    pa.cam = c->cam; pa.R = c->R; pa.image = c->image; pa.lut = c->d_lut;
    pa.queue = reinterpret_cast<Deferred *>(static_cast<char *>(c->scratch) + L.queue_off);
    pa.counters = c->d_counters;
    pa.n_total = c->n;

    // Each launch may defer at most queue_cap particles: split the call into chunks of <= queue_cap particles.
    // (With scratch sized by tsplat_scratch_bytes(R, max block) this is a single chunk.)
    const int64_t chunk_cap = L.queue_cap;
    int rc = TSPLAT_OK;
    auto submit = [&](const ProjectArgs &args, int64_t n_groups, int64_t n_particles) -> int {
        ProjectArgs a2 = args;
        a2.queue_cap = (unsigned)(chunk_cap < n_particles ? chunk_cap : n_particles);
        a2.small_call = a2.queue_cap < SMALL_QUEUE;
        a2.call_particles = (unsigned)n_particles;
        if (!a2.small_call && mode != TSPLAT_MODE_SURFACE) {
            // hint only (K3b handles any queue length): deferred fraction of the most recent call whose read-back has
            // landed x this call's particles.  The two words are read without synchronisation: a torn pair can only
            // mis-steer one call between two correct paths.
            const unsigned q_prev = reinterpret_cast<volatile unsigned *>(c->h_qhint)[0];
            const unsigned n_prev = reinterpret_cast<volatile unsigned *>(c->h_qhint)[5];
            if (n_prev > 0) a2.small_call = (double)q_prev / (double)n_prev * (double)n_particles < SMALL_QUEUE / 2;
        }
        a2.n_groups = n_groups;
        switch (mode) {
        case TSPLAT_MODE_SURFACE: return launch_render_surface(c, a2, n_groups, st);
        case TSPLAT_MODE_DENSITY: return launch_render<TSPLAT_MODE_DENSITY>(c, a2, n_groups, L, st);
        case TSPLAT_MODE_WEIGHTED: return launch_render<TSPLAT_MODE_WEIGHTED>(c, a2, n_groups, L, st);
        case TSPLAT_MODE_RGB: return launch_render<TSPLAT_MODE_RGB>(c, a2, n_groups, L, st);
        default: return launch_render<TSPLAT_MODE_DEPTH>(c, a2, n_groups, L, st);
        }
    };

    if (n_ranges == 1 || total <= 0) {
        // single range: split by chunk_cap particles
        int64_t s = starts[0];
        const int64_t e = starts[0] + lens[0];
        while (s < e) {
            const int64_t ce = (e - s > chunk_cap) ? s + chunk_cap : e;
            pa.start = s; pa.end = ce; pa.g0 = s >> 2;
            const int64_t ng = ((ce + 3) >> 2) - (s >> 2);
            pa.table.n = 0;
            rc = submit(pa, ng, ce - s);
            if (rc) return rc;
            s = ce;
        }
    } else {
        // multi-range: stage (start, end, gprefix) through a pinned ring slot; chunk by cumulative particle count.
        // A launch takes at most chunk_cap particles and MAX_RANGES ranges; a range longer than what is left of the
        // chunk is split (cursor = range r0, offset off0 into it), so neither a huge cell nor more than MAX_RANGES cells
        // per call is an error.
        int r0 = 0;
        int64_t off0 = 0;
        while (r0 < n_ranges) {
            const int slot = c->range_slot;
            c->range_slot = (c->range_slot + 1) % RANGE_SLOTS;
            CUDA_TRY(cudaEventSynchronize(c->range_evt[slot]));     // previous use of this slot has been copied
            int64_t *hs = c->h_ranges[slot];
            int64_t acc = 0, groups = 0;
            int m = 0;
            // layout: start[cap_m] | end[cap_m] | gprefix[cap_m + 1]
            const int cap_m = (int)std::min<int64_t>((int64_t)n_ranges - r0, (int64_t)MAX_RANGES);
            int64_t *h_start = hs, *h_end = hs + cap_m, *h_pref = hs + 2 * (int64_t)cap_m;
            while (r0 < n_ranges && m < cap_m && acc < chunk_cap) {
                const int64_t s0 = starts[r0] + off0;
                const int64_t left = lens[r0] - off0;
                const int64_t take = std::min(left, chunk_cap - acc);
                if (take > 0) {
                    h_start[m] = s0; h_end[m] = s0 + take; h_pref[m] = groups;
                    groups += ((s0 + take + 3) >> 2) - (s0 >> 2);
                    ++m;
                    acc += take;
                }
                if (take == left) { ++r0; off0 = 0; } else off0 += take;
            }
            h_pref[m] = groups;
            if (m > 0) {
                int64_t *ds = c->d_ranges[slot];
                CUDA_TRY(cudaMemcpyAsync(ds, hs, sizeof(int64_t) * (3 * (size_t)cap_m + 1), cudaMemcpyHostToDevice, st));
                CUDA_TRY(cudaEventRecord(c->range_evt[slot], st));
                pa.table.start = ds; pa.table.end = ds + cap_m; pa.table.gprefix = ds + 2 * (int64_t)cap_m; pa.table.n = m;
                rc = submit(pa, groups, acc);
                if (rc) return rc;
            }
        }
    }
    return TSPLAT_OK;
}

extern "C" int tsplat_colormap(tsplat_ctx *c, const float *image, int image_res, int channels,
                               const tsplat_colormap_params *params, const float *lut, int lut_w, int lut_h,
                               void *out, int out_w, int out_h, int out_fmt, void *stream)
{
    if (!c || !image || !params || !out) return set_err(TSPLAT_ERR_INVALID, "NULL argument");
    if (channels != 1 && channels != 2 && channels != 4) return set_err(TSPLAT_ERR_INVALID, "channels must be 1, 2 or 4");
    if (image_res <= 0 || out_w <= 0 || out_h <= 0) return set_err(TSPLAT_ERR_INVALID, "bad size");
    if (out_fmt < TSPLAT_FMT_RGBA8 || out_fmt > TSPLAT_FMT_RGBA32F) return set_err(TSPLAT_ERR_INVALID, "bad output format");
    const int kind = params->kind;
    if (kind < TSPLAT_CMAP_DENSITY || kind > TSPLAT_CMAP_RGB) return set_err(TSPLAT_ERR_INVALID, "bad colormap kind");
    if (kind == TSPLAT_CMAP_RGB && channels < 4) return set_err(TSPLAT_ERR_INVALID, "RGB map needs a 4-channel image");
    if ((kind == TSPLAT_CMAP_WEIGHTED || kind == TSPLAT_CMAP_BIVARIATE_WEIGHTED) && channels < 2)
        return set_err(TSPLAT_ERR_INVALID, "weighted map needs a 2-channel image");
    if (kind != TSPLAT_CMAP_RGB) {
        if (!lut || lut_w <= 0 || lut_h <= 0) return set_err(TSPLAT_ERR_INVALID, "colormap LUT missing");
        if (!aligned16(lut)) return set_err(TSPLAT_ERR_INVALID, "colormap LUT must be 16-byte aligned");
    }
    CUDA_TRY(cudaSetDevice(c->device));
    CmapArgs a;
    a.image = image; a.res = image_res; a.channels = channels; a.p = *params;
    a.lut = lut; a.lut_w = lut_w; a.lut_h = lut_h; a.out = out; a.out_w = out_w; a.out_h = out_h; a.out_fmt = out_fmt;
    dim3 grid((out_w + 255) / 256, out_h);
    k_colormap<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    c->launches++;
    c->last_stream = (cudaStream_t)stream;
    CUDA_TRY(cudaGetLastError());
    return TSPLAT_OK;
}

extern "C" int tsplat_set_surface(tsplat_ctx *c, const float *host_lut, int n_floats, float density_cut)
{
    if (!c) return set_err(TSPLAT_ERR_INVALID, "NULL context");
    if (host_lut) {
        if (n_floats != LUT_TOTAL) return set_err(TSPLAT_ERR_INVALID, "surface LUT must have %d floats, got %d", LUT_TOTAL, n_floats);
        CUDA_TRY(cudaSetDevice(c->device));
        CUDA_TRY(cudaMemcpy(c->d_slut, host_lut, LUT_TOTAL * sizeof(float), cudaMemcpyHostToDevice));
        c->surface_set = true;
    } else if (!c->surface_set) {
        return set_err(TSPLAT_ERR_STATE, "surface LUT not set yet: pass it with the first call");
    }
    c->density_cut = density_cut;
    return TSPLAT_OK;
}

extern "C" int tsplat_bilateral_filter(tsplat_ctx *c, const float *in, float *out, int width, int height, float spatial_sigma,
                                       float range_sigma, int kernel_size, void *stream)
{
    if (!c || !in || !out) return set_err(TSPLAT_ERR_INVALID, "NULL argument");
    if (in == out) return set_err(TSPLAT_ERR_INVALID, "in and out must differ");
    if (width <= 0 || height <= 0 || kernel_size < 0 || kernel_size > 4096) return set_err(TSPLAT_ERR_INVALID, "bad size");
    if (!(spatial_sigma > 0.0f) || !(range_sigma > 0.0f)) return set_err(TSPLAT_ERR_INVALID, "sigmas must be positive");
    CUDA_TRY(cudaSetDevice(c->device));
    tsplat_surface::BilateralArgs a;
    a.in = in; a.out = out; a.width = width; a.height = height; a.spatial_sigma = spatial_sigma; a.range_sigma = range_sigma;
    a.kernel_size = kernel_size;
    dim3 grid((width + 31) / 32, (height + 7) / 8);
    const size_t smem = tsplat_surface::bilateral_smem_bytes(kernel_size / 2);
    static const bool force_direct = getenv("TSPLAT_BILATERAL_DIRECT") != nullptr;      // A/B timing of the round-1 kernel
    if (smem <= 200 * 1024 && !force_direct) {
        if (smem > 48 * 1024 && smem > c->bilateral_smem_set) {
            CUDA_TRY(cudaFuncSetAttribute(tsplat_surface::k_bilateral_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            c->bilateral_smem_set = smem;
        }
        tsplat_surface::k_bilateral_filter<<<grid, tsplat_surface::BF_TX * tsplat_surface::BF_TY, smem, (cudaStream_t)stream>>>(a);
    } else {
        tsplat_surface::k_bilateral_filter_direct<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    }
    c->launches++;
    c->last_stream = (cudaStream_t)stream;
    CUDA_TRY(cudaGetLastError());
    return TSPLAT_OK;
}

extern "C" int tsplat_surface_shade(tsplat_ctx *c, const float *smoothed, int res, const tsplat_surface_params *params,
                                    const float *lut, int lut_w, void *out, int out_w, int out_h, int out_fmt, void *stream)
{
    if (!c || !smoothed || !params || !out) return set_err(TSPLAT_ERR_INVALID, "NULL argument");
    if (res <= 0 || out_w <= 0 || out_h <= 0) return set_err(TSPLAT_ERR_INVALID, "bad size");
    if (out_fmt < TSPLAT_FMT_RGBA8 || out_fmt > TSPLAT_FMT_RGBA32F) return set_err(TSPLAT_ERR_INVALID, "bad output format");
    if (params->material_colormap && (!lut || lut_w <= 0)) return set_err(TSPLAT_ERR_INVALID, "material colormap LUT missing");
    CUDA_TRY(cudaSetDevice(c->device));
    tsplat_surface::ShadeArgs a;
    a.image = smoothed; a.res = res; a.p = *params; a.lut = lut; a.lut_w = lut_w; a.out = out; a.out_w = out_w; a.out_h = out_h;
    a.out_fmt = out_fmt;
    dim3 grid((out_w + 255) / 256, out_h);
    k_surface_shade<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    c->launches++;
    c->last_stream = (cudaStream_t)stream;
    CUDA_TRY(cudaGetLastError());
    return TSPLAT_OK;
}

extern "C" int tsplat_reduce_colormap(tsplat_ctx *c, const float *const *peer_images, int n_peers, int channels,
                                      int row0, int nrows, const tsplat_colormap_params *params, const float *lut,
                                      int lut_w, int lut_h, void *out, int out_fmt, float *sum_out, void *stream)
{
    if (!c || !peer_images) return set_err(TSPLAT_ERR_INVALID, "NULL argument");
    if (n_peers < 1 || n_peers > MAX_PEERS) return set_err(TSPLAT_ERR_INVALID, "n_peers must be in [1, %d]", MAX_PEERS);
    if (channels != 1 && channels != 2 && channels != 4) return set_err(TSPLAT_ERR_INVALID, "channels must be 1, 2 or 4");
    if (row0 < 0 || nrows < 0 || row0 + nrows > c->R) return set_err(TSPLAT_ERR_INVALID, "row slab outside the image");
    if (out && !params) return set_err(TSPLAT_ERR_INVALID, "colormap parameters missing");
    if (out && (out_fmt < TSPLAT_FMT_RGBA8 || out_fmt > TSPLAT_FMT_RGBA32F)) return set_err(TSPLAT_ERR_INVALID, "bad output format");
    if (out && params->kind != TSPLAT_CMAP_RGB && (!lut || lut_w <= 0 || lut_h <= 0))
        return set_err(TSPLAT_ERR_INVALID, "colormap LUT missing");
    if (out && params->kind == TSPLAT_CMAP_RGB && channels < 4) return set_err(TSPLAT_ERR_INVALID, "RGB map needs a 4-channel image");
    CUDA_TRY(cudaSetDevice(c->device));
    if (nrows == 0) return TSPLAT_OK;
    ReduceArgs a;
    memset(&a, 0, sizeof(a));
    for (int r = 0; r < n_peers; ++r) {
        if (!peer_images[r]) return set_err(TSPLAT_ERR_INVALID, "NULL peer image %d", r);
        a.peer[r] = peer_images[r];
    }
    a.n_peers = n_peers; a.res = c->R; a.channels = channels; a.row0 = row0; a.nrows = nrows;
    if (params) a.p = *params;
    a.lut = lut; a.lut_w = lut_w; a.lut_h = lut_h; a.out = out; a.out_fmt = out_fmt; a.sum_out = sum_out;
    const int64_t quads = ((int64_t)nrows * c->R * channels + 3) / 4;
    const unsigned rgrid = (unsigned)((quads + 255) / 256);
    cudaStream_t rst = (cudaStream_t)stream;
    switch (n_peers) {
    case 1: k_reduce_colormap<1><<<rgrid, 256, 0, rst>>>(a); break;
    case 2: k_reduce_colormap<2><<<rgrid, 256, 0, rst>>>(a); break;
    case 3: k_reduce_colormap<3><<<rgrid, 256, 0, rst>>>(a); break;
    case 4: k_reduce_colormap<4><<<rgrid, 256, 0, rst>>>(a); break;
    case 5: k_reduce_colormap<5><<<rgrid, 256, 0, rst>>>(a); break;
    case 6: k_reduce_colormap<6><<<rgrid, 256, 0, rst>>>(a); break;
    case 7: k_reduce_colormap<7><<<rgrid, 256, 0, rst>>>(a); break;
    case 8: k_reduce_colormap<8><<<rgrid, 256, 0, rst>>>(a); break;
    default: k_reduce_colormap<MAX_PEERS><<<rgrid, 256, 0, rst>>>(a); break;
    }
    c->launches++;
    c->last_stream = (cudaStream_t)stream;
    CUDA_TRY(cudaGetLastError());
    return TSPLAT_OK;
}

extern "C" int tsplat_enable_kernel_timing(tsplat_ctx *c, int enable)
{
    if (!c) return set_err(TSPLAT_ERR_INVALID, "NULL context");
    CUDA_TRY(cudaSetDevice(c->device));
    if (enable && !c->timing_events_created) {
        for (int i = 0; i < TSPLAT_TIMING_SLOTS; ++i) {
            CUDA_TRY(cudaEventCreate(&c->t_begin[i]));
            CUDA_TRY(cudaEventCreate(&c->t_end[i]));
        }
        c->timing_events_created = true;
    }
    c->timing = enable != 0;
    c->timing_read = c->timing_write = 0;
    return TSPLAT_OK;
}

extern "C" int tsplat_kernel_timing(tsplat_ctx *c, int64_t *n_launches, double *total_ms)
{
    if (!c || !n_launches || !total_ms) return set_err(TSPLAT_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    *n_launches = 0; *total_ms = 0.0;
    for (; c->timing_read < c->timing_write; ++c->timing_read) {
        const int i = (int)(c->timing_read % TSPLAT_TIMING_SLOTS);
        CUDA_TRY(cudaEventSynchronize(c->t_end[i]));
        float ms = 0.0f;
        CUDA_TRY(cudaEventElapsedTime(&ms, c->t_begin[i], c->t_end[i]));
        *total_ms += ms; ++*n_launches;
    }
    return TSPLAT_OK;
}

extern "C" int tsplat_enable_peer_access(int device_ordinal, int peer_ordinal)
{
    CUDA_TRY(cudaSetDevice(device_ordinal));
    int can = 0;
    CUDA_TRY(cudaDeviceCanAccessPeer(&can, device_ordinal, peer_ordinal));
    if (!can) return set_err(TSPLAT_ERR_STATE, "device %d cannot access device %d", device_ordinal, peer_ordinal);
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_ordinal, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return TSPLAT_OK; }
    CUDA_TRY(e);
    return TSPLAT_OK;
}

extern "C" int tsplat_allreduce_image(tsplat_ctx *c, const float *const *peer_images, float *const *peer_out,
                                      const float *const *peer_scale, int n_peers, int channels, int row0, int nrows, int op,
                                      void *stream)
{
    if (!c || !peer_images || !peer_out) return set_err(TSPLAT_ERR_INVALID, "NULL argument");
    if (n_peers < 1 || n_peers > MAX_PEERS) return set_err(TSPLAT_ERR_INVALID, "n_peers must be in [1, %d]", MAX_PEERS);
    if (channels != 1 && channels != 2 && channels != 4) return set_err(TSPLAT_ERR_INVALID, "channels must be 1, 2 or 4");
    if (row0 < 0 || nrows < 0 || row0 + nrows > c->R) return set_err(TSPLAT_ERR_INVALID, "row slab outside the image");
    if (op != TSPLAT_REDUCE_SUM && op != TSPLAT_REDUCE_ZMAX) return set_err(TSPLAT_ERR_INVALID, "unknown reduction %d", op);
    if (op == TSPLAT_REDUCE_ZMAX && channels != 2) return set_err(TSPLAT_ERR_INVALID, "the z-buffer reduction needs a 2-channel image");
    if (op == TSPLAT_REDUCE_SUM && !peer_scale) return set_err(TSPLAT_ERR_INVALID, "peer scales missing");
    CUDA_TRY(cudaSetDevice(c->device));
    if (nrows == 0) return TSPLAT_OK;
    AllReduceArgs a;
    memset(&a, 0, sizeof(a));
    for (int r = 0; r < n_peers; ++r) {
        if (!peer_images[r] || !peer_out[r] || (op == TSPLAT_REDUCE_SUM && !peer_scale[r]))
            return set_err(TSPLAT_ERR_INVALID, "NULL peer pointer %d", r);
        a.peer[r] = peer_images[r]; a.out[r] = peer_out[r]; a.scale[r] = op == TSPLAT_REDUCE_SUM ? peer_scale[r] : nullptr;
    }
    a.n_peers = n_peers; a.op = op;
    a.first = (int64_t)row0 * c->R * channels;
    a.count = (int64_t)nrows * c->R * channels;
    int64_t blocks = (a.count / 4 + 255) / 256;
    if (blocks > (int64_t)c->sm_count * 8) blocks = (int64_t)c->sm_count * 8;
    if (blocks < 1) blocks = 1;
    cudaStream_t ast = (cudaStream_t)stream;
    switch (n_peers) {
    case 2: k_allreduce_image<2><<<(unsigned)blocks, 256, 0, ast>>>(a); break;
    case 4: k_allreduce_image<4><<<(unsigned)blocks, 256, 0, ast>>>(a); break;
    case 8: k_allreduce_image<8><<<(unsigned)blocks, 256, 0, ast>>>(a); break;
    default: k_allreduce_image<MAX_PEERS><<<(unsigned)blocks, 256, 0, ast>>>(a); break;
    }
    c->launches++;
    c->last_stream = (cudaStream_t)stream;
    CUDA_TRY(cudaGetLastError());
    return TSPLAT_OK;
}

extern "C" int tsplat_periodic_accumulate(tsplat_ctx *c, const float *src, float *dst, int channels, const float *offsets_xy,
                                          const float *weights, int n, void *stream)
{
    if (!c || !src || !dst || (n > 0 && (!offsets_xy || !weights))) return set_err(TSPLAT_ERR_INVALID, "NULL argument");
    if (src == dst) return set_err(TSPLAT_ERR_INVALID, "src and dst must differ");
    if (channels != 1 && channels != 2 && channels != 4) return set_err(TSPLAT_ERR_INVALID, "channels must be 1, 2 or 4");
    if (n < 0 || n > MAX_REPLICAS) return set_err(TSPLAT_ERR_INVALID, "replica count %d outside [0, %d]", n, MAX_REPLICAS);
    CUDA_TRY(cudaSetDevice(c->device));
    PeriodicArgs a;
    memset(&a, 0, sizeof(a));
    a.src = src; a.dst = dst; a.res = c->R; a.channels = channels; a.n = n;
    for (int i = 0; i < n; ++i) { a.ox[i] = offsets_xy[2 * i]; a.oy[i] = offsets_xy[2 * i + 1]; a.w[i] = weights[i]; }
    dim3 grid((c->R + 255) / 256, c->R);
    k_periodic_accumulate<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    c->launches++;
    c->last_stream = (cudaStream_t)stream;
    CUDA_TRY(cudaGetLastError());
    return TSPLAT_OK;
}

static int content_args(tsplat_ctx *c, const float *image, int res, int channels, int content, float scale, ContentArgs *a)
{
    if (!c || !image) return set_err(TSPLAT_ERR_INVALID, "NULL argument");
    if (channels != 1 && channels != 2 && channels != 4) return set_err(TSPLAT_ERR_INVALID, "channels must be 1, 2 or 4");
    if (res <= 0) return set_err(TSPLAT_ERR_INVALID, "bad resolution");
    if (content < CONTENT_CH0 || content > CONTENT_CH0_WHERE_CH1) return set_err(TSPLAT_ERR_INVALID, "bad content kind");
    if ((content == CONTENT_RATIO || content == CONTENT_CH0_WHERE_CH1) && channels < 2)
        return set_err(TSPLAT_ERR_INVALID, "this content kind needs two channels");
    a->image = image; a->channels = channels; a->content = content; a->scale = scale; a->use_log = 0;
    a->n_values = (int64_t)res * res * (content == CONTENT_ALL ? channels : 1);
    return TSPLAT_OK;
}

extern "C" int tsplat_content_stats(tsplat_ctx *c, const float *image, int res, int channels, int content, float scale,
                                    tsplat_content_stats_t *out, void *stream)
{
    if (!out) return set_err(TSPLAT_ERR_INVALID, "NULL argument");
    ContentArgs a;
    int rc = content_args(c, image, res, channels, content, scale, &a);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    ContentStats init;
    memset(&init, 0, sizeof(init));
    init.lin_min_key = init.log_min_key = 0xffffffffu;
    ContentStats *d = reinterpret_cast<ContentStats *>(c->d_select);
    CUDA_TRY(cudaMemcpyAsync(d, &init, sizeof(init), cudaMemcpyHostToDevice, st));
    k_content_stats<<<c->sm_count * 8, 256, 0, st>>>(a, d);
    c->launches++;
    ContentStats h;
    CUDA_TRY(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const float nanf_ = nanf("");
    out->n_finite_lin = (int64_t)h.n_finite_lin; out->n_finite_log = (int64_t)h.n_finite_log;
    out->any_negative = (int32_t)h.any_negative;
    out->lin_min = h.n_finite_lin ? key_to_float(h.lin_min_key) : nanf_;
    out->lin_max = h.n_finite_lin ? key_to_float(h.lin_max_key) : nanf_;
    out->log_min = h.n_finite_log ? key_to_float(h.log_min_key) : nanf_;
    out->log_max = h.n_finite_log ? key_to_float(h.log_max_key) : nanf_;
    return TSPLAT_OK;
}

extern "C" int tsplat_content_select(tsplat_ctx *c, const float *image, int res, int channels, int content, float scale,
                                     int use_log, const int64_t *ranks, int n_ranks, float *out_values, void *stream)
{
    if (!ranks || !out_values) return set_err(TSPLAT_ERR_INVALID, "NULL argument");
    if (n_ranks < 1 || n_ranks > SELECT_MAX_RANKS) return set_err(TSPLAT_ERR_INVALID, "n_ranks must be in [1, %d]", SELECT_MAX_RANKS);
    SelectArgs a;
    memset(&a, 0, sizeof(a));
    int rc = content_args(c, image, res, channels, content, scale, &a.c);
    if (rc) return rc;
    a.c.use_log = use_log ? 1 : 0;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    a.n_ranks = n_ranks;
    a.hist = reinterpret_cast<unsigned *>(c->d_select);
    int64_t remaining[SELECT_MAX_RANKS];
    for (int r = 0; r < n_ranks; ++r) { if (ranks[r] < 0) return set_err(TSPLAT_ERR_INVALID, "negative rank"); remaining[r] = ranks[r]; }
    unsigned *h_hist = c->h_select;
    const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
    for (int pass = 0; pass < 3; ++pass) {
        a.shift = shifts[pass]; a.bits = bits[pass];
        CUDA_TRY(cudaMemsetAsync(a.hist, 0, sizeof(unsigned) * SELECT_MAX_RANKS * 2048, st));
        k_content_select<<<c->sm_count * 4, 256, 0, st>>>(a);
        c->launches++;
        CUDA_TRY(cudaMemcpyAsync(h_hist, a.hist, sizeof(unsigned) * n_ranks * 2048, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        for (int r = 0; r < n_ranks; ++r) {
            const unsigned *hh = h_hist + r * 2048;
            int64_t acc = 0;
            int bin = -1;
            for (int b = 0; b < (1 << bits[pass]); ++b) {
                if (remaining[r] < acc + (int64_t)hh[b]) { bin = b; break; }
                acc += hh[b];
            }
            if (bin < 0) return set_err(TSPLAT_ERR_INVALID, "rank %lld beyond the number of finite values", (long long)ranks[r]);
            remaining[r] -= acc;
            a.prefix[r] |= (unsigned)bin << shifts[pass];
        }
        a.prefix_mask |= ((1u << bits[pass]) - 1u) << shifts[pass];
    }
    for (int r = 0; r < n_ranks; ++r) out_values[r] = key_to_float(a.prefix[r]);
    return TSPLAT_OK;
}

extern "C" int tsplat_image_axpy(tsplat_ctx *c, float *dst, const float *src, float scale, int64_t n, void *stream)
{
    if (!c || !dst || !src || n < 0) return set_err(TSPLAT_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    if (n == 0) return TSPLAT_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > c->sm_count * 32) blocks = c->sm_count * 32;
    k_axpy<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, src, scale, n);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return TSPLAT_OK;
}

extern "C" int tsplat_memcpy_h2d(void *dst_dev, const void *src_host, int64_t bytes, void *stream)
{
    if (bytes < 0 || (bytes > 0 && (!dst_dev || !src_host))) return set_err(TSPLAT_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaMemcpyAsync(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return TSPLAT_OK;
}

extern "C" int tsplat_memcpy_d2h(void *dst_host, const void *src_dev, int64_t bytes, void *stream)
{
    if (bytes < 0 || (bytes > 0 && (!dst_host || !src_dev))) return set_err(TSPLAT_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return TSPLAT_OK;
}

extern "C" int tsplat_stream_sync(void *stream)
{
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return TSPLAT_OK;
}

extern "C" int tsplat_get_stats(tsplat_ctx *c, tsplat_stats *out)
{
    if (!c || !out) return set_err(TSPLAT_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    Counters h;
    CUDA_TRY(cudaStreamSynchronize(c->last_stream));
    CUDA_TRY(cudaMemcpy(&h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < STAT_SLOTS; ++i) {
        h.culled += h.slots[i].culled_direct & 0xffffffffull; h.direct += h.slots[i].culled_direct >> 32; h.reds += h.slots[i].reds;
    }
    out->particles_culled = (int64_t)h.culled;
    out->particles_direct = (int64_t)h.direct;
    out->particles_tiled = (int64_t)h.tiled;
    out->particles_huge = (int64_t)h.huge;
    out->tile_pairs = (int64_t)h.pairs;
    out->particles_submitted = (int64_t)(h.culled + h.direct + h.tiled + h.huge);
    out->direct_vector_reds = (int64_t)h.reds;
    out->kernel_launches = c->launches;
    return TSPLAT_OK;
}

// cell layout entry points are in tsplat_cells.cu
