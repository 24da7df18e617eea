// tsplat_cells.cu -- K4: CellLayout.from_positions on the device.
//
// Replaces the numpy pipeline of src/topsy/cell_layout.py:87-112 for device-resident positions:
//   pos_indices  = floor((pos - box_min) / cell_size).astype(intp)          (:97)   -- in the position dtype
//   cell_indices = iz + nside * (iy + nside * ix)                            (:104)
//   ordering     = argsort(cell_indices)                                      (:106)  -- here: the STABLE argsort
//   lengths      = bincount(cell_indices, minlength=nside^3)                  (:109)
// (offsets = cumsum(lengths) - lengths and the mgrid centres are O(nside^3) host work.)
//
// numpy's default argsort is not stable, so the order of particles *inside* a cell is unspecified by the reference
// (and is re-randomised by randomize_within_cells, cell_layout.py:17-24); what is bit-exact is the cell assignment.
// This implementation returns the stable order (== np.argsort(kind='stable')), via a hand-written counting sort:
//   k_cell_keys    keys + per-(cell, warp) histogram
//   k_cell_units   per-cell exclusive scan over warps  -> lengths
//   k_cell_offsets exclusive scan over cells           -> offsets
//   k_cell_scatter each warp walks its contiguous slice in order; rank inside a 32-particle step via match_any
// Compiled with -fmad=false: (pos - box_min) / cell_size is an IEEE subtract and an IEEE divide, as in numpy.
#include "../../include/tsplat.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int WARPS_PER_CTA = 8;

struct CellPlan {
    int64_t per_warp;      // particles per warp slice (multiple of 32)
    int units;             // number of warps
    int64_t keys_off, table_off, offsets_off, total;
};

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

CellPlan make_plan(int64_t n, int nside)
{
    CellPlan p;
    const int64_t ncell = (int64_t)nside * nside * nside;
    int64_t units = (n + 2047) / 2048;
    if (units < 1) units = 1;
    if (units > 148 * 32) units = 148 * 32;
    units = align_up(units, WARPS_PER_CTA);
    p.units = (int)units;
    p.per_warp = align_up((n + units - 1) / units, 32);
    if (p.per_warp < 32) p.per_warp = 32;
    p.keys_off = 0;
    p.table_off = align_up(n * 4, 256);
    p.offsets_off = p.table_off + align_up(ncell * units * 4, 256);
    p.total = p.offsets_off + align_up(ncell * 8, 256);
    return p;
}

template <typename T>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_cell_keys(const T *__restrict__ pos, int64_t n, T box_min, T cell_size,
                                                               int nside, int64_t per_warp, int units,
                                                               int *__restrict__ keys, unsigned int *__restrict__ table,
                                                               int32_t *__restrict__ status)
{
    const int unit = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int64_t begin = (int64_t)unit * per_warp;
    int64_t end = begin + per_warp;
    if (end > n) end = n;
    for (int64_t i = begin + lane; i < end; i += 32) {
        const T px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2];
        const long long ix = (long long)floor((px - box_min) / cell_size);
        const long long iy = (long long)floor((py - box_min) / cell_size);
        const long long iz = (long long)floor((pz - box_min) / cell_size);
        int key;
        if (ix < 0 || iy < 0 || iz < 0 || ix >= nside || iy >= nside || iz >= nside) {
            *status = 1;          // reference: ValueError("... too close to edge of box")
            key = 0;
        } else {
            key = (int)(iz + nside * (iy + (long long)nside * ix));
        }
        keys[i] = key;
        atomicAdd(&table[(size_t)key * units + unit], 1u);
    }
}

// one CTA per cell: exclusive scan over the `units` per-warp counts, total -> lengths[cell]
__global__ void __launch_bounds__(256) k_cell_units(unsigned int *__restrict__ table, int units, int64_t *__restrict__ lengths)
{
    __shared__ unsigned int s_part[256];
    unsigned int *row = table + (size_t)blockIdx.x * units;
    const int per = (units + 255) / 256;
    const int b = threadIdx.x * per;
    unsigned int sum = 0;
    for (int i = b; i < b + per && i < units; ++i) sum += row[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int run = 0;
        for (int i = 0; i < 256; ++i) { const unsigned int t = s_part[i]; s_part[i] = run; run += t; }
        lengths[blockIdx.x] = (int64_t)run;
    }
    __syncthreads();
    unsigned int run = s_part[threadIdx.x];
    for (int i = b; i < b + per && i < units; ++i) { const unsigned int t = row[i]; row[i] = run; run += t; }
}

__global__ void __launch_bounds__(1024) k_cell_offsets(const int64_t *__restrict__ lengths, int ncell, int64_t *__restrict__ offsets)
{
    __shared__ int64_t s_part[1024];
    const int per = (ncell + 1023) / 1024;
    const int b = threadIdx.x * per;
    int64_t sum = 0;
    for (int i = b; i < b + per && i < ncell; ++i) sum += lengths[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t run = 0;
        for (int i = 0; i < 1024; ++i) { const int64_t t = s_part[i]; s_part[i] = run; run += t; }
    }
    __syncthreads();
    int64_t run = s_part[threadIdx.x];
    for (int i = b; i < b + per && i < ncell; ++i) { offsets[i] = run; run += lengths[i]; }
}

__device__ __forceinline__ unsigned mix32(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// A keyed pseudo-random PERMUTATION of [0, n): 4-round Feistel network on the smallest even number of bits that holds n,
// cycle-walked back into range (at most 4 n candidates, so < 4 walks on average).  Every slot computes its image on its
// own: the within-cell shuffle of CellLayout.randomize_within_cells (cell_layout.py:17-24) without a sort.
__device__ __forceinline__ unsigned feistel_permute(unsigned r, unsigned n, unsigned key)
{
    if (n <= 1u) return 0u;
    unsigned bits = 32u - (unsigned)__clz(n - 1u);
    if (bits < 2u) bits = 2u;
    bits += bits & 1u;
    const unsigned half = bits >> 1, mask = (1u << half) - 1u;
    do {
        unsigned L = r >> half, R = r & mask;
#pragma unroll
        for (unsigned round = 0; round < 4u; ++round) {
            const unsigned t = L ^ (mix32(R ^ key ^ (round * 0x9e3779b9u)) & mask);
            L = R; R = t;
        }
        r = (L << half) | R;
    } while (r >= n);
    return r;
}

// shuffle_seed == 0: stable order (== np.argsort(kind='stable')); otherwise slot s of cell c goes to
// feistel_permute(s, lengths[c], hash(seed, c)): a pseudo-random permutation inside every cell, never across cells
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_cell_scatter(const int *__restrict__ keys, int64_t n, int64_t per_warp, int units,
                                                                  unsigned int *__restrict__ table,
                                                                  const int64_t *__restrict__ offsets,
                                                                  const int64_t *__restrict__ lengths, unsigned shuffle_seed,
                                                                  int64_t *__restrict__ order)
{
    const int unit = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int64_t begin = (int64_t)unit * per_warp;
    int64_t end = begin + per_warp;
    if (end > n) end = n;
    for (int64_t i0 = begin; i0 < end; i0 += 32) {
        const int64_t i = i0 + lane;
        const bool valid = i < end;
        const unsigned active = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const int key = keys[i];
            const unsigned peers = __match_any_sync(active, key);
            const int leader = __ffs(peers) - 1;
            const int rank = __popc(peers & ((1u << lane) - 1u));
            unsigned int base = 0;
            if (lane == leader) base = atomicAdd(&table[(size_t)key * units + unit], (unsigned)__popc(peers));
            base = __shfl_sync(peers, base, leader);
            unsigned slot = base + (unsigned)rank;
            if (shuffle_seed) slot = feistel_permute(slot, (unsigned)lengths[key], mix32(shuffle_seed ^ mix32((unsigned)key + 0x632be5abu)));
            order[offsets[key] + slot] = i;
        }
        __syncwarp();
    }
}

}  // namespace

extern "C" int64_t tsplat_cell_layout_work_bytes(int64_t n, int nside)
{
    if (n < 0 || nside <= 0 || nside > 1024) return -1;
    return make_plan(n, nside).total;
}

static int cell_layout_impl(int device_ordinal, const void *pos, int64_t n, int dtype_bytes, double box_min,
                            double cell_size, int nside, unsigned shuffle_seed, int64_t *order, int64_t *lengths, int32_t *status,
                            void *work, int64_t work_bytes, void *stream)
{
    if (n < 0 || nside <= 0 || nside > 1024 || (dtype_bytes != 4 && dtype_bytes != 8)) return TSPLAT_ERR_INVALID;
    if (!lengths || !status || !work || (n > 0 && (!pos || !order))) return TSPLAT_ERR_INVALID;
    const CellPlan p = make_plan(n, nside);
    if (work_bytes < p.total) return TSPLAT_ERR_INVALID;
    if (cudaSetDevice(device_ordinal) != cudaSuccess) return TSPLAT_ERR_CUDA;
    cudaStream_t st = (cudaStream_t)stream;
    const int ncell = nside * nside * nside;
    char *w = static_cast<char *>(work);
    int *keys = reinterpret_cast<int *>(w + p.keys_off);
    unsigned int *table = reinterpret_cast<unsigned int *>(w + p.table_off);
    int64_t *offsets = reinterpret_cast<int64_t *>(w + p.offsets_off);
    if (cudaMemsetAsync(table, 0, (size_t)ncell * p.units * 4, st) != cudaSuccess) return TSPLAT_ERR_CUDA;
    if (cudaMemsetAsync(status, 0, sizeof(int32_t), st) != cudaSuccess) return TSPLAT_ERR_CUDA;
    const int grid = p.units / WARPS_PER_CTA;
    if (n > 0) {
        if (dtype_bytes == 4)
            k_cell_keys<float><<<grid, WARPS_PER_CTA * 32, 0, st>>>(static_cast<const float *>(pos), n, (float)box_min,
                                                                  (float)cell_size, nside, p.per_warp, p.units, keys, table, status);
        else
            k_cell_keys<double><<<grid, WARPS_PER_CTA * 32, 0, st>>>(static_cast<const double *>(pos), n, box_min, cell_size,
                                                                   nside, p.per_warp, p.units, keys, table, status);
    }
    k_cell_units<<<ncell, 256, 0, st>>>(table, p.units, lengths);
    k_cell_offsets<<<1, 1024, 0, st>>>(lengths, ncell, offsets);
    if (n > 0)
        k_cell_scatter<<<grid, WARPS_PER_CTA * 32, 0, st>>>(keys, n, p.per_warp, p.units, table, offsets, lengths, shuffle_seed, order);
    if (cudaGetLastError() != cudaSuccess) return TSPLAT_ERR_CUDA;
    return TSPLAT_OK;
}

extern "C" int tsplat_cell_layout(int device_ordinal, const void *pos, int64_t n, int dtype_bytes, double box_min,
                                  double cell_size, int nside, int64_t *order, int64_t *lengths, int32_t *status,
                                  void *work, int64_t work_bytes, void *stream)
{
    return cell_layout_impl(device_ordinal, pos, n, dtype_bytes, box_min, cell_size, nside, 0u, order, lengths, status, work,
                            work_bytes, stream);
}

extern "C" int tsplat_cell_layout_shuffled(int device_ordinal, const void *pos, int64_t n, int dtype_bytes, double box_min,
                                           double cell_size, int nside, uint32_t shuffle_seed, int64_t *order, int64_t *lengths,
                                           int32_t *status, void *work, int64_t work_bytes, void *stream)
{
    if (n >= ((int64_t)1 << 31)) return TSPLAT_ERR_INVALID;      // slots inside a cell are permuted as 32-bit numbers
    return cell_layout_impl(device_ordinal, pos, n, dtype_bytes, box_min, cell_size, nside, shuffle_seed ? shuffle_seed : 1u, order,
                            lengths, status, work, work_bytes, stream);
}

// dst[i] = (float) src[order[i] * stride + offset]: the loaders' "array[ordering]" (loader.py:100-110 of the reference) on
// the device; src is float32 (src_bytes 4) or float64 (8)
template <typename T>
__global__ void __launch_bounds__(256) k_gather_f32(float *__restrict__ dst, const T *__restrict__ src, const int64_t *__restrict__ order,
                                                   int64_t n, int stride, int offset)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = (float)src[order[i] * stride + offset];
}

extern "C" int tsplat_gather_f32(int device_ordinal, float *dst, const void *src, int src_bytes, int stride, int offset,
                                 const int64_t *order, int64_t n, void *stream)
{
    if (n < 0 || stride < 1 || offset < 0 || offset >= stride || (src_bytes != 4 && src_bytes != 8)) return TSPLAT_ERR_INVALID;
    if (n > 0 && (!dst || !src || !order)) return TSPLAT_ERR_INVALID;
    if (cudaSetDevice(device_ordinal) != cudaSuccess) return TSPLAT_ERR_CUDA;
    if (n == 0) return TSPLAT_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (src_bytes == 4)
        k_gather_f32<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, static_cast<const float *>(src), order, n, stride, offset);
    else
        k_gather_f32<double><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, static_cast<const double *>(src), order, n, stride, offset);
    if (cudaGetLastError() != cudaSuccess) return TSPLAT_ERR_CUDA;
    return TSPLAT_OK;
}
