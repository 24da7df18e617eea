// Surface render mode: kernels K9-K11.  Included by tsplat.cu (shares its Camera / ProjectArgs / queue machinery).
//
//   K9  k_project_surface / k_queue_surface   z-buffered splat of the particles above a density cut
//         replaces DepthSPHWithOcclusion's render pass: vertex_depth_with_cut + fragment_raw with depth_compare =
//         greater, blending off (src/topsy/sph.py:457-601, src/topsy/shaders/sph.wgsl:93-158).
//         The (quantity, depth) float2 pixel IS the 64-bit z-buffer key: depth (non-negative float, so its bit pattern
//         orders like the value) in the high word, quantity in the low word; one RED.MAX.U64 per fragment.  The result
//         is independent of draw order (ties in depth go to the larger quantity bit pattern), unlike the reference,
//         where the first fragment drawn wins a tie and depths beyond 1.0 are clamped before the test.
//   K10 k_bilateral_filter                    replaces shaders/smooth.wgsl (colormap/surface.py:262-297)
//   K11 k_surface_shade                       replaces shaders/surface.wgsl (colormap/surface.py:341-369)
#pragma once

namespace tsplat_surface {
using namespace tsplat;

constexpr int SURF_INLINE_SPAN = 4;          // footprints up to 4 x 4 pixel centres are drawn by the projecting thread

__device__ __forceinline__ void zbuffer_fragment(float *__restrict__ image, size_t pix, float q, float depth)
{
    if (depth > 0.0f) {                       // the depth buffer is cleared to 0 and the test is "greater"
        const unsigned long long key = ((unsigned long long)__float_as_uint(depth) << 32) | __float_as_uint(q);
        atomicMax(reinterpret_cast<unsigned long long *>(image) + pix, key);
    }
}

// one fragment of fragment_raw (sph.wgsl:148-158)
__device__ __forceinline__ void surface_fragment(float *__restrict__ image, const float *__restrict__ lut, int R, int j, int k,
                                                 float px0, float py1, float wpx, float inv, float q, float cz, float hz)
{
    const float K = sample_lut(lut, wpx, inv, px0, py1, (float)j + 0.5f, (float)k + 0.5f);
    if (K < 0.0f) return;                     // outside the sphere: discard
    const float t = hz * K;
    zbuffer_fragment(image, (size_t)k * R + j, q, cz + t);
}

// ProjectArgs: w0 = mass, w1 = quantity; image = (R, R, 2); lut = local-sphere mip chain (device)
__global__ void __launch_bounds__(256) k_project_surface(const ProjectArgs a, const float density_cut)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t group = 0;
    int e_first = 0, e_last = 0;
    const bool active = gi < a.n_groups;
    if (active) {
        int64_t lo, hi;
        if (a.table.n > 0) {
            int l = 0, r = a.table.n;            // invariant: gprefix[l] <= gi < gprefix[r]
            while (r - l > 1) {
                const int m = (l + r) >> 1;
                if (a.table.gprefix[m] <= gi) l = m; else r = m;
            }
            lo = a.table.start[l];
            hi = a.table.end[l];
            group = (lo >> 2) + (gi - a.table.gprefix[l]);
        } else {
            lo = a.start; hi = a.end; group = a.g0 + gi;
        }
        const int64_t base = group << 2;
        e_first = (int)max((int64_t)0, lo - base);
        e_last = (int)min((int64_t)4, hi - base);
    }
    float xs[4], ys[4], zs[4], hs[4], ms[4], qs[4];
    {
        float4 X, Y, Z, H, W0, W1;
        X = Y = Z = H = W0 = W1 = make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t base = group << 2;
        if (active) {
            if (base + 4 <= a.n_total) {
                const uint64_t pol = l2_policy_evict_first();
                X = ld4(a.x, group, pol); Y = ld4(a.y, group, pol); Z = ld4(a.z, group, pol); H = ld4(a.h, group, pol);
                W0 = ld4(a.w0, group, pol); W1 = ld4(a.w1, group, pol);
            } else {
                const int64_t left = a.n_total - base;
                X = ld4_tail(a.x, base, left); Y = ld4_tail(a.y, base, left); Z = ld4_tail(a.z, base, left);
                H = ld4_tail(a.h, base, left); W0 = ld4_tail(a.w0, base, left); W1 = ld4_tail(a.w1, base, left);
            }
        }
        xs[0] = X.x; xs[1] = X.y; xs[2] = X.z; xs[3] = X.w;   ys[0] = Y.x; ys[1] = Y.y; ys[2] = Y.z; ys[3] = Y.w;
        zs[0] = Z.x; zs[1] = Z.y; zs[2] = Z.z; zs[3] = Z.w;   hs[0] = H.x; hs[1] = H.y; hs[2] = H.z; hs[3] = H.w;
        ms[0] = W0.x; ms[1] = W0.y; ms[2] = W0.z; ms[3] = W0.w;   qs[0] = W1.x; qs[1] = W1.y; qs[2] = W1.z; qs[3] = W1.w;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        bool defer = false;
        float4 dq0 = make_float4(0.f, 0.f, 0.f, 0.f), dq1 = dq0;
        if (e >= e_first && e < e_last) {
            const float h = hs[e];
            const float rho = ms[e] / ((h * h) * h);          // vertex_depth_with_cut: m / pow(h, 3)
            if (rho > density_cut) {
                const Proj p = project(xs[e], ys[e], zs[e], h, a.cam);
                int j0, j1, k0, k1;
                pixel_range(p.px0, p.px1, a.R, j0, j1);
                pixel_range(p.py0, p.py1, a.R, k0, k1);
                if (p.keep && j1 >= j0 && k1 >= k0) {
                    const float hz = (h * a.cam.sf) * 0.5f;    // depth extent of the sphere in clip space
                    if (j1 - j0 < SURF_INLINE_SPAN && k1 - k0 < SURF_INLINE_SPAN) {
                        const float inv = 1.0f / p.wpx;
                        for (int k = k0; k <= k1; ++k)
                            for (int j = j0; j <= j1; ++j)
                                surface_fragment(a.image, a.lut, a.R, j, k, p.px0, p.py1, p.wpx, inv, qs[e], p.cz, hz);
                    } else {
                        defer = true;
                        dq0 = make_float4(p.px0, p.px1, p.py0, p.py1);
                        dq1 = make_float4(p.wpx, qs[e], p.cz, hz);
                    }
                }
            }
        }
        const unsigned dm = __ballot_sync(0xffffffffu, defer);
        if (dm) {
            unsigned qb = 0;
            if (lane == 0) qb = atomicAdd(&a.counters->q_count, (unsigned)__popc(dm));
            qb = __shfl_sync(0xffffffffu, qb, 0);
            const unsigned slot = qb + __popc(dm & lt_mask);
            if (defer && slot < a.queue_cap) {
                float4 *q = reinterpret_cast<float4 *>(a.queue + slot);
                q[0] = dq0;
                q[1] = dq1;
            }
        }
    }
}

__device__ __forceinline__ void surface_splat_record(const QueueArgs &a, const float4 q0, const float4 q1, unsigned tid,
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
        surface_fragment(a.image, a.lut, a.R, j0 + (int)dj, k0 + (int)dk, px0, py1, wpx, inv, q1.y, q1.z, q1.w);
    }
}

// deferred records: a warp per record up to COOP_MIN_WPX pixels across, a whole CTA beyond
__global__ void __launch_bounds__(256) k_queue_surface(const QueueArgs a)
{
    const unsigned count = min(*a.count, a.cap);
    if (blockIdx.x >= count) return;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned w = blockIdx.x * 8u + warp; w < count; w += gridDim.x * 8u) {
        const float4 q1 = __ldg(reinterpret_cast<const float4 *>(a.queue + w) + 1);
        if (q1.x > COOP_MIN_WPX) continue;
        surface_splat_record(a, __ldg(reinterpret_cast<const float4 *>(a.queue + w)), q1, lane, 32u);
    }
    for (unsigned w = blockIdx.x; w < count; w += gridDim.x) {
        const float4 q1 = __ldg(reinterpret_cast<const float4 *>(a.queue + w) + 1);
        if (!(q1.x > COOP_MIN_WPX)) continue;
        surface_splat_record(a, __ldg(reinterpret_cast<const float4 *>(a.queue + w)), q1, threadIdx.x, 256u);
    }
}

// ------------------------------------------------------------------------------------------------------------
// K10: bilateral filter of the depth channel (replaces shaders/smooth.wgsl, colormap/surface.py:262-297); channel 0
// passes through.  out = sum_taps s w / sum_taps w over the (2 half + 1)^2 clamp-to-edge window, w = w_spatial(dx, dy) *
// exp(-(s - centre)^2 / (2 sigma_r^2)).
//
// B200 design: a CTA produces a 32 x 8 block of outputs from a shared-memory tile of the depth channel with its halo
// ((8 + 2 half) x (32 + 2 half) floats, clamp-to-edge resolved once while staging), so every input texel is fetched from
// global memory once per CTA instead of (2 half + 1)^2 times per pixel; the spatial weights depend only on (|dx|, |dy|)
// and are tabulated once per CTA in shared memory with the reference's own operations (sqrt, square, exp), which leaves
// ONE exp and one division per tap instead of two exps, a sqrt and two divisions.  Taps are accumulated per pixel in the
// reference's order (rows outer, columns inner) with separate multiplies and adds, so the result is bit-identical to the
// direct evaluation (kept below as the fallback for windows whose tile does not fit in shared memory).
// ------------------------------------------------------------------------------------------------------------
struct BilateralArgs {
    const float *in;
    float *out;
    int width, height;
    float spatial_sigma, range_sigma;
    int kernel_size;
};

constexpr int BF_TX = 32, BF_TY = 8;

__host__ __device__ inline size_t bilateral_smem_bytes(int half)
{
    return sizeof(float) * ((size_t)(BF_TY + 2 * half) * (BF_TX + 2 * half) + (size_t)(half + 1) * (half + 1));
}

__global__ void __launch_bounds__(BF_TX * BF_TY) k_bilateral_filter(const BilateralArgs a)
{
    extern __shared__ float bf_smem[];
    const int half = a.kernel_size / 2;
    const int tw = BF_TX + 2 * half, th = BF_TY + 2 * half;
    float *tile = bf_smem, *wtab = bf_smem + (size_t)tw * th;
    const int x0 = blockIdx.x * BF_TX, y0 = blockIdx.y * BF_TY;
    const float2 *in = reinterpret_cast<const float2 *>(a.in);
    const float two_ss = 2.0f * a.spatial_sigma * a.spatial_sigma, two_rs = 2.0f * a.range_sigma * a.range_sigma;
    for (int i = threadIdx.x; i < tw * th; i += BF_TX * BF_TY) {
        const int r = i / tw, c = i - r * tw;
        const int sy = min(max(y0 - half + r, 0), a.height - 1), sx = min(max(x0 - half + c, 0), a.width - 1);
        tile[i] = __ldg(&in[(size_t)sy * a.width + sx].y);
    }
    for (int i = threadIdx.x; i < (half + 1) * (half + 1); i += BF_TX * BF_TY) {
        const int ady = i / (half + 1), adx = i - ady * (half + 1);
        const float dist = sqrtf((float)(adx * adx + ady * ady));
        wtab[i] = expf(-(dist * dist) / two_ss);
    }
    __syncthreads();
    const int tx = threadIdx.x & (BF_TX - 1), ty = threadIdx.x / BF_TX;
    const int x = x0 + tx, y = y0 + ty;
    if (x >= a.width || y >= a.height) return;
    const float centre = tile[(ty + half) * tw + tx + half];
    float vsum = 0.0f, wsum = 0.0f;
    for (int dy = -half; dy <= half; ++dy) {
        const float *row = tile + (ty + half + dy) * tw + tx + half;
        const float *wrow = wtab + abs(dy) * (half + 1);
        for (int dx = -half; dx <= half; ++dx) {
            const float s = row[dx];
            const float diff = fabsf(s - centre);
            const float w = wrow[abs(dx)] * expf(-(diff * diff) / two_rs);
            vsum += s * w;
            wsum += w;
        }
    }
    const size_t o = (size_t)y * a.width + x;
    reinterpret_cast<float2 *>(a.out)[o] = make_float2(in[o].x, vsum / wsum);
}

// direct evaluation from global memory: windows too large for a shared-memory tile (kernel_size > ~200)
__global__ void __launch_bounds__(256) k_bilateral_filter_direct(const BilateralArgs a)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= a.width || y >= a.height) return;
    const float2 *in = reinterpret_cast<const float2 *>(a.in);
    const float2 centre = in[(size_t)y * a.width + x];
    const int half = a.kernel_size / 2;
    const float two_ss = 2.0f * a.spatial_sigma * a.spatial_sigma, two_rs = 2.0f * a.range_sigma * a.range_sigma;
    float vsum = 0.0f, wsum = 0.0f;
    for (int dy = -half; dy <= half; ++dy) {
        const int sy = min(max(y + dy, 0), a.height - 1);
        for (int dx = -half; dx <= half; ++dx) {
            const int sx = min(max(x + dx, 0), a.width - 1);
            const float s = __ldg(&in[(size_t)sy * a.width + sx].y);
            const float dist = sqrtf((float)(dx * dx + dy * dy));
            const float diff = fabsf(s - centre.y);
            const float w = expf(-(dist * dist) / two_ss) * expf(-(diff * diff) / two_rs);
            vsum += s * w;
            wsum += w;
        }
    }
    reinterpret_cast<float2 *>(a.out)[(size_t)y * a.width + x] = make_float2(centre.x, vsum / wsum);
}

// ------------------------------------------------------------------------------------------------------------
// K11: surface lighting (surface.wgsl:24-123)
// ------------------------------------------------------------------------------------------------------------
struct ShadeArgs {
    const float *image;          // (res, res, 2): material value, smoothed depth
    int res;
    tsplat_surface_params p;
    const float *lut;            // 1-D colormap (device RGBA float), used when p.material_colormap
    int lut_w;
    void *out;
    int out_w, out_h, out_fmt;
};

// textureSample(colorTexture, textureSampler, (u, v)): mag linear / min nearest, clamp-to-edge
__device__ __forceinline__ float2 sample_rg(const float *__restrict__ img, int res, float u, float v, bool linear)
{
    const float2 *im = reinterpret_cast<const float2 *>(img);
    if (!linear) {
        const int x = min(max((int)floorf(u * res), 0), res - 1), y = min(max((int)floorf(v * res), 0), res - 1);
        return im[(size_t)y * res + x];
    }
    const float px = u * res - 0.5f, py = v * res - 0.5f;
    const float ix = floorf(px), iy = floorf(py);
    const float fx = px - ix, fy = py - iy;
    const int x0 = min(max((int)ix, 0), res - 1), x1 = min(max((int)ix + 1, 0), res - 1);
    const int y0 = min(max((int)iy, 0), res - 1), y1 = min(max((int)iy + 1, 0), res - 1);
    const float2 t00 = im[(size_t)y0 * res + x0], t01 = im[(size_t)y0 * res + x1];
    const float2 t10 = im[(size_t)y1 * res + x0], t11 = im[(size_t)y1 * res + x1];
    const float2 top = make_float2(t00.x + (t01.x - t00.x) * fx, t00.y + (t01.y - t00.y) * fx);
    const float2 bot = make_float2(t10.x + (t11.x - t10.x) * fx, t10.y + (t11.y - t10.y) * fx);
    return make_float2(top.x + (bot.x - top.x) * fy, top.y + (bot.y - top.y) * fy);
}

}  // namespace tsplat_surface
