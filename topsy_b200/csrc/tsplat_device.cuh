// Device-side arithmetic contract of the SPH splat (shared by every kernel in tsplat.cu).
//
// Restates, for one particle and one pixel centre, what the reference's vertex shader, the fixed-function
// rasteriser and the fragment shader compute:
//   src/topsy/shaders/sph.wgsl:54-66   vertex_calculate_positions  (clip = M p, half-size 2 h / scale)
//   src/topsy/shaders/sph.wgsl:68-91   vertex_rgb / vertex_weighting / vertex_depth   (per-particle intensities)
//   src/topsy/shaders/sph.wgsl:138-165 fragment_weighting / fragment_rgb              (K * intensities)
//   src/topsy/sph.py:409-426           kernel texture: 4 mip levels, mag linear / min+mip nearest, clamp-to-edge
// The same operations, in the same order and with the same FMA placement, are in oracle/splat_oracle.c and
// oracle/topsy_oracle.py (the checkers).  This translation unit is compiled with -fmad=false so that ONLY the
// fmaf() calls below fuse; divisions are IEEE (-prec-div=true, the nvcc default; never --use_fast_math).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tsplat {

constexpr int LUT_TOTAL = 5440;                      // 64^2 + 32^2 + 16^2 + 8^2
__host__ __device__ constexpr int lut_offset(int level) { return level == 0 ? 0 : level == 1 ? 4096 : level == 2 ? 5120 : 5376; }

constexpr float LEVEL_T0 = 45.254834f;               // 64/sqrt(2): wpx above -> level 0 nearest; wpx >= 64 -> bilinear
constexpr float LEVEL_T1 = 22.627417f;
constexpr float LEVEL_T2 = 11.313708f;               // wpx <= this -> level 3 (8x8)

struct Camera {
    float m[12];        // rows 0..2 of the row-major 4x4 transform
    float sf;           // 1/scale
    float R;            // render resolution as float
    float halfR;        // 0.5 * R (exact)
};

struct Proj {
    float cz;           // clip-space depth in [0,1] when kept
    float px0, px1;     // left / right quad edge in pixel units
    float py0, py1;     // top / bottom quad edge in pixel-row units (row 0 = +y)
    float wpx;          // quad width in pixels = a * R
    bool keep;
};

__device__ __forceinline__ Proj project(float x, float y, float z, float h, const Camera &c)
{
    Proj p;
    const float cx = fmaf(c.m[0], x, fmaf(c.m[1], y, fmaf(c.m[2], z, c.m[3])));
    const float cy = fmaf(c.m[4], x, fmaf(c.m[5], y, fmaf(c.m[6], z, c.m[7])));
    p.cz = fmaf(c.m[8], x, fmaf(c.m[9], y, fmaf(c.m[10], z, c.m[11])));
    const float a = (c.sf * h) * 2.0f;
    const float x0 = cx - a, x1 = cx + a, y0 = cy - a, y1 = cy + a;
    p.px0 = fmaf(x0, c.halfR, c.halfR);
    p.px1 = fmaf(x1, c.halfR, c.halfR);
    p.py0 = fmaf(-y1, c.halfR, c.halfR);
    p.py1 = fmaf(-y0, c.halfR, c.halfR);
    p.wpx = a * c.R;
    // fixed-function clip: the whole quad shares one z, so it is kept iff 0 <= z <= 1.
    // a > 0 && finite: zero-size quads cover nothing; NaN / negative smoothing lengths are dropped (documented).
    p.keep = (p.cz >= 0.0f) && (p.cz <= 1.0f) && (a > 0.0f) && (a <= 3.4028234664e38f);
    return p;
}

// Integer pixel ranges [lo, hi] such that lo <= j <= hi  <=>  (j + 0.5f >= e0) && (j + 0.5f < e1), clipped to
// [0, R-1] (empty ranges come out as lo > hi).  The ceil form is equivalent to the comparison form for every j inside
// the image (tests/test_oracle_golden.py::test_ceil_bounds_equal_comparisons checks the fp32 corner cases on the CPU).
// cvt.rpi.s32.f32 saturates and maps NaN to 0, so no float clamps are needed.
__device__ __forceinline__ void pixel_range(float e0, float e1, int R, int &lo, int &hi)
{
    lo = min(max(__float2int_ru(e0 - 0.5f), 0), R);
    hi = min(max(__float2int_ru(e1 - 0.5f), 0), R) - 1;
}

// Kernel value for a quad of width wpx (inv = 1/wpx) at pixel centre (fx, fy).  `lut` holds all four levels.
__device__ __forceinline__ float sample_lut(const float *__restrict__ lut, float wpx, float inv, float px0, float py1,
                                            float fx, float fy)
{
    const float u = (fx - px0) * inv;
    const float v = (py1 - fy) * inv;
    if (wpx >= 64.0f) {
        const float tu = fmaf(u, 64.0f, -0.5f), tv = fmaf(v, 64.0f, -0.5f);
        const float iu = floorf(tu), iv = floorf(tv);
        const float fu = tu - iu, fv = tv - iv;
        const int a = (int)iu, b = (int)iv;
        const int a0 = min(max(a, 0), 63), a1 = min(max(a + 1, 0), 63);
        const int b0 = min(max(b, 0), 63), b1 = min(max(b + 1, 0), 63);
        const float t00 = lut[b0 * 64 + a0], t01 = lut[b0 * 64 + a1];
        const float t10 = lut[b1 * 64 + a0], t11 = lut[b1 * 64 + a1];
        const float top = fmaf(fu, t01 - t00, t00);
        const float bot = fmaf(fu, t11 - t10, t10);
        return fmaf(fv, bot - top, top);
    }
    const int level = wpx > LEVEL_T0 ? 0 : wpx > LEVEL_T1 ? 1 : wpx > LEVEL_T2 ? 2 : 3;
    const int n = 64 >> level;
    const int iu = min(max(__float2int_rd(u * (float)n), 0), n - 1);
    const int iv = min(max(__float2int_rd(v * (float)n), 0), n - 1);
    return lut[lut_offset(level) + iv * n + iu];
}

// Level-3-only variant (wpx <= LEVEL_T2): `lut8` points at the 8x8 table.
__device__ __forceinline__ float sample_lut8(const float *__restrict__ lut8, float inv, float px0, float py1,
                                             float fx, float fy)
{
    const float u = (fx - px0) * inv;
    const float v = (py1 - fy) * inv;
    const int iu = min(max(__float2int_rd(u * 8.0f), 0), 7);
    const int iv = min(max(__float2int_rd(v * 8.0f), 0), 7);
    return lut8[iv * 8 + iu];
}

// A deferred (large-footprint) particle, written by the project kernel and consumed by the tile / cooperative
// kernels.  32 bytes = one DRAM sector.
struct __align__(16) Deferred {
    float px0, px1, py0, py1;
    float wpx, v0, v1, v2;      // v*: DENSITY m/h^2 | WEIGHTED m/h^2, q | RGB r/h^2, g/h^2, b/h^2 | DEPTH m/h^2, cz
};

}  // namespace tsplat
