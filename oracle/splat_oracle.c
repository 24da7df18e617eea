/*
 * CPU oracle for topsy's SPH splat -- TEST INFRASTRUCTURE ONLY (never linked or called by topsy_b200).
 *
 * Plain-C restatement of the reference's vertex + rasteriser + fragment + additive-blend path:
 *   src/topsy/shaders/sph.wgsl:54-91 (vertex_calculate_positions / vertex_weighting / vertex_rgb / vertex_depth)
 *   src/topsy/shaders/sph.wgsl:138-165 (fragment_weighting / fragment_rgb)
 *   src/topsy/sph.py:31-42 (ONE/ONE additive blend), :268-299 (transform), :409-426 (LUT mip chain + sampler)
 * Sampling / coverage rules that the reference inherits from the WebGPU defaults are written out in
 * SURVEY.md section 8 rows a12/a13 and DESIGN.md section 3.
 *
 * The arithmetic contract (fp32 geometry with explicit FMAs, comparisons on pixel centres) is shared
 * bit-for-bit with oracle/topsy_oracle.py and topsy_b200/csrc/tsplat_device.cuh.  Build with
 * -ffp-contract=off so that only the fmaf() calls below fuse.
 *
 * Parity status: PINNED through tests/test_oracle_golden.py (this file == numpy oracle == reference goldens).
 *
 * Two accumulation flavours:
 *   oracle_splat_f64  : fp64 accumulators  -> the checker
 *   oracle_splat_f32  : fp32 accumulators  -> the timed CPU baseline (the reference blends in fp32)
 * Both use thread-private images followed by a parallel sum (the strategy BASELINE.md section 3 prescribes).
 * The private images live in a workspace that is allocated once and reused by later calls (oracle_release_workspace()
 * frees it): a call pays one parallel clear and one parallel sum of the T images, not T page-faulting allocations, so
 * the timed CPU baseline does not depend on how few particles a step holds.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* persistent thread-private accumulation images (one per OpenMP thread beyond thread 0, which uses the output) */
#define ORACLE_MAX_THREADS 1024
static void *g_priv[ORACLE_MAX_THREADS];
static size_t g_priv_bytes[ORACLE_MAX_THREADS];

static void *workspace_get(int t, size_t bytes)
{
    if (t < 0 || t >= ORACLE_MAX_THREADS) return NULL;
    if (g_priv_bytes[t] < bytes) {
        free(g_priv[t]);
        g_priv[t] = malloc(bytes);          /* first touched (cleared) by the owning thread: NUMA-local pages */
        g_priv_bytes[t] = g_priv[t] ? bytes : 0;
    }
    return g_priv[t];
}

void oracle_release_workspace(void)
{
    for (int t = 0; t < ORACLE_MAX_THREADS; ++t) { free(g_priv[t]); g_priv[t] = NULL; g_priv_bytes[t] = 0; }
}

enum { MODE_DENSITY = 0, MODE_WEIGHTED = 1, MODE_RGB = 2, MODE_DEPTH = 3 };
static const int MODE_CHANNELS[4] = {1, 2, 4, 2};
static const int LUT_OFF[4] = {0, 4096, 5120, 5376};

#define LEVEL_T0 45.254834f
#define LEVEL_T1 22.627417f
#define LEVEL_T2 11.313708f

typedef struct { float cz, px0, px1, py0, py1, wpx, inv; int keep; } proj_t;

static inline proj_t project(float x, float y, float z, float h, const float *M, float sf, float R)
{
    proj_t p;
    float cx = fmaf(M[0], x, fmaf(M[1], y, fmaf(M[2], z, M[3])));
    float cy = fmaf(M[4], x, fmaf(M[5], y, fmaf(M[6], z, M[7])));
    p.cz = fmaf(M[8], x, fmaf(M[9], y, fmaf(M[10], z, M[11])));
    float a = (sf * h) * 2.0f;
    float half = 0.5f * R;
    float x0 = cx - a, x1 = cx + a, y0 = cy - a, y1 = cy + a;
    p.px0 = fmaf(x0, half, half);
    p.px1 = fmaf(x1, half, half);
    p.py0 = fmaf(-y1, half, half);
    p.py1 = fmaf(-y0, half, half);
    p.wpx = a * R;
    p.inv = 1.0f / p.wpx;
    p.keep = (p.cz >= 0.0f) && (p.cz <= 1.0f) && (a > 0.0f) && isfinite(a);
    return p;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static inline float sample_lut(const float *lut, const proj_t *p, float fx, float fy)
{
    float u = (fx - p->px0) * p->inv;
    float v = (p->py1 - fy) * p->inv;
    if (p->wpx >= 64.0f) {
        float tu = fmaf(u, 64.0f, -0.5f), tv = fmaf(v, 64.0f, -0.5f);
        float iu = floorf(tu), iv = floorf(tv);
        float fu = tu - iu, fv = tv - iv;
        int a0 = clampi((int)iu, 0, 63), a1 = clampi((int)iu + 1, 0, 63);
        int b0 = clampi((int)iv, 0, 63), b1 = clampi((int)iv + 1, 0, 63);
        float t00 = lut[b0 * 64 + a0], t01 = lut[b0 * 64 + a1], t10 = lut[b1 * 64 + a0], t11 = lut[b1 * 64 + a1];
        float top = fmaf(fu, t01 - t00, t00);
        float bot = fmaf(fu, t11 - t10, t10);
        return fmaf(fv, bot - top, top);
    }
    int level = p->wpx > LEVEL_T0 ? 0 : p->wpx > LEVEL_T1 ? 1 : p->wpx > LEVEL_T2 ? 2 : 3;
    int n = 64 >> level;
    int iu = clampi((int)floorf(u * (float)n), 0, n - 1);
    int iv = clampi((int)floorf(v * (float)n), 0, n - 1);
    return lut[LUT_OFF[level] + iv * n + iu];
}

/* conservative integer bounds; exact membership is decided by the fp32 comparisons inside the loop */
static inline void bounds(float lo, float hi, int R, int *a, int *b)
{
    float l = floorf(lo) - 1.0f, h = ceilf(hi) + 1.0f;
    if (!(l > 0.0f)) l = 0.0f;
    if (!(h < (float)(R - 1))) h = (float)(R - 1);
    if (l > (float)(R - 1)) { *a = 1; *b = 0; return; }
    if (h < 0.0f) { *a = 1; *b = 0; return; }
    *a = (int)l; *b = (int)h;
}

#define SPLAT_BODY(ACC_T)                                                                                      \
    const int C = MODE_CHANNELS[mode];                                                                         \
    const float Rf = (float)R;                                                                                 \
    proj_t p = project(x[i], y[i], z[i], h[i], M, sf, Rf);                                                     \
    if (!p.keep) continue;                                                                                     \
    int j0, j1, k0, k1;                                                                                        \
    bounds(p.px0, p.px1, R, &j0, &j1);                                                                         \
    bounds(p.py0, p.py1, R, &k0, &k1);                                                                         \
    if (j1 < j0 || k1 < k0) continue;                                                                          \
    float rhh = 1.0f / (h[i] * h[i]);          /* one IEEE reciprocal, then multiplies (arithmetic contract) */ \
    float i0 = w0[i] * rhh, i1 = 0.0f, i2 = 0.0f;                                                              \
    if (mode == MODE_RGB) { i1 = w1[i] * rhh; i2 = w2[i] * rhh; }                                              \
    for (int k = k0; k <= k1; ++k) {                                                                           \
        float fy = (float)k + 0.5f;                                                                            \
        if (!(fy >= p.py0 && fy < p.py1)) continue;                                                            \
        for (int j = j0; j <= j1; ++j) {                                                                       \
            float fx = (float)j + 0.5f;                                                                        \
            if (!(fx >= p.px0 && fx < p.px1)) continue;                                                        \
            float K = sample_lut(lut, &p, fx, fy);                                                             \
            ACC_T *px = mine + ((size_t)k * R + j) * C;                                                        \
            if (mode == MODE_RGB) {                                                                            \
                px[0] += (ACC_T)(i0 * K); px[1] += (ACC_T)(i1 * K); px[2] += (ACC_T)(i2 * K); px[3] += (ACC_T)1; \
            } else {                                                                                           \
                float v0 = K * i0;                                                                             \
                px[0] += (ACC_T)v0;                                                                            \
                if (mode == MODE_WEIGHTED) px[1] += (ACC_T)(v0 * w1[i]);                                       \
                else if (mode == MODE_DEPTH) px[1] += (ACC_T)(v0 * p.cz);                                      \
            }                                                                                                  \
        }                                                                                                      \
    }

#define DEFINE_SPLAT(NAME, ACC_T)                                                                              \
int NAME(const float *x, const float *y, const float *z, const float *h,                                      \
         const float *w0, const float *w1, const float *w2, int64_t n,                                         \
         const int64_t *starts, const int64_t *lens, int nranges,                                              \
         const float *M, float sf, const float *lut, int R, int mode, ACC_T *img, int clear, int nthreads)     \
{                                                                                                              \
    if (mode < 0 || mode > 3 || R <= 0) return -1;                                                             \
    const size_t npx = (size_t)R * R * MODE_CHANNELS[mode];                                                    \
    int64_t one_start = 0, one_len = n;                                                                        \
    if (nranges <= 0 || !starts) { starts = &one_start; lens = &one_len; nranges = 1; }                        \
    for (int r = 0; r < nranges; ++r) if (starts[r] < 0 || starts[r] + lens[r] > n) return -2;                 \
    if (clear) memset(img, 0, npx * sizeof(ACC_T));                                                            \
    int nt = nthreads;                                                                                         \
    if (nt <= 0) {                                                                                             \
        nt = 1;                                                                                                \
        _Pragma("omp parallel") { _Pragma("omp single") nt = omp_get_num_threads(); }                         \
    }                                                                                                          \
    if (nt > ORACLE_MAX_THREADS) nt = ORACLE_MAX_THREADS;                                                      \
    ACC_T **priv = (ACC_T **)calloc(nt, sizeof(ACC_T *));                                                      \
    if (!priv) return -3;                                                                                      \
    int fail = 0;                                                                                              \
    _Pragma("omp parallel num_threads(nt)")                                                                    \
    {                                                                                                          \
        int t = omp_get_thread_num();                                                                          \
        ACC_T *mine = (t == 0) ? img : (ACC_T *)workspace_get(t, npx * sizeof(ACC_T));                         \
        priv[t] = mine;                                                                                        \
        if (!mine) { _Pragma("omp atomic write") fail = 1; }                                                   \
        else if (t != 0) memset(mine, 0, npx * sizeof(ACC_T));                                                 \
        _Pragma("omp barrier")                                                                                 \
        if (!fail) {                                                                                           \
            for (int r = 0; r < nranges; ++r) {                                                                \
                const int64_t s = starts[r], e = starts[r] + lens[r];                                          \
                _Pragma("omp for schedule(dynamic, 4096) nowait")                                              \
                for (int64_t i = s; i < e; ++i) { SPLAT_BODY(ACC_T) }                                          \
            }                                                                                                  \
        }                                                                                                      \
        _Pragma("omp barrier")                                                                                 \
        if (!fail) {                                                                                           \
            _Pragma("omp for schedule(static)")                                                                \
            for (size_t q = 0; q < npx; ++q) {                                                                 \
                ACC_T s = img[q];                                                                              \
                for (int tt = 1; tt < nt; ++tt) s += priv[tt][q];                                              \
                img[q] = s;                                                                                    \
            }                                                                                                  \
        }                                                                                                      \
    }                                                                                                          \
    free(priv);                                                                                                \
    return fail ? -3 : 0;                                                                                      \
}

DEFINE_SPLAT(oracle_splat_f64, double)
DEFINE_SPLAT(oracle_splat_f32, float)


/*
 * Surface render mode: z-buffered splat of the front-most particles above a density cut.
 *   src/topsy/shaders/sph.wgsl:93-120   vertex_depth_with_cut: rho = m / h^3 > density_cut, intensities = (q, clip z, h*sf*0.5)
 *   src/topsy/shaders/sph.wgsl:148-158  fragment_raw: K = textureSample(local-sphere kernel); discard if K < 0;
 *                                       depth = z + (h*sf*0.5)*K; output (q, depth), frag_depth = depth
 *   src/topsy/sph.py:455-470, :592-599  blend = replace, depth_compare = greater, depth cleared to 0
 * `lut` is the local-sphere mip chain (sph.py:446-455: sqrt(4 - d^2) inside the sphere, -0.01 outside; no normalisation).
 * img is (R, R, 2) float32 = (q, depth); 0 where nothing was drawn.
 *
 * clamp_depth != 0: the reference's depth test on frag_depth clamped to the viewport range [0, 1] (WebGPU), first
 *                   drawn fragment wins ties, `zbuf` (R*R floats, caller-provided) holds the depth attachment.
 * clamp_depth == 0: what the CUDA path does: per pixel the fragment with the largest (unclamped) depth wins, ties go
 *                   to the larger bit pattern of q -- order independent, so it can be compared bit-for-bit.  The two
 *                   differ only where a pixel receives more than one fragment deeper than 1.0.
 * Arithmetic contract: rho = m / ((h*h)*h) with fp32 multiplies (WGSL's pow(h, 3.0) is only accurate to a few ULP and
 * implementation dependent), depth = z + hz*K with separate multiply and add.
 */
int oracle_splat_surface(const float *x, const float *y, const float *z, const float *h, const float *m, const float *q,
                         int64_t n, const int64_t *starts, const int64_t *lens, int nranges, const float *M, float sf,
                         const float *lut, int R, float density_cut, float *img, float *zbuf, int clear, int clamp_depth)
{
    if (R <= 0 || (clamp_depth && !zbuf)) return -1;
    int64_t one_start = 0, one_len = n;
    if (nranges <= 0 || !starts) { starts = &one_start; lens = &one_len; nranges = 1; }
    for (int r = 0; r < nranges; ++r) if (starts[r] < 0 || starts[r] + lens[r] > n) return -2;
    if (clear) {
        memset(img, 0, (size_t)R * R * 2 * sizeof(float));
        if (zbuf) memset(zbuf, 0, (size_t)R * R * sizeof(float));
    }
    const float Rf = (float)R;
    for (int r = 0; r < nranges; ++r) {
        for (int64_t i = starts[r]; i < starts[r] + lens[r]; ++i) {
            const float rho = m[i] / ((h[i] * h[i]) * h[i]);
            if (!(rho > density_cut)) continue;
            proj_t p = project(x[i], y[i], z[i], h[i], M, sf, Rf);
            if (!p.keep) continue;
            int j0, j1, k0, k1;
            bounds(p.px0, p.px1, R, &j0, &j1);
            bounds(p.py0, p.py1, R, &k0, &k1);
            if (j1 < j0 || k1 < k0) continue;
            const float hz = (h[i] * sf) * 0.5f;
            for (int k = k0; k <= k1; ++k) {
                float fy = (float)k + 0.5f;
                if (!(fy >= p.py0 && fy < p.py1)) continue;
                for (int j = j0; j <= j1; ++j) {
                    float fx = (float)j + 0.5f;
                    if (!(fx >= p.px0 && fx < p.px1)) continue;
                    float K = sample_lut(lut, &p, fx, fy);
                    if (K < 0.0f) continue;
                    float t = hz * K;
                    float depth = p.cz + t;
                    float *px = img + ((size_t)k * R + j) * 2;
                    if (clamp_depth) {
                        float d = depth < 0.0f ? 0.0f : (depth > 1.0f ? 1.0f : depth);
                        if (d > zbuf[(size_t)k * R + j]) { zbuf[(size_t)k * R + j] = d; px[0] = q[i]; px[1] = depth; }
                    } else {
                        uint32_t db, qb, odb, oqb;
                        memcpy(&db, &depth, 4); memcpy(&qb, &q[i], 4); memcpy(&odb, &px[1], 4); memcpy(&oqb, &px[0], 4);
                        uint64_t key = ((uint64_t)db << 32) | qb, old = ((uint64_t)odb << 32) | oqb;
                        if (depth > 0.0f && key > old) { px[0] = q[i]; px[1] = depth; }
                    }
                }
            }
        }
    }
    return 0;
}

/* Number of (particle, pixel) updates and culled particles: the "work" term of the second roofline. */
int oracle_count_updates(const float *x, const float *y, const float *z, const float *h, int64_t n,
                         const float *M, float sf, int R, int64_t *n_updates, int64_t *n_culled)
{
    int64_t upd = 0, cul = 0;
    const float Rf = (float)R;
#pragma omp parallel for reduction(+ : upd, cul) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        proj_t p = project(x[i], y[i], z[i], h[i], M, sf, Rf);
        if (!p.keep) { cul++; continue; }
        int j0, j1, k0, k1;
        bounds(p.px0, p.px1, R, &j0, &j1);
        bounds(p.py0, p.py1, R, &k0, &k1);
        if (j1 < j0 || k1 < k0) continue;
        int64_t nj = 0, nk = 0;
        for (int j = j0; j <= j1; ++j) { float fx = (float)j + 0.5f; nj += (fx >= p.px0 && fx < p.px1); }
        for (int k = k0; k <= k1; ++k) { float fy = (float)k + 0.5f; nk += (fy >= p.py0 && fy < p.py1); }
        upd += nj * nk;
    }
    *n_updates = upd; *n_culled = cul;
    return 0;
}

int oracle_num_threads(void)
{
    int nt = 1;
#pragma omp parallel
    {
#pragma omp single
        nt = omp_get_num_threads();
    }
    return nt;
}
