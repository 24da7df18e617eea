/*
 * tsplat -- C ABI of the B200-native SPH projection path (drop-in for topsy's wgpu render passes).
 *
 * The reference (pynbody/topsy 0.8.1) has no FFI for this path: its boundary is the Python classes
 * SPH / RGBSPH / DepthSPH (src/topsy/sph.py), ParticleBuffers (src/topsy/particle_buffers.py) and the colormap
 * implementations (src/topsy/colormap/implementation.py), which drive wgpu.  Each entry point below names the wgpu
 * call sequence it replaces; INTEGRATION.md shows the ctypes stub a topsy maintainer would add.
 *
 * Conventions: extern "C", plain pointers and sizes, no exceptions.  Every function returns 0 on success or a
 * negative tsplat_status; tsplat_last_error() returns a thread-local message.  All device buffers (particles,
 * image, scratch, colormap LUT, output) are owned by the caller (PyTorch tensors in the Python host layer) and only
 * borrowed; the context owns a few KB of constants/counters.  A context is bound to one device and is not
 * thread-safe; different contexts may be used from different threads.  `stream` is a cudaStream_t (NULL = legacy
 * default stream); all work is enqueued asynchronously on it.
 */
#ifndef TSPLAT_H
#define TSPLAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSPLAT_ABI_VERSION 1

typedef struct tsplat_ctx tsplat_ctx;

typedef enum {
    TSPLAT_OK = 0,
    TSPLAT_ERR_INVALID = -1,   /* bad argument (Python layer raises ValueError)   */
    TSPLAT_ERR_STATE = -2,     /* call order violated (Python layer raises RuntimeError) */
    TSPLAT_ERR_CUDA = -3,      /* CUDA runtime failure, message has the cudaError string */
    TSPLAT_ERR_NOMEM = -4
} tsplat_status;

/* Accumulation modes = the reference's vertex/fragment entry-point pairs (src/topsy/shaders/sph.wgsl). */
typedef enum {
    TSPLAT_MODE_DENSITY = 0,   /* vertex_weighting/fragment_weighting with quantity == 0: 1 channel  K*m/h^2      */
    TSPLAT_MODE_WEIGHTED = 1,  /* vertex_weighting/fragment_weighting (sph.wgsl:75-83,138-146): (K m/h^2, K m/h^2 q) */
    TSPLAT_MODE_RGB = 2,       /* vertex_rgb/fragment_rgb (sph.wgsl:68-73,160-165): (K r/h^2, K g/h^2, K b/h^2, 1)   */
    TSPLAT_MODE_DEPTH = 3,     /* vertex_depth/fragment_weighting (sph.wgsl:85-91): (K m/h^2, K m/h^2 z_clip)        */
    TSPLAT_MODE_SURFACE = 4    /* vertex_depth_with_cut/fragment_raw (sph.wgsl:93-158) with depth_compare = greater and
                                  blending off (sph.py:457-601): z-buffered (quantity, depth) of the front-most particle
                                  above the density cut; w0 = mass, w1 = quantity; needs tsplat_set_surface()          */
} tsplat_mode;

/* channels per pixel for a mode: 1, 2, 4, 2, 2 */
int tsplat_mode_channels(int mode);

typedef enum {
    TSPLAT_FMT_RGBA8 = 0,      /* rgba8unorm canvas (visualizer.py:170-181)  */
    TSPLAT_FMT_RGBA16F = 1,    /* rgba16float canvas of the 'rgb-hdr' mode   */
    TSPLAT_FMT_RGBA32F = 2     /* Colormap.sph_raw_output_to_image float path (implementation.py:143-148) */
} tsplat_format;

typedef enum {
    TSPLAT_CMAP_DENSITY = 0,        /* fragment_main, DENSITY        (colormap.wgsl:109-127) */
    TSPLAT_CMAP_WEIGHTED = 1,       /* fragment_main, WEIGHTED_MEAN                           */
    TSPLAT_CMAP_BIVARIATE = 2,      /* fragment_main, BIVARIATE + DENSITY  (colormap.wgsl:84-107) */
    TSPLAT_CMAP_BIVARIATE_WEIGHTED = 3,
    TSPLAT_CMAP_RGB = 4             /* fragment_main_tri (colormap.wgsl:131-159) */
} tsplat_cmap_kind;

/* The uniform block of colormap.wgsl:1-8 (already shifted for mass_scale by the host, implementation.py:427-453),
 * plus the preprocessor switches the reference bakes into the shader source. */
typedef struct {
    float vmin, vmax;
    float density_vmin, density_vmax;
    float window_aspect_ratio;      /* width/height of the output; 1 for get_sph_presentation_image */
    float gamma;
    int32_t kind;                   /* tsplat_cmap_kind */
    int32_t log_scale;              /* LOG_SCALE */
} tsplat_colormap_params;

typedef struct {
    int64_t particles_submitted;    /* particles inside the submitted ranges since the last clear           */
    int64_t particles_culled;       /* dropped by the z clip (0 <= z_clip <= 1) or non-positive smoothing     */
    int64_t particles_direct;       /* splatted by the per-particle atomic path                                */
    int64_t particles_tiled;        /* deferred to the tile-gather path                                        */
    int64_t particles_huge;         /* deferred to the cooperative atomic path (footprint or pair overflow)    */
    int64_t tile_pairs;             /* (particle, tile) pairs processed by the gather path                     */
    int64_t kernel_launches;        /* CUDA kernels launched by this context since creation                    */
    int64_t direct_vector_reds;     /* 128-bit (or narrower) REDs issued by the direct path since the last clear */
} tsplat_stats;

/* Live timing of the dominant kernel (K1, the project + direct-splat pass): with timing enabled every tsplat_render brackets
 * its K1 launch with CUDA events on the caller's stream (a ring of TSPLAT_TIMING_SLOTS pairs); tsplat_kernel_timing
 * synchronises those events and returns the number of K1 launches recorded since the last query and their summed
 * duration.  For bench.py's roofline (achieved = algorithmic bytes per launch / this duration); off by default. */
#define TSPLAT_TIMING_SLOTS 256
int tsplat_enable_kernel_timing(tsplat_ctx *ctx, int enable);
int tsplat_kernel_timing(tsplat_ctx *ctx, int64_t *n_launches, double *total_ms);

const char *tsplat_last_error(void);
int tsplat_abi_version(void);

/* Replaces: device.create_texture(rg32float/rgba32float render target) + pipeline/bind-group setup
 * (sph.py:50-88,156-259).  `resolution` is the square render resolution R. */
int tsplat_create(int device_ordinal, int resolution, tsplat_ctx **out);
int tsplat_destroy(tsplat_ctx *ctx);

/* Replaces SPH._setup_kernel_texture (sph.py:396-426): 5440 host floats = levels 64^2,32^2,16^2,8^2 row-major. */
int tsplat_set_kernel_lut(tsplat_ctx *ctx, const float *host_lut, int n_floats);

/* Replaces SPH._update_transform_buffer (sph.py:262-299).  M is the ROW-major 4x4 float32 matrix
 * clipdisp @ (R/scale (+) 1) @ translate(offset), i.e. the transpose of what the reference uploads for WGSL;
 * scale_factor = 1/scale. */
int tsplat_set_camera(tsplat_ctx *ctx, const float *M16, float scale_factor);

/* Replaces set_vertex_buffer(0, pos_smooth) (particle_buffers.py:62-68) -- device SoA float32, borrowed. */
int tsplat_set_particles(tsplat_ctx *ctx, const float *x, const float *y, const float *z, const float *h, int64_t n);

/* Replaces set_vertex_buffer(1, mass_and_quantity | rgb).  DENSITY/DEPTH: w0=m.  WEIGHTED: w0=m,w1=q.
 * RGB: w0,w1,w2 = r,g,b.  Unused pointers may be NULL. */
int tsplat_set_weights(tsplat_ctx *ctx, const float *w0, const float *w1, const float *w2);

/* Render target: caller-owned device buffer of R*R*channels float32, pixel-interleaved, row 0 = top (+y). */
int tsplat_set_image(tsplat_ctx *ctx, float *image, int channels);

/* Caller-owned device scratch for the deferred (large-footprint) particle queue and tile bins.
 * tsplat_scratch_bytes() gives the size needed to defer up to max_particles_per_call particles. */
int64_t tsplat_scratch_bytes(int resolution, int64_t max_particles_per_call);
int tsplat_set_scratch(tsplat_ctx *ctx, void *scratch, int64_t bytes);

/* Replaces ParticleBuffers.update_particle_ranges + multi_draw_indirect + queue.submit of one render pass
 * (particle_buffers.py:70-82, sph.py:318-326,337-362).  starts/lens are HOST arrays of n_ranges global particle
 * ranges inside the current particle buffers (n_ranges == 0 -> everything).  clear != 0 == LoadOp.clear. */
int tsplat_render(tsplat_ctx *ctx, const int64_t *starts, const int64_t *lens, int n_ranges, int mode, int clear,
                  void *stream);

/* Replaces ColormapBase.encode_render_pass (implementation.py:352-367, colormap.wgsl).  lut: device float32 RGBA,
 * lut_w x lut_h texels (lut_h == 1 for the 1-D maps; ignored for TSPLAT_CMAP_RGB).  `image`/`channels` is the
 * accumulation image to present (pass the buffer given to tsplat_set_image, or any other R_in x R_in image).
 * out: device buffer out_w*out_h*4 elements of out_fmt. */
int tsplat_colormap(tsplat_ctx *ctx, const float *image, int image_res, int channels,
                    const tsplat_colormap_params *params, const float *lut, int lut_w, int lut_h,
                    void *out, int out_w, int out_h, int out_fmt, void *stream);

/* Multi-GPU presentation (no reference counterpart: topsy is single-device).  Sums rows [row0, row0+nrows) of the
 * n_peers partial accumulation images (device pointers, typically NVLink peer mappings of every rank's image, passed
 * as a HOST array), in rank order; optionally stores the fp32 sum into sum_out (full-image base pointer, may be a peer
 * mapping) and/or the colormapped RGBA into out (full-image base pointer of out_fmt, may be a peer mapping).
 * Fuses reduce-scatter + colormap + gather into one kernel over peer memory. */
int tsplat_reduce_colormap(tsplat_ctx *ctx, const float *const *peer_images, int n_peers, int channels,
                           int row0, int nrows, const tsplat_colormap_params *params, const float *lut,
                           int lut_w, int lut_h, void *out, int out_fmt, float *sum_out, void *stream);

/* Multi-GPU image all-reduce for the drop-in classes (no reference counterpart: topsy is single-device; the wiring it
 * extends is visualizer.py:75-80 loader -> ParticleBuffers and sph.py:306-332 render).  Every rank holds a partial
 * accumulation image of its particle stripe; this call reduces rows [row0, row0+nrows) over the n_peers partial images
 * (NVLink peer mappings, HOST array of device pointers) and stores the result into the same rows of EVERY peer's
 * reduced image (peer_out, HOST array of device pointers): reduce-scatter + all-gather over peer memory in one kernel.
 *   op == TSPLAT_REDUCE_SUM  out = sum_r (*peer_scale[r]) * peer_images[r]   -- peer_scale[r] is a DEVICE pointer to rank
 *                            r's mass scale (RenderProgression.end_frame_get_scalefactor) in peer memory, so ranks whose
 *                            progressive frames covered different fractions of their stripes combine correctly;
 *   op == TSPLAT_REDUCE_ZMAX surface mode: per pixel the 64-bit (depth, quantity) key maximum (channels must be 2).
 * The caller brackets the call with barriers (all partial images complete before, all slabs stored after). */
enum { TSPLAT_REDUCE_SUM = 0, TSPLAT_REDUCE_ZMAX = 1 };
int tsplat_allreduce_image(tsplat_ctx *ctx, const float *const *peer_images, float *const *peer_out,
                           const float *const *peer_scale, int n_peers, int channels, int row0, int nrows, int op,
                           void *stream);

/* A process that drives several GPUs itself (no torchrun) lets `device_ordinal`'s kernels load / store `peer_ordinal`'s
 * memory (cudaDeviceEnablePeerAccess); already-enabled is not an error.  Multi-process runs get their peer mappings from
 * PyTorch symmetric memory instead. */
int tsplat_enable_peer_access(int device_ordinal, int peer_ordinal);

/* Replaces PeriodicSPH's accumulation pass (periodic_sph.py:59-88, overlay.py, shaders/overlay.wgsl): dst (R x R x
 * channels, device) = sum over n <= 128 replicas of weights[i] * src sampled bilinearly at the pixel shifted by the
 * clip-space offset (offsets_xy[2i], offsets_xy[2i+1]); a replica contributes only inside its own [-1,1]^2 + offset
 * quad.  offsets_xy / weights are HOST arrays. */
int tsplat_periodic_accumulate(tsplat_ctx *ctx, const float *src, float *dst, int channels, const float *offsets_xy,
                               const float *weights, int n, void *stream);

/* Surface render mode.  tsplat_set_surface replaces DepthSPHWithOcclusion's kernel texture (the LocalSphereKernel mip
 * chain, sph.py:446-455,497-501: 5440 host floats, same layout as tsplat_set_kernel_lut) and its density_cut uniform
 * (sph.py:503-506).  TSPLAT_MODE_SURFACE renders use this LUT instead of the SPH kernel LUT. */
int tsplat_set_surface(tsplat_ctx *ctx, const float *host_lut, int n_floats, float density_cut);

/* Replaces ColorAsSurfaceMap._encode_smoothing_filter_pass (colormap/surface.py:262-297, shaders/smooth.wgsl): bilateral
 * filter of channel 1 of a width x height x 2 device image (channel 0 is copied); in != out. */
int tsplat_bilateral_filter(tsplat_ctx *ctx, const float *in, float *out, int width, int height, float spatial_sigma,
                            float range_sigma, int kernel_size, void *stream);

/* The uniform block of shaders/surface.wgsl:14-24 (colormap/surface.py:341-354) plus its preprocessor switches. */
typedef struct {
    float depth_scale;
    float light_direction[3];
    float light_color[3];
    float ambient_color[3];
    float vmin, vmax;
    float window_aspect_ratio;      /* width/height of the output */
    int32_t material_colormap;      /* MATERIAL_COLORMAP: colour the surface by channel 0 through the 1-D LUT */
    int32_t log_scale;              /* MATERIAL_LOG */
} tsplat_surface_params;

/* Replaces ColorAsSurfaceMap's render pass (shaders/surface.wgsl): normals from central differences of the smoothed depth
 * (channel 1 of the res x res x 2 device image), Lambert + ambient lighting dimmed with depth, optional material colormap
 * (lut: device float32 RGBA, lut_w texels).  out: device buffer out_w*out_h*4 elements of out_fmt. */
int tsplat_surface_shade(tsplat_ctx *ctx, const float *smoothed, int res, const tsplat_surface_params *params,
                         const float *lut, int lut_w, void *out, int out_w, int out_h, int out_fmt, void *stream);

/* Device-side autorange (replaces the host read-back + np.percentile of colormap/implementation.py:381-425,
 * :512-531, :576-588).  "content" selects what the colormap looks at: 0 = channel 0 * scale (density), 1 = channel 1 /
 * channel 0 (weighted mean), 2 = every channel * scale (RGB maps, which include the count channel like the reference),
 * 3 = channel 0 * scale of the pixels whose channel 1 is positive (surface maps, colormap/surface.py:256-259).
 * tsplat_content_stats: finite min/max of the values and of their log10, whether any value is negative, counts.
 * tsplat_content_select: exact order statistics (0-based ranks among the finite values, ascending) of the values or of
 * their log10 -- a 3-pass radix select; the host interpolates percentiles from them like np.percentile.
 * Both calls synchronise the stream (they return host values). */
typedef struct {
    float lin_min, lin_max, log_min, log_max;     /* NaN when there is no finite value */
    int64_t n_finite_lin, n_finite_log;
    int32_t any_negative;
} tsplat_content_stats_t;
int tsplat_content_stats(tsplat_ctx *ctx, const float *image, int res, int channels, int content, float scale,
                         tsplat_content_stats_t *out, void *stream);
int tsplat_content_select(tsplat_ctx *ctx, const float *image, int res, int channels, int content, float scale,
                          int use_log, const int64_t *ranks, int n_ranks, float *out_values, void *stream);

/* out[i] = a[i] + b[i] * scale, used by PeriodicSPH-style accumulation and by tests (device pointers). */
int tsplat_image_axpy(tsplat_ctx *ctx, float *dst, const float *src, float scale, int64_t n, void *stream);

/* Replaces CellLayout.from_positions' numpy pipeline (cell_layout.py:87-112) for device-resident positions.
 * pos: device N x 3 array-of-structures in `dtype_bytes` (4 or 8) precision; box_min/cell_size are passed as doubles
 * that already hold the value numpy would compute in the position dtype.  Outputs (device): order[N] int64 = stable
 * argsort of the cell index, lengths[nside^3] int64.  status[0] (device int32) is set to 1 if any particle falls
 * outside [0, nside) (the reference raises ValueError, cell_layout.py:100-101).  work: device scratch of
 * tsplat_cell_layout_work_bytes(n, nside) bytes. */
int64_t tsplat_cell_layout_work_bytes(int64_t n, int nside);
int tsplat_cell_layout(int device_ordinal, const void *pos, int64_t n, int dtype_bytes, double box_min,
                       double cell_size, int nside, int64_t *order, int64_t *lengths, int32_t *status,
                       void *work, int64_t work_bytes, void *stream);

/* The same with the within-cell shuffle of CellLayout.randomize_within_cells (cell_layout.py:17-24) fused into the
 * scatter: inside every cell the stable order is permuted by a keyed pseudo-random permutation (4-round Feistel network,
 * cycle-walked), never across cells, so that any leading fraction of a cell is a fair subsample.  n < 2^31. */
int tsplat_cell_layout_shuffled(int device_ordinal, const void *pos, int64_t n, int dtype_bytes, double box_min,
                                double cell_size, int nside, uint32_t shuffle_seed, int64_t *order, int64_t *lengths,
                                int32_t *status, void *work, int64_t work_bytes, void *stream);

/* dst[i] = (float) src[order[i] * stride + offset] on the device: what the reference's loaders do on the host with
 * `array.astype(np.float32)[self._particle_order]` (loader.py:100-110).  src is float32 (src_bytes 4) or float64 (8). */
int tsplat_gather_f32(int device_ordinal, float *dst, const void *src, int src_bytes, int stride, int offset,
                      const int64_t *order, int64_t n, void *stream);

/* Host <-> device staging for the end-to-end (host buffer) path: thin cudaMemcpyAsync wrappers so the Python host
 * layer needs no other CUDA binding.  Replaces queue.write_buffer / queue.read_texture. */
int tsplat_memcpy_h2d(void *dst_dev, const void *src_host, int64_t bytes, void *stream);
int tsplat_memcpy_d2h(void *dst_host, const void *src_dev, int64_t bytes, void *stream);
int tsplat_stream_sync(void *stream);

/* Counters (forces a stream sync on the context's last stream). */
int tsplat_get_stats(tsplat_ctx *ctx, tsplat_stats *out);

#ifdef __cplusplus
}
#endif
#endif /* TSPLAT_H */
