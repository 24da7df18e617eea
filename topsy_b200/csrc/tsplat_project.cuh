// K1 (round 2): k_project_stream<MODE, CELL_W> -- project / cull / classify / direct splat as a per-warp software pipeline.
// Included by tsplat.cu (uses its ProjectArgs, Counters, ld4 / red_v* helpers).
//
// Replaces the per-thread-batch k_project_splat of round 1 (234 thread-instructions per particle on sub-pixel
// footprints, profiles/r01/prof_k1_c4s_v3.summary.txt).  What changed, and why:
//   * warps are persistent: a warp walks batches of 128 particles (batch = warp index + k * warps in the grid), so the
//     per-thread set-up (LUT staging, statistics flush) is paid once per warp instead of once per batch, and batch b+1
//     is already on its way into the warp's shared-memory stage (cp.async, 16 bytes per lane and array, L2 evict-first)
//     while batch b is classified: the prefetch holds no registers;
//   * stage 1 (classify) is branch-free per particle: project, pixel ranges, class.  Nothing that only covered
//     particles need (1/(h*h), 1/wpx, intensities) is computed here;
//   * covered small-footprint particles are COMPACTED (ballot + popc) into a ring of 32-byte raw records in shared
//     memory; stage 2 (finalize) runs once 32 of them are pending, one record per lane, fully converged -- so a batch in
//     which 28 % of the particles touch a pixel centre costs 28 % of the reciprocal / intensity work, not 100 %;
//   * stage 3 (work items = one cell column x two pixel rows, as in round 1) consumes a flattened item list 32 at a
//     time; the list lives in a ring (bit vector of record heads + popc to find the owner) and leftovers carry over to
//     the next group, so no RED instruction is issued with a mostly empty warp.  In the sub-pixel regime (at least
//     KP_INLINE_MIN_SINGLE of the 32 records cover exactly one cell) those single-cell records are emitted straight from
//     registers -- one RED, no list traffic -- and only the others are queued.
// Arithmetic is unchanged (project(), pixel_range(), sample_lut8(), one IEEE reciprocal of h*h and of wpx, the same
// products in the same order): images are bit-compatible with round 1 up to fp32 accumulation order.
#pragma once

constexpr int KP_THREADS = 128;
constexpr int KP_WARPS = KP_THREADS / 32;
constexpr int KP_Q1N = 160;                  // raw-record ring: < 32 pending + <= 128 pushed per batch
constexpr int KP_FRN = 64;                   // final-record ring: <= 31 records with pending items + <= 32 new (power of two)
constexpr int KP_BWN = 64;                   // item bit-vector ring, words: < 32 pending items + <= 32 * 32 new (power of two)
constexpr unsigned KP_INLINE_MIN_SINGLE = 12; // groups with at least this many single-cell records emit those from registers
constexpr int KP_MAX_R = 8192;               // j0, k0 are packed into 13 bits each; larger images defer every covered particle
constexpr int KP_MAX_SPAN = 8;               // direct particles cover at most 8 x 8 pixel centres

struct KpWarpSmem {
    float4 rawA[KP_Q1N];                     // px0 py1 wpx h
    float4 rawB[KP_Q1N];                     // w0 (w1 | cz) w2 packed(j0 | k0 << 13 | (ncols-1) << 26 | (nrows-1) << 29)
    float4 finA[KP_FRN];                     // px0 py1 inv v0
    float4 finB[KP_FRN];                     // v1 v2 packed  position of the record's item 0 in the item stream (13 bits)
    unsigned bits[KP_BWN];                   // bit p set: a record's first queued item sits at stream position p
};

struct KpBatch {
    float4 X, Y, Z, H, W0, W1, W2;
};

struct KpStage { float4 v[7][32]; };         // one batch in flight per warp: [array][lane] (conflict-free 128-bit accesses)

__device__ __forceinline__ void kp_cp_async16(void *smem_dst, const void *gsrc, unsigned src_bytes, uint64_t pol)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" :: "r"(d), "l"(gsrc), "r"(src_bytes), "l"(pol) : "memory");
}

// Issue the asynchronous copies of one batch into the warp's stage and return the lane's in-range mask (bit e set:
// particle e of the lane's 4-particle group is inside the requested range).  Lanes beyond the work list copy nothing.
template <int MODE>
__device__ __forceinline__ unsigned kp_issue(const ProjectArgs &a, int64_t batch, int lane, uint64_t pol, KpStage &st)
{
    unsigned mask = 0;
    const int64_t gi = batch * 32 + lane;
    if (gi < a.n_groups) {
        int64_t lo, hi, group;
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
        const int e_first = (int)max((int64_t)0, lo - base), e_last = (int)min((int64_t)4, hi - base);
        mask = ((1u << e_last) - 1u) & ~((1u << e_first) - 1u);
        // the last group of the buffer may be partial: copy only the bytes that exist, the rest of the 16 is zero-filled
        const int64_t left = a.n_total - base;
        const unsigned nbytes = left >= 4 ? 16u : (unsigned)left * 4u;
        kp_cp_async16(&st.v[0][lane], a.x + base, nbytes, pol);
        kp_cp_async16(&st.v[1][lane], a.y + base, nbytes, pol);
        kp_cp_async16(&st.v[2][lane], a.z + base, nbytes, pol);
        kp_cp_async16(&st.v[3][lane], a.h + base, nbytes, pol);
        kp_cp_async16(&st.v[4][lane], a.w0 + base, nbytes, pol);
        if (MODE == TSPLAT_MODE_WEIGHTED || MODE == TSPLAT_MODE_RGB) kp_cp_async16(&st.v[5][lane], a.w1 + base, nbytes, pol);
        if (MODE == TSPLAT_MODE_RGB) kp_cp_async16(&st.v[6][lane], a.w2 + base, nbytes, pol);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    return mask;
}

template <int MODE>
__device__ __forceinline__ void kp_take(const KpStage &st, int lane, KpBatch &b)
{
    asm volatile("cp.async.wait_group 0;" ::: "memory");      // the lane reads only what it copied itself
    b.X = st.v[0][lane]; b.Y = st.v[1][lane]; b.Z = st.v[2][lane]; b.H = st.v[3][lane]; b.W0 = st.v[4][lane];
    b.W1 = b.W2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == TSPLAT_MODE_WEIGHTED || MODE == TSPLAT_MODE_RGB) b.W1 = st.v[5][lane];
    if (MODE == TSPLAT_MODE_RGB) b.W2 = st.v[6][lane];
}

__device__ __forceinline__ float kp_get(const float4 &v, int e) { return e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w; }

// One cell (CELL_W horizontally adjacent pixels = one 128-bit RED: RGB 1, WEIGHTED / DEPTH 2, DENSITY 4) of pixel row k.
// [j0, j1] = the record's covered pixel columns, cj = cell column.
template <int MODE, int CELL_W>
__device__ __forceinline__ void kp_emit_cell(const ProjectArgs &a, const float *__restrict__ s_lut8, uint64_t pol_image,
                                             float px0, float py1, float inv, float v0, float v1, float v2, unsigned j0,
                                             unsigned j1, unsigned cj, unsigned k)
{
    const float fy = (float)k + 0.5f;
    const unsigned pix = k * (unsigned)a.R + cj * CELL_W;                  // R <= 8192: fits 32 bits
    if (CELL_W == 1) {
        const float K = sample_lut8(s_lut8, inv, px0, py1, (float)cj + 0.5f, fy);
        if (MODE == TSPLAT_MODE_RGB) {                                    // RGB counts fragments even where K == 0
            red_v4(a.image + 4 * (size_t)pix, v0 * K, v1 * K, v2 * K, 1.0f, pol_image);
        } else if (K != 0.0f) {                                           // adding +0 is a no-op
            const float val = K * v0;
            if (MODE == TSPLAT_MODE_DENSITY) red_v1(a.image + pix, val, pol_image);
            else red_v2(a.image + 2 * (size_t)pix, val, val * v1, pol_image);
        }
    } else {
        float Ks[CELL_W];
        bool any = false;
#pragma unroll
        for (int c = 0; c < CELL_W; ++c) {
            const unsigned j = cj * CELL_W + c;
            const bool in = (j >= j0) && (j <= j1);
            Ks[c] = in ? sample_lut8(s_lut8, inv, px0, py1, (float)j + 0.5f, fy) : 0.0f;
            any |= (Ks[c] != 0.0f);
        }
        if (any) {
            if (MODE == TSPLAT_MODE_DENSITY) {
                red_v4(a.image + pix, Ks[0] * v0, Ks[1] * v0, Ks[2 % CELL_W] * v0, Ks[3 % CELL_W] * v0, pol_image);
            } else {                                                      // two pixels x (val, val * q|cz)
                const float a0 = Ks[0] * v0, a1 = Ks[1] * v0;
                red_v4(a.image + 2 * (size_t)pix, a0, a0 * v1, a1, a1 * v1, pol_image);
            }
        }
    }
}

// One work item of a finalized record: cell column `local % ncj`, pixel rows k0 + 2 * (local / ncj) (+ 1).
template <int MODE, int CELL_W>
__device__ __forceinline__ void kp_emit_item(const ProjectArgs &a, const float *__restrict__ s_lut8, uint64_t pol_image,
                                             float px0, float py1, float inv, float v0, float v1, float v2, unsigned packed,
                                             unsigned local)
{
    constexpr int CELL_SHIFT = CELL_W == 4 ? 2 : CELL_W == 2 ? 1 : 0;
    const unsigned j0 = packed & 0x1fffu, k0 = (packed >> 13) & 0x1fffu, nc1 = (packed >> 26) & 7u, nr1 = packed >> 29;
    const unsigned j1 = j0 + nc1;
    const unsigned cj0 = j0 >> CELL_SHIFT, ncj = (j1 >> CELL_SHIFT) - cj0 + 1u;
    const unsigned dk2 = (unsigned)(local >= ncj) + (unsigned)(local >= 2u * ncj) + (unsigned)(local >= 3u * ncj);
    const unsigned cj = cj0 + (local - dk2 * ncj);
    const unsigned k = k0 + 2u * dk2;
    kp_emit_cell<MODE, CELL_W>(a, s_lut8, pol_image, px0, py1, inv, v0, v1, v2, j0, j1, cj, k);
    if (2u * dk2 < nr1) kp_emit_cell<MODE, CELL_W>(a, s_lut8, pol_image, px0, py1, inv, v0, v1, v2, j0, j1, cj, k + 1u);
}

template <int MODE, int CELL_W>
__global__ void __launch_bounds__(KP_THREADS) k_project_stream(const ProjectArgs a)
{
    constexpr int CELL_SHIFT = CELL_W == 4 ? 2 : CELL_W == 2 ? 1 : 0;
    __shared__ float s_lut8[64];
    __shared__ KpWarpSmem s_warp[KP_WARPS];
    __shared__ KpStage s_stage[KP_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    KpWarpSmem &S = s_warp[warp];
    KpStage &stage = s_stage[warp];
    if (threadIdx.x < 64) s_lut8[threadIdx.x] = __ldg(a.lut + lut_offset(3) + threadIdx.x);
    for (int i = lane; i < KP_BWN; i += 32) S.bits[i] = 0u;
    __syncthreads();                              // the only block-wide barrier: warps are independent from here on
    const uint64_t pol_stream = l2_policy_evict_first(), pol_image = l2_policy_evict_last();
    const unsigned lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);
    const bool direct_ok = a.R <= KP_MAX_R;

    // pipeline state (warp-uniform)
    unsigned q1_head = 0, q1_count = 0;           // raw ring
    unsigned rec_tail = 0, rec_base = 0;          // final records enqueued / records whose first item was consumed
    unsigned items_tail = 0, items_done = 0;      // item stream positions (items_done is a multiple of 32)
    // statistics (per lane, flushed once per warp): direct = in range - culled - deferred
    unsigned n_in = 0, n_culled = 0, n_deferred = 0, n_reds = 0;

    // ---- stage 3: consume 32 items of the stream -------------------------------------------------------------
    auto consume = [&]() {
        const unsigned wi = (items_done >> 5) & (KP_BWN - 1);
        const unsigned word = S.bits[wi];
        const unsigned t = items_done + lane;
        const unsigned rk = rec_base + __popc(word & le_mask) - 1u;
        if (t < items_tail) {
            const float4 ra = S.finA[rk & (KP_FRN - 1)];
            const float4 rb = S.finB[rk & (KP_FRN - 1)];
            const unsigned local = (t - __float_as_uint(rb.w)) & 0x1fffu;      // position of the record's item 0 is stored
            kp_emit_item<MODE, CELL_W>(a, s_lut8, pol_image, ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, __float_as_uint(rb.z), local);
        }
        rec_base += __popc(word);
        items_done += 32u;
        __syncwarp();
        if (lane == 0) S.bits[wi] = 0u;
        __syncwarp();
    };

    // ---- stage 2: finalize up to 32 raw records (one per lane) -------------------------------------------------
    auto finalize = [&](const unsigned nrec) {
        const bool act = (unsigned)lane < nrec;
        unsigned idx = q1_head + lane;
        if (idx >= KP_Q1N) idx -= KP_Q1N;
        float4 A = make_float4(0.f, 0.f, 1.f, 1.f), B = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act) { A = S.rawA[idx]; B = S.rawB[idx]; }
        const unsigned packed = __float_as_uint(B.w);
        const float rhh = 1.0f / (A.w * A.w);
        const float inv = 1.0f / A.z;
        const float v0 = B.x * rhh;
        const float v1 = MODE == TSPLAT_MODE_RGB ? B.y * rhh : B.y;
        const float v2 = MODE == TSPLAT_MODE_RGB ? B.z * rhh : 0.0f;
        const unsigned j0 = packed & 0x1fffu, nc1 = (packed >> 26) & 7u, nr1 = packed >> 29;
        const unsigned ncj = ((j0 + nc1) >> CELL_SHIFT) - (j0 >> CELL_SHIFT) + 1u;
        const unsigned n_items = act ? ncj * ((nr1 + 2u) >> 1) : 0u;
        if (act) n_reds += ncj * (nr1 + 1u);
        // sub-pixel regime: records that cover exactly one cell are emitted from registers (one RED, no list traffic)
        const bool single = act && ncj == 1u && nr1 == 0u;
        const bool inline_mode = (unsigned)__popc(__ballot_sync(0xffffffffu, single)) >= KP_INLINE_MIN_SINGLE;
        const unsigned n_enq = (inline_mode && single) ? 0u : n_items;
        const unsigned enq = __ballot_sync(0xffffffffu, n_enq > 0u);
        if (enq) {
            unsigned incl = n_enq;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
            if (n_enq) {
                const unsigned pos = items_tail + incl - n_enq;
                const unsigned slot = (rec_tail + __popc(enq & lt_mask)) & (KP_FRN - 1);
                S.finA[slot] = make_float4(A.x, A.y, inv, v0);
                S.finB[slot] = make_float4(v1, v2, B.w, __uint_as_float(pos & 0x1fffu));
                atomicOr(&S.bits[(pos >> 5) & (KP_BWN - 1)], 1u << (pos & 31u));
            }
            items_tail += total;
            rec_tail += (unsigned)__popc(enq);
        }
        if (inline_mode && single) {
            const unsigned k0 = (packed >> 13) & 0x1fffu;
            kp_emit_cell<MODE, CELL_W>(a, s_lut8, pol_image, A.x, A.y, inv, v0, v1, v2, j0, j0 + nc1, j0 >> CELL_SHIFT, k0);
        }
        q1_head += nrec;
        if (q1_head >= KP_Q1N) q1_head -= KP_Q1N;
        q1_count -= nrec;
        __syncwarp();
    };

    // ---- stage 1: one particle per lane and call, branch-free ---------------------------------------------------
    unsigned defer_mask = 0;
    auto classify = [&](const KpBatch &b, const unsigned in_mask, const int e) {
        const float x = kp_get(b.X, e), y = kp_get(b.Y, e), z = kp_get(b.Z, e), h = kp_get(b.H, e);
        const Proj p = project(x, y, z, h, a.cam);
        int j0, j1, k0, k1;
        pixel_range(p.px0, p.px1, a.R, j0, j1);
        pixel_range(p.py0, p.py1, a.R, k0, k1);
        const bool in = (in_mask >> e) & 1u;
        const bool keep = in && p.keep;
        // nx1 = columns - 1, ny1 = rows - 1 (negative: no pixel centre covered); one unsigned compare per axis tests
        // "non-empty and at most KP_MAX_SPAN wide"
        const int nx1 = j1 - j0, ny1 = k1 - k0;
        const bool covered = keep && nx1 >= 0 && ny1 >= 0;
        const bool direct = keep && direct_ok && p.wpx <= DIRECT_MAX_WPX && (unsigned)nx1 < (unsigned)KP_MAX_SPAN &&
                            (unsigned)ny1 < (unsigned)KP_MAX_SPAN;
        const bool defer = covered && !direct;
        n_culled += (unsigned)(in && !p.keep);
        defer_mask |= (unsigned)defer << e;
        const unsigned dm = __ballot_sync(0xffffffffu, direct);
        if (direct) {
            unsigned idx = q1_head + q1_count + (unsigned)__popc(dm & lt_mask);
            if (idx >= KP_Q1N) idx -= KP_Q1N;
            const unsigned packed = (unsigned)j0 | ((unsigned)k0 << 13) | ((unsigned)nx1 << 26) | ((unsigned)ny1 << 29);
            const float w1 = MODE == TSPLAT_MODE_DEPTH ? p.cz : kp_get(b.W1, e);
            S.rawA[idx] = make_float4(p.px0, p.py1, p.wpx, h);
            S.rawB[idx] = make_float4(kp_get(b.W0, e), w1, kp_get(b.W2, e), __uint_as_float(packed));
        }
        q1_count += (unsigned)__popc(dm);
    };

    // deferred (large-footprint) particles of the batch: ONE queue reservation per warp (same-address atomics serialise in
    // the L2 at ~1 per ns), then every lane re-projects its deferred particles and writes the 32-byte queue records
    auto append_deferred = [&](const KpBatch &b) {
        const unsigned cnt = (unsigned)__popc(defer_mask);
        unsigned incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned qb = 0;
        if (lane == 0) qb = atomicAdd(&a.counters->q_count, total);
        qb = __shfl_sync(0xffffffffu, qb, 0);
        unsigned slot = qb + incl - cnt;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (defer_mask & (1u << e)) {
                if (slot < a.queue_cap) {
                    const float h = kp_get(b.H, e);
                    const Proj p = project(kp_get(b.X, e), kp_get(b.Y, e), kp_get(b.Z, e), h, a.cam);
                    const float rhh = 1.0f / (h * h);
                    const float v0 = kp_get(b.W0, e) * rhh;
                    float v1, v2 = 0.0f;
                    if (MODE == TSPLAT_MODE_RGB) { v1 = kp_get(b.W1, e) * rhh; v2 = kp_get(b.W2, e) * rhh; }
                    else if (MODE == TSPLAT_MODE_DEPTH) v1 = p.cz;
                    else v1 = kp_get(b.W1, e);
                    float4 *q = reinterpret_cast<float4 *>(a.queue + slot);
                    q[0] = make_float4(p.px0, p.px1, p.py0, p.py1);
                    q[1] = make_float4(p.wpx, v0, v1, v2);
                }
                ++slot;
            }
        }
    };

    // ---- main loop over the warp's batches -----------------------------------------------------------------------
    const int64_t n_batches = (a.n_groups + 31) >> 5;
    const int64_t stride = (int64_t)gridDim.x * KP_WARPS;
    int64_t batch = (int64_t)blockIdx.x * KP_WARPS + warp;
    unsigned nxt_mask = kp_issue<MODE>(a, batch, lane, pol_stream, stage);
#pragma unroll 1
    for (; batch < n_batches; batch += stride) {
        KpBatch cur;
        kp_take<MODE>(stage, lane, cur);
        const unsigned in_mask = nxt_mask;
        nxt_mask = 0;
        if (batch + stride < n_batches) nxt_mask = kp_issue<MODE>(a, batch + stride, lane, pol_stream, stage);
        n_in += (unsigned)__popc(in_mask);
        defer_mask = 0;
        classify(cur, in_mask, 0);
        classify(cur, in_mask, 1);
        classify(cur, in_mask, 2);
        classify(cur, in_mask, 3);
        if (__any_sync(0xffffffffu, defer_mask != 0u)) { n_deferred += (unsigned)__popc(defer_mask); append_deferred(cur); }
        __syncwarp();
#pragma unroll 1
        while (q1_count >= 32u) {
            finalize(32u);
#pragma unroll 1
            while (items_tail - items_done >= 32u) consume();
        }
    }
    // ---- flush ----------------------------------------------------------------------------------------------------
    if (q1_count) finalize(q1_count);
#pragma unroll 1
    while (items_done < items_tail) consume();

    // warp-aggregated statistics, spread over STAT_SLOTS counter slots (tsplat_get_stats sums them)
#ifndef TSPLAT_NO_STATS
    const unsigned w_culled = __reduce_add_sync(0xffffffffu, n_culled), w_defer = __reduce_add_sync(0xffffffffu, n_deferred);
    const unsigned w_direct = __reduce_add_sync(0xffffffffu, n_in) - w_culled - w_defer, w_reds = __reduce_add_sync(0xffffffffu, n_reds);
    if (lane == 0) {
        StatSlot *slot = a.counters->slots + ((blockIdx.x * KP_WARPS + warp) & (STAT_SLOTS - 1));
        const unsigned long long cd = (unsigned long long)w_culled | ((unsigned long long)w_direct << 32);
        if (cd) atomicAdd(&slot->culled_direct, cd);
        if (w_reds) atomicAdd(&slot->reds, (unsigned long long)w_reds);
        if (a.small_call && w_defer) atomicAdd(&a.counters->huge, (unsigned long long)w_defer);
    }
#endif
}

// resident CTAs per SM the launch sizes its grid for: what the occupancy calculator says for this instantiation
// (register-limited: 4-5 CTAs of 4 warps); TSPLAT_KP_CTAS overrides it for tuning runs
template <int MODE, int CELL_W>
static int kp_ctas_per_sm()
{
    static const int v = [] {
        const char *e = getenv("TSPLAT_KP_CTAS");
        int n = e ? atoi(e) : 0;
        if (n <= 0 && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_project_stream<MODE, CELL_W>, KP_THREADS, 0) != cudaSuccess) n = 4;
        return n > 0 ? n : 4;
    }();
    return v;
}
