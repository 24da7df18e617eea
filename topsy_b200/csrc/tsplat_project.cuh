// K1 (round 2): k_project_stream<MODE, CELL_W> -- project / cull / classify / direct splat as a per-warp software pipeline.
// Included by tsplat.cu (uses its ProjectArgs, Counters, ld4 / red_v* helpers).
//
// Replaces the per-thread-batch k_project_splat of round 1 (234 thread-instructions per particle on sub-pixel
// footprints, profiles/r01/prof_k1_c4s_v3.summary.txt).  What changed, and why:
//   * warps are persistent: a warp walks batches of 128 particles (batch = warp index + k * warps in the grid), so the
//     per-thread set-up (LUT staging, statistics flush) is paid once per warp instead of once per batch, and batch b+1
//     is already on its way into the warp's shared-memory stage (cp.async, 16 bytes per lane and array, L2 evict-first)
//     while batch b is classified: the prefetch holds no registers;
//   * stage 1 (classify) is branch-free per particle: project, pixel ranges, class, predicated stores.  Nothing that
//     only covered particles need (1/(h*h), 1/wpx, intensities) is computed here;
//   * covered small-footprint particles are COMPACTED (ballot + popc) into two rings of 32-byte raw records in shared
//     memory -- stack A: the footprint lies inside ONE cell (the sub-pixel regime), stack B: several cells.  A stack is
//     drained whenever it holds 32 records, one record per lane, fully converged: a batch in which 28 % of the particles
//     touch a pixel centre pays 28 % of the reciprocal / intensity work, not 100 %;
//   * stack A records are finished straight from registers: two reciprocals, one LUT fetch, one RED;
//   * stack B records are expanded into work items (one cell column x two pixel rows, as in round 1) on a flattened item
//     list that is consumed 32 at a time; the list lives in a ring (bit vector of record heads + popc to find the owner)
//     and leftovers carry over to the next group, so no RED instruction is issued with a mostly empty warp.
// Arithmetic is unchanged (project(), pixel_range(), sample_lut8(), one IEEE reciprocal of h*h and of wpx, the same
// products in the same order; inv * 8 is an exact scaling): images are bit-compatible with round 1 up to fp32
// accumulation order.
#pragma once

constexpr int KP_THREADS = 128;
constexpr int KP_WARPS = KP_THREADS / 32;
constexpr int KP_QN = 128;                   // raw-record store shared by two stacks: 2 x (< 32 pending) + <= 64 pushed per pair of particles
constexpr int KP_FRN = 64;                   // final-record ring: <= 31 records with pending items + <= 32 new (power of two)
constexpr int KP_BWN = 64;                   // item bit-vector ring, words: < 32 pending items + <= 32 * 32 new (power of two)
constexpr int KP_MAX_R = 8192;               // j0, k0 are packed into 13 bits each; larger images defer every covered particle
constexpr int KP_MAX_SPAN = 8;               // direct particles cover at most 8 x 8 pixel centres

// Raw records wait in two STACKS that share one array (the order in which records are finished does not matter):
// stack A (footprint inside one cell) grows up from slot 0, stack B (several cells) grows down from slot KP_QN - 1.
struct KpRaw {
    float4 a[KP_QN];                         // px0 py1 wpx h
    float4 b[KP_QN];                         // w0 (w1 | cz) w2 packed(j0 | k0 << 13 | (ncols-1) << 26 | (nrows-1) << 29)
};

struct KpWarpSmem {
    KpRaw raw;
    float4 finA[KP_FRN];                     // px0 py1 inv*8 v0
    float4 finB[KP_FRN];                     // v1 v2 packed  position of the record's item 0 in the item stream (13 bits)
    unsigned bits[KP_BWN];                   // bit p set: a record's first item sits at stream position p
};

struct KpBatch { float4 X, Y, Z, H, W0, W1, W2; };

struct KpStage { float4 v[7][32]; };         // one batch in flight per warp: [array][lane] (conflict-free 128-bit accesses)

__device__ __forceinline__ void kp_cp_async16(void *smem_dst, const void *gsrc, uint64_t pol)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(d), "l"(gsrc), "l"(pol) : "memory");
}

__device__ __forceinline__ void kp_cp_async16_partial(void *smem_dst, const void *gsrc, unsigned src_bytes, uint64_t pol)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" :: "r"(d), "l"(gsrc), "r"(src_bytes), "l"(pol) : "memory");
}

// Issue the asynchronous copies of group index gi (a group = 4 consecutive particles = one 128-bit load per array; gi
// counts groups within the call's work list and fits 32 bits) into the warp's stage and return the lane's in-range mask
// (bit e set: particle e of the lane's group is inside the requested range).  Lanes beyond the work list copy nothing.
template <int MODE>
__device__ __forceinline__ unsigned kp_issue(const ProjectArgs &a, unsigned gi, int lane, uint64_t pol, KpStage &st)
{
    unsigned mask = 0;
    if (gi < (unsigned)a.n_groups) {
        unsigned group;
        if (a.table.n > 0) {
            int l = 0, r = a.table.n;            // invariant: gprefix[l] <= gi < gprefix[r]
            while (r - l > 1) {
                const int m = (l + r) >> 1;
                if (a.table.gprefix[m] <= (int64_t)gi) l = m; else r = m;
            }
            const int64_t lo = a.table.start[l], hi = a.table.end[l];
            const int64_t g64 = (lo >> 2) + ((int64_t)gi - a.table.gprefix[l]);
            const int64_t base = g64 << 2;
            const int e_first = (int)max((int64_t)0, lo - base), e_last = (int)min((int64_t)4, hi - base);
            mask = ((1u << e_last) - 1u) & ~((1u << e_first) - 1u);
            group = (unsigned)g64;
        } else {
            // single range [start, end): only its first and last group can be partial
            group = (unsigned)a.g0 + gi;
            mask = 0xfu;
            if (gi - 1u >= (unsigned)a.n_groups - 2u) {      // gi == 0 or gi == n_groups - 1 (one compare, rarely taken)
                if (gi == 0u) mask &= ~((1u << (unsigned)(a.start & 3)) - 1u);
                if (gi == (unsigned)a.n_groups - 1u) mask &= (2u << (unsigned)((a.end - 1) & 3)) - 1u;
            }
        }
        const float4 *px = reinterpret_cast<const float4 *>(a.x) + group, *py = reinterpret_cast<const float4 *>(a.y) + group;
        const float4 *pz = reinterpret_cast<const float4 *>(a.z) + group, *ph = reinterpret_cast<const float4 *>(a.h) + group;
        const float4 *p0 = reinterpret_cast<const float4 *>(a.w0) + group, *p1 = reinterpret_cast<const float4 *>(a.w1) + group;
        const float4 *p2 = reinterpret_cast<const float4 *>(a.w2) + group;
        if (group < (unsigned)(a.n_total >> 2)) {      // the group lies wholly inside the buffers (n_total < 2^33 per call)
            kp_cp_async16(&st.v[0][lane], px, pol); kp_cp_async16(&st.v[1][lane], py, pol); kp_cp_async16(&st.v[2][lane], pz, pol);
            kp_cp_async16(&st.v[3][lane], ph, pol); kp_cp_async16(&st.v[4][lane], p0, pol);
            if (MODE == TSPLAT_MODE_WEIGHTED || MODE == TSPLAT_MODE_RGB) kp_cp_async16(&st.v[5][lane], p1, pol);
            if (MODE == TSPLAT_MODE_RGB) kp_cp_async16(&st.v[6][lane], p2, pol);
        } else {
            // the last group of the buffer is partial: copy only the bytes that exist, the rest of the 16 is zero-filled
            const unsigned nb = (unsigned)(a.n_total & 3) * 4u;
            kp_cp_async16_partial(&st.v[0][lane], px, nb, pol); kp_cp_async16_partial(&st.v[1][lane], py, nb, pol);
            kp_cp_async16_partial(&st.v[2][lane], pz, nb, pol); kp_cp_async16_partial(&st.v[3][lane], ph, nb, pol);
            kp_cp_async16_partial(&st.v[4][lane], p0, nb, pol);
            if (MODE == TSPLAT_MODE_WEIGHTED || MODE == TSPLAT_MODE_RGB) kp_cp_async16_partial(&st.v[5][lane], p1, nb, pol);
            if (MODE == TSPLAT_MODE_RGB) kp_cp_async16_partial(&st.v[6][lane], p2, nb, pol);
        }
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

// Level-3 LUT fetch for a pixel centre the footprint COVERS (px0 <= fx < px1, py0 <= fy < py1).  Same operations as
// sample_lut8(): u = (fx - px0) * inv, floor(u * 8) -- with inv8 = inv * 8 (an exact scaling, so (fx - px0) * inv8 ==
// ((fx - px0) * inv) * 8 bit for bit); covered centres have u, v >= 0, so only the upper clamp can ever act.
__device__ __forceinline__ float kp_lut8_covered(const float (&s_lut8)[64], float inv8, float px0, float py1, float fx, float fy)
{
    const int iu = min(__float2int_rd((fx - px0) * inv8), 7);
    const int iv = min(__float2int_rd((py1 - fy) * inv8), 7);
    return s_lut8[iv * 8 + iu];
}

// One cell (CELL_W horizontally adjacent pixels = one 128-bit RED: RGB 1, WEIGHTED / DEPTH 2, DENSITY 4) of pixel row k.
// [j0, j1] = the record's covered pixel columns, cj = cell column; row k is covered.
template <int MODE, int CELL_W>
__device__ __forceinline__ void kp_emit_cell(const ProjectArgs &a, const float (&s_lut8)[64], uint64_t pol_image,
                                             float px0, float py1, float inv8, float v0, float v1, float v2, unsigned j0,
                                             unsigned j1, unsigned cj, unsigned k)
{
    const float fy = (float)k + 0.5f;
    const unsigned pix = k * (unsigned)a.R + cj * CELL_W;                  // R <= 8192: fits 32 bits
    if (CELL_W == 1) {
        const float K = kp_lut8_covered(s_lut8, inv8, px0, py1, (float)cj + 0.5f, fy);
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
            // columns of the cell outside [j0, j1] are not covered: clamp the column so that the fetch stays in the table
            const unsigned jc = min(max(j, j0), j1);
            const float Kc = kp_lut8_covered(s_lut8, inv8, px0, py1, (float)jc + 0.5f, fy);
            Ks[c] = (j == jc) ? Kc : 0.0f;
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

template <int MODE, int CELL_W>
__global__ void __launch_bounds__(KP_THREADS, 5) k_project_stream(const ProjectArgs a)
{
    constexpr int CELL_SHIFT = CELL_W == 4 ? 2 : CELL_W == 2 ? 1 : 0;
    __shared__ float s_lut8[64];
    __shared__ KpWarpSmem s_warp[KP_WARPS];
    __shared__ KpStage s_stage[KP_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    KpWarpSmem &S = s_warp[warp];
    KpStage &stage = s_stage[warp];
    if (threadIdx.x < 64) s_lut8[threadIdx.x] = __ldg(a.lut + lut_offset(3) + threadIdx.x);
    if (blockIdx.x == 0 && threadIdx.x == 64) a.counters->call_particles = a.call_particles;
    for (int i = lane; i < KP_BWN; i += 32) S.bits[i] = 0u;
    __syncthreads();                              // the only block-wide barrier: warps are independent from here on
    const uint64_t pol_stream = l2_policy_evict_first(), pol_image = l2_policy_evict_last();
    const unsigned lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);
    const float direct_max_wpx = a.R <= KP_MAX_R ? DIRECT_MAX_WPX : -1.0f;      // larger images defer every covered particle

    // pipeline state (warp-uniform)
    unsigned qa_count = 0, qb_count = 0;          // records on the raw stacks A / B
    unsigned rec_tail = 0, rec_base = 0;          // final records enqueued / records whose first item was consumed
    unsigned items_tail = 0, items_done = 0;      // item stream positions (items_done is a multiple of 32)
    // statistics (per lane, flushed once per warp): direct = in range - culled - deferred
    unsigned n_in = 0, n_culled = 0, n_deferred = 0, n_reds = 0;

    // ---- stage 3: consume 32 work items of stack-B records ----------------------------------------------------------
    // item `local` of a record = cell column local % ncj, pixel rows k0 + 2 * (local / ncj) (+ 1)
    auto emit_item = [&](const unsigned t, const unsigned rk) {
        const float4 ra = S.finA[rk & (KP_FRN - 1)];      // px0 py1 inv8 v0
        const float4 rb = S.finB[rk & (KP_FRN - 1)];      // v1 v2 packed pos
        const unsigned packed = __float_as_uint(rb.z);
        const unsigned local = (t - __float_as_uint(rb.w)) & 0x1fffu;
        const unsigned j0 = packed & 0x1fffu, k0 = (packed >> 13) & 0x1fffu, nc1 = (packed >> 26) & 7u, nr1 = packed >> 29;
        const unsigned j1 = j0 + nc1;
        const unsigned cj0 = j0 >> CELL_SHIFT, ncj = (j1 >> CELL_SHIFT) - cj0 + 1u;
        const unsigned dk2 = (unsigned)(local >= ncj) + (unsigned)(local >= 2u * ncj) + (unsigned)(local >= 3u * ncj);
        const unsigned cj = cj0 + (local - dk2 * ncj);
        const unsigned k = k0 + 2u * dk2;
        kp_emit_cell<MODE, CELL_W>(a, s_lut8, pol_image, ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, j0, j1, cj, k);
        if (2u * dk2 < nr1) kp_emit_cell<MODE, CELL_W>(a, s_lut8, pol_image, ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, j0, j1, cj, k + 1u);
    };
    auto consume = [&]() {
        const unsigned wi = (items_done >> 5) & (KP_BWN - 1);
        const unsigned word = S.bits[wi];
        const unsigned t = items_done + lane;
        const unsigned rk = rec_base + __popc(word & le_mask) - 1u;
        if (t < items_tail) emit_item(t, rk);
        rec_base += __popc(word);
        items_done += 32u;
        __syncwarp();
        if (lane == 0) S.bits[wi] = 0u;
        __syncwarp();
    };
    // 64 items per call (both words are full): two independent record fetch -> LUT -> RED chains in flight per lane
    auto consume2 = [&]() {
        const unsigned wi0 = (items_done >> 5) & (KP_BWN - 1), wi1 = (wi0 + 1u) & (KP_BWN - 1);
        const unsigned word0 = S.bits[wi0], word1 = S.bits[wi1];
        const unsigned t0 = items_done + lane;
        const unsigned n0 = (unsigned)__popc(word0);
        const unsigned rk0 = rec_base + __popc(word0 & le_mask) - 1u;
        const unsigned rk1 = rec_base + n0 + __popc(word1 & le_mask) - 1u;
        emit_item(t0, rk0);
        emit_item(t0 + 32u, rk1);
        rec_base += n0 + (unsigned)__popc(word1);
        items_done += 64u;
        __syncwarp();
        if (lane < 2) S.bits[lane ? wi1 : wi0] = 0u;
        __syncwarp();
    };

    // what stage 2 computes for every raw record, whichever stack it came from
    struct Fin { float px0, py1, inv8, v0, v1, v2; unsigned packed; };
    auto finish = [&](const unsigned idx, const bool act) -> Fin {
        float4 A = make_float4(0.f, 0.f, 1.f, 1.f), B = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act) { A = S.raw.a[idx]; B = S.raw.b[idx]; }
        const float rhh = 1.0f / (A.w * A.w);
        Fin f;
        f.px0 = A.x; f.py1 = A.y;
        f.inv8 = (1.0f / A.z) * 8.0f;
        f.v0 = B.x * rhh;
        f.v1 = MODE == TSPLAT_MODE_RGB ? B.y * rhh : B.y;
        f.v2 = MODE == TSPLAT_MODE_RGB ? B.z * rhh : 0.0f;
        f.packed = __float_as_uint(B.w);
        return f;
    };

    // ---- stage 2, stack A: up to 32 records whose footprint is one cell of one pixel row -- one RED from registers ------
    auto finalize_a = [&](const unsigned nrec) {
        const bool act = (unsigned)lane < nrec;
        const Fin f = finish(qa_count - nrec + (unsigned)lane, act);      // the top nrec records of stack A
        if (act) {
            const unsigned j0 = f.packed & 0x1fffu, nc1 = (f.packed >> 26) & 7u;
            n_reds += 1u;
            kp_emit_cell<MODE, CELL_W>(a, s_lut8, pol_image, f.px0, f.py1, f.inv8, f.v0, f.v1, f.v2, j0, j0 + nc1, j0 >> CELL_SHIFT,
                                       (f.packed >> 13) & 0x1fffu);
        }
        qa_count -= nrec;
        __syncwarp();
    };

    // ---- stage 2, stack B: up to 32 multi-cell records become final records + work items ---------------------------------
    auto finalize_b = [&](const unsigned nrec) {
        const bool act = (unsigned)lane < nrec;
        const Fin f = finish((unsigned)KP_QN - qb_count + (unsigned)lane, act);      // the top nrec records of stack B
        const unsigned j0 = f.packed & 0x1fffu, nc1 = (f.packed >> 26) & 7u, nr1 = f.packed >> 29;
        const unsigned ncj = ((j0 + nc1) >> CELL_SHIFT) - (j0 >> CELL_SHIFT) + 1u;
        const unsigned n_items = act ? ncj * ((nr1 + 2u) >> 1) : 0u;
        if (act) n_reds += ncj * (nr1 + 1u);
        unsigned incl = n_items;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
        if (act) {
            const unsigned pos = items_tail + incl - n_items;
            const unsigned slot = (rec_tail + (unsigned)lane) & (KP_FRN - 1);
            S.finA[slot] = make_float4(f.px0, f.py1, f.inv8, f.v0);
            S.finB[slot] = make_float4(f.v1, f.v2, __uint_as_float(f.packed), __uint_as_float(pos & 0x1fffu));
            atomicOr(&S.bits[(pos >> 5) & (KP_BWN - 1)], 1u << (pos & 31u));
        }
        items_tail += total;
        rec_tail += nrec;
        qb_count -= nrec;
        __syncwarp();
    };

    // ---- stage 1: one particle per lane and call; branch-free up to the stack store ----------------------------------
    unsigned defer_mask = 0;
    auto classify = [&](const KpBatch &b, const unsigned in_mask, const int e) {
        const float h = kp_get(b.H, e);
        const Proj p = project(kp_get(b.X, e), kp_get(b.Y, e), kp_get(b.Z, e), h, a.cam);
        // first covered column / row, clipped below at 0; one-past-last, clipped to [0, R]  (== pixel_range())
        const int j0 = max(__float2int_ru(p.px0 - 0.5f), 0), k0 = max(__float2int_ru(p.py0 - 0.5f), 0);
        const int jend = min(max(__float2int_ru(p.px1 - 0.5f), 0), a.R), kend = min(max(__float2int_ru(p.py1 - 0.5f), 0), a.R);
        const bool in = (in_mask >> e) & 1u;
        const bool keep = in && p.keep;
        // nx1 = columns - 1, ny1 = rows - 1 (negative: no pixel centre covered; no overflow: 0 <= jend <= R, j0 >= 0);
        // one unsigned compare per axis tests "non-empty and at most KP_MAX_SPAN wide"
        const int nx1 = jend - 1 - j0, ny1 = kend - 1 - k0;
        const bool covered = keep && nx1 >= 0 && ny1 >= 0;
        const bool direct = keep && p.wpx <= direct_max_wpx && (unsigned)nx1 < (unsigned)KP_MAX_SPAN &&
                            (unsigned)ny1 < (unsigned)KP_MAX_SPAN;
        const bool defer = covered && !direct;
        // stack A: one pixel row and all columns inside one cell
        const bool one = CELL_W == 1 ? (nx1 | ny1) == 0
                                     : (ny1 == 0 && ((unsigned)j0 >> CELL_SHIFT) == ((unsigned)(jend - 1) >> CELL_SHIFT));
        const bool to_a = direct && one, to_b = direct && !one;
        n_culled += (unsigned)(in && !p.keep);
        defer_mask |= (unsigned)defer << e;
        const unsigned ma = __ballot_sync(0xffffffffu, to_a), mb = __ballot_sync(0xffffffffu, to_b);
        if (direct) {
            const unsigned rank = (unsigned)__popc((to_a ? ma : mb) & lt_mask);
            const unsigned idx = to_a ? qa_count + rank : (unsigned)(KP_QN - 1) - qb_count - rank;
            KpRaw &Q = S.raw;
            // disjoint bit fields, written as a sum so that the compiler may fold shifts and adds into multiply-adds
            const unsigned packed = (unsigned)j0 + (unsigned)k0 * 8192u + (unsigned)nx1 * 67108864u + (unsigned)ny1 * 536870912u;
            Q.a[idx] = make_float4(p.px0, p.py1, p.wpx, h);
            Q.b[idx] = make_float4(kp_get(b.W0, e), MODE == TSPLAT_MODE_DEPTH ? p.cz : kp_get(b.W1, e), kp_get(b.W2, e),
                                   __uint_as_float(packed));
        }
        qa_count += (unsigned)__popc(ma);
        qb_count += (unsigned)__popc(mb);
    };

    auto drain = [&]() {
        __syncwarp();
#pragma unroll 1
        while (qa_count >= 32u) finalize_a(32u);
#pragma unroll 1
        while (qb_count >= 32u) {
            finalize_b(32u);
#pragma unroll 1
            while (items_tail - items_done >= 64u) consume2();
            if (items_tail - items_done >= 32u) consume();
        }
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
    const unsigned n_groups = (unsigned)a.n_groups;                     // < 2^31 per launch
    const unsigned gstride = gridDim.x * (unsigned)(KP_WARPS * 32);      // groups between two batches of a warp
    unsigned gi = (blockIdx.x * (unsigned)KP_WARPS + (unsigned)warp) * 32u + (unsigned)lane;
    const unsigned gi_warp_end = n_groups + (unsigned)lane;             // (gi - lane) < n_groups  <=>  the batch exists
    unsigned nxt_mask = kp_issue<MODE>(a, gi, lane, pol_stream, stage);
#pragma unroll 1
    for (; gi < gi_warp_end; gi += gstride) {
        KpBatch cur;
        kp_take<MODE>(stage, lane, cur);
        const unsigned in_mask = nxt_mask;
        nxt_mask = 0;
        if (gi + gstride < gi_warp_end) nxt_mask = kp_issue<MODE>(a, gi + gstride, lane, pol_stream, stage);
        n_in += (unsigned)__popc(in_mask);
        defer_mask = 0;
        classify(cur, in_mask, 0);
        classify(cur, in_mask, 1);
        drain();
        classify(cur, in_mask, 2);
        classify(cur, in_mask, 3);
        if (__any_sync(0xffffffffu, defer_mask != 0u)) { n_deferred += (unsigned)__popc(defer_mask); append_deferred(cur); }
        drain();
    }
    // ---- flush ----------------------------------------------------------------------------------------------------
    if (qa_count) finalize_a(qa_count);
    if (qb_count) finalize_b(qb_count);
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
