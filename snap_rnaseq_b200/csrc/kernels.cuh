// kernels.cuh -- the __global__ entry points.  One warp per work item, persistent grid, dynamic work fetch.
#pragma once
#include "paired.cuh"

#ifndef WARPS_PER_CTA
#define WARPS_PER_CTA 8
#endif
#define CTA_THREADS (WARPS_PER_CTA * 32)

struct DevBatch {  // a snapb200_read_batch resident in HBM
    const uint32_t *offsets;
    const uint8_t *bases, *quals;
    uint32_t n;
};

struct Counters {  // device-side counters read back after each launch group
    uint32_t work;        // next work item
    uint32_t n_retry;     // items that overflowed the small scratch tier
    uint32_t n_fallback;  // pairs that need the single-end fallback
    uint32_t n_fix;       // mapq fix-up requests
    uint32_t n_limit;     // items that hit a reference pool limit
    uint32_t pad[3];
};

__device__ __forceinline__ uint32_t round8(uint32_t x) { return (x + 7u) & ~7u; }
// the L buffer: the triangular table of the warp mode (961 cells) or, in kernels that use it, the rolling row pairs of
// the lane mode (lane_roll_cells(lane_k) cells x 32 lanes)
__host__ __device__ inline size_t lv_shared_bytes(int lane_k = 0)
{
    size_t lane_bytes = lane_k > 0 ? (size_t)lane_roll_cells(lane_k) * 32 * sizeof(lane_cell_t) : 0;
    size_t bytes = (size_t)LV_CELLS * 2 > lane_bytes ? (size_t)LV_CELLS * 2 : lane_bytes;
    return (bytes + 15) & ~(size_t)15;
}

__device__ __forceinline__ uint32_t fetch_work(uint32_t *counter)
{
    uint32_t w = 0;
    if (lane_id() == 0) w = atomicAdd(counter, 1u);
    return __shfl_sync(FULL_MASK, w, 0);
}

// ---- single end ---------------------------------------------------------------------------------------
// Work item p in [0, n_items): result slot = positions ? positions[p] : p; read = items ? items[slot] : slot, where a
// read id is idx*2+mate when two batches are given (fallback of pairs) and idx otherwise.
struct SingleArgs {
    int ix_slot;  // c_index[] entry of the index
    SingleCfg cfg;
    DevBatch b[2];
    int two_batches;
    const uint32_t *positions, *items;
    uint32_t n_items;
    snapb200_single_result *results;
    int32_t *mh_counts; uint32_t *mh_locs; uint8_t *mh_rcs; int32_t *mh_scores;
    // scratch, laid out [warp slot][...]
    Elem *pool; int2 *anchors; int *lists; uint32_t *epochs; uint32_t *hit_count, *hit_loc; uint8_t *hit_rc;
    Counters *ctr; uint32_t *retry_list; MapqFix *fix; uint32_t fix_cap;
    unsigned long long *stats;
    int mapq_divisor;
};

__host__ __device__ inline size_t single_warp_shared(uint32_t rl)
{
    size_t s = (sizeof(SingleSm) + 15) & ~(size_t)15;
    s += lv_shared_bytes();
    s += 4 * (size_t)rl;                       // D[2], Q[2]
    s += ((size_t)rl + 2 * WIN_SLACK + 15) & ~(size_t)15;  // W
    return s;
}

__global__ void __launch_bounds__(CTA_THREADS, 24 / WARPS_PER_CTA) single_kernel(const SingleArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const size_t per_warp = single_warp_shared(a.cfg.rl);
    uint8_t *base = smem + per_warp * warp;
    SingleSm *sm = (SingleSm *)base;
    base += (sizeof(SingleSm) + 15) & ~(size_t)15;
    int16_t *L = (int16_t *)base;
    base += lv_shared_bytes();
    ReadView v;
    v.base = base; v.rl = a.cfg.rl; v.len = 0;
    uint8_t *W = base + 4 * a.cfg.rl;
    const uint32_t slot = blockIdx.x * WARPS_PER_CTA + warp;
    SingleScratch sc;
    sc.pool = a.pool + (size_t)slot * a.cfg.pool_cap;
    sc.anchors[0] = a.anchors + (size_t)slot * 2 * (a.cfg.tmask + 1);
    sc.anchors[1] = sc.anchors[0] + (a.cfg.tmask + 1);
    sc.list_head = a.lists + (size_t)slot * 2 * a.cfg.n_lists;
    sc.list_tail = sc.list_head + a.cfg.n_lists;
    sc.epoch = a.epochs + slot;
    sc.hit_count = a.hit_count ? a.hit_count + (size_t)slot * MAXK : nullptr;
    sc.hit_loc = a.hit_loc ? a.hit_loc + (size_t)slot * MAXK * 512 : nullptr;
    sc.hit_rc = a.hit_rc ? a.hit_rc + (size_t)slot * MAXK * 512 : nullptr;
    MapqFixList fix = {a.fix, &a.ctr->n_fix, a.fix_cap};
    #pragma unroll 1
    for (;;) {
        const uint32_t p = fetch_work(&a.ctr->work);
        if (p >= a.n_items) break;
        const uint32_t rslot = a.positions ? a.positions[p] : p;
        const uint32_t id = a.items ? a.items[rslot] : rslot;
        const DevBatch &b = a.two_batches ? a.b[id & 1] : a.b[0];
        const uint32_t ridx = a.two_batches ? id >> 1 : id;
        const uint32_t off = b.offsets[ridx], len = b.offsets[ridx + 1] - off;
        const uint32_t mh = a.cfg.max_hits_to_get;
        bool ok = single_align_warp(a.ix_slot, a.cfg, sc, sm, v, W, L, b.bases + off, b.quals + off, len, rslot, fix, a.mapq_divisor,
                                    mh ? a.mh_counts + rslot : nullptr, mh ? a.mh_locs + (size_t)rslot * mh : nullptr,
                                    mh ? a.mh_rcs + (size_t)rslot * mh : nullptr, mh ? a.mh_scores + (size_t)rslot * mh : nullptr,
                                    a.stats + 11);
        if (lane == 0) {
            snapb200_single_result *r = &a.results[rslot];
            if (!ok) {
                r->status = STATUS_RETRY;
                a.retry_list[atomicAdd(&a.ctr->n_retry, 1u)] = rslot;
            } else {
                r->location = sm->out_loc;
                r->score = sm->out_score;
                r->mapq = sm->out_mapq;
                r->status = (uint8_t)sm->out_status;
                r->direction = (uint8_t)sm->out_dir;
                r->popular_seeds_skipped = (uint16_t)sm->popular;
                r->n_lookups = sm->n_lookups;
                r->n_scored = sm->n_scored;
                r->p_all = sm->p_all;
                r->p_best = sm->p_best;
                atomicAdd(a.stats + 8, (unsigned long long)sm->n_lookups);
                atomicAdd(a.stats + 9, (unsigned long long)sm->n_scored);
                atomicAdd(a.stats + 10, (unsigned long long)sm->popular);
                atomicAdd(a.stats + 12, (unsigned long long)sm->n_probes);
                atomicAdd(a.stats + 13, (unsigned long long)sm->n_hit_words);
            }
        }
        __syncwarp();
    }
}

// ---- paired end ------------------------------------------------------------------------------------------
struct PairedArgs {
    int ix_slot;  // c_index[] entry of the index
    PairedCfg cfg;
    DevBatch b[2];
    const uint32_t *positions;  // retry list or null
    uint32_t n_items;
    snapb200_paired_result *results;
    uint32_t force_spacing;
    Cand *cands; Mate *mates; Anchor *anchors; lane_cell_t *lane_tables; uint32_t *order;  // [warp slot][...]
    Counters *ctr; uint32_t *retry_list, *fallback_list; MapqFix *fix; uint32_t fix_cap;
    unsigned long long *stats;
    unsigned long long *prof;  // optional cycle accounting [8] (builds with -DSNAPB200_PROFILE)
    uint32_t smem_per_warp;    // paired_warp_shared(cfg.rl, cfg.lane_k), computed on the host
};

__host__ __device__ inline size_t paired_warp_shared(uint32_t rl, uint32_t lane_k)
{
    size_t s = (sizeof(PairedSm) + 15) & ~(size_t)15;
    s += lv_shared_bytes((int)lane_k);
    s += 8 * (size_t)rl;
    s += ((size_t)rl + 2 * WIN_SLACK + 15) & ~(size_t)15;
    return s;
}

#ifndef PAIRED_WARPS_PER_SM
// Measured on C3 (scripts/ab_variants.py, profiles/r2_ab_*.log): 16 / 18 / 20 / 24 / 28 / 32 warps per SM = 73.3 / 69.8 / 66.9 / 67.8 / 69.0 /
// 79.8 ms per million pairs.  The kernel is instruction-fetch and L1 bound: beyond 24 warps the registers drop to 64 (spills) and the
// shared memory of the extra warps comes out of the L1 that the per-warp scratch in HBM is served from.
#define PAIRED_WARPS_PER_SM 24
#endif
__global__ void __launch_bounds__(CTA_THREADS, PAIRED_WARPS_PER_SM / WARPS_PER_CTA) paired_kernel(const PairedArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    uint8_t *base = smem + a.smem_per_warp * (uint32_t)warp;
    PairedSm *sm = (PairedSm *)base;
    base += (sizeof(PairedSm) + 15) & ~(size_t)15;
    int16_t *L = (int16_t *)base;
    base += lv_shared_bytes((int)a.cfg.lane_k);
    uint8_t *const rbase = base;  // mate w: four arrays of rl bytes at rbase + 4*w*rl
    uint8_t *W = base + 8 * a.cfg.rl;
    const uint32_t slot = blockIdx.x * WARPS_PER_CTA + warp;
    PairedScratch sc;
    sc.cands = a.cands + (size_t)slot * a.cfg.cand_cap;
    sc.mates = a.mates + (size_t)slot * 2 * a.cfg.mate_cap;
    sc.mate_cap = a.cfg.mate_cap;
    sc.anchors = a.anchors + (size_t)slot * a.cfg.anchor_cap;
    sc.lane_table = a.lane_tables + (size_t)slot * lane_table_cells((int)a.cfg.lane_k) * 32;
    sc.order = a.order + (size_t)slot * a.cfg.cand_cap;
    MapqFixList fix = {a.fix, &a.ctr->n_fix, a.fix_cap};
    if (lane < 2) sm->sc_key[lane] = 0xffffffffu;  // no cached seed schedule yet
    if (lane < 5) sm->acc[lane] = 0;  // the run counters are summed per warp and flushed once (five global atomics per pair otherwise)
    __syncwarp();
    PROF(if (lane == 0 && a.prof) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); atomicMin(a.prof + 11, t & 0xffffffffffull); })
    #pragma unroll 1
    for (;;) {
        const uint32_t p = fetch_work(&a.ctr->work);
        if (p >= a.n_items) {
            // when this warp found the queue empty: the earliest such time is where the kernel's tail begins, the latest is its end
            PROF(if (lane == 0 && a.prof) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); t &= 0xffffffffffull; atomicMin(a.prof + 12, t); atomicMax(a.prof + 13, t); atomicAdd(a.prof + 14, t); })
            break;
        }
        const uint32_t pi = a.positions ? a.positions[p] : p;
        snapb200_paired_result *r = &a.results[pi];
        const uint32_t off0 = a.b[0].offsets[pi], len0 = a.b[0].offsets[pi + 1] - off0;
        const uint32_t off1 = a.b[1].offsets[pi], len1 = a.b[1].offsets[pi + 1] - off1;
        if (lane == 0) {  // ChimericPairedEndAligner::align prologue (:74-80); untouched fields read as zero
            r->location[0] = r->location[1] = INVALID_LOC;
            r->score[0] = r->score[1] = 0; r->mapq[0] = r->mapq[1] = 0;
            r->status[0] = r->status[1] = SNAPB200_NOT_FOUND;
            r->direction[0] = r->direction[1] = 0;
            r->from_align_together = 0; r->aligned_as_pair = 0; r->pad = 0;
            r->n_lv_calls = 0; r->n_lookups = 0; r->p_all = 0; r->p_best = 0;
        }
        __syncwarp();
        if (len0 < 50 && len1 < 50) continue;
        uint32_t n_bad0, n_bad1;
        const ReadView v0 = {rbase, a.cfg.rl, len0}, v1 = {rbase + 4 * a.cfg.rl, a.cfg.rl, len1};
        uint32_t ns = stage_read(v0, a.b[0].bases + off0, a.b[0].quals + off0, &n_bad0);
        ns += stage_read(v1, a.b[1].bases + off1, a.b[1].quals + off1, &n_bad1);
        PROF(if (lane == 0) for (int q = 0; q < 12; q++) sm->t_phase[q] = 0; long long t_s = clock64();)
        int rc = paired_intersect_warp(a.ix_slot, a.cfg, sc, sm, rbase, len0, len1, ns, n_bad0, n_bad1, W, L, r, pi, fix);
        PROF(if (lane == 0 && a.prof) {
            atomicAdd(a.prof + 0, (unsigned long long)(clock64() - t_s));
            for (int q = 1; q < 5; q++) atomicAdd(a.prof + q, (unsigned long long)sm->t_phase[q]);
            atomicAdd(a.prof + 5, 1ull);
            for (int q = 5; q < 10; q++) atomicAdd(a.prof + q + 1, (unsigned long long)sm->t_phase[q]);
        })
        if (lane == 0) {
            if (rc == 2) {
                if (a.cfg.hard_limit) {
                    r->status[0] = r->status[1] = STATUS_LIMIT;
                    atomicAdd(&a.ctr->n_limit, 1u);
                } else {
                    r->status[0] = r->status[1] = STATUS_RETRY;
                    a.retry_list[atomicAdd(&a.ctr->n_retry, 1u)] = pi;
                }
            } else {
                if (rc == 1) {
                    r->n_lv_calls = sm->n_lv;
                    r->n_lookups = sm->n_look[0] + sm->n_look[1];
                    sm->acc[0] += r->n_lookups; sm->acc[1] += sm->n_lv; sm->acc[2] += sm->popular[0] + sm->popular[1];
                    sm->acc[3] += sm->n_probes; sm->acc[4] += sm->n_hit_words;
                }
                r->from_align_together = 1;
                r->aligned_as_pair = 1;
                bool fallback;
                if (a.force_spacing) {  // ChimericPairedEndAligner.cpp:92-99
                    if (r->status[0] == SNAPB200_NOT_FOUND) r->from_align_together = 0;
                    fallback = false;
                } else {
                    fallback = r->status[0] == SNAPB200_NOT_FOUND || r->status[1] == SNAPB200_NOT_FOUND;
                }
                if (fallback) a.fallback_list[atomicAdd(&a.ctr->n_fallback, 1u)] = pi;
            }
        }
        __syncwarp();
    }
    __syncwarp();
    if (lane < 5 && sm->acc[lane]) {
        // words 8, 9, 10, 12, 13: n_hash_table_lookups, n_locations_scored, n_hits_ignored_popularity, n_table_probes, n_hit_words_read
        atomicAdd(a.stats + (lane < 3 ? 8 + lane : 9 + lane), sm->acc[lane]);
    }
}

// ---- work ordering for the paired path ------------------------------------------------------------------------------------
// Pairs differ in cost by four orders of magnitude (median 2 scored locations, maximum several thousand: reads from repeat
// families), and a warp keeps one pair until it is done.  Served in input order, the heavy pairs that happen to come late
// leave most SMs idle at the end of a launch, and heavy and ordinary pairs executing side by side on an SM compete for the
// instruction cache.  So before the aligner runs, this kernel estimates each pair's weight from eight index probes (four seeds spread over
// each mate) and the host sorts the pairs heaviest first (stable: ordinary pairs stay in input order).
// The order changes nothing in the results: pairs are independent and every pair writes only its own record.
// Thread t: item t / (8 * MATES), mate (t>>3) & (MATES-1), seed number t&7 of eight spread evenly over the read.  MATES = 2
// for pairs, 1 for the single-end aligner (whose heavy reads vote and score thousands of locations just the same).
#define WEIGH_SEEDS 8  // per read
template <int MATES>
__global__ void weigh_kernel(const DevIndex ix, const DevBatch b0, const DevBatch b1, uint32_t n, uint32_t max_hits,
                             uint32_t *keys, uint32_t *vals)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t item = t / (WEIGH_SEEDS * MATES);
    uint32_t w = 0;
    if (item < n) {
        const DevBatch &b = (MATES == 2 && ((t >> 3) & 1)) ? b1 : b0;
        const uint32_t off = b.offsets[item], len = b.offsets[item + 1] - off;
        if (len >= ix.seed_len) {
            const uint8_t *seed = b.bases + off + (size_t)(len - ix.seed_len) * (t & 7) / (WEIGH_SEEDS - 1);
            uint64_t f, r;
            if (pack_seed(seed, ix.seed_len, &f, &r)) {
                HitList hl[2];
                lookup_seed(ix, f, r, hl, nullptr);
                w = min(hl[0].n, max_hits) + min(hl[1].n, max_hits);
            }
        }
    }
    w += __shfl_xor_sync(FULL_MASK, w, 1);
    w += __shfl_xor_sync(FULL_MASK, w, 2);
    w += __shfl_xor_sync(FULL_MASK, w, 4);
    if (MATES == 2) w += __shfl_xor_sync(FULL_MASK, w, 8);
    if (item < n && t % (WEIGH_SEEDS * MATES) == 0) {
        // 16-bit sort key, ascending = heaviest first; everything light shares the last key and keeps its input order
        const uint32_t q = w >> (MATES == 2 ? 5 : 4);
        keys[item] = q == 0 ? 0xffffu : 0xfffeu - min(q, 0xfffeu);
        vals[item] = item;
    }
}

// Fold the single-end fallback results into the pair records (ChimericPairedEndAligner.cpp:110-119):
// thread t handles fallback entry t>>1, end t&1.
__global__ void merge_fallback_kernel(const uint32_t *fallback_list, uint32_t n_fallback, const snapb200_single_result *sr,
                                      snapb200_paired_result *results)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n_fallback) return;
    uint32_t pi = fallback_list[t >> 1], e = t & 1;
    const snapb200_single_result &s = sr[pi * 2 + e];
    snapb200_paired_result *r = &results[pi];
    r->status[e] = s.status;
    r->location[e] = s.location;
    r->direction[e] = s.direction;
    r->score[e] = s.score;
    r->mapq[e] = s.mapq / 4;  // heavy quality penalty for chimeric reads
    if (e == 0) { r->from_align_together = 0; r->aligned_as_pair = 0; }
}

// status / mapq histogram of the final records (one thread per read).  Counted in shared memory per block and flushed with
// one global atomic per non-zero counter: a global atomic per read on the same handful of addresses took 3 ms per million
// pairs (ncu launch list), 4 % of a step.
__device__ __forceinline__ void stats_count(unsigned int *blk, bool valid, uint8_t st, int mapq, bool as_pair)
{
    if (valid) {
        atomicAdd(&blk[0], 1u);
        atomicAdd(&blk[st == SNAPB200_SINGLE_HIT ? 2 : st == SNAPB200_MULTIPLE_HITS ? 3 : 4], 1u);
        if (as_pair) atomicAdd(&blk[6], 1u);
        if (st != SNAPB200_NOT_FOUND && mapq >= 0 && mapq <= 70) atomicAdd(&blk[14 + mapq], 1u);
    }
}
__global__ void stats_single_kernel(const snapb200_single_result *r, uint32_t n, unsigned long long *stats)
{
    __shared__ unsigned int blk[SNAPB200_STATS_WORDS];
    for (int i = threadIdx.x; i < SNAPB200_STATS_WORDS; i += blockDim.x) blk[i] = 0;
    __syncthreads();
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t < n;
    stats_count(blk, valid, valid ? r[t].status : 0, valid ? r[t].mapq : 0, false);
    __syncthreads();
    for (int i = threadIdx.x; i < SNAPB200_STATS_WORDS; i += blockDim.x) if (blk[i]) atomicAdd(stats + i, (unsigned long long)blk[i]);
}
__global__ void stats_paired_kernel(const snapb200_paired_result *r, uint32_t n, unsigned long long *stats)
{
    __shared__ unsigned int blk[SNAPB200_STATS_WORDS];
    for (int i = threadIdx.x; i < SNAPB200_STATS_WORDS; i += blockDim.x) blk[i] = 0;
    __syncthreads();
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t < 2 * n;
    const int e = t & 1;
    stats_count(blk, valid, valid ? r[t >> 1].status[e] : 0, valid ? r[t >> 1].mapq[e] : 0, valid && r[t >> 1].aligned_as_pair);
    __syncthreads();
    for (int i = threadIdx.x; i < SNAPB200_STATS_WORDS; i += blockDim.x) if (blk[i]) atomicAdd(stats + i, (unsigned long long)blk[i]);
}

// ---- CIGAR against the resident genome (SAM.cpp:1159-1189) ---------------------------------------------------
struct CigarArgs {
    DevIndex ix;
    DevBatch b;
    const uint32_t *locations;
    const uint8_t *directions;
    int use_m;
    char *cigars;
    uint32_t stride;
    int32_t *edit_distance;
    uint32_t rl;
    Counters *ctr;
};

__host__ __device__ inline size_t cigar_warp_shared(uint32_t rl)
{
    return lv_shared_bytes() + (size_t)rl + (((size_t)rl + 2 * WIN_SLACK + 15) & ~(size_t)15);
}

__global__ void __launch_bounds__(CTA_THREADS) cigar_kernel(const CigarArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    uint8_t *base = smem + cigar_warp_shared(a.rl) * warp;
    int16_t *L = (int16_t *)base;
    uint8_t *P = base + lv_shared_bytes();
    uint8_t *W = P + a.rl;
    #pragma unroll 1
    for (;;) {
        const uint32_t i = fetch_work(&a.ctr->work);
        if (i >= a.b.n) break;
        const uint32_t off = a.b.offsets[i], len = a.b.offsets[i + 1] - off;
        const uint32_t loc = a.locations[i];
        char *out = a.cigars + (size_t)i * a.stride;
        #pragma unroll 1
        for (uint32_t j = lane; j < a.stride; j += 32) out[j] = 0;
        __syncwarp();
        if (loc == INVALID_LOC || !substring_ok(a.ix, loc, len)) {
            if (lane == 0) a.edit_distance[i] = -3;
            continue;
        }
        const bool rc = a.directions[i] == SNAPB200_RC;
        #pragma unroll 1
        for (uint32_t j = lane; j < len; j += 32) P[j] = rc ? rc_base(a.b.bases[off + len - 1 - j]) : a.b.bases[off + j];
        stage_window(a.ix, loc, len, W);
        LvStr s;
        s.p = P; s.ps = 1; s.plen = (int)len;
        s.t = W + WIN_SLACK; s.ts = 1; s.tlen = (int)len;
        s.t_lo = -WIN_SLACK; s.t_hi = (int)len + WIN_SLACK;
        int e = lv_cigar_warp(s, MAXK - 1, L, out, (int)a.stride, a.use_m != 0);
        if (lane == 0) a.edit_distance[i] = e;
        __syncwarp();
    }
}

// ---- building blocks on explicit strings (known-answer tests) ----------------------------------------------
struct LvArgs {
    int ix_slot;  // c_index[] entry; only the probability tables are used
    int dir;
    uint32_t n;
    const uint32_t *text_off, *pat_off;
    const uint8_t *texts, *pats, *quals;
    const int32_t *k;
    int32_t *score, *indel;
    double *prob;
    int use_m; char *cigars; uint32_t stride;  // CIGAR variant when cigars != null
    uint32_t max_text, max_pat;
    Counters *ctr;
};

__host__ __device__ inline size_t lvtest_warp_shared(uint32_t max_text, uint32_t max_pat)
{
    return lv_shared_bytes() + ((max_text + 15) & ~15u) + 2 * (size_t)((max_pat + 15) & ~15u);
}

__global__ void __launch_bounds__(CTA_THREADS) lv_kernel(const LvArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    uint8_t *base = smem + lvtest_warp_shared(a.max_text, a.max_pat) * warp;
    int16_t *L = (int16_t *)base;
    uint8_t *T = base + lv_shared_bytes();
    uint8_t *P = T + ((a.max_text + 15) & ~15u);
    uint8_t *Q = P + ((a.max_pat + 15) & ~15u);
    #pragma unroll 1
    for (;;) {
        const uint32_t i = fetch_work(&a.ctr->work);
        if (i >= a.n) break;
        const uint32_t to = a.text_off[i], tl = a.text_off[i + 1] - to, po = a.pat_off[i], pl = a.pat_off[i + 1] - po;
        #pragma unroll 1
        for (uint32_t j = lane; j < tl; j += 32) T[j] = a.texts[to + j];
        #pragma unroll 1
        for (uint32_t j = lane; j < pl; j += 32) { P[j] = a.pats[po + j]; if (a.quals) Q[j] = a.quals[po + j]; }
        __syncwarp();
        LvStr s;
        s.p = P; s.ps = 1; s.plen = (int)pl;
        s.tlen = (int)tl; s.t_lo = 0; s.t_hi = (int)tl;
        if (a.dir > 0) { s.t = T; s.ts = 1; } else { s.t = T + tl - 1; s.ts = -1; }  // backward: text(i) = T[tl-1-i]
        if (a.cigars) {
            char *out = a.cigars + (size_t)i * a.stride;
            #pragma unroll 1
            for (uint32_t j = lane; j < a.stride; j += 32) out[j] = 0;
            __syncwarp();
            int e = lv_cigar_warp(s, a.k[i], L, out, (int)a.stride, a.use_m != 0);
            if (lane == 0) a.score[i] = e;
        } else {
            double prob;
            int indel;
            int e = lv_score_warp(s, a.quals ? Q : nullptr, 1, a.k[i], a.ix_slot, L, &prob, &indel);
            if (lane == 0) { a.score[i] = e; a.prob[i] = prob; a.indel[i] = indel; }
        }
        __syncwarp();
    }
}

__global__ void lookup_kernel(const DevIndex ix, uint32_t n, const uint8_t *seeds, uint32_t max_out, uint32_t *n_hits, uint32_t *hits)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t f, r;
    HitList hl[2] = {{nullptr, 0}, {nullptr, 0}};
    if (pack_seed(seeds + (size_t)i * ix.seed_len, ix.seed_len, &f, &r)) lookup_seed(ix, f, r, hl, nullptr);
    #pragma unroll 1
    for (int d = 0; d < 2; d++) {
        n_hits[i * 2 + d] = hl[d].n;
        #pragma unroll 1
        for (uint32_t j = 0; j < hl[d].n && j < max_out; j++) hits[((size_t)i * 2 + d) * max_out + j] = hl[d].hits[j];
    }
}

__global__ void mapq_kernel(uint32_t n, const double *p_all, const double *p_best, const int32_t *score, const int32_t *popular,
                            int32_t *mapq, uint8_t *near_integer)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool ni;
    mapq[i] = compute_mapq_dev(p_all[i], p_best[i], score[i], popular[i], &ni);
    near_integer[i] = ni;
}

// ---- stage-2 roofline diagnostics -----------------------------------------------------------------------------------
// setup (not timed): pack the seed at each genome position the way stage 1 does from the staged read
__global__ void probe_pack_kernel(const DevIndex ix, uint32_t n, const uint32_t *positions, ulonglong2 *packed)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t f = 0, r = 0;
    bool ok = pack_seed(ix.genome + positions[i], ix.seed_len, &f, &r);
    packed[i] = ok ? make_ulonglong2(f, r) : make_ulonglong2(~0ull, ~0ull);
}

// timed: stage 2 alone -- one lane per packed seed, both directions resolved, overflow count words read
__global__ void probe_bench_kernel(const DevIndex ix, uint32_t n, const ulonglong2 *packed, unsigned long long *totals)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t probes = 0, counts = 0, hits = 0;
    if (i < n) {
        ulonglong2 s = packed[i];
        if (s.x != ~0ull) {
            HitList hl[2] = {{nullptr, 0}, {nullptr, 0}};
            lookup_seed(ix, s.x, s.y, hl, &probes);
            counts = (hl[0].n > 1) + (hl[1].n > 1 && s.x != s.y);
            hits = hl[0].n + hl[1].n;
        }
    }
    // block-level totals: same-address atomics are ~1 per cycle at L2, so one set per block, not per warp
    __shared__ unsigned int blk[3];
    if (threadIdx.x < 3) blk[threadIdx.x] = 0;
    __syncthreads();
    probes = __reduce_add_sync(FULL_MASK, probes);
    counts = __reduce_add_sync(FULL_MASK, counts);
    hits = __reduce_add_sync(FULL_MASK, hits);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&blk[0], probes); atomicAdd(&blk[1], counts); atomicAdd(&blk[2], hits); }
    __syncthreads();
    if (threadIdx.x < 3) atomicAdd(totals + threadIdx.x, (unsigned long long)blk[threadIdx.x]);
}

__global__ void gather_bench_kernel(const uint4 *buf, unsigned long long n_sectors, uint32_t n, uint32_t salt, unsigned long long *sink)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // one uniformly random 32-byte sector per thread (two 16-byte loads of the same sector)
    unsigned long long h = ((unsigned long long)i + 1) * 0x9E3779B97F4A7C15ull + (unsigned long long)salt * 0xD1B54A32D192ED03ull;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    unsigned long long sector = h % n_sectors;
    uint4 a = __ldg(buf + sector * 2), b = __ldg(buf + sector * 2 + 1);
    unsigned x = a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
    if (x == 0x12345678u) atomicAdd(sink, 1ull);  // keeps the loads alive
}
