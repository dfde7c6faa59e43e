// single.cuh -- single-end seed-and-vote aligner, one warp per read.
//
// Replaces BaseAligner::AlignRead / score / findCandidate / allocateNewCandidate / incrementWeight /
// fillHitsFound (SNAPLib/BaseAligner.cpp:510-1568, 1679-1727).  Outputs are bit-identical to the reference:
// the order in which seeds are looked up, hits are voted, weight lists are served and candidates are scored
// is the reference's; what changes is where the time goes:
//   warp sections   -- read staging (fwd + reverse complement) in shared memory, up to 32 index probes in
//                      flight per warp, coalesced hit-list loads, 32 candidate-bucket probes at once, genome
//                      window staging, Landau-Vishkin rows across lanes;
//   leader sections -- the weight-list / merge bookkeeping, which is a dependent chain by definition.
#pragma once
#include "common.cuh"
#include "lookup.cuh"
#include "lv.cuh"

struct __align__(16) Elem {  // BaseAligner::HashTableElement (BaseAligner.h:198-226), 160 bytes instead of 464
    uint64_t used, scored;
    double best_prob;
    uint32_t base, weight, lowest_possible, best_score, best_loc;
    int32_t w_next, w_prev, h_next;
    uint8_t dir, all_scored;
    uint16_t pad;
    uint16_t seed_offset[BUCKET];
};

struct SingleCfg {  // kernel-wide configuration derived from snapb200_single_params
    uint32_t max_hits, max_k, num_seeds, extra, explore, stop_first, max_hits_to_get;
    double seed_coverage;
    uint32_t n_lists;      // maxSeedsToUse(ctor)+1, BaseAligner.cpp:127
    uint32_t pool_cap;     // elements per warp in this scratch tier
    uint32_t tmask;        // candidate table size - 1 (power of two)
    uint32_t rl;           // shared-memory stride per read orientation (>= longest read in the batch)
};

struct SingleScratch {  // per-warp slices of global scratch
    Elem *pool;
    int2 *anchors[2];   // {head element, epoch}
    int *list_head, *list_tail;
    uint32_t *epoch;    // persists across reads (BaseAligner::hashTableEpoch)
    uint32_t *hit_count, *hit_loc;
    uint8_t *hit_rc;
};

struct SingleSm {  // per-warp shared state: everything leader and warp sections exchange
    double p_all, p_best, prob;
    unsigned long long cand_mask;
    uint32_t lowest_unseen[2], n_applied[2];
    uint32_t most_seeds, best_score, best_loc, second_best, second_loc, score_limit, popular;
    uint32_t n_used, highest_list, n_lookups, n_scored, epoch, n_probes, n_hit_words;
    uint32_t out_loc;
    int out_score, out_mapq, out_dir, out_status;
    int action, cand_elem, cand_dir, cand_any_nearby, sc, loc_off, overflow, force, list, alloc_in_chunk;
    uint32_t cand_loc, cand_seedoff, cand_idx;
    uint32_t next, wrap, n_sched, terminal;
    uint32_t used[16];               // seedUsed bitmap (BaseAligner.h:177-186), 512 bits
    uint16_t sched_off[32], sched_wrap[32];
    uint32_t hit_locs[32];
    int hit_elem[32];
};

struct ReadView {  // a read staged in shared memory in both orientations: four arrays of `rl` bytes back to back
    uint8_t *base;
    uint32_t rl, len;
    // D(FORWARD) = read, D(RC) = reverse complement (BaseAligner.cpp:638-650); Q(RC) = reversed quality.  Computed, not
    // stored: a view is three registers, is passed by value and never has to live in local memory.
    __device__ __forceinline__ uint8_t *D(int dir) const { return base + (uint32_t)dir * rl; }
    __device__ __forceinline__ uint8_t *Q(int dir) const { return base + (2u + (uint32_t)dir) * rl; }
};

// all lanes; returns the number of 'N' bases; *n_not_acgt (optional): bases that cannot be part of a seed
__device__ __forceinline__ uint32_t stage_read(const ReadView v, const uint8_t *bases, const uint8_t *quals, uint32_t *n_not_acgt = nullptr)
{
    const int lane = lane_id();
    uint32_t ns = 0, bad = 0;
    #pragma unroll 1
    for (uint32_t base = 0; base < v.len; base += 32) {
        uint32_t i = base + lane;
        bool is_n = false, is_bad = false;
        if (i < v.len) {
            uint8_t b = bases[i], q = quals[i];
            v.D(0)[i] = b;
            v.Q(0)[i] = q;
            v.D(1)[v.len - 1 - i] = rc_base(b);
            v.Q(1)[v.len - 1 - i] = q;
            is_n = b == 'N';
            is_bad = base2(b) < 0;
        }
        ns += __popc(__ballot_sync(FULL_MASK, is_n));
        bad += __popc(__ballot_sync(FULL_MASK, is_bad));
    }
    __syncwarp();
    if (n_not_acgt) *n_not_acgt = bad;
    return ns;
}

// Stage genome[loc-WIN_SLACK, loc+rlen+WIN_SLACK) into W.  Bytes the reference could not legally read
// (outside [-100, nBases+100), Genome.h:175) become 0x01, which matches nothing.
__device__ __forceinline__ void stage_window(const DevIndex &ix, uint32_t loc, uint32_t rlen, uint8_t *W)
{
    const int lane = lane_id();
    const int n = (int)rlen + 2 * WIN_SLACK;
    const long long g0 = (long long)loc - WIN_SLACK;
    #pragma unroll 1
    for (int j = lane; j < n; j += 32) {
        long long g = g0 + j;
        W[j] = (g >= -100 && g < (long long)ix.n_bases + 100) ? __ldg(ix.genome + g) : (uint8_t)0x01;
    }
    __syncwarp();
}

__device__ __forceinline__ bool substring_ok(const DevIndex &ix, uint32_t offset, uint32_t len)
{  // Genome::getSubstring for lengthNeeded <= chromosomePadding (Genome.h:78-86)
    return !((uint64_t)offset > ix.n_bases || (uint64_t)offset + len > (uint64_t)ix.n_bases + 100);
}

// The scoring step shared by BaseAligner::score (BaseAligner.cpp:1158-1242) and
// IntersectingPairedEndAligner::scoreLocation (:755-841).  All lanes; uniform results.
__device__ __noinline__ int score_location_warp(int ix_slot, const ReadView v, int dir, uint32_t loc, uint32_t seed_offset,
                                   int score_limit, bool single_variant, uint8_t *W, int16_t *L, double *match_prob,
                                   int *loc_offset)
{
    const DevIndex &ix = c_index[ix_slot];
    const uint32_t rlen = v.len;
    uint32_t glen = rlen + MAXK;
    bool have = substring_ok(ix, loc, glen);
    *match_prob = 0;
    *loc_offset = 0;
    if (!have) {  // up against the end of the genome / a contig
        uint32_t end_off;
        if ((uint64_t)loc + rlen + MAXK >= ix.n_bases) {
            end_off = ix.n_bases;
        } else if (single_variant) {  // getNextPieceAfterLocation(loc)
            end_off = ix.n_bases;
            #pragma unroll 1
            for (uint32_t i = 0; i < ix.n_pieces; i++) if (ix.piece_begin[i] > loc) { end_off = ix.piece_begin[i]; break; }
        } else {  // getPieceAtLocation(loc + rlen + MAX_K)
            end_off = 0;
            #pragma unroll 1
            for (uint32_t i = 0; i < ix.n_pieces; i++) if (ix.piece_begin[i] <= loc + rlen + MAXK) end_off = ix.piece_begin[i];
        }
        glen = end_off - loc - 1;
        if (glen >= rlen - MAXK) have = substring_ok(ix, loc, glen);
        if (!have) return -1;
    }
    stage_window(ix, loc, rlen, W);
    const int seed_len = (int)ix.seed_len;
    const int tail = (int)seed_offset + seed_len;
    const int wn = (int)rlen + 2 * WIN_SLACK;
    LvStr s;
    double p1, p2;
    int dummy;
    // forward: read tail against the genome after the seed
    s.p = v.D(dir) + tail; s.ps = 1; s.plen = (int)rlen - tail;
    s.t = W + WIN_SLACK + tail; s.ts = 1; s.tlen = (int)glen - tail;
    s.t_lo = -(WIN_SLACK + tail); s.t_hi = wn - (WIN_SLACK + tail);
    int s1 = lv_score_warp(s, v.Q(dir) + tail, 1, score_limit, ix_slot, L, &p1, &dummy);
    if (s1 == -1) return -1;
    __syncwarp();
    // backward: read head (reversed) against the genome before the seed
    s.p = v.D(dir) + seed_offset - 1; s.ps = -1; s.plen = (int)seed_offset;
    s.t = W + WIN_SLACK + seed_offset - 1; s.ts = -1; s.tlen = (int)seed_offset + MAXK;
    s.t_lo = (int)seed_offset - (int)rlen - WIN_SLACK; s.t_hi = WIN_SLACK + (int)seed_offset;
    int s2 = lv_score_warp(s, v.Q(dir) + seed_offset - 1, -1, score_limit - s1, ix_slot, L, &p2, loc_offset);
    if (s2 == -1) { *loc_offset = 0; return -1; }
    *match_prob = p1 * p2 * ix.seed_prob;
    return s1 + s2;
}

// The same scoring step for 32 candidates at once, one per lane (see lv_lane in lv.cuh).  v: the read (shared memory);
// dir/loc/seed_offset: this lane's candidate; K <= kl: the score limit (kl: the launch's lane-mode limit); R: this lane's column of the
// interleaved rolling rows (shared) and T its column of the full table (HBM scratch).  Lanes whose genome window is not entirely inside the genome return SC_NONE_LANE and are
// left to score_location_warp.  All 32 lanes must call this together.
#define SC_NONE_LANE (-3)
PROF(__device__ unsigned long long g_prof_lane[4];)  // lane-mode calls, live lanes forward, live lanes backward, lanes that scored
__device__ __noinline__ void score_location_lane(int ix_slot, const ReadView v, int dir, uint32_t loc, uint32_t seed_offset, int K, int kl,
                                    lane_cell_t *R, lane_cell_t *T, bool active, int *score, double *match_prob, int *loc_offset)
{
    const DevIndex &ix = c_index[ix_slot];
    const uint32_t rlen = v.len;
    // same test as getSubstring(loc, rlen + MAX_K) != NULL; the 4-byte loads stay within +-16 bytes of that window
    bool ok = active && substring_ok(ix, loc, rlen + MAXK);
    *score = SC_NONE_LANE;
    *match_prob = 0;
    *loc_offset = 0;
    const int seed_len = (int)ix.seed_len;
    const int tail = (int)seed_offset + seed_len;
    const uint8_t *g = ix.genome + loc;
#ifndef NO_LANE_PREFETCH  // A/B on C3: 67.8 -> 65.0 ms per million pairs
    // every lane asks for the 128-byte lines of its own window up front, so that their HBM latencies overlap instead of being paid one
    // line at a time along the walk (the window is [loc - MAXK, loc + rlen + MAXK): two or three lines)
    if (ok) {
        const uintptr_t lo = ((uintptr_t)g - MAXK) & ~(uintptr_t)127, hi = (uintptr_t)g + rlen + MAXK;
        #pragma unroll 1
        for (uintptr_t a = lo; a < hi; a += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
    }
#endif
    double p1 = 0, p2 = 0;
    int dummy, off = 0;
#ifndef NO_LANE_DUAL  // A/B on C3: 61.9 -> 60.3 ms per million pairs (profiles/r2_ab_9.log)
    // At most 16 locations, all in the lower half of the warp: lane i + 16 runs the backward pass of lane i's location while lane i
    // runs its forward pass -- one trip through the row loop instead of two.  The backward pass gets the whole limit K instead of
    // K - s1; LV finds the same distance d2 (the limit only decides when it gives up), and d2 > K - s1 is turned into -1 below.
    {
        const unsigned okmask = __ballot_sync(FULL_MASK, ok);
        if (okmask != 0 && (okmask >> 16) == 0) {
            const int lane = lane_id(), src = lane & 15;
            const bool upper = lane >= 16;
            const int dir_s = __shfl_sync(FULL_MASK, dir, src), K_s = __shfl_sync(FULL_MASK, K, src);
            const uint32_t loc_s = __shfl_sync(FULL_MASK, loc, src), so_s = __shfl_sync(FULL_MASK, seed_offset, src);
            const bool ok_s = __shfl_sync(FULL_MASK, (int)ok, src) != 0;
            const uint8_t *gs = ix.genome + loc_s;
            const int tail_s = (int)so_s + seed_len;
            const int at = upper ? (int)so_s - 1 : tail_s;
            double pp = 0;
            int oo = 0;
            const int sd = lv_lane_rt(upper ? -1 : 1, v.D(dir_s) + at, upper ? (int)so_s : (int)rlen - tail_s, gs + at, v.Q(dir_s) + at, K_s, kl, R, T, ix_slot, ok_s, &pp, &oo);
            const int sb = __shfl_down_sync(FULL_MASK, sd, 16), ob = __shfl_down_sync(FULL_MASK, oo, 16);
            const double pb = shfl_f64(pp, (lane + 16) & 31);
            if (!ok) return;
            if (sd == -1 || sb == -1 || sb > K - (sd > 0 ? sd : 0)) { *score = -1; return; }
            *score = sd + sb;
            *match_prob = pp * pb * ix.seed_prob;
            *loc_offset = ob;
            return;
        }
    }
#endif
#ifndef LANE_TEMPLATE_DIR  // one run-time-direction body for both passes: 7 KB less code to fetch, 66.4 -> 63.8 ms per million C3 pairs
    int s1 = lv_lane_rt(1, v.D(dir) + tail, (int)rlen - tail, g + tail, v.Q(dir) + tail, K, kl, R, T, ix_slot, ok, &p1, &dummy);
#else
    int s1 = lv_lane<1>(v.D(dir) + tail, (int)rlen - tail, g + tail, v.Q(dir) + tail, K, kl, R, T, ix_slot, ok, &p1, &dummy);
#endif
    const bool ok2 = ok && s1 != -1;
    PROF({ const unsigned a = __ballot_sync(FULL_MASK, ok), b = __ballot_sync(FULL_MASK, ok2);
           if (lane_id() == 0) { atomicAdd(&g_prof_lane[0], 1ull); atomicAdd(&g_prof_lane[1], (unsigned long long)__popc(a)); atomicAdd(&g_prof_lane[2], (unsigned long long)__popc(b)); } })
#ifndef LANE_TEMPLATE_DIR
    int s2 = lv_lane_rt(-1, v.D(dir) + (int)seed_offset - 1, (int)seed_offset, g + (int)seed_offset - 1, v.Q(dir) + (int)seed_offset - 1,
                        K - (s1 > 0 ? s1 : 0), kl, R, T, ix_slot, ok2, &p2, &off);
#else
    int s2 = lv_lane<-1>(v.D(dir) + (int)seed_offset - 1, (int)seed_offset, g + (int)seed_offset - 1, v.Q(dir) + (int)seed_offset - 1,
                         K - (s1 > 0 ? s1 : 0), kl, R, T, ix_slot, ok2, &p2, &off);
#endif
    if (!ok) return;
    if (s1 == -1 || s2 == -1) { *score = -1; return; }
    *score = s1 + s2;
    *match_prob = p1 * p2 * ix.seed_prob;
    *loc_offset = off;
}

// computeMAPQ (SNAPLib/mapq.h:32-65).  *near_integer is set when the log10 value is so close to an integer that
// the truncation could depend on the last ulp of log10; the host then repeats the evaluation with libm.
__device__ __forceinline__ int compute_mapq_dev(double p_all, double p_best, int score, int popular, bool *near_integer)
{
    *near_integer = false;
    if (!(p_all > p_best)) p_all = p_best;
    if (p_all == p_best && popular == 0 && score < 5) return 70;
    double correct = p_best / p_all;
    int base;
    if (correct >= 1) {
        base = 69;
    } else {
        double v = -10 * log10(1 - correct);
        double r = rint(v);
        if (fabs(v - r) <= 1e-9 * fmax(1.0, fabs(v))) *near_integer = true;
        int iv = (v >= 2147483647.0) ? 2147483647 : (int)v;  // +inf when correct rounds 1-correct to 0 is excluded above
        base = iv < 69 ? iv : 69;
    }
    int pen = popular - 10;
    if (pen < 0) pen = 0;
    base -= pen / 2;
    return base > 0 ? base : 0;
}

// ---- candidate table / weight lists (leader lane only unless noted) -------------------------------------
__device__ __forceinline__ uint32_t cand_slot(uint32_t base, uint32_t tmask) { return ((base / BUCKET) * 2654435761u >> 7) & tmask; }

// findElement (BaseAligner.cpp:1415-1442); read-only, callable from any lane
__device__ __forceinline__ int find_element(const SingleScratch &sc, uint32_t tmask, uint32_t epoch, uint32_t loc, int dir)
{
    uint32_t base = loc - loc % BUCKET;
    int2 a = sc.anchors[dir][cand_slot(base, tmask)];
    if ((uint32_t)a.y != epoch) return -1;
    int e = a.x;
    #pragma unroll 1
    while (e >= 0 && sc.pool[e].base != base) e = sc.pool[e].h_next;
    return e;
}

__device__ __forceinline__ void list_unlink(const SingleScratch &sc, int e)
{
    Elem *el = &sc.pool[e];
    if (el->w_next == e && el->w_prev == e) return;  // self-linked = on no list (BaseAligner.cpp:1394)
    uint32_t w = el->weight;
    if (el->w_prev >= 0) sc.pool[el->w_prev].w_next = el->w_next; else sc.list_head[w] = el->w_next;
    if (el->w_next >= 0) sc.pool[el->w_next].w_prev = el->w_prev; else sc.list_tail[w] = el->w_prev;
    el->w_next = el->w_prev = e;
}
__device__ __forceinline__ void list_append(const SingleScratch &sc, int e, uint32_t w)
{
    Elem *el = &sc.pool[e];
    int tail = sc.list_tail[w];
    el->w_next = -1;
    el->w_prev = tail;
    if (tail >= 0) sc.pool[tail].w_next = e; else sc.list_head[w] = e;
    sc.list_tail[w] = e;
}

// leader: one seed hit (the body of the loop at BaseAligner.cpp:844-868)
__device__ __forceinline__ void vote_hit(const SingleCfg &cfg, const SingleScratch &sc, SingleSm *sm, uint32_t loc, int e,
                                         int dir, uint32_t offset)
{
    if (e < 0 && sm->alloc_in_chunk) e = find_element(sc, cfg.tmask, sm->epoch, loc, dir);
    const uint32_t low = loc % BUCKET;
    if (e >= 0) {  // findCandidate + incrementWeight
        Elem *el = &sc.pool[e];
        unsigned long long bit = 1ull << low;
        el->all_scored = el->all_scored && (el->used & bit) != 0;
        el->used |= bit;
        if (!el->all_scored && el->weight < cfg.n_lists - 1) {
            list_unlink(sc, e);
            el->weight++;
            if (el->weight > sm->highest_list) sm->highest_list = el->weight;
            list_append(sc, e, el->weight);
        }
        el->seed_offset[low] = (uint16_t)offset;
    } else if (sm->lowest_unseen[dir] <= sm->score_limit) {  // allocateNewCandidate
        if (sm->n_used >= cfg.pool_cap) { sm->overflow = 1; return; }
        e = (int)sm->n_used++;
        Elem *el = &sc.pool[e];
        uint32_t base = loc - low;
        el->used = 1ull << low;
        el->scored = 0;
        el->lowest_possible = sm->lowest_unseen[dir];
        el->dir = (uint8_t)dir;
        el->weight = 1;
        el->base = base;
        el->best_score = UNUSED_SCORE;
        el->all_scored = 0;
        el->best_prob = 0;
        el->best_loc = 0;
        list_append(sc, e, 1);
        el->seed_offset[low] = (uint16_t)offset;
        if (sm->highest_list < 1) sm->highest_list = 1;
        int2 *a = &sc.anchors[dir][cand_slot(base, cfg.tmask)];
        int2 cur = *a;
        el->h_next = ((uint32_t)cur.y == sm->epoch) ? cur.x : -1;
        *a = make_int2(e, (int)sm->epoch);
        sm->alloc_in_chunk = 1;
    }
}

enum { ACT_NONE = 0, ACT_RETURN_TRUE = 1, ACT_RETURN_FALSE = 2, ACT_ELEMENT = 3 };

// leader: merge bookkeeping after one candidate was scored (BaseAligner.cpp:1253-1384).  Returns true when
// score() must return true immediately (stopOnFirstHit).
__device__ __forceinline__ bool after_score(const SingleCfg &cfg, const SingleScratch &sc, SingleSm *sm)
{
    Elem *el = &sc.pool[sm->cand_elem];
    const int s = sm->sc;
    const uint32_t score = (uint32_t)s;  // -1 -> 0xffffffff as in the reference's unsigned
    const double prob = sm->prob;
    const uint32_t elem_loc = sm->cand_loc;
    uint32_t loc = elem_loc;
    if (s != -1) loc += (uint32_t)sm->loc_off;
    if (cfg.max_hits_to_get > 0 && s != -1 && sc.hit_count[score] < cfg.max_hits_to_get) {
        uint32_t flat = score * 512 + sc.hit_count[score];  // hitLocations[MAX_K][512], BaseAligner.h:149-152
        if (flat < MAXK * 512) { sc.hit_loc[flat] = loc; sc.hit_rc[flat] = el->dir; }
        sc.hit_count[score]++;
    }
    sm->n_scored++;
    bool any_nearby = sm->cand_any_nearby != 0;
    if (any_nearby) {
        if (el->best_score < score || (el->best_score == score && prob <= el->best_prob)) return false;
    }
    el->best_loc = loc;
    int near = -1;
    if (s != -1) {
        const uint32_t half = BUCKET / 2;
        uint32_t near_loc = elem_loc + (2 * (elem_loc % BUCKET / half) - 1) * half;
        near = find_element(sc, cfg.tmask, sm->epoch, near_loc, el->dir);
    }
    if (near >= 0 && sc.pool[near].scored != 0) {
        Elem *ne = &sc.pool[near];
        if (!((ne->base > el->base && loc - ne->best_loc <= BUCKET) || (ne->base < el->base && ne->best_loc <= BUCKET))) near = -1;
        if (near >= 0) {
            if (ne->best_score < score || (ne->best_score == score && ne->best_prob >= prob)) return false;
            any_nearby = true;
            double t = sm->p_all - ne->best_prob;
            sm->p_all = t > 0.0 ? t : 0.0;
            ne->best_prob = 0;
        }
    }
    {
        double t = sm->p_all - el->best_prob;
        sm->p_all = t > 0.0 ? t : 0.0;
    }
    sm->p_all += prob;
    el->best_prob = prob;
    el->best_score = score;
    if (sm->best_score > score || (sm->best_score == score && prob > sm->p_best)) {
        if ((sm->second_best == UNUSED_SCORE || !(sm->second_loc + BUCKET > loc && sm->second_loc < loc + BUCKET)) &&
            (sm->best_score == UNUSED_SCORE || !(sm->best_loc + BUCKET > loc && sm->best_loc < loc + BUCKET)) &&
            (!any_nearby || (sm->best_loc / BUCKET != loc / BUCKET && sm->second_loc / BUCKET != loc / BUCKET))) {
            sm->second_best = sm->best_score;
            sm->second_loc = sm->best_loc;
        }
        sm->best_score = score;
        sm->p_best = prob;
        sm->best_loc = loc;
        sm->out_loc = loc;
        sm->out_score = (int)score;
        sm->out_dir = el->dir;
    } else if (sm->second_best > score) {
        sm->second_best = score;
        sm->second_loc = loc;
    }
    if (cfg.stop_first && sm->best_score <= cfg.max_k) {
        sm->out_status = SNAPB200_MULTIPLE_HITS;
        sm->out_mapq = 0;
        return true;
    }
    sm->score_limit = min(sm->best_score, cfg.max_k) + cfg.extra;
    return false;
}

struct MapqFixList { MapqFix *items; uint32_t *count; uint32_t cap; };

// BaseAligner::score (BaseAligner.cpp:977-1399).  All lanes; returns true when a final answer was produced.
__device__ __noinline__ bool single_score(int ix_slot, const SingleCfg &cfg, const SingleScratch &sc, SingleSm *sm,
                             const ReadView v, uint8_t *W, int16_t *L, bool force_in, uint32_t read_index,
                             const MapqFixList &fix, int mapq_divisor)
{
    const int lane = lane_id();
    if (lane == 0) {
        #pragma unroll 1
        for (int d = 0; d < 2; d++) {
            uint32_t q = sm->n_applied[d] / sm->most_seeds;
            if (q > sm->lowest_unseen[d]) sm->lowest_unseen[d] = q;
        }
        sm->list = (int)sm->highest_list;
        sm->force = force_in;
    }
    #pragma unroll 1
    for (;;) {
        if (lane == 0) {
            int list = sm->list;
            #pragma unroll 1
            while (list > 0 && sc.list_head[list] < 0) { list--; sm->highest_list = (uint32_t)list; }
            sm->list = list;
            int action = ACT_ELEMENT;
            uint32_t lo = min(sm->lowest_unseen[0], sm->lowest_unseen[1]);
            if (lo > sm->score_limit || sm->force) {
                if (list == 0) {
                    sm->out_score = (int)sm->best_score;
                    if (sm->best_score <= cfg.max_k) {
                        sm->out_loc = sm->best_loc;
                        bool near_int;
                        int mq = compute_mapq_dev(sm->p_all, sm->p_best, (int)sm->best_score, (int)sm->popular, &near_int);
                        if (near_int) {
                            uint32_t slot = atomicAdd(fix.count, 1u);
                            if (slot < fix.cap) {
                                MapqFix f;
                                f.index = read_index; f.end = 0; f.p_all = sm->p_all; f.p_best = sm->p_best;
                                f.score = (int)sm->best_score; f.popular = (int)sm->popular; f.divisor = mapq_divisor;
                                f.is_paired_rule = 0;
                                fix.items[slot] = f;
                            }
                        }
                        sm->out_mapq = mq;
                        sm->out_status = mq >= 10 ? SNAPB200_SINGLE_HIT : SNAPB200_MULTIPLE_HITS;
                    } else {
                        sm->out_status = (sm->n_applied[0] == 0 && sm->n_applied[1] == 0) ? SNAPB200_MULTIPLE_HITS : SNAPB200_NOT_FOUND;
                        sm->out_mapq = 0;
                    }
                    action = ACT_RETURN_TRUE;
                }
                sm->force = 1;
            } else if (list == 0) {
                action = ACT_RETURN_FALSE;
            }
            if (action == ACT_ELEMENT) {
                int ei = sc.list_head[list];
                Elem *el = &sc.pool[ei];
                sm->cand_elem = ei;
                sm->cand_mask = (el->lowest_possible <= sm->score_limit) ? el->used : 0ull;  // snapshot (:1132)
            }
            sm->action = action;
        }
        __syncwarp();
        const int action = sm->action;
        if (action == ACT_RETURN_TRUE) return true;
        if (action == ACT_RETURN_FALSE) return false;
        // candidates of this element, ascending bit order
        #pragma unroll 1
        for (;;) {
            if (lane == 0) {
                Elem *el = &sc.pool[sm->cand_elem];
                unsigned long long mask = sm->cand_mask;
                int have = 0;
                #pragma unroll 1
                while (mask) {
                    int idx = __ffsll((long long)mask) - 1;
                    unsigned long long bit = 1ull << idx;
                    mask &= ~bit;
                    if (el->scored & bit) continue;
                    sm->cand_any_nearby = el->scored != 0;
                    el->scored |= bit;
                    sm->cand_idx = (uint32_t)idx;
                    sm->cand_loc = el->base + (uint32_t)idx;
                    sm->cand_seedoff = el->seed_offset[idx];
                    sm->cand_dir = el->dir;
                    have = 1;
                    break;
                }
                sm->cand_mask = mask;
                sm->action = have;
            }
            __syncwarp();
            if (!sm->action) break;
            double prob;
            int loc_off;
            int s = score_location_warp(ix_slot, v, sm->cand_dir, sm->cand_loc, sm->cand_seedoff, (int)sm->score_limit, true, W, L,
                                        &prob, &loc_off);
            __syncwarp();
            if (lane == 0) {
                sm->sc = s;
                sm->prob = prob;
                sm->loc_off = loc_off;
                sm->action = after_score(cfg, sc, sm) ? 1 : 0;
            }
            __syncwarp();
            if (sm->action) return true;
        }
        if (lane == 0) {  // remove the element from its weight list (:1391-1394)
            sc.pool[sm->cand_elem].all_scored = 1;
            list_unlink(sc, sm->cand_elem);
        }
        __syncwarp();
        if (!sm->force) return false;
    }
}

// leader: produce the next <= 32 seed offsets in the reference's order (BaseAligner.cpp:686-744, 876)
__device__ __forceinline__ void schedule_seeds_single(SingleSm *sm, const uint8_t *read, uint32_t len, uint32_t seed_len)
{
    const uint32_t n_possible = len - seed_len + 1;
    uint32_t next = sm->next, wrap = sm->wrap, n = 0;
    sm->terminal = 0;
    #pragma unroll 1
    while (n < 32) {
        if (next >= n_possible) {
            wrap++;
            if (wrap >= seed_len) { sm->terminal = 1; break; }
            next = wrapped_seed(seed_len, wrap);
        }
        #pragma unroll 1
        while (next < n_possible && (sm->used[next >> 5] >> (next & 31) & 1)) next++;
        if (next >= n_possible) continue;
        sm->used[next >> 5] |= 1u << (next & 31);
        bool ok = true;
        #pragma unroll 1
        for (uint32_t i = 0; i < seed_len; i++) ok &= base2(read[next + i]) >= 0;
        if (!ok) continue;  // seeds with N are skipped without counting (:742-744)
        sm->sched_off[n] = (uint16_t)next;
        sm->sched_wrap[n] = (uint16_t)wrap;
        n++;
        next += seed_len;
    }
    sm->next = next;
    sm->wrap = wrap;
    sm->n_sched = n;
}

// leader: fillHitsFound (BaseAligner.cpp:940-975)
__device__ __forceinline__ void fill_hits(const SingleCfg &cfg, const SingleScratch &sc, int32_t *found, uint32_t *locs,
                                          uint8_t *rcs, int32_t *scores)
{
    uint32_t want = cfg.max_hits_to_get;
    if (want == 0) return;
    int nf = 0;
    int first = 0;
    #pragma unroll 1
    while (first < MAXK && sc.hit_count[first] == 0) first++;
    int last = min(first + 4, MAXK);
    #pragma unroll 1
    for (int dist = first; dist < last; dist++) {
        #pragma unroll 1
        for (uint32_t i = 0; i < sc.hit_count[dist]; i++) {
            uint32_t flat = (uint32_t)dist * 512 + i;
            locs[nf] = flat < MAXK * 512 ? sc.hit_loc[flat] : 0;
            rcs[nf] = flat < MAXK * 512 ? sc.hit_rc[flat] : 0;
            scores[nf] = dist;
            nf++;
            if ((uint32_t)nf == want) { *found = nf; return; }
        }
    }
    *found = nf;
}

// BaseAligner::AlignRead (searchRadius == 0).  All lanes.  Result fields are left in *sm (out_*, p_all, ...).
// Returns false if the scratch tier overflowed (the read must be rerun with a larger tier).
__device__ bool single_align_warp(int ix_slot, const SingleCfg &cfg, const SingleScratch &sc, SingleSm *sm,
                                  ReadView &v, uint8_t *W, int16_t *L, const uint8_t *bases, const uint8_t *quals,
                                  uint32_t len, uint32_t read_index, const MapqFixList &fix, int mapq_divisor,
                                  int32_t *mh_found, uint32_t *mh_locs, uint8_t *mh_rcs, int32_t *mh_scores,
                                  unsigned long long *stat_ns_ignored)
{
    const int lane = lane_id();
    const DevIndex &ix = c_index[ix_slot];
    const uint32_t seed_len = ix.seed_len;
    if (lane == 0) {
        sm->out_loc = INVALID_LOC; sm->out_dir = SNAPB200_FORWARD; sm->out_score = UNUSED_SCORE; sm->out_mapq = 0;
        sm->out_status = SNAPB200_NOT_FOUND;
        sm->p_all = sm->p_best = 0; sm->popular = 0; sm->n_lookups = sm->n_scored = 0; sm->overflow = 0;
        sm->n_probes = sm->n_hit_words = 0;
        if (cfg.max_hits_to_get > 0) {
            #pragma unroll 1
            for (int i = 0; i < MAXK; i++) sc.hit_count[i] = 0;
            *mh_found = 0;
        }
    }
    __syncwarp();
    if (len < seed_len) return true;
    v.len = len;
    uint32_t ns = stage_read(v, bases, quals);
    if (ns > cfg.max_k) {
        if (lane == 0) atomicAdd(stat_ns_ignored, 1ull);
        return true;
    }
    uint32_t max_seeds = cfg.num_seeds ? cfg.num_seeds : (uint32_t)(int)(cfg.seed_coverage * len / seed_len);
    if (lane == 0) {  // clearCandidates + per-read state (BaseAligner.cpp:663-684)
        sm->epoch = ++(*sc.epoch);
        sm->n_used = 0;
        sm->highest_list = 0;
        #pragma unroll 1
        for (uint32_t i = 0; i < cfg.n_lists; i++) sc.list_head[i] = sc.list_tail[i] = -1;
        #pragma unroll 1
        for (int i = 0; i < 16; i++) sm->used[i] = 0;
        sm->next = 0; sm->wrap = 0;
        sm->lowest_unseen[0] = sm->lowest_unseen[1] = 0;
        sm->most_seeds = 1;
        sm->best_score = sm->second_best = UNUSED_SCORE;
        sm->best_loc = sm->second_loc = 0;
        sm->n_applied[0] = sm->n_applied[1] = 0;
        sm->score_limit = cfg.max_k + cfg.extra;
    }
    __syncwarp();
    bool answered = false, skip_fill = false;
    #pragma unroll 1
    for (;;) {
        if (sm->n_applied[0] + sm->n_applied[1] >= max_seeds) break;
        if (lane == 0) schedule_seeds_single(sm, v.D(0), len, seed_len);
        __syncwarp();
        const uint32_t n_sched = sm->n_sched;
        // warp section: all scheduled seeds are probed at once, one per lane
        HitList my[2] = {{nullptr, 0}, {nullptr, 0}};
        uint32_t my_probes = 0;
        if ((uint32_t)lane < n_sched) {
            uint64_t sf, sr;
            pack_seed(v.D(0) + sm->sched_off[lane], seed_len, &sf, &sr);
            lookup_seed(ix, sf, sr, my, &my_probes);
        }
        bool out = false;
        #pragma unroll 1
        for (uint32_t j = 0; j < n_sched; j++) {
            if (sm->n_applied[0] + sm->n_applied[1] >= max_seeds) { out = true; break; }
            const uint32_t seed_at = sm->sched_off[j];
            const uint32_t probes_j = __shfl_sync(FULL_MASK, my_probes, (int)j);
            if (lane == 0) { sm->most_seeds = (uint32_t)sm->sched_wrap[j] + 1; sm->n_lookups++; sm->n_probes += probes_j; }
            bool applied = false;
            #pragma unroll 1
            for (int dir = 0; dir < 2; dir++) {
                const uint32_t n = __shfl_sync(FULL_MASK, my[dir].n, (int)j);
                const uint32_t *hits = (const uint32_t *)shfl_u64((uint64_t)my[dir].hits, (int)j);
                if (n > cfg.max_hits && !cfg.explore) {
                    if (lane == 0) sm->popular++;
                    continue;
                }
                const uint32_t offset = dir == 0 ? seed_at : len - seed_len - seed_at;
                const uint32_t lim = min(n, cfg.max_hits);
                #pragma unroll 1
                for (uint32_t base = 0; base < lim; base += 32) {
                    uint32_t i = base + lane;
                    int e = -2;  // -2: not a usable hit (BaseAligner.cpp:848-853)
                    uint32_t loc = 0;
                    if (i < lim) {
                        uint32_t hit = __ldg(&hits[i]);
                        if (hit >= offset) {
                            loc = hit - offset;
                            e = find_element(sc, cfg.tmask, sm->epoch, loc, dir);
                        }
                    }
                    sm->hit_locs[lane] = loc;
                    sm->hit_elem[lane] = e;
                    __syncwarp();
                    if (lane == 0) {
                        uint32_t cnt = min(32u, lim - base);
                        sm->alloc_in_chunk = 0;
                        #pragma unroll 1
                        for (uint32_t q = 0; q < cnt && !sm->overflow; q++) {
                            if (sm->hit_elem[q] == -2) continue;
                            vote_hit(cfg, sc, sm, sm->hit_locs[q], sm->hit_elem[q], dir, offset);
                        }
                    }
                    __syncwarp();
                    if (sm->overflow) return false;
                }
                if (lane == 0) { sm->n_applied[dir]++; sm->n_hit_words += lim + (n > 1 ? 1 : 0); }
                applied = true;
            }
            __syncwarp();
            if (applied && single_score(ix_slot, cfg, sc, sm, v, W, L, false, read_index, fix, mapq_divisor)) { answered = true; break; }
        }
        if (answered || out) break;
        if (sm->terminal) {
            // every seed offset has been tried (wrapCount >= seedLen, BaseAligner.cpp:697-719): force a result;
            // this exit of the reference does not call fillHitsFound
            if (sm->n_applied[0] + sm->n_applied[1] < max_seeds) {
                // by now the reference has wrapped seedLen-1 times, each time setting
                // mostSeedsContainingAnyParticularBase = wrapCount+1 (:722), whether or not a lookup followed
                if (lane == 0) sm->most_seeds = seed_len;
                __syncwarp();
                single_score(ix_slot, cfg, sc, sm, v, W, L, true, read_index, fix, mapq_divisor);
                answered = true;
                skip_fill = true;
            }
            break;
        }
    }
    if (!answered) single_score(ix_slot, cfg, sc, sm, v, W, L, true, read_index, fix, mapq_divisor);
    if (lane == 0 && !skip_fill) fill_hits(cfg, sc, mh_found, mh_locs, mh_rcs, mh_scores);
    __syncwarp();
    return true;
}
